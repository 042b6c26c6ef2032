"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

CPU restatement of the eval-mode arithmetic of ``vector_quantize_pytorch.LFQ``
(lookup-free quantisation, arXiv 2310.05737).

The reference does NOT vendor this dependency: it is pinned only as
``vector-quantize-pytorch>=1.11.8`` (/root/reference/setup.py:34) and is absent
from this container (no network), so its published algorithm is restated here
and parity is anchored on the reference's own call sites:

  * constructor     /root/reference/phenaki_pytorch/cvivit.py:318-319
                    ``LFQ(dim = dim, codebook_size = codebook_size, **kwargs)``
  * forward         /root/reference/phenaki_pytorch/cvivit.py:568-570
                    ``tokens, indices, vq_aux_loss = self.vq(tokens)``
  * decode lookup   /root/reference/phenaki_pytorch/cvivit.py:437-439
                    ``codes = self.vq.indices_to_codes(indices)``

PARITY UNPINNED for this one step: the reference holds no test / golden vector
for LFQ, and upstream cannot be imported here, so this file *defines* the VQ
step of the oracle.  State-dict names follow upstream
(``mask``, ``project_in.*``, ``project_out.*``).

Published algorithm (eval mode):
    codebook_dim = log2(codebook_size)
    x   = project_in(tokens)                      # Linear(dim, codebook_dim) + bias
    q   = where(x > 0, +scale, -scale)
    idx = sum_d (x_d > 0) * 2**(codebook_dim-1-d) # MSB first
    out = project_out(q)                          # Linear(codebook_dim, dim) + bias
    aux = 0
"""
import math

import torch
from torch import nn


class LFQ(nn.Module):
    def __init__(self, *, dim, codebook_size, codebook_scale=1.0, **unused_training_kwargs):
        super().__init__()
        codebook_dim = int(math.log2(codebook_size))
        assert 2 ** codebook_dim == codebook_size, "codebook size must be a power of two"
        self.dim = dim
        self.codebook_dim = codebook_dim
        self.codebook_size = codebook_size
        self.codebook_scale = codebook_scale
        has_projections = dim != codebook_dim
        self.project_in = nn.Linear(dim, codebook_dim) if has_projections else nn.Identity()
        self.project_out = nn.Linear(codebook_dim, dim) if has_projections else nn.Identity()
        self.register_buffer("mask", 2 ** torch.arange(codebook_dim - 1, -1, -1))

    def indices_to_codes(self, indices, project_out=True):
        bits = (indices[..., None].int() & self.mask) != 0
        codes = torch.where(bits, self.codebook_scale, -self.codebook_scale).float()
        return self.project_out(codes) if project_out else codes

    def forward(self, x, **unused):
        x = self.project_in(x)
        positive = x > 0
        quantized = torch.where(positive, self.codebook_scale, -self.codebook_scale).to(x.dtype)
        indices = (positive.int() * self.mask.int()).sum(dim=-1)
        out = self.project_out(quantized)
        return out, indices, torch.zeros((), device=x.device)


class _CosineSimCodebook(nn.Module):
    """Buffers of upstream ``CosineSimCodebook`` for one codebook without k-means init (restated from memory of
    vector-quantize-pytorch 1.11-1.14, unverifiable here): ``embed`` = l2-normalised kaiming-uniform rows."""

    def __init__(self, dim, codebook_size):
        super().__init__()
        embed = torch.empty(1, codebook_size, dim)
        nn.init.kaiming_uniform_(embed)
        embed = torch.nn.functional.normalize(embed, dim=-1)
        self.register_buffer("initted", torch.Tensor([True]))
        self.register_buffer("cluster_size", torch.zeros(1, codebook_size))
        self.register_buffer("embed_avg", embed.clone())
        self.register_buffer("embed", embed)


class VectorQuantize(nn.Module):
    """Eval-mode arithmetic of ``vector_quantize_pytorch.VectorQuantize(dim, codebook_size, use_cosine_sim=True)``
    (the reference's non-LFQ tokenizer, cvivit.py:321; call sites :441 ``vq.codebook[indices]`` and :570
    ``vq(tokens, mask=...)``).  PARITY UNPINNED like LFQ above: the dependency is absent, so this restatement -- from
    memory of upstream 1.11-1.14 -- DEFINES the step:
        x_n   = l2norm(x)
        idx   = argmax_c  x_n . embed_c            (embed rows are unit vectors)
        out   = embed[idx]                          (project_in / project_out are Identity when codebook_dim == dim)
        loss  = 0 (the commitment loss, EMA update and straight-through estimator are training-only)
    State-dict names follow upstream: ``_codebook.{initted, cluster_size, embed_avg, embed}``."""

    def __init__(self, *, dim, codebook_size, use_cosine_sim=True, **unused_training_kwargs):
        super().__init__()
        assert use_cosine_sim, "the reference only builds the cosine-sim codebook (cvivit.py:321)"
        self.dim, self.codebook_size = dim, codebook_size
        self._codebook = _CosineSimCodebook(dim, codebook_size)

    @property
    def codebook(self):
        return self._codebook.embed[0]

    def forward(self, x, mask=None, **unused):
        flat = torch.nn.functional.normalize(x.float(), dim=-1)
        dist = flat @ self.codebook.t()
        indices = dist.argmax(dim=-1)
        return self.codebook[indices], indices, torch.zeros((1,), device=x.device)
