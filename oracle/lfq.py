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


class VectorQuantize(nn.Module):
    """Placeholder so ``from vector_quantize_pytorch import VectorQuantize`` resolves.
    The cosine-sim codebook path (cvivit.py:321) is a next-tier row."""

    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("cosine-sim VectorQuantize is out of scope (SURVEY 8f-3)")
