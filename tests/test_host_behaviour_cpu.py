"""CPU: host-side behaviour a user of the reference relies on (no kernels involved): copy.deepcopy of the modules (EMA
wrappers), out-of-range token ids raise like nn.Embedding, the weight signature notices replaced Parameters, the
default-precision switch."""
import copy
import importlib

import pytest
import torch

import phenaki_pytorch_b200 as P
from phenaki_pytorch_b200 import _lib as L
from phenaki_pytorch_b200.modules import weights_signature
from tests import cases as C


def test_deepcopy_gives_an_independent_module_with_the_same_weights():
    torch.manual_seed(0)
    mg = P.MaskGit(**C.SAMPLE_MASKGIT)
    mg.precision = L.PREC_BF16
    cp = copy.deepcopy(mg)
    assert cp is not mg and cp.precision == L.PREC_BF16 and cp._tables is None
    for (k, a), (_, b) in zip(mg.state_dict().items(), cp.state_dict().items()):
        assert torch.equal(a, b) and (a.numel() == 0 or a.data_ptr() != b.data_ptr()), k
    ph = P.Phenaki(cvivit=P.CViViT(**C.SAMPLE_CVIVIT), maskgit=mg, critic=P.TokenCritic(**C.SAMPLE_CRITIC), text_embed_dim=48)
    ph2 = copy.deepcopy(ph)
    assert ph2.maskgit is not ph.maskgit and torch.equal(ph2.critic.pos_emb.weight, ph.critic.pos_emb.weight)


def test_out_of_range_token_ids_raise_like_nn_embedding():
    torch.manual_seed(0)
    mg = P.MaskGit(**C.SAMPLE_MASKGIT)
    bad = torch.full((1, 18), C.SAMPLE_MASKGIT["num_tokens"] + 1, dtype=torch.int64)  # one past the mask id
    with pytest.raises(IndexError):
        mg(bad, video_patch_shape=(3, 2, 3))
    with pytest.raises(IndexError):
        P.TokenCritic(**C.SAMPLE_CRITIC)(torch.full((1, 3, 2, 3), -1, dtype=torch.int64), cond_drop_prob=0.0)


def test_weight_signature_sees_replaced_and_modified_parameters():
    torch.manual_seed(0)
    mg = P.MaskGit(**C.SAMPLE_MASKGIT)
    s0 = weights_signature(mg)
    assert weights_signature(mg) == s0
    with torch.no_grad():
        mg.to_logits.bias.add_(1.0)            # in-place update bumps _version
    s1 = weights_signature(mg)
    assert s1 != s0
    mg.to_logits.weight = torch.nn.Parameter(mg.to_logits.weight.detach().clone())  # replaced Parameter object: seen at once
    assert weights_signature(mg) != s1
    mg._sig = ("stale",)
    P.invalidate_weights(mg)
    assert mg._sig is None


def test_default_precision_follows_the_environment(monkeypatch):
    monkeypatch.setenv("PHK_PREC", "bf16x3")
    assert P.CViViT(**C.SAMPLE_CVIVIT).precision == L.PREC_BF16X3
    monkeypatch.setenv("PHK_PREC", "bf16")
    assert P.MaskGit(**C.SAMPLE_MASKGIT).precision == L.PREC_BF16
    monkeypatch.delenv("PHK_PREC")
    assert P.MaskGit(**C.SAMPLE_MASKGIT).precision == L.default_precision()


def test_gradient_groups_cover_the_bucket_once_in_completion_order():
    """The slices the overlapped all-reduce launches as the backward finishes them: every element of the flat gradient
    bucket in exactly one span, head first, layers top-down, embeddings + position-bias MLP last -- for the MaskGit of
    BASELINE configs[3] and for a SelfCritic owner (MaskGit parameters followed by to_pred).  (A zero-element parameter --
    the self-attention's null_kv (heads, 0, dim_head) -- has no address in the bucket and must not drag a group's span to offset "minus base".)"""
    import phenaki_pytorch_b200 as P
    from phenaki_pytorch_b200.modules import GradKeep
    mg = P.MaskGit(dim=64, num_tokens=128, max_seq_len=64, heads=2, dim_head=32, depth=3, dim_context=48)
    assert any(p.numel() == 0 for p in mg.parameters())
    for owner in (mg, P.SelfCritic(mg)):
        gk = GradKeep(owner.parameters())
        groups = mg._gradient_groups(gk, owner)
        assert groups is not None and len(groups) == mg.transformer.depth + 2
        covered = torch.zeros(gk.flat.numel(), dtype=torch.int32)
        for spans in groups:
            for lo, hi in spans:
                covered[lo:hi] += 1
        assert int(covered.min()) == 1 and int(covered.max()) == 1
        names = dict(owner.named_parameters())
        where = lambda n: (gk.views[names[n]].data_ptr() - gk.flat.data_ptr()) // 4
        inside = lambda n, g: any(lo <= where(n) < hi for lo, hi in groups[g])
        prefix = "maskgit." if owner is not mg else ""
        assert inside(prefix + "to_logits.weight", 0) and inside(prefix + "transformer.norm_out.gamma", 0)
        assert inside(prefix + "transformer.layers.2.0.dsconv.weight", 1) and inside(prefix + "transformer.layers.0.1.q_scale", 3)
        assert inside(prefix + "token_emb.weight", 4) and inside(prefix + "continuous_pos_bias.net.0.0.weight", 4)
