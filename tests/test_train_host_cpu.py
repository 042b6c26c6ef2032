"""CPU: host-side logic of the training step -- the random-subset mask, the gradient table / flat gradient bucket, and
the autograd bridge of Phenaki.forward -- with the C call replaced by the CPU restatement of the kernels
(tests/train_mirror.py), against the reference's gradients (tests/golden/train_*.pt)."""
import pytest
import torch

import phenaki_pytorch_b200 as P
from oracle import phenaki_oracle as O
from phenaki_pytorch_b200 import modules as M
from phenaki_pytorch_b200.phenaki import get_mask_subset_with_prob
from tests import cases as C
from tests import train_mirror as TM
from tests.test_oracle_golden import _train_modules


def test_mask_subset_matches_the_reference_formula():
    g = torch.Generator().manual_seed(7)
    for b, n in ((3, 18), (2, 48), (4, 7)):
        mask = torch.ones((b, n), dtype=torch.bool)
        mask[0, n - 3:] = False  # padding only shifts the ranks (phenaki_pytorch.py:49-51)
        prob = torch.rand((b,), generator=g)
        u = torch.rand((b, n), generator=g)
        assert torch.equal(get_mask_subset_with_prob(mask, prob, u), O.mask_subset_with_prob(mask, prob, u))


@pytest.mark.parametrize("name", list(C.TRAIN_CASES))
def test_gradient_table_covers_exactly_the_parameters_the_reference_differentiates(golden, name):
    case, g = C.TRAIN_CASES[name], golden(f"train_{name}")
    maskgit, critic = _train_modules(case)
    for module, ref in ((maskgit, g["maskgit_grads"]), (critic, g["critic_grads"])):
        if module is None:
            continue
        table, gk = module._grad_table(with_cross=True)
        named = dict(module.named_parameters())
        used = {k for k, p in named.items() if gk.grad_of(p) is not None}
        assert used == set(ref), (used ^ set(ref))
        assert gk.flat.numel() >= sum(p.numel() for p in named.values()) and gk.grad_of(w_any := next(iter(named.values()))).data_ptr() % 256 == gk.flat.data_ptr() % 256
        assert table.transformer.depth == module.transformer.depth
        # every gradient view aliases the one flat bucket (the unit of the data-parallel all-reduce)
        w = module.transformer.layers[0][0].dsconv.weight
        gk.flat.fill_(3.0)
        assert gk.grad_of(w).shape == w.shape and bool((gk.grad_of(w) == 3.0).all())


def _fake_train_step(module, heads):
    """Stands in for phk_maskgit_train_step: same contract as _TokenTransformer.train_step, computed by the CPU
    restatement of the kernels."""

    def run(ids_in, patch_shape, *, targets=None, token_mask=None, labels=None, context=None, text_mask=None,
            video_mask=None, loss_scale=1.0, keep_logits=False, overlap_all_reduce=False):
        sd = {k: v.detach() for k, v in module.state_dict().items()}
        with torch.no_grad():
            loss, grads, logits = TM.train_step(sd, ids_in, targets, token_mask, labels, patch_shape=patch_shape,
                                                heads=heads, context=context, text_mask=text_mask,
                                                is_critic=module.is_critic, loss_scale=loss_scale, video_mask=video_mask)
        _, gk = module._grad_table(with_cross=context is not None)
        for k, p in module.named_parameters():
            if gk.grad_of(p) is not None:
                gk.views[p].copy_(grads[k])
        return loss, gk, (logits.reshape(*ids_in.shape, -1) if keep_logits and logits is not None else None)

    return run


@pytest.mark.parametrize("name", ["generator", "with_critic"])
def test_phenaki_forward_autograd_bridge_with_the_kernels_restated_on_cpu(golden, name, monkeypatch):
    case, g = C.TRAIN_CASES[name], golden(f"train_{name}")
    torch.manual_seed(case["seed"])
    cvivit = P.CViViT(**C.SAMPLE_CVIVIT)
    maskgit = P.MaskGit(**case["maskgit"])
    critic = P.TokenCritic(**case["critic"]) if case["critic"] else None
    phenaki = P.Phenaki(cvivit=cvivit, maskgit=maskgit, critic=critic, steps=case["steps"],
                        text_embed_dim=case["maskgit"]["dim_context"]).train()
    heads = case["maskgit"].get("heads", 8)
    monkeypatch.setattr(phenaki.maskgit, "train_step", _fake_train_step(phenaki.maskgit, heads))
    ids, ctx = C.train_inputs(case)
    b, n = ids.shape[0], ids[0].numel()
    torch.manual_seed(case["noise_seed"])
    rand_step, u = O.train_draws(b, n, case["steps"])
    draws = {"rand_step": rand_step, "perm": u}
    loss = phenaki(video_codebook_ids=ids, text_embeds=ctx, only_train_generator=True,
                   draw_fn=lambda shape, tag: draws[tag])
    (2.0 * loss).backward()  # the upstream gradient scales what libphk computed
    ref_loss = g["ce"] if critic is not None else g["loss"]
    torch.testing.assert_close(loss.detach(), ref_loss, rtol=1e-5, atol=1e-6)
    for k, p in phenaki.maskgit.named_parameters():
        if k in g["maskgit_grads"]:
            torch.testing.assert_close(p.grad, 2.0 * g["maskgit_grads"][k], rtol=2e-4, atol=4e-6, msg=lambda m, k=k: f"{k}: {m}")
        else:
            assert p.grad is None


def test_training_entry_refuses_cpu_tensors():
    """No CPU fallback: the training entry raises on CPU inputs like every other entry point."""
    from phenaki_pytorch_b200 import _lib as L
    torch.manual_seed(0)
    phenaki = P.Phenaki(cvivit=P.CViViT(**C.SAMPLE_CVIVIT), maskgit=P.MaskGit(**C.SAMPLE_MASKGIT), text_embed_dim=48)
    with pytest.raises((L.PhkError, RuntimeError, AssertionError)):
        phenaki(video_codebook_ids=torch.zeros((1, 3, 2, 3), dtype=torch.int64), text_embeds=torch.zeros((1, 4, 48)))
