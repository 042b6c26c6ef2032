"""CPU: the hand-derived backward formulas that csrc/train.cu implements (restated in tests/train_mirror.py, no
autograd) reproduce the reference's autograd loss and gradients (tests/golden/train_*.pt, written by the UNMODIFIED
reference).  A wrong formula in the CUDA kernels would therefore already show up here, without a GPU."""
import pytest
import torch

from oracle import phenaki_oracle as O
from tests import cases as C
from tests import train_mirror as M
from tests.test_oracle_golden import _train_modules


@pytest.mark.parametrize("name", list(C.TRAIN_CASES))
def test_backward_formulas_match_reference_autograd(golden, name):
    case, g = C.TRAIN_CASES[name], golden(f"train_{name}")
    maskgit, critic = _train_modules(case)
    mg_sd = {k: v.detach() for k, v in maskgit.state_dict().items()}
    ids, ctx = C.train_inputs(case)
    b, n = ids.shape[0], ids[0].numel()
    flat = ids.reshape(b, n)
    token_mask = g["token_mask"]
    mask_id = case["maskgit"]["num_tokens"]
    heads = case["maskgit"].get("heads", 8)
    tmask = torch.any(ctx != 0, dim=-1)
    vmask = torch.ones((b, n), dtype=torch.bool)
    with torch.no_grad():
        loss, grads, logits = M.train_step(mg_sd, torch.where(token_mask, mask_id, flat), flat, token_mask, None,
                                           patch_shape=case["patch_shape"], heads=heads, context=ctx, text_mask=tmask,
                                           is_critic=False, video_mask=vmask)
        total = loss
        if critic is not None:
            cr_sd = {k: v.detach() for k, v in critic.state_dict().items()}
            pred = g["pred_ids"]
            closs, cgrads, _ = M.train_step(cr_sd, torch.where(token_mask, pred, flat), None, None,
                                            (flat != pred).float(), patch_shape=case["patch_shape"], heads=heads,
                                            context=ctx, text_mask=tmask, is_critic=True, video_mask=vmask)
            torch.testing.assert_close(loss, g["ce"], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(closs, g["bce"], rtol=1e-5, atol=1e-6)
            total = loss + closs
            for k, ref in g["critic_grads"].items():
                torch.testing.assert_close(cgrads[k], ref, rtol=2e-4, atol=2e-6, msg=lambda m, k=k: f"critic.{k}: {m}")
        if case.get("self_critic"):
            pred = g["pred_ids"]
            closs, sgrads, _ = M.train_step(mg_sd, torch.where(token_mask, pred, flat), None, None, (flat != pred).float(),
                                            patch_shape=case["patch_shape"], heads=heads, context=ctx, text_mask=tmask,
                                            is_critic=False, video_mask=vmask,
                                            head=(g["to_pred_weight"], g["to_pred_bias"]))
            torch.testing.assert_close(closs, g["bce"], rtol=1e-5, atol=1e-6)
            total = loss + closs
            torch.testing.assert_close(sgrads["to_pred.weight"], g["to_pred_grads"]["weight"], rtol=2e-4, atol=2e-6)
            torch.testing.assert_close(sgrads["to_pred.bias"], g["to_pred_grads"]["bias"], rtol=2e-4, atol=2e-6)
            grads = {k: v + sgrads[k] for k, v in grads.items()}  # both losses differentiate MaskGit
    torch.testing.assert_close(total, g["loss"], rtol=1e-5, atol=1e-6)
    for k, ref in g["maskgit_grads"].items():
        torch.testing.assert_close(grads[k], ref, rtol=2e-4, atol=2e-6, msg=lambda m, k=k: f"maskgit.{k}: {m}")
