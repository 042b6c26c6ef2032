"""Runs the CUDA training step (Phenaki.forward -> phk_maskgit_train_step) on cuda:0 against the reference's
autograd loss and gradients stored in tests/golden/train_*.pt and prints one line per case.

Stand-alone on purpose (python tests/gpu_train_check.py [case ...]): tests/test_gpu_train.py runs it in a child
process, so a fault in the not-yet-validated kernels cannot take the rest of the GPU suite down with it.
"""
import os
import sys

os.environ["PHK_EXPERIMENTAL"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import phenaki_pytorch_b200 as P  # noqa: E402
from oracle import phenaki_oracle as O  # noqa: E402  (the checker)
from tests import cases as C  # noqa: E402


def check_case(name, verbose=True, bf16=False, device="cuda:0"):
    """bf16=True: tcgen05 products (PHK_PREC_BF16); the bar is then closeness to the fp32 reference (5 % of each
    gradient tensor's largest entry), not parity."""
    case = C.TRAIN_CASES[name]
    g = torch.load(os.path.join(ROOT, "tests", "golden", f"train_{name}.pt"), weights_only=False)
    torch.manual_seed(case["seed"])
    cvivit = P.CViViT(**C.SAMPLE_CVIVIT)
    maskgit = P.MaskGit(**case["maskgit"])
    critic = P.TokenCritic(**case["critic"]) if case["critic"] else None
    assert C.state_digest(maskgit.state_dict()) == g["maskgit_digest"]
    dev = torch.device(device)
    phenaki = P.Phenaki(cvivit=cvivit, maskgit=maskgit, critic=critic, steps=case["steps"],
                        self_token_critic=case.get("self_critic", False),
                        text_embed_dim=case["maskgit"]["dim_context"]).to(dev).train()
    if bf16:
        from phenaki_pytorch_b200 import _lib as L
        phenaki.maskgit.precision = L.PREC_BF16
        if critic is not None:
            critic.precision = L.PREC_BF16
    ids, ctx = C.train_inputs(case)
    b, n = ids.shape[0], ids[0].numel()
    vocab = case["maskgit"]["num_tokens"]
    # the reference's draws, in its order, from the CPU generator it used (tests/golden/make_golden.py::make_train)
    torch.manual_seed(case["noise_seed"])
    rand_step, u = O.train_draws(b, n, case["steps"])
    draws = {"rand_step": rand_step, "perm": u}
    if phenaki.critic is not None:
        draws["gumbel"] = torch.zeros((b, n, vocab)).uniform_(0, 1)
    loss = phenaki(video_codebook_ids=ids.to(dev), text_embeds=ctx.to(dev), draw_fn=lambda shape, tag: draws[tag])
    loss.backward()
    if dev.type == "cuda":
        torch.cuda.synchronize()
    worst = 0.0
    if bf16 and phenaki.critic is not None:
        # the sampled predictions (an argmax over noisy bf16-product logits) may differ from the fp32 run, and with them
        # the critic's inputs: only the generator half is comparable
        print("  (bf16: critic half not compared -- its inputs are sampled from the logits)")
    torch.testing.assert_close(loss.detach().cpu(), g["loss"], rtol=5e-2 if bf16 else 1e-4, atol=1e-5)

    def compare(module, ref_grads, who):
        nonlocal worst
        named = dict(module.named_parameters())
        for k, p in named.items():
            if k not in ref_grads:
                assert p.grad is None, f"{who}.{k}: the reference leaves this gradient unset"
                continue
            assert p.grad is not None, f"{who}.{k}: no gradient"
            ref = ref_grads[k]
            got = p.grad.detach().cpu()
            if ref.numel() == 0:
                assert got.shape == ref.shape
                continue
            scale = max(ref.abs().max().item(), 1e-12)
            err = (got - ref).abs().max().item() / scale
            worst = max(worst, err)
            if verbose:
                print(f"  {who}.{k:60s} max|ref| {scale:.3e}  max err / max|ref| {err:.2e}")
            if bf16:
                assert scale < 1e-7 or err < 5e-2, f"{who}.{k}: max err / max|ref| = {err:.3e}"
            else:
                torch.testing.assert_close(got, ref, rtol=2e-3, atol=2e-4 * scale + 1e-7, msg=lambda m, k=k: f"{who}.{k}: {m}")

    if not (bf16 and case.get("self_critic")):
        compare(phenaki.maskgit, g["maskgit_grads"], "maskgit")
    if critic is not None and not bf16:
        compare(phenaki.critic, g["critic_grads"], "critic")
    if case.get("self_critic") and not bf16:
        compare(phenaki.critic.to_pred[0], g["to_pred_grads"], "to_pred")
    print(f"TRAIN_OK {name}{' bf16' if bf16 else ''} loss {loss.item():.6f} worst relative gradient error {worst:.2e}")


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "--bf16"]
    for nm in args or list(C.TRAIN_CASES):
        check_case(nm, bf16="--bf16" in sys.argv)
