"""Shared body of the sample-tail checks (CPU emulator and GPU): phk_sample_tail against plain torch arithmetic."""
import torch

from phenaki_pytorch_b200 import _lib as L
from tests import cases as C


def make_inputs(b, n, k, dim, V, seed, dev):
    g = torch.Generator().manual_seed(seed)
    xc = torch.randn((b * n, dim), generator=g) * 2 + 0.3
    xn = torch.randn((b * n, dim), generator=g) * 2 - 0.1
    gamma, beta = torch.randn((dim,), generator=g), torch.randn((dim,), generator=g) * 0.1
    W = (torch.randn((V, dim), generator=g) / dim ** 0.5).bfloat16()
    bias = torch.randn((V,), generator=g)
    mask = torch.zeros((b, n), dtype=torch.uint8)
    for i in range(b):
        mask[i, torch.randperm(n, generator=g)[:k]] = 1
    ids0 = torch.randint(0, V, (b, n), generator=g)
    return [t.to(dev) for t in (xc, xn, gamma, beta, W, bias, mask, ids0)]


def reference(xc, xn, gamma, beta, W, bias, scale):
    """e = norm(xn) + s (norm(xc) - norm(xn)) in bf16 -> logits in fp32 -> argmax, 1 - softmax[argmax]."""
    ln = lambda x: torch.nn.functional.layer_norm(x, x.shape[-1:], gamma, beta)
    e = (scale * ln(xc) + (1 - scale) * ln(xn)).bfloat16().float()
    logits = e @ W.float().t() + bias
    pred = logits.argmax(-1)
    score = 1 - logits.softmax(-1).gather(1, pred[:, None]).squeeze(1)
    return pred, score, logits


def run_tail(lib, xc, xn, gamma, beta, W, bias, mask, ids, *, b, n, k, V, dim, scale, temperature, seed=5, offset=11):
    dev = xc.device
    pred = torch.empty((b, n), dtype=torch.int64, device=dev)
    score = torch.empty((b, n), dtype=torch.float32, device=dev)
    nb = lib.phk_sample_tail_scratch_bytes(b, k, dim)
    scratch = torch.empty(int(nb), dtype=torch.uint8, device=dev)
    L.check(lib.phk_sample_tail(L.ptr(xc), L.ptr(xn), L.ptr(gamma), L.ptr(beta), scale, L.ptr(W), dim, L.ptr(bias), b, n, k,
                                V, dim, temperature, seed, offset, None, L.ptr(mask), L.ptr(ids), L.ptr(pred), L.ptr(score),
                                L.ptr(scratch), nb, L.stream_ptr()), "phk_sample_tail")
    return pred, score


def check_exact_k(lib, dev, b, n, k, dim, V, sync=lambda: None):
    """Exactly k masked tokens per sequence, temperature 0 (pure argmax, :493): masked rows get the argmax of the guided
    logits and 1 - p; the others keep their id and score -1e4."""
    xc, xn, gamma, beta, W, bias, mask, ids0 = make_inputs(b, n, k, dim, V, 100 + k, dev)
    ids = ids0.clone()
    pred, score = run_tail(lib, xc, xn, gamma, beta, W, bias, mask, ids, b=b, n=n, k=k, V=V, dim=dim, scale=3.0,
                           temperature=0.0)
    sync()
    rp, rs, logits = reference(xc.cpu(), xn.cpu(), gamma.cpu(), beta.cpu(), W.cpu(), bias.cpu(), 3.0)
    m = mask.cpu().reshape(-1).bool()
    ids, pred, score, ids0 = ids.cpu().reshape(-1), pred.cpu().reshape(-1), score.cpu().reshape(-1), ids0.cpu().reshape(-1)
    # a differing argmax is only acceptable between logits that tie to within bf16-operand rounding of the embedding
    differ = (ids[m] != rp[m]).nonzero().flatten()
    rows = m.nonzero().flatten()
    for j in differ.tolist():
        r = rows[j]
        assert abs(float(logits[r, ids[r]] - logits[r, rp[r]])) < 2e-2, f"row {int(r)}: wrong argmax"
    assert len(differ) <= max(1, int(0.02 * m.sum()))
    same = m.clone()
    same[rows[differ]] = False
    torch.testing.assert_close(score[same], rs[same], rtol=2e-3, atol=2e-4)
    assert torch.equal(pred[m], ids[m])
    assert torch.equal(ids[~m], ids0[~m]) and torch.equal(pred[~m], ids0[~m])
    assert bool((score[~m] == -1e4).all())
