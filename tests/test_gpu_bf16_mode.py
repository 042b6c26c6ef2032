"""GPU: PHK_PREC_BF16 (tcgen05 GEMMs, bf16 operands, fp32 accumulation / residual / LayerNorm / softmax)
against the fp32 reference goldens.  This is the dtype flow of the reference under
torch.autocast(bfloat16), so the comparison is reference-fp32 vs bf16 contraction noise.

Stated tolerances (normalised activations are O(1); bf16 has 8 mantissa bits, error grows ~sqrt(layers)):
  * activations after patch-embed / each transformer:  |err| <= 0.06 + 0.03*|ref|
  * MaskGit logits / embeds:                           |err| <= 0.08 + 0.03*|ref|
  * LFQ token ids: a bit may flip only where the reference pre-sign margin is below 0.12
    (the bf16 noise floor on the 16 projections); bits with larger margins must agree, and
    overall bit agreement must exceed 97 %.
"""
import pytest
import torch

import phenaki_pytorch_b200 as P
from phenaki_pytorch_b200 import _lib as L
from tests import cases as C

pytestmark = pytest.mark.gpu
DEV = "cuda"


def close(a, b, atol, rtol, what):
    err = (a - b).abs()
    bound = atol + rtol * b.abs()
    worst = float((err - bound).max())
    assert worst <= 0, f"{what}: max |err| {float(err.max()):.4f} exceeds {atol} + {rtol}*|ref| by {worst:.4f}"
    return float(err.max())


@pytest.mark.parametrize("name", ["cfg1", "rect"])
def test_cvivit_bf16_mode_against_fp32_reference_golden(golden, name):
    case, g = C.CVIVIT_CASES[name], golden(f"cvivit_{name}")
    torch.manual_seed(case["seed"])
    model = P.CViViT(**case["ctor"]).to(DEV).eval()
    model.precision = L.PREC_BF16
    video = C.seeded_randn(case["video"], case["video_seed"]).to(DEV)
    taps = {}
    ids = model.encode_ids(video, taps=taps)
    b, t, h, w, d = g["patch"].shape
    close(taps["patch"].cpu(), g["patch"], 0.06, 0.03, "patch embed")
    close(taps["spatial"].cpu().reshape(b * t, h * w, d), g["spatial"], 0.06, 0.03, "spatial transformer")
    close(taps["temporal"].cpu().permute(0, 2, 3, 1, 4).reshape(b * h * w, t, d), g["temporal"], 0.06, 0.03,
          "temporal transformer")
    bits = model.vq.codebook_dim
    proj = g["proj"].reshape(-1, bits)
    diff = (ids.cpu().reshape(-1) ^ g["ids"].reshape(-1))
    flipped = 0
    for r in range(diff.numel()):
        for dbit in range(bits):
            if (int(diff[r]) >> (bits - 1 - dbit)) & 1:
                flipped += 1
                assert abs(float(proj[r, dbit])) < 0.12, f"bit flipped with reference margin {float(proj[r, dbit]):.3f}"
    assert flipped <= 0.03 * diff.numel() * bits, f"{flipped} of {diff.numel() * bits} LFQ bits differ"


def test_maskgit_bf16_mode_against_fp32_reference_golden(golden):
    case, g = C.MASKGIT_CASES["wide"], golden("maskgit_wide")
    torch.manual_seed(case["seed"])
    model = P.MaskGit(**case["ctor"]).to(DEV).eval()
    model.precision = L.PREC_BF16
    ids, ctx = C.token_inputs(case, case["ctor"]["num_tokens"])
    ids, ctx = ids.to(DEV), ctx.to(DEV)
    kw = dict(text_mask=torch.any(ctx != 0, dim=-1), video_patch_shape=case["patch_shape"], context=ctx)
    close(model(ids, return_embeds=True, **kw).cpu(), g["embeds"], 0.08, 0.03, "embeds")
    close(model(ids, **kw).cpu(), g["cond"], 0.08, 0.03, "logits (cond)")
    close(model(ids, cond_drop_prob=1.0, **kw).cpu(), g["null"], 0.08, 0.03, "logits (null)")


def test_sampling_runs_in_bf16_mode_and_is_deterministic():
    case = C.SAMPLE_CASES["critic_primed"]
    torch.manual_seed(case["seed"])
    cv, mg, cr = P.CViViT(**C.SAMPLE_CVIVIT), P.MaskGit(**C.SAMPLE_MASKGIT), P.TokenCritic(**C.SAMPLE_CRITIC)
    for m in (cv, mg, cr):
        m.precision = L.PREC_BF16
    ph = P.Phenaki(cvivit=cv.to(DEV), maskgit=mg.to(DEV), critic=cr.to(DEV), steps=4,
                   text_embed_dim=C.SAMPLE_MASKGIT["dim_context"])
    ph.cvivit.precision = L.PREC_BF16
    ctx = C.synthetic_text_embeds(2, 6, C.SAMPLE_MASKGIT["dim_context"], (6, 3), 3).to(DEV)
    outs = []
    for _ in range(2):
        tape = C.NoiseTape(8)
        outs.append(ph.sample(num_frames=7, text_embeds=ctx, return_token_ids=True,
                              noise_fn=lambda s, t: tape(s, t).to(DEV)).cpu())
    assert torch.equal(outs[0], outs[1])
    assert (outs[0] < C.SAMPLE_MASKGIT["num_tokens"]).all()
