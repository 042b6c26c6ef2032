"""GPU parity of the C-ViViT decode row: `decode_from_codebook_indices` / `decode` / the video that
`Phenaki.sample` returns, against the reconstruction the UNMODIFIED reference produced (tests/golden) and the oracle.

Bars: fp32 parity mode |err| <= 2e-4 + 2e-4*|ref| (same as the encoder activations); bf16 mode
|err| <= 0.06 + 0.03*|ref| on the decoder activations and pixels (pixels are a Linear of a LayerNorm output, O(1))."""
import pytest
import torch

from oracle import phenaki_oracle as O
import phenaki_pytorch_b200 as P
from phenaki_pytorch_b200 import _lib as L
from tests import cases as C

pytestmark = pytest.mark.gpu
DEV = "cuda"
ATOL, RTOL = 2e-4, 2e-4


def pair(v):
    return v if isinstance(v, tuple) else (v, v)


def close(a, b, atol, rtol, what):
    err = (a - b).abs()
    worst = float((err - (atol + rtol * b.abs())).max())
    assert worst <= 0, f"{what}: max |err| {float(err.max()):.5f} exceeds {atol} + {rtol}*|ref| by {worst:.5f}"


def test_lfq_codes_kernel_matches_oracle():
    torch.manual_seed(0)
    for rows, dim, bits in [(48, 256, 16), (577, 512, 16), (36, 128, 10), (5, 64, 8)]:
        ids = torch.randint(0, 2 ** bits, (rows,), dtype=torch.int64)
        sd = {"vq.mask": 2 ** torch.arange(bits - 1, -1, -1), "vq.project_out.weight": torch.randn(dim, bits),
              "vq.project_out.bias": torch.randn(dim)}
        ref = O.lfq_indices_to_codes(ids[None], sd)[0]
        out = torch.empty(rows, dim, device=DEV)
        d = {k: v.to(DEV) for k, v in sd.items()}
        idd = ids.to(DEV)   # named: device copies must outlive the launch
        L.check(L.lib().phk_lfq_codes(L.ptr(idd), L.ptr(d["vq.project_out.weight"]), L.ptr(d["vq.project_out.bias"]),
                                      L.ptr(out), rows, dim, bits, L.stream_ptr()))
        torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,C_,F,H,W,f0,nt,pt,p1,p2", [(2, 3, 7, 32, 48, 1, 2, 3, 8, 16), (2, 3, 7, 32, 48, 0, 1, 1, 8, 16),
                                                      (3, 1, 1, 32, 32, 0, 1, 1, 8, 8), (1, 3, 5, 12, 10, 1, 2, 2, 3, 5)])
def test_unpatchify_is_the_inverse_rearrange(B, C_, F, H, W, f0, nt, pt, p1, p2):
    hh, ww = H // p1, W // p2
    K = C_ * pt * p1 * p2
    g = torch.Generator().manual_seed(1)
    Pm = torch.randn(B * nt * hh * ww, K, generator=g)
    video = torch.full((B, C_, F, H, W), -7.0, device=DEV)
    Pd = Pm.to(DEV)
    L.check(L.lib().phk_unpatchify(L.ptr(Pd), K, L.ptr(video), B, C_, F, H, W, f0, nt, pt, p1, p2, L.stream_ptr()))
    # 'b t h w (c pt p1 p2) -> b c (t pt) (h p1) (w p2)'  (cvivit.py:286-295)
    ref = Pm.reshape(B, nt, hh, ww, C_, pt, p1, p2).permute(0, 4, 1, 5, 2, 6, 3, 7).reshape(B, C_, nt * pt, H, W)
    assert torch.equal(video[:, :, f0:f0 + nt * pt].cpu(), ref)          # pure data movement: bit-exact
    rest = torch.ones(F, dtype=torch.bool)
    rest[f0:f0 + nt * pt] = False
    assert (video[:, :, rest.to(DEV)] == -7.0).all()                      # other frames untouched


def test_layernorm_gather_mode_picks_input_rows():
    torch.manual_seed(2)
    B, T, hw, D = 3, 4, 6, 128
    x = torch.randn(B * T * hw, D)
    g, b = torch.randn(D), torch.randn(D)
    ref = torch.nn.functional.layer_norm(x, (D,), g, b).reshape(B, T, hw, D)
    xd, gd, bd = x.to(DEV), g.to(DEV), b.to(DEV)
    first = torch.empty(B * hw, D, device=DEV)
    rest = torch.empty(B * (T - 1) * hw, D, device=DEV)
    lib = L.lib()
    L.check(lib.phk_layernorm(L.ptr(xd), L.ptr(gd), L.ptr(bd), L.ptr(first), None, B * hw, D, 0, -hw, T * hw, 0, L.stream_ptr()))
    L.check(lib.phk_layernorm(L.ptr(xd), L.ptr(gd), L.ptr(bd), L.ptr(rest), None, B * (T - 1) * hw, D, 0, -(T - 1) * hw,
                              T * hw, hw, L.stream_ptr()))
    torch.testing.assert_close(first.cpu().reshape(B, 1, hw, D), ref[:, :1], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(rest.cpu().reshape(B, T - 1, hw, D), ref[:, 1:], rtol=1e-5, atol=1e-5)


def _model(name, prec):
    case = C.CVIVIT_CASES[name]
    torch.manual_seed(case["seed"])
    model = P.CViViT(**case["ctor"]).to(DEV).eval()
    model.precision = prec
    return case, model


@pytest.mark.parametrize("name", [n for n in C.CVIVIT_CASES if n != "cosine_vq"])  # cosine_vq: test_gpu_zz_*
def test_decode_from_codebook_indices_matches_reference_golden(golden, name):
    case, model = _model(name, L.PREC_F32)
    g = golden(f"cvivit_{name}")
    ids = g["ids"].to(DEV)
    taps = {}
    rec = model.decode_from_codebook_indices(ids.reshape(ids.shape[0], -1), taps=taps)
    assert rec.dtype == torch.float32 and tuple(rec.shape) == tuple(g["recon"].shape)
    # stage by stage against the oracle so that a failure names the kernel
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    b, t, h, w = g["ids"].shape
    with torch.no_grad():
        if "vq._codebook.embed" in sd:  # cosine-sim codebook: codes = vq.codebook[indices] (cvivit.py:441)
            codes = sd["vq._codebook.embed"][0][g["ids"].reshape(b, -1)].reshape(b, t, h, w, -1)
        else:
            codes = O.lfq_indices_to_codes(g["ids"].reshape(b, -1), sd).reshape(b, t, h, w, -1)
    torch.testing.assert_close(taps["codes"].cpu(), codes, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(rec.cpu(), g["recon"], rtol=RTOL, atol=ATOL)
    # 4-D ids and the float-token entry point give the same video
    assert torch.equal(model.decode_from_codebook_indices(ids), rec)
    torch.testing.assert_close(model.decode(taps["codes"]), rec, rtol=0, atol=0)
    # return_recons_only = encode + decode (cvivit.py:576-579)
    video = C.seeded_randn(case["video"], case["video_seed"]).to(DEV)
    rr = model(video, return_recons_only=True)
    assert rr.ndim == video.ndim                                          # images come back as (b, c, h, w)
    torch.testing.assert_close(rr.cpu().reshape(g["recon"].shape), g["recon"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("name", ["cfg1", "rect"])
def test_decode_bf16_mode_against_fp32_reference_golden(golden, name):
    case, model = _model(name, L.PREC_BF16)
    g = golden(f"cvivit_{name}")
    rec = model.decode_from_codebook_indices(g["ids"].to(DEV))
    close(rec.cpu(), g["recon"], 0.06, 0.03, "reconstruction")
    again = model.decode_from_codebook_indices(g["ids"].to(DEV))
    assert torch.equal(rec, again)


def test_decode_token_count_contract():
    case, model = _model("rect", L.PREC_F32)
    with pytest.raises(AssertionError):
        model.decode_from_codebook_indices(torch.zeros((1, 13), dtype=torch.int64, device=DEV))  # not k * (h*w)
    with pytest.raises(L.PhkError):
        model.decode_from_codebook_indices(torch.zeros((1, 12), dtype=torch.int64))              # CPU tensor: no fallback


def _phenaki(case):
    torch.manual_seed(case["seed"])
    cv = P.CViViT(**C.SAMPLE_CVIVIT)
    mg = P.MaskGit(**C.SAMPLE_MASKGIT)
    cr = P.TokenCritic(**C.SAMPLE_CRITIC) if case["critic"] else None
    return P.Phenaki(cvivit=cv.to(DEV), maskgit=mg.to(DEV), critic=cr.to(DEV) if cr else None, steps=case["steps"],
                     text_embed_dim=C.SAMPLE_MASKGIT["dim_context"])


@pytest.mark.parametrize("name", list(C.SAMPLE_CASES))
def test_sampled_video_matches_reference_golden(golden, name):
    """Phenaki.sample end to end (token loop with the reference's uniform draws replayed, then C-ViViT decode and
    the prime-frame crop, phenaki_pytorch.py:552-560): pixels against the reference's video."""
    case, g = C.SAMPLE_CASES[name], golden(f"sample_{name}")
    ph = _phenaki(case)
    ctx = C.synthetic_text_embeds(case["batch"], case["ctx_len"], C.SAMPLE_MASKGIT["dim_context"], case["ctx_valid"],
                                  case["seed"] + 1000).to(DEV)
    tape = C.NoiseTape(case["noise_seed"])
    prime = None
    if case["prime"]:
        prime = C.seeded_randn((case["batch"], 3, case["prime_frames"], *C.SAMPLE_CVIVIT["image_size"]),
                               case["seed"] + 2000).to(DEV)
    video = ph.sample(num_frames=case["num_frames"], text_embeds=ctx, prime_frames=prime, cond_scale=case["cond_scale"],
                      noise_fn=lambda shape, tag: tape(shape, tag).to(DEV))
    assert tuple(video.shape) == tuple(g["video"].shape)
    torch.testing.assert_close(video.cpu(), g["video"], rtol=RTOL, atol=ATOL)


def test_make_video_scene_chain_matches_reference_golden(golden):
    """make_video end to end (BASELINE configs[4] in miniature): 3 scenes, TokenCritic, each scene primed with the
    last 4 decoded frames of the previous one (re-encoded by C-ViViT), uniform draws replayed, pixels vs the reference."""
    case, g = C.MAKE_VIDEO_CASE, golden("make_video")
    torch.manual_seed(case["seed"])
    cv, mg, cr = P.CViViT(**C.SAMPLE_CVIVIT), P.MaskGit(**C.SAMPLE_MASKGIT), P.TokenCritic(**C.SAMPLE_CRITIC)
    assert C.state_digest(cv.state_dict()) == g["cvivit_digest"] and C.state_digest(cr.state_dict()) == g["critic_digest"]
    ph = P.Phenaki(cvivit=cv.to(DEV), maskgit=mg.to(DEV), critic=cr.to(DEV), steps=case["steps"],
                   text_embed_dim=C.SAMPLE_MASKGIT["dim_context"])
    table = C.make_video_text_table(case)
    ph.encode_texts = lambda texts, output_device=None: table[texts[0]].to(DEV)    # stands in for T5 (no weights here)
    tape = C.NoiseTape(case["noise_seed"])
    sample = ph.sample
    ph.sample = lambda **kw: sample(noise_fn=lambda shape, tag: tape(shape, tag).to(DEV), **kw)
    video, scenes = P.make_video(ph, texts=list(case["texts"]), num_frames=case["num_frames"],
                                 prime_lengths=case["prime_lengths"])
    assert tuple(video.shape) == tuple(g["video"].shape) and len(scenes) == 3
    torch.testing.assert_close(video.cpu(), g["video"], rtol=RTOL, atol=ATOL)
    for a, b in zip(scenes, g["scenes"]):
        torch.testing.assert_close(a.cpu(), b, rtol=RTOL, atol=ATOL)
