"""GPU parity, module granularity, through the reference-facing classes (which call the C ABI):
C-ViViT token ids / MaskGit logits / TokenCritic scores / the full demasking loop against the golden
vectors produced by the UNMODIFIED reference (tests/golden) and against the oracle.

Bars (fp32 parity mode, PHK_PREC_F32):
  * token ids, masks: identical (a differing LFQ bit is tolerated only where the REFERENCE's own
    pre-sign value is below 2e-5, i.e. inside fp32 summation-order noise -- reported, none in the goldens);
  * activations / logits: |err| <= 2e-4 + 2e-4*|ref| after up to 8 transformer layers.
"""
import pytest
import torch

from oracle import phenaki_oracle as O
import phenaki_pytorch_b200 as P
from tests import cases as C

pytestmark = pytest.mark.gpu
DEV = "cuda"
ATOL, RTOL = 2e-4, 2e-4
MARGIN = 2e-5


def pair(v):
    return v if isinstance(v, tuple) else (v, v)


def assert_ids_match(ids, ref_ids, ref_proj, bits):
    ids, ref_ids = ids.cpu().reshape(-1), ref_ids.reshape(-1)
    diff = ids ^ ref_ids
    bad = torch.nonzero(diff).flatten().tolist()
    proj = ref_proj.reshape(-1, bits)
    for r in bad:
        for d in range(bits):
            if (int(diff[r]) >> (bits - 1 - d)) & 1:
                assert abs(float(proj[r, d])) < MARGIN, f"token {r} bit {d} flipped with margin {float(proj[r, d])}"
    return len(bad)


@pytest.mark.parametrize("name", [n for n in C.CVIVIT_CASES if n != "cosine_vq"])  # cosine_vq: test_gpu_zz_*
def test_cvivit_token_ids_match_reference_golden(golden, name):
    case, g = C.CVIVIT_CASES[name], golden(f"cvivit_{name}")
    torch.manual_seed(case["seed"])
    model = P.CViViT(**case["ctor"])
    assert C.state_digest(model.state_dict()) == g["state_digest"]
    model = model.to(DEV).eval()
    video = C.seeded_randn(case["video"], case["video_seed"]).to(DEV)
    taps = {}
    v5 = video if video.ndim == 5 else video.unsqueeze(2)
    ids = model.encode_ids(v5, taps=taps)
    assert ids.dtype == torch.int64 and tuple(ids.shape) == tuple(g["ids"].shape)
    lfq = case["ctor"].get("lookup_free_quantization", True)
    bits = model.vq.codebook_dim if lfq else 0
    # stage by stage, so a failure names the kernel
    torch.testing.assert_close(taps["patch"].cpu(), g["patch"], rtol=RTOL, atol=ATOL)
    b, t, h, w, d = g["patch"].shape
    torch.testing.assert_close(taps["spatial"].cpu().reshape(b * t, h * w, d), g["spatial"], rtol=RTOL, atol=ATOL)
    temporal = taps["temporal"].cpu().permute(0, 2, 3, 1, 4).reshape(b * h * w, t, d)  # -> '(b h w) t d'
    torch.testing.assert_close(temporal, g["temporal"], rtol=RTOL, atol=ATOL)
    if not lfq:
        # cosine-sim codebook (cvivit.py:321): an id may only differ where the reference's two best similarities tie to
        # within fp32 summation noise (none in the golden case)
        got, want = ids.cpu().reshape(-1), g["ids"].reshape(-1)
        sims = g["proj"].reshape(-1, g["proj"].shape[-1])
        for r in torch.nonzero(got != want).flatten().tolist():
            assert abs(float(sims[r, got[r]] - sims[r, want[r]])) < MARGIN, f"token {r}: wrong nearest code"
        assert torch.equal(got, want), "the golden case has margins far above fp32 noise: ids must be identical"
        return
    torch.testing.assert_close(taps["proj"].cpu().reshape(g["proj"].shape), g["proj"], rtol=RTOL, atol=ATOL)
    flipped = assert_ids_match(ids, g["ids"], g["proj"], bits)
    assert flipped == 0, "golden cases have margins far above fp32 noise: ids must be identical"
    # the module-level API returns the same thing
    out = model(video, return_only_codebook_ids=True)
    assert torch.equal(out, ids)


def test_cvivit_state_dict_roundtrip_changes_nothing(golden):
    case, g = C.CVIVIT_CASES["rect"], golden("cvivit_rect")
    torch.manual_seed(case["seed"])
    src = P.CViViT(**case["ctor"])
    torch.manual_seed(999)
    dst = P.CViViT(**case["ctor"]).to(DEV)
    video = C.seeded_randn(case["video"], case["video_seed"]).to(DEV)
    before = dst(video, return_only_codebook_ids=True)
    dst.load_state_dict(src.state_dict(), strict=True)   # in-place parameter update must invalidate the tables
    after = dst(video, return_only_codebook_ids=True)
    assert torch.equal(after.cpu(), g["ids"]) and not torch.equal(before, after)


def test_cvivit_shape_contract_errors():
    torch.manual_seed(0)
    m = P.CViViT(**C.SAMPLE_CVIVIT).to(DEV)
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 3, 5, 16, 24, device=DEV), return_only_codebook_ids=True)  # (5-1) % 3 != 0
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 3, 4, 16, 16, device=DEV), return_only_codebook_ids=True)  # wrong image size


@pytest.mark.parametrize("name", list(C.MASKGIT_CASES))
def test_maskgit_logits_match_reference_golden(golden, name):
    case, g = C.MASKGIT_CASES[name], golden(f"maskgit_{name}")
    torch.manual_seed(case["seed"])
    model = P.MaskGit(**case["ctor"])
    assert C.state_digest(model.state_dict()) == g["state_digest"]
    model = model.to(DEV).eval()
    ids, ctx = C.token_inputs(case, case["ctor"]["num_tokens"])
    ids, ctx = ids.to(DEV), ctx.to(DEV)
    tmask = torch.any(ctx != 0, dim=-1)
    kw = dict(text_mask=tmask, video_patch_shape=case["patch_shape"], context=ctx)
    bias = model._pos_bias(model._table(), case["patch_shape"], ids.device)
    torch.testing.assert_close(bias.cpu(), g["bias"], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(model(ids, return_embeds=True, **kw).cpu(), g["embeds"], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(model(ids, cond_drop_prob=0.0, **kw).cpu(), g["cond"], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(model(ids, cond_drop_prob=1.0, **kw).cpu(), g["null"], rtol=RTOL, atol=ATOL)
    # CFG: null + (cond - null) * 3 amplifies the error by <= 5x
    cfg = model.forward_with_cond_scale(ids, cond_scale=3.0, **kw)
    torch.testing.assert_close(cfg.cpu(), g["cfg"], rtol=RTOL, atol=5 * ATOL)
    # 4-D ids (b, t, h, w) carry their own patch shape (phenaki_pytorch.py:175-177)
    out4 = model(ids.reshape(ids.shape[0], *case["patch_shape"]), text_mask=tmask, context=ctx)
    torch.testing.assert_close(out4.cpu(), g["cond"], rtol=RTOL, atol=ATOL)


def test_maskgit_sequence_length_contract():
    case = C.MASKGIT_CASES["small"]
    torch.manual_seed(1)
    m = P.MaskGit(**{**case["ctor"], "max_seq_len": 16}).to(DEV)
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 48, dtype=torch.long, device=DEV), video_patch_shape=(3, 4, 4))


def test_token_critic_scores_match_reference_golden(golden):
    case, g = C.CRITIC_CASES["small"], golden("critic_small")
    torch.manual_seed(case["seed"])
    model = P.TokenCritic(**case["ctor"])
    assert C.state_digest(model.state_dict()) == g["state_digest"]
    model = model.to(DEV).eval()
    ids, ctx = C.token_inputs(case, case["ctor"]["num_tokens"])
    ids, ctx = ids.to(DEV), ctx.to(DEV)
    tmask = torch.any(ctx != 0, dim=-1)
    kw = dict(text_mask=tmask, video_patch_shape=case["patch_shape"], context=ctx)
    torch.testing.assert_close(model(ids, cond_drop_prob=0.0, **kw).cpu(), g["cond"], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(model.forward_with_cond_scale(ids, cond_scale=5.0, **kw).cpu(), g["cfg"], rtol=RTOL,
                               atol=9 * ATOL)


def _phenaki(case):
    torch.manual_seed(case["seed"])
    cv = P.CViViT(**C.SAMPLE_CVIVIT)
    mg = P.MaskGit(**C.SAMPLE_MASKGIT)
    cr = P.TokenCritic(**C.SAMPLE_CRITIC) if case["critic"] else None
    ph = P.Phenaki(cvivit=cv.to(DEV), maskgit=mg.to(DEV), critic=cr.to(DEV) if cr else None, steps=case["steps"],
                   text_embed_dim=C.SAMPLE_MASKGIT["dim_context"])
    return ph


@pytest.mark.parametrize("name", list(C.SAMPLE_CASES))
def test_sampling_loop_token_ids_match_reference_golden(golden, name):
    """Phenaki.sample's demasking loop, free running, with the reference's uniform draws replayed:
    every step's mask / prediction / ids and the final ids must equal the reference's (integers)."""
    case, g = C.SAMPLE_CASES[name], golden(f"sample_{name}")
    ph = _phenaki(case)
    ctx = C.synthetic_text_embeds(case["batch"], case["ctx_len"], C.SAMPLE_MASKGIT["dim_context"], case["ctx_valid"],
                                  case["seed"] + 1000).to(DEV)
    tape = C.NoiseTape(case["noise_seed"])
    noise_fn = lambda shape, tag: tape(shape, tag).to(DEV)
    prime = None
    if case["prime"]:
        prime = C.seeded_randn((case["batch"], 3, case["prime_frames"], *C.SAMPLE_CVIVIT["image_size"]),
                               case["seed"] + 2000).to(DEV)
        pids = ph.cvivit(prime, return_only_codebook_ids=True)
        assert torch.equal(pids.reshape(case["batch"], -1).cpu(), g["prime_ids"])
    trace = []
    orig = ph.sample_token_ids
    ph.sample_token_ids = lambda **kw: orig(trace=trace, **kw)
    ids = ph.sample(num_frames=case["num_frames"], text_embeds=ctx, prime_frames=prime, cond_scale=case["cond_scale"],
                    return_token_ids=True, noise_fn=noise_fn)
    for mine, ref in zip(trace, g["trace"]):
        s = ref["step"]
        assert torch.equal(mine["mask"].cpu(), ref["mask"]), f"step {s}: mask differs"
        assert torch.equal(mine["pred"].cpu(), ref["pred"]), f"step {s}: sampled ids differ"
        assert torch.equal(mine["ids"].cpu(), ref["ids"]), f"step {s}: ids differ"
        if "scores" in ref:
            torch.testing.assert_close(mine["scores"].cpu(), ref["scores"], rtol=1e-3, atol=1e-3)
    assert torch.equal(ids.cpu(), g["final_ids"])


def test_sampling_is_seed_deterministic_and_fills_every_token():
    case = C.SAMPLE_CASES["confidence"]
    ph = _phenaki(case)
    ctx = C.synthetic_text_embeds(2, 6, C.SAMPLE_MASKGIT["dim_context"], (6, 3), 7).to(DEV)
    outs = []
    for seed in (5, 5, 6):
        torch.manual_seed(seed)
        outs.append(ph.sample(num_frames=7, text_embeds=ctx, return_token_ids=True).cpu())
    assert torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], outs[2])
    assert (outs[0] >= 0).all() and (outs[0] < C.SAMPLE_MASKGIT["num_tokens"]).all()  # no mask id left


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_encode_host_stream_matches_device_encode(depth):
    """CViViT.encode_host_iter (phk_encode_pipe_*): pinned host batches in, host ids out, copies overlapped with the
    previous batch's encode -- every batch must carry exactly the ids of the plain device call, in order."""
    case = C.CVIVIT_CASES["rect"]
    torch.manual_seed(case["seed"])
    model = P.CViViT(**case["ctor"]).to(DEV).eval()
    batches = [C.seeded_randn(case["video"], 100 + i).pin_memory() for i in range(7)]
    want = [model(v.to(DEV), return_only_codebook_ids=True).cpu() for v in batches]
    got = list(model.encode_host_iter(iter(batches), depth=depth))
    assert len(got) == len(want)
    for i, (a, b) in enumerate(zip(got, want)):
        assert not a.is_cuda and a.dtype == torch.int64 and torch.equal(a, b), f"batch {i}"
    assert torch.equal(model.encode_host(batches[3]), want[3])
    with pytest.raises(P._lib.PhkError):
        next(model.encode_host_iter([batches[0].to(DEV)]))            # device tensors go through forward()
    with pytest.raises(AssertionError):
        list(model.encode_host_iter([batches[0], batches[1][:1]]))    # one shape per stream


def test_encode_graph_replay_equals_eager_launches():
    """The library replays a captured CUDA graph from the third call with identical buffers on (the first runs eagerly,
    the second captures): ids stay identical, the launch counter advances by the same amount per call, and new input
    DATA in the same buffer is honoured."""
    case = C.CVIVIT_CASES["cfg1"]
    torch.manual_seed(case["seed"])
    model = P.CViViT(**case["ctor"]).to(DEV).eval()
    lib = P._lib.lib()
    video = C.seeded_randn(case["video"], case["video_seed"]).to(DEV)
    outs, counts = [], []
    for _ in range(5):
        before = lib.phk_launch_count()
        outs.append(model(video, return_only_codebook_ids=True))
        counts.append(lib.phk_launch_count() - before)
    assert all(torch.equal(o, outs[0]) for o in outs)
    assert len(set(counts[1:])) == 1 and counts[1] > 10, counts   # the first call also builds the cached position bias
    other = C.seeded_randn(case["video"], 777).to(DEV)
    expect = model(other, return_only_codebook_ids=True)       # eager (new buffer)
    video.copy_(other)                                          # same buffer as the captured graph, new contents
    assert torch.equal(model(video, return_only_codebook_ids=True), expect)
    assert not torch.equal(expect, outs[0])
