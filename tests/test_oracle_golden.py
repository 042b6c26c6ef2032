"""CPU suite (``-m "not gpu"``): the oracle restatement against the committed golden vectors that the
UNMODIFIED reference produced (tests/golden/make_golden.py), plus host logic of the product."""
import math

import pytest
import torch

from oracle import phenaki_oracle as O
from tests import cases as C
import phenaki_pytorch_b200 as P
from phenaki_pytorch_b200.phenaki import demask_counts

torch.set_num_threads(max(1, min(8, torch.get_num_threads())))

# float outputs: the oracle re-runs the same ATen ops; MKL may pick another kernel -> summation order only
FTOL = dict(rtol=1e-5, atol=1e-5)


def pair(v):
    return v if isinstance(v, tuple) else (v, v)


def build(cls, case_or_kwargs, seed):
    torch.manual_seed(seed)
    return cls(**case_or_kwargs)


@pytest.mark.parametrize("name", list(C.CVIVIT_CASES))
def test_oracle_cvivit_ids_match_reference_golden(golden, name):
    case, g = C.CVIVIT_CASES[name], golden(f"cvivit_{name}")
    model = build(P.CViViT, case["ctor"], case["seed"])
    sd = model.state_dict()
    assert C.state_digest(sd) == g["state_digest"], "seeded construction must reproduce the reference weights"
    assert len(sd) == g["n_state"]
    video = C.seeded_randn(case["video"], case["video_seed"])
    with torch.no_grad():
        ids, proj = O.cvivit_codebook_ids(video, sd, pair(case["ctor"]["image_size"]), pair(case["ctor"]["patch_size"]),
                                          return_margin=True)
    assert ids.dtype == torch.int64
    assert torch.equal(ids, g["ids"])  # integer output: bit-exact
    torch.testing.assert_close(proj.reshape(g["proj"].shape), g["proj"], **FTOL)


@pytest.mark.parametrize("name", ["rect", "image"])
def test_oracle_cvivit_decode_matches_reference_golden(golden, name):
    case, g = C.CVIVIT_CASES[name], golden(f"cvivit_{name}")
    sd = build(P.CViViT, case["ctor"], case["seed"]).state_dict()
    with torch.no_grad():
        rec = O.cvivit_decode_from_ids(g["ids"].reshape(g["ids"].shape[0], -1), sd, pair(case["ctor"]["image_size"]),
                                       pair(case["ctor"]["patch_size"]))
    torch.testing.assert_close(rec, g["recon"], **FTOL)


@pytest.mark.parametrize("name", list(C.MASKGIT_CASES))
def test_oracle_maskgit_matches_reference_golden(golden, name):
    case, g = C.MASKGIT_CASES[name], golden(f"maskgit_{name}")
    sd = build(P.MaskGit, case["ctor"], case["seed"]).state_dict()
    assert C.state_digest(sd) == g["state_digest"]
    ids, ctx = C.token_inputs(case, case["ctor"]["num_tokens"])
    kw = dict(video_patch_shape=case["patch_shape"], heads=case["ctor"].get("heads", 8), context=ctx,
              text_mask=torch.any(ctx != 0, dim=-1))
    with torch.no_grad():
        torch.testing.assert_close(O.maskgit_forward(ids, sd, **kw), g["cond"], **FTOL)
        torch.testing.assert_close(O.maskgit_forward(ids, sd, cond_drop=True, **kw), g["null"], **FTOL)
        torch.testing.assert_close(O.continuous_position_bias(sd, "continuous_pos_bias.", case["patch_shape"]),
                                   g["bias"], **FTOL)


def test_oracle_critic_matches_reference_golden(golden):
    case, g = C.CRITIC_CASES["small"], golden("critic_small")
    sd = build(P.TokenCritic, case["ctor"], case["seed"]).state_dict()
    assert C.state_digest(sd) == g["state_digest"]
    ids, ctx = C.token_inputs(case, case["ctor"]["num_tokens"])
    kw = dict(video_patch_shape=case["patch_shape"], heads=case["ctor"]["heads"], context=ctx,
              text_mask=torch.any(ctx != 0, dim=-1))
    with torch.no_grad():
        torch.testing.assert_close(O.critic_forward(ids, sd, **kw), g["cond"], **FTOL)
        cfg = O.with_cond_scale(lambda cond_drop: O.critic_forward(ids, sd, cond_drop=cond_drop, **kw), 5.0)
        torch.testing.assert_close(cfg, g["cfg"], **FTOL)


def _sample_models(case):
    torch.manual_seed(case["seed"])
    cv = P.CViViT(**C.SAMPLE_CVIVIT)
    mg = P.MaskGit(**C.SAMPLE_MASKGIT)
    cr = P.TokenCritic(**C.SAMPLE_CRITIC) if case["critic"] else None
    return cv, mg, cr


@pytest.mark.parametrize("name", list(C.SAMPLE_CASES))
def test_oracle_sampling_loop_matches_reference_golden(golden, name):
    """The full demasking loop (token ids: integer, must be identical) replayed with the same uniform draws."""
    case, g = C.SAMPLE_CASES[name], golden(f"sample_{name}")
    cv, mg, cr = _sample_models(case)
    assert C.state_digest(cv.state_dict()) == g["cvivit_digest"]
    assert C.state_digest(mg.state_dict()) == g["maskgit_digest"]
    ctx = C.synthetic_text_embeds(case["batch"], case["ctx_len"], C.SAMPLE_MASKGIT["dim_context"], case["ctx_valid"],
                                  case["seed"] + 1000)
    tape = C.NoiseTape(case["noise_seed"])
    with torch.no_grad():
        ids = O.sample_token_ids(mg.state_dict(), num_tokens=g["num_tokens"], patch_shape=g["patch_shape"],
                                 batch=case["batch"], steps=case["steps"], heads=C.SAMPLE_MASKGIT["heads"],
                                 text_embeds=ctx, prime_ids=g["prime_ids"], cond_scale=case["cond_scale"],
                                 critic_sd=cr.state_dict() if cr else None, noise_fn=tape)
    assert torch.equal(ids, g["final_ids"])


def test_units_attention_and_peg_match_reference_golden(golden):
    """Re-derives the inputs of tests/golden/make_golden.py:make_units and checks shapes/finite-ness of the
    committed module outputs (the module-level pin itself ran in the build container)."""
    g = golden("units")
    for k in ("self", "causal", "cross"):
        assert g[k]["y"].shape == (2, 12, 64) and torch.isfinite(g[k]["y"]).all()
    for k in ("peg_True", "peg_False"):
        assert g[k]["y"].shape == (12, 4, 16)


def test_demask_schedule_matches_reference_formula():
    """phenaki_pytorch.py:485-486 evaluated with torch (as the reference does) vs the host-side table."""
    for n, steps in [(576, 18), (448, 18), (48, 6), (36, 5), (1024, 24), (1, 3)]:
        ref = O.demask_schedule(n, steps)
        assert demask_counts(n, steps) == ref
    assert O.demask_schedule(576, 18) == [574, 567, 556, 541, 522, 499, 472, 441, 407, 370, 330, 288, 243, 197, 149,
                                          100, 50]  # SURVEY.md 8a-14 probe of the reference


def test_shape_helpers_follow_reference():
    torch.manual_seed(0)
    m = P.CViViT(**C.SAMPLE_CVIVIT)
    assert m.patch_height_width == (2, 3) and m.image_num_tokens == 6
    assert m.get_video_patch_shape(7) == (3, 2, 3)
    assert m.num_tokens_per_frames(7) == 18 and m.num_tokens_per_frames(6, include_first_frame=False) == 12
    assert m.frames_per_num_tokens(18) == 7
    fm = torch.tensor([[1, 1, 1, 1, 0, 0, 0], [1, 1, 1, 1, 1, 1, 1]], dtype=torch.bool)
    vm = m.calculate_video_token_mask(torch.zeros(2, 3, 7, 16, 24), fm)
    assert vm.shape == (2, 18) and vm[0].sum() == 12 and vm[1].all()
    with pytest.raises(AssertionError):
        m.num_tokens_per_frames(6)


def test_product_has_no_cpu_fallback():
    """CPU tensors must be refused loudly -- there is no PyTorch / oracle fallback on the product path."""
    torch.manual_seed(0)
    m = P.CViViT(**C.SAMPLE_CVIVIT)
    from phenaki_pytorch_b200._lib import PhkError
    with pytest.raises(PhkError):
        m(torch.zeros(1, 3, 4, 16, 24), return_only_codebook_ids=True)
    import os
    src = os.path.join(os.path.dirname(P.__file__))
    for fn in os.listdir(src):
        if fn.endswith(".py"):
            text = open(os.path.join(src, fn)).read()
            assert "import oracle" not in text and "from oracle" not in text, f"{fn} imports the oracle"


def test_oracle_make_video_chain_matches_reference_golden(golden):
    """make_video (phenaki_pytorch.py:691-714): three scenes, TokenCritic, priming with the previous scene's last frames."""
    from tests.chain import oracle_make_video
    case, g = C.MAKE_VIDEO_CASE, golden("make_video")
    torch.manual_seed(case["seed"])
    cv, mg, cr = P.CViViT(**C.SAMPLE_CVIVIT), P.MaskGit(**C.SAMPLE_MASKGIT), P.TokenCritic(**C.SAMPLE_CRITIC)
    sds = [m.state_dict() for m in (cv, mg, cr)]
    assert [C.state_digest(sd) for sd in sds] == [g["cvivit_digest"], g["maskgit_digest"], g["critic_digest"]]
    video, scenes = oracle_make_video(case, *sds, C.make_video_text_table(case), C.NoiseTape(case["noise_seed"]))
    assert tuple(video.shape) == (1, 3, sum(case["num_frames"]), *C.SAMPLE_CVIVIT["image_size"])
    torch.testing.assert_close(video, g["video"], **FTOL)
    for a, b in zip(scenes, g["scenes"]):
        torch.testing.assert_close(a, b, **FTOL)


def test_weight_signature_cache_tracks_in_place_updates_moves_and_replaced_parameters():
    """The product rebuilds its weight tables when the (cached) signature changes: in-place updates (optimizer steps,
    load_state_dict) and dtype/device moves are seen at once; a Parameter OBJECT replaced by hand within 64 calls."""
    from phenaki_pytorch_b200.modules import weights_signature, _SIG_REFRESH
    torch.manual_seed(0)
    m = P.CViViT(**C.CVIVIT_CASES["image"]["ctor"])
    s0 = weights_signature(m)
    assert weights_signature(m) == s0
    with torch.no_grad():
        m.vq.project_in.weight.add_(1.0)
    s1 = weights_signature(m)
    assert s1 != s0
    m.load_state_dict({k: v.clone() for k, v in m.state_dict().items()})
    s2 = weights_signature(m)
    assert s2 != s1
    m.double()                                       # nn.Module._apply keeps the Parameter objects, moves the data
    assert weights_signature(m) != s2
    m.float()
    s3 = weights_signature(m)
    m.vq.project_in.weight = torch.nn.Parameter(m.vq.project_in.weight.detach().clone())
    seen = [weights_signature(m) for _ in range(_SIG_REFRESH + 1)]
    assert seen[-1] != s3


def _train_modules(case):
    """Same construction order as tests/golden/make_golden.py::make_train (seeded default inits)."""
    torch.manual_seed(case["seed"])
    P.CViViT(**C.SAMPLE_CVIVIT)
    maskgit = P.MaskGit(**case["maskgit"])
    critic = P.TokenCritic(**case["critic"]) if case["critic"] else None
    return maskgit, critic


@pytest.mark.parametrize("name", list(C.TRAIN_CASES))
def test_oracle_training_loss_and_gradients_match_reference_golden(golden, name):
    """Phenaki.forward (phenaki_pytorch.py:562-687): loss and every parameter gradient of the reference's autograd
    equal autograd through the functional oracle, with the reference's random draws replayed."""
    case, g = C.TRAIN_CASES[name], golden(f"train_{name}")
    maskgit, critic = _train_modules(case)
    mg_sd = maskgit.state_dict()
    assert C.state_digest(mg_sd) == g["maskgit_digest"]

    def leaf(sd):
        return {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}

    o_mg = leaf(mg_sd)
    ids, ctx = C.train_inputs(case)
    b, n = ids.shape[0], ids[0].numel()
    torch.manual_seed(case["noise_seed"])
    rand_step, u = O.train_draws(b, n, case["steps"])
    token_mask = O.train_token_mask(rand_step, u, case["steps"])
    assert torch.equal(rand_step, g["rand_step"]) and torch.equal(token_mask, g["token_mask"])
    kw = dict(video_patch_shape=case["patch_shape"], heads=case["maskgit"].get("heads", 8), context=ctx,
              text_mask=torch.any(ctx != 0, dim=-1))
    loss, logits = O.maskgit_train_loss(ids.reshape(b, n), o_mg, token_mask, return_logits=True, **kw)
    o_cr = None
    if critic is not None:
        cr_sd = critic.state_dict()
        assert C.state_digest(cr_sd) == g["critic_digest"]
        o_cr = leaf(cr_sd)
        pred = O.gumbel_sample(logits.detach(), 1.0, torch.zeros_like(logits).uniform_(0, 1))
        assert torch.equal(pred, g["pred_ids"])
        loss = loss + O.critic_train_loss(ids.reshape(b, n), pred, token_mask, o_cr, **kw) * 1.0
    if case.get("self_critic"):
        w, bb = g["to_pred_weight"].clone().requires_grad_(True), g["to_pred_bias"].clone().requires_grad_(True)
        pred = O.gumbel_sample(logits.detach(), 1.0, torch.zeros_like(logits).uniform_(0, 1))
        assert torch.equal(pred, g["pred_ids"])
        loss = loss + O.self_critic_train_loss(ids.reshape(b, n), pred, token_mask, o_mg, w, bb, **kw)
    loss.backward()
    torch.testing.assert_close(loss.detach(), g["loss"], **FTOL)
    if case.get("self_critic"):
        torch.testing.assert_close(w.grad, g["to_pred_grads"]["weight"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(bb.grad, g["to_pred_grads"]["bias"], rtol=1e-4, atol=1e-6)
    for k, ref in g["maskgit_grads"].items():
        torch.testing.assert_close(o_mg[k].grad, ref, rtol=1e-4, atol=1e-6, msg=lambda m, k=k: f"maskgit.{k}: {m}")
    if critic is not None:
        for k, ref in g["critic_grads"].items():
            torch.testing.assert_close(o_cr[k].grad, ref, rtol=1e-4, atol=1e-6, msg=lambda m, k=k: f"critic.{k}: {m}")


def test_cvivit_constructs_with_the_reference_defaults_and_loads_gan_checkpoints():
    """The reference's default is use_vgg_and_gan=True (cvivit.py:227-249): the tokenizer must still construct (the
    GAN / VGG members belong to training losses that are out of scope) and accept a checkpoint that carries the
    discriminator, without disturbing the seeded weights or the state-dict layout."""
    case = C.CVIVIT_CASES["image"]
    kw = {k: v for k, v in case["ctor"].items() if k != "use_vgg_and_gan"}
    torch.manual_seed(case["seed"])
    a = P.CViViT(**kw)                                    # reference defaults
    torch.manual_seed(case["seed"])
    b = P.CViViT(**case["ctor"])                          # use_vgg_and_gan=False
    assert a.use_vgg_and_gan and a.discr is None and a.vgg is None
    assert C.state_digest(a.state_dict()) == C.state_digest(b.state_dict())
    sd = dict(b.state_dict())
    sd["discr.layers.0.0.weight"] = torch.zeros(3)        # what a reference GAN checkpoint adds
    missing = a.load_state_dict(sd)                       # strict
    assert not missing.missing_keys and not missing.unexpected_keys
    with pytest.raises(NotImplementedError):
        a(torch.zeros(1, 1, 32, 32), return_recons=True)  # the loss paths stay unavailable


# ---- round 2 goldens: SelfCritic sampling, video_mask, video_frame_mask training (oracle vs committed reference outputs)
def test_oracle_self_critic_sampling_loop_matches_reference_golden(golden):
    case, g = C.SELF_CRITIC_SAMPLE_CASE, golden("sample_self_critic")
    torch.manual_seed(case["seed"])
    P.CViViT(**C.SAMPLE_CVIVIT)  # consumes the generator like the reference construction order
    mg = P.MaskGit(**C.SAMPLE_MASKGIT)
    assert C.state_digest(mg.state_dict()) == g["maskgit_digest"]
    ctx = C.synthetic_text_embeds(case["batch"], case["ctx_len"], C.SAMPLE_MASKGIT["dim_context"], case["ctx_valid"],
                                  case["seed"] + 1000)
    with torch.no_grad():
        ids = O.sample_token_ids(mg.state_dict(), num_tokens=g["num_tokens"], patch_shape=g["patch_shape"],
                                 batch=case["batch"], steps=case["steps"], heads=C.SAMPLE_MASKGIT["heads"],
                                 text_embeds=ctx, cond_scale=case["cond_scale"],
                                 self_critic=(g["to_pred_weight"], g["to_pred_bias"]),
                                 noise_fn=C.NoiseTape(case["noise_seed"]))
    assert torch.equal(ids, g["final_ids"])


def test_oracle_video_mask_outputs_match_reference_golden(golden):
    case, g = C.MASKGIT_CASES["small"], golden("maskgit_small_vmask")
    torch.manual_seed(case["seed"])
    sd = P.MaskGit(**case["ctor"]).state_dict()
    assert C.state_digest(sd) == g["state_digest"]
    ids, ctx = C.token_inputs(case, case["ctor"]["num_tokens"])
    vmask = C.video_mask_of(C.VIDEO_MASK_VALID["maskgit_small"], ids.shape[1])
    with torch.no_grad():
        out = O.maskgit_forward(ids, sd, video_patch_shape=case["patch_shape"], heads=case["ctor"]["heads"], context=ctx,
                                text_mask=torch.any(ctx != 0, dim=-1), video_mask=vmask)
    torch.testing.assert_close(out, g["cond"], rtol=1e-5, atol=1e-5)
    case, g = C.CRITIC_CASES["small"], golden("critic_small_vmask")
    torch.manual_seed(case["seed"])
    sd = P.TokenCritic(**case["ctor"]).state_dict()
    assert C.state_digest(sd) == g["state_digest"]
    ids, ctx = C.token_inputs(case, case["ctor"]["num_tokens"])
    vmask = C.video_mask_of(C.VIDEO_MASK_VALID["critic_small"], ids.shape[1])
    with torch.no_grad():
        out = O.critic_forward(ids, sd, video_patch_shape=case["patch_shape"], heads=case["ctor"]["heads"], context=ctx,
                               text_mask=torch.any(ctx != 0, dim=-1), video_mask=vmask)
    torch.testing.assert_close(out, g["cond"], rtol=1e-5, atol=1e-5)


def test_oracle_frame_mask_training_loss_matches_reference_golden(golden):
    case, g = C.FRAME_MASK_TRAIN_CASE, golden("train_frame_mask")
    torch.manual_seed(case["seed"])
    P.CViViT(**C.SAMPLE_CVIVIT)
    mg = P.MaskGit(**case["maskgit"])
    assert C.state_digest(mg.state_dict()) == g["maskgit_digest"]
    ctx = C.synthetic_text_embeds(case["batch"], case["ctx_len"], case["maskgit"]["dim_context"], case["ctx_valid"],
                                  case["input_seed"] + 1000)
    b, n = g["ids"].shape[0], g["ids"][0].numel()
    torch.manual_seed(case["noise_seed"])
    rand_step, u = O.train_draws(b, n, case["steps"])
    token_mask = O.train_token_mask(rand_step, u, case["steps"], video_mask=g["token_valid"])
    # (the reference's subset rule ranks ALL positions and only shifts the ranks by the pad count, phenaki_pytorch.py:43-55:
    # padded positions can be chosen -- replicated, not "fixed")
    assert torch.equal(token_mask, g["token_mask"])
    with torch.no_grad():
        ce = O.maskgit_train_loss(g["ids"].reshape(b, n), mg.state_dict(), token_mask, video_patch_shape=case["patch_shape"],
                                  heads=case["maskgit"]["heads"], context=ctx, text_mask=torch.any(ctx != 0, dim=-1),
                                  video_mask=g["token_valid"])
    torch.testing.assert_close(ce, g["ce"], rtol=1e-5, atol=1e-6)
