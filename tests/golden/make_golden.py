"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference
with the stubs in oracle/reference_loader.py) and, in the same pass, PINS the oracle restatement
(oracle/phenaki_oracle.py) against it: every integer output (token ids, masks) must be identical and every
float output equal to within fp32 summation order (<=1e-5; bit-identical in all but one case, see
``same``) to the reference's on CPU fp32, otherwise this script aborts.

Run in the build container only:   python tests/golden/make_golden.py
The GPU box never runs this (no /root/reference there); it only reads the committed fixtures.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import phenaki_oracle as O  # noqa: E402
from oracle.reference_loader import load_reference  # noqa: E402
from tests import cases as C  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def same(a, b, what):
    """Integer outputs must be identical.  Float outputs are bit-identical in practice when both sides
    hit the same BLAS kernel; MKL picks kernels by operand address/stride, so a <=1e-5 (fp32 rounding
    order) difference is accepted and REPORTED."""
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if torch.equal(a, b):
        print(f"  pinned: {what} {tuple(a.shape)} bit-identical")
        return
    if not a.is_floating_point():
        raise SystemExit(f"ORACLE != REFERENCE for integer output {what}")
    d = (a - b).abs().max().item()
    if not torch.allclose(a, b, rtol=1e-5, atol=1e-5):
        raise SystemExit(f"ORACLE != REFERENCE for {what}: max abs diff {d}")
    print(f"  pinned: {what} {tuple(a.shape)} max abs diff {d:.2e} (fp32 summation order only)")


def pair(v):
    return v if isinstance(v, tuple) else (v, v)


def hook_outputs(mod_map):
    store, handles = {}, []
    for name, mod in mod_map.items():
        handles.append(mod.register_forward_hook(
            lambda m, i, o, name=name: store.__setitem__(name, o.detach().clone())))
    return store, handles


def make_cvivit(ref):
    for name, case in C.CVIVIT_CASES.items():
        print(f"[cvivit/{name}]")
        torch.manual_seed(case["seed"])
        model = ref.CViViT(**case["ctor"]).eval()
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        video = C.seeded_randn(case["video"], case["video_seed"])
        lfq = case["ctor"].get("lookup_free_quantization", True)
        mods = {"patch_first": model.to_patch_emb_first_frame, "patch_rest": model.to_patch_emb,
                "spatial": model.enc_spatial_transformer, "temporal": model.enc_temporal_transformer}
        if lfq:
            mods["proj"] = model.vq.project_in
        taps, handles = hook_outputs(mods)
        with torch.no_grad():
            ids = model(video, return_only_codebook_ids=True)
            recon = model.decode_from_codebook_indices(ids.reshape(ids.shape[0], -1))
        for h in handles:
            h.remove()
        image_size, patch_size = pair(case["ctor"]["image_size"]), pair(case["ctor"]["patch_size"])
        with torch.no_grad():
            o_ids, o_proj = O.cvivit_codebook_ids(video, sd, image_size, patch_size, return_margin=True)
            dim, heads, pt, ch = O.cvivit_geometry(sd, image_size, patch_size)
            v5 = video if video.ndim == 5 else video.unsqueeze(2)
            o_patch = O.cvivit_patch_embed(v5, sd, patch_size, pt)
            o_recon = O.cvivit_decode_from_ids(ids.reshape(ids.shape[0], -1), sd, image_size, patch_size)
        same(o_ids, ids, "codebook ids")
        if lfq:
            same(o_proj.reshape(taps["proj"].shape), taps["proj"], "LFQ pre-sign projection")
        else:  # cosine-sim codebook: the similarities (their top-2 gap is the margin of an id) come from the oracle
            taps["proj"] = o_proj.reshape(o_proj.shape[0], -1, o_proj.shape[-1])
        same(o_patch[:, :1], taps["patch_first"], "first-frame patch embed")
        if "patch_rest" in taps:
            same(o_patch[:, 1:], taps["patch_rest"], "rest-frames patch embed")
        same(o_recon, recon, "decode_from_codebook_indices")
        gold = dict(state_digest=C.state_digest(sd), n_state=len(sd), ids=ids, proj=taps["proj"],
                    patch=o_patch, spatial=taps["spatial"], temporal=taps["temporal"], recon=recon)
        torch.save(gold, os.path.join(OUT, f"cvivit_{name}.pt"))


def make_maskgit(ref):
    for name, case in C.MASKGIT_CASES.items():
        print(f"[maskgit/{name}]")
        torch.manual_seed(case["seed"])
        model = ref.MaskGit(**case["ctor"]).eval()
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        ids, ctx = C.token_inputs(case, case["ctor"]["num_tokens"])
        tmask = torch.any(ctx != 0, dim=-1)
        heads = case["ctor"].get("heads", 8)
        with torch.no_grad():
            cond = model(ids, cond_drop_prob=0.0, text_mask=tmask, video_patch_shape=case["patch_shape"], context=ctx)
            null = model(ids, cond_drop_prob=1.0, text_mask=tmask, video_patch_shape=case["patch_shape"], context=ctx)
            emb = model(ids, text_mask=tmask, video_patch_shape=case["patch_shape"], context=ctx, return_embeds=True)
            cfg = model.forward_with_cond_scale(ids, cond_scale=3.0, text_mask=tmask,
                                                video_patch_shape=case["patch_shape"], context=ctx)
            bias = model.continuous_pos_bias(*case["patch_shape"])
            kw = dict(video_patch_shape=case["patch_shape"], heads=heads, context=ctx, text_mask=tmask)
            o_cond = O.maskgit_forward(ids, sd, **kw)
            o_null = O.maskgit_forward(ids, sd, cond_drop=True, **kw)
            o_emb = O.maskgit_forward(ids, sd, return_embeds=True, **kw)
            o_cfg = O.with_cond_scale(lambda cond_drop: O.maskgit_forward(ids, sd, cond_drop=cond_drop, **kw), 3.0)
            o_bias = O.continuous_position_bias(sd, "continuous_pos_bias.", case["patch_shape"])
        same(o_cond, cond, "logits (cond)")
        same(o_null, null, "logits (null)")
        same(o_emb, emb, "embeds")
        same(o_cfg, cfg, "CFG logits")
        same(o_bias, bias, "3-D continuous position bias")
        torch.save(dict(state_digest=C.state_digest(sd), cond=cond, null=null, embeds=emb, cfg=cfg,
                        bias=bias), os.path.join(OUT, f"maskgit_{name}.pt"))


def make_critic(ref):
    for name, case in C.CRITIC_CASES.items():
        print(f"[critic/{name}]")
        torch.manual_seed(case["seed"])
        model = ref.TokenCritic(**case["ctor"]).eval()
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        ids, ctx = C.token_inputs(case, case["ctor"]["num_tokens"])
        tmask = torch.any(ctx != 0, dim=-1)
        heads = case["ctor"].get("heads", 8)
        with torch.no_grad():
            cond = model(ids, cond_drop_prob=0.0, text_mask=tmask, video_patch_shape=case["patch_shape"], context=ctx)
            cfg = model.forward_with_cond_scale(ids, cond_scale=5.0, text_mask=tmask,
                                                video_patch_shape=case["patch_shape"], context=ctx)
            kw = dict(video_patch_shape=case["patch_shape"], heads=heads, context=ctx, text_mask=tmask)
            o_cond = O.critic_forward(ids, sd, **kw)
            o_cfg = O.with_cond_scale(lambda cond_drop: O.critic_forward(ids, sd, cond_drop=cond_drop, **kw), 5.0)
        same(o_cond, cond, "critic scores (cond)")
        same(o_cfg, cfg, "critic scores (CFG 5)")
        torch.save(dict(state_digest=C.state_digest(sd), cond=cond, cfg=cfg),
                   os.path.join(OUT, f"critic_{name}.pt"))


def make_sample(ref):
    for name, case in C.SAMPLE_CASES.items():
        print(f"[sample/{name}]")
        torch.manual_seed(case["seed"])
        cvivit = ref.CViViT(**C.SAMPLE_CVIVIT)
        maskgit = ref.MaskGit(**C.SAMPLE_MASKGIT)
        critic = ref.TokenCritic(**C.SAMPLE_CRITIC) if case["critic"] else None
        phenaki = ref.Phenaki(cvivit=cvivit, maskgit=maskgit, critic=critic, steps=case["steps"],
                              text_embed_dim=C.SAMPLE_MASKGIT["dim_context"]).eval()
        ctx = C.synthetic_text_embeds(case["batch"], case["ctx_len"], C.SAMPLE_MASKGIT["dim_context"],
                                      case["ctx_valid"], case["seed"] + 1000)
        phenaki.encode_texts = lambda texts, output_device=None: ctx
        captured = {}
        orig_decode = phenaki.cvivit.decode_from_codebook_indices
        phenaki.cvivit.decode_from_codebook_indices = lambda ids: (captured.__setitem__("ids", ids.clone()), orig_decode(ids))[1]
        prime = None
        if case["prime"]:
            prime = C.seeded_randn((case["batch"], 3, case["prime_frames"], *C.SAMPLE_CVIVIT["image_size"]),
                                   case["seed"] + 2000)
        texts = [f"prompt {i}" for i in range(case["batch"])]
        torch.manual_seed(case["noise_seed"])
        video = phenaki.sample(texts=texts, num_frames=case["num_frames"], prime_frames=prime,
                               cond_scale=case["cond_scale"])
        # oracle replay
        cv_sd = {k: v.detach().clone() for k, v in phenaki.cvivit.state_dict().items()}
        mg_sd = {k: v.detach().clone() for k, v in maskgit.state_dict().items()}
        cr_sd = {k: v.detach().clone() for k, v in critic.state_dict().items()} if critic else None
        image_size, patch_size = C.SAMPLE_CVIVIT["image_size"], C.SAMPLE_CVIVIT["patch_size"]
        pt = C.SAMPLE_CVIVIT["temporal_patch_size"]
        hh, ww = image_size[0] // patch_size[0], image_size[1] // patch_size[1]
        tape = C.NoiseTape(case["noise_seed"])
        trace = []
        with torch.no_grad():
            prime_ids, pf = None, 0
            if prime is not None:
                prime_ids = O.cvivit_codebook_ids(prime, cv_sd, image_size, patch_size).reshape(case["batch"], -1)
                pf = prime.shape[2]
            nf = case["num_frames"]
            num_tokens = (hh * ww) * ((nf - 1) // pt + 1) if prime is None else (hh * ww) * (nf // pt)
            patch_shape = (1 + (nf + pf - 1) // pt, hh, ww)
            ids = O.sample_token_ids(mg_sd, num_tokens=num_tokens, patch_shape=patch_shape,
                                     batch=case["batch"], steps=case["steps"], heads=C.SAMPLE_MASKGIT["heads"],
                                     text_embeds=ctx, prime_ids=prime_ids, cond_scale=case["cond_scale"],
                                     critic_sd=cr_sd, noise_fn=tape, trace=trace)
            full = ids if prime_ids is None else torch.cat((prime_ids, ids), dim=-1)
            o_video = O.cvivit_decode_from_ids(full, cv_sd, image_size, patch_size)[:, :, pf:]
        same(full, captured["ids"], "final token ids of Phenaki.sample")
        same(o_video, video, "sampled video")
        torch.save(dict(cvivit_digest=C.state_digest(cv_sd), maskgit_digest=C.state_digest(mg_sd),
                        critic_digest=C.state_digest(cr_sd) if cr_sd else None,
                        final_ids=ids, prime_ids=prime_ids, video=video, patch_shape=patch_shape,
                        num_tokens=num_tokens,
                        trace=[{k: v for k, v in t.items()} for t in trace]),
                   os.path.join(OUT, f"sample_{name}.pt"))


def make_makevideo(ref):
    """make_video (phenaki_pytorch.py:691-714): scenes chained by priming; reference vs oracle replay of the whole chain."""
    case = C.MAKE_VIDEO_CASE
    print("[make_video]")
    torch.manual_seed(case["seed"])
    cvivit = ref.CViViT(**C.SAMPLE_CVIVIT)
    maskgit = ref.MaskGit(**C.SAMPLE_MASKGIT)
    critic = ref.TokenCritic(**C.SAMPLE_CRITIC)
    phenaki = ref.Phenaki(cvivit=cvivit, maskgit=maskgit, critic=critic, steps=case["steps"],
                          text_embed_dim=C.SAMPLE_MASKGIT["dim_context"]).eval()
    table = C.make_video_text_table(case)
    phenaki.encode_texts = lambda texts, output_device=None: table[texts[0]]
    torch.manual_seed(case["noise_seed"])
    video, scenes = ref.make_video(phenaki, texts=list(case["texts"]), num_frames=case["num_frames"],
                                   prime_lengths=case["prime_lengths"])
    # oracle replay of the chain with the same uniform stream
    cv_sd = {k: v.detach().clone() for k, v in phenaki.cvivit.state_dict().items()}
    mg_sd = {k: v.detach().clone() for k, v in maskgit.state_dict().items()}
    cr_sd = {k: v.detach().clone() for k, v in critic.state_dict().items()}
    from tests.chain import oracle_make_video
    o_video, o_scenes = oracle_make_video(case, cv_sd, mg_sd, cr_sd, table, C.NoiseTape(case["noise_seed"]))
    for i, (a, b) in enumerate(zip(o_scenes, scenes)):
        same(a, b, f"make_video scene {i}")
    same(o_video, video, "make_video whole video")
    torch.save(dict(cvivit_digest=C.state_digest(cv_sd), maskgit_digest=C.state_digest(mg_sd),
                    critic_digest=C.state_digest(cr_sd), video=video, scenes=scenes),
               os.path.join(OUT, "make_video.pt"))


def make_units(ref):
    """Kernel-granularity goldens straight from reference attention.py modules."""
    from phenaki_pytorch import attention as A
    print("[units]")
    gold = {}
    torch.manual_seed(40)
    # self attention with bias + key mask
    att = A.Attention(dim=64, dim_head=32, heads=2).eval()
    x = C.seeded_randn((2, 12, 64), 41)
    bias = C.seeded_randn((2, 12, 12), 42)
    m = torch.ones(2, 12, dtype=torch.bool)
    m[1, 9:] = False
    sd = {"a." + k: v.detach() for k, v in att.state_dict().items()}
    with torch.no_grad():
        y = att(x, attn_bias=bias, mask=m)
        same(O.attention(x, sd, "a.", heads=2, attn_bias=bias, mask=m), y, "unit self-attn bias+mask")
    gold["self"] = dict(y=y)
    # causal + alibi
    att = A.Attention(dim=64, dim_head=32, heads=2, causal=True).eval()
    sd = {"a." + k: v.detach() for k, v in att.state_dict().items()}
    with torch.no_grad():
        y = att(x)
        same(O.attention(x, sd, "a.", heads=2, causal=True), y, "unit causal+alibi attn")
    gold["causal"] = dict(y=y)
    # cross attention with null kv + context mask
    att = A.Attention(dim=64, dim_context=48, dim_head=32, heads=2, num_null_kv=2).eval()
    sd = {"a." + k: v.detach() for k, v in att.state_dict().items()}
    ctx = C.seeded_randn((2, 5, 48), 43)
    cm = torch.tensor([[True] * 5, [True, True, False, False, False]])
    with torch.no_grad():
        y = att(x, context=ctx, mask=cm)
        same(O.attention(x, sd, "a.", heads=2, num_null_kv=2, context=ctx, mask=cm), y, "unit cross attn")
    gold["cross"] = dict(y=y)
    # PEG, both paddings, with the raw-reshape quirk ((b h w) t d buffer, shape (b,t,h,w))
    for causal in (True, False):
        pg = A.PEG(dim=16, causal=causal).eval()
        sd = {"p." + k: v.detach() for k, v in pg.state_dict().items()}
        xx = C.seeded_randn((2 * 3 * 2, 4, 16), 44)  # (b h w) t d with b=2,h=3,w=2,t=4
        with torch.no_grad():
            y = pg(xx, shape=(2, 4, 3, 2))
            same(O.peg(xx, (2, 4, 3, 2), sd, "p.", causal), y, f"unit PEG causal={causal}")
        gold[f"peg_{causal}"] = dict(y=y)
    torch.save(gold, os.path.join(OUT, "units.pt"))


def make_train(ref):
    """Phenaki.forward (training loss, phenaki_pytorch.py:562-687) and its gradients: the reference's autograd vs
    autograd through the functional oracle, on the same weights, inputs and random draws."""
    for name, case in C.TRAIN_CASES.items():
        print(f"[train/{name}]")
        torch.manual_seed(case["seed"])
        cvivit = ref.CViViT(**C.SAMPLE_CVIVIT)
        maskgit = ref.MaskGit(**case["maskgit"])
        critic = ref.TokenCritic(**case["critic"]) if case["critic"] else None
        self_critic = case.get("self_critic", False)
        phenaki = ref.Phenaki(cvivit=cvivit, maskgit=maskgit, critic=critic, steps=case["steps"],
                              self_token_critic=self_critic,
                              text_embed_dim=case["maskgit"]["dim_context"]).train()
        ids, ctx = C.train_inputs(case)
        heads = case["maskgit"].get("heads", 8)
        b, n = ids.shape[0], ids[0].numel()
        torch.manual_seed(case["noise_seed"])
        loss = phenaki(video_codebook_ids=ids, text_embeds=ctx)
        loss.backward()
        mg_sd = {k: v.detach().clone() for k, v in maskgit.state_dict().items()}
        cr_sd = {k: v.detach().clone() for k, v in critic.state_dict().items()} if critic else None
        mg_grads = {k: p.grad.detach().clone() for k, p in maskgit.named_parameters() if p.grad is not None}
        cr_grads = ({k: p.grad.detach().clone() for k, p in critic.named_parameters() if p.grad is not None}
                    if critic else None)
        no_grad = sorted(k for k, p in maskgit.named_parameters() if p.grad is None)
        print("  maskgit parameters without gradient:", no_grad)

        # ---- oracle replay: same draws (global generator, reference order), autograd over the functional restatement
        def leaf(sd):
            return {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}

        o_mg, o_cr = leaf(mg_sd), (leaf(cr_sd) if cr_sd else None)
        tmask = torch.any(ctx != 0, dim=-1)
        torch.manual_seed(case["noise_seed"])
        rand_step, u = O.train_draws(b, n, case["steps"])
        token_mask = O.train_token_mask(rand_step, u, case["steps"])
        flat = ids.reshape(b, n)
        kw = dict(video_patch_shape=case["patch_shape"], heads=heads, context=ctx, text_mask=tmask)
        o_loss, logits = O.maskgit_train_loss(flat, o_mg, token_mask, return_logits=True, **kw)
        gold = dict(maskgit_digest=C.state_digest(mg_sd), token_mask=token_mask, rand_step=rand_step)
        if critic is not None:
            gu = torch.zeros_like(logits).uniform_(0, 1)  # gumbel_noise (:83-86) draws from the global generator
            pred = O.gumbel_sample(logits.detach(), phenaki.critic_train_sample_temperature, gu)
            bce = O.critic_train_loss(flat, pred, token_mask, o_cr, **kw)
            gold.update(critic_digest=C.state_digest(cr_sd), pred_ids=pred, ce=o_loss.detach().clone(),
                        bce=bce.detach().clone())
            o_loss = o_loss + bce * phenaki.critic_loss_weight
        if self_critic:
            lin = phenaki.critic.to_pred[0]
            o_w, o_b = lin.weight.detach().clone().requires_grad_(True), lin.bias.detach().clone().requires_grad_(True)
            gu = torch.zeros_like(logits).uniform_(0, 1)
            pred = O.gumbel_sample(logits.detach(), phenaki.critic_train_sample_temperature, gu)
            bce = O.self_critic_train_loss(flat, pred, token_mask, o_mg, o_w, o_b, **kw)
            gold.update(pred_ids=pred, ce=o_loss.detach().clone(), bce=bce.detach().clone(),
                        to_pred_weight=lin.weight.detach().clone(), to_pred_bias=lin.bias.detach().clone(),
                        to_pred_grads=dict(weight=lin.weight.grad.detach().clone(), bias=lin.bias.grad.detach().clone()))
            o_loss = o_loss + bce * phenaki.critic_loss_weight
        o_loss.backward()
        if self_critic:
            same(o_w.grad, lin.weight.grad, "d loss / d to_pred.weight")
            same(o_b.grad, lin.bias.grad, "d loss / d to_pred.bias")
        same(o_loss.detach(), loss.detach(), "training loss")
        for k, g in mg_grads.items():
            same(o_mg[k].grad, g, f"d loss / d maskgit.{k}")
        if critic is not None:
            for k, g in cr_grads.items():
                same(o_cr[k].grad, g, f"d loss / d critic.{k}")
        gold.update(loss=loss.detach().clone(), maskgit_grads=mg_grads, critic_grads=cr_grads)
        torch.save(gold, os.path.join(OUT, f"train_{name}.pt"))


def make_selfcritic(ref):
    """Phenaki(self_token_critic=True).sample: reference vs oracle replay (SelfCritic = Linear(dim,1) on MaskGit embeds)."""
    case = C.SELF_CRITIC_SAMPLE_CASE
    print("[sample/self_critic]")
    torch.manual_seed(case["seed"])
    cvivit = ref.CViViT(**C.SAMPLE_CVIVIT)
    maskgit = ref.MaskGit(**C.SAMPLE_MASKGIT)
    phenaki = ref.Phenaki(cvivit=cvivit, maskgit=maskgit, self_token_critic=True, steps=case["steps"],
                          text_embed_dim=C.SAMPLE_MASKGIT["dim_context"]).eval()
    ctx = C.synthetic_text_embeds(case["batch"], case["ctx_len"], C.SAMPLE_MASKGIT["dim_context"], case["ctx_valid"],
                                  case["seed"] + 1000)
    phenaki.encode_texts = lambda texts, output_device=None: ctx
    captured = {}
    orig_decode = phenaki.cvivit.decode_from_codebook_indices
    phenaki.cvivit.decode_from_codebook_indices = lambda ids: (captured.__setitem__("ids", ids.clone()), orig_decode(ids))[1]
    torch.manual_seed(case["noise_seed"])
    video = phenaki.sample(texts=[f"prompt {i}" for i in range(case["batch"])], num_frames=case["num_frames"],
                           cond_scale=case["cond_scale"])
    mg_sd = {k: v.detach().clone() for k, v in maskgit.state_dict().items()}
    cv_sd = {k: v.detach().clone() for k, v in phenaki.cvivit.state_dict().items()}
    lin = phenaki.critic.to_pred[0]
    w, b = lin.weight.detach().clone(), lin.bias.detach().clone()
    image_size, patch_size, pt = C.SAMPLE_CVIVIT["image_size"], C.SAMPLE_CVIVIT["patch_size"], C.SAMPLE_CVIVIT["temporal_patch_size"]
    hh, ww = image_size[0] // patch_size[0], image_size[1] // patch_size[1]
    nf = case["num_frames"]
    num_tokens, patch_shape = hh * ww * ((nf - 1) // pt + 1), (1 + (nf - 1) // pt, hh, ww)
    trace = []
    with torch.no_grad():
        ids = O.sample_token_ids(mg_sd, num_tokens=num_tokens, patch_shape=patch_shape, batch=case["batch"],
                                 steps=case["steps"], heads=C.SAMPLE_MASKGIT["heads"], text_embeds=ctx,
                                 cond_scale=case["cond_scale"], self_critic=(w, b),
                                 noise_fn=C.NoiseTape(case["noise_seed"]), trace=trace)
        o_video = O.cvivit_decode_from_ids(ids, cv_sd, image_size, patch_size)
    same(ids, captured["ids"], "final token ids of Phenaki.sample (self critic)")
    same(o_video, video, "sampled video (self critic)")
    torch.save(dict(cvivit_digest=C.state_digest(cv_sd), maskgit_digest=C.state_digest(mg_sd), to_pred_weight=w,
                    to_pred_bias=b, final_ids=ids, video=video, patch_shape=patch_shape, num_tokens=num_tokens,
                    trace=[{k: v for k, v in t.items()} for t in trace]), os.path.join(OUT, "sample_self_critic.pt"))


def make_vmask(ref):
    """MaskGit / TokenCritic forward with a video_mask (key mask of the self-attention, attention.py:164-167)."""
    case = C.MASKGIT_CASES["small"]
    print("[maskgit/small + video_mask]")
    torch.manual_seed(case["seed"])
    model = ref.MaskGit(**case["ctor"]).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ids, ctx = C.token_inputs(case, case["ctor"]["num_tokens"])
    tmask = torch.any(ctx != 0, dim=-1)
    vmask = C.video_mask_of(C.VIDEO_MASK_VALID["maskgit_small"], ids.shape[1])
    heads = case["ctor"].get("heads", 8)
    with torch.no_grad():
        cond = model(ids, cond_drop_prob=0.0, text_mask=tmask, video_mask=vmask, video_patch_shape=case["patch_shape"], context=ctx)
        cfg = model.forward_with_cond_scale(ids, cond_scale=3.0, text_mask=tmask, video_mask=vmask,
                                            video_patch_shape=case["patch_shape"], context=ctx)
        kw = dict(video_patch_shape=case["patch_shape"], heads=heads, context=ctx, text_mask=tmask, video_mask=vmask)
        o_cond = O.maskgit_forward(ids, sd, **kw)
        o_cfg = O.with_cond_scale(lambda cond_drop: O.maskgit_forward(ids, sd, cond_drop=cond_drop, **kw), 3.0)
    same(o_cond, cond, "logits (cond, video_mask)")
    same(o_cfg, cfg, "CFG logits (video_mask)")
    torch.save(dict(state_digest=C.state_digest(sd), cond=cond, cfg=cfg), os.path.join(OUT, "maskgit_small_vmask.pt"))

    case = C.CRITIC_CASES["small"]
    print("[critic/small + video_mask]")
    torch.manual_seed(case["seed"])
    model = ref.TokenCritic(**case["ctor"]).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ids, ctx = C.token_inputs(case, case["ctor"]["num_tokens"])
    tmask = torch.any(ctx != 0, dim=-1)
    vmask = C.video_mask_of(C.VIDEO_MASK_VALID["critic_small"], ids.shape[1])
    heads = case["ctor"].get("heads", 8)
    with torch.no_grad():
        cond = model(ids, cond_drop_prob=0.0, text_mask=tmask, video_mask=vmask, video_patch_shape=case["patch_shape"], context=ctx)
        cfg = model.forward_with_cond_scale(ids, cond_scale=5.0, text_mask=tmask, video_mask=vmask,
                                            video_patch_shape=case["patch_shape"], context=ctx)
        kw = dict(video_patch_shape=case["patch_shape"], heads=heads, context=ctx, text_mask=tmask, video_mask=vmask)
        o_cond = O.critic_forward(ids, sd, **kw)
        o_cfg = O.with_cond_scale(lambda cond_drop: O.critic_forward(ids, sd, cond_drop=cond_drop, **kw), 5.0)
    same(o_cond, cond, "critic scores (cond, video_mask)")
    same(o_cfg, cfg, "critic scores (CFG 5, video_mask)")
    torch.save(dict(state_digest=C.state_digest(sd), cond=cond, cfg=cfg), os.path.join(OUT, "critic_small_vmask.pt"))


def make_framemask(ref):
    """Phenaki.forward(videos, video_frame_mask=...): live tokenisation, frame mask -> token mask
    (cvivit.py:365-373), masked-subset sampling among the valid tokens, key-masked attention; loss + every gradient."""
    case = C.FRAME_MASK_TRAIN_CASE
    print("[train/frame_mask]")
    torch.manual_seed(case["seed"])
    cvivit = ref.CViViT(**C.SAMPLE_CVIVIT)
    maskgit = ref.MaskGit(**case["maskgit"])
    critic = ref.TokenCritic(**case["critic"])
    phenaki = ref.Phenaki(cvivit=cvivit, maskgit=maskgit, critic=critic, steps=case["steps"],
                          text_embed_dim=case["maskgit"]["dim_context"]).train()
    videos = C.seeded_randn(case["video"], case["input_seed"])
    ctx = C.synthetic_text_embeds(case["batch"], case["ctx_len"], case["maskgit"]["dim_context"], case["ctx_valid"],
                                  case["input_seed"] + 1000)
    fmask = C.frame_mask_of(case["frames_valid"], case["video"][2])
    torch.manual_seed(case["noise_seed"])
    loss = phenaki(videos, text_embeds=ctx, video_frame_mask=fmask)
    loss.backward()
    heads = case["maskgit"].get("heads", 8)
    mg_sd = {k: v.detach().clone() for k, v in maskgit.state_dict().items()}
    cr_sd = {k: v.detach().clone() for k, v in critic.state_dict().items()}
    cv_sd = {k: v.detach().clone() for k, v in phenaki.cvivit.state_dict().items()}
    mg_grads = {k: p.grad.detach().clone() for k, p in maskgit.named_parameters() if p.grad is not None}
    cr_grads = {k: p.grad.detach().clone() for k, p in critic.named_parameters() if p.grad is not None}

    def leaf(sd):
        return {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}

    o_mg, o_cr = leaf(mg_sd), leaf(cr_sd)
    with torch.no_grad():
        ids = O.cvivit_codebook_ids(videos, cv_sd, C.SAMPLE_CVIVIT["image_size"], C.SAMPLE_CVIVIT["patch_size"])
        token_valid = phenaki.cvivit.calculate_video_token_mask(videos, video_frame_mask=fmask)
    b, n = ids.shape[0], ids[0].numel()
    flat = ids.reshape(b, n)
    tmask = torch.any(ctx != 0, dim=-1)
    torch.manual_seed(case["noise_seed"])
    rand_step, u = O.train_draws(b, n, case["steps"])
    token_mask = O.train_token_mask(rand_step, u, case["steps"], video_mask=token_valid)
    kw = dict(video_patch_shape=case["patch_shape"], heads=heads, context=ctx, text_mask=tmask, video_mask=token_valid)
    o_loss, logits = O.maskgit_train_loss(flat, o_mg, token_mask, return_logits=True, **kw)
    gu = torch.zeros_like(logits).uniform_(0, 1)
    pred = O.gumbel_sample(logits.detach(), phenaki.critic_train_sample_temperature, gu)
    bce = O.critic_train_loss(flat, pred, token_mask, o_cr, **kw)
    total = o_loss + bce * phenaki.critic_loss_weight
    total.backward()
    same(total.detach(), loss.detach(), "training loss (frame mask)")
    for k, g in mg_grads.items():
        same(o_mg[k].grad, g, f"d loss / d maskgit.{k}")
    for k, g in cr_grads.items():
        same(o_cr[k].grad, g, f"d loss / d critic.{k}")
    torch.save(dict(maskgit_digest=C.state_digest(mg_sd), critic_digest=C.state_digest(cr_sd), ids=ids,
                    token_valid=token_valid, token_mask=token_mask, rand_step=rand_step, pred_ids=pred,
                    ce=o_loss.detach().clone(), bce=bce.detach().clone(), loss=loss.detach().clone(),
                    maskgit_grads=mg_grads, critic_grads=cr_grads), os.path.join(OUT, "train_frame_mask.pt"))


if __name__ == "__main__":
    ref = load_reference()
    which = sys.argv[1:] or ["units", "cvivit", "maskgit", "critic", "sample", "makevideo", "train", "selfcritic", "vmask", "framemask"]
    for w in which:
        globals()["make_" + w](ref)
    print("golden fixtures written to", OUT)
