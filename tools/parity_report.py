"""Parity of the benchmarked modes against the reference's own GPU paths, at BASELINE configs[1] / configs[2] sizes.

Runs on the GPU box (no /root/reference there): the functional oracle (bit-identical to the unmodified reference on the
CPU, tests/golden/make_golden.py) is executed ON CUDA, once in eager fp32 and once under torch.autocast(bfloat16) --
the reference's only bf16 path (Linear / einsum / conv in bf16, LayerNorm / softmax / norm in fp32, fp32
residual stream).  Every product precision mode is compared with both.

    python tools/parity_report.py [out.json]           # prints one JSON object, also written to out.json

Used by tests/test_gpu_parity_at_size.py (thresholds) and quoted in DESIGN.md section 2.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import phenaki_oracle as O  # noqa: E402  (test infrastructure: this tool is a checker)
import phenaki_pytorch_b200 as P  # noqa: E402
from phenaki_pytorch_b200 import _lib as L  # noqa: E402

DEV = "cuda"
CFG2 = dict(dim=512, codebook_size=65536, image_size=256, patch_size=32, temporal_patch_size=2, spatial_depth=4,
            temporal_depth=4, dim_head=64, heads=8, use_vgg_and_gan=False)
CFG3 = dict(dim=512, num_tokens=65536, max_seq_len=1024, dim_context=768, depth=6)


def _exact_fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def bit_agreement(a, b, bits=16):
    x = (a ^ b).reshape(-1)
    return 1.0 - sum(int(((x >> k) & 1).sum()) for k in range(bits)) / (x.numel() * bits)


def id_agreement(a, b):
    return float((a == b).float().mean())


def flips_by_margin(ids, ref_ids, ref_proj, bits=16):
    """Flipped LFQ bits against the fp32 reference, with the reference's own pre-sign margin |x| of every flipped bit."""
    diff = (ids ^ ref_ids).reshape(-1)
    proj = ref_proj.reshape(-1, bits).abs()
    shifts = torch.arange(bits - 1, -1, -1, device=diff.device)
    flipped = ((diff[:, None] >> shifts[None, :]) & 1).bool()
    m = proj[flipped]
    if m.numel() == 0:
        return dict(flipped_bits=0, max_margin=0.0, p99_margin=0.0)
    return dict(flipped_bits=int(m.numel()), max_margin=float(m.max()), p99_margin=float(m.float().quantile(0.99)))


def modes():
    out = {"fp32": L.PREC_F32, "bf16": L.PREC_BF16}
    if hasattr(L, "PREC_BF16X3"):
        out["bf16x3"] = L.PREC_BF16X3
    return out


def cfg2_report(batch=8):
    _exact_fp32()
    torch.manual_seed(0)
    model = P.CViViT(**CFG2).eval()
    sd = {k: v.detach().clone().to(DEV) for k, v in model.state_dict().items()}
    video = torch.randn((batch, 3, 17, 256, 256), generator=torch.Generator().manual_seed(1)).to(DEV)
    model = model.to(DEV)
    with torch.no_grad():
        ref32, proj32 = O.cvivit_codebook_ids(video, sd, (256, 256), (32, 32), return_margin=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ref16, proj16 = O.cvivit_codebook_ids(video, sd, (256, 256), (32, 32), return_margin=True)
    ref32, ref16 = ref32.long(), ref16.long()
    rep = dict(videos=batch, tokens=int(ref32.numel()),
               reference_autocast_bf16_vs_reference_fp32=dict(
                   bit_agreement=bit_agreement(ref16, ref32), id_agreement=id_agreement(ref16, ref32),
                   **flips_by_margin(ref16, ref32, proj32.float())))
    for name, prec in modes().items():
        model.precision = prec
        ids = model(video, return_only_codebook_ids=True)
        rep[f"ours_{name}_vs_reference_fp32"] = dict(bit_agreement=bit_agreement(ids, ref32),
                                                     id_agreement=id_agreement(ids, ref32),
                                                     **flips_by_margin(ids, ref32, proj32.float()))
        rep[f"ours_{name}_vs_reference_autocast_bf16"] = dict(bit_agreement=bit_agreement(ids, ref16),
                                                              id_agreement=id_agreement(ids, ref16))
    return rep


def cfg2_cosine_vq_report(batch=8):
    """SURVEY 8f-3 at size: the cosine-sim VectorQuantize tokenizer (lookup_free_quantization=False, cvivit.py:321,
    564-570) with the configs[1] encoder and K = 65536 codes: ids of all videos against the oracle on CUDA (fp32 and
    autocast-bf16), plus the time of the whole encode call and of the nearest-code search alone."""
    _exact_fp32()
    torch.manual_seed(6)
    model = P.CViViT(**dict(CFG2, lookup_free_quantization=False)).eval()
    sd = {k: v.detach().clone().to(DEV) for k, v in model.state_dict().items()}
    video = torch.randn((batch, 3, 17, 256, 256), generator=torch.Generator().manual_seed(7)).to(DEV)
    model = model.to(DEV)
    with torch.no_grad():
        ref32, sims = O.cvivit_codebook_ids(video, sd, (256, 256), (32, 32), return_margin=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ref16 = O.cvivit_codebook_ids(video, sd, (256, 256), (32, 32))
    sims = sims.reshape(-1, sims.shape[-1]).float()
    top2 = sims.topk(2, dim=-1).values
    rep = dict(videos=batch, tokens=int(ref32.numel()), codebook_size=int(sims.shape[-1]),
               reference_top1_top2_similarity_gap=dict(median=float((top2[:, 0] - top2[:, 1]).median()),
                                                       min=float((top2[:, 0] - top2[:, 1]).min())),
               reference_autocast_bf16_vs_reference_fp32=dict(id_agreement=id_agreement(ref16, ref32)))
    for name, prec in modes().items():
        model.precision = prec
        ids = model(video, return_only_codebook_ids=True)
        flat, want = ids.reshape(-1), ref32.reshape(-1)
        differ = torch.nonzero(flat != want).flatten()
        worst = 0.0
        if differ.numel():  # how much worse (in cosine units) is the chosen code than the reference's, at worst
            rows = sims[differ]
            worst = float((rows.gather(1, want[differ, None]) - rows.gather(1, flat[differ, None])).max())
        for _ in range(3):
            model(video, return_only_codebook_ids=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            model(video, return_only_codebook_ids=True)
        e1.record()
        torch.cuda.synchronize()
        rep[f"ours_{name}"] = dict(id_agreement_vs_reference_fp32=id_agreement(ids, ref32),
                                   id_agreement_vs_reference_autocast_bf16=id_agreement(ids, ref16),
                                   worst_similarity_loss_of_a_differing_id=worst,
                                   encode_ms=e0.elapsed_time(e1) / 10,
                                   frames_per_s=batch * 17 / (e0.elapsed_time(e1) / 10) * 1e3)
    return rep


def _err(a, ref):
    e = (a.float() - ref.float()).abs()
    return dict(max_abs=float(e.max()), mean_abs=float(e.mean()), ref_rms=float(ref.float().pow(2).mean().sqrt()),
                argmax_agreement=float((a.argmax(-1) == ref.argmax(-1)).float().mean()))


def cfg3_logits_report(batch=4):
    _exact_fp32()
    torch.manual_seed(2)
    mg = P.MaskGit(**CFG3).eval()
    sd = {k: v.detach().clone().to(DEV) for k, v in mg.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 65537, (batch, 576), generator=g).to(DEV)
    ctx = torch.randn((batch, 16, 768), generator=g).to(DEV)
    tmask = torch.ones((batch, 16), dtype=torch.bool, device=DEV)
    tmask[1 % batch, 8:] = False
    kw = dict(video_patch_shape=(9, 8, 8), context=ctx, text_mask=tmask)
    with torch.no_grad():
        ref32 = O.maskgit_forward(ids, sd, **kw)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ref16 = O.maskgit_forward(ids, sd, **kw).float()
    mg = mg.to(DEV)
    rep = dict(batch=batch, reference_autocast_bf16_vs_reference_fp32=_err(ref16, ref32))
    for name, prec in modes().items():
        mg.precision = prec
        out = mg(ids, **kw)
        rep[f"ours_{name}_vs_reference_fp32"] = _err(out, ref32)
        rep[f"ours_{name}_vs_reference_autocast_bf16"] = _err(out, ref16)
        del out
    return rep


class _Tape:
    """Seeded uniform draws by tag, generated on the device: both sides of a comparison see the same numbers."""

    def __init__(self, seed):
        self.seed = seed

    def __call__(self, shape, tag):
        g = torch.Generator(device=DEV).manual_seed(self.seed * 1000 + sum(ord(c) * (i + 1) for i, c in enumerate(tag)))
        return torch.rand(shape, generator=g, device=DEV)


def cfg3_loop_report(batch=1, steps=18):
    """One full 18-step demasking loop at configs[2] sizes against the oracle on CUDA (fp32) with the uniform draws
    injected on both sides: the product's fp32 path through the unfused step, then the fused bf16 step at
    temperature 0 (no noise anywhere: both sides are pure arg-max decoders)."""
    _exact_fp32()
    torch.manual_seed(4)
    cv = P.CViViT(**CFG2)
    mg = P.MaskGit(**CFG3).eval()
    sd = {k: v.detach().clone().to(DEV) for k, v in mg.state_dict().items()}
    ctx = torch.randn((batch, 16, 768), generator=torch.Generator().manual_seed(5)).to(DEV)
    ph = P.Phenaki(cvivit=cv.to(DEV), maskgit=mg.to(DEV), steps=steps, text_embed_dim=768)
    rep = dict(batch=batch, steps=steps)

    def oracle(temp, trace, autocast=False):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            return O.sample_token_ids(sd, num_tokens=576, patch_shape=(9, 8, 8), batch=batch, steps=steps,
                                      text_embeds=ctx, cond_scale=3.0, starting_temperature=temp, noise_fn=_Tape(7),
                                      trace=trace)

    def ours(prec, temp, trace, noise):
        mg.precision = prec
        return ph.sample_token_ids(num_tokens=576, patch_shape=(9, 8, 8), batch_size=batch, text_embeds=ctx,
                                   cond_scale=3.0, starting_temperature=temp, noise_fn=_Tape(7) if noise else None,
                                   trace=trace)

    # (1) fp32 parity mode, reference temperature schedule, injected gumbel noise
    tr_ref, tr = [], []
    ref = oracle(0.9, tr_ref)
    out = ours(L.PREC_F32, 0.9, tr, True)
    first_bad = next((i for i, (a, b) in enumerate(zip(tr, tr_ref))
                      if not (torch.equal(a["mask"], b["mask"]) and torch.equal(a["ids"], b["ids"]))), None)
    rep["fp32_unfused_vs_reference_fp32_injected_noise"] = dict(
        final_id_agreement=id_agreement(out, ref), first_step_with_any_difference=first_bad,
        per_step_id_agreement=[id_agreement(a["ids"], b["ids"]) for a, b in zip(tr, tr_ref)])
    # (2) temperature 0: the fused bf16 step (production path) and the other modes against both reference dtypes
    tr32, tr16 = [], []
    ref32 = oracle(0.0, tr32)
    ref16 = oracle(0.0, tr16, autocast=True)
    rep["temperature0_reference_autocast_bf16_vs_reference_fp32"] = dict(
        final_id_agreement=id_agreement(ref16, ref32), step0_pred_agreement=id_agreement(tr16[0]["pred"], tr32[0]["pred"]))
    for name, prec in modes().items():
        t = []
        if prec == L.PREC_BF16:
            out = ours(prec, 0.0, None, False)   # fused head + masked-rows tail (no trace: the production path)
            t = None
        else:
            out = ours(prec, 0.0, t, True)
        d = dict(final_id_agreement_vs_reference_fp32=id_agreement(out, ref32),
                 final_id_agreement_vs_reference_autocast_bf16=id_agreement(out, ref16))
        if t:
            d["step0_pred_agreement_vs_reference_fp32"] = id_agreement(t[0]["pred"], tr32[0]["pred"])
        rep[f"temperature0_ours_{name}"] = d
    return rep


def main():
    rep = dict(device=torch.cuda.get_device_name(0), torch=torch.__version__)
    rep["cfg2_encode_ids"] = cfg2_report()
    torch.cuda.empty_cache()
    rep["cfg2_cosine_vq_K65536"] = cfg2_cosine_vq_report()
    torch.cuda.empty_cache()
    rep["cfg3_logits"] = cfg3_logits_report()
    torch.cuda.empty_cache()
    rep["cfg3_demask_loop"] = cfg3_loop_report()
    txt = json.dumps(rep, indent=1)
    print(txt)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
