"""TEST INFRASTRUCTURE: a CPU restatement of the training-step kernels of csrc/train.cu, formula by formula and in the
driver's order (forward with saved activations, then the hand-derived backward -- NO autograd).  It exists so that the
backward MATH the CUDA kernels implement is checked on the CPU against the reference's autograd gradients
(tests/golden/train_*.pt) before any GPU time is spent: tests/test_train_mirror_cpu.py.

Each ``*_bwd`` function below corresponds to one kernel (or one GEMM call) of csrc/train.cu and uses the same inputs
(saved activations) and produces the same outputs; names match the kernels.
"""
import math

import torch
import torch.nn.functional as F

EPS_LN = 1e-5
EPS_L2 = 1e-12


# ---------------------------------------------------------------- forward pieces (same values the forward kernels save)
def ln_fwd(x, g, b):
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + EPS_LN) * g + b


def l2n(t):
    return t / t.norm(dim=-1, keepdim=True).clamp(min=EPS_L2)


def attn_fwd(q, kv, null_kv, q_scale, k_scale, bias, key_mask, heads):
    """q (b,n,I), kv (b,m,2I) -> o (b,n,I); returns saved (qh, kh, vv, P).  attention.py:146-181."""
    b, n, I = q.shape
    dh = I // heads
    m = kv.shape[1]
    nn_ = null_kv.shape[1] // 2
    qr = q.reshape(b, n, heads, dh).permute(0, 2, 1, 3)
    kr = kv[..., :I].reshape(b, m, heads, dh).permute(0, 2, 1, 3)
    vr = kv[..., I:].reshape(b, m, heads, dh).permute(0, 2, 1, 3)
    nk = null_kv[:, 0::2].unsqueeze(0).expand(b, -1, -1, -1)
    nv = null_kv[:, 1::2].unsqueeze(0).expand(b, -1, -1, -1)
    kraw = torch.cat((nk, kr), dim=2)
    vv = torch.cat((nv, vr), dim=2)
    qh = l2n(qr) * q_scale
    kh = l2n(kraw) * k_scale
    s = torch.einsum("bhid,bhjd->bhij", qh, kh) * 8.0
    if bias is not None:
        s = s + F.pad(bias, (nn_, 0))
    if key_mask is not None:
        km = F.pad(key_mask, (nn_, 0), value=True)
        s = s.masked_fill(~km[:, None, None, :], -torch.finfo(s.dtype).max)
    P = s.softmax(-1)
    o = torch.einsum("bhij,bhjd->bhid", P, vv).permute(0, 2, 1, 3).reshape(b, n, I)
    return o, (qr, kraw, qh, kh, vv, P)


def peg_fwd(x, w27, bias, shape, causal):
    """x (b, n, D) viewed as (b,t,h,w,D) (layout 0); w27 tap-major [27, D]; y = x + conv(x) + b."""
    b, t, h, w = shape
    D = x.shape[-1]
    xv = x.reshape(b, t, h, w, D)
    pad_t0 = 2 if causal else 1
    xp = F.pad(xv, (0, 0, 1, 1, 1, 1, pad_t0, 2 - pad_t0))
    y = xv + bias
    for kt in range(3):
        for kh in range(3):
            for kw in range(3):
                y = y + xp[:, kt:kt + t, kh:kh + h, kw:kw + w] * w27[(kt * 3 + kh) * 3 + kw]
    return y.reshape(b, t * h * w, D)


def cpb_coords(dims):
    """distinct coordinate deltas: rows u of the table, and the (n, n) -> u index map (cpb_table / cpb_expand kernels)."""
    d = list(dims) + [1] * (3 - len(dims))
    s = [2 * v - 1 for v in d]
    us = torch.arange(s[0] * s[1] * s[2])
    delta = torch.stack((us // (s[1] * s[2]) - (d[0] - 1), (us // s[2]) % s[1] - (d[1] - 1), us % s[2] - (d[2] - 1)), dim=-1)
    inp = torch.sign(delta) * torch.log(delta.abs().float() + 1)
    n = d[0] * d[1] * d[2]
    idx = torch.arange(n)
    c = torch.stack((idx // (d[1] * d[2]), (idx // d[2]) % d[1], idx % d[2]), dim=-1)
    diff = c[:, None, :] - c[None, :, :]
    umap = ((diff[..., 0] + d[0] - 1) * s[1] + (diff[..., 1] + d[1] - 1)) * s[2] + (diff[..., 2] + d[2] - 1)
    return inp[:, :len(dims)], umap


def lrelu(x):
    return torch.where(x > 0, x, 0.1 * x)


# ---------------------------------------------------------------- backward kernels
def ln_bwd(x, g, dy):
    """ln_bwd_dx_kernel + ln_bwd_dgb_kernel: -> dx, dgamma, dbeta."""
    D = x.shape[-1]
    mean = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(((x - mean) ** 2).mean(-1, keepdim=True) + EPS_LN)
    xhat = (x - mean) * rstd
    dxh = dy * g
    c1 = dxh.sum(-1, keepdim=True) / D
    c2 = (dxh * xhat).sum(-1, keepdim=True) / D
    dx = rstd * (dxh - c1 - xhat * c2)
    red = tuple(range(x.ndim - 1))
    return dx, (dy * xhat).sum(red), dy.sum(red)


def geglu_bwd(h, dg):
    """geglu_bwd_kernel: h = [val | gate] -> dh."""
    inner = h.shape[-1] // 2
    val, gate = h[..., :inner], h[..., inner:]
    cdf = 0.5 * (1 + torch.erf(gate * 0.7071067811865476))
    pdf = torch.exp(-0.5 * gate * gate) * 0.3989422804014327
    return torch.cat((dg * gate * cdf, dg * val * (cdf + gate * pdf)), dim=-1)


def attn_bwd(saved, do, q_scale, k_scale, heads, nnull):
    """attn_bwd_probs / attn_bwd_dq / attn_bwd_dkv / attn_bwd_norm kernels.
    -> dq (b,n,I), dkv (b,m,2I), dnull_kv (h, 2*nnull, dh), dq_scale, dk_scale, dS (b,h,n,nnull+m)."""
    qr, kraw, qh, kh, vv, P = saved
    b, H, n, dh = qr.shape
    dO = do.reshape(b, n, H, dh).permute(0, 2, 1, 3)
    dP = torch.einsum("bhid,bhjd->bhij", dO, vv)
    dS = P * (dP - (P * dP).sum(-1, keepdim=True))
    dqh = 8.0 * torch.einsum("bhij,bhjd->bhid", dS, kh)
    dkh = 8.0 * torch.einsum("bhij,bhid->bhjd", dS, qh)
    dvv = torch.einsum("bhij,bhid->bhjd", P, dO)

    def norm_bwd(raw, dhat, scale):
        r = raw.norm(dim=-1, keepdim=True).clamp(min=EPS_L2)
        u = raw / r
        dscale = (dhat * u).sum((0, 1, 2))
        gvec = dhat * scale
        return (gvec - u * (u * gvec).sum(-1, keepdim=True)) / r, dscale

    dq_raw, dq_scale = norm_bwd(qr, dqh, q_scale)
    dk_raw, dk_scale = norm_bwd(kraw, dkh, k_scale)
    I = H * dh
    dq = dq_raw.permute(0, 2, 1, 3).reshape(b, n, I)
    dk = dk_raw[:, :, nnull:].permute(0, 2, 1, 3).reshape(b, -1, I)
    dv = dvv[:, :, nnull:].permute(0, 2, 1, 3).reshape(b, -1, I)
    dnull = torch.zeros(H, 2 * nnull, dh)
    if nnull:
        dnull[:, 0::2] = dk_raw[:, :, :nnull].sum(0)
        dnull[:, 1::2] = dvv[:, :, :nnull].sum(0)
    return dq, torch.cat((dk, dv), dim=-1), dnull, dq_scale, dk_scale, dS


def peg_bwd(x, w27, dy, shape, causal):
    """peg_bwd_dx_kernel / peg_bwd_dw_kernel: -> dx, dw27, dbias."""
    b, t, h, w = shape
    D = x.shape[-1]
    pad_t0 = 2 if causal else 1
    dyv = dy.reshape(b, t, h, w, D)
    xv = x.reshape(b, t, h, w, D)
    # source p receives from output o = p - (kt - pad_t0, kh - 1, kw - 1)
    dyp = F.pad(dyv, (0, 0, 1, 1, 1, 1, 2 - pad_t0, pad_t0))
    xp = F.pad(xv, (0, 0, 1, 1, 1, 1, pad_t0, 2 - pad_t0))
    dx = dyv.clone()
    dw = torch.zeros(27, D)
    for kt in range(3):
        for kh in range(3):
            for kw in range(3):
                tap = (kt * 3 + kh) * 3 + kw
                dx = dx + dyp[:, 2 - kt:2 - kt + t, 2 - kh:2 - kh + h, 2 - kw:2 - kw + w] * w27[tap]
                dw[tap] = (dyv * xp[:, kt:kt + t, kh:kh + h, kw:kw + w]).sum((0, 1, 2, 3))
    return dx.reshape(b, t * h * w, D), dw, dyv.sum((0, 1, 2, 3))


def ce_fwd_bwd(logits, targets, token_mask):
    """ce_rows_kernel + reduce: mean CE over masked rows; dlogits = (softmax - onehot) / n_masked on masked rows."""
    R, V = logits.shape
    cnt = token_mask.sum().clamp(min=1).float()
    lse = torch.logsumexp(logits, dim=-1)
    row = lse - logits.gather(1, targets[:, None]).squeeze(1)
    loss = (row * token_mask).sum() / cnt
    d = torch.softmax(logits, -1)
    d[torch.arange(R), targets] -= 1.0
    return loss, d * (token_mask[:, None] / cnt)


def bce_fwd_bwd(scores, labels):
    """bce_rows_kernel: mean over all rows of softplus-form BCE with logits; dscore = (sigmoid - y) / R."""
    R = scores.numel()
    loss = (torch.clamp(scores, min=0) - scores * labels + torch.log1p(torch.exp(-scores.abs()))).mean()
    return loss, (torch.sigmoid(scores) - labels) / R


# ---------------------------------------------------------------- the driver (phk_maskgit_train_step), mirrored
def train_step(sd, ids_in, targets, token_mask, labels, *, patch_shape, heads, context, text_mask, is_critic,
               shrink_alpha=0.1, loss_scale=1.0, video_mask=None, head=None):
    """Returns (loss, grads dict keyed like the state dict, logits or None).  sd: plain state dict (no autograd).
    ``labels`` given -> Linear(dim,1) + BCE head (``head`` = (weight, bias) of SelfCritic.to_pred on a MaskGit body,
    its gradients come back as 'to_pred.weight' / 'to_pred.bias')."""
    b, n = ids_in.shape
    p = "transformer."
    D = sd["token_emb.weight"].shape[1]
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith(p + "layers."))
    shape = (b, *patch_shape)
    grads = {k: torch.zeros_like(v) for k, v in sd.items() if v.is_floating_point()}
    # ---- forward
    x = sd["pos_emb.weight"][:n] + sd["token_emb.weight"][ids_in]
    if not is_critic:
        x = x * shrink_alpha + x * (1 - shrink_alpha)
    bias, cpb_saved = None, None
    if not is_critic:
        inp, umap = cpb_coords(patch_shape)
        c = "continuous_pos_bias.net."
        a1 = lrelu(inp @ sd[c + "0.0.weight"].t() + sd[c + "0.0.bias"])
        a2 = lrelu(a1 @ sd[c + "1.0.weight"].t() + sd[c + "1.0.bias"])
        table = a2 @ sd[c + "2.weight"].t() + sd[c + "2.bias"]  # [U, heads]
        bias = table[umap].permute(2, 0, 1)  # (heads, n, n)
        cpb_saved = (inp, umap, a1, a2)
    saved = []
    for l in range(depth):
        lp = f"{p}layers.{l}."
        S = {"x0": x}
        w27 = sd[lp + "0.dsconv.weight"].reshape(D, 27).t()
        x1 = peg_fwd(x, w27, sd[lp + "0.dsconv.bias"], shape, False)
        S["x1"] = x1
        xn1 = ln_fwd(x1, sd[lp + "1.norm.gamma"], sd[lp + "1.norm.beta"])
        q = xn1 @ sd[lp + "1.to_q.weight"].t()
        kv = x1 @ sd[lp + "1.to_kv.weight"].t()  # raw x (attention.py:140-144)
        o, S["att1"] = attn_fwd(q, kv, sd[lp + "1.null_kv"], sd[lp + "1.q_scale"], sd[lp + "1.k_scale"], bias,
                                video_mask, heads)
        S.update(xn1=xn1, o1=o)
        x2 = x1 + o @ sd[lp + "1.to_out.weight"].t()
        S["x2"] = x2
        has_cross = (lp + "2.to_q.weight") in sd and context is not None
        if has_cross:
            ctxn = ln_fwd(context, sd[lp + "2.context_norm.gamma"], sd[lp + "2.context_norm.beta"])
            ckv = ctxn @ sd[lp + "2.to_kv.weight"].t()
            xn2 = ln_fwd(x2, sd[lp + "2.norm.gamma"], sd[lp + "2.norm.beta"])
            q2 = xn2 @ sd[lp + "2.to_q.weight"].t()
            o2, S["att2"] = attn_fwd(q2, ckv, sd[lp + "2.null_kv"], sd[lp + "2.q_scale"], sd[lp + "2.k_scale"], None,
                                     text_mask, heads)
            S.update(ctxn=ctxn, xn2=xn2, o2=o2)
            x3 = x2 + o2 @ sd[lp + "2.to_out.weight"].t()
        else:
            x3 = x2
        S["x3"] = x3
        xn3 = ln_fwd(x3, sd[lp + "3.0.weight"], sd[lp + "3.0.bias"])
        h = xn3 @ sd[lp + "3.1.weight"].t()
        inner = h.shape[-1] // 2
        g = F.gelu(h[..., inner:]) * h[..., :inner]
        S.update(xn3=xn3, h=h, g=g, has_cross=has_cross)
        x = x3 + g @ sd[lp + "3.4.weight"].t()
        saved.append(S)
    xf = x
    emb = ln_fwd(xf, sd[p + "norm_out.gamma"], sd[p + "norm_out.beta"])
    R = b * n
    logits = None
    if labels is not None:
        w, bb = head if head is not None else (sd["to_logits.0.weight"], sd["to_logits.0.bias"])
        hk = "to_pred." if head is not None else "to_logits.0."
        scores = (emb.reshape(R, D) @ w.t()).squeeze(-1) + bb
        loss, dsc = bce_fwd_bwd(scores, labels.reshape(R))
        dsc = dsc * loss_scale
        grads[hk + "weight"] = (dsc[:, None] * emb.reshape(R, D)).sum(0, keepdim=True)
        grads[hk + "bias"] = dsc.sum().reshape(1)
        demb = (dsc[:, None] * w).reshape(b, n, D)
    else:
        w, bb = sd["to_logits.weight"], sd["to_logits.bias"]
        logits = emb.reshape(R, D) @ w.t() + bb
        loss, dl = ce_fwd_bwd(logits, targets.reshape(R), token_mask.reshape(R).float())
        dl = dl * loss_scale
        grads["to_logits.weight"] = dl.t() @ emb.reshape(R, D)       # wgrad: dY^T X
        grads["to_logits.bias"] = dl.sum(0)
        demb = (dl @ w).reshape(b, n, D)                                # dgrad: dY W
    # ---- backward
    dx, dgm, _ = ln_bwd(xf, sd[p + "norm_out.gamma"], demb)
    grads[p + "norm_out.gamma"] = dgm
    dbias_total = torch.zeros_like(bias) if bias is not None else None
    for l in reversed(range(depth)):
        lp = f"{p}layers.{l}."
        S = saved[l]
        # feed-forward
        w1, w2 = sd[lp + "3.1.weight"], sd[lp + "3.4.weight"]
        dg = dx @ w2
        grads[lp + "3.4.weight"] = dx.reshape(R, D).t() @ S["g"].reshape(R, -1)
        dh = geglu_bwd(S["h"], dg)
        grads[lp + "3.1.weight"] = dh.reshape(R, -1).t() @ S["xn3"].reshape(R, D)
        dxn3 = dh @ w1
        d3, dgm, dbt = ln_bwd(S["x3"], sd[lp + "3.0.weight"], dxn3)
        grads[lp + "3.0.weight"], grads[lp + "3.0.bias"] = dgm, dbt
        dx = dx + d3
        # cross attention
        if S["has_cross"]:
            do2 = dx @ sd[lp + "2.to_out.weight"]
            grads[lp + "2.to_out.weight"] = dx.reshape(R, D).t() @ S["o2"].reshape(R, -1)
            nnull = sd[lp + "2.null_kv"].shape[1] // 2
            dq2, dckv, dnull, dqs, dks, _ = attn_bwd(S["att2"], do2, sd[lp + "2.q_scale"], sd[lp + "2.k_scale"], heads, nnull)
            grads[lp + "2.null_kv"], grads[lp + "2.q_scale"], grads[lp + "2.k_scale"] = dnull, dqs, dks
            grads[lp + "2.to_q.weight"] = dq2.reshape(R, -1).t() @ S["xn2"].reshape(R, D)
            dxn2 = dq2 @ sd[lp + "2.to_q.weight"]
            d2, dgm, _ = ln_bwd(S["x2"], sd[lp + "2.norm.gamma"], dxn2)
            grads[lp + "2.norm.gamma"] = dgm
            dx = dx + d2
            Lc = context.shape[1]
            grads[lp + "2.to_kv.weight"] = dckv.reshape(b * Lc, -1).t() @ S["ctxn"].reshape(b * Lc, -1)
            dctxn = dckv @ sd[lp + "2.to_kv.weight"]
            _, dgm, _ = ln_bwd(context, sd[lp + "2.context_norm.gamma"], dctxn)
            grads[lp + "2.context_norm.gamma"] = dgm
        # self attention
        do1 = dx @ sd[lp + "1.to_out.weight"]
        grads[lp + "1.to_out.weight"] = dx.reshape(R, D).t() @ S["o1"].reshape(R, -1)
        dq, dkv, dnull, dqs, dks, dS = attn_bwd(S["att1"], do1, sd[lp + "1.q_scale"], sd[lp + "1.k_scale"], heads, 0)
        grads[lp + "1.q_scale"], grads[lp + "1.k_scale"] = dqs, dks
        if dbias_total is not None:
            dbias_total += dS.sum(0)
        grads[lp + "1.to_q.weight"] = dq.reshape(R, -1).t() @ S["xn1"].reshape(R, D)
        grads[lp + "1.to_kv.weight"] = dkv.reshape(R, -1).t() @ S["x1"].reshape(R, D)
        dxn1 = dq @ sd[lp + "1.to_q.weight"]
        d1, dgm, _ = ln_bwd(S["x1"], sd[lp + "1.norm.gamma"], dxn1)
        grads[lp + "1.norm.gamma"] = dgm
        dx = dx + d1 + dkv @ sd[lp + "1.to_kv.weight"]
        # PEG
        w27 = sd[lp + "0.dsconv.weight"].reshape(D, 27).t()
        dx, dw27, dbp = peg_bwd(S["x0"], w27, dx, shape, False)
        grads[lp + "0.dsconv.weight"] = dw27.t().reshape(D, 1, 3, 3, 3)
        grads[lp + "0.dsconv.bias"] = dbp
    # ---- embeddings (gradient shrink: only the alpha branch carries gradient, phenaki_pytorch.py:199)
    a = 1.0 if is_critic else shrink_alpha
    dflat = (dx * a).reshape(R, D)
    grads["token_emb.weight"].index_add_(0, ids_in.reshape(R), dflat)
    grads["pos_emb.weight"][:n] += (dx * a).sum(0)
    # ---- continuous position bias MLP
    if bias is not None:
        inp, umap, a1, a2 = cpb_saved
        c = "continuous_pos_bias.net."
        U = inp.shape[0]
        dtable = torch.zeros(U, bias.shape[0])
        dtable.index_add_(0, umap.reshape(-1), dbias_total.permute(1, 2, 0).reshape(-1, bias.shape[0]))
        grads[c + "2.weight"] = dtable.t() @ a2
        grads[c + "2.bias"] = dtable.sum(0)
        da2 = (dtable @ sd[c + "2.weight"]) * torch.where(a2 > 0, 1.0, 0.1)
        grads[c + "1.0.weight"] = da2.t() @ a1
        grads[c + "1.0.bias"] = da2.sum(0)
        da1 = (da2 @ sd[c + "1.0.weight"]) * torch.where(a1 > 0, 1.0, 0.1)
        grads[c + "0.0.weight"] = da1.t() @ inp
        grads[c + "0.0.bias"] = da1.sum(0)
    return loss, grads, logits
