"""CPU: csrc/train.cu itself -- the CUDA source of the training step -- compiled by g++ against tests/cuda_emu (every
CUDA thread a fiber, __syncthreads / warp shuffles as barriers) and EXECUTED on the CPU through the product's own
Python path (MaskGit.train_step -> ctypes tables -> phk_maskgit_train_step), against the reference's autograd loss and
gradients (tests/golden/train_*.pt).  The forward building blocks the driver calls are CPU statements of the
include/phk.h contracts here (their CUDA versions are covered by the -m gpu suite); everything else -- the driver's
buffer wiring, the gradient table, every backward kernel's indexing and synchronisation -- is the shipped code."""
import contextlib
import ctypes
import os
import subprocess

import pytest
import torch

import phenaki_pytorch_b200 as P
from phenaki_pytorch_b200 import _lib as L
from phenaki_pytorch_b200 import modules as M
from tests import cases as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "cuda_emu")
EMU_LIB = os.path.join(EMU_DIR, "_build", "libphk_train_emu.so")
SOURCES = [os.path.join(ROOT, "phenaki_pytorch_b200", "csrc", "train.cu"), os.path.join(EMU_DIR, "cuda_emu.cpp")]


@pytest.fixture(scope="module")
def emu():
    deps = SOURCES + [os.path.join(EMU_DIR, "cuda_emu.h"), os.path.join(ROOT, "include", "phk.h")]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-DPHK_CUDA_EMU", "-x", "c++", *SOURCES,
                               "-o", EMU_LIB])
    lib = ctypes.CDLL(EMU_LIB)
    for name in ("phk_maskgit_train_workspace_bytes", "phk_maskgit_train_step"):
        fn = getattr(lib, name)
        fn.argtypes = L.PROTOTYPES[name]
        fn.restype = L._RESTYPES.get(name, ctypes.c_int)
    lib.phk_last_error.restype = ctypes.c_char_p
    return lib


@pytest.fixture
def on_cpu(emu, monkeypatch):
    """Routes the product's host path to the emulated library with CPU tensors (test only)."""
    monkeypatch.setattr(L, "lib", lambda: emu)
    monkeypatch.setattr(L, "require_cuda", lambda t, name, dtype=None: t.contiguous())
    monkeypatch.setattr(L, "stream_ptr", lambda: None)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())

    def keep_t(self, tensor):
        tensor = tensor.detach().float().contiguous()
        self.refs.append(tensor)
        return tensor.data_ptr()

    monkeypatch.setattr(M.Keep, "t", keep_t)


def _modules(case):
    torch.manual_seed(case["seed"])
    P.CViViT(**C.SAMPLE_CVIVIT)
    maskgit = P.MaskGit(**case["maskgit"])
    critic = P.TokenCritic(**case["critic"]) if case["critic"] else None
    return maskgit, critic


def _compare(module, gk, ref, who):
    for k, p in module.named_parameters():
        got = gk.grad_of(p)
        if k not in ref:
            assert got is None, f"{who}.{k}"
            continue
        if ref[k].numel() == 0:
            assert got.shape == ref[k].shape
            continue
        scale = max(ref[k].abs().max().item(), 1e-12)
        torch.testing.assert_close(got, ref[k], rtol=1e-3, atol=1e-4 * scale + 1e-7, msg=lambda m, k=k: f"{who}.{k}: {m}")


@pytest.mark.parametrize("name", list(C.TRAIN_CASES))
def test_cuda_training_step_executed_on_the_cpu_matches_reference_autograd(golden, on_cpu, name):
    case, g = C.TRAIN_CASES[name], golden(f"train_{name}")
    maskgit, critic = _modules(case)
    ids, ctx = C.train_inputs(case)
    b, n = ids.shape[0], ids[0].numel()
    flat = ids.reshape(b, n)
    token_mask = g["token_mask"]
    tmask = torch.any(ctx != 0, dim=-1)
    vmask = torch.ones((b, n), dtype=torch.bool)
    mask_id = case["maskgit"]["num_tokens"]
    loss, gk, logits = maskgit.train_step(torch.where(token_mask, mask_id, flat), case["patch_shape"], targets=flat,
                                          token_mask=token_mask, context=ctx, text_mask=tmask, video_mask=vmask,
                                          keep_logits=critic is not None)
    ref_loss = g["ce"] if critic is not None else g["loss"]
    torch.testing.assert_close(loss, ref_loss, rtol=1e-5, atol=1e-6)
    _compare(maskgit, gk, g["maskgit_grads"], "maskgit")
    if critic is not None:
        # the logits handed back for the critic's sampling are the forward logits, not their gradient
        from oracle import phenaki_oracle as O
        with torch.no_grad():
            ref_logits = O.maskgit_forward(torch.where(token_mask, mask_id, flat), maskgit.state_dict(),
                                           video_patch_shape=case["patch_shape"], heads=case["maskgit"]["heads"],
                                           context=ctx, text_mask=tmask, video_mask=vmask)
        torch.testing.assert_close(logits, ref_logits, rtol=1e-4, atol=1e-4)
        pred = g["pred_ids"]
        closs, cgk, _ = critic.train_step(torch.where(token_mask, pred, flat), case["patch_shape"],
                                          labels=(flat != pred).float(), context=ctx, text_mask=tmask, video_mask=vmask,
                                          loss_scale=1.0)
        torch.testing.assert_close(closs, g["bce"], rtol=1e-5, atol=1e-6)
        _compare(critic, cgk, g["critic_grads"], "critic")


def test_emulated_step_with_a_gradient_scale_and_without_text(golden, on_cpu):
    """loss_scale scales every gradient but not the reported loss; without context the cross-attention is skipped
    (attention.py:327) and its parameters get no gradient."""
    case = C.TRAIN_CASES["with_critic"]
    g = golden("train_with_critic")
    maskgit, _ = _modules(case)
    ids, ctx = C.train_inputs(case)
    b, n = ids.shape[0], ids[0].numel()
    flat, token_mask = ids.reshape(b, n), g["token_mask"]
    inp = torch.where(token_mask, case["maskgit"]["num_tokens"], flat)
    kw = dict(targets=flat, token_mask=token_mask, context=ctx, text_mask=torch.any(ctx != 0, dim=-1))
    l1, g1, _ = maskgit.train_step(inp, case["patch_shape"], **kw)
    l2, g2, _ = maskgit.train_step(inp, case["patch_shape"], loss_scale=0.25, **kw)
    torch.testing.assert_close(l1, l2, rtol=0, atol=0)
    torch.testing.assert_close(g2.flat, 0.25 * g1.flat, rtol=1e-5, atol=1e-9)
    l3, g3, _ = maskgit.train_step(inp, case["patch_shape"], targets=flat, token_mask=token_mask)
    cross = maskgit.transformer.layers[0][2]
    assert g3.grad_of(cross.to_q.weight) is None and g3.grad_of(maskgit.to_logits.weight) is not None
    from oracle import phenaki_oracle as O
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in maskgit.state_dict().items()}
    ref = O.maskgit_train_loss(flat, sd, token_mask, video_patch_shape=case["patch_shape"], heads=case["maskgit"]["heads"])
    ref.backward()
    torch.testing.assert_close(l3, ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(g3.grad_of(maskgit.to_logits.weight), sd["to_logits.weight"].grad, rtol=1e-3, atol=1e-7)
    torch.testing.assert_close(g3.grad_of(maskgit.token_emb.weight), sd["token_emb.weight"].grad, rtol=1e-3, atol=1e-7)
