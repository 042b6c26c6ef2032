"""TEST INFRASTRUCTURE: builds tests/cuda_emu (csrc/train.cu compiled by g++ for the CPU executor) and routes the
product's host path to it with CPU tensors.  Used by test_train_emulated_cpu.py and the two-rank gloo test."""
import contextlib
import ctypes
import os
import subprocess

import torch

from phenaki_pytorch_b200 import _lib as L
from phenaki_pytorch_b200 import modules as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "cuda_emu")
EMU_LIB = os.path.join(EMU_DIR, "_build", "libphk_train_emu.so")
SOURCES = [os.path.join(ROOT, "phenaki_pytorch_b200", "csrc", "train.cu"), os.path.join(EMU_DIR, "cuda_emu.cpp")]


def build_emu():
    deps = SOURCES + [os.path.join(EMU_DIR, "cuda_emu.h"), os.path.join(ROOT, "include", "phk.h")]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
        tmp = EMU_LIB + f".{os.getpid()}.tmp"
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-DPHK_CUDA_EMU", "-x", "c++", *SOURCES,
                               "-o", tmp])
        os.replace(tmp, EMU_LIB)
    lib = ctypes.CDLL(EMU_LIB)
    for name in ("phk_maskgit_train_workspace_bytes", "phk_maskgit_train_step"):
        fn = getattr(lib, name)
        fn.argtypes = L.PROTOTYPES[name]
        fn.restype = L._RESTYPES.get(name, ctypes.c_int)
    lib.phk_last_error.restype = ctypes.c_char_p
    lib.phk_emu_set_shuffle.argtypes = [ctypes.c_uint64]
    lib.phk_emu_set_shuffle.restype = None
    return lib


class _Setter:
    """monkeypatch-shaped setter for processes that never need to undo (spawned workers)."""

    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)


def route_product_to_emulator(lib, patch=_Setter):
    """The product refuses CPU tensors by design; under test the train-step entry points are served by the emulated
    library instead of libphk.so and the CUDA-only plumbing (device guard, stream) becomes a no-op."""
    patch.setattr(L, "lib", lambda: lib)
    patch.setattr(L, "require_cuda", lambda t, name, dtype=None: t.contiguous())
    patch.setattr(L, "stream_ptr", lambda: None)
    patch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())

    def keep_t(self, tensor):
        tensor = tensor.detach().float().contiguous()
        self.refs.append(tensor)
        return tensor.data_ptr()

    patch.setattr(M.Keep, "t", keep_t)

    def keep_h(self, tensor):
        t = tensor.detach().to(torch.bfloat16).contiguous()
        self.refs.append(t)
        return t.data_ptr()

    patch.setattr(M.Keep, "h", keep_h)
