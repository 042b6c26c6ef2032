"""CPU: the bf16-mode parity tests of tests/test_gpu_bf16_mode.py with the DRIVERS executed from the shipped source
(csrc/api.cu, sample_tail.cu and every plain kernel they launch) and the tcgen05 / TMA kernels represented by their
include/phk.h contracts (tests/cuda_emu/cuda_emu.cpp: bf16 operands, fp32 accumulation).  What this covers is the host
logic of bf16 mode -- bf16 weight packing (GEGLU row interleave, padding), buffer wiring, the CFG-pair sharing of the
first layer, the masked-rows tail of the demasking step -- not the tensor-core kernels themselves, which only the B200
run exercises."""
import pytest
import torch

from tests import emu_runtime
from tests import test_gpu_bf16_mode as G

_NAMES = ["test_cvivit_bf16_mode_against_fp32_reference_golden", "test_maskgit_bf16_mode_against_fp32_reference_golden",
          "test_sampling_runs_in_bf16_mode_and_is_deterministic", "test_layernorm_cfg_combination",
          "test_fused_sample_step_agrees_with_unfused_path", "test_bf16_sampling_with_fused_head_is_deterministic",
          "test_fused_sample_step_on_masked_rows_equals_the_all_rows_step"]
for _n in _NAMES:
    globals()[_n] = getattr(G, _n)


@pytest.fixture(scope="module")
def _emu_lib():
    return emu_runtime.build_emu()


@pytest.fixture(autouse=True)
def _product_on_the_cpu(_emu_lib, monkeypatch):
    emu_runtime.route_product_to_emulator(_emu_lib, monkeypatch)
    monkeypatch.setattr(G, "DEV", "cpu")
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "manual_seed_all", lambda *a, **k: None, raising=False)
