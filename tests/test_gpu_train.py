"""GPU: the training step (Phenaki.forward -> phk_maskgit_train_step: forward, loss and the hand-written backward
kernels of csrc/train.cu) against the reference's autograd loss and gradients (tests/golden/train_*.pt).

The kernels were written after round 1's GPU budget was spent: their math is pinned on the CPU
(tests/test_train_mirror_cpu.py) but they have not run on a GPU yet.  Until they have, each case runs in a child
process with a time limit (a fault cannot disturb the validated suite) and is a non-strict xfail: a pass shows up
as XPASS, a failure does not turn the suite red.
"""
import os
import subprocess
import sys

import pytest

from tests import cases as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="csrc/train.cu has not been run on a GPU yet (written after the GPU budget of "
                                        "round 1 was spent); math pinned on the CPU in test_train_mirror_cpu.py")
@pytest.mark.parametrize("name", list(C.TRAIN_CASES))
def test_training_step_matches_reference_autograd(name):
    env = dict(os.environ, PHK_EXPERIMENTAL="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_train_check.py"), name], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=180)
    sys.stdout.write(out.stdout[-4000:])
    sys.stderr.write(out.stderr[-4000:])
    assert out.returncode == 0 and f"TRAIN_OK {name}" in out.stdout
