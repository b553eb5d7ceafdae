"""compute-sanitizer over the kernels that synchronise through shared memory, mbarriers, TMEM or cross-GPU flag
words (racecheck was missing): the tcgen05 GEMM with its stream-K fix-up, the SFU field kernel
(TMA-staged chunks, gather mode, ballot compaction), the 4D-STEM kernels (shared-memory CAS atomics), the ray
kernel's bulk store and the peer barrier.  tests/sanitizer_subset.py holds the invocations (each checks its own
result against the oracle); this test runs it under racecheck and synccheck and requires a clean summary."""
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sanitizer():
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    return exe if os.path.exists(exe) else None


@pytest.mark.parametrize("tool,clean", [("racecheck", "RACECHECK SUMMARY: 0 hazards displayed (0 errors, 0 warnings)"),
                                        ("synccheck", "ERROR SUMMARY: 0 errors")])
def test_sanitizer_subset(tool, clean):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = _sanitizer()
    if exe is None:
        pytest.skip("compute-sanitizer not installed")
    r = subprocess.run([exe, "--tool", tool, sys.executable, os.path.join(ROOT, "tests", "sanitizer_subset.py"), "all"],
                       cwd=ROOT, capture_output=True, text=True, timeout=1200)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    for sec in ("gemm", "field", "binned", "stem4d", "trace", "jets", "peer"):
        assert f"section {sec}: ok" in out, out[-4000:]
    assert clean in out, out[-4000:]
