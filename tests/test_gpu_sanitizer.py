"""compute-sanitizer over a small exercise of every kernel (tools/sanitize_small.py: ragged sizes, all modes, a 4,096-env
chained run, zero-copy host inputs).  ONE tool per process and per GPU lease: the B200 profiling guide reports that running
several sanitizer tools in one call can leave the device unusable for everyone, so this test only runs when
FPV_RUN_SANITIZER names the tool (memcheck / racecheck / initcheck / synccheck); the builder runs each tool in its own
gpurun call.  Round 2: the pool's wrapper answered "compute-sanitizer is closed on this pool and stays closed" (exit 86), so
the test skips there; what stands in for it are the guard-cell tests (sentinel padding around every state plane and output
of every kernel form at ragged sizes, tests/test_gpu_parity.py, test_gpu_acro.py, test_gpu_chase.py, test_gpu_ring_modes.py)
and the bit-identity tests between kernel forms."""
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_compute_sanitizer_is_clean():
    tool = os.environ.get("FPV_RUN_SANITIZER", "")
    if tool not in ("memcheck", "racecheck", "initcheck", "synccheck"):
        pytest.skip("set FPV_RUN_SANITIZER=memcheck|racecheck|initcheck|synccheck (one tool per GPU lease)")
    cs = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.isfile(cs):
        pytest.skip("compute-sanitizer not available")
    r = subprocess.run([cs, "--tool", tool, "--error-exitcode", "9", sys.executable, os.path.join(ROOT, "tools", "sanitize_small.py")],
                       capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    if r.returncode == 86 or "closed on this pool" in tail:
        pytest.skip("compute-sanitizer is closed on this GPU pool (the wrapper refuses to run it): " + tail.strip()[:160])
    assert r.returncode == 0 and "sanitize_small: ok" in r.stdout, tail
    assert "ERROR SUMMARY: 0 errors" in r.stdout + r.stderr, tail
