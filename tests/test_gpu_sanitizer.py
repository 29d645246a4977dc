"""compute-sanitizer over a small exercise of every kernel (tools/sanitize_small.py: ragged sizes, all modes, a 4,096-env
chained run, zero-copy host inputs).  ONE tool per process and per GPU lease: the B200 profiling guide reports that running
several sanitizer tools in one call can leave the device unusable for everyone, so this test only runs when
FPV_RUN_SANITIZER names the tool (memcheck / racecheck / initcheck / synccheck); the builder runs each tool in its own
gpurun call and commits the summaries under profiles/ (profiles/r2_sanitizer_*.txt)."""
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
    assert r.returncode == 0 and "sanitize_small: ok" in r.stdout, tail
    assert "ERROR SUMMARY: 0 errors" in r.stdout + r.stderr, tail
