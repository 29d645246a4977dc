"""Every script under examples/ runs to completion on the GPU box (they are the documentation's runnable form)."""
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = sorted(glob.glob(os.path.join(ROOT, "examples", "*.py")))


@pytest.mark.gpu
@pytest.mark.parametrize("script", EXAMPLES, ids=[os.path.basename(p) for p in EXAMPLES])
def test_example_runs(script):
    r = subprocess.run([sys.executable, script], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.strip(), "the example printed nothing"
