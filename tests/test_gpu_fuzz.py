"""Short randomised differential run (tools/fuzz_gpu.py): random sizes / degrees / option settings / interleavings,
everything compared bit for bit with the oracle. Longer runs: `python tools/fuzz_gpu.py 90 <seed>`."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz_against_oracle(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_gpu.py"), "6", str(seed)], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and "fuzz ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
    assert "starved 0" in r.stdout
