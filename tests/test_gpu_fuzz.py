"""GPU (-m gpu): randomised parity sweep (tools/fuzz_gpu.py) -- random shapes (odd, tiny, wide), all 72 wavelets, every
transform mode, batches and proximal operators through the public class, every buffer bit-identical to the CPU oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12])
def test_random_shapes_and_modes_match_the_oracle(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_gpu.py"), "500", str(seed)], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    tail = "\n".join(l for l in r.stdout.splitlines() if not l.startswith(("Warning", "Forcing")))[-2000:]
    assert r.returncode == 0, tail + r.stderr[-2000:]
    assert "0 failures" in tail
