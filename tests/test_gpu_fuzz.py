"""B200: a short run of the randomised CUDA-vs-oracle check (scratch/gpu_fuzz.py) -- random joint counts,
odd heights, widths that are not multiples of 4, group sizes above and below the NMS pair-matrix limit.
Seed 1 is the sequence whose first (longer) run exposed the odd-height alignment bug."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_randomised_cuda_vs_oracle_short():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scratch", "gpu_fuzz.py"), "8", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    tail = (p.stdout + p.stderr)[-3000:]
    m = re.search(r"gpu fuzz: (.*) \| failures: (\d+)", p.stdout)
    assert p.returncode == 0 and m, tail
    assert int(m.group(2)) == 0, tail
    counts = dict((k, int(v)) for k, v in re.findall(r"(\w+) (\d+)", m.group(1)))
    assert counts.get("encode", 0) >= 10 and counts.get("decode", 0) >= 5, counts
