"""CPU: the measurement contract that can be checked without a GPU -- the reference arm
(`bench.py --impl reference`, the oracle port timed on the host cores) prints ONE JSON line with the
contract's keys, and the product arm refuses to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "persons/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("persons/sec encode+decode") and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "persons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    # both arms print the same config object (same workload, same sizes, derived from the same flags)
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    want = bench.bench_config(argparse.Namespace(height=64, width=48, batch=1024, persons=8192, steps=1), 1)
    assert d["config"] == want and want["persons_per_gpu_per_step"] == 8192 * bench.step_cycles(1)
    assert bench.step_cycles(20) == 25 and bench.step_cycles(200) == 3 and bench.step_cycles(10000) == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_arm_needs_cuda():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
    assert not [ln for ln in p.stdout.splitlines() if ln.startswith("{")]


def test_source_fingerprint_does_not_depend_on_where_the_tree_lives(tmp_path, monkeypatch):
    """bench.py trusts profiles/*_traffic.json only when its fingerprint equals the one of the sources in use; the GPU box
    runs from a scratch path, so the hash must cover relative names and contents, not absolute paths."""
    import shutil
    from simple_pose_b200 import build
    here = build._fingerprint()
    root = tmp_path / "elsewhere"
    shutil.copytree(build.CSRC, root / "simple_pose_b200" / "csrc")
    shutil.copytree(os.path.join(os.path.dirname(build.PKG), "include"), root / "include")
    monkeypatch.setattr(build, "PKG", str(root / "simple_pose_b200"))
    monkeypatch.setattr(build, "CSRC", str(root / "simple_pose_b200" / "csrc"))
    assert build._fingerprint() == here
    with open(root / "simple_pose_b200" / "csrc" / "sp_step.cu", "a") as fh:
        fh.write("// changed\n")
    assert build._fingerprint() != here
