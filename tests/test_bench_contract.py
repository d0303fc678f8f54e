"""bench.py's reference arm runs on the CPU, so its half of the driver contract can be checked without a GPU: exactly one
JSON line on stdout, the keys the driver reads, and the synthetic image generators."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--batch", "4096", "--steps", "1",
                          "--warmup", "3"], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "images/s" and d["value"] > 0
    assert d["metric"].startswith("MNIST-16x16 images/sec") and d["config"]["workload"].startswith("fc:")
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--batch", "4096",
                          "--steps", "1"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_synthetic_image_generators():
    sys.path.insert(0, ROOT)
    import bench
    u = bench.synth_images(1000, 256, 7, "uniform")
    m = bench.synth_images(1000, 256, 7, "mnist")
    assert u.shape == m.shape == (1000, 256) and u.dtype == m.dtype == np.int8
    assert np.array_equal(u, bench.synth_images(1000, 256, 7, "uniform"))           # seeded
    assert abs(int(np.median(m)) + 20) <= 2 and m.max() == 127 and u.min() == -128   # background near -20, strokes to 127
