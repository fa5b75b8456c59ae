"""bench.py contract checks that need no GPU: the reference arm (the oracle port timed on the host cores) prints the
JSON line the driver parses, and the product arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--config", "2", "--batch", "4"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "G+D+GP step images/sec at 64x64" and line["unit"] == "images/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert abs(line["cpu_baseline"]["value"] - line["value"]) < 1e-9
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None
    # the reference arm runs on OUR arm's config: the config object is built by the same function in both arms
    assert line["config"]["batch_per_gpu"] == 4 and line["config"]["baseline_config"] == 2
    assert "batch 4/GPU" in line["config"]["workload"]


def test_default_workload_is_the_14_class_batch_128_config():
    sys.path.insert(0, ROOT)
    import bench
    old = sys.argv
    try:
        sys.argv = ["bench.py"]
        args = bench.parse()
    finally:
        sys.argv = old
    cfg, multiclass, B, H, W, label = bench.resolve(args)
    assert (cfg, multiclass, B, H, W) == (3, True, 128, 64, 128)
    c = bench.workload_config(cfg, label, B, 8)
    assert c["global_batch"] == 1024 and "14-class 64x64" in c["workload"]        # --gpus 8 is BASELINE configs[3]


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert "CUDA device" in (out.stderr + out.stdout)
