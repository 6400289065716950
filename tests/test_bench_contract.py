"""CPU test of bench.py's contract for the arm that can run without a GPU: `--impl reference` (the CPU port
of the path, oracle/).  One JSON line with the keys the driver parses; values self-consistent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys(oracle_lib):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "geodesic maps/sec" and d["unit"] == "maps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["data"] == "synthetic"
    assert d["dtype"] == "f32" and d["vs_baseline"] is None and d["scaling"] == "weak"
    cfg = d["config"]
    assert cfg["N"] == 50000 and cfg["Q"] == 128 and cfg["k"] == 8 and cfg["max_step"] == 32 and "workload" in cfg
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == "maps/s"
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # value = maps per second of one scene's worth of seeds
    assert abs(d["value"] - cfg["Q"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]


def test_bench_refuses_to_run_the_product_arm_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--no-e2e",
                          "--no-cpu-baseline"], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
