"""bench.py's reference arm on the CPU: the JSON line carries the keys the driver's contract names (the GPU arm cannot run here)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-1500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_line():
    j = _line("--impl", "reference", "--config", "tiny", "--steps", "2", "--warmup", "1", "--cpu-steps", "1")
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["impl"] == "reference" and j["metric"] == "bpr_triples_per_sec_hot_step" and j["unit"] == "triples/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and "workload" in j["config"] and j["vs_baseline"] is None
    cb, e2e = j["cpu_baseline"], j["e2e"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    assert e2e == {"value": j["value"], "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_stock_torch_comparator_line_on_cpu_device():
    env_backup = os.environ.get("MMSSL_STOCK_DEVICE")
    os.environ["MMSSL_STOCK_DEVICE"] = "cpu"
    try:
        j = _line("--impl", "stock-gpu", "--config", "tiny", "--steps", "2", "--warmup", "1")
    finally:
        if env_backup is None:
            os.environ.pop("MMSSL_STOCK_DEVICE", None)
        else:
            os.environ["MMSSL_STOCK_DEVICE"] = env_backup
    assert j["impl"] == "stock-torch-gpu" and j["value"] > 0 and j["comparator"]["ms_per_step"] > 0
