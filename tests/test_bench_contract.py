"""bench.py's output contract on the CPU side: the reference arm prints exactly one JSON line on stdout
with the keys the driver reads, whatever libraries write to file descriptor 1 meanwhile."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"}


def run_reference(*extra):
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "4", *extra], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line(native):
    line = run_reference()
    assert REQUIRED <= set(line)
    assert line["impl"] == "reference" and line["unit"] == "solves/s" and line["higher_is_better"] is True
    assert line["metric"] == "batched OCP SQP solves/sec (H=20)" and "workload" in line["config"]
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1


def test_reference_arm_other_ranks_stay_silent(native, monkeypatch):
    monkeypatch.setenv("RANK", "1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_workload_switch_changes_metric_and_config():
    import importlib
    sys.path.insert(0, str(ROOT))
    bench = importlib.import_module("bench")

    class Args:
        workload, batch = "cartpole", 0
    try:
        bench.select_workload(Args)
        assert bench.METRIC == "batched OCP SQP solves/sec (H=200)" and Args.batch == 592
        assert "cart-pole" in bench.workload_config(Args.batch, 1)["workload"]
    finally:
        Args.workload, Args.batch = "quadrotor", 0
        bench.select_workload(Args)
    assert bench.METRIC == "batched OCP SQP solves/sec (H=20)" and Args.batch == 4096
