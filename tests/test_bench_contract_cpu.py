"""bench.py --impl reference runs on the CPU and prints one JSON line with the contract's keys (the GPU arm is
exercised on the GPU box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line(oracle):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] and d["dtype"] == "f64"


def test_reference_arm_other_ranks_exit_quietly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_stdout_is_exactly_the_json_line_even_when_libraries_write_to_fd_1():
    """NCCL prints its version banner on fd 1 when NCCL_DEBUG is set on the box: bench.py keeps a private duplicate of fd 1
    for the result and sends everything else to stderr."""
    import subprocess
    import sys
    code = ("import bench, os; bench.claim_stdout(); os.write(1, b'NCCL version 2.x\\n'); print('noise'); "
            "bench.emit({'metric': 'm', 'value': 1})")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0
    assert r.stdout == '{"metric": "m", "value": 1}\n'
    assert "NCCL version" in r.stderr and "noise" in r.stderr
