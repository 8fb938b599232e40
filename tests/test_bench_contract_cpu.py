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


def test_committed_gpu_lines_carry_the_contract():
    """The GPU arm's lines kept under profiles/ (written by bench.py on B200 boxes this round): contract keys, the roofline
    arithmetic, parity of the timed result against the reference's f(), a launch count, clocks without a thermal / hardware flag."""
    full = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "parity", "sustained")
    for name, n in (("r02_bench_n1_final3.json", 1), ("r02_bench_n1_final4.json", 1), ("r02_bench_n2_final2.json", 2),
                    ("r02_bench_n4_final2.json", 4), ("r02_bench_n8_final2.json", 8)):
        d = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        for k in full:
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["dtype"] == "f64" and d["higher_is_better"] is True and d["vs_baseline"] is None
        pts = d["roofline"]["points_per_launch"]
        assert abs(d["value"] - n * pts / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
        r = d["roofline"]
        assert r["bound"] == "hbm" and abs(r["achieved"] - 32 * pts / (d["ms_per_step"] * 1e-3) / 1e9) <= 1e-6 * r["achieved"]
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.7 < r["frac"] < 1.2
        assert d["parity"]["bitwise"] is True and d["parity"]["rows_checked"] == 40 * n and d["parity"]["ranks"] == n
        assert d["gpu_launches"] == d["steps"]                       # one launch per evaluation, at every N
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert d["e2e"]["h2d_bytes_per_step"] == 16 * pts and d["e2e"]["matches_device_result"] is True
    d = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_n1_final3.json")).read().strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] > 0
    fp = [json.loads(open(os.path.join(ROOT, "profiles", f)).read().strip().splitlines()[-1])["integrator_fingerprint"]
          for f in ("r02_bench_n1_final4.json", "r02_bench_n2_final2.json", "r02_bench_n4_final2.json")]
    assert all(f["sum64"] == fp[0]["sum64"] and f["xor64"] == fp[0]["xor64"] and f["t"] == fp[0]["t"] for f in fp)
