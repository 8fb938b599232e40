"""The drivers' rank processes (crdmodel_b200/host/crd_workers.hpp): ranks wait for each other on a shared barrier, so a rank
that dies must end the whole job instead of leaving the others waiting forever (the reference relies on mpirun for this)."""
import os
import subprocess
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("workers") / "workers_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "crdmodel_b200", "host"),
                    os.path.join(ROOT, "tests", "cpp", "workers_check.cpp"), "-o", out, "-lpthread"], check=True)
    return out


@pytest.mark.parametrize("mode,rc", [("ok", 0), ("fail", 1), ("crash", 1), ("rank0", 5)])
def test_a_failing_rank_ends_the_job(exe, mode, rc):
    t0 = time.time()
    r = subprocess.run([exe, mode, "4"], capture_output=True, text=True, timeout=20)
    assert r.returncode == rc, (mode, r.returncode, r.stderr)
    assert time.time() - t0 < 5.0
    if mode in ("fail", "crash"):
        assert "WORKER_ERROR" in r.stderr
    # no worker is left behind, blocked on the barrier
    time.sleep(0.2)
    left = subprocess.run(["pgrep", "-f", exe], capture_output=True, text=True).stdout.split()
    assert left == [], left


def test_multi_gpu_request_without_enough_gpus_fails_fast(tmp_path):
    """bin/FHNmodel_torus with System.gpus = 3 on a machine with fewer GPUs: every rank refuses; exit 1, no hang."""
    from crdmodel_b200 import build as B
    B.build(); B.build_drivers()
    try:
        import ctypes
        n = ctypes.CDLL(B.LIB).crd_device_count()
    except Exception:
        n = 0
    if n >= 3:
        pytest.skip("3 or more GPUs visible")
    ini = tmp_path / "a.ini"
    ini.write_text("[Parameters]\ndiffusion = 0.12\nbeta = 1.25\nbetaMin = 0.7\nbetaMax = 1.7\nsurfaceLength = 80\nsurfaceWidth = 20\n"
                   "waveLength = 0.1\nwaveWidth = 0.5\nwaveInside = 1\noutputTimestep = 2\ntBoundary = 1\ntFinal = 1\nthetaMesh = 32\n"
                   "[System]\nincludeAllVars = 0\nvaryBeta = 0\ngpus = 3\n")
    r = subprocess.run([os.path.join(ROOT, "bin", "FHNmodel_torus"), str(ini)], capture_output=True, text=True, timeout=30, cwd=tmp_path)
    assert r.returncode == 1
