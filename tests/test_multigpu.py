"""Multi-GPU path (one process per GPU, IPC halo ring over NVLink).  Needs >= 2 GPUs; on a single-GPU box the
ring is covered by tests/test_rhs_gpu.py::test_phi_split_is_bitwise_invariant (emulated ranks, one process)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ring_across_gpus(crd, oracle):
    n = crd.lib().crd_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (found %d)" % n)
    world = 2 if n < 4 else (4 if n < 8 else 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "mgpu_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_driver_on_two_gpus_matches_one_gpu(crd, tmp_path):
    """bin/FHNmodel_torus with System.gpus = 2 (forked workers, IPC halo ring, shared-memory allreduce): the two
    subdomain file sets, stacked by their js/je, equal the single-GPU output digit for digit (System.resident = 0:
    both runs take the host-driven loop)."""
    import numpy as np
    if crd.lib().crd_device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    from crdmodel_b200 import build as B
    B.build_drivers()
    ini = """[Parameters]
diffusion = 0.12
beta = 1.25
surfaceWidth = 20
surfaceLength = 80
waveLength = 0.1
waveWidth = 0.5
waveInside = 1
outputTimestep = 3
tBoundary = 0.5
tFinal = 1.5
thetaMesh = 40
betaMin = 0.7
betaMax = 1.7

[System]
includeAllVars = 1
varyBeta = 0
gpus = %d
resident = 0
"""
    outs = {}
    for ng in (1, 2):
        d = tmp_path / ("g%d" % ng)
        d.mkdir()
        (d / "a.ini").write_text(ini % ng)
        r = subprocess.run([os.path.join(ROOT, "bin", "FHNmodel_torus"), str(d / "a.ini")], cwd=str(d), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert ("nprocs = %d" % ng) in r.stdout
        rows = []
        for rk in range(ng):
            sub = (d / ("FHNmodel_torus_subdomain.%03d.txt" % rk)).read_text().split()
            nx, ny, js, je = int(sub[0]), int(sub[1]), int(sub[4]), int(sub[5])
            u = np.loadtxt(d / ("FHNmodel_torus_u.%03d.txt" % rk)).reshape(4, je - js + 1, nx)
            rows.append((js, u))
        outs[ng] = np.concatenate([u for _, u in sorted(rows, key=lambda t: t[0])], axis=1)
        assert outs[ng].shape == (4, 160, 40)
    assert np.array_equal(outs[1], outs[2])      # EXACT arithmetic: the split does not change a bit of any output line
