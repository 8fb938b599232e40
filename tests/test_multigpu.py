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
