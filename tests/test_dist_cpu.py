"""Host-side logic of the multi-process phi split on CPU: world_size-2 gloo ranks exchange the (here fake) halo
handles, pick ring neighbours, cover the mesh with SetupDecomp's formula, and run the allreduce hook."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from crdmodel_b200 import dist as cdist
    from crdmodel_b200.api import CRD_SUM, CRD_MAX, CRD_MIN
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    handles = cdist.exchange_handles(bytes([rank]) * 64)
    assert [h[0] for h in handles] == list(range(world)) and all(len(h) == 64 for h in handles)
    prev, nxt = cdist.ring_neighbours(rank, world)
    assert prev == (rank - 1) %% world and nxt == (rank + 1) %% world
    ext = cdist.slab_extents(1601, world)
    rows = [j for (a, b) in ext for j in range(a, b + 1)]
    assert rows == list(range(1601)) and ext[rank] == (1601 * rank // world, 1601 * (rank + 1) // world - 1)
    ar = cdist.make_allreduce()
    assert ar([1.0 + rank, 10.0], CRD_SUM) == [sum(1.0 + r for r in range(world)), 10.0 * world]
    assert ar([float(rank)], CRD_MAX) == [world - 1.0] and ar([float(rank)], CRD_MIN) == [0.0]
    dist.barrier()
    dist.destroy_process_group()
    print("DIST_OK", rank)
""") % ROOT


def test_two_rank_gloo_plumbing(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29623", str(script)],
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0 and r.stdout.count("DIST_OK") == 2, r.stdout[-2000:] + r.stderr[-2000:]
