"""Worker for tests/test_multigpu.py: run under torchrun, one rank per GPU.
Checks (1) the phi-split RHS over the IPC halo ring is bit-identical to the single-slab CPU checker,
(2) a multi-rank integration ends on the bits of the single-rank CPU integration."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crdmodel_b200 as crd  # noqa: E402
from crdmodel_b200 import dist as cdist  # noqa: E402
import oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ctx = crd.Context(local)
    ctx.set_comm(rank, world, cdist.make_allreduce())
    if os.environ.get("CRD_TEST_HOST_ALLREDUCE") != "1":
        cdist.comm_connect(ctx, rank, world)       # device-side allreduce (default); the host hook stays as the fallback
    ok = True
    for model in ("fhn_torus", "gb_torus", "fhn_flat"):
        nx, ny = 320, 1003
        y = O.fill_state(model, 2 * nx * ny, seed=17)
        js, je = crd.decomp_phi(ny, world, rank)
        for arith in (crd.ARITH_EXACT,):
            grid = crd.Grid(ctx, crd.make_params(model, nx, ny, js=js, je=je, arith=arith, t_boundary=38.0))
            cdist.ring_connect(grid, rank, world, cdist.exchange_handles(grid.halo_handle()))
            yv = crd.NVector.from_numpy(ctx, y[2 * nx * js: 2 * nx * (je + 1)], 2 * nx * ny)
            dv = grid.new_vector()
            for t in (10.0, 50.0, 10.0, 50.0, 50.0):      # several epochs through the double-buffered ghosts
                grid.f(t, yv, dv)
            got = dv.to_numpy()
            ref = O.rhs(O.make_params(model, nx, ny, t_boundary=38.0), 50.0, y)[2 * nx * js: 2 * nx * (je + 1)]
            if model.startswith("fhn"):
                good = got.tobytes() == ref.tobytes()
            else:
                good = bool(np.abs(got - ref).max() <= 1e-12 * (1 + np.abs(ref).max()))
            # fused stage assembly over the ring: f(t, sum c_j X_j) == lincomb then f, bit for bit
            X2 = O.fill_state(model, 2 * nx * ny, seed=18)
            xv = crd.NVector.from_numpy(ctx, X2[2 * nx * js: 2 * nx * (je + 1)], 2 * nx * ny)
            zv, d1, d2 = grid.new_vector(), grid.new_vector(), grid.new_vector()
            sv = grid.new_vector()                       # EXACT grid: the op-by-op stage assembly of the RK driver
            crd.N_VConst(0.0, sv); crd.N_VLinearSum(0.01, xv, 1.0, sv, sv); crd.N_VLinearSum(1.0, yv, 1.0, sv, zv)
            grid.f(50.0, zv, d1)
            grid.f_lincomb(50.0, [1.0, 0.01], [yv, xv], d2)
            good = good and d1.to_numpy().tobytes() == d2.to_numpy().tobytes()
            # host-buffer entry over the ring
            out = np.empty_like(got)
            grid.f_host(50.0, np.ascontiguousarray(y[2 * nx * js: 2 * nx * (je + 1)]), out)
            good = good and out.tobytes() == got.tobytes()
            # global reductions
            import math
            nrm = crd.N_VWrmsNorm(yv, yv)
            good = good and nrm == math.sqrt(math.fsum((y * y) ** 2) / y.size)     # exactly rounded, whatever the split
            good = good and crd.N_VMaxNorm(yv) == np.abs(y).max() and crd.N_VMin(yv) == y.min()
            if not good:
                print("rank %d: FAILED %s" % (rank, model), flush=True)
            ok = ok and good
            dist.barrier()
            ctx.sync()
            grid.close()
    # last RK stage fused with the step finish over the ring (interior + two edge-band launches, each with its own region
    # of partial sums) == stage evaluation then N_VErkFinish: ynew bit for bit, the global sums to rounding
    for model in ("fhn_torus", "gb_torus"):
        nx, rows = 320, 3400                       # > 1 Mi points per rank: the streaming kernel applies
        ny = rows * world
        js, je = crd.decomp_phi(ny, world, rank)
        grid = crd.Grid(ctx, crd.make_params(model, nx, ny, js=js, je=je, t_boundary=38.0))
        cdist.ring_connect(grid, rank, world, cdist.exchange_handles(grid.halo_handle()))
        X = []
        for k in range(5):
            v = grid.new_vector()
            grid.fill_synthetic(v, seed=0x5EED + k)
            if k > 0:
                crd.N_VScale(0.25, v, v)
            X.append(v)
        h = 1e-3
        c = [1.0, h * 5 / 32, h * 7 / 32, h * 13 / 32, -h / 32]
        hb = [h / 6, h / 3, h / 3, h / 6, 0.0]
        hd = [h * (1 / 6 + 0.5), h * (1 / 3 - 7 / 3), h * (1 / 3 - 7 / 3), h * (1 / 6 - 13 / 6), h * 16 / 3]
        good = True
        for t in (10.0, 50.0):
            F5, want, got = grid.new_vector(), grid.new_vector(), grid.new_vector()
            grid.f_lincomb(t, c, X, F5)
            e2, y2 = crd.N_VErkFinish(hb, hd, X[0], X[1:] + [F5], want, 1e-5, 1e-10, exact=True)
            rc, fe2, fy2 = grid.f_lincomb_finish(t, c, hb, hd, X, got, 1e-5, 1e-10)
            good = good and rc == 0 and got.to_numpy().tobytes() == want.to_numpy().tobytes()
            good = good and fe2 == e2 and abs(fy2 - y2) <= 1e-11 * y2
        if not good:
            print("rank %d: FAILED fused stage finish %s" % (rank, model), flush=True)
        ok = ok and good
        dist.barrier()
        ctx.sync()
        grid.close()
    # trajectory: phi-split integration == single-slab CPU integration of the same driver
    nx, ny = 32, 128
    beta = 1.25
    js, je = crd.decomp_phi(ny, world, rank)
    grid = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, js=js, je=je, beta=beta, vary_beta=0, t_boundary=1.0))
    cdist.ring_connect(grid, rank, world, cdist.exchange_handles(grid.halo_handle()))
    yv = grid.new_vector()
    grid.fill_initial_conditions(yv, 0.1, 0.5, 1, -beta, beta ** 3 - 3 * beta)
    solver = crd.ARKodeSolver(grid, yv)        # EXACT grid, fused operations: the bits of the op-by-op sequence
    flag, t = solver.ARKode(2.0)
    st = solver.stats()
    mine = yv.to_numpy()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    full = np.concatenate(gathered)
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from test_integrate_gpu import cpu_trajectory, reference_ics
        y0, _ = reference_ics("fhn_torus", nx, ny, beta)
        cpu, nst_c, nfe_c = cpu_trajectory(O, O.make_params("fhn_torus", nx, ny, beta=beta, vary_beta=0, t_boundary=1.0), y0, [2.0], 1e-5, 1e-10)
        # same RHS bits, same element-wise bits, an error norm that does not depend on the summation order: the phi-split
        # GPU run takes the steps of the single-rank CPU run and ends on the same bits
        good = flag == 0 and full.tobytes() == cpu[0].tobytes() and st["nst"] == nst_c
        print("trajectory: world=%d nst=%d nfe=%d (cpu nst=%d) bitwise=%s" % (world, st["nst"], st["nfe"], nst_c, good), flush=True)
        ok = ok and good
    solver.free(); grid.close()
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        print("MGPU_OK" if all(flags) else "MGPU_FAIL", flush=True)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
