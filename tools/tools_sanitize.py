"""Scratch: one small pass over every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import crdmodel_b200 as crd
ctx = crd.Context(0)
nx, ny = 300, 77
for model in ("fhn_torus", "gb_flat"):
    for arith in (0, 1):
        for variant in (1, 5, 10, 13, 15, 20):
            g = crd.Grid(ctx, crd.make_params(model, nx, ny, arith=arith, t_boundary=38.0))
            g.set_variant(variant)
            V = [g.new_vector() for _ in range(5)]
            for j, v in enumerate(V):
                ctx.fill_synthetic(model, 2 * nx * ny, v.device_ptr, seed=7 + j)
            d = g.new_vector()
            g.f(10.0, V[0], d)
            for n in (2, 3, 5):
                g.f_lincomb(50.0, [1.0, 0.01, 0.02, 0.03, 0.04][:n], V[:n], d)
            ctx.sync()
            g.close()
# ring of 3 emulated ranks (thin slabs and overlapped path), N_Vector ops, fused ops, integrator step
for ny2 in (77, 600):
    grids, ys, ds = [], [], []
    for r in range(3):
        js, je = crd.decomp_phi(ny2, 3, r)
        g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny2, js=js, je=je))
        grids.append(g); ys.append(g.new_vector()); ds.append(g.new_vector()); g.fill_synthetic(ys[-1])
    for r in range(3):
        grids[r].halo_connect_local(grids[(r - 1) % 3], grids[(r + 1) % 3])
    for rep in range(2):
        for r in range(3): grids[r].post_halo(ys[r])
        for r in range(3): grids[r].compute(50.0, ys[r], ds[r])
    ctx.sync()
    for g in grids: g.close()
g = crd.Grid(ctx, crd.make_params("fhn_torus", 64, 96, vary_beta=0, t_boundary=0.0))
y = g.new_vector(); g.fill_initial_conditions(y, 0.1, 0.5, 1, -1.25, 1.25 ** 3 - 3.75)
s = crd.ARKodeSolver(g, y); print(s.ARKode(0.2)); s.free()
a, b, c = crd.NVector(ctx, 1001), crd.NVector(ctx, 1001), crd.NVector(ctx, 1001)
crd.N_VConst(1.5, a); crd.N_VConst(2.0, b); crd.N_VLinearSum(2.0, a, 3.0, b, c); crd.N_VInv(b, c)
print(crd.N_VWrmsNorm(a, b), crd.N_VMaxNorm(c), crd.N_VMin(c), crd.N_VDotProd(a, b), crd.N_VInvTest(a, c))
out = np.empty(2 * 64 * 96); g.f_host(1.0, y.to_numpy(), out)
print("sanitize pass done")
