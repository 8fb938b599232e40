"""Scratch: kernel-variant sweep of the fused stage RHS f(t, sum c_j X_j), 16384 x 16384 FHN torus, EXACT, sustained."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
nx = ny = 16384
ctx = crd.Context(0)
g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, arith=0))
V = [g.new_vector() for _ in range(5)]
for j, v in enumerate(V):
    ctx.fill_synthetic("fhn_torus", 2 * nx * ny, v.device_ptr, seed=100 + j)
d = g.new_vector()
for n, c in ((2, [1.0, 0.01]), (3, [1.0, 0.01, 0.02]), (5, [1.0, 0.01, 0.02, 0.03, -0.01])):
    for variant in (0, 20):
        g.set_variant(variant)
        for _ in range(3): g.f_lincomb(50.0, c, V[:n], d)
        ctx.sync(); ctx.timer_start()
        reps = 40
        for _ in range(reps): g.f_lincomb(50.0, c, V[:n], d)
        ms = ctx.timer_stop() / reps
        byt = nx * ny * 16 * (n + 1)
        print(dict(n=n, variant=variant, ms=round(ms, 3), GBs=round(byt / ms / 1e6)), flush=True)
