"""Scratch: throughput of the fused stage-assembly RHS (f_lincomb) vs lincomb + f, 16384 x 16384 FHN torus."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
nx = ny = 16384
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ctx = crd.Context(0)
for arith in (0, 1):
    g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, arith=arith))
    g.set_variant(variant)
    V = [g.new_vector() for _ in range(5)]
    for j, v in enumerate(V):
        ctx.fill_synthetic("fhn_torus", 2 * nx * ny, v.device_ptr, seed=100 + j)
    z, d = g.new_vector(), g.new_vector()
    for n, c in ((1, [1.0]), (2, [1.0, 0.01]), (3, [1.0, 0.01, 0.02]), (5, [1.0, 0.01, 0.02, 0.03, -0.01])):
        for _ in range(3): g.f_lincomb(50.0, c, V[:n], d)
        ctx.sync(); ctx.timer_start()
        reps = 60
        for _ in range(reps): g.f_lincomb(50.0, c, V[:n], d)
        ms = ctx.timer_stop() / reps
        byt = nx * ny * 16 * (n + 1)
        for _ in range(3): crd.N_VLinearCombination(c, V[:n], z); g.f(50.0, z, d)
        ctx.sync(); ctx.timer_start()
        for _ in range(reps): crd.N_VLinearCombination(c, V[:n], z); g.f(50.0, z, d)
        ms2 = ctx.timer_stop() / reps
        print(dict(arith="exact" if arith == 0 else "fast", n=n, fused_ms=round(ms, 3), fused_GBs=round(byt / ms / 1e6), separate_ms=round(ms2, 3),
                   separate_GBs=round((byt + nx * ny * 32) / ms2 / 1e6)), flush=True)
    g.close()
