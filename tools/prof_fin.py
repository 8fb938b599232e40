"""Round-2 driver for the last stage + finish kernel (EXACT / FAST): times it per grid variant (0 = default; 23 / 24 need a
library built with --profiling-variants) and checks ynew against the separate stage evaluation + N_VErkFinish.
python tools/prof_fin.py [rows] [reps] [variants, comma separated]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
ctx = crd.Context(0)
h = 1e-3
c5 = [1.0, h * 5 / 32, h * 7 / 32, h * 13 / 32, -h / 32]
hb = [h / 6, h / 3, h / 3, h / 6, 0.0]
hd = [h * (1 / 6 + 0.5), h * (1 / 3 - 7 / 3), h * (1 / 3 - 7 / 3), h * (1 / 6 - 13 / 6), h * 16 / 3]
for model, nx in (("fhn_torus", 16384), ("gb_torus", 8192)):
  for arith in (crd.ARITH_EXACT, crd.ARITH_FAST):
    g = crd.Grid(ctx, crd.make_params(model, nx, rows, arith=arith))
    X = [g.new_vector() for _ in range(5)]
    out, want, F5 = g.new_vector(), g.new_vector(), g.new_vector()
    for j, v in enumerate(X):
        g.fill_synthetic(v, seed=0x5EED + j)
        if j:
            crd.N_VScale(0.25, v, v)
    g.f_lincomb(50.0, c5, X, F5)
    e2, _ = crd.N_VErkFinish(hb, hd, X[0], X[1:] + [F5], want, 1e-5, 1e-10, exact=(arith == crd.ARITH_EXACT))
    for rnd in range(2):
        for var in variants:
            g.set_variant(var)
            for _ in range(2):
                rc, fe2, _ = g.f_lincomb_finish(50.0, c5, hb, hd, X, out, 1e-5, 1e-10)
            ctx.sync(); ctx.timer_start()
            for _ in range(reps):
                g.f_lincomb_finish(50.0, c5, hb, hd, X, out, 1e-5, 1e-10)
            ms = ctx.timer_stop() / reps
            crd.N_VLinearSum(1.0, out, -1.0, want, F5)
            same = crd.N_VMaxNorm(F5) == 0.0
            print("%s %s variant %2d: %.3f ms  %.0f GB/s  ynew identical %s  err sum equal %s" %
                  (model, "exact" if arith == crd.ARITH_EXACT else "fast ", var, ms, 96.0 * nx * rows / ms / 1e6, same, fe2 == e2), flush=True)
    for v in X + [out, want, F5]:
        v.destroy()
    g.close()
ctx.close()
