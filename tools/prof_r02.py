"""Round-2 profiling driver (for ncu under gpurun): the kernels of one large-mesh integrator step and the Goldbeter RHS,
a few launches each.  python tools/prof_r02.py [rows]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = crd.Context(0)
h = 1e-3
c5 = [1.0, h * 5 / 32, h * 7 / 32, h * 13 / 32, -h / 32]
hb = [h / 6, h / 3, h / 3, h / 6, 0.0]
hd = [h * (1 / 6 + 0.5), h * (1 / 3 - 7 / 3), h * (1 / 3 - 7 / 3), h * (1 / 6 - 13 / 6), h * 16 / 3]
for model, nx in (("fhn_torus", 16384), ("gb_torus", 8192)):
    for arith in (crd.ARITH_EXACT, crd.ARITH_FAST):
        g = crd.Grid(ctx, crd.make_params(model, nx, rows, arith=arith))
        X = [g.new_vector() for _ in range(5)]
        out = g.new_vector()
        for j, v in enumerate(X):
            g.fill_synthetic(v, seed=0x5EED + j)
            if j:
                crd.N_VScale(0.25, v, v)
        for _ in range(reps):
            g.f(50.0, X[0], out)                                   # plain RHS (tiled kernel)
        for _ in range(reps):
            g.f_lincomb(50.0, [1.0, 0.5 * h], X[:2], out)          # 2-vector stage (streaming kernel, 3 CTAs/SM)
        for _ in range(reps):
            g.f_lincomb_finish(50.0, c5, hb, hd, X, out, 1e-5, 1e-10)   # last stage + finish
        ctx.sync()
        for v in X + [out]:
            v.destroy()
        g.close()
ctx.close()
print("done")
