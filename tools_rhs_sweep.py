"""Scratch: time RHS kernel variants on the GPU box (not part of the product)."""
import sys, json
import crdmodel_b200 as crd
nx = ny = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ctx = crd.Context(0)
res = []
for model in ("fhn_torus", "gb_torus", "fhn_flat"):
    for arith in (0, 1):
        g = crd.Grid(ctx, crd.make_params(model, nx, ny, arith=arith))
        y, d = g.new_vector(), g.new_vector()
        g.fill_synthetic(y)
        for variant in (0, 1, 2, 3):
            g.set_variant(variant)
            for _ in range(3): g.f(50.0, y, d)
            ctx.sync()
            ctx.timer_start()
            reps = 20
            for _ in range(reps): g.f(50.0, y, d)
            ms = ctx.timer_stop() / reps
            gbs = nx * ny * 32 / ms / 1e6
            res.append(dict(model=model, arith=arith, variant=variant, ms=ms, GBs=gbs, Gpts=nx*ny/ms/1e6))
            print(res[-1], flush=True)
        y.destroy(); d.destroy(); g.close()
# streaming vector ops for reference
n = 2 * nx * ny
a, b, c = crd.NVector(ctx, n), crd.NVector(ctx, n), crd.NVector(ctx, n)
crd.N_VConst(1.0, a); crd.N_VConst(2.0, b)
for name, fn, bytes_per in (("linearsum", lambda: crd.N_VLinearSum(2.0, a, 3.0, b, c), 24), ("scale", lambda: crd.N_VScale(2.0, a, c), 16), ("const", lambda: crd.N_VConst(1.0, c), 8)):
    for _ in range(3): fn()
    ctx.sync(); ctx.timer_start()
    for _ in range(20): fn()
    ms = ctx.timer_stop() / 20
    print(name, ms, "ms", n * bytes_per / ms / 1e6, "GB/s", flush=True)
import time
t0=time.time(); r = crd.N_VWrmsNorm(a, b); ctx.sync(); 
ctx.timer_start()
for _ in range(10): crd.N_VWrmsNorm(a, b)
ms = ctx.timer_stop()/10
print("wrmsnorm", ms, "ms", n*16/ms/1e6, "GB/s")
json.dump(res, open("gpurun_out/sweep_%d.json" % nx, "w"))
