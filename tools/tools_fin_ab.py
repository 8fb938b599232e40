"""Scratch: A/B of the last-stage + finish kernel (crd_rhs_lincomb_finish) on the 16384 x 16384 FHN mesh:
grid variant 24 (raw vectors in registers, 2 CTAs/SM) vs 22 / 23 (partial sums, 3 / 2 CTAs/SM; 22 is the default); ynew must be bit-identical."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import crdmodel_b200 as crd
nx = ny = 16384
ctx = crd.Context(0)
h = 1e-3
c = [1.0, h * 5 / 32, h * 7 / 32, h * 13 / 32, -h / 32]
hb = [h / 6, h / 3, h / 3, h / 6, 0.0]
hd = [h * (1 / 6 + 0.5), h * (1 / 3 - 7 / 3), h * (1 / 3 - 7 / 3), h * (1 / 6 - 13 / 6), h * 16 / 3]
out = []
for model, arith in (("fhn_torus", 0), ("gb_torus", 0), ("fhn_torus", 1)):
    g = crd.Grid(ctx, crd.make_params(model, nx, ny, arith=arith))
    X = [g.new_vector() for _ in range(5)]
    for j, v in enumerate(X):
        ctx.fill_synthetic(model, 2 * nx * ny, v.device_ptr, seed=100 + j)
        if j:
            crd.N_VScale(0.25, v, v)
    got, ref = g.new_vector(), g.new_vector()
    sums = {}
    for variant in (24, 22, 23, 24, 22):
        g.set_variant(variant)
        dst = ref if variant == 24 else got
        for _ in range(3):
            rc, fe2, fy2 = g.f_lincomb_finish(50.0, c, hb, hd, X, dst, 1e-5, 1e-10)
        assert rc == 0
        ctx.sync(); ctx.timer_start()
        reps = 25
        for _ in range(reps):
            g.f_lincomb_finish(50.0, c, hb, hd, X, dst, 1e-5, 1e-10)
        ms = ctx.timer_stop() / reps
        same = None
        if variant != 24:
            crd.N_VLinearSum(1.0, got, -1.0, ref, got)
            same = crd.N_VMaxNorm(got) == 0.0
        rec = dict(model=model, arith="exact" if arith == 0 else "fast", variant=variant, ms=round(ms, 4),
                   GBs=round(nx * ny * 16 * 6 / ms / 1e6), fe2=fe2, fy2=fy2, ynew_identical=same)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    for v in X + [got, ref]:
        v.destroy()
    g.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/s3_fin_ab.json", "w"), indent=1)
