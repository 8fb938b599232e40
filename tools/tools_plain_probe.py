"""Scratch: plain RHS f(t, y) at 16384 x 16384, sustained: TMA tile kernel (default) vs the streaming kernel (20: 2 CTAs/SM, 21: 3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
ctx = crd.Context(0)
nx = ny = 16384
for model in ("fhn_torus", "gb_torus"):
    for arith in (0, 1):
        g = crd.Grid(ctx, crd.make_params(model, nx, ny, arith=arith))
        y, d = g.new_vector(), g.new_vector()
        g.fill_synthetic(y)
        for variant in (0, 20, 21, 0, 21):
            g.set_variant(variant)
            for _ in range(10): g.f(50.0, y, d)
            ctx.sync(); ctx.timer_start()
            reps = 300
            for _ in range(reps): g.f(50.0, y, d)
            ms = ctx.timer_stop() / reps
            print(dict(model=model, arith="exact" if arith == 0 else "fast", variant=variant, ms=round(ms, 4), GBs=round(nx * ny * 32 / ms / 1e6)), flush=True)
        y.destroy(); d.destroy(); g.close()
