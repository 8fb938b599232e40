"""Round-2 driver: plain RHS per kernel variant (default build: 0 = automatic, 10, 13, 15, 20, 21).  python tools/prof_variants.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
ctx = crd.Context(0)
for model, nx, rows in (("gb_torus", 8192, 4096), ("fhn_torus", 16384, 8192)):
    for arith in (crd.ARITH_EXACT, crd.ARITH_FAST):
        g = crd.Grid(ctx, crd.make_params(model, nx, rows, arith=arith))
        y, d = g.new_vector(), g.new_vector()
        g.fill_synthetic(y)
        for rnd in range(2):
            for var in (0, 13, 15, 10, 21, 20):
                g.set_variant(var)
                for _ in range(5):
                    g.f(50.0, y, d)
                ctx.sync(); ctx.timer_start()
                for _ in range(200):
                    g.f(50.0, y, d)
                ms = ctx.timer_stop() / 200
                print("%s %s variant %2d: %.4f ms  %.0f GB/s (%.3f of 6553)" % (model, "exact" if arith == crd.ARITH_EXACT else "fast ", var, ms,
                      32.0 * nx * rows / ms / 1e6, 32.0 * nx * rows / ms / 1e6 / 6553.3), flush=True)
        y.destroy(); d.destroy(); g.close()
ctx.close()
