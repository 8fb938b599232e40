"""f(t1, y) and the next step's second stage f(t2, y + c f1): one pass (crd_rhs_pair) against the two launches it replaces.
python tools/prof_pair.py [NXxNY ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
ctx = crd.Context(0)
sizes = [(16384, 16384), (4096, 16384), (2048, 8192), (1024, 4096), (600, 2400)]
if len(sys.argv) > 1:
    sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for model in ("fhn_torus", "gb_torus"):
    for arith in (crd.ARITH_EXACT, crd.ARITH_FAST):
        for nx, ny in sizes:
            if model == "gb_torus" and nx * ny > (1 << 27):
                continue
            g = crd.Grid(ctx, crd.make_params(model, nx, ny, arith=arith))
            y, f1, f2 = g.new_vector(), g.new_vector(), g.new_vector()
            g.fill_synthetic(y)
            pts = nx * ny
            reps = max(10, min(500, int(2e9 / pts)))
            def timeit(call):
                for _ in range(3):
                    call()
                ctx.sync(); ctx.timer_start()
                for _ in range(reps):
                    call()
                return ctx.timer_stop() / reps
            def separate():
                g.f(50.0, y, f1)
                g.f_lincomb(50.1, [1.0, 5e-4], [y, f1], f2)
            for rnd in range(2):
                ms_p = timeit(lambda: g.f_pair(50.0, 50.1, 5e-4, y, f1, f2))
                ms_s = timeit(separate)
                print("%s %s %5d x %5d: pair %8.1f us (%4.0f GB/s at 48 B/pt)   separate %8.1f us (%4.0f GB/s at 80 B/pt)   x%.2f" % (
                    model, "exact" if arith == crd.ARITH_EXACT else "fast ", nx, ny, 1e3 * ms_p, 48.0 * pts / ms_p / 1e6, 1e3 * ms_s,
                    80.0 * pts / ms_s / 1e6, ms_s / ms_p), flush=True)
            for v in (y, f1, f2):
                v.destroy()
            g.close()
ctx.close()
