"""Mid-size meshes (1-70 Mi points): which kernel evaluates f(), the fused 2-vector stage and the last stage + finish fastest.
python tools/prof_mid.py   (CRD_STREAM_SEG_ROWS=128 restores the fixed segment length of the streaming kernel)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
ctx = crd.Context(0)
h = 1e-3
c5 = [1.0, h * 5 / 32, h * 7 / 32, h * 13 / 32, -h / 32]
hb = [h / 6, h / 3, h / 3, h / 6, 0.0]
hd = [h * (1 / 6 + 0.5), h * (1 / 3 - 7 / 3), h * (1 / 3 - 7 / 3), h * (1 / 6 - 13 / 6), h * 16 / 3]
sizes = [(600, 2400), (800, 3200), (1000, 4000), (1024, 4096), (1400, 5600), (2048, 8192), (4096, 16384)]
if len(sys.argv) > 1:
    sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for nx, ny in sizes:
    g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny))
    X = [g.new_vector() for _ in range(5)]
    d, F5 = g.new_vector(), g.new_vector()
    for j, v in enumerate(X):
        g.fill_synthetic(v, seed=0x5EED + j)
        if j:
            crd.N_VScale(0.25, v, v)
    pts = nx * ny
    reps = max(20, min(1000, int(1e9 / pts)))
    def timeit(call):
        for _ in range(3):
            call()
        ctx.sync(); ctx.timer_start()
        for _ in range(reps):
            call()
        return ctx.timer_stop() / reps
    line = "%5d x %5d (%5.1f MB/vec):" % (nx, ny, 16 * pts / 1e6)
    for name, nvec, call in (("f", 2, lambda: g.f(50.0, X[0], d)), ("lc2", 3, lambda: g.f_lincomb(50.0, [1.0, 5e-4], X[:2], d))):
        for var in (0, 1, 13, 21):
            g.set_variant(var)
            ms = timeit(call)
            line += "  %s v%-2d %6.1f us %4.0f GB/s |" % (name, var, 1e3 * ms, 16.0 * nvec * pts / ms / 1e6)
        g.set_variant(0)
    rc = g.f_lincomb_finish(50.0, c5, hb, hd, X, d, 1e-5, 1e-10)[0]
    if rc == 0:
        ms = timeit(lambda: g.f_lincomb_finish(50.0, c5, hb, hd, X, d, 1e-5, 1e-10))
        line += "  fin fused %6.1f us %4.0f GB/s |" % (1e3 * ms, 96.0 * pts / ms / 1e6)
    def separate():
        g.f_lincomb(50.0, c5, X, F5)
        crd.N_VErkFinish(hb, hd, X[0], X[1:] + [F5], d, 1e-5, 1e-10, exact=True)
    ms = timeit(separate)
    line += "  fin separate %6.1f us" % (1e3 * ms)
    print(line, flush=True)
    for v in X + [d, F5]:
        v.destroy()
    g.close()
ctx.close()
