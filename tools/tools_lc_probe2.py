"""Scratch: is the fused stage kernel bound by the DRAM access pattern?  Same point count, different row lengths:
nx = 256 makes every tile row adjacent in memory to the next (linear streams), nx = 16384 jumps 256 KB between tile rows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
ctx = crd.Context(0)
for nx, ny in ((16384, 16384), (4096, 65536), (1024, 262144), (256, 1048576)):
    g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, arith=0))
    V = [g.new_vector() for _ in range(2)]
    for j, v in enumerate(V):
        ctx.fill_synthetic("fhn_torus", 2 * nx * ny, v.device_ptr, seed=100 + j)
    d = g.new_vector()
    def t(fn, reps=30):
        for _ in range(3): fn()
        ctx.sync(); ctx.timer_start()
        for _ in range(reps): fn()
        return round(ctx.timer_stop() / reps, 3)
    res = {"nx": nx, "ny": ny, "plain": t(lambda: g.f(50.0, V[0], d))}
    for variant in (0, 13, 21):
        g.set_variant(variant)
        res["lc2_v%d" % variant] = t(lambda: g.f_lincomb(50.0, [1.0, 0.01], V, d))
    g.set_variant(0)
    res["linearsum"] = t(lambda: crd.N_VLinearSum(1.0, V[0], 0.01, V[1], d))
    print(res, flush=True)
    for v in V + [d]: v.destroy()
    g.close()
