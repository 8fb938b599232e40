"""Scratch: A/B the RHS kernel variants on the GPU box (sustained runs, interleaved rounds, median)."""
import sys, json, statistics
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
nx = int(sys.argv[1].split("x")[0]) if len(sys.argv) > 1 else 16384
ny = int(sys.argv[1].split("x")[-1]) if len(sys.argv) > 1 else 16384
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 10, 11, 13, 14]
models = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fhn_torus", "gb_torus"]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 150
ctx = crd.Context(0)
out = []
for model in models:
    for arith in (0, 1):
        g = crd.Grid(ctx, crd.make_params(model, nx, ny, arith=arith))
        y, d = g.new_vector(), g.new_vector()
        g.fill_synthetic(y)
        times = {v: [] for v in variants}
        for rnd in range(3):
            for v in variants:
                g.set_variant(v)
                for _ in range(5): g.f(50.0, y, d)
                ctx.sync(); ctx.timer_start()
                for _ in range(reps): g.f(50.0, y, d)
                times[v].append(ctx.timer_stop() / reps)
        for v in variants:
            ms = statistics.median(times[v])
            r = dict(model=model, arith="exact" if arith == 0 else "fast", variant=v, ms=round(ms, 4), GBs=round(nx * ny * 32 / ms / 1e6, 1),
                     Gpts=round(nx * ny / ms / 1e6, 1), spread=[round(nx * ny * 32 / t / 1e6) for t in times[v]])
            out.append(r); print(r, flush=True)
        y.destroy(); d.destroy(); g.close()
json.dump(out, open("gpurun_out/sweep_%d.json" % nx, "w"))
