"""Scratch: a short launch sequence for ncu (FHN / Goldbeter torus, exact / fast, one variant)."""
import sys
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
nx = ny = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = crd.Context(0)
for model in ("fhn_torus", "gb_torus"):
    for arith in (0, 1):
        g = crd.Grid(ctx, crd.make_params(model, nx, ny, arith=arith))
        g.set_variant(variant)
        y, d = g.new_vector(), g.new_vector()
        g.fill_synthetic(y)
        for _ in range(2):
            g.f(50.0, y, d)
        ctx.sync()
        y.destroy(); d.destroy(); g.close()
print("prof ok")
