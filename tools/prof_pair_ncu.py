"""One launch of the pair pass per arithmetic for ncu (tools: ncu -k regex:rhs_pair_tile --metrics ... python tools/prof_pair_ncu.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
ctx = crd.Context(0)
nx, ny = 16384, 4096
for arith in (crd.ARITH_EXACT, crd.ARITH_FAST):
    g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, arith=arith))
    y, f1, f2 = g.new_vector(), g.new_vector(), g.new_vector()
    g.fill_synthetic(y)
    for _ in range(3):
        g.f_pair(50.0, 50.1, 5e-4, y, f1, f2)
    ctx.sync()
    for v in (y, f1, f2):
        v.destroy()
    g.close()
ctx.close()
