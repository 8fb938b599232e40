"""Scratch: a few host-driven integrator steps on the 16384 x 16384 FHN mesh (for ncu: the fused last-stage kernel)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
ctx = crd.Context(0)
nx = ny = 16384
g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny))
y = g.new_vector(); g.fill_synthetic(y)
s = crd.ARKodeSolver(g, y, t0=50.0)
s.set_init_step(1e-9)
for _ in range(3):
    print(s.ARKode(51.0, crd.ARK_ONE_STEP))
ctx.sync(); t0 = time.time()
for _ in range(5):
    s.ARKode(51.0, crd.ARK_ONE_STEP)
ctx.sync(); print("ms/step", (time.time() - t0) / 5 * 1e3, s.stats())
