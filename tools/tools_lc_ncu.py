"""Scratch: one launch each of the 2-vector fused stage RHS with the staged tile (13) and the streaming kernel (21), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
ctx = crd.Context(0)
nx = ny = 16384
g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, arith=0))
V = [g.new_vector() for _ in range(2)]
for j, v in enumerate(V):
    ctx.fill_synthetic("fhn_torus", 2 * nx * ny, v.device_ptr, seed=100 + j)
d = g.new_vector()
for variant in (13, 21):
    g.set_variant(variant)
    for _ in range(2):
        g.f_lincomb(50.0, [1.0, 0.01], V, d)
ctx.sync()
