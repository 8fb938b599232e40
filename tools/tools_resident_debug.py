"""Scratch: step-by-step comparison of the resident and the host-driven loop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import crdmodel_b200 as crd
from test_resident_gpu import smooth_state
ctx = crd.Context(0)
nx, ny = 64, 96
y0 = smooth_state("fhn_torus", nx, ny)
seq = {}
for resident in (False, True):
    grid = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, vary_beta=0, t_boundary=0.0))
    grid.set_resident(1 if resident else -1)
    y = crd.NVector.from_numpy(ctx, y0)
    s = crd.ARKodeSolver(grid, y, fused="full", resident=resident)
    rows = []
    for _ in range(8):
        flag, t = s.ARKode(10.0, crd.ARK_ONE_STEP)
        st = s.stats()
        rows.append((t, st["hlast"], st["nst_attempts"], st["nfe"], y.to_numpy().copy()))
    seq[resident] = rows
    s.free(); grid.close()
for a, b in zip(seq[False], seq[True]):
    print("t %.17g %.17g  h %.17g %.17g rel %.3g  att %d %d nfe %d %d  dy %.3g" % (a[0], b[0], a[1], b[1], abs(a[1] - b[1]) / a[1], a[2], b[2], a[3], b[3],
                                                                         np.abs(a[4] - b[4]).max()))
