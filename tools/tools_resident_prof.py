"""Scratch: one resident ARKode call on a default mesh (for ncu)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
model = sys.argv[1] if len(sys.argv) > 1 else "fhn_torus"
nx, ny = (400, 1600) if model.startswith("fhn") else (100, 400)
tf = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 150   # 150: the instantiation with per-phase cycle counters; 0: the default kernel
ctx = crd.Context(0)
beta = 1.25 if model.startswith("fhn") else 0.4
g = crd.Grid(ctx, crd.make_params(model, nx, ny, beta=beta, vary_beta=0, t_boundary=0.0))
y = g.new_vector()
if variant: g.set_variant(variant)
s0, s1 = (-beta, beta**3 - 3*beta) if model.startswith("fhn") else (0.392, 1.6469)
g.fill_initial_conditions(y, 0.1, 0.5, 1, s0, s1)
s = crd.ARKodeSolver(g, y, fused="full", resident=True)
s.ARKode(tf * 1e-2); ctx.sync(); n0 = s.stats()["nst"]
t0 = time.time(); s.ARKode(tf); ctx.sync(); dt = time.time() - t0
st = s.stats()
cy = g.resident_cycles()
passes = st["nfe"]
print(model, "variant", variant, "steps", st["nst"] - n0, "us/step %.2f" % (1e6 * dt / (st["nst"] - n0)), st)
print("   cycles per step:", {k: round(v / (st["nst"] - n0)) for k, v in cy.items()})
