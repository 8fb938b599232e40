"""Scratch: RHS and integrator throughput on the reference's default (L2-resident, launch-bound) grids and
the Goldbeter BASELINE config, printed as JSON lines (for profiles/README.md)."""
import json, sys, time
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
ctx = crd.Context(0)
def rhs_rate(model, nx, ny, arith, reps=300):
    g = crd.Grid(ctx, crd.make_params(model, nx, ny, arith=arith))
    y, d = g.new_vector(), g.new_vector()
    g.fill_synthetic(y)
    for _ in range(10): g.f(50.0, y, d)
    ctx.sync(); ctx.timer_start()
    for _ in range(reps): g.f(50.0, y, d)
    ms = ctx.timer_stop() / reps
    y.destroy(); d.destroy(); g.close()
    return ms
def integ_rate(model, nx, ny, tfinal, fused, reuse, resident=False, variant=0, arith=0):
    beta = 1.25 if model.startswith("fhn") else 0.4
    g = crd.Grid(ctx, crd.make_params(model, nx, ny, beta=beta, vary_beta=0, t_boundary=0.0, arith=arith))
    if variant: g.set_variant(variant)
    y = g.new_vector()
    s0, s1 = (-beta, beta**3 - 3*beta) if model.startswith("fhn") else (0.392, 1.6469)
    g.fill_initial_conditions(y, 0.1, 0.5, 1, s0, s1)
    s = crd.ARKodeSolver(g, y, fused=fused, reuse_first_stage=reuse, resident=resident)
    s.ARKode(tfinal * 1e-3); ctx.sync(); n0 = s.stats()["nst"]   # set-up, first-step estimate, first launch: not timed
    t0 = time.time(); flag, t = s.ARKode(tfinal); ctx.sync(); dt = time.time() - t0
    st = s.stats(); s.free(); y.destroy(); g.close()
    return dict(flag=flag, seconds=dt, nst=st["nst"], nfe=st["nfe"], netf=st["netf"], steps_per_s=(st["nst"] - n0)/dt,
                us_per_step=1e6 * dt / max(1, st["nst"] - n0))
for model, nx, ny in (("fhn_torus", 400, 1600), ("gb_torus", 100, 400), ("fhn_flat", 400, 1600)):
    for arith in (0, 1):
        ms = rhs_rate(model, nx, ny, arith, reps=300 if nx * ny < 1e7 else 60)
        print(json.dumps(dict(kind="rhs", model=model, nx=nx, ny=ny, arith="exact" if arith == 0 else "fast", us_per_rhs=round(ms*1e3, 2),
                              Gpts=round(nx*ny/ms/1e6, 2), GBs=round(nx*ny*32/ms/1e6, 1))), flush=True)
for model, nx, ny, tf in (("fhn_torus", 400, 1600, 2.0), ("gb_torus", 100, 400, 0.5)):
    for fused, reuse, resident, variant, arith in ((False, False, False, 0, 0), ("ops", False, False, 0, 0), ("full", False, False, 0, 0),
                                                   ("full", True, False, 0, 0), ("full", False, True, 0, 0), ("full", False, True, 0, 1),
                                                   ("full", False, True, 120, 0),
                                                   ("full", False, True, 122, 0)):
        r = integ_rate(model, nx, ny, tf, fused, reuse, resident, variant, arith)
        r.update(kind="integrate", model=model, nx=nx, ny=ny, tfinal=tf, fused=fused, reuse_first_stage=reuse, resident=resident,
                 variant=variant, arith="exact" if arith == 0 else "fast")
        print(json.dumps(r), flush=True)
