"""Mesh-size sweep (FHN torus, EXACT): f() and the fused 2-vector stage per evaluation, and integrator steps/s, from the
reference's default mesh up to the headline's scale — looks for holes between the L2-resident and the HBM-bound regimes.
python tools/prof_sizes.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crdmodel_b200 as crd
ctx = crd.Context(0)
rows = []
sizes = ((400, 1600), (600, 2400), (800, 3200), (1000, 4000), (1024, 4096), (1400, 5600), (2048, 8192), (4096, 16384), (8192, 16384))
if len(sys.argv) > 1:
    sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for nx, ny in sizes:
    g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny))
    y, d, z = g.new_vector(), g.new_vector(), g.new_vector()
    g.fill_synthetic(y); g.fill_synthetic(z, seed=7)
    pts = nx * ny
    reps = max(20, min(2000, int(2e9 / pts)))
    out = {"nx": nx, "ny": ny, "points": pts, "MB_per_vector": 16 * pts / 1e6}
    for name, nvec, call in (("f", 2, lambda: g.f(50.0, y, d)), ("f_lincomb2", 3, lambda: g.f_lincomb(50.0, [1.0, 5e-4], [y, z], d))):
        for _ in range(5):
            call()
        ctx.sync(); ctx.timer_start()
        for _ in range(reps):
            call()
        ms = ctx.timer_stop() / reps
        out[name] = {"us": 1e3 * ms, "GBs": 16.0 * nvec * pts / ms / 1e6}
    # integrator: 40 steps from the reference's initial conditions, host-driven fused and (where it applies) resident
    for label, resident in (("host_driven_fused", False), ("resident", True), ("resident_forced", True)):
        g.set_resident(1 if label == "resident_forced" else 0)
        g.fill_initial_conditions(y, 0.1, 0.5, 1, -1.25, 1.25 ** 3 - 3 * 1.25)
        try:
            s = crd.ARKodeSolver(g, y, fused="full", resident=resident, max_steps=40)
        except Exception as e:
            out[label] = {"error": str(e)[:80]}
            continue
        s.ARKode(1e-4); ctx.sync()
        n0 = s.stats(); t0 = time.time()
        s.ARKode(1e9); ctx.sync()
        dt = time.time() - t0; n1 = s.stats()
        att = max(1, n1["nst_attempts"] - n0["nst_attempts"])
        out[label] = {"us_per_attempt": 1e6 * dt / att, "attempts": att, "resident_launches": g.resident_launches,
                      "GBs_at_272B_per_point": 272.0 * pts * att / dt / 1e9}
        s.free()
    rows.append(out)
    print(json.dumps(out), flush=True)
    y.destroy(); d.destroy(); z.destroy(); g.close()
ctx.close()
