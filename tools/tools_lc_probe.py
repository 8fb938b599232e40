import os, sys
sys.path.insert(0, "/root/repo")
import crdmodel_b200 as crd
nx = ny = 16384
ctx = crd.Context(0)
g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, arith=0))
V = [g.new_vector() for _ in range(3)]
for j, v in enumerate(V):
    ctx.fill_synthetic("fhn_torus", 2 * nx * ny, v.device_ptr, seed=100 + j)
d = g.new_vector()
def t(label, fn, reps=40):
    for _ in range(3): fn()
    ctx.sync(); ctx.timer_start()
    for _ in range(reps): fn()
    print(label, round(ctx.timer_stop() / reps, 3), "ms", flush=True)
t("plain f", lambda: g.f(50.0, V[0], d))
t("lincomb n=1", lambda: g.f_lincomb(50.0, [1.0], V[:1], d))
t("lincomb n=2 distinct", lambda: g.f_lincomb(50.0, [1.0, 0.01], V[:2], d))
t("lincomb n=2 same vector twice", lambda: g.f_lincomb(50.0, [1.0, 0.01], [V[0], V[0]], d))
t("linearsum 2R+1W", lambda: crd.N_VLinearSum(1.0, V[0], 0.01, V[1], d))
t("lincomb kernel n=2", lambda: crd.N_VLinearCombination([1.0, 0.01], V[:2], d))
