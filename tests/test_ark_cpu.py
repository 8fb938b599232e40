"""The explicit RK driver (crdmodel_b200/host/crd_ark.cpp) on the CPU: host N_Vector + the reference's
f() restated in C.  Checks the integrator's own contract (order, tolerances, counters, dense output,
max-steps, the reference's whole main() running end to end behind the ARKode-legacy names)."""
import ctypes as C
import os

import numpy as np
import pytest

P_, D_, L_ = C.c_void_p, C.c_double, C.c_long
RHS = C.CFUNCTYPE(C.c_int, D_, P_, P_, P_)


@pytest.fixture(scope="module")
def K(oracle):
    L = oracle.lib()
    L.N_VMake_Parallel.restype = P_
    L.N_VMake_Parallel.argtypes = [C.c_int, L_, L_, P_]
    L.N_VGetArrayPointer.restype = C.POINTER(D_)
    L.N_VGetArrayPointer.argtypes = [P_]
    L.ARKodeCreate.restype = P_
    L.ARKodeInit.argtypes = [P_, P_, P_, D_, P_]
    L.ARKodeSStolerances.argtypes = [P_, D_, D_]
    L.ARKodeSetUserData.argtypes = [P_, P_]
    L.ARKodeSetMaxNumSteps.argtypes = [P_, L_]
    L.ARKode.argtypes = [P_, D_, P_, C.POINTER(D_), C.c_int]
    L.ARKodeFree.argtypes = [C.POINTER(P_)]
    L.ARKodeGetNumSteps.argtypes = [P_, C.POINTER(L_)]
    L.ARKodeGetNumRhsEvals.argtypes = [P_, C.POINTER(L_), C.POINTER(L_)]
    L.ARKodeGetNumErrTestFails.argtypes = [P_, C.POINTER(L_)]
    L.crd_ARKodeSetReuseFirstStage.argtypes = [P_, C.c_int]
    L.crd_ARKodeSetFixedStep.argtypes = [P_, D_]
    return L


def integrate(K, f, y0, touts, rtol=1e-5, atol=1e-10, user=None, mxsteps=200000, reuse=False, fixed=0.0):
    y = np.array(y0, dtype=np.float64)
    Y = K.N_VMake_Parallel(0, y.size, y.size, y.ctypes.data)
    mem = P_(K.ARKodeCreate())
    assert K.ARKodeInit(mem, C.cast(f, P_), None, 0.0, Y) == 0
    assert K.ARKodeSStolerances(mem, rtol, atol) == 0
    K.ARKodeSetUserData(mem, user)
    K.ARKodeSetMaxNumSteps(mem, mxsteps)
    K.crd_ARKodeSetReuseFirstStage(mem, 1 if reuse else 0)
    if fixed:
        K.crd_ARKodeSetFixedStep(mem, fixed)
    outs, flags = [], []
    t = D_()
    for tout in touts:
        flags.append(K.ARKode(mem, tout, Y, C.byref(t), 1))
        outs.append(y.copy())
        if flags[-1] < 0:
            break
    nst, nfe, nfi, netf = L_(), L_(), L_(), L_()
    K.ARKodeGetNumSteps(mem, C.byref(nst)); K.ARKodeGetNumRhsEvals(mem, C.byref(nfe), C.byref(nfi)); K.ARKodeGetNumErrTestFails(mem, C.byref(netf))
    K.ARKodeFree(C.byref(mem))
    return outs, flags, dict(nst=nst.value, nfe=nfe.value, netf=netf.value, t=t.value)


def make_rhs(K, fun):
    def cb(t, y, ydot, _u):
        n = 2
        yp, dp = K.N_VGetArrayPointer(y), K.N_VGetArrayPointer(ydot)
        d = fun(t, np.array([yp[i] for i in range(n)]))
        for i in range(n):
            dp[i] = d[i]
        return 0
    return RHS(cb)


def test_harmonic_oscillator_tolerance_and_dense_output(K):
    f = make_rhs(K, lambda t, y: np.array([y[1], -y[0]]))
    touts = [0.5, 1.0, 2.5, 7.0]
    outs, flags, st = integrate(K, f, [1.0, 0.0], touts, rtol=1e-8, atol=1e-12)
    assert flags == [0, 0, 0, 0]
    for tout, y in zip(touts, outs):
        assert abs(y[0] - np.cos(tout)) < 2e-6 and abs(y[1] + np.sin(tout)) < 2e-6
    assert st["nfe"] >= 6 * st["nst"]          # 5 stages + 1 for the dense-output derivative per step
    outs2, _, st2 = integrate(K, f, [1.0, 0.0], touts, rtol=1e-8, atol=1e-12, reuse=True)
    assert all(a.tobytes() == b.tobytes() for a, b in zip(outs, outs2))   # reuse of f(tn, yn) changes nothing
    assert st2["nst"] == st["nst"] and st2["nfe"] < st["nfe"]


def test_butcher_table_satisfies_the_order_conditions(K):
    """The method restated from ARKode 1.x's documentation (Zonneveld 5-3-4; SUNDIALS itself is not available here): the
    weights b must satisfy all eight conditions of order 4, the embedding b2 the four of order 3 (and NOT those of order 4, or the
    difference would not estimate the error), c must be the row sums of a strictly lower-triangular A."""
    K.crd_ARKodeGetButcherTable.argtypes = [P_] + [C.POINTER(C.c_int)] * 3 + [C.POINTER(D_)] * 4
    mem = P_(K.ARKodeCreate())
    s, q, p = C.c_int(), C.c_int(), C.c_int()
    A, c, b, b2 = (D_ * 64)(), (D_ * 8)(), (D_ * 8)(), (D_ * 8)()
    assert K.crd_ARKodeGetButcherTable(mem, C.byref(s), C.byref(q), C.byref(p), A, c, b, b2) == 0
    K.ARKodeFree(C.byref(mem))
    assert (s.value, q.value, p.value) == (5, 4, 3)
    n = s.value
    A = np.array(A[:]).reshape(8, 8)[:n, :n]
    c, b, b2 = np.array(c[:n]), np.array(b[:n]), np.array(b2[:n])
    assert np.all(np.triu(A) == 0.0)                      # explicit
    assert np.allclose(A.sum(1), c, atol=1e-15)

    def conditions(w):
        return [w.sum() - 1, w @ c - 1 / 2, w @ c ** 2 - 1 / 3, w @ (A @ c) - 1 / 6,                       # orders 1-3
                w @ c ** 3 - 1 / 4, (w * c) @ (A @ c) - 1 / 8, w @ (A @ c ** 2) - 1 / 12, w @ (A @ (A @ c)) - 1 / 24]   # order 4
    assert np.max(np.abs(conditions(b))) < 1e-15
    e = np.abs(conditions(b2))
    assert np.max(e[:4]) < 1e-14 and np.max(e[4:]) > 1e-3


def test_fourth_order_convergence_fixed_step(K):
    f = make_rhs(K, lambda t, y: np.array([y[1], -y[0]]))
    errs = []
    for h in (0.1, 0.05):
        outs, flags, _ = integrate(K, f, [1.0, 0.0], [1.0], fixed=h)
        errs.append(abs(outs[0][0] - np.cos(1.0)))
    assert 10.0 < errs[0] / errs[1] < 24.0      # ~2^4


def test_max_steps_returns_too_much_work(K):
    f = make_rhs(K, lambda t, y: np.array([y[1], -y[0]]))
    outs, flags, st = integrate(K, f, [1.0, 0.0], [100.0], rtol=1e-10, atol=1e-14, mxsteps=5)
    assert flags == [-1] and st["nst"] == 5


def test_rhs_failure_is_reported(K):
    bad = RHS(lambda t, y, ydot, u: -1)
    outs, flags, _ = integrate(K, bad, [1.0, 0.0], [1.0])
    assert flags[0] < 0


def test_fhn_torus_trajectory_cpu(K, oracle):
    """Reference ICs (uniform 1,1 with varyBeta = 1) on a coarse torus: the trajectory stays bounded,
    the frozen rows do not move before tBoundary, counters are consistent."""
    nx, ny = 16, 64
    Pm = oracle.make_params("fhn_torus", nx, ny, t_boundary=1.0)
    y0 = np.ones(2 * nx * ny)
    f = C.cast(oracle.lib().crd_oracle_f, P_)
    outs, flags, st = integrate(K, f, y0, [0.5, 2.0], user=C.cast(C.pointer(Pm), P_))
    assert flags == [0, 0]
    a = outs[0].reshape(ny, nx, 2)
    assert np.all(a[0] == 1.0) and np.all(a[-1] == 1.0)       # t < tBoundary: rows 0 and ny-1 frozen
    assert np.abs(outs[1]).max() < 5.0 and st["nst"] > 3
    b = outs[1].reshape(ny, nx, 2)
    assert np.any(b[0] != 1.0)


def test_reference_main_runs_end_to_end(oracle, tmp_path):
    """The reference's whole program (ini parse -> ICs -> ARKode loop -> text files), compiled in place,
    with this repository's integrator behind the ARKode names: the output-file contract of SURVEY.md §5.5."""
    if not oracle.ref_available("fhn_torus"):
        pytest.skip("oracle/_ref not built")
    ini = tmp_path / "fhn.ini"
    ini.write_text("[Parameters]\ndiffusion = 0.12\nbeta = 1.25\nsurfaceWidth = 20\nsurfaceLength = 80\nwaveLength = 0.1\n"
                   "waveWidth = 0.5\nwaveInside = 0\noutputTimestep = 3\ntBoundary = 38\ntFinal = 1.5\nthetaMesh = 12\n"
                   "betaMin = 0.7\nbetaMax = 1.7\n[System]\nincludeAllVars = 1\nvaryBeta = 0\n")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        rc = oracle.ref_lib("fhn_torus").crd_ref_main(str(ini).encode(), 1)
    finally:
        os.chdir(cwd)
    assert rc == 0
    sub = (tmp_path / "FHNmodel_torus_subdomain.000.txt").read_text().split()
    assert sub[:6] == ["12", "48", "0", "11", "0", "47"]
    u = np.loadtxt(tmp_path / "FHNmodel_torus_u.000.txt")
    v = np.loadtxt(tmp_path / "FHNmodel_torus_v.000.txt")
    assert u.shape == (4, 12 * 48) and v.shape == (4, 12 * 48)     # ICs + 3 outputs
    assert np.isfinite(u).all() and set(np.unique(u[0])) == {-1.25, 0.75}


# ---- the erk_evolve hook (crd_ark.h): the vector implementation may run the whole step loop itself ----------------
S_MAX = 8


class ErkState(C.Structure):
    """crd_erk_state of include/crd_ark.h"""
    _fields_ = [("s", C.c_int), ("p", C.c_int),
                ("A", D_ * S_MAX * S_MAX), ("b", D_ * S_MAX), ("d", D_ * S_MAX), ("c", D_ * S_MAX),
                ("rtol", D_), ("atol", D_),
                ("k1", D_), ("k2", D_), ("k3", D_), ("bias", D_), ("safety", D_), ("growth", D_), ("etamxf", D_), ("etamin", D_),
                ("lbound", D_), ("ubound", D_),
                ("small_nef", C.c_int), ("maxnef", C.c_int), ("nglobal", L_),
                ("tout", D_), ("itask", C.c_int), ("max_steps", L_),
                ("tn", D_), ("next_h", D_), ("hold", D_), ("eta", D_), ("etamax", D_), ("ehist", D_ * 2), ("ynorm_sq", D_),
                ("nst", L_), ("nst_attempts", L_), ("nfe", L_), ("netf", L_),
                ("yn", P_), ("yold", P_), ("ycur", P_), ("fnew", P_), ("fold", P_), ("F", P_ * S_MAX),
                ("flag", C.c_int), ("h_failed", D_)]


EVOLVE = C.CFUNCTYPE(C.c_int, C.POINTER(ErkState), P_)


class FusedOps(C.Structure):
    _fields_ = [("lincomb", P_), ("erk_finish", P_), ("rhs_lincomb", P_), ("erk_evolve", EVOLVE), ("rhs_lincomb_finish", P_)]


def integrate_with_evolve(K, f, y0, touts, evolve, mxsteps=200000):
    K.crd_ARKodeSetFusedOps.argtypes = [P_, C.POINTER(FusedOps)]
    ops = FusedOps(None, None, None, EVOLVE(evolve), None)
    y = np.array(y0, dtype=np.float64)
    Y = K.N_VMake_Parallel(0, y.size, y.size, y.ctypes.data)
    mem = P_(K.ARKodeCreate())
    assert K.ARKodeInit(mem, C.cast(f, P_), None, 0.0, Y) == 0
    assert K.ARKodeSStolerances(mem, 1e-8, 1e-12) == 0
    K.ARKodeSetMaxNumSteps(mem, mxsteps)
    assert K.crd_ARKodeSetFusedOps(mem, C.byref(ops)) == 0
    outs, flags, ts = [], [], []
    t = D_()
    for tout in touts:
        flags.append(K.ARKode(mem, tout, Y, C.byref(t), 1))
        outs.append(y.copy()); ts.append(t.value)
    nst, nfe, nfi = L_(), L_(), L_()
    K.ARKodeGetNumSteps(mem, C.byref(nst)); K.ARKodeGetNumRhsEvals(mem, C.byref(nfe), C.byref(nfi))
    K.ARKodeFree(C.byref(mem))
    return outs, flags, ts, dict(nst=nst.value, nfe=nfe.value)


def test_evolve_hook_not_applicable_falls_back_to_the_host_loop(K):
    f = make_rhs(K, lambda t, y: np.array([y[1], -y[0]]))
    calls = []
    def evolve(st, user):
        calls.append(st.contents.tn)
        return 1          # "does not apply": the integrator must carry on by itself and stop asking
    touts = [0.5, 1.0, 2.5]
    a, fa, _, sa = integrate_with_evolve(K, f, [1.0, 0.0], touts, evolve)
    b, fb, sb = integrate(K, f, [1.0, 0.0], touts, rtol=1e-8, atol=1e-12)
    assert fa == fb == [0, 0, 0] and len(calls) == 1
    assert all(x.tobytes() == y.tobytes() for x, y in zip(a, b)) and sa["nst"] == sb["nst"]


def test_evolve_hook_owns_the_loop_and_returns_permuted_vectors(K):
    """A stand-in evolve that jumps to the exact solution just past tout: ARKode must take the state, the vector roles
    (handles are swapped, not copied), the counters and the step size from it and interpolate to tout itself."""
    f = make_rhs(K, lambda t, y: np.array([y[1], -y[0]]))
    def put(v, vals):
        p = K.N_VGetArrayPointer(v)
        for i, x in enumerate(vals):
            p[i] = x
    def evolve(stp, user):
        st = stp.contents
        assert st.s == 5 and st.itask == 1 and st.next_h > 0 and st.nglobal == 2
        assert abs(st.A[4][3] + 1.0 / 32.0) < 1e-16 and abs(st.d[4] - 16.0 / 3.0) < 1e-15     # Zonneveld 5-3-4, d = b - b_embedded
        h = 1e-3
        t1 = st.tout + 0.4 * h          # the accepted step [t1 - h, t1] brackets tout
        t0 = t1 - h
        # the new state goes into the caller's ycur array, the previous one stays where yn was: roles are permuted
        put(st.ycur, [np.cos(t1), -np.sin(t1)]); put(st.yn, [np.cos(t0), -np.sin(t0)])
        put(st.fold, [-np.sin(t1), -np.cos(t1)]); put(st.fnew, [-np.sin(t0), -np.cos(t0)])
        st.yn, st.yold, st.ycur = st.ycur, st.yn, st.yold
        st.fnew, st.fold = st.fold, st.fnew
        st.tn, st.hold, st.next_h = t1, h, h
        st.nst += 7; st.nst_attempts += 8; st.nfe += 40; st.netf += 1
        st.flag = 0
        return 0
    touts = [0.5, 1.0, 2.5]
    outs, flags, ts, st = integrate_with_evolve(K, f, [1.0, 0.0], touts, evolve)
    assert flags == [0, 0, 0] and ts == touts
    for tout, y in zip(touts, outs):
        assert abs(y[0] - np.cos(tout)) < 1e-12 and abs(y[1] + np.sin(tout)) < 1e-12      # cubic Hermite over h = 1e-3
    assert st["nst"] == 21 and st["nfe"] >= 120


def test_evolve_hook_failures_are_reported_like_the_host_loop(K):
    f = make_rhs(K, lambda t, y: np.array([y[1], -y[0]]))
    def too_much_work(stp, user):
        st = stp.contents
        st.nst += st.max_steps
        st.flag = -1
        return 0
    outs, flags, ts, st = integrate_with_evolve(K, f, [1.0, 0.0], [100.0], too_much_work, mxsteps=5)
    assert flags == [-1] and ts == [0.0] and st["nst"] == 5 and outs[0].tolist() == [1.0, 0.0]
    outs, flags, ts, st = integrate_with_evolve(K, f, [1.0, 0.0], [1.0], lambda stp, user: -1)
    assert flags == [-20]       # ARK_MEM_FAIL: the device loop could not be launched


def test_controller_pow_is_accurate_and_independent_of_libm(tmp_path):
    """crd_pow.h: x^y as one fixed sequence of IEEE operations shared by the host controller (crd_ark.cpp) and the device-resident
    loop (crd_resident.cu), so both choose the same step sizes bit for bit.  Here: accuracy against libm over the controller's
    domain (bases max(bias*dsm, 1e-10) .. 1e5, exponents -k1/p, k2/p, -k3/p and a few others)."""
    import ctypes
    import math
    import subprocess
    import numpy as np
    src = tmp_path / "p.c"
    src.write_text('#include "crd_pow.h"\ndouble crd_test_pow(double x, double y) { return crd_pow_pos(x, y); }\n')
    so = tmp_path / "p.so"
    subprocess.run(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fPIC", "-shared", "-I" + os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "crdmodel_b200", "host"),
                    str(src), "-o", str(so)], check=True)
    f = ctypes.CDLL(str(so)).crd_test_pow
    f.restype = ctypes.c_double
    f.argtypes = [ctypes.c_double, ctypes.c_double]
    rng = np.random.default_rng(0)
    worst = 0.0
    for y in (-0.58 / 3, 0.21 / 3, -0.1 / 3, -0.25, 0.5, 1.0, -1.0, 2.0 / 3):
        for x in np.concatenate([10.0 ** rng.uniform(-10, 5, 4000), [1e-10, 1.0, 2.0, 0.5, 1.5, math.sqrt(2.0), 1e5]]):
            got, want = f(float(x), y), math.pow(float(x), y)
            worst = max(worst, abs(got - want) / want)
    assert worst < 4e-15, worst
    assert f(1.0, -0.58 / 3) == 1.0 and f(4.0, 0.5) == 2.0
