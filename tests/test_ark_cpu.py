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
