"""Trajectory parity: the same explicit RK driver runs (a) on the CPU with host vectors, the op-by-op sequence of N_Vector
operations and the reference's f() restated in C, (b) on the GPU with device vectors, the CUDA f() and either the same
op-by-op sequence or the fused kernels.  With EXACT arithmetic the RHS is bit-identical (FHN), the element-wise operations
are bit-identical, the fused kernels reproduce the op-by-op bits, and the error norm is an exactly rounded sum on both sides:
the GPU loop takes the SAME steps as the CPU run and every output is equal bit for bit (Goldbeter: libm pow differs from the
device's single-rounded x^4 by an ulp on a few points, so <= 1 x (rtol |y| + atol), the north-star bound).  That holds for all
three ways of driving the method on the device: op by op, host-driven with fused kernels, and the device-resident loop (whose
step controller evaluates x^y through the host controller's own sequence of IEEE operations, host/crd_pow.h).  nst / nfe / netf
reported."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P_, D_, L_ = C.c_void_p, C.c_double, C.c_long


def cpu_trajectory(oracle, Pm, y0, touts, rtol, atol):
    K = oracle.lib()
    K.N_VMake_Parallel.restype = P_
    K.N_VMake_Parallel.argtypes = [C.c_int, L_, L_, P_]
    K.ARKodeCreate.restype = P_
    K.ARKodeInit.argtypes = [P_, P_, P_, D_, P_]
    K.ARKodeSStolerances.argtypes = [P_, D_, D_]
    K.ARKodeSetUserData.argtypes = [P_, P_]
    K.ARKodeSetMaxNumSteps.argtypes = [P_, L_]
    K.ARKode.argtypes = [P_, D_, P_, C.POINTER(D_), C.c_int]
    K.ARKodeFree.argtypes = [C.POINTER(P_)]
    K.ARKodeGetNumSteps.argtypes = [P_, C.POINTER(L_)]
    K.ARKodeGetNumRhsEvals.argtypes = [P_, C.POINTER(L_), C.POINTER(L_)]
    K.ARKodeGetNumErrTestFails.argtypes = [P_, C.POINTER(L_)]
    y = y0.copy()
    Y = K.N_VMake_Parallel(0, y.size, y.size, y.ctypes.data)
    mem = P_(K.ARKodeCreate())
    assert K.ARKodeInit(mem, C.cast(K.crd_oracle_f, P_), None, 0.0, Y) == 0
    K.ARKodeSStolerances(mem, rtol, atol)
    K.ARKodeSetUserData(mem, C.cast(C.pointer(Pm), P_))
    K.ARKodeSetMaxNumSteps(mem, 200000)
    t = D_()
    outs = []
    for tout in touts:
        assert K.ARKode(mem, tout, Y, C.byref(t), 1) == 0
        outs.append(y.copy())
    nst, nfe, nfi, netf = L_(), L_(), L_(), L_()
    K.ARKodeGetNumSteps(mem, C.byref(nst)); K.ARKodeGetNumRhsEvals(mem, C.byref(nfe), C.byref(nfi))
    K.ARKodeGetNumErrTestFails(mem, C.byref(netf))
    K.ARKodeFree(C.byref(mem))
    cpu_trajectory.netf = netf.value
    return outs, nst.value, nfe.value


def reference_ics(model, nx, ny, beta):
    """FHNmodel_torus.cpp:285-354 / GoldbeterModel_torus.cpp:313-414 for varyBeta = 0, waveInside = 1."""
    dx, dy = 2 * np.pi / (nx - 1), 2 * np.pi / (ny - 1)
    xx = (np.arange(nx) * dx)[None, :]
    yy = (np.arange(ny) * dy)[:, None]
    wl, ww = 2 * np.pi * 0.1, 2 * np.pi * 0.5
    box = (xx >= np.pi - ww / 2) & (xx <= np.pi + ww / 2) & (yy >= wl) & (yy <= 2 * wl)
    y = np.empty((ny, nx, 2))
    if model == "fhn_torus":
        us, vs = -beta, beta ** 3 - 3 * beta
        y[..., 0] = np.where(box, us + 2, us); y[..., 1] = np.where(box, vs + 1.5, vs)
        return y.ravel(), (us, vs)
    zs, ys = 0.392, 1.6469     # near the Goldbeter steady state for beta = 0.4
    y[..., 0] = np.where(box, zs + 1, zs); y[..., 1] = np.where(box, ys + 1, ys)
    return y.ravel(), (zs, ys)


@pytest.mark.parametrize("model,nx,ny,touts", [("fhn_torus", 32, 128, [0.5, 1.0, 2.0, 4.0]), ("gb_torus", 24, 96, [0.05, 0.1, 0.2])])
@pytest.mark.parametrize("mode", ["resident", "fused", "opbyop"])
def test_trajectory_parity(crd, ctx, oracle, model, nx, ny, touts, mode):
    # resident: the whole step loop in one persistent kernel; fused: one launch per stage + fused finish, host loop;
    # opbyop: the SUNDIALS 2.x vector-op sequence
    fused, resident = mode != "opbyop", mode == "resident"
    rtol, atol = 1e-5, 1e-10
    beta = 1.25 if model == "fhn_torus" else 0.4
    y0, (s0, s1) = reference_ics(model, nx, ny, beta)
    Pm = oracle.make_params(model, nx, ny, beta=beta, vary_beta=0, t_boundary=1.0)
    cpu, nst_c, nfe_c = cpu_trajectory(oracle, Pm, y0, touts, rtol, atol)

    grid = crd.Grid(ctx, crd.make_params(model, nx, ny, beta=beta, vary_beta=0, t_boundary=1.0, arith=crd.ARITH_EXACT))
    y = grid.new_vector()
    grid.fill_initial_conditions(y, 0.1, 0.5, 1, s0, s1)
    assert y.to_numpy().tobytes() == y0.tobytes()          # device IC generator == the reference's IC loop
    solver = crd.ARKodeSolver(grid, y, rtol=rtol, atol=atol, fused=fused, resident=resident)
    for tout, want in zip(touts, cpu):
        flag, t = solver.ARKode(tout)
        assert flag == 0 and t == tout
        got = y.to_numpy()
        if model == "fhn_torus":
            assert got.tobytes() == want.tobytes(), (model, mode, tout, np.abs(got - want).max())
        else:
            assert np.all(np.abs(got - want) <= 1.0 * (rtol * np.abs(want) + atol)), (model, tout, np.abs(got - want).max())
    st = solver.stats()
    print("\n%s %s: GPU nst=%d nfe=%d netf=%d | CPU nst=%d nfe=%d netf=%d" % (model, mode, st["nst"], st["nfe"], st["netf"], nst_c, nfe_c, cpu_trajectory.netf))
    if model == "fhn_torus":
        assert st["nst"] == nst_c and st["netf"] == cpu_trajectory.netf
        assert st["nfe"] == (nfe_c if mode == "opbyop" else nfe_c - st["nst_attempts"])   # the fused loop reuses f(tn, yn) as stage 1
    else:
        assert abs(st["nst"] - nst_c) <= 1
    assert grid.resident_launches == (len(touts) if resident else 0)
    solver.free(); grid.close()


def test_reuse_first_stage_is_bitwise_neutral(crd, ctx):
    nx, ny = 32, 128
    y0, (s0, s1) = reference_ics("fhn_torus", nx, ny, 1.25)
    res = []
    for reuse in (False, True):
        grid = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, beta=1.25, vary_beta=0, t_boundary=0.0))
        y = crd.NVector.from_numpy(ctx, y0)
        s = crd.ARKodeSolver(grid, y, reuse_first_stage=reuse, resident=False)
        assert s.ARKode(1.0)[0] == 0
        res.append((y.to_numpy(), s.stats()))
        s.free(); grid.close()
    assert res[0][0].tobytes() == res[1][0].tobytes()
    assert res[0][1]["nst"] == res[1][1]["nst"] and res[1][1]["nfe"] < res[0][1]["nfe"]


def test_fused_stage_rhs_does_not_change_the_trajectory(crd, ctx):
    """fused="full" (stage assembly inside the RHS kernel) and fused="ops" (stage states materialised by
    lincomb_kernel) run the same arithmetic: identical bits, identical counters, fewer kernel launches."""
    nx, ny = 48, 192
    y0, _ = reference_ics("fhn_torus", nx, ny, 1.25)
    res = []
    for mode in ("ops", "full"):
        grid = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, beta=1.25, vary_beta=0, t_boundary=0.5))
        y = crd.NVector.from_numpy(ctx, y0)
        l0 = ctx.launches
        s = crd.ARKodeSolver(grid, y, fused=mode, resident=False)
        assert s.ARKode(1.5)[0] == 0
        res.append((y.to_numpy(), s.stats(), ctx.launches - l0))
        s.free(); grid.close()
    assert res[0][0].tobytes() == res[1][0].tobytes()
    assert res[0][1]["nst"] == res[1][1]["nst"] and res[0][1]["nfe"] == res[1][1]["nfe"]
    assert res[1][2] < res[0][2]


def test_fused_last_stage_finish_in_the_integrator(crd, ctx):
    """Large single-GPU meshes: the host-driven loop issues the last stage and the finish as one kernel: the same bits as with
    the two separate kernels at every step (EXACT arithmetic: the error sum does not depend on the order of summation), and the
    bits of the CPU run of the same driver (op-by-op vector operations, the reference's f() restated in C)."""
    nx, ny = 512, 2304      # > 1 Mi points: the fused kernel applies; > 4 Mi would be needed for nothing else
    res = []
    for sf in (False, True):
        grid = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, t_boundary=0.0))
        grid.set_resident(-1)
        y = grid.new_vector()
        grid.fill_synthetic(y)
        s = crd.ARKodeSolver(grid, y, fused="full", resident=False, stage_finish=sf)
        s.set_init_step(1e-6)
        flag, t = s.ARKode(1.0, crd.ARK_ONE_STEP)
        assert flag == 0 and t == 1e-6
        first = y.to_numpy()
        for _ in range(5):
            assert s.ARKode(1.0, crd.ARK_ONE_STEP)[0] == 0
        res.append((first, y.to_numpy(), s.stats()))
        s.free(); grid.close()
    assert res[0][0].tobytes() == res[1][0].tobytes()
    assert res[0][2]["nst"] == res[1][2]["nst"] == 6 and res[0][2]["nfe"] == res[1][2]["nfe"]
    assert res[0][1].tobytes() == res[1][1].tobytes()


def test_large_mesh_fused_loop_takes_the_steps_of_the_cpu_run(crd, ctx, oracle):
    """512 x 2304 (the streaming kernels: fused stages, last stage + finish in one pass) from the synthetic state, ARK_NORMAL to a
    tout a dozen steps away: the GPU's output, nst and netf equal the CPU run's bit for bit."""
    nx, ny = 512, 2304
    tout = 4e-6
    Pm = oracle.make_params("fhn_torus", nx, ny, t_boundary=0.0)
    y0 = oracle.fill_state("fhn_torus", 2 * nx * ny)
    cpu, nst_c, nfe_c = cpu_trajectory(oracle, Pm, y0, [tout], 1e-5, 1e-10)
    grid = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, t_boundary=0.0))
    grid.set_resident(-1)
    y = grid.new_vector()
    grid.fill_synthetic(y)
    s = crd.ARKodeSolver(grid, y, fused="full", resident=False)
    flag, t = s.ARKode(tout)
    st = s.stats()
    print("\nlarge mesh: GPU nst=%d nfe=%d netf=%d | CPU nst=%d nfe=%d netf=%d" % (st["nst"], st["nfe"], st["netf"], nst_c, nfe_c, cpu_trajectory.netf))
    assert flag == 0 and t == tout and nst_c >= 3
    assert st["nst"] == nst_c and st["netf"] == cpu_trajectory.netf
    assert y.to_numpy().tobytes() == cpu[0].tobytes()
    s.free(); grid.close()


def test_accepted_state_and_next_second_stage_in_one_pass(crd, ctx):
    """FAST grid beyond 2 Mi points: after every accepted step f(tn, ynew) and the next step's second stage come from one pass
    over ynew (crd_fused_ops.rhs_pair).  Same steps and the same number of evaluations per step as the loop that issues them one
    by one (one more at the very end: the stage prepared for a step that is never taken), states equal to rounding; the device-
    side count of evaluations shows that the pass really ran."""
    nx, ny = 1030, 2100
    out = {}
    for pair in (True, False):
        g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, arith=crd.ARITH_FAST, vary_beta=0, t_boundary=0.0))
        y = g.new_vector()
        g.fill_initial_conditions(y, 0.1, 0.5, 1, -1.25, 1.25 ** 3 - 3 * 1.25)
        s = crd.ARKodeSolver(g, y, fused="full", resident=False, stage_pair=pair, max_steps=12)
        flag, t = s.ARKode(1e9)
        assert flag == -1          # the step limit
        st = s.stats()
        out[pair] = (y.to_numpy(), st, t)
        s.free(); g.close()
    (ya, sa, ta), (yb, sb, tb) = out[True], out[False]
    assert sa["nst"] == sb["nst"] == 12 and sa["netf"] == sb["netf"] and abs(ta - tb) <= 1e-9 * abs(tb)
    assert sa["nfe"] == sb["nfe"] + 1
    assert np.abs(ya - yb).max() <= 1e-11 * (1.0 + np.abs(yb).max())


def test_one_pass_pair_keeps_exact_trajectories_bit_identical(crd, ctx):
    """EXACT FHN grid of a few million points, several ARKode calls with dense output in between (the prepared second stage
    survives the return to the caller), an error-test failure on the way (the prepared stage is discarded and evaluated again for
    the smaller step): states returned at every output time, step counts and the final state are bit-identical with and without
    the pass."""
    nx, ny = 1030, 1100
    out = {}
    for pair in (True, False):
        g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, vary_beta=0, t_boundary=0.0))
        y = g.new_vector()
        g.fill_initial_conditions(y, 0.1, 0.5, 1, -1.25, 1.25 ** 3 - 3 * 1.25)
        s = crd.ARKodeSolver(g, y, fused="full", resident=False, stage_pair=pair)
        s.set_init_step(2e-2)        # too large for the front: the first attempts fail the error test
        res = []
        for tout in (0.05, 0.11, 0.2):
            flag, t = s.ARKode(tout)
            assert flag == 0 and t == tout
            res.append(y.to_numpy().copy())
        st = s.stats()
        out[pair] = (res, st)
        s.free(); g.close()
    (ra, sa), (rb, sb) = out[True], out[False]
    assert sa["nst"] == sb["nst"] and sa["netf"] == sb["netf"] and sa["netf"] > 0 and sa["nst_attempts"] == sb["nst_attempts"]
    for a, b in zip(ra, rb):
        assert a.tobytes() == b.tobytes()
    assert sb["nfe"] <= sa["nfe"] <= sb["nfe"] + 1
