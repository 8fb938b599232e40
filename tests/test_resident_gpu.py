"""The device-resident step loop (crd_erk_evolve: stages, finish, error test and step controller inside one
persistent cooperative kernel) against the host-driven loop of crd_ark.cpp (one launch per stage, host round trip
per step), which the other tests pin to the CPU checker.  Same arithmetic per point, an error norm that does not depend on
the order of summation (double-double), and a step controller that evaluates x^y through the same sequence of IEEE operations on
the device as on the host (host/crd_pow.h): in EXACT arithmetic the two loops take the same steps and produce the same bits,
for a single step and over whole trajectories."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ARK_TOO_MUCH_WORK = -1
MODELS = ["fhn_torus", "gb_torus", "fhn_flat", "gb_flat"]


def smooth_state(model, nx, ny):
    """A smooth state with O(1) gradients, inside the physical range of the model."""
    x = np.linspace(0, 2 * np.pi, nx)[None, :]
    y = np.linspace(0, 2 * np.pi, ny)[:, None]
    s = np.empty((ny, nx, 2))
    if model.startswith("fhn"):
        s[..., 0] = -1.2 + 1.5 * np.sin(x) * np.cos(2 * y)
        s[..., 1] = -0.6 + 0.4 * np.cos(x + y)
    else:
        s[..., 0] = 0.6 + 0.3 * np.sin(x) * np.cos(2 * y)
        s[..., 1] = 1.4 + 0.4 * np.cos(x + y)
    return s.ravel()


def run(crd, ctx, model, nx, ny, y0, touts, resident, arith=None, itask=None, h0=None, t_boundary=0.0, max_steps=200000,
        variant=0):
    grid = crd.Grid(ctx, crd.make_params(model, nx, ny, vary_beta=0, t_boundary=t_boundary,
                                         arith=crd.ARITH_EXACT if arith is None else arith))
    grid.set_resident(1 if resident else -1)
    if variant:
        grid.set_variant(variant)
    y = crd.NVector.from_numpy(ctx, y0)
    s = crd.ARKodeSolver(grid, y, fused="full", resident=resident, max_steps=max_steps)
    if h0 is not None:
        s.set_init_step(h0)
    outs = []
    for tout in touts:
        flag, t = s.ARKode(tout, crd.ARK_NORMAL if itask is None else itask)
        outs.append((flag, t, y.to_numpy()))
    st = s.stats()
    st["resident_launches"] = grid.resident_launches
    s.free(); grid.close()
    return outs, st


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("arith", ["exact", "fast"])
@pytest.mark.parametrize("nx,ny", [(70, 53), (32, 9), (3, 4), (129, 40)])
def test_one_step_is_bit_identical_to_the_host_driven_step(crd, ctx, model, arith, nx, ny):
    """Given the step size, stage states, stage derivatives and the new state are the same bits (ragged strips, meshes
    narrower than a warp, fewer rows than a work unit)."""
    ar = crd.ARITH_EXACT if arith == "exact" else crd.ARITH_FAST
    y0 = smooth_state(model, nx, ny)
    h0 = 1e-4
    a, sa = run(crd, ctx, model, nx, ny, y0, [1.0], False, ar, crd.ARK_ONE_STEP, h0, t_boundary=0.5)
    b, sb = run(crd, ctx, model, nx, ny, y0, [1.0], True, ar, crd.ARK_ONE_STEP, h0, t_boundary=0.5)
    assert sb["resident_launches"] == 1 and sa["resident_launches"] == 0
    assert a[0][0] == 0 and b[0][0] == 0
    assert a[0][1] == b[0][1] == h0
    assert sa["nst"] == sb["nst"] == 1 and sa["netf"] == sb["netf"] == 0
    if arith == "exact":
        assert a[0][2].tobytes() == b[0][2].tobytes()
    else:   # FMA contraction is the compiler's choice per kernel: equal to rounding, not necessarily to the bit
        assert np.abs(a[0][2] - b[0][2]).max() <= 1e-13 * (1.0 + np.abs(a[0][2]).max())
    assert not np.array_equal(a[0][2], y0)


@pytest.mark.parametrize("variant", [120, 121, 122, 124, 126])
@pytest.mark.parametrize("nx,ny", [(100, 77), (40, 300), (700, 160)])
def test_kernel_variants_agree(crd, ctx, variant, nx, ny):
    """Where a vector lives (120 + n: at most n of the 7 storages in shared memory, the rest in their global arrays) is a
    tuning knob: it does not change a bit.  Meshes with fewer rows than SMs, several rows per band, one row per band."""
    y0 = smooth_state("fhn_torus", nx, ny)
    a, sa = run(crd, ctx, "fhn_torus", nx, ny, y0, [0.02], True, None, None, 1e-4, t_boundary=0.01)
    b, sb = run(crd, ctx, "fhn_torus", nx, ny, y0, [0.02], True, None, None, 1e-4, t_boundary=0.01, variant=variant)
    assert sa["nst"] == sb["nst"] > 3 and sa["nfe"] == sb["nfe"]
    assert a[0][2].tobytes() == b[0][2].tobytes()


@pytest.mark.parametrize("model", MODELS)
def test_trajectory_matches_the_host_driven_loop(crd, ctx, model):
    """Adaptive run across the boundary release (tBoundary inside the interval), several ARKode calls."""
    nx, ny = 48, 160
    rtol, atol = 1e-5, 1e-10
    y0 = smooth_state(model, nx, ny)
    touts = [0.05, 0.3, 0.6] if model.startswith("fhn") else [0.02, 0.05, 0.1]
    tb = touts[1] * 0.9
    a, sa = run(crd, ctx, model, nx, ny, y0, touts, False, t_boundary=tb)
    b, sb = run(crd, ctx, model, nx, ny, y0, touts, True, t_boundary=tb)
    assert sb["resident_launches"] == len(touts)
    for (fa, ta, ya), (fb, tb_, yb) in zip(a, b):
        assert fa == 0 and fb == 0 and ta == tb_
        assert ya.tobytes() == yb.tobytes(), np.abs(ya - yb).max()
    print("\n%s host-driven nst=%d nfe=%d netf=%d | resident nst=%d nfe=%d netf=%d" %
          (model, sa["nst"], sa["nfe"], sa["netf"], sb["nst"], sb["nfe"], sb["netf"]))
    assert sa["nst"] == sb["nst"] and sa["netf"] == sb["netf"] and sa["hlast"] == sb["hlast"]
    # neither loop re-evaluates f(tn, yn) as stage 1: s evaluations per attempt (the host loop also counts its set-up
    # evaluations: f(t0, y0) and the probes of the initial-step estimate)
    assert 0 <= sa["nfe"] - sb["nfe"] <= 12


def test_first_steps_track_the_host_loop_closely(crd, ctx):
    """Step by step (ARK_ONE_STEP): the same attempts, the same times, the same bits."""
    nx, ny = 64, 96
    y0 = smooth_state("fhn_torus", nx, ny)
    ra, rb = [], []
    for resident, res in ((False, ra), (True, rb)):
        grid = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, vary_beta=0, t_boundary=0.0))
        grid.set_resident(1 if resident else -1)
        y = crd.NVector.from_numpy(ctx, y0)
        s = crd.ARKodeSolver(grid, y, fused="full", resident=resident)
        for _ in range(12):
            flag, t = s.ARKode(10.0, crd.ARK_ONE_STEP)
            assert flag == 0
            res.append((t, s.stats()["nst_attempts"], y.to_numpy()))
        s.free(); grid.close()
    assert ra[0][0] == rb[0][0] and ra[0][2].tobytes() == rb[0][2].tobytes()   # first step: same h, same bits
    for (ta, na, ya), (tb, nb, yb) in zip(ra, rb):
        assert na == nb and ta == tb
        assert ya.tobytes() == yb.tobytes()


def test_max_steps_is_reported_like_arkode(crd, ctx):
    nx, ny = 40, 64
    y0 = smooth_state("fhn_torus", nx, ny)
    for resident in (False, True):
        out, st = run(crd, ctx, "fhn_torus", nx, ny, y0, [50.0], resident, max_steps=7)
        flag, t, y = out[0]
        assert flag == ARK_TOO_MUCH_WORK and 0.0 < t < 50.0 and st["nst"] == 7
        assert np.all(np.isfinite(y))


def test_default_meshes_use_the_resident_loop_and_large_ones_do_not(crd, ctx):
    y0 = smooth_state("gb_torus", 100, 400)
    grid = crd.Grid(ctx, crd.make_params("gb_torus", 100, 400))
    y = crd.NVector.from_numpy(ctx, y0)
    s = crd.ARKodeSolver(grid, y)
    assert s.ARKode(0.01)[0] == 0 and grid.resident_launches == 1
    s.free(); grid.close()
    # a band whose stage tile does not fit shared memory (14 rows x 1024 columns x 16 B = 229 KB): not applicable even when
    # asked for; the integrator carries on with the launch-per-stage loop by itself
    grid = crd.Grid(ctx, crd.make_params("fhn_torus", 1024, 2048))
    grid.set_resident(1)
    y = grid.new_vector()
    grid.fill_synthetic(y)
    s = crd.ARKodeSolver(grid, y)
    assert s.ARKode(1.0, crd.ARK_ONE_STEP)[0] == 0 and s.ARKode(1.0, crd.ARK_ONE_STEP)[0] == 0 and grid.resident_launches == 0
    assert s.stats()["nst"] == 2
    s.free(); grid.close()
    nx, ny = 2048, 2304    # > 1 Mi points: beyond L2, stays with the launch-per-stage path
    grid = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny))
    y = grid.new_vector()
    grid.fill_synthetic(y)
    s = crd.ARKodeSolver(grid, y)
    assert s.ARKode(1.0, crd.ARK_ONE_STEP)[0] == 0 and grid.resident_launches == 0
    s.free(); grid.close()


def test_default_ini_mesh_rate(crd, ctx):
    """The reference's default FHN mesh (400 x 1600): the same bits from both loops, and the point of the exercise — steps per
    second — printed for both."""
    nx, ny = 400, 1600
    rtol, atol = 1e-5, 1e-10
    y0 = smooth_state("fhn_torus", nx, ny)
    res = {}
    for resident in (False, True):
        t0 = time.time()
        out, st = run(crd, ctx, "fhn_torus", nx, ny, y0, [0.5], resident, t_boundary=0.2)
        dt = time.time() - t0
        res[resident] = (out[0], st, dt)
        print("\n400x1600 %s: nst=%d nfe=%d netf=%d  %.0f steps/s (incl. set-up)" %
              ("resident" if resident else "host-driven", st["nst"], st["nfe"], st["netf"], st["nst"] / dt))
    (fa, ta, ya), sa, _ = res[False]
    (fb, tb, yb), sb, _ = res[True]
    assert fa == 0 and fb == 0
    assert ya.tobytes() == yb.tobytes()
    assert sa["nst"] == sb["nst"] and sa["netf"] == sb["netf"]
