"""The whole phi-split path on ONE GPU: R ranks = R host threads, each with its own context (stream), grid slab and
integrator, wired exactly like R processes on R GPUs — boundary rows pushed into the neighbours' ghost rows from inside the
evaluation's launch, the integrator's norms exchanged between the contexts by the device-side allreduce (mailboxes), nothing
crossing the host.  The kernels of the ranks run concurrently on the one device and really wait for each other.  With EXACT
arithmetic the split must not change a bit: every rank's slab of the final state equals the single-rank run's, which equals
the CPU run of the same driver with the reference's f() (tests/test_integrate_gpu.py)."""
import math
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_ranks(nr, body):
    """body(rank, barrier) on nr threads; returns the list of results, re-raises the first exception"""
    out, err = [None] * nr, []
    bar = threading.Barrier(nr)

    def wrap(r):
        try:
            out[r] = body(r, bar)
        except BaseException as e:   # noqa: BLE001
            err.append(e)
            bar.abort()
    th = [threading.Thread(target=wrap, args=(r,)) for r in range(nr)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if err:
        raise err[0]
    return out


@pytest.mark.parametrize("nr", [2, 3])
def test_reductions_are_exchanged_on_the_device_and_identical_on_every_rank(crd, oracle, nr):
    n = 200003
    rng = np.random.default_rng(nr)
    x, w = rng.standard_normal(n) * 10.0 ** rng.integers(-6, 3, n), rng.random(n) + 0.5
    cuts = [n * r // nr for r in range(nr + 1)]
    ctxs = [crd.Context(0) for _ in range(nr)]

    def body(r, bar):
        c = ctxs[r]
        c.set_halo_timeout(20000.0)
        c.comm_connect_local(r, ctxs)
        bar.wait()
        X = crd.NVector.from_numpy(c, x[cuts[r]:cuts[r + 1]], n)
        W = crd.NVector.from_numpy(c, w[cuts[r]:cuts[r + 1]], n)
        res = []
        for _ in range(3):                      # several reductions: the mailboxes' double buffering
            res.append((crd.N_VWrmsNorm(X, W), crd.N_VMaxNorm(X), crd.N_VMin(X), crd.N_VWL2Norm(X, W), crd.N_VDotProd(X, W)))
        return res
    out = run_ranks(nr, body)
    s2 = math.fsum((x * w) ** 2)
    for r in range(nr):
        for wr, mx, mn, wl2, dot in out[r]:
            assert wr == math.sqrt(s2 / n) and wl2 == math.sqrt(s2)     # exactly rounded, whatever the split
            assert mx == np.abs(x).max() and mn == x.min()
            assert dot == out[0][0][4] and abs(dot - math.fsum(x * w)) <= 1e-12 * math.fsum(np.abs(x * w))
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("model,nx,ny,tout,nr", [("fhn_torus", 64, 256, 1.5, 2), ("fhn_torus", 64, 256, 1.5, 3), ("gb_torus", 48, 192, 0.1, 2)])
def test_split_integration_equals_the_single_rank_run_bit_for_bit(crd, oracle, model, nx, ny, tout, nr):
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_integrate_gpu import reference_ics
    beta = 1.25 if model == "fhn_torus" else 0.4
    y0, (s0, s1) = reference_ics(model, nx, ny, beta)
    kw = dict(beta=beta, vary_beta=0, t_boundary=0.5)
    # single rank
    c1 = crd.Context(0)
    g1 = crd.Grid(c1, crd.make_params(model, nx, ny, **kw))
    y1 = crd.NVector.from_numpy(c1, y0)
    s = crd.ARKodeSolver(g1, y1, resident=False)
    assert s.ARKode(tout)[0] == 0
    one, st1 = y1.to_numpy(), s.stats()
    s.free(); g1.close(); c1.close()
    # nr ranks on the same GPU
    ctxs = [crd.Context(0) for _ in range(nr)]
    grids = [None] * nr

    def body(r, bar):
        c = ctxs[r]
        c.set_halo_timeout(30000.0)
        c.comm_connect_local(r, ctxs)
        js, je = crd.decomp_phi(ny, nr, r)
        grids[r] = crd.Grid(c, crd.make_params(model, nx, ny, js=js, je=je, **kw))
        bar.wait()
        grids[r].halo_connect_local(grids[(r - 1) % nr], grids[(r + 1) % nr])
        bar.wait()
        y = crd.NVector.from_numpy(c, y0[2 * nx * js: 2 * nx * (je + 1)], 2 * nx * ny)
        sv = crd.ARKodeSolver(grids[r], y, resident=False)
        flag, t = sv.ARKode(tout)
        res = (flag, y.to_numpy(), sv.stats())
        bar.wait()
        sv.free()
        return res
    out = run_ranks(nr, body)
    assert all(o[0] == 0 for o in out)
    assert np.concatenate([o[1] for o in out]).tobytes() == one.tobytes()
    for o in out:
        assert o[2]["nst"] == st1["nst"] and o[2]["netf"] == st1["netf"] and o[2]["nfe"] == st1["nfe"]
    for g in grids:
        g.close()
    for c in ctxs:
        c.close()
