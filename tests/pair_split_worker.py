"""Worker of tests/test_rhs_gpu.py::test_two_evaluations_in_one_pass_on_a_phi_split (a process of its own: see there).
python tests/pair_split_worker.py MODEL exact|fast"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crdmodel_b200 as crd  # noqa: E402
import oracle as O  # noqa: E402


def main():
    model, arith = sys.argv[1], sys.argv[2]
    nx, ny = 520, 6300
    ar = crd.ARITH_EXACT if arith == "exact" else crd.ARITH_FAST
    y = O.fill_state(model, 2 * nx * ny, seed=9)
    c = 0.5e-3
    times = ((10.0, 10.3), (37.9, 38.2), (50.0, 50.1), (10.0, 10.3))   # frozen boundary rows in both, one, neither evaluation
    c0 = crd.Context(0)
    g1 = crd.Grid(c0, crd.make_params(model, nx, ny, arith=ar, t_boundary=38.0))
    Y = crd.NVector.from_numpy(c0, y)
    W1, W2 = g1.new_vector(), g1.new_vector()
    want = {}
    for t1, t2 in times[:3]:
        g1.f(t1, Y, W1)
        g1.f_lincomb(t2, [1.0, c], [Y, W1], W2)
        want[t1] = (W1.to_numpy(), W2.to_numpy())
    assert g1.f_pair(50.0, 50.1, c, Y, W1, W2) == 0       # (the pass itself is loaded before any rank spins on another)
    c0.sync()
    g1.close()
    for nr in (2, 3):
        ctxs = [crd.Context(0) for _ in range(nr)]
        grids, ys, f1s, f2s = [], [], [], []
        for r in range(nr):
            ctxs[r].set_halo_timeout(20000.0)
            js, je = crd.decomp_phi(ny, nr, r)
            g = crd.Grid(ctxs[r], crd.make_params(model, nx, ny, js=js, je=je, arith=ar, t_boundary=38.0))
            grids.append(g)
            ys.append(crd.NVector.from_numpy(ctxs[r], y[2 * nx * js: 2 * nx * (je + 1)], 2 * nx * ny))
            f1s.append(g.new_vector()); f2s.append(g.new_vector())
        for r in range(nr):
            grids[r].halo_connect_local(grids[(r - 1) % nr], grids[(r + 1) % nr])
        for t1, t2 in times:
            rcs = [grids[r].f_pair(t1, t2, c, ys[r], f1s[r], f2s[r]) for r in range(nr)]    # asynchronous: all in flight
            for cx in ctxs:
                cx.sync()
            assert rcs == [0] * nr, rcs
            got1 = np.concatenate([v.to_numpy() for v in f1s]); got2 = np.concatenate([v.to_numpy() for v in f2s])
            w1, w2 = want[t1]
            if arith == "exact":
                assert got1.tobytes() == w1.tobytes() and got2.tobytes() == w2.tobytes(), (model, nr, t1)
            else:
                assert np.abs(got1 - w1).max() <= 1e-12 * (1 + np.abs(w1).max()), (model, nr, t1)
                assert np.abs(got2 - w2).max() <= 1e-12 * (1 + np.abs(w2).max()), (model, nr, t2)
        # an ordinary evaluation after it still finds its ghost rows (epochs stay in step)
        for r in range(nr):
            grids[r].f(50.0, ys[r], f1s[r])
        for cx in ctxs:
            cx.sync()
        got = np.concatenate([v.to_numpy() for v in f1s])
        if arith == "exact":
            assert got.tobytes() == want[50.0][0].tobytes()
        for g in grids:
            g.close()
        for cx in ctxs:
            cx.close()
    c0.close()
    print("PAIR SPLIT OK", model, arith)


if __name__ == "__main__":
    main()
