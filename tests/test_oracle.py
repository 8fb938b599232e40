"""Pins the CPU checker (oracle/): restatement vs the reference compiled in place (oracle/_ref) and vs
the committed golden vectors generated from it, plus the analytic invariants of SURVEY.md §4."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rhs_golden.npz")
NAMES = {0: "fhn_torus", 1: "gb_torus", 2: "fhn_flat", 3: "gb_flat"}


def golden_cases():
    g = np.load(GOLDEN)
    for k, row in enumerate(g["meta"]):
        yield k, int(row[0]), int(row[1]), int(row[2]), float(row[3]), int(row[4]), int(row[5]), int(row[6]), g["ydot_%03d" % k]


def test_restatement_matches_golden_bitwise(oracle):
    n = 0
    for k, m, nx, ny, t, vb, jd, seed, ref in golden_cases():
        P = oracle.make_params(m, nx, ny, vary_beta=vb, just_diffusion=jd, t_boundary=38.0)
        y = oracle.fill_state(m, 2 * nx * ny, seed=seed)
        got = oracle.rhs(P, t, y)
        assert got.tobytes() == ref.tobytes(), (k, NAMES[m], nx, ny, t, vb, jd)
        n += 1
    assert n == 72


@pytest.mark.parametrize("model", ["fhn_torus", "gb_torus", "fhn_flat", "gb_flat"])
def test_restatement_matches_reference_bitwise(oracle, model):
    if not oracle.ref_available(model):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for (nx, ny) in ((48, 96), (5, 4), (100, 400)):
        for t in (10.0, 50.0):
            P = oracle.make_params(model, nx, ny, t_boundary=38.0)
            y = oracle.fill_state(model, 2 * nx * ny, seed=7)
            ref, _ = oracle.ref_rhs(P, t, y)
            assert oracle.rhs(P, t, y).tobytes() == ref.tobytes()


def test_reference_multirank_matches_survey(oracle):
    """the reference's exchange is only right with <= 2 ranks per dimension."""
    if not oracle.ref_available("fhn_torus"):
        pytest.skip("oracle/_ref not built")
    P = oracle.make_params("fhn_torus", 48, 96)
    y = oracle.fill_state("fhn_torus", 2 * 48 * 96)
    one, _ = oracle.ref_rhs(P, 50.0, y)
    expect = {2: 0, 4: 0, 3: 576, 8: 768}
    for nr, bad in expect.items():
        got, _ = oracle.ref_rhs(P, 50.0, y, nranks=nr)
        assert int((got != one).sum()) == bad


def test_golden_decomposition_formula(oracle):
    """SetupDecomp extents (FHNmodel_torus.cpp:750-755) as recorded from the reference."""
    g = np.load(GOLDEN)["decomp_400x1600"]
    for nr, r, is_, ie, js, je, d0, d1 in g:
        c0, c1 = r // d1, r % d1
        assert is_ == 400 * c0 // d0 and ie == 400 * (c0 + 1) // d0 - 1
        assert js == 1600 * c1 // d1 and je == 1600 * (c1 + 1) // d1 - 1
        assert d0 * d1 == nr and d0 >= d1


def test_fill_state_jump_ahead(oracle):
    a = oracle.fill_state("fhn_torus", 1000)
    b = oracle.fill_state("fhn_torus", 300, first_elem=700)
    assert a[700:].tobytes() == b.tobytes()
    assert a.min() >= -2.0 and a.max() < 2.0
    g = oracle.fill_state("gb_torus", 1000)
    assert g.min() >= 0.1 and g.max() < 1.6


@pytest.mark.parametrize("model", ["fhn_torus", "fhn_flat"])
def test_fhn_steady_state_is_stationary(oracle, model):
    beta = 1.25
    nx, ny = 40, 64
    P = oracle.make_params(model, nx, ny, beta=beta, vary_beta=0, t_boundary=0.0)
    y = np.empty((ny, nx, 2))
    y[..., 0] = -beta
    y[..., 1] = beta ** 3 - 3 * beta
    d = oracle.rhs(P, 1.0, y)
    assert np.abs(d).max() < 1e-13


@pytest.mark.parametrize("model", ["gb_torus", "gb_flat"])
def test_goldbeter_conservation_identity(oracle, model):
    """Z' + Y' = v0 + v1*beta - k*Z on a uniform field (diffusion of a constant is 0)."""
    nx, ny = 16, 24
    P = oracle.make_params(model, nx, ny, beta=0.4, vary_beta=0)
    y = np.empty((ny, nx, 2))
    y[..., 0] = 0.37
    y[..., 1] = 1.1
    d = oracle.rhs(P, 1.0, y).reshape(ny, nx, 2)
    np.testing.assert_allclose(d[..., 0] + d[..., 1], 1.0 + 7.3 * 0.4 - 10.0 * 0.37, rtol=0, atol=1e-12)


def test_freeze_rows(oracle):
    nx, ny = 12, 9
    P = oracle.make_params("fhn_torus", nx, ny, t_boundary=38.0)
    y = oracle.fill_state("fhn_torus", 2 * nx * ny)
    before = oracle.rhs(P, 10.0, y).reshape(ny, nx, 2)
    after = oracle.rhs(P, 50.0, y).reshape(ny, nx, 2)
    assert np.all(before[0] == 0) and np.all(before[-1] == 0)
    assert np.array_equal(before[1:-1], after[1:-1])
    assert np.any(after[0] != 0)


def test_torus_operator_second_order(oracle):
    """Laplace-Beltrami of cos(m phi) is -m^2/(R + r cos theta)^2 cos(m phi): discrete error is O(h^2)."""
    errs = []
    for nx in (32, 64):
        ny = 4 * nx
        P = oracle.make_params("gb_torus", nx, ny, diff=1.0, just_diffusion=1)
        dx, dy = 2 * np.pi / (nx - 1), 2 * np.pi / (ny - 1)
        th = np.arange(nx) * dx
        ph = np.arange(ny) * dy
        R, r = 80 / (2 * np.pi), 20 / (2 * np.pi)
        m = 2
        y = np.zeros((ny, nx, 2))
        y[..., 0] = np.cos(m * ph)[:, None]
        d = oracle.rhs(P, 0.0, y).reshape(ny, nx, 2)[..., 0]
        exact = -m * m / (R + r * np.cos(th))[None, :] ** 2 * np.cos(m * ph)[:, None]
        errs.append(np.abs(d[2:-2] - exact[2:-2]).max())   # away from the (n-1)-spacing seam
    assert errs[1] < errs[0] / 3.0


def test_restatement_matches_reference_digests_at_baseline_meshes(oracle):
    """BASELINE configs[0..4] mesh sizes: the checker against SHA-256 digests / sampled values of the reference's own f()
    (tests/golden/rhs_baseline_digests.json + rhs_baseline_samples.npz, generated from oracle/_ref by
    tests/golden/make_baseline_digests.py).  The two headline meshes (2.7e8 points, ~10-30 s of one core each) are checked
    through their sampled values: the row of each of 48 samples is recomputed in band form."""
    import hashlib
    import json
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    cases = json.load(open(os.path.join(here, "rhs_baseline_digests.json")))
    samples = np.load(os.path.join(here, "rhs_baseline_samples.npz"))
    assert len(cases) >= 13
    for c in cases:
        nx, ny = c["nx"], c["ny"]
        if "sample_set" in c:
            P = oracle.make_params(c["model"], nx, ny, just_diffusion=c["just_diffusion"], t_boundary=38.0)
            el, val = samples["idx_" + c["sample_set"]], samples["val_" + c["sample_set"]]
            pick = np.concatenate([np.arange(0, 8), np.arange(len(el) - 8, len(el)), np.linspace(8, len(el) - 9, 32).astype(int)])
            for k in pick:
                j, off = divmod(int(el[k]), 2 * nx)
                row = oracle.rhs_band(P, c["t"], j, 1, oracle.band_state(c["model"], nx, ny, j, 1, seed=c["seed"]))
                assert row[off] == val[k], (c["name"], k)
            continue
        y = oracle.fill_state(c["model"], 2 * nx * ny, seed=c["seed"])
        got = oracle.rhs(oracle.make_params(c["model"], nx, ny, just_diffusion=c["just_diffusion"], t_boundary=38.0), c["t"], y)
        assert hashlib.sha256(got.tobytes()).hexdigest() == c["sha256"], c["name"]      # same libm here: bit for bit, pow included
        for i, hx in c["samples"].items():
            assert got[int(i)] == float.fromhex(hx)


def test_band_form_equals_the_whole_mesh(oracle):
    """crd_oracle_rhs_band / crd_ref_rhs_band (the reference's f() and Exchange() on the rows one rank of a phi split owns, the
    neighbours' rows delivered as messages) reproduce the same rows of the np = 1 result bit for bit: bands at the global
    boundary rows (frozen while t < tBoundary, periodic wrap), in the interior, and the whole mesh."""
    nx, ny = 37, 29
    for model in ("fhn_torus", "gb_torus", "fhn_flat", "gb_flat"):
        P = oracle.make_params(model, nx, ny, t_boundary=38.0)
        y = oracle.fill_state(model, 2 * nx * ny)
        for t in (10.0, 50.0):
            ref = oracle.rhs(P, t, y).reshape(ny, 2 * nx)
            for j0, n in ((0, 2), (0, 3), (5, 4), (ny - 2, 2), (ny - 3, 3), (0, ny), (4, 1), (0, 1), (ny - 1, 1)):
                yb = oracle.band_state(model, nx, ny, j0, n)
                assert oracle.rhs_band(P, t, j0, n, yb).tobytes() == ref[j0:j0 + n].tobytes(), (model, t, j0, n)
                if n >= 2 and oracle.ref_available(model):      # the reference's face / corner code needs two rows
                    assert oracle.ref_rhs_band(P, t, j0, n, yb).tobytes() == ref[j0:j0 + n].tobytes(), (model, t, j0, n)
