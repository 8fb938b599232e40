"""The per-point arithmetic of the CUDA kernels, executed on the host.

crd_rhs_point.cuh (stencil_exact, react, the reciprocal-refinement division, pow4_rn) and crd_tables.hpp (geometry, scalars,
per-theta / per-phi tables) are the product's own source; tests/point_harness.cpp compiles them for the CPU with shims for the
intrinsics (each operation separately rounded) and sweeps a periodic grid.  Against the reference's f() (golden vectors
generated from oracle/_ref, and the checker on more shapes) the result must be what the GPU tests demand of the kernels:
bit-identical for FHN and for the diffusion-only Goldbeter run, <= 4e-16 of the summed terms where libm's pow differs from
the single-rounded x^4.  No GPU and no library needed: an arithmetic regression shows up here first."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "rhs_golden.npz")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = tmp_path_factory.mktemp("point") / "point_harness.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I/usr/local/cuda/include",
                    "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "crdmodel_b200", "csrc"),
                    os.path.join(ROOT, "tests", "point_harness.cpp"), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    lib.point_harness_rhs.restype = C.c_int
    lib.point_harness_rhs.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def host_rhs(harness, P, t, y):
    """P: the product's ctypes crd_params (crdmodel_b200.api.make_params)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.full_like(y, -777.0)
    react_on = 0 if (P.model in (1, 3) and P.just_diffusion) else 1
    assert harness.point_harness_rhs(C.byref(P), t, react_on, y.ctypes.data, out.ctypes.data, None) == 0
    return out


def scale_of(P, y):
    # magnitude of the summed terms (same bound as tests/test_rhs_gpu.py)
    ny, nx = P.ny, P.nx
    a = np.abs(y.reshape(ny, nx, 2))
    u = a[..., 0]
    nb = u + np.roll(u, 1, 0) + np.roll(u, -1, 0) + np.roll(u, 1, 1) + np.roll(u, -1, 1)
    torus = P.model in (0, 1)
    dx = (2 * np.pi if torus else P.surface_width) / (nx - 1)
    dy = (2 * np.pi if torus else P.surface_length) / (ny - 1)
    r2 = (P.surface_width / (2 * np.pi)) ** 2 if torus else 1.0
    Rm = (P.surface_length - P.surface_width) / (2 * np.pi) if torus else 1.0
    s = P.diff * nb * (4.0 / (r2 * dx * dx) + 4.0 / (Rm * Rm * dy * dy)) + 700.0 * (1 + a[..., 0] + a[..., 1])
    return np.repeat(s[..., None], 2, axis=2).ravel()


def test_device_point_source_reproduces_the_golden_vectors(harness, oracle):
    import crdmodel_b200.api as api      # parameter struct only; nothing is computed by the library here
    g = np.load(GOLDEN)
    exact_cases = 0
    for k, row in enumerate(g["meta"]):
        m, nx, ny, t, vb, jd, seed = int(row[0]), int(row[1]), int(row[2]), float(row[3]), int(row[4]), int(row[5]), int(row[6])
        ref = g["ydot_%03d" % k]
        y = oracle.fill_state(m, 2 * nx * ny, seed=seed)
        P = api.make_params(m, nx, ny, vary_beta=vb, just_diffusion=jd, t_boundary=38.0)
        got = host_rhs(harness, P, t, y)
        if m in (0, 2) or jd == 1:
            assert got.tobytes() == ref.tobytes(), (k, m, nx, ny, t, vb, jd)
            exact_cases += 1
        else:
            assert np.all(np.abs(got - ref) <= 4e-16 * scale_of(P, y)), (k, m)
            assert (got != ref).mean() < 0.05
    assert exact_cases >= 4


@pytest.mark.parametrize("model", ["fhn_torus", "gb_torus", "fhn_flat", "gb_flat"])
def test_device_point_source_against_the_checker(harness, oracle, model):
    import crdmodel_b200.api as api
    for (nx, ny) in ((100, 400), (3, 2), (2, 3), (257, 31), (31, 257)):
        for t in (10.0, 50.0):          # frozen boundary rows / released
            y = oracle.fill_state(model, 2 * nx * ny, seed=11 + nx)
            ref = oracle.rhs(oracle.make_params(model, nx, ny, t_boundary=38.0), t, y)
            P = api.make_params(model, nx, ny, t_boundary=38.0)
            got = host_rhs(harness, P, t, y)
            if model.startswith("fhn"):
                assert got.tobytes() == ref.tobytes(), (model, nx, ny, t)
            else:
                assert np.all(np.abs(got - ref) <= 4e-16 * scale_of(P, y)), (model, nx, ny, t)
    if model.startswith("gb"):          # diffusion only: no pow, bit for bit
        nx, ny = 64, 96
        y = oracle.fill_state(model, 2 * nx * ny, seed=3)
        ref = oracle.rhs(oracle.make_params(model, nx, ny, just_diffusion=1), 50.0, y)
        got = host_rhs(harness, api.make_params(model, nx, ny, just_diffusion=1), 50.0, y)
        assert got.tobytes() == ref.tobytes()


def test_exact_division_edge_values_on_the_host(harness, oracle):
    """States that drive the three numerators of a point to zero, to tiny / huge magnitudes and to non-finite values: the guarded
    out-of-line IEEE path must take over (same cases as tests/test_rhs_gpu.py::test_exact_division_edge_values)."""
    import crdmodel_b200.api as api
    nx, ny = 16, 12
    rng = np.random.default_rng(1)
    for model in ("fhn_torus", "gb_torus"):
        for mag in (0.0, 1e-300, 1e-250, 1e-160, 1.0, 1e150, 1e250, 1e300):
            y = rng.uniform(-1.0, 1.0, 2 * nx * ny) * mag
            y[::7] = 0.0
            y[5::11] = -0.0
            ref = oracle.rhs(oracle.make_params(model, nx, ny, just_diffusion=1, t_boundary=0.0), 50.0, y) if model == "gb_torus" \
                else oracle.rhs(oracle.make_params(model, nx, ny, t_boundary=0.0), 50.0, y)
            P = api.make_params(model, nx, ny, just_diffusion=1 if model == "gb_torus" else 0, t_boundary=0.0)
            with np.errstate(all="ignore"):
                got = host_rhs(harness, P, 50.0, y)
            assert got.tobytes() == ref.tobytes(), (model, mag)


def test_device_point_source_at_baseline_meshes(harness, oracle):
    """BASELINE configs[0..2] mesh sizes against the digests of the reference's own f() (tests/golden/rhs_baseline_digests.json)."""
    import hashlib
    import json
    import crdmodel_b200.api as api
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "rhs_baseline_digests.json")))
    for c in cases:
        nx, ny = c["nx"], c["ny"]
        if nx * ny >= (1 << 26):      # the headline meshes: checked on the GPU (tests/test_rhs_gpu.py) and through samples (test_oracle.py)
            continue
        y = oracle.fill_state(c["model"], 2 * nx * ny, seed=c["seed"])
        P = api.make_params(c["model"], nx, ny, just_diffusion=c["just_diffusion"], t_boundary=38.0)
        got = host_rhs(harness, P, c["t"], y)
        if c["model"].startswith("fhn") or c["just_diffusion"] == 1:
            assert hashlib.sha256(got.tobytes()).hexdigest() == c["sha256"], c["name"]
        else:
            sc = scale_of(P, y)
            for i, hx in c["samples"].items():
                assert abs(got[int(i)] - float.fromhex(hx)) <= 4e-16 * sc[int(i)], (c["name"], i)


def test_beta_table_of_a_phi_split_carries_the_neighbours_rows(harness):
    """grid_host_tables of a rank of a phi split: brow[1 + j] is local row j, brow[0] / brow[nyl + 1] the rows just south / north of
    the slab in the periodic global mesh — the very values the owning ranks hold for them (the pass that forms two evaluations at
    once evaluates f on those rows, src/FHNmodel_torus.cpp:623,631: beta(phi) from the GLOBAL row index)."""
    import crdmodel_b200.api as api
    harness.point_harness_brow.restype = C.c_int
    harness.point_harness_brow.argtypes = [C.c_void_p, C.c_void_p]
    for model in ("fhn_torus", "gb_flat", "gb_torus"):
        nx, ny = 40, 101
        whole = np.zeros(ny + 2)
        P = api.make_params(model, nx, ny, vary_beta=1)
        assert harness.point_harness_brow(C.byref(P), whole.ctypes.data) == ny + 2
        assert whole[0] == whole[ny] and whole[ny + 1] == whole[1]            # one slab: its own last / first row
        if model != "gb_torus":                                                # (the Goldbeter torus program ignores varyBeta = 1)
            assert len(set(whole[1:ny + 1])) == ny                             # beta really varies with phi
        for nr in (2, 3, 7):
            for r in range(nr):
                js, je = ny * r // nr, ny * (r + 1) // nr - 1
                nyl = je - js + 1
                part = np.zeros(nyl + 2)
                Pr = api.make_params(model, nx, ny, js=js, je=je, vary_beta=1)
                assert harness.point_harness_brow(C.byref(Pr), part.ctypes.data) == nyl + 2
                assert part[1:nyl + 1].tobytes() == whole[1 + js:2 + je].tobytes()
                assert part[0] == whole[1 + (js - 1) % ny] and part[nyl + 1] == whole[1 + (je + 1) % ny]
