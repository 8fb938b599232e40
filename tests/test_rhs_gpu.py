"""Parity of the CUDA right-hand side (through the C ABI) with the CPU checker.

Bar: FHN (torus, flat) and Goldbeter-flat stencil in EXACT arithmetic are BIT-IDENTICAL to the
reference's f(); Goldbeter EXACT differs only where libm's pow(x,2|4) is not correctly rounded
(<= a few ulp of the summed terms); FAST arithmetic is within 1e-12 relative to the magnitude of the
summed terms (the north-star tolerance, SURVEY.md §7 "cancellation in the parity metric")."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rhs_golden.npz")
MODELS = ["fhn_torus", "gb_torus", "fhn_flat", "gb_flat"]


def gpu_rhs(crd, ctx, model, nx, ny, t, y, arith, variant=0, **kw):
    P = crd.make_params(model, nx, ny, arith=arith, **kw)
    g = crd.Grid(ctx, P)
    g.set_variant(variant)
    yv = crd.NVector.from_numpy(ctx, y)
    dv = g.new_vector()
    crd.N_VConst(-777.0, dv)
    g.f(t, yv, dv)
    out = dv.to_numpy()
    g.close()
    return out


def scale_of(oracle, P, y):
    """Per-point magnitude of the summed terms: |f| evaluated term-wise is bounded by f on |y| with the
    diffusion weights; a cheap upper bound is max(|ref|, local |y| stencil sum)."""
    ny, nx = P.ny, P.nx
    a = np.abs(y.reshape(ny, nx, 2))
    u = a[..., 0]
    nb = u + np.roll(u, 1, 0) + np.roll(u, -1, 0) + np.roll(u, 1, 1) + np.roll(u, -1, 1)
    dx = (2 * np.pi if P.model in (0, 1) else P.surface_width) / (nx - 1)
    dy = (2 * np.pi if P.model in (0, 1) else P.surface_length) / (ny - 1)
    r2 = (P.surface_width / (2 * np.pi)) ** 2 if P.model in (0, 1) else 1.0
    Rm = (P.surface_length - P.surface_width) / (2 * np.pi) if P.model in (0, 1) else 1.0
    s = P.diff * nb * (4.0 / (r2 * dx * dx) + 4.0 / (Rm * Rm * dy * dy)) + 700.0 * (1 + a[..., 0] + a[..., 1])
    return np.repeat(s[..., None], 2, axis=2).ravel()


def test_golden_vectors(crd, ctx, oracle):
    g = np.load(GOLDEN)
    for k, row in enumerate(g["meta"]):
        m, nx, ny, t, vb, jd, seed = int(row[0]), int(row[1]), int(row[2]), float(row[3]), int(row[4]), int(row[5]), int(row[6])
        ref = g["ydot_%03d" % k]
        y = oracle.fill_state(m, 2 * nx * ny, seed=seed)
        kw = dict(vary_beta=vb, just_diffusion=jd, t_boundary=38.0)
        got = gpu_rhs(crd, ctx, m, nx, ny, t, y, crd.ARITH_EXACT, **kw)
        P = oracle.make_params(m, nx, ny, **kw)
        if m in (0, 2) or jd == 1:
            assert got.tobytes() == ref.tobytes(), ("exact", k, m, nx, ny, t, vb, jd)
        else:
            assert np.all(np.abs(got - ref) <= 4e-16 * scale_of(oracle, P, y)), ("gb exact", k)
        fast = gpu_rhs(crd, ctx, m, nx, ny, t, y, crd.ARITH_FAST, **kw)
        assert np.all(np.abs(fast - ref) <= 1e-12 * scale_of(oracle, P, y)), ("fast", k, np.abs(fast - ref).max())


def gb_uniform_scale(P):
    """Upper bound of the summed |terms| of the Goldbeter torus RHS over the synthetic state (Z, Y in [0.1, 1.6))."""
    dx, dy = 2 * np.pi / (P.nx - 1), 2 * np.pi / (P.ny - 1)
    r2 = (P.surface_width / (2 * np.pi)) ** 2
    Rm = (P.surface_length - P.surface_width) / (2 * np.pi)
    return P.diff * 5 * 1.6 * (4.0 / (r2 * dx * dx) + 4.0 / (Rm * Rm * dy * dy)) + 700.0 * (1 + 1.6 + 1.6)


def test_baseline_mesh_digests(crd, ctx, oracle):
    """BASELINE configs[0..4] mesh sizes — the 16384 x 16384 FHN torus and the theta 8192 x phi 32768 Goldbeter torus the
    throughput numbers are quoted on included — against digests of the reference's own f() (tests/golden/
    rhs_baseline_digests.json + rhs_baseline_samples.npz, generated from oracle/_ref by make_baseline_digests.py): SHA-256 of the
    ydot bytes where the path is bit-exact (FHN flat / torus, diffusion-only Goldbeter), sampled values within 4e-16 of the
    summed terms for the Goldbeter kinetics.  Full-width bands are also forced through every kernel the automatic choice
    can pick for large slabs (variants 13, 15, 20, 21)."""
    import hashlib
    import json
    here = os.path.dirname(GOLDEN)
    cases = json.load(open(os.path.join(here, "rhs_baseline_digests.json")))
    samples = np.load(os.path.join(here, "rhs_baseline_samples.npz"))
    assert len(cases) >= 13
    big_seen = 0
    for c in cases:
        nx, ny, model = c["nx"], c["ny"], c["model"]
        kw = dict(just_diffusion=c["just_diffusion"], t_boundary=38.0)
        exact_bits = model.startswith("fhn") or c["just_diffusion"] == 1
        g = crd.Grid(ctx, crd.make_params(model, nx, ny, arith=crd.ARITH_EXACT, **kw))
        yv, dv = g.new_vector(), g.new_vector()
        g.fill_synthetic(yv, seed=c["seed"])       # the oracle's LCG stream, generated on the device (test_synthetic_state_matches_oracle_stream)
        for variant in [0] + list(c.get("variants", [])):
            g.set_variant(variant)
            crd.N_VConst(-777.0, dv)
            g.f(c["t"], yv, dv)
            got = dv.to_numpy()
            if exact_bits:
                assert hashlib.sha256(got.data).hexdigest() == c["sha256"], (c["name"], variant)
            if "sample_set" in c:
                big_seen += 1
                el, val = samples["idx_" + c["sample_set"]], samples["val_" + c["sample_set"]]
                if exact_bits:
                    assert got[el].tobytes() == val.tobytes(), c["name"]
                else:
                    tol = 4e-16 * gb_uniform_scale(oracle.make_params(model, nx, ny, **kw))
                    assert np.all(np.abs(got[el] - val) <= tol), (c["name"], np.abs(got[el] - val).max(), tol)
                    assert (got[el] != val).mean() < 0.05
            elif not exact_bits:
                y = oracle.fill_state(model, 2 * nx * ny, seed=c["seed"])
                sc = scale_of(oracle, oracle.make_params(model, nx, ny, **kw), y)
                for i, hx in c["samples"].items():
                    assert abs(got[int(i)] - float.fromhex(hx)) <= 4e-16 * sc[int(i)], (c["name"], i)
            del got
        yv.destroy(); dv.destroy(); g.close()
    assert big_seen >= 4


def test_band_of_the_headline_meshes_matches_the_reference_on_the_box(crd, ctx, oracle):
    """The reference's f() itself (oracle/_ref, Exchange included) run HERE on bands of the headline meshes — the rows a rank of
    the phi split owns, the neighbours' rows supplied as messages — against the same rows of the device result: FHN torus
    16384 wide bit for bit, Goldbeter torus 8192 wide within 4e-16 of the summed terms.  Bands: the slab's first and last rows
    (frozen at t < tBoundary) and rows straddling the 16-row tiles and 128-row segments of the large-slab kernels."""
    if not oracle.ref_available("fhn_torus"):
        pytest.skip("oracle/_ref not built")
    for model, nx, ny in (("fhn_torus", 16384, 16384), ("gb_torus", 8192, 32768)):
        g = crd.Grid(ctx, crd.make_params(model, nx, ny, arith=crd.ARITH_EXACT, t_boundary=38.0))
        yv, dv = g.new_vector(), g.new_vector()
        g.fill_synthetic(yv)
        P = oracle.make_params(model, nx, ny, t_boundary=38.0)
        for t in (10.0, 50.0):
            g.f(t, yv, dv)
            for j0, n in ((0, 3), (ny - 3, 3), (14, 4), (126, 4), (ny // 2 - 1, 2), (4097, 2)):
                ref = oracle.ref_rhs_band(P, t, j0, n, oracle.band_state(model, nx, ny, j0, n))
                got = np.empty(2 * nx * n)
                crd._lib.check(crd.lib().crd_memcpy_d2h(ctx._h, got.ctypes.data, dv.device_ptr + 16 * nx * j0, got.nbytes), "rows")
                if model == "fhn_torus":
                    assert got.tobytes() == ref.tobytes(), (model, t, j0)
                else:
                    assert np.all(np.abs(got - ref) <= 4e-16 * gb_uniform_scale(P)), (model, t, j0, np.abs(got - ref).max())
        yv.destroy(); dv.destroy(); g.close()


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 10, 13, 15, 20, 21])
def test_parity_vs_oracle(crd, ctx, oracle, model, variant):
    for (nx, ny) in ((400, 1600) if variant in (0, 10) else (100, 400), (3, 2), (2, 3), (257, 31), (31, 257), (128, 16), (129, 17)):
        for t in (10.0, 50.0):
            P = oracle.make_params(model, nx, ny, t_boundary=38.0)
            y = oracle.fill_state(model, 2 * nx * ny, seed=11 + nx)
            ref = oracle.rhs(P, t, y)
            got = gpu_rhs(crd, ctx, model, nx, ny, t, y, crd.ARITH_EXACT, variant, t_boundary=38.0)
            if model.startswith("fhn"):
                assert got.tobytes() == ref.tobytes(), (model, variant, nx, ny, t)
            else:
                sc = scale_of(oracle, P, y)
                assert np.all(np.abs(got - ref) <= 4e-16 * sc)
                assert (got != ref).mean() < 0.05      # pow() is correctly rounded almost everywhere
            fast = gpu_rhs(crd, ctx, model, nx, ny, t, y, crd.ARITH_FAST, variant, t_boundary=38.0)
            assert np.all(np.abs(fast - ref) <= 1e-12 * scale_of(oracle, P, y))


def test_fhn_steady_state_and_constant_field(crd, ctx):
    beta = 1.25
    nx, ny = 400, 1600
    y = np.empty((ny, nx, 2))
    y[..., 0] = -beta
    y[..., 1] = beta ** 3 - 3 * beta
    for model in ("fhn_torus", "fhn_flat"):
        for arith in (crd.ARITH_EXACT, crd.ARITH_FAST):
            d = gpu_rhs(crd, ctx, model, nx, ny, 1.0, y, arith, beta=beta, vary_beta=0, t_boundary=0.0)
            assert np.abs(d).max() < 1e-13


def test_phi_split_is_bitwise_invariant(crd, ctx, oracle):
    """Emulated ranks on one GPU: all slabs post their halo rows first, then all compute (the push never
    waits, so this order cannot deadlock).  Gathered ydot must equal the single-slab result bit for bit."""
    # (96, 203): thin slabs, everything on the main stream; (300, 1100): slabs tall enough for the overlapped
    # path (interior rows on the main stream, halo + edge rows on the auxiliary stream)
    # forced variants 13 / 20 / 21: the tiled and the streaming kernels, which acquire the ghost rows' flags inside the launch
    cases = [(m, 96, 203, (1, 2, 3, 8), None) for m in MODELS] + [("fhn_torus", 300, 1100, (2, 3), None), ("gb_flat", 300, 1100, (2,), None)]
    cases += [("fhn_torus", 300, 1100, (2, 3), v) for v in (13, 20, 21)] + [("gb_torus", 520, 700, (2,), 15)]
    for model, nx, ny, ranks, forced in cases:
        y = oracle.fill_state(model, 2 * nx * ny, seed=5)
        for t in (10.0, 50.0):
            one = gpu_rhs(crd, ctx, model, nx, ny, t, y, crd.ARITH_EXACT, t_boundary=38.0)
            for nr in ranks:
                grids, ys, ds = [], [], []
                for r in range(nr):
                    js, je = crd.decomp_phi(ny, nr, r)
                    g = crd.Grid(ctx, crd.make_params(model, nx, ny, js=js, je=je, t_boundary=38.0))
                    g.set_variant(forced if forced else (10 if (nr + r) % 2 else 0))      # default: mix the tiled and the direct kernel
                    grids.append(g)
                    ys.append(crd.NVector.from_numpy(ctx, y[2 * nx * js: 2 * nx * (je + 1)], 2 * nx * ny))
                    ds.append(g.new_vector())
                for r in range(nr):
                    grids[r].halo_connect_local(grids[(r - 1) % nr], grids[(r + 1) % nr])
                for rep in range(3):   # several epochs: exercises the ghost double-buffering
                    for r in range(nr):
                        grids[r].post_halo(ys[r])
                    for r in range(nr):
                        grids[r].compute(t, ys[r], ds[r])
                got = np.concatenate([d.to_numpy() for d in ds])
                assert got.tobytes() == one.tobytes(), (model, t, nr)
                ctx.sync()
                for g in grids:
                    g.close()


@pytest.mark.parametrize("variant,mesh", [(13, (520, 1400)), (15, (520, 1400)), (20, (520, 1400)), (21, (520, 1400)), (0, (520, 1400)),
                                          (21, (1030, 8300)), (0, (1030, 8300))])
def test_exchange_inside_the_launch(crd, oracle, variant, mesh):
    """crd_rhs on a connected grid: ONE launch per evaluation — its first CTAs push the boundary rows to the neighbours, the
    tiles / row segments that touch a ghost row come last and acquire their strip's flag.  Two and three ranks emulated on one
    GPU, each with its own context (stream), so the launches run concurrently and really wait for each other; plain states
    and fused stage states (their combination is what gets pushed); several epochs through the double-buffered ghost rows.
    Gathered result bit-identical to the single slab."""
    # (1030 x 8300: slabs of a few million points, whose streaming-kernel segments are sized to fill whole waves; the automatic
    # choice there is the tiled kernel for the plain state and the streaming kernel for the fused stage)
    model, (nx, ny) = "fhn_torus", mesh
    y = oracle.fill_state(model, 2 * nx * ny, seed=5)
    x2 = oracle.fill_state(model, 2 * nx * ny, seed=6)
    c0 = crd.Context(0)
    one = gpu_rhs(crd, c0, model, nx, ny, 10.0, y, crd.ARITH_EXACT, t_boundary=38.0)
    one50 = gpu_rhs(crd, c0, model, nx, ny, 50.0, y, crd.ARITH_EXACT, t_boundary=38.0)
    g1 = crd.Grid(c0, crd.make_params(model, nx, ny, t_boundary=38.0))
    Y1, X1, D1 = crd.NVector.from_numpy(c0, y), crd.NVector.from_numpy(c0, x2), g1.new_vector()
    g1.f_lincomb(50.0, [1.0, 0.01], [Y1, X1], D1)
    one_lc = D1.to_numpy()
    g1.close()
    for nr in (2, 3):
        ctxs = [crd.Context(0) for _ in range(nr)]
        grids, ys, xs, ds = [], [], [], []
        for r in range(nr):
            ctxs[r].set_halo_timeout(20000.0)
            js, je = crd.decomp_phi(ny, nr, r)
            g = crd.Grid(ctxs[r], crd.make_params(model, nx, ny, js=js, je=je, t_boundary=38.0))
            g.set_variant(variant)
            grids.append(g)
            ys.append(crd.NVector.from_numpy(ctxs[r], y[2 * nx * js: 2 * nx * (je + 1)], 2 * nx * ny))
            xs.append(crd.NVector.from_numpy(ctxs[r], x2[2 * nx * js: 2 * nx * (je + 1)], 2 * nx * ny))
            ds.append(g.new_vector())
        for r in range(nr):
            grids[r].halo_connect_local(grids[(r - 1) % nr], grids[(r + 1) % nr])
        for t, want in ((10.0, one), (50.0, one50), (10.0, one), (10.0, one), (50.0, one50)):
            l0 = [c.launches for c in ctxs]
            for r in range(nr):
                grids[r].f(t, ys[r], ds[r])           # asynchronous: every rank's launch is in flight before anyone is waited for
            for c in ctxs:
                c.sync()
            if variant != 0 or nx * ny // nr >= (2 << 20):   # (automatic on the small mesh: the direct kernel and separate launches)
                assert [c.launches - a for c, a in zip(ctxs, l0)] == [1] * nr
            got = np.concatenate([d.to_numpy() for d in ds])
            assert got.tobytes() == want.tobytes(), (variant, nr, t)
        if variant in (13, 20, 21, 0):
            for r in range(nr):
                grids[r].f_lincomb(50.0, [1.0, 0.01], [ys[r], xs[r]], ds[r])
            for c in ctxs:
                c.sync()
            assert np.concatenate([d.to_numpy() for d in ds]).tobytes() == one_lc.tobytes(), (variant, nr)
        for g in grids:
            g.close()
        for c in ctxs:
            c.close()
    c0.close()


def test_a_neighbour_that_never_posts_fails_the_run_instead_of_hanging_or_integrating_on(crd):
    """A rank of a phi split whose neighbour has died: the wait for its boundary rows gives up after the halo timeout, the
    evaluation finishes with stale rows — and must not be used: the next wait for the stream reports it, the context stays
    failed, crd_f returns -1 and ARKode ends with ARK_RHSFUNC_FAIL (the reference would hang in MPI_Wait,
    FHNmodel_torus.cpp:904-946).  Both ways of running the exchange (inside the launch, as separate launches)."""
    nx, ny = 300, 1100
    for variant in (13, 1):
        ca, cb = crd.Context(0), crd.Context(0)
        ca.set_halo_timeout(300.0)
        js, je = crd.decomp_phi(ny, 2, 0)
        ga = crd.Grid(ca, crd.make_params("fhn_torus", nx, ny, js=js, je=je))
        js, je = crd.decomp_phi(ny, 2, 1)
        gb = crd.Grid(cb, crd.make_params("fhn_torus", nx, ny, js=js, je=je))      # never evaluates anything
        ga.halo_connect_local(gb, gb)
        ga.set_variant(variant)
        y, d = ga.new_vector(), ga.new_vector()
        ga.fill_synthetic(y)
        assert ca.failed == 0
        s = crd.ARKodeSolver(ga, y, fused="full", resident=False)
        flag, _ = s.ARKode(1.0)
        assert flag < 0, flag                      # ARK_FIRST_RHSFUNC_ERR / ARK_RHSFUNC_FAIL, never success
        assert ca.failed in (100, 101, 102, 103, 104, 105)
        with pytest.raises(crd.CrdError):
            ca.sync()
        with pytest.raises(crd.CrdError):
            ga.f(0.0, y, d)                        # sticky: nothing runs on a failed context
        with pytest.raises(crd.CrdError):
            ga.new_vector()
        assert not np.isfinite(crd.N_VWrmsNorm(y, y))
        ca.clear_error()
        assert ca.failed == 0 and np.isfinite(crd.N_VMaxNorm(y))
        s.free(); ga.close(); gb.close(); ca.close(); cb.close()


def test_host_entry_matches_device_entry(crd, ctx, oracle):
    nx, ny = 1024, 4099
    for model in ("fhn_torus", "gb_torus"):
        y = oracle.fill_state(model, 2 * nx * ny, seed=3)
        dev = gpu_rhs(crd, ctx, model, nx, ny, 10.0, y, crd.ARITH_EXACT, t_boundary=38.0)
        g = crd.Grid(ctx, crd.make_params(model, nx, ny, t_boundary=38.0))
        out = np.empty_like(y)
        g.f_host(10.0, y, out)
        assert out.tobytes() == dev.tobytes()
        g.close()


def test_synthetic_state_matches_oracle_stream(crd, ctx, oracle):
    for model in ("fhn_torus", "gb_torus"):
        n = 100003
        v = crd.NVector(ctx, n)
        ctx.fill_synthetic(model, n, v.device_ptr, seed=0x5EED, first_elem=12345)
        assert v.to_numpy().tobytes() == oracle.fill_state(model, n, seed=0x5EED, first_elem=12345).tobytes()


def test_full_size_properties(crd, ctx):
    """BASELINE size (16384 x 16384, 4.29 GB per vector): properties that need no CPU pass.
    (1) uniform steady state -> ydot == 0 to rounding; (2) diffusion-only operator is linear;
    (3) the frozen rows are exactly zero and nothing else changes between t < tB and t > tB."""
    nx = ny = 16384
    n = 2 * nx * ny
    beta = 1.25
    g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny, arith=crd.ARITH_EXACT, beta=beta, vary_beta=0, t_boundary=38.0))
    y, d = g.new_vector(), g.new_vector()
    # (1)
    import ctypes as C
    half = crd.NVector(ctx, 2)
    half.set(np.array([-beta, beta ** 3 - 3 * beta]))
    # build the uniform state with two strided fills: y = const pair repeated -> use lincomb of synthetic? simpler:
    crd.N_VConst(-beta, y)                      # u = v = -beta, then fix v through f's linearity in v: f_u = ... - v
    g.f(50.0, y, d)
    # with v = -beta instead of Vs: du = 3u - u^3 - v = -3b + b^3 + b, dv = eps*(u + b) = 0
    assert crd.N_VMaxNorm(d) == pytest.approx(abs(-3 * beta + beta ** 3 + beta), rel=1e-12)
    # (3)
    g.fill_synthetic(y)
    d2 = g.new_vector()
    g.f(50.0, y, d)
    g.f(10.0, y, d2)
    crd.N_VLinearSum(1.0, d, -1.0, d2, d2)     # differs only on rows 0 and ny-1
    diff = d2.to_numpy().reshape(ny, nx * 2)
    assert not diff[1:-1].any()
    assert diff[0].any() and diff[-1].any()
    del diff
    # (2) linearity of the diffusion-only Goldbeter operator: f(a*y1 + b*y2) = a f(y1) + b f(y2)
    g2 = crd.Grid(ctx, crd.make_params("gb_torus", nx, ny, arith=crd.ARITH_FAST, just_diffusion=1))
    y2 = g.new_vector()
    ctx.fill_synthetic("gb_torus", n, y2.device_ptr, seed=99)
    g2.f(0.0, y, d)
    g2.f(0.0, y2, d2)
    crd.N_VLinearSum(2.0, d, -0.5, d2, d)       # a f(y1) + b f(y2)
    crd.N_VLinearSum(2.0, y, -0.5, y2, y)       # a y1 + b y2
    g2.f(0.0, y, d2)
    scale = crd.N_VMaxNorm(d)
    crd.N_VLinearSum(1.0, d, -1.0, d2, d)
    assert crd.N_VMaxNorm(d) <= 1e-11 * scale
    g.close(); g2.close()


def test_exact_division_edge_values(crd, ctx, oracle):
    """The EXACT torus stencil divides by 2dx, dx*dx, dy*dy through a reciprocal + two FMA corrections
    (div_const_rn); zeros (uniform patches, signed), subnormal-range and huge values must take the IEEE
    fallback and still match the reference bit for bit."""
    nx, ny = 64, 48
    base = oracle.fill_state("fhn_torus", 2 * nx * ny, seed=21).reshape(ny, nx, 2)
    cases = []
    a = base.copy(); a[10:30, 5:40, :] = 0.0; cases.append(a)                    # exact zeros -> zero numerators
    a = base.copy(); a[..., 0] = 0.75; cases.append(a)                           # uniform u: all stencil terms are +-0
    a = base.copy(); a[..., 0] = -0.0; cases.append(a)
    cases.append(base * 1e-300)                                                  # quotients in the subnormal range
    cases.append(base * 1e-160)
    a = base.copy(); a[::2] *= 1e-310; cases.append(a)                           # subnormal inputs next to normal ones
    cases.append(base * 1e90)                                                    # u^3 ~ 1e270, no overflow
    for model in ("fhn_torus", "gb_torus"):
        for k, yy in enumerate(cases):
            yy = np.ascontiguousarray(yy).ravel()
            if model == "gb_torus":
                yy = np.abs(yy)
            P = oracle.make_params(model, nx, ny, just_diffusion=1 if model == "gb_torus" else 0)
            ref = oracle.rhs(P, 50.0, yy)
            for variant in (0, 10, 20):
                got = gpu_rhs(crd, ctx, model, nx, ny, 50.0, yy, crd.ARITH_EXACT, variant, just_diffusion=1 if model == "gb_torus" else 0)
                assert got.tobytes() == ref.tobytes(), (model, k, variant)


def stage_state_op_by_op(crd, grid, coefs, V):
    """assemble() of the RK driver without fused operations: sdata = 0; sdata = c_j X_j + sdata; z = X_0 + sdata (c_0 = 1)."""
    z, sdata = grid.new_vector(), grid.new_vector()
    crd.N_VConst(0.0, sdata)
    for c, x in zip(coefs[1:], V[1:]):
        crd.N_VLinearSum(c, x, 1.0, sdata, sdata)
    if len(coefs) > 1:
        crd.N_VLinearSum(1.0, V[0], 1.0, sdata, z)
    else:
        crd.N_VScale(1.0, V[0], z)
    sdata.destroy()
    return z


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("arith", ["exact", "fast"])
def test_fused_stage_rhs_equals_assembly_then_rhs(crd, ctx, oracle, model, arith):
    """crd_rhs_lincomb: f(t, sum c_j X_j) with the stage state formed inside the kernel must give the bits of the state
    assembled first and evaluated second — on an EXACT grid the op-by-op assembly of the RK driver (N_VLinearSum chain), on
    a FAST grid N_VLinearCombination's chain of fused multiply-adds — for every kernel (tiled / direct / streaming)."""
    ar = crd.ARITH_EXACT if arith == "exact" else crd.ARITH_FAST
    for (nx, ny) in ((300, 700), (64, 37), (2, 3)):
        n_el = 2 * nx * ny
        X = [oracle.fill_state(model, n_el, seed=40 + j) for j in range(5)]
        for ncomb, coefs in ((1, [1.0]), (2, [1.0, 0.013]), (3, [1.0, 0.02, -0.007]), (5, [1.0, 0.004, 0.005, 0.009, -0.001])):
            for variant in (0, 1, 5, 10, 13, 20, 21):
                g = crd.Grid(ctx, crd.make_params(model, nx, ny, t_boundary=38.0, arith=ar))
                g.set_variant(variant)
                V = [crd.NVector.from_numpy(ctx, x) for x in X[:ncomb]]
                d1, d2 = g.new_vector(), g.new_vector()
                if arith == "exact":
                    z = stage_state_op_by_op(crd, g, coefs, V)
                else:
                    z = g.new_vector()
                    crd.N_VLinearCombination(coefs, V, z)
                for t in (10.0, 50.0):
                    g.f(t, z, d1)
                    g.f_lincomb(t, coefs, V, d2)
                    assert d1.to_numpy().tobytes() == d2.to_numpy().tobytes(), (model, nx, ny, ncomb, variant, t)
                g.close()


@pytest.mark.parametrize("model,arith", [("fhn_torus", "exact"), ("fhn_torus", "fast"), ("gb_torus", "exact"), ("fhn_flat", "exact"), ("gb_flat", "fast")])
def test_last_stage_fused_with_the_finish_equals_stage_then_finish(crd, ctx, model, arith):
    """crd_rhs_lincomb_finish (streaming kernel, F_5 never stored) against crd_rhs_lincomb followed by N_VErkFinish_Crd:
    ynew bit-identical; the error sum identical too on an EXACT grid (double-double accumulation), equal to summation-order
    rounding on a FAST one; ragged strip and row counts; frozen rows (t < tBoundary)."""
    nx, ny = 520, 2100          # > 1 Mi points, 3 strips (the last one 8 columns wide), rows not a multiple of the segments
    ar = crd.ARITH_EXACT if arith == "exact" else crd.ARITH_FAST
    grid = crd.Grid(ctx, crd.make_params(model, nx, ny, arith=ar, t_boundary=38.0))
    rng = np.random.default_rng(7)
    n = 2 * nx * ny
    lo, hi = (-2.0, 2.0) if model.startswith("fhn") else (0.1, 1.6)
    X = [crd.NVector.from_numpy(ctx, rng.uniform(lo, hi, n))] + [crd.NVector.from_numpy(ctx, rng.uniform(-1.0, 1.0, n)) for _ in range(4)]
    h = 1e-3
    c = [1.0, h * 5 / 32, h * 7 / 32, h * 13 / 32, -h / 32]
    hb = [h / 6, h / 3, h / 3, h / 6, 0.0]
    hd = [h * (1 / 6 + 0.5), h * (1 / 3 - 7 / 3), h * (1 / 3 - 7 / 3), h * (1 / 6 - 13 / 6), h * 16 / 3]
    rtol, atol = 1e-5, 1e-10
    for t in (10.0, 50.0):
        F5, want = grid.new_vector(), grid.new_vector()
        grid.f_lincomb(t, c, X, F5)
        e2, y2 = crd.N_VErkFinish(hb, hd, X[0], X[1:] + [F5], want, rtol, atol, exact=(arith == "exact"))
        got = grid.new_vector()
        rc, fe2, fy2 = grid.f_lincomb_finish(t, c, hb, hd, X, got, rtol, atol)
        assert rc == 0
        assert got.to_numpy().tobytes() == want.to_numpy().tobytes()
        if arith == "exact":
            assert fe2 == e2          # double-double error sum: the same bits whatever the order of summation
        assert abs(fe2 - e2) <= 1e-11 * e2 and abs(fy2 - y2) <= 1e-11 * y2
        got.destroy()
        for v in (F5, want):
            v.destroy()
    # does not apply: small meshes (the caller issues the two operations)
    small = crd.Grid(ctx, crd.make_params(model, 64, 96, arith=ar))
    Xs = [small.new_vector() for _ in range(5)]
    for v in Xs:
        crd.N_VConst(0.5, v)
    assert small.f_lincomb_finish(50.0, c, hb, hd, Xs, small.new_vector(), rtol, atol)[0] == 1
    small.close(); grid.close()


@pytest.mark.parametrize("model", ["fhn_torus", "gb_flat"])
def test_streaming_kernel_with_wave_filling_segments(crd, ctx, model):
    """A mesh of a few million points: the streaming kernel cuts its strips into segments that fill whole waves of the
    persistent CTAs (stream_seg_rows: here 47-row segments, the last one 11 rows, 5 strips with the last one 6 columns wide).
    Plain state, 2- and 3-vector stages through the streaming kernel (2 and 3 CTAs per SM) and the automatic choice against the
    direct kernel, and the fused finish against stage + N_VErkFinish: bit for bit (EXACT)."""
    nx, ny = 1030, 4100
    grid = crd.Grid(ctx, crd.make_params(model, nx, ny, t_boundary=38.0))
    rng = np.random.default_rng(11)
    n = 2 * nx * ny
    lo, hi = (-2.0, 2.0) if model.startswith("fhn") else (0.1, 1.6)
    X = [crd.NVector.from_numpy(ctx, rng.uniform(lo, hi, n))] + [crd.NVector.from_numpy(ctx, rng.uniform(-1.0, 1.0, n)) for _ in range(4)]
    d1, d2 = grid.new_vector(), grid.new_vector()
    h = 1e-3
    for t in (10.0, 50.0):
        for ncomb, coefs in ((0, None), (2, [1.0, 0.5 * h]), (3, [1.0, 0.25 * h, 0.25 * h])):
            def ev(out):
                if ncomb == 0:
                    grid.f(t, X[0], out)
                else:
                    grid.f_lincomb(t, coefs, X[:ncomb], out)
            grid.set_variant(1)
            ev(d1)
            want = d1.to_numpy().tobytes()
            for variant in (0, 20, 21):
                grid.set_variant(variant)
                ev(d2)
                assert d2.to_numpy().tobytes() == want, (model, t, ncomb, variant)
        grid.set_variant(0)
        c = [1.0, h * 5 / 32, h * 7 / 32, h * 13 / 32, -h / 32]
        hb = [h / 6, h / 3, h / 3, h / 6, 0.0]
        hd = [h * (1 / 6 + 0.5), h * (1 / 3 - 7 / 3), h * (1 / 3 - 7 / 3), h * (1 / 6 - 13 / 6), h * 16 / 3]
        grid.f_lincomb(t, c, X, d1)
        e2, y2 = crd.N_VErkFinish(hb, hd, X[0], X[1:] + [d1], d2, 1e-5, 1e-10, exact=True)
        got = grid.new_vector()
        rc, fe2, fy2 = grid.f_lincomb_finish(t, c, hb, hd, X, got, 1e-5, 1e-10)
        assert rc == 0 and got.to_numpy().tobytes() == d2.to_numpy().tobytes() and fe2 == e2
        got.destroy()
    grid.close()


@pytest.mark.parametrize("model,arith", [("fhn_torus", "exact"), ("fhn_torus", "fast"), ("gb_torus", "exact"), ("fhn_flat", "exact"), ("gb_flat", "fast")])
@pytest.mark.parametrize("nx,ny", [(1030, 2100), (481, 4400)])
def test_two_evaluations_in_one_pass(crd, ctx, model, arith, nx, ny):
    """crd_rhs_pair: f1 = f(t1, y) and f2 = f(t2, y + c f1) from one pass over y, against crd_rhs followed by crd_rhs_lincomb —
    bit for bit in EXACT arithmetic (same device functions), to rounding in FAST.  Strips of 240 columns whose last one is 70 / 1 columns wide, the
    periodic columns fetched as one or two wrapped pieces, segments that do not divide the rows, frozen boundary rows in both,
    one or neither evaluation."""
    ar = crd.ARITH_EXACT if arith == "exact" else crd.ARITH_FAST
    grid = crd.Grid(ctx, crd.make_params(model, nx, ny, arith=ar, t_boundary=38.0))
    rng = np.random.default_rng(3)
    lo, hi = (-2.0, 2.0) if model.startswith("fhn") else (0.1, 1.6)
    y = crd.NVector.from_numpy(ctx, rng.uniform(lo, hi, 2 * nx * ny))
    f1, f2, w1, w2 = (grid.new_vector() for _ in range(4))
    c = 0.5e-3
    for t1, t2 in ((10.0, 10.3), (37.9, 38.2), (50.0, 50.1)):
        crd.N_VConst(-7.0, f1); crd.N_VConst(-7.0, f2)
        assert grid.f_pair(t1, t2, c, y, f1, f2) == 0
        grid.f(t1, y, w1)
        grid.f_lincomb(t2, [1.0, c], [y, w1], w2)
        if arith == "exact":
            assert f1.to_numpy().tobytes() == w1.to_numpy().tobytes(), (model, arith, t1)
            assert f2.to_numpy().tobytes() == w2.to_numpy().tobytes(), (model, arith, t2)
        else:       # FAST: the compiler contracts each kernel's expressions on its own; equal to rounding
            for got, want in ((f1, w1), (f2, w2)):
                a, b = got.to_numpy(), want.to_numpy()
                assert np.abs(a - b).max() <= 1e-12 * (1.0 + np.abs(b).max()), (model, arith, t1)
    # does not apply: a small mesh, a phi-split grid (the caller evaluates one by one)
    small = crd.Grid(ctx, crd.make_params(model, 300, 400, arith=ar))
    v = [small.new_vector() for _ in range(3)]
    assert small.f_pair(1.0, 1.1, c, *v) == 1
    small.close(); grid.close()


@pytest.mark.parametrize("model,arith", [("fhn_torus", "exact"), ("gb_torus", "exact"), ("fhn_flat", "fast")])
def test_two_evaluations_in_one_pass_on_a_phi_split(model, arith):
    """crd_rhs_pair on connected grids (tests/pair_split_worker.py): the ranks, emulated on one GPU with a context and a stream
    each, exchange TWO rows of y per side, then every rank forms F1 and F2 of its slab in one pass; gathered F1 and F2 are the
    bits of the single-slab evaluations.  Runs in a process of its own with CUDA_MODULE_LOADING = EAGER and
    CUDA_DEVICE_MAX_CONNECTIONS = 32, because the emulation has hazards that R ranks on R GPUs do not have: a rank's sequence is
    push, wait, pass, all ranks are driven by ONE host thread, and (a) the first launch of a kernel loads it lazily, which waits
    for the device — where another rank's wait kernel is spinning on rows this thread has not pushed yet; (b) with the default 8
    hardware queues two ranks' streams can share one, and one rank's push then queues up behind the other's pass, which waits for
    exactly that push."""
    import subprocess
    import sys
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32", CUDA_MODULE_LOADING="EAGER")
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "pair_split_worker.py"), model, arith],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "PAIR SPLIT OK" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:])
