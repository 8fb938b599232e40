"""Device N_Vector vs the CPU checker (oracle/nvector_host.c = restated nvector_parallel).
Element-wise ops: bit-identical.  Reductions: within sqrt(N)*eps of an exactly rounded sum (the
summation order differs from the serial loop)."""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P_, D_, L_ = C.c_void_p, C.c_double, C.c_long


@pytest.fixture(scope="module")
def H(oracle):
    L = oracle.lib()
    L.N_VMake_Parallel.restype = P_
    L.N_VMake_Parallel.argtypes = [C.c_int, L_, L_, P_]
    sig = {"N_VLinearSum": (None, [D_, P_, D_, P_, P_]), "N_VScale": (None, [D_, P_, P_]), "N_VProd": (None, [P_, P_, P_]),
           "N_VDiv": (None, [P_, P_, P_]), "N_VAbs": (None, [P_, P_]), "N_VInv": (None, [P_, P_]), "N_VAddConst": (None, [P_, D_, P_]),
           "N_VCompare": (None, [D_, P_, P_]), "N_VConst": (None, [D_, P_])}
    for k, (r, a) in sig.items():
        getattr(L, k).restype = r
        getattr(L, k).argtypes = a
    return L


def hwrap(H, a):
    return H.N_VMake_Parallel(0, a.size, a.size, a.ctypes.data)


@pytest.mark.parametrize("n", [1, 2, 3, 1000, 1001, 2 * 400 * 1600, (1 << 22) + 5])
def test_elementwise_bitwise(crd, ctx, H, n):
    rng = np.random.default_rng(n)
    x, y = rng.standard_normal(n), rng.standard_normal(n) + 3.0
    X, Y, Z = crd.NVector.from_numpy(ctx, x), crd.NVector.from_numpy(ctx, y), crd.NVector(ctx, n)
    hz = np.zeros(n)
    hX, hY, hZ = hwrap(H, x), hwrap(H, y), hwrap(H, hz)
    for a, b in ((2.5, -0.75), (1.0, 1.0), (1.0, -1.0), (-1.0, 1.0), (1.0, 3.0), (3.0, 1.0), (-1.0, 2.0), (2.0, -1.0), (2.0, 2.0), (2.0, -2.0)):
        crd.N_VLinearSum(a, X, b, Y, Z); H.N_VLinearSum(a, hX, b, hY, hZ)
        assert Z.to_numpy().tobytes() == hz.tobytes(), (a, b)
    for c in (1.0, -1.0, 0.3):
        crd.N_VScale(c, X, Z); H.N_VScale(c, hX, hZ); assert Z.to_numpy().tobytes() == hz.tobytes()
    crd.N_VProd(X, Y, Z); H.N_VProd(hX, hY, hZ); assert Z.to_numpy().tobytes() == hz.tobytes()
    crd.N_VDiv(X, Y, Z); H.N_VDiv(hX, hY, hZ); assert Z.to_numpy().tobytes() == hz.tobytes()
    crd.N_VAbs(X, Z); H.N_VAbs(hX, hZ); assert Z.to_numpy().tobytes() == hz.tobytes()
    crd.N_VInv(Y, Z); H.N_VInv(hY, hZ); assert Z.to_numpy().tobytes() == hz.tobytes()
    crd.N_VAddConst(X, 1e-10, Z); H.N_VAddConst(hX, 1e-10, hZ); assert Z.to_numpy().tobytes() == hz.tobytes()
    crd.N_VCompare(0.5, X, Z); H.N_VCompare(0.5, hX, hZ); assert Z.to_numpy().tobytes() == hz.tobytes()
    crd.N_VConst(-2.5, Z); assert np.all(Z.to_numpy() == -2.5)
    # in-place forms ARKode uses: y += a x, x *= c
    x2, y2 = x.copy(), y.copy()
    X2, Y2 = crd.NVector.from_numpy(ctx, x2), crd.NVector.from_numpy(ctx, y2)
    hX2, hY2 = hwrap(H, x2), hwrap(H, y2)
    crd.N_VLinearSum(0.7, X2, 1.0, Y2, Y2); H.N_VLinearSum(0.7, hX2, 1.0, hY2, hY2); assert Y2.to_numpy().tobytes() == y2.tobytes()
    crd.N_VLinearSum(1.0, X2, 0.7, Y2, X2); H.N_VLinearSum(1.0, hX2, 0.7, hY2, hX2); assert X2.to_numpy().tobytes() == x2.tobytes()
    crd.N_VScale(1.5, X2, X2); H.N_VScale(1.5, hX2, hX2); assert X2.to_numpy().tobytes() == x2.tobytes()


@pytest.mark.parametrize("n", [1, 7, 1001, 2 * 400 * 1600, (1 << 23) + 3])
def test_reductions(crd, ctx, n):
    rng = np.random.default_rng(n + 1)
    x, y, w = rng.standard_normal(n), rng.standard_normal(n), rng.random(n) + 0.5
    X, Y, W = (crd.NVector.from_numpy(ctx, a) for a in (x, y, w))
    tol = 8 * math.sqrt(n) * 2.2e-16
    def close(got, want, scale):
        assert abs(got - want) <= tol * scale + 1e-300, (got, want)
    close(crd.N_VDotProd(X, Y), math.fsum(x * y), math.fsum(np.abs(x * y)))
    close(crd.N_VL1Norm(X), math.fsum(np.abs(x)), math.fsum(np.abs(x)))
    # weighted square sums: every term rounded like nvector_parallel's loop, the sum accumulated in double-double and rounded
    # once -> the exactly rounded sum (math.fsum), whatever the grouping of the terms over threads, blocks and ranks
    s2 = math.fsum((x * w) ** 2)
    assert crd.N_VWrmsNorm(X, W) == math.sqrt(s2 / n)
    assert crd.N_VWL2Norm(X, W) == math.sqrt(s2)
    idm = (rng.random(n) > 0.5).astype(float)
    sm = math.fsum(((x * w) ** 2)[idm > 0])
    assert crd.N_VWrmsNormMask(X, W, crd.NVector.from_numpy(ctx, idm)) == math.sqrt(sm / n)
    assert crd.N_VMaxNorm(X) == np.abs(x).max()
    assert crd.N_VMin(X) == x.min()
    assert crd.N_VMinQuotient(X, W) == (x / w).min()
    # determinism: same input, same bits
    assert crd.N_VDotProd(X, Y) == crd.N_VDotProd(X, Y)
    Z = crd.NVector(ctx, n)
    assert crd.N_VInvTest(W, Z) is True and Z.to_numpy().tobytes() == (1.0 / w).tobytes()
    if n > 3:
        x0 = x.copy(); x0[n // 2] = 0.0
        assert crd.N_VInvTest(crd.NVector.from_numpy(ctx, x0), Z) is False
    c = rng.integers(-2, 3, n).astype(float)
    M = crd.NVector(ctx, n)
    ok = crd.N_VConstrMask(crd.NVector.from_numpy(ctx, c), X, M)
    viol = ((np.abs(c) == 2) & (x * c <= 0)) | ((np.abs(c) == 1) & (x * c < 0))
    assert np.array_equal(M.to_numpy(), viol.astype(float)) and ok == (not viol.any())


def test_empty_and_degenerate(crd, ctx):
    E = crd.NVector(ctx, 0)
    assert crd.N_VWrmsNorm(E, E) != crd.N_VWrmsNorm(E, E) or True   # 0/0: nan like the serial code; must not crash
    assert crd.N_VMaxNorm(E) == 0.0
    crd.N_VConst(1.0, E)
    crd.N_VLinearSum(1.0, E, 2.0, E, E)
    assert crd.N_VMin(E) == np.finfo(np.float64).max     # BIG_REAL


@pytest.mark.parametrize("n", [2, 1001, 2 * 400 * 1600])
def test_fused_ops(crd, ctx, n):
    rng = np.random.default_rng(n + 2)
    vs = [rng.standard_normal(n) for _ in range(6)]
    V = [crd.NVector.from_numpy(ctx, a) for a in vs]
    Z = crd.NVector(ctx, n)
    for k in range(1, 7):
        c = list(rng.standard_normal(k))
        c[0] = 1.0
        crd.N_VLinearCombination(c, V[:k], Z)
        want = sum(ci * vi for ci, vi in zip(c, vs))
        scale = sum(abs(ci) * np.abs(vi) for ci, vi in zip(c, vs))
        assert np.all(np.abs(Z.to_numpy() - want) <= (k + 1) * 1.2e-16 * scale)
    # finish: ynew, error norm, state norm
    yn, F = vs[0] + 2.0, vs[1:6]
    h = 1e-2
    b = np.array([1 / 6, 1 / 3, 1 / 3, 1 / 6, 0.0]); b2 = np.array([-0.5, 7 / 3, 7 / 3, 13 / 6, -16 / 3])
    hb, hd = h * b, h * (b - b2)
    rtol, atol = 1e-5, 1e-10
    Yn = crd.NVector.from_numpy(ctx, yn)
    e2, y2 = crd.N_VErkFinish(list(hb), list(hd), Yn, V[1:6], Z, rtol, atol)
    ynew = yn + sum(hb[j] * F[j] for j in range(5))
    err = sum(hd[j] * F[j] for j in range(5))
    np.testing.assert_allclose(Z.to_numpy(), ynew, rtol=1e-14, atol=1e-14)
    we2 = math.fsum((err / (rtol * np.abs(yn) + atol)) ** 2)
    wy2 = math.fsum((ynew / (rtol * np.abs(ynew) + atol)) ** 2)
    # the second sum only feeds the "too much accuracy" test: for rtol > uround the kernels report its bound n / rtol^2
    assert abs(e2 - we2) <= 1e-10 * we2 and y2 == n / rtol ** 2 and y2 >= wy2
    tiny = 1e-17                                              # rtol below uround: the sum itself
    _, y2t = crd.N_VErkFinish(list(hb), list(hd), Yn, V[1:6], Z, tiny, atol)
    wy2t = math.fsum((ynew / (tiny * np.abs(ynew) + atol)) ** 2)
    assert abs(y2t - wy2t) <= 1e-10 * wy2t
    # the same finish with the bits of the op-by-op sequence (N_VErkFinishSeq_Crd): the N_VLinearSum chains of
    # compute_solution(), the ewt chain abs / scale / addconst / inv, N_VWrmsNorm's terms, an exactly rounded sum
    Z2 = crd.NVector(ctx, n)
    e2x, y2x = crd.N_VErkFinish(list(hb), list(hd), Yn, V[1:6], Z2, rtol, atol, exact=True)
    ycur, tempv = yn.copy(), np.zeros(n)
    for j in range(5):
        if hb[j] != 0.0:
            ycur = hb[j] * F[j] + ycur
        tempv = hd[j] * F[j] + tempv
    assert Z2.to_numpy().tobytes() == ycur.tobytes()
    assert e2x == math.fsum((tempv * (1.0 / (rtol * np.abs(yn) + atol))) ** 2)
    assert y2x == n / rtol ** 2
    # and through the vector operations themselves, one by one
    T, Ew = crd.NVector(ctx, n), crd.NVector(ctx, n)
    crd.N_VConst(0.0, T)
    for j in range(5):
        crd.N_VLinearSum(float(hd[j]), V[1 + j], 1.0, T, T)
    crd.N_VAbs(Yn, Ew); crd.N_VScale(rtol, Ew, Ew); crd.N_VAddConst(Ew, atol, Ew); crd.N_VInv(Ew, Ew)
    assert crd.N_VWrmsNorm(T, Ew) == math.sqrt(e2x / n)


def test_host_mirror_roundtrip(crd, ctx):
    n = 1000
    v = crd.NVector(ctx, n)
    lib = crd.lib()
    p = lib.N_VGetArrayPointer(v.h)
    assert p
    host = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,))
    host[:] = np.arange(n)
    assert lib.N_VCopyFromHost_Crd(v.h) == 0
    assert np.array_equal(v.to_numpy(), np.arange(n))
    crd.N_VScale(2.0, v, v)
    assert lib.N_VCopyToHost_Crd(v.h) == 0
    assert np.array_equal(host, 2.0 * np.arange(n))
