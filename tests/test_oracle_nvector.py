"""Pins oracle/nvector_host.c (the restated nvector_parallel) against closed-form numpy results."""
import ctypes as C
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def H(oracle):
    L = oracle.lib()
    P, D, Lg = C.c_void_p, C.c_double, C.c_long
    L.N_VMake_Parallel.restype = P
    L.N_VMake_Parallel.argtypes = [C.c_int, Lg, Lg, P]
    L.N_VDestroy_Parallel.argtypes = [P]
    sig = {"N_VLinearSum": (None, [D, P, D, P, P]), "N_VConst": (None, [D, P]), "N_VProd": (None, [P, P, P]),
           "N_VDiv": (None, [P, P, P]), "N_VScale": (None, [D, P, P]), "N_VAbs": (None, [P, P]), "N_VInv": (None, [P, P]),
           "N_VAddConst": (None, [P, D, P]), "N_VDotProd": (D, [P, P]), "N_VMaxNorm": (D, [P]), "N_VWrmsNorm": (D, [P, P]),
           "N_VWrmsNormMask": (D, [P, P, P]), "N_VMin": (D, [P]), "N_VWL2Norm": (D, [P, P]), "N_VL1Norm": (D, [P]),
           "N_VCompare": (None, [D, P, P]), "N_VInvTest": (C.c_int, [P, P]), "N_VConstrMask": (C.c_int, [P, P, P]),
           "N_VMinQuotient": (D, [P, P])}
    for k, (r, a) in sig.items():
        getattr(L, k).restype = r
        getattr(L, k).argtypes = a
    return L


def wrap(H, a):
    return H.N_VMake_Parallel(0, a.size, a.size, a.ctypes.data)


def test_host_nvector_ops(H):
    rng = np.random.default_rng(1)
    n = 1001
    x, y, w = rng.standard_normal(n), rng.standard_normal(n), rng.random(n) + 0.5
    z = np.zeros(n)
    X, Y, W, Z = (wrap(H, a) for a in (x, y, w, z))
    for a, b in ((2.5, -0.75), (1.0, 1.0), (1.0, -1.0), (-1.0, 1.0), (1.0, 3.0), (3.0, 1.0), (-1.0, 2.0), (2.0, -1.0), (2.0, 2.0), (2.0, -2.0)):
        H.N_VLinearSum(a, X, b, Y, Z)
        np.testing.assert_allclose(z, a * x + b * y, rtol=1e-15, atol=1e-15)
    H.N_VConst(3.25, Z); assert np.all(z == 3.25)
    H.N_VProd(X, Y, Z); assert np.array_equal(z, x * y)
    H.N_VDiv(X, W, Z); assert np.array_equal(z, x / w)
    H.N_VScale(-1.0, X, Z); assert np.array_equal(z, -x)
    H.N_VScale(0.3, X, Z); assert np.array_equal(z, 0.3 * x)
    H.N_VAbs(X, Z); assert np.array_equal(z, np.abs(x))
    H.N_VInv(W, Z); assert np.array_equal(z, 1.0 / w)
    H.N_VAddConst(X, 2.0, Z); assert np.array_equal(z, x + 2.0)
    assert H.N_VDotProd(X, Y) == pytest.approx(math.fsum(x * y), rel=1e-13)
    assert H.N_VMaxNorm(X) == np.abs(x).max()
    assert H.N_VMin(X) == x.min()
    assert H.N_VWrmsNorm(X, W) == pytest.approx(math.sqrt(math.fsum((x * w) ** 2) / n), rel=1e-13)
    assert H.N_VWL2Norm(X, W) == pytest.approx(math.sqrt(math.fsum((x * w) ** 2)), rel=1e-13)
    assert H.N_VL1Norm(X) == pytest.approx(math.fsum(np.abs(x)), rel=1e-13)
    idm = (rng.random(n) > 0.5).astype(float)
    assert H.N_VWrmsNormMask(X, W, wrap(H, idm)) == pytest.approx(math.sqrt(math.fsum(((x * w) ** 2)[idm > 0]) / n), rel=1e-13)
    H.N_VCompare(0.5, X, Z); assert np.array_equal(z, (np.abs(x) >= 0.5).astype(float))
    assert H.N_VInvTest(W, Z) == 1 and np.array_equal(z, 1.0 / w)
    x0 = x.copy(); x0[17] = 0.0
    assert H.N_VInvTest(wrap(H, x0), Z) == 0
    assert H.N_VMinQuotient(X, W) == (x / w).min()
    c = rng.integers(-2, 3, n).astype(float)
    m = np.zeros(n)
    ok = H.N_VConstrMask(wrap(H, c), X, wrap(H, m))
    viol = ((np.abs(c) == 2) & (x * c <= 0)) | ((np.abs(c) == 1) & (x * c < 0))
    assert np.array_equal(m, viol.astype(float)) and ok == (0 if viol.any() else 1)
