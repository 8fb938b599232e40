"""TEST INFRASTRUCTURE — see oracle/__init__.py."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODELS = {"fhn_torus": 0, "gb_torus": 1, "fhn_flat": 2, "gb_flat": 3}


class OracleParams(C.Structure):
    """Mirror of crd_oracle_params (oracle/crd_oracle.h)."""
    _fields_ = [("model", C.c_int), ("nx", C.c_long), ("ny", C.c_long), ("diff", C.c_double),
                ("beta", C.c_double), ("beta_min", C.c_double), ("beta_max", C.c_double),
                ("vary_beta", C.c_int), ("just_diffusion", C.c_int), ("t_boundary", C.c_double),
                ("surface_length", C.c_double), ("surface_width", C.c_double)]


def make_params(model, nx, ny, diff=0.12, beta=None, beta_min=0.7, beta_max=1.7, vary_beta=None,
                just_diffusion=0, t_boundary=None, surface_length=80.0, surface_width=20.0):
    """Defaults = SURVEY.md §8(d): FHN beta 1.25 / varyBeta 1 / tBoundary 38, Goldbeter beta 0.4 / 0 / 0."""
    m = MODELS[model] if isinstance(model, str) else int(model)
    fhn = m in (0, 2)
    return OracleParams(m, nx, ny, diff, (1.25 if fhn else 0.4) if beta is None else beta, beta_min, beta_max,
                        (1 if fhn else 0) if vary_beta is None else vary_beta, just_diffusion,
                        (38.0 if fhn else 0.0) if t_boundary is None else t_boundary, surface_length, surface_width)


def build(ref="/root/reference"):
    """Compile the checker (and oracle/_ref when the reference tree is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "REF=" + ref, "all"], check=True)


_dp = C.POINTER(C.c_double)


def _ptr(a):
    return a.ctypes.data_as(_dp)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.crd_oracle_rhs.argtypes = [C.POINTER(OracleParams), C.c_double, _dp, _dp]
        L.crd_oracle_rhs_rows.argtypes = [C.POINTER(OracleParams), C.c_double, _dp, _dp, C.c_long, C.c_long]
        L.crd_oracle_rhs_band.argtypes = [C.POINTER(OracleParams), C.c_double, C.c_long, C.c_long, _dp, _dp]
        L.crd_oracle_fill_state.argtypes = [C.c_int, C.c_ulonglong, C.c_long, C.c_long, _dp]
        L.crd_oracle_fill_state.restype = None
        _lib = L
    return _lib


def rhs(P, t, y):
    """ydot of the plain-C restatement; y is the AoS state [ny, nx, 2] (any shape, C order)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(y)
    rc = lib().crd_oracle_rhs(C.byref(P), float(t), _ptr(y), _ptr(out))
    if rc != 0:
        raise RuntimeError("crd_oracle_rhs failed: %d" % rc)
    return out


def rhs_rows(P, t, y, j0, j1):
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty((j1 - j0) * P.nx * 2)
    rc = lib().crd_oracle_rhs_rows(C.byref(P), float(t), _ptr(y), _ptr(out), j0, j1)
    if rc != 0:
        raise RuntimeError("crd_oracle_rhs_rows failed: %d" % rc)
    return out


def fill_state(model, n_elems, seed=0x5EED, first_elem=0):
    m = MODELS[model] if isinstance(model, str) else int(model)
    out = np.empty(n_elems)
    lib().crd_oracle_fill_state(m, seed, first_elem, n_elems, _ptr(out))
    return out


def band_state(model, nx, ny, j0, nrows, seed=0x5EED):
    """Rows j0-1 .. j0+nrows (nrows + 2 rows, periodic in the global mesh) of the synthetic global state [ny][nx][2]."""
    out = np.empty((nrows + 2, 2 * nx))
    lo, hi = j0 - 1, j0 + nrows
    if lo >= 0 and hi < ny:
        return fill_state(model, 2 * nx * (nrows + 2), seed, 2 * nx * lo).reshape(nrows + 2, 2 * nx)
    for r, j in enumerate(range(lo, hi + 1)):
        out[r] = fill_state(model, 2 * nx, seed, 2 * nx * (j % ny))
    return out


def rhs_band(P, t, j0, nrows, yband):
    """Plain-C restatement on a band of phi rows of the global mesh (what one rank of a phi split computes)."""
    yband = np.ascontiguousarray(yband, dtype=np.float64)
    out = np.empty(nrows * P.nx * 2)
    rc = lib().crd_oracle_rhs_band(C.byref(P), float(t), j0, nrows, _ptr(yband), _ptr(out))
    if rc != 0:
        raise RuntimeError("crd_oracle_rhs_band failed: %d" % rc)
    return out


_ref = {}


def ref_available(model):
    name = [k for k, v in MODELS.items() if v == (MODELS[model] if isinstance(model, str) else model)][0]
    return os.path.exists(os.path.join(HERE, "_ref", "libcrd_ref_%s.so" % name))


def ref_lib(model):
    """The compiled-in-place reference for one program (oracle/_ref/libcrd_ref_<model>.so)."""
    name = [k for k, v in MODELS.items() if v == (MODELS[model] if isinstance(model, str) else model)][0]
    if name not in _ref:
        L = C.CDLL(os.path.join(HERE, "_ref", "libcrd_ref_%s.so" % name))
        L.crd_ref_rhs.argtypes = [C.POINTER(OracleParams), C.c_int, C.c_double, _dp, _dp, C.c_int, _dp]
        L.crd_ref_rhs_band.argtypes = [C.POINTER(OracleParams), C.c_double, C.c_long, C.c_long, _dp, _dp]
        L.crd_ref_decomp.argtypes = [C.POINTER(OracleParams), C.c_int, C.c_int, C.POINTER(C.c_long)]
        L.crd_ref_main.argtypes = [C.c_char_p, C.c_int]
        _ref[name] = L
    return _ref[name]


def ref_rhs(P, t, y, nranks=1, reps=1, want_out=True):
    """The reference's own f() on `nranks` emulated MPI ranks. Returns (ydot, seconds)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(y) if want_out else None
    sec = C.c_double(0)
    rc = ref_lib(P.model).crd_ref_rhs(C.byref(P), nranks, float(t), _ptr(y), _ptr(out) if want_out else None,
                                      reps, C.byref(sec))
    if rc != 0:
        raise RuntimeError("reference f() returned %d" % rc)
    return out, sec.value


def ref_rhs_band(P, t, j0, nrows, yband):
    """The reference's own f() (Exchange included) on the band of global phi rows [j0, j0+nrows); yband as band_state gives it."""
    yband = np.ascontiguousarray(yband, dtype=np.float64)
    out = np.empty(nrows * P.nx * 2)
    rc = ref_lib(P.model).crd_ref_rhs_band(C.byref(P), float(t), j0, nrows, _ptr(yband), _ptr(out))
    if rc != 0:
        raise RuntimeError("reference f() on a band returned %d" % rc)
    return out


def ref_decomp(P, nranks, rank):
    out = (C.c_long * 8)()
    ref_lib(P.model).crd_ref_decomp(C.byref(P), nranks, rank, out)
    return dict(zip(("is", "ie", "js", "je", "nxl", "nyl", "dims0", "dims1"), list(out)))
