"""Error behaviour at the C ABI (include/crd_b200.h: int flags, negative = failure, NULL = allocation / argument failure,
crd_last_error() holds the reason; ARKode-legacy flags as in the reference's check_flag, src/FHNmodel_torus.cpp:681-705).
No exception and no crash may cross the boundary: the calls run in a child process so that a crash would be a test failure."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r"""
import ctypes as C, sys
sys.path.insert(0, %r)
import crdmodel_b200 as crd
L = crd.lib()
L.crd_last_error.restype = C.c_char_p
js, je = C.c_int64(), C.c_int64()
assert L.crd_decomp_phi(10, 0, 0, C.byref(js), C.byref(je)) == -1 and b"crd_decomp_phi" in L.crd_last_error()
assert L.crd_decomp_phi(3, 4, 0, C.byref(js), C.byref(je)) == -1          # fewer rows than ranks
assert L.crd_decomp_phi(10, 2, 2, C.byref(js), C.byref(je)) == -1         # rank out of range
L.crd_grid_create.restype = C.c_void_p
assert L.crd_grid_create(None, None) is None
assert L.crd_f(C.c_double(0.0), None, None, None) == -1                   # ARKRhsFn: < 0 = unrecoverable
assert L.crd_rhs_host(None, C.c_double(0.0), None, None) == -1 and b"crd_rhs_host" in L.crd_last_error()
L.N_VNew_Crd.restype = C.c_void_p
assert L.N_VNew_Crd(None, 10, 10) is None                                 # NULL = MEMORY_ERROR for check_flag(opt 0)
L.crd_grid_destroy(None)                                                  # destructors accept NULL
ARK_MEM_NULL, ARK_ILL_INPUT, ARK_NO_MALLOC = -21, -22, -23
L.ARKodeCreate.restype = C.c_void_p
assert L.ARKodeInit(None, None, None, C.c_double(0.0), None) == ARK_MEM_NULL
m = C.c_void_p(L.ARKodeCreate())
assert m.value
t = C.c_double()
assert L.ARKode(m, C.c_double(1.0), None, C.byref(t), 1) == ARK_NO_MALLOC
assert L.ARKodeSStolerances(m, C.c_double(1e-5), C.c_double(1e-10)) == ARK_NO_MALLOC
assert L.ARKodeInit(m, None, None, C.c_double(0.0), None) == ARK_ILL_INPUT
assert L.ARKodeSetMaxNumSteps(None, 10) == ARK_MEM_NULL and L.ARKodeSetUserData(None, None) == ARK_MEM_NULL
L.ARKodeFree(C.byref(m))
assert m.value is None
L.ARKodeFree(C.byref(m))                                                  # twice is harmless
if L.crd_device_count() == 0:
    L.crd_ctx_create.restype = C.c_void_p
    assert L.crd_ctx_create(0, None) is None and len(L.crd_last_error()) > 0   # no device: fail loudly, no CPU fallback
print("ABI_ERRORS_OK")
"""


def test_error_paths_return_flags_and_never_crash(crd):
    r = subprocess.run([sys.executable, "-c", SNIPPET % ROOT], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ABI_ERRORS_OK" in r.stdout, (r.returncode, r.stdout, r.stderr[-2000:])
