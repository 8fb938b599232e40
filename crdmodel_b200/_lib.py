"""ctypes binding of libcrd_b200.so (include/crd_b200.h).  There is no Python or CPU fallback: if the
library is missing, or no B200 is present when a context is created, this raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# CRD_B200_LIB: another build of the same library (kernel experiments); there is still no fallback of any kind
LIB_PATH = os.environ.get("CRD_B200_LIB") or os.path.join(HERE, "lib", "libcrd_b200.so")
# the explicit RK driver + generic N_VXxx dispatchers live in their own host-only library (so that linking the real SUNDIALS
# beside libcrd_b200.so never defines ARKode twice); beside whichever libcrd_b200.so is loaded
ARK_LIB_PATH = os.environ.get("CRD_ARK_LIB") or os.path.join(os.path.dirname(LIB_PATH), "libcrd_ark.so")

c_double_p = C.POINTER(C.c_double)
c_long_p = C.POINTER(C.c_long)
HALO_HANDLE_BYTES = 64


class CrdError(RuntimeError):
    pass


class Params(C.Structure):
    """crd_params: everything f() depends on (ini keys / UserData of the reference)."""
    _fields_ = [("model", C.c_int32), ("arith", C.c_int32), ("nx", C.c_int64), ("ny", C.c_int64),
                ("js", C.c_int64), ("je", C.c_int64), ("diff", C.c_double), ("beta", C.c_double),
                ("beta_min", C.c_double), ("beta_max", C.c_double), ("vary_beta", C.c_int32),
                ("just_diffusion", C.c_int32), ("t_boundary", C.c_double), ("surface_length", C.c_double),
                ("surface_width", C.c_double)]


class IcParams(C.Structure):
    _fields_ = [("wave_length", C.c_double), ("wave_width", C.c_double), ("wave_inside", C.c_int32),
                ("ic_type", C.c_int32), ("s0", C.c_double), ("s1", C.c_double)]


class FusedOps(C.Structure):
    _fields_ = [("lincomb", C.c_void_p), ("erk_finish", C.c_void_p), ("rhs_lincomb", C.c_void_p), ("erk_evolve", C.c_void_p),
                ("rhs_lincomb_finish", C.c_void_p)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, c_double_p, C.c_int, C.c_int, C.c_void_p)
ARK_RHS_FN = C.CFUNCTYPE(C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p)

# name -> (restype, argtypes); every symbol include/crd_b200.h, crd_ark.h and crd_sundials_compat.h declare
V, I, L, D, P = C.c_void_p, C.c_int, C.c_long, C.c_double, C.c_void_p
SIGNATURES = {
    "crd_last_error": (C.c_char_p, []),
    "crd_device_count": (I, []),
    "crd_ctx_create": (P, [I, P]),
    "crd_ctx_destroy": (None, [P]),
    "crd_ctx_set_comm": (I, [P, I, I, ALLREDUCE_FN, P]),
    "crd_ctx_comm_handle": (I, [P, C.c_char_p]),
    "crd_ctx_comm_connect_ipc": (I, [P, I, I, C.c_char_p]),
    "crd_ctx_comm_connect_local": (I, [P, I, I, C.POINTER(P)]),
    "crd_ctx_stream": (P, [P]),
    "crd_ctx_device": (I, [P]),
    "crd_ctx_sync": (I, [P]),
    "crd_ctx_set_halo_timeout": (I, [P, D]),
    "crd_ctx_failed": (I, [P]),
    "crd_ctx_clear_error": (I, [P]),
    "crd_ctx_launch_count": (C.c_int64, [P]),
    "crd_timer_start": (I, [P]),
    "crd_timer_stop": (I, [P, c_double_p]),
    "crd_malloc": (P, [P, C.c_size_t]),
    "crd_free": (I, [P, P]),
    "crd_malloc_host": (P, [C.c_size_t]),
    "crd_free_host": (I, [P]),
    "crd_memcpy_h2d": (I, [P, P, P, C.c_size_t]),
    "crd_memcpy_d2h": (I, [P, P, P, C.c_size_t]),
    "crd_memset_zero": (I, [P, P, C.c_size_t]),
    "crd_flush_l2": (I, [P]),
    "crd_decomp_phi": (I, [C.c_int64, I, I, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "crd_grid_create": (P, [P, C.POINTER(Params)]),
    "crd_grid_destroy": (None, [P]),
    "crd_grid_params": (I, [P, C.POINTER(Params)]),
    "crd_grid_local_length": (C.c_int64, [P]),
    "crd_grid_global_length": (C.c_int64, [P]),
    "crd_grid_dx": (D, [P]),
    "crd_grid_dy": (D, [P]),
    "crd_grid_halo_handle": (I, [P, C.c_char_p]),
    "crd_grid_halo_connect_ipc": (I, [P, C.c_char_p, C.c_char_p]),
    "crd_grid_halo_connect_local": (I, [P, P, P]),
    "crd_rhs": (I, [P, D, P, P]),
    "crd_rhs_post_halo": (I, [P, P]),
    "crd_rhs_compute": (I, [P, D, P, P]),
    "crd_rhs_host": (I, [P, D, P, P]),
    "crd_f": (I, [D, P, P, P]),
    "crd_rhs_lincomb": (I, [P, D, I, c_double_p, C.POINTER(P), P]),
    "crd_f_lincomb": (I, [D, I, c_double_p, C.POINTER(P), P, P]),
    "crd_rhs_lincomb_finish": (I, [P, D, I, c_double_p, c_double_p, c_double_p, C.POINTER(P), P, D, D, c_double_p]),
    "crd_f_lincomb_finish": (I, [D, I, c_double_p, c_double_p, c_double_p, C.POINTER(P), P, D, D, c_double_p, P]),
    "crd_rhs_pair": (I, [P, D, D, D, P, P, P]),
    "crd_f_pair": (I, [D, D, D, P, P, P, P]),
    "crd_grid_rhs_count": (C.c_int64, [P]),
    "crd_grid_set_variant": (I, [P, I]),
    "crd_grid_set_overlap": (I, [P, I]),
    "crd_erk_evolve": (I, [P, P]),
    "crd_grid_set_resident": (I, [P, I]),
    "crd_grid_resident_launches": (C.c_int64, [P]),
    "crd_grid_resident_cycles": (I, [P, C.POINTER(C.c_int64)]),
    "crd_fill_synthetic": (I, [P, I, C.c_uint64, C.c_int64, C.c_int64, P]),
    "crd_fill_initial_conditions": (I, [P, C.POINTER(IcParams), P]),
    "crd_snapshot_create": (P, [P, C.c_int64, I, I]),
    "crd_snapshot_destroy": (None, [P]),
    "crd_snapshot_begin": (I, [P, P]),
    "crd_snapshot_wait": (I, [P, I, C.POINTER(P), C.POINTER(P)]),
    "crd_snapshot_release": (I, [P, I]),
    # device N_Vector
    "N_VNew_Crd": (P, [P, L, L]),
    "N_VNewEmpty_Crd": (P, [P, L, L]),
    "N_VMake_Crd": (P, [P, L, L, P]),
    "N_VDestroy_Crd": (None, [P]),
    "N_VGetDeviceArrayPointer_Crd": (P, [P]),
    "N_VCopyToHost_Crd": (I, [P]),
    "N_VCopyFromHost_Crd": (I, [P]),
    "N_VGetLocalLength_Crd": (L, [P]),
    "N_VGetContext_Crd": (P, [P]),
    "N_VClone_Crd": (P, [P]),
    "N_VCloneEmpty_Crd": (P, [P]),
    "N_VSpace_Crd": (None, [P, c_long_p, c_long_p]),
    "N_VGetArrayPointer_Crd": (P, [P]),
    "N_VSetArrayPointer_Crd": (None, [P, P]),
    "N_VLinearSum_Crd": (None, [D, P, D, P, P]),
    "N_VConst_Crd": (None, [D, P]),
    "N_VProd_Crd": (None, [P, P, P]),
    "N_VDiv_Crd": (None, [P, P, P]),
    "N_VScale_Crd": (None, [D, P, P]),
    "N_VAbs_Crd": (None, [P, P]),
    "N_VInv_Crd": (None, [P, P]),
    "N_VAddConst_Crd": (None, [P, D, P]),
    "N_VDotProd_Crd": (D, [P, P]),
    "N_VMaxNorm_Crd": (D, [P]),
    "N_VWrmsNorm_Crd": (D, [P, P]),
    "N_VWrmsNormMask_Crd": (D, [P, P, P]),
    "N_VMin_Crd": (D, [P]),
    "N_VWL2Norm_Crd": (D, [P, P]),
    "N_VL1Norm_Crd": (D, [P]),
    "N_VCompare_Crd": (None, [D, P, P]),
    "N_VInvTest_Crd": (I, [P, P]),
    "N_VConstrMask_Crd": (I, [P, P, P]),
    "N_VMinQuotient_Crd": (D, [P, P]),
    "N_VLinearCombination_Crd": (I, [I, c_double_p, C.POINTER(P), P]),
    "N_VErkFinish_Crd": (I, [I, c_double_p, c_double_p, P, C.POINTER(P), P, D, D, c_double_p]),
    "N_VErkFinishSeq_Crd": (I, [I, c_double_p, c_double_p, P, C.POINTER(P), P, D, D, c_double_p]),
    "crd_nv_fused_ops": (C.POINTER(FusedOps), []),
    "crd_nv_fused_vector_ops": (C.POINTER(FusedOps), []),
    "crd_nv_fused_ops_exact": (C.POINTER(FusedOps), []),
    "crd_nv_fused_vector_ops_exact": (C.POINTER(FusedOps), []),
    "crd_nv_fused_ops_for": (C.POINTER(FusedOps), [P]),
    # generic dispatchers + ARKode-legacy interface (crd_sundials_compat.h, crd_ark.h)
    "N_VClone": (P, [P]),
    "N_VDestroy": (None, [P]),
    "N_VGetArrayPointer": (P, [P]),
    "N_VLinearSum": (None, [D, P, D, P, P]),
    "N_VConst": (None, [D, P]),
    "N_VProd": (None, [P, P, P]),
    "N_VDiv": (None, [P, P, P]),
    "N_VScale": (None, [D, P, P]),
    "N_VAbs": (None, [P, P]),
    "N_VInv": (None, [P, P]),
    "N_VAddConst": (None, [P, D, P]),
    "N_VDotProd": (D, [P, P]),
    "N_VMaxNorm": (D, [P]),
    "N_VWrmsNorm": (D, [P, P]),
    "N_VWrmsNormMask": (D, [P, P, P]),
    "N_VMin": (D, [P]),
    "N_VWL2Norm": (D, [P, P]),
    "N_VL1Norm": (D, [P]),
    "N_VCompare": (None, [D, P, P]),
    "N_VInvTest": (I, [P, P]),
    "N_VConstrMask": (I, [P, P, P]),
    "N_VMinQuotient": (D, [P, P]),
    "ARKodeCreate": (P, []),
    "ARKodeInit": (I, [P, P, P, D, P]),
    "ARKodeSStolerances": (I, [P, D, D]),
    "ARKodeSetUserData": (I, [P, P]),
    "ARKodeSetMaxNumSteps": (I, [P, L]),
    "ARKode": (I, [P, D, P, c_double_p, I]),
    "ARKodeFree": (None, [C.POINTER(P)]),
    "ARKodeGetNumSteps": (I, [P, c_long_p]),
    "ARKodeGetNumRhsEvals": (I, [P, c_long_p, c_long_p]),
    "ARKodeGetNumErrTestFails": (I, [P, c_long_p]),
    "ARKodeGetNumStepAttempts": (I, [P, c_long_p]),
    "ARKodeGetCurrentStep": (I, [P, c_double_p]),
    "ARKodeGetLastStep": (I, [P, c_double_p]),
    "ARKodeGetCurrentTime": (I, [P, c_double_p]),
    "crd_ARKodeSetFusedOps": (I, [P, P]),
    "crd_ARKodeSetReuseFirstStage": (I, [P, I]),
    "crd_ARKodeSetResident": (I, [P, I]),
    "crd_ARKodeSetStageFinish": (I, [P, I]),
    "crd_ARKodeSetStagePair": (I, [P, I]),
    "crd_ARKodeSetInitStep": (I, [P, D]),
    "crd_ARKodeSetFixedStep": (I, [P, D]),
    "crd_ARKodeGetButcherTable": (I, [P, P, P, P, c_double_p, c_double_p, c_double_p, c_double_p]),
}

_lib = None


class _Libs:
    """Both libraries behind one namespace: libcrd_b200.so (device path, include/crd_b200.h) and libcrd_ark.so (ARKode-legacy
    names, generic N_VXxx; include/crd_sundials_compat.h, crd_ark.h)."""

    def __init__(self, handles):
        self._handles = handles

    def __getattr__(self, name):
        for h in self._handles:
            try:
                fn = getattr(h, name)
            except AttributeError:
                continue
            setattr(self, name, fn)
            return fn
        raise AttributeError(name)


def which_library(name):
    """'b200' / 'ark': the library that exports `name` (tests)."""
    L = lib()
    for tag, h in zip(("b200", "ark"), L._handles):
        if hasattr(h, name):
            return tag
    return None


def bind(libs, signatures=SIGNATURES):
    for name, (res, args) in signatures.items():
        fn = getattr(libs, name)  # AttributeError if neither library exports the symbol
        fn.restype = res
        fn.argtypes = args
    return libs


def lib():
    """The loaded libraries.  Raises CrdError if they have not been built (python -m crdmodel_b200.build)."""
    global _lib
    if _lib is None:
        for p in (LIB_PATH, ARK_LIB_PATH):
            if not os.path.exists(p):
                raise CrdError("%s is missing: build it with `python -m crdmodel_b200.build` "
                               "(there is no CPU fallback)" % p)
        _lib = bind(_Libs([C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL), C.CDLL(ARK_LIB_PATH, mode=C.RTLD_GLOBAL)]))
    return _lib


def last_error():
    return lib().crd_last_error().decode()


def check(rc, what):
    if rc != 0:
        raise CrdError("%s failed (%d): %s" % (what, rc, last_error()))


def check_ptr(p, what):
    if not p:
        raise CrdError("%s failed: %s" % (what, last_error()))
    return p
