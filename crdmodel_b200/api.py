"""Host-side mirror of the reference's operator interface for the hot path, over the C ABI.

Names follow the reference / SUNDIALS (UserData -> Grid, N_Vector ops -> N_V*, ARKode* -> ARKodeSolver)
so tests read like the reference's own call sites (src/FHNmodel_torus.cpp:281,356-373,423,504-667).
Nothing here computes: every method is one call into libcrd_b200.so.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import CrdError, IcParams, Params, check, check_ptr, lib

MODELS = {"fhn_torus": 0, "gb_torus": 1, "fhn_flat": 2, "gb_flat": 3}
MODEL_NAMES = {v: k for k, v in MODELS.items()}
ARITH_EXACT, ARITH_FAST = 0, 1
CRD_SUM, CRD_MAX, CRD_MIN = 1, 2, 3
ARK_NORMAL, ARK_ONE_STEP = 1, 2


def make_params(model, nx, ny, js=None, je=None, arith=ARITH_EXACT, diff=0.12, beta=None, beta_min=0.7,
                beta_max=1.7, vary_beta=None, just_diffusion=0, t_boundary=None, surface_length=80.0,
                surface_width=20.0):
    """crd_params with the synthetic-benchmark defaults of SURVEY.md §8(d)."""
    m = MODELS[model] if isinstance(model, str) else int(model)
    fhn = m in (0, 2)
    return Params(m, arith, nx, ny, 0 if js is None else js, ny - 1 if je is None else je, diff,
                  (1.25 if fhn else 0.4) if beta is None else beta, beta_min, beta_max,
                  (1 if fhn else 0) if vary_beta is None else vary_beta, just_diffusion,
                  (38.0 if fhn else 0.0) if t_boundary is None else t_boundary, surface_length, surface_width)


def decomp_phi(ny, nranks, rank):
    """(js, je) of SetupDecomp (FHNmodel_torus.cpp:752-753) for a 1-D phi split."""
    js, je = C.c_int64(), C.c_int64()
    check(lib().crd_decomp_phi(ny, nranks, rank, C.byref(js), C.byref(je)), "crd_decomp_phi")
    return js.value, je.value


class Context:
    """One GPU + one stream.  stream: a raw cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""

    def __init__(self, device=0, stream=None):
        self._h = check_ptr(lib().crd_ctx_create(device, stream), "crd_ctx_create")
        self._cb = None
        self.device = device
        # everything created from this context (solvers, grids, vectors) is destroyed with it, in that order: the C objects
        # hold a pointer to the context and must not outlive it whatever order Python finalises the wrappers in
        self._children = weakref.WeakSet()

    def _adopt(self, child):
        self._children.add(child)

    def close(self):
        if self._h:
            kids = list(self._children)
            for rank_ in (0, 1, 2):
                for k in kids:
                    if k._close_rank == rank_:
                        k._release()
            lib().crd_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_comm(self, rank, nranks, allreduce):
        """allreduce(values: list[float], op) -> list[float] across ranks (host level)."""
        def _cb(vals, n, op, _user):
            try:
                out = allreduce([vals[i] for i in range(n)], op)
                for i in range(n):
                    vals[i] = out[i]
                return 0
            except Exception:  # never let an exception cross the C boundary
                return -1
        self._cb = _lib.ALLREDUCE_FN(_cb)
        check(lib().crd_ctx_set_comm(self._h, rank, nranks, self._cb, None), "crd_ctx_set_comm")

    def comm_handle(self):
        """64-byte handle of this context's mailbox block for the device-side allreduce."""
        buf = C.create_string_buffer(_lib.HALO_HANDLE_BYTES)
        check(lib().crd_ctx_comm_handle(self._h, buf), "crd_ctx_comm_handle")
        return buf.raw

    def comm_connect_ipc(self, rank, nranks, handles):
        """handles: every rank's comm_handle(), in rank order.  Reductions are finished on the device from then on."""
        blob = b"".join(bytes(h) for h in handles)
        check(lib().crd_ctx_comm_connect_ipc(self._h, rank, nranks, blob), "crd_ctx_comm_connect_ipc")

    def comm_connect_local(self, rank, ctxs):
        """The same for contexts of this process (any devices with peer access)."""
        arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
        check(lib().crd_ctx_comm_connect_local(self._h, rank, len(ctxs), arr), "crd_ctx_comm_connect_local")

    def sync(self):
        check(lib().crd_ctx_sync(self._h), "crd_ctx_sync")

    def set_halo_timeout(self, milliseconds):
        """How long a phi-split evaluation waits for a neighbour's boundary rows before the context fails."""
        check(lib().crd_ctx_set_halo_timeout(self._h, float(milliseconds)), "crd_ctx_set_halo_timeout")

    @property
    def failed(self):
        """0, or the device's error code once a halo wait has timed out (every call then fails until clear_error)."""
        return lib().crd_ctx_failed(self._h)

    def clear_error(self):
        check(lib().crd_ctx_clear_error(self._h), "crd_ctx_clear_error")

    @property
    def launches(self):
        return lib().crd_ctx_launch_count(self._h)

    def timer_start(self):
        check(lib().crd_timer_start(self._h), "crd_timer_start")

    def timer_stop(self):
        ms = C.c_double()
        check(lib().crd_timer_stop(self._h, C.byref(ms)), "crd_timer_stop")
        return ms.value

    def flush_l2(self):
        check(lib().crd_flush_l2(self._h), "crd_flush_l2")

    def malloc(self, nbytes):
        return check_ptr(lib().crd_malloc(self._h, nbytes), "crd_malloc")

    def free(self, ptr):
        check(lib().crd_free(self._h, ptr), "crd_free")

    def fill_synthetic(self, model, n, out_ptr, seed=0x5EED, first_elem=0):
        m = MODELS[model] if isinstance(model, str) else int(model)
        check(lib().crd_fill_synthetic(self._h, m, seed, first_elem, n, out_ptr), "crd_fill_synthetic")


def _ptr(x):
    """device pointer of an NVector, a torch tensor, or a raw int."""
    if isinstance(x, NVector):
        return x.device_ptr
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return x


class NVector:
    """Device-resident N_Vector (N_VNew_Crd).  `.h` is the N_Vector handle for the N_V* functions."""

    def __init__(self, ctx, local_length, global_length=None, _handle=None):
        self.ctx = ctx
        if _handle is not None:
            self.h = _handle
        else:
            g = local_length if global_length is None else global_length
            self.h = check_ptr(lib().N_VNew_Crd(ctx._h, local_length, g), "N_VNew_Crd")
        self.n = lib().N_VGetLocalLength_Crd(self.h)
        ctx._adopt(self)

    _close_rank = 2

    def _release(self):
        self.destroy()

    @classmethod
    def from_numpy(cls, ctx, a, global_length=None):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        v = cls(ctx, a.size, global_length)
        v.set(a)
        return v

    def set(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        assert a.size == self.n
        check(lib().crd_memcpy_h2d(self.ctx._h, self.device_ptr, a.ctypes.data, a.nbytes), "crd_memcpy_h2d")

    def to_numpy(self):
        out = np.empty(self.n)
        if self.n:
            check(lib().crd_memcpy_d2h(self.ctx._h, out.ctypes.data, self.device_ptr, out.nbytes), "crd_memcpy_d2h")
        return out

    @property
    def device_ptr(self):
        return lib().N_VGetDeviceArrayPointer_Crd(self.h)

    def clone(self):
        return NVector(self.ctx, 0, _handle=check_ptr(lib().N_VClone(self.h), "N_VClone"))

    def destroy(self):
        if self.h:
            if self.ctx._h:                 # (a vector that outlived its context was released with it)
                lib().N_VDestroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def _h(v):
    return v.h if isinstance(v, NVector) else v


# the generic SUNDIALS names, dispatching through the ops table like the reference's calls do
def N_VLinearSum(a, x, b, y, z): lib().N_VLinearSum(a, _h(x), b, _h(y), _h(z))
def N_VConst(c, z): lib().N_VConst(c, _h(z))
def N_VProd(x, y, z): lib().N_VProd(_h(x), _h(y), _h(z))
def N_VDiv(x, y, z): lib().N_VDiv(_h(x), _h(y), _h(z))
def N_VScale(c, x, z): lib().N_VScale(c, _h(x), _h(z))
def N_VAbs(x, z): lib().N_VAbs(_h(x), _h(z))
def N_VInv(x, z): lib().N_VInv(_h(x), _h(z))
def N_VAddConst(x, b, z): lib().N_VAddConst(_h(x), b, _h(z))
def N_VDotProd(x, y): return lib().N_VDotProd(_h(x), _h(y))
def N_VMaxNorm(x): return lib().N_VMaxNorm(_h(x))
def N_VWrmsNorm(x, w): return lib().N_VWrmsNorm(_h(x), _h(w))
def N_VWrmsNormMask(x, w, id_): return lib().N_VWrmsNormMask(_h(x), _h(w), _h(id_))
def N_VMin(x): return lib().N_VMin(_h(x))
def N_VWL2Norm(x, w): return lib().N_VWL2Norm(_h(x), _h(w))
def N_VL1Norm(x): return lib().N_VL1Norm(_h(x))
def N_VCompare(c, x, z): lib().N_VCompare(c, _h(x), _h(z))
def N_VInvTest(x, z): return bool(lib().N_VInvTest(_h(x), _h(z)))
def N_VConstrMask(c, x, m): return bool(lib().N_VConstrMask(_h(c), _h(x), _h(m)))
def N_VMinQuotient(num, denom): return lib().N_VMinQuotient(_h(num), _h(denom))


def N_VLinearCombination(c, X, z):
    n = len(c)
    cc = (C.c_double * n)(*c)
    xx = (C.c_void_p * n)(*[_h(x) for x in X])
    check(lib().N_VLinearCombination_Crd(n, cc, xx, _h(z)), "N_VLinearCombination_Crd")


def N_VErkFinish(hb, hd, yn, F, ynew, rtol, atol, exact=False):
    """exact: the bits of the op-by-op sequence (N_VErkFinishSeq_Crd) instead of fused multiply-adds."""
    s = len(hb)
    out = (C.c_double * 2)()
    fn = lib().N_VErkFinishSeq_Crd if exact else lib().N_VErkFinish_Crd
    check(fn(s, (C.c_double * s)(*hb), (C.c_double * s)(*hd), _h(yn),
             (C.c_void_p * s)(*[_h(f) for f in F]), _h(ynew), rtol, atol, out), "N_VErkFinish_Crd")
    return out[0], out[1]


class Grid:
    """One phi slab of one model on one GPU: the device-side UserData (+ ini parameters) of the reference."""

    def __init__(self, ctx, params):
        self.ctx = ctx
        self.params = params
        self._h = check_ptr(lib().crd_grid_create(ctx._h, C.byref(params)), "crd_grid_create")
        self.nx, self.ny = params.nx, params.ny
        self.js, self.je = params.js, params.je
        self.nyl = self.je - self.js + 1
        self.local_length = lib().crd_grid_local_length(self._h)
        self.global_length = lib().crd_grid_global_length(self._h)
        ctx._adopt(self)

    _close_rank = 1

    def _release(self):
        self.close()

    def close(self):
        if self._h:
            if self.ctx._h:
                lib().crd_grid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def new_vector(self):
        return NVector(self.ctx, self.local_length, self.global_length)

    def set_variant(self, v):
        check(lib().crd_grid_set_variant(self._h, v), "crd_grid_set_variant")

    def set_resident(self, mode):
        """Device-resident step loop: 0 automatic (meshes of up to 1 Mi points), 1 whenever it applies, -1 never."""
        check(lib().crd_grid_set_resident(self._h, mode), "crd_grid_set_resident")

    @property
    def resident_launches(self):
        return lib().crd_grid_resident_launches(self._h)

    def resident_cycles(self):
        """SM cycles of the last resident launch by part (one CTA's view): phase1, interior, wait, edges, rest, total."""
        a = (C.c_int64 * 6)()
        check(lib().crd_grid_resident_cycles(self._h, a), "crd_grid_resident_cycles")
        return dict(zip(("phase1", "interior", "wait", "edges", "rest", "total"), list(a)))

    def set_overlap(self, on):
        check(lib().crd_grid_set_overlap(self._h, 1 if on else 0), "crd_grid_set_overlap")

    # f(t, y, ydot, user_data)  — the ARKRhsFn of the reference (FHNmodel_torus.cpp:504)
    def f(self, t, y, ydot):
        check(lib().crd_rhs(self._h, t, _ptr(y), _ptr(ydot)), "crd_rhs")

    def f_lincomb(self, t, c, X, ydot):
        """ydot = f(t, sum_j c[j]*X[j]) without materialising the combination (fused RK stage assembly)."""
        n = len(c)
        check(lib().crd_rhs_lincomb(self._h, t, n, (C.c_double * n)(*c), (C.c_void_p * n)(*[_ptr(x) for x in X]), _ptr(ydot)),
              "crd_rhs_lincomb")

    def f_lincomb_finish(self, t, c, hb, hd, X, ynew, rtol, atol):
        """Last RK stage fused with the step finish: returns (rc, sum (err w)^2, sum (ynew w')^2); rc = 1: does not apply."""
        n = len(c)
        out = (C.c_double * 2)()
        rc = lib().crd_rhs_lincomb_finish(self._h, t, n, (C.c_double * n)(*c), (C.c_double * n)(*hb), (C.c_double * n)(*hd),
                                          (C.c_void_p * n)(*[_ptr(x) for x in X]), _ptr(ynew), rtol, atol, out)
        if rc < 0:
            check(rc, "crd_rhs_lincomb_finish")
        return rc, out[0], out[1]

    def f_pair(self, t1, t2, c, y, f1, f2):
        """f1 = f(t1, y) and f2 = f(t2, y + c f1) in one pass over y; returns 0, or 1 when it does not apply to this grid."""
        rc = lib().crd_rhs_pair(self._h, t1, t2, c, _ptr(y), _ptr(f1), _ptr(f2))
        if rc < 0:
            check(rc, "crd_rhs_pair")
        return rc

    def post_halo(self, y):
        check(lib().crd_rhs_post_halo(self._h, _ptr(y)), "crd_rhs_post_halo")

    def compute(self, t, y, ydot):
        check(lib().crd_rhs_compute(self._h, t, _ptr(y), _ptr(ydot)), "crd_rhs_compute")

    def f_host(self, t, y_host, ydot_host):
        """Host arrays in, host arrays out (numpy arrays or raw pointers to pinned memory)."""
        yp = y_host.ctypes.data if isinstance(y_host, np.ndarray) else y_host
        dp = ydot_host.ctypes.data if isinstance(ydot_host, np.ndarray) else ydot_host
        check(lib().crd_rhs_host(self._h, t, yp, dp), "crd_rhs_host")

    def halo_handle(self):
        buf = C.create_string_buffer(_lib.HALO_HANDLE_BYTES)
        check(lib().crd_grid_halo_handle(self._h, buf), "crd_grid_halo_handle")
        return buf.raw

    def halo_connect_ipc(self, prev_handle, next_handle):
        check(lib().crd_grid_halo_connect_ipc(self._h, prev_handle, next_handle), "crd_grid_halo_connect_ipc")

    def halo_connect_local(self, prev, nxt):
        check(lib().crd_grid_halo_connect_local(self._h, prev._h, nxt._h), "crd_grid_halo_connect_local")

    def fill_initial_conditions(self, y, wave_length, wave_width, wave_inside, s0, s1, ic_type=0):
        ic = IcParams(wave_length, wave_width, wave_inside, ic_type, s0, s1)
        check(lib().crd_fill_initial_conditions(self._h, C.byref(ic), _ptr(y)), "crd_fill_initial_conditions")

    def fill_synthetic(self, y, seed=0x5EED):
        """Slab of the global synthetic state (same stream as oracle.fill_state)."""
        self.ctx.fill_synthetic(self.params.model, self.local_length, _ptr(y), seed, 2 * self.js * self.nx)


class Snapshot:
    """Output snapshots off the time loop's critical path (crd_snapshot_*): begin() enqueues a gather of the written
    variables + an asynchronous copy into a page-locked buffer and returns the slot; wait(slot) blocks until the copy has
    landed and returns numpy views of the values; release(slot) hands the buffer back."""

    _close_rank = 0

    def __init__(self, ctx, npoints, nvars=1, nslots=2):
        self.ctx, self.n, self.nvars = ctx, npoints, nvars
        self._h = check_ptr(lib().crd_snapshot_create(ctx._h, npoints, nvars, nslots), "crd_snapshot_create")
        ctx._adopt(self)

    def begin(self, y):
        slot = lib().crd_snapshot_begin(self._h, _ptr(y))
        if slot == -1:
            check(slot, "crd_snapshot_begin")
        return slot          # -2: every slot is still held

    def wait(self, slot):
        p0, p1 = C.c_void_p(), C.c_void_p()
        check(lib().crd_snapshot_wait(self._h, slot, C.byref(p0), C.byref(p1)), "crd_snapshot_wait")
        v0 = np.ctypeslib.as_array(C.cast(p0, C.POINTER(C.c_double)), shape=(self.n,))
        v1 = np.ctypeslib.as_array(C.cast(p1, C.POINTER(C.c_double)), shape=(self.n,)) if p1.value else None
        return v0, v1

    def release(self, slot):
        check(lib().crd_snapshot_release(self._h, slot), "crd_snapshot_release")

    def close(self):
        if self._h:
            if self.ctx._h:
                lib().crd_snapshot_destroy(self._h)
            self._h = None

    def _release(self):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ARKodeSolver:
    """The reference's ARKode call sequence (FHNmodel_torus.cpp:356-373,423,491) over the device path."""

    def __init__(self, grid, y, t0=0.0, rtol=1e-5, atol=1e-10, max_steps=200000, fused=True, reuse_first_stage=None,
                 resident=True, stage_finish=True, stage_pair=True):
        L = lib()
        self.grid, self.y = grid, y
        self.mem = C.c_void_p(check_ptr(L.ARKodeCreate(), "ARKodeCreate"))
        grid.ctx._adopt(self)
        f = C.cast(L.crd_f, C.c_void_p)
        check(L.ARKodeInit(self.mem, f, None, t0, y.h), "ARKodeInit")
        check(L.ARKodeSStolerances(self.mem, rtol, atol), "ARKodeSStolerances")
        check(L.ARKodeSetUserData(self.mem, grid.handle), "ARKodeSetUserData")
        check(L.ARKodeSetMaxNumSteps(self.mem, max_steps), "ARKodeSetMaxNumSteps")
        # fused: True / "full" = fused vector ops + stage assembly inside the RHS; "ops" = fused vector ops only;
        # False = the op-by-op SUNDIALS 2.x sequence.  The table follows the grid's arithmetic: on an EXACT grid every fused
        # entry reproduces the bits of the op-by-op sequence (crd_nv_fused_ops_exact), on a FAST grid they are fma chains.
        if fused:
            exact = grid.params.arith == ARITH_EXACT
            if fused == "ops":
                table = L.crd_nv_fused_vector_ops_exact() if exact else L.crd_nv_fused_vector_ops()
            else:
                table = L.crd_nv_fused_ops_for(grid.handle)
            check(L.crd_ARKodeSetFusedOps(self.mem, C.cast(table, C.c_void_p)), "crd_ARKodeSetFusedOps")
        # stage 1 of every step is f(tn, yn), which the previous step has just evaluated for its dense output: with the
        # fused operations it is reused by default (bit-identical results, 5 instead of 6 evaluations per step); the
        # op-by-op sequence re-evaluates it like ARKode 1.x does
        if reuse_first_stage is None:
            reuse_first_stage = bool(fused)
        check(L.crd_ARKodeSetReuseFirstStage(self.mem, 1 if reuse_first_stage else 0), "crd_ARKodeSetReuseFirstStage")
        # resident: with the full fused table the whole step loop runs as one persistent kernel when it applies
        # (one GPU, mesh within the grid's size limit: Grid.set_resident); False keeps one launch per stage
        check(L.crd_ARKodeSetResident(self.mem, 1 if resident else 0), "crd_ARKodeSetResident")
        # stage_finish: the last stage and the step finish in one pass over memory where the grid offers it (one GPU, large mesh)
        check(L.crd_ARKodeSetStageFinish(self.mem, 1 if stage_finish else 0), "crd_ARKodeSetStageFinish")
        # stage_pair: f(tn, ynew) of an accepted step and the next step's second stage in one pass where the grid offers it
        check(L.crd_ARKodeSetStagePair(self.mem, 1 if stage_pair else 0), "crd_ARKodeSetStagePair")

    _close_rank = 0

    def _release(self):
        self.free()

    def set_init_step(self, h):
        check(lib().crd_ARKodeSetInitStep(self.mem, h), "crd_ARKodeSetInitStep")

    def set_fixed_step(self, h):
        check(lib().crd_ARKodeSetFixedStep(self.mem, h), "crd_ARKodeSetFixedStep")

    def ARKode(self, tout, itask=ARK_NORMAL):
        t = C.c_double()
        flag = lib().ARKode(self.mem, tout, self.y.h, C.byref(t), itask)
        return flag, t.value

    def stats(self):
        L = lib()
        a, b, c_, d = C.c_long(), C.c_long(), C.c_long(), C.c_long()
        nfi = C.c_long()
        L.ARKodeGetNumSteps(self.mem, C.byref(a))
        L.ARKodeGetNumRhsEvals(self.mem, C.byref(b), C.byref(nfi))
        L.ARKodeGetNumErrTestFails(self.mem, C.byref(c_))
        L.ARKodeGetNumStepAttempts(self.mem, C.byref(d))
        h = C.c_double()
        L.ARKodeGetLastStep(self.mem, C.byref(h))
        return {"nst": a.value, "nfe": b.value, "netf": c_.value, "nst_attempts": d.value, "hlast": h.value}

    def free(self):
        if self.mem:
            if self.grid.ctx._h:
                lib().ARKodeFree(C.byref(self.mem))
            self.mem = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
