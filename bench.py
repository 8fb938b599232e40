#!/usr/bin/env python
"""bench.py — grid-point RHS evaluations per second of the FHN-torus hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            (N = 1)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      the reference's own f() on the host cores

A "step" is ONE evaluation ydot = f(t, y) over this rank's phi slab through the C ABI (crd_rhs):
for N > 1 that is the halo exchange with the two ring neighbours + the fused stencil+reaction
kernel; for N = 1 it is exactly one kernel launch.  Workload: BASELINE.json configs[3], FHN on the
torus, synthetic LCG state, theta 16384 x phi 16384 PER GPU (4.29 GB per vector, far larger than the
126 MB L2), phi-split over the N GPUs (global phi mesh 16384*N): weak scaling.  EXACT arithmetic: the
device result is bit-identical to the reference's f() — and the line says so itself: after the timed
region every rank compares rows of its ydot (the slab's first / last rows, whose neighbours live on
another GPU, and random interior rows) with the reference's own f() run on the host for the same
rows (`parity`).

Besides the headline the line carries, at every N: `sustained` (the same launches for >= 2 s), `e2e`
(host buffers, copies inside the timed region, with the pure-copy ceiling measured beside it), `cfg5`
(BASELINE configs[4], Goldbeter torus theta 8192 x 4096 phi rows per GPU), `integrator` (steps/s of
the explicit RK loop on the headline mesh); at N = 1 also the reference's default meshes, the stage
kernels timed alone, and the CPU baseline.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX = 16384
ROWS_PER_GPU = 16384
BYTES_PER_POINT = 32          # read u,v + write u',v' 
T_EVAL = 50.0                 # t > tBoundary: no frozen rows
T_FROZEN = 10.0               # t < tBoundary: global rows 0 and ny-1 held at zero (parity check only)
METRIC = "FHN-torus grid-point RHS evals/sec (fp64)"
UNIT = "point-RHS/s"


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner on fd 1 when NCCL_DEBUG is
    set on the box): keep a private duplicate of fd 1 for the result and point fd 1 at stderr for everything else."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


def workload_config(model, nx, nyl, ny, world, arith="exact"):
    """`config` of the JSON line; the reference arm prints the same dictionary when it runs the same mesh."""
    fhn = model == "fhn_torus"
    return {"workload": "BASELINE configs[%d]: %s torus RHS f(t,y), synthetic LCG state, theta %d x phi %d per GPU "
                        "(global phi %d), phi-split ring of %d GPU(s)" % (3 if fhn else 4, "FHN" if fhn else "Goldbeter", nx, nyl, ny, world),
            "arith": arith + ((" (bit-identical to the reference f())" if fhn else " (reference operation order; libm pow differs by <= a few ulp)")
                              if arith == "exact" else " (<=1e-12)"),
            "nx": nx, "rows_per_gpu": nyl, "ny_global": ny, "t": T_EVAL,
            "l2": "inputs larger than L2 (%.2f GB per vector vs 126 MB)" % (16 * nx * nyl / 1e9),
            "parallelism": "phi-split x%d, P2P halo rows over NVLink" % world}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc = [], None
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:   # nvidia-smi numbers physical GPUs: map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it is a plain list
            ids = [int(x) for x in vis.split(",")] if vis else []
            if device < len(ids):
                device = ids[device]
        except ValueError:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in list(self.rows):
            if ts < t0 or ts > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's threads to the cores of its GPU's NUMA node (read from sysfs) before anything is allocated, so the
    page-locked host buffers of the e2e leg live next to the GPU's PCIe root.  Returns what happened (goes into the line)."""
    out = {"bound": False}
    try:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [int(x) for x in vis.split(",")] if vis and all(x.strip().isdigit() for x in vis.split(",")) else []
        phys = ids[local_rank] if local_rank < len(ids) else local_rank
        r = subprocess.run(["nvidia-smi", "-i", str(phys), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                           capture_output=True, text=True, timeout=20)
        bdf = r.stdout.strip().splitlines()[0].strip().lower()
        if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:      # nvidia-smi prints an 8-digit domain, sysfs uses 4
            bdf = bdf[4:]
        out["pci"] = bdf
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        out["numa_node"] = node
        if node < 0:
            out["why"] = "sysfs reports no NUMA affinity for the device (single-node host or a VM that hides it)"
            return out
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            out["why"] = "none of the node's cores are in this process's affinity mask"
            return out
        os.sched_setaffinity(0, cpus)
        out.update(bound=True, cpus=len(cpus))
    except Exception as e:
        out["why"] = "%s: %s" % (type(e).__name__, str(e)[:120])
    return out


# ---- CHECKER (untimed; the only use of oracle/ besides the CPU baseline): rows of ydot against the reference's f() ----------
def parity_rows(crd, ctx, grid, y, ydot, model, nx, ny, js, je, arith, rank, times=(T_EVAL, T_FROZEN), n_interior=8):
    """Compares rows of this rank's device result with the reference's own f() (oracle/_ref: its Exchange(), stencil and
    kinetics, unmodified) run on the host for the same global rows — the rank's first two and last two rows, whose south /
    north neighbours live on the neighbouring GPU (the checker generates those rows from the LCG stream itself, so a wrong
    halo shows), and n_interior random pairs of interior rows.  Every rank evaluates f once per time (the ring is collective)."""
    import numpy as np
    import oracle as O
    P = O.make_params(model, nx, ny)
    have_ref = O.ref_available(model)
    fn = O.ref_rhs_band if have_ref else O.rhs_band
    rng = np.random.default_rng(77 + rank)
    bands = [(js, 2), (je - 1, 2)] + [(int(j), 2) for j in rng.integers(js + 2, je - 2, n_interior)]
    bitwise, worst, rows = True, 0.0, 0
    # tolerance when the kinetics go through libm pow (Goldbeter) or the arithmetic is FAST: relative to the summed |terms|
    if model.endswith("torus"):
        dx, dy = 2 * np.pi / (nx - 1), 2 * np.pi / (ny - 1)
        r2, Rm = (20.0 / (2 * np.pi)) ** 2, 60.0 / (2 * np.pi)
        scale = 0.12 * 5 * 2.0 * (4.0 / (r2 * dx * dx) + 4.0 / (Rm * Rm * dy * dy)) + 700.0 * 5.0
    else:                                                   # flat: D (uW + uE)/dx^2 + D (uS + uN)/dy^2 - 2 D (1/dx^2 + 1/dy^2) u
        dx, dy = 20.0 / (nx - 1), 80.0 / (ny - 1)
        scale = 0.12 * 2.0 * 8.0 * (1.0 / (dx * dx) + 1.0 / (dy * dy)) + 700.0 * 5.0
    tol = 0.0 if (model.startswith("fhn") and arith == "exact") else (4e-16 if arith == "exact" else 1e-12) * scale
    for t in times:
        grid.f(t, y, ydot)
        for j0, n in bands:
            ref = fn(P, t, j0, n, O.band_state(model, nx, ny, j0, n))
            got = np.empty(2 * nx * n)
            crd._lib.check(crd.lib().crd_memcpy_d2h(ctx._h, got.ctypes.data, ydot.device_ptr + 16 * nx * (j0 - js), got.nbytes), "rows")
            rows += n
            if got.tobytes() != ref.tobytes():
                bitwise = False
                worst = max(worst, float(np.abs(got - ref).max()))
    return {"rows_checked": rows, "bitwise": bitwise, "ok": bool(bitwise or worst <= tol), "max_abs_diff": worst, "tolerance": tol,
            "times": list(times), "checker": "reference f() (oracle/_ref) on the same global rows" if have_ref else "plain-C restatement (oracle/_ref absent)"}


def run_reference(args, rank):
    """--impl reference: the reference's own f() (oracle/_ref, compiled in place from its sources; the plain-C restatement if
    that library is absent) on all host cores, emulated MPI ranks = threads, on the SAME mesh as the GPU arm (theta 16384 x
    phi 16384, one f() per step) when the box's cores finish K + W calls within a few minutes, else on a phi band of it."""
    if rank != 0:
        return
    import oracle as O
    cores = os.cpu_count() or 1
    have_ref = O.ref_available("fhn_torus")
    kind = "reference" if have_ref else "port"
    nranks = cores if have_ref else 1
    # probe the rate on a small band, then take the full mesh if (K + W) calls fit ~150 s
    Pp = O.make_params("fhn_torus", NX, max(64, 2 * nranks))
    yp = O.fill_state("fhn_torus", 2 * NX * Pp.ny)
    if have_ref:
        O.ref_rhs(Pp, T_EVAL, yp, nranks=nranks, reps=1, want_out=False)
        _, sec = O.ref_rhs(Pp, T_EVAL, yp, nranks=nranks, reps=2, want_out=False)
        sec /= 2
    else:
        t0 = time.time(); O.rhs(Pp, T_EVAL, yp); sec = time.time() - t0
    rate = NX * Pp.ny / max(sec, 1e-6)
    calls = max(1, args.steps + args.warmup)
    full = NX * ROWS_PER_GPU * calls / rate <= 150.0
    rows = ROWS_PER_GPU if full else int(min(ROWS_PER_GPU, max(2 * nranks, rate * (100.0 / calls) / NX)))
    P = O.make_params("fhn_torus", NX, rows)
    y = O.fill_state("fhn_torus", 2 * NX * rows)

    def step(reps):
        if have_ref:
            return O.ref_rhs(P, T_EVAL, y, nranks=nranks, reps=reps, want_out=False)[1]
        t0 = time.time()
        for _ in range(reps):
            O.rhs(P, T_EVAL, y)
        return time.time() - t0
    step(max(1, args.warmup))
    sec = step(args.steps)
    value = NX * rows * args.steps / sec
    sample = "FHN torus theta %d x phi %d%s, %d f() calls, %d emulated MPI ranks (threads)%s" % (
        NX, rows, "" if full else " (a phi band of the 16384-row mesh: the full mesh would not fit the time limit on these cores)", args.steps, nranks,
        "; dims from MPI_Dims_create; timing only: the reference's exchange is wrong for >2 ranks per dimension" if nranks > 2 else "")
    config = workload_config("fhn_torus", NX, rows, rows, 1)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nranks, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def cpu_baseline(model="fhn_torus", NX=NX, budget_s=20.0):
    """Bounded CPU sample beside the GPU number (rank 0, N = 1): the reference's f() on all host cores."""
    import oracle as O
    cores = os.cpu_count() or 1
    have_ref = O.ref_available(model)
    nranks = cores if have_ref else 1
    rows = max(256, 2 * nranks)
    P = O.make_params(model, NX, rows)
    y = O.fill_state(model, 2 * NX * rows)
    if have_ref:
        _, sec1 = O.ref_rhs(P, T_EVAL, y, nranks=nranks, reps=1, want_out=False)
        reps = int(max(1, min(200, budget_s / max(sec1, 1e-3) / 2)))
        _, sec = O.ref_rhs(P, T_EVAL, y, nranks=nranks, reps=reps, want_out=False)
        _, s1 = O.ref_rhs(P, T_EVAL, y, nranks=1, reps=1, want_out=False)
        single = NX * rows / s1
    else:
        t0 = time.time(); O.rhs(P, T_EVAL, y); sec1 = time.time() - t0
        reps = int(max(1, min(50, budget_s / max(sec1, 1e-3))))
        t0 = time.time()
        for _ in range(reps):
            O.rhs(P, T_EVAL, y)
        sec = time.time() - t0
        single = NX * rows * reps / sec
    return {"value": NX * rows * reps / sec, "unit": UNIT, "cores": nranks, "kind": "reference" if have_ref else "port",
            "single_core_value": single,
            "sample": "reference f() of %s (oracle/_ref, -O2) on theta %d x phi %d, %d calls, %d emulated MPI ranks on %d host cores"
                      % (model, NX, rows, reps, nranks, cores)}


def default_mesh_integrations(crd, ctx):
    """Integrator steps/s on the meshes the reference ships (data/*.ini: FHN torus 400 x 1600, Goldbeter torus 100 x 400;
    they live in L2, a step is bound by latency, not bandwidth): reference initial conditions, t in [0, tf], rtol 1e-5,
    atol 1e-10, three ways of driving the same method — the device-resident step loop (one persistent kernel per
    ARKode call), the host-driven loop with fused kernels (one launch per stage), and the op-by-op SUNDIALS 2.x
    sequence of N_Vector operations."""
    out = []
    for model, nx, ny, tf in (("fhn_torus", 400, 1600, 2.0), ("gb_torus", 100, 400, 0.5)):
        fhn = model.startswith("fhn")
        beta = 1.25 if fhn else 0.4
        row = {"model": model, "nx": nx, "ny": ny, "t_final": tf}
        # op_by_op re-evaluates f(tn, yn) as stage 1 like ARKode 1.x (6 evaluations per step); the fused loops reuse it (5)
        for name, fused, resident in (("resident", "full", True), ("host_driven_fused", "full", False), ("op_by_op", False, False)):
            g = crd.Grid(ctx, crd.make_params(model, nx, ny, beta=beta, vary_beta=0, t_boundary=0.0))
            y = g.new_vector()
            s0, s1 = (-beta, beta ** 3 - 3 * beta) if fhn else (0.392, 1.6469)
            g.fill_initial_conditions(y, 0.1, 0.5, 1, s0, s1)
            s = crd.ARKodeSolver(g, y, fused=fused, resident=resident)
            s.ARKode(tf * 1e-3); ctx.sync()          # set-up, initial-step estimate, first launch: not timed
            n0 = s.stats()
            l0 = ctx.launches
            t0 = time.time()
            flag, _ = s.ARKode(tf)
            ctx.sync()
            dt = time.time() - t0
            n1 = s.stats()
            row[name] = {"steps_per_s": (n1["nst"] - n0["nst"]) / dt, "us_per_step": 1e6 * dt / max(1, n1["nst"] - n0["nst"]),
                         "nst": n1["nst"] - n0["nst"], "nfe": n1["nfe"] - n0["nfe"], "netf": n1["netf"] - n0["netf"], "flag": flag,
                         "kernel_launches": int(ctx.launches - l0)}
            s.free(); y.destroy(); g.close()
        out.append(row)
    return out


def stage_kernel_times(crd, ctx, grid, y, ydot, model, nx, nyl, reps=20):
    """The kernels a step of the explicit 5-stage method issues on a mesh beyond L2, besides the plain f(tn, yn) of the headline:
    the 2-vector stage f(yn + c F) (three per step) and the last stage fused with the step finish (yn, F1..F4 in; ynew out, F5
    never stored).  Algorithmic bytes: 16 B per point and vector read or written."""
    h = 1e-3
    c5 = [1.0, h * 5 / 32, h * 7 / 32, h * 13 / 32, -h / 32]
    hb = [h / 6, h / 3, h / 3, h / 6, 0.0]
    hd = [h * (1 / 6 + 0.5), h * (1 / 3 - 7 / 3), h * (1 / 3 - 7 / 3), h * (1 / 6 - 13 / 6), h * 16 / 3]
    X = [y]
    try:
        for j in range(1, 5):
            v = grid.new_vector()
            X.append(v)
            grid.fill_synthetic(v, seed=0x5EED + j)
            crd.N_VScale(0.25, v, v)
        out = []
        points = nx * nyl
        for name, nvec, call in (
                ("stage f(yn + c*F): rhs_stream_kernel<NV=2>, 3 CTAs/SM", 3, lambda: grid.f_lincomb(T_EVAL, [1.0, 0.5 * h], X[:2], ydot)),
                ("last stage + finish: rhs_stream_kernel<NV=5, FIN>, 3 CTAs/SM", 6,
                 lambda: grid.f_lincomb_finish(T_EVAL, c5, hb, hd, X, ydot, 1e-5, 1e-10))):
            rc = None
            for _ in range(3):
                rc = call()
            if isinstance(rc, tuple) and rc[0] != 0:      # the fused finish does not apply to this mesh (small slab)
                out.append({"kernel": name, "skipped": "does not apply to this mesh"})
                continue
            ctx.sync(); ctx.timer_start()
            for _ in range(reps):
                call()
            ms = ctx.timer_stop() / reps
            out.append({"kernel": name, "ms": ms, "algorithmic_bytes": 16 * nvec * points, "GBs": 16 * nvec * points / ms / 1e6,
                        "launches_timed": reps, "note": "CUDA events around %d consecutive calls (the fused finish returns its two sums to the host: one stream synchronisation per call is inside)" % reps})
        return out
    finally:
        for v in X[1:]:
            v.destroy()


def nvector_op_times(crd, ctx, grid, y, ydot, peaks, reps=10):
    """The N_Vector operations the explicit integrator calls (SUNDIALS 2.x table order) on vectors of the headline
    mesh, each timed alone with CUDA events over `reps` back-to-back calls through the C ABI.  Algorithmic bytes per element: 8 per
    vector read or written.  A reduction returns its value to the caller, so one stream synchronisation (and, for N > 1, the
    exchange between the GPUs) per call is inside its time."""
    z, w = grid.new_vector(), grid.new_vector()
    try:
        n = y.n
        crd.N_VScale(0.5, y, w)
        crd.N_VAbs(y, z); crd.N_VAddConst(z, 0.5, z)        # z > 0: a weight vector / a safe divisor
        h = 1e-3
        ops = [
            ("N_VLinearSum", 3, lambda: crd.N_VLinearSum(1.5, y, -0.25, z, ydot)),
            ("N_VLinearSum in place (z = a x + z)", 3, lambda: crd.N_VLinearSum(1e-9, y, 1.0, ydot, ydot)),
            ("N_VConst", 1, lambda: crd.N_VConst(0.0, ydot)),
            ("N_VProd", 3, lambda: crd.N_VProd(y, z, ydot)),
            ("N_VDiv", 3, lambda: crd.N_VDiv(y, z, ydot)),
            ("N_VScale", 2, lambda: crd.N_VScale(0.75, y, ydot)),
            ("N_VAbs", 2, lambda: crd.N_VAbs(y, ydot)),
            ("N_VInv", 2, lambda: crd.N_VInv(z, ydot)),
            ("N_VAddConst", 2, lambda: crd.N_VAddConst(y, 1e-10, ydot)),
            ("N_VDotProd", 2, lambda: crd.N_VDotProd(y, z)),
            ("N_VMaxNorm", 1, lambda: crd.N_VMaxNorm(y)),
            ("N_VWrmsNorm", 2, lambda: crd.N_VWrmsNorm(y, z)),
            ("N_VMin", 1, lambda: crd.N_VMin(z)),
            ("N_VWL2Norm", 2, lambda: crd.N_VWL2Norm(y, z)),
            ("N_VL1Norm", 1, lambda: crd.N_VL1Norm(y)),
            ("N_VLinearCombination (3 vectors, fused stage assembly)", 4, lambda: crd.N_VLinearCombination([1.0, 0.5 * h, 0.25 * h], [y, z, w], ydot)),
        ]
        out = []
        for name, nvec, call in ops:
            call(); call()
            ctx.sync(); ctx.timer_start()
            for _ in range(reps):
                call()
            ms = ctx.timer_stop() / reps
            gbs = 8.0 * nvec * n / ms / 1e6
            out.append({"op": name, "ms": ms, "bytes_per_element": 8 * nvec, "GBs": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]})
        return {"elements_per_gpu": n, "launches_timed_per_op": reps, "ops": out,
                "note": "CUDA events around %d consecutive calls of each operation on the headline mesh's vectors (%.2f GB each)" % (reps, 8 * n / 1e9)}
    finally:
        z.destroy(); w.destroy()


def flat_models_block(crd, ctx, peaks, arith, arith_name, steps, warmup):
    """The two flat programs of the reference (BASELINE configs[0] is the FHN flat mesh on the CPU) on a mesh
    beyond L2, one GPU: same timing as the headline, rows of the result against the reference's own f()."""
    out = []
    for model, nx, ny in (("fhn_flat", 16384, 8192), ("gb_flat", 8192, 8192)):
        g = crd.Grid(ctx, crd.make_params(model, nx, ny, arith=arith))
        y, d = g.new_vector(), g.new_vector()
        try:
            g.fill_synthetic(y)
            for _ in range(max(warmup, 5)):
                g.f(T_EVAL, y, d)
            ctx.sync()
            k = max(steps, 50)
            ms = time_rhs(g, ctx, y, d, k)
            row = {"model": model, "nx": nx, "ny": ny, "steps": k, "ms_per_step": ms / k, "value": nx * ny * k / (ms * 1e-3), "unit": UNIT,
                   "frac_of_hbm_peak": BYTES_PER_POINT * nx * ny * k / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
            try:
                row["parity"] = parity_rows(crd, ctx, g, y, d, model, nx, ny, 0, ny - 1, arith_name, 0, n_interior=4)
            except Exception as e:
                row["parity"] = {"ok": False, "error": "%s: %s" % (type(e).__name__, str(e)[:160])}
            out.append(row)
        finally:
            y.destroy(); d.destroy(); g.close()
    return out


def copy_ceiling(torch, nbytes, reps=3, host_in=None, host_out=None):
    """What the PCIe / host-memory path allows with no kernel in between: one H2D and one D2H cudaMemcpyAsync of `nbytes` each,
    from / into page-locked host memory, in flight together on two streams (the traffic of one e2e step).  Timed like the e2e
    leg: one warm-up, then the MEAN over `reps` back-to-back repetitions (the caller takes the max over ranks).
    host_in / host_out: addresses of the e2e leg's own page-locked buffers (same pages, same NUMA placement); fresh ones if absent."""
    import ctypes as C
    n = nbytes // 8

    def wrap(addr):
        t = torch.frombuffer((C.c_double * n).from_address(addr), dtype=torch.float64)
        return t if t.is_pinned() else None
    h_in = wrap(host_in) if host_in else None
    h_out = wrap(host_out) if host_out else None
    if h_in is None or h_out is None:
        h_in = torch.empty(n, dtype=torch.float64, pin_memory=True)
        h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
        h_in.zero_()
    d_in = torch.empty(n, dtype=torch.float64, device="cuda")
    d_out = torch.zeros(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
    once()
    t0 = time.time()
    for _ in range(reps):
        once()
    dt = (time.time() - t0) / reps
    del h_in, h_out, d_in, d_out
    return dt


def integrator_fingerprint(crd, ctx, g, y, world, gloo_group, steps=5):
    """A few steps of the fused EXACT integrator from the synthetic state on the 16384 x 16384 mesh (one slab at N = 1, split
    over the GPUs at N > 1), then a fingerprint of the state that does not depend on the partition: the sum (mod 2^64) and the
    XOR of all elements' bit patterns, with the step counters and the time reached.  EXACT arithmetic makes the integration
    independent of the split bit for bit, so this dictionary must be the same at every N."""
    import numpy as np
    g.fill_synthetic(y)
    solver = crd.ARKodeSolver(g, y, t0=T_EVAL, fused="full", max_steps=steps)
    solver.set_init_step(1e-9)
    flag, t = solver.ARKode(T_EVAL + 1.0, crd.ARK_NORMAL)          # stops at the step limit (flag -1) with the state in y
    st = solver.stats()
    solver.free()
    bits = y.to_numpy().view(np.uint64)
    part = (int(np.add.reduce(bits, dtype=np.uint64)), int(np.bitwise_xor.reduce(bits)))
    del bits
    parts = [part]
    if world > 1:
        import torch.distributed as dist
        parts = [None] * world
        dist.all_gather_object(parts, part, group=gloo_group)
    ssum, sxor = 0, 0
    for a, b in parts:
        ssum = (ssum + a) & 0xFFFFFFFFFFFFFFFF
        sxor ^= b
    g.fill_synthetic(y)
    # (the count of evaluations is left out: one slab prepares the next step's second stage together with f(tn, ynew), a phi-split
    # grid does not, so the last, unused preparation shows up as one more evaluation at N = 1)
    return {"mesh": "theta 16384 x phi 16384 in total", "steps": st["nst"], "netf": st["netf"], "t": float(t).hex(), "flag": flag,
            "sum64": "%016x" % ssum, "xor64": "%016x" % sxor,
            "note": "same dictionary at every N = the phi split does not change a bit of the integration"}


def time_rhs(grid, ctx, y, ydot, steps):
    ctx.timer_start()
    for _ in range(steps):
        grid.f(T_EVAL, y, ydot)
    return ctx.timer_stop()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="crd")
    ap.add_argument("--arith", default="exact", choices=["exact", "fast"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg5"],
                    help="cfg4: FHN torus 16384 x 16384 per GPU (BASELINE configs[3], the headline); "
                         "cfg5: Goldbeter torus theta 8192 x 4096 phi rows per GPU (configs[4], global 8192 x 32768 at 8 GPUs)")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: keep the GLOBAL mesh of the workload (cfg4: 16384 x 16384) and give each GPU 1/N of its rows")
    ap.add_argument("--rows-per-gpu", type=int, default=None)
    ap.add_argument("--nx", type=int, default=None)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-integrator", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline, parity and e2e only (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    claim_stdout()

    if args.impl == "reference":
        run_reference(args, rank)
        return

    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local_rank)

    import numpy as np  # noqa: F401
    import torch
    import crdmodel_b200 as crd
    from crdmodel_b200 import dist as cdist

    torch.cuda.set_device(local_rank)
    use_dist = world > 1
    gloo = None
    if use_dist:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x, op="max"):
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}[op])
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, "max")

    def merge_parity(p):
        """one dictionary for the job: every rank's rows, the worst rank's verdict"""
        p = dict(p)
        p["rows_checked"] = int(reduce_ranks(p["rows_checked"], "sum"))
        p["bitwise"] = bool(reduce_ranks(1.0 if p["bitwise"] else 0.0, "min") > 0.5)
        p["ok"] = bool(reduce_ranks(1.0 if p["ok"] else 0.0, "min") > 0.5)
        p["max_abs_diff"] = reduce_ranks(p["max_abs_diff"], "max")
        p["ranks"] = world
        return p

    model = "fhn_torus" if args.workload == "cfg4" else "gb_torus"
    nx = args.nx or (NX if args.workload == "cfg4" else 8192)
    nyl = args.rows_per_gpu or (ROWS_PER_GPU if args.workload == "cfg4" else 4096)
    if args.strong:
        nyl = (ROWS_PER_GPU if args.workload == "cfg4" else 32768) // world
    ny = nyl * world
    arith = crd.ARITH_EXACT if args.arith == "exact" else crd.ARITH_FAST
    ctx = crd.Context(local_rank)
    if use_dist:
        ctx.set_comm(rank, world, cdist.make_allreduce(gloo))
        cdist.comm_connect(ctx, rank, world, gloo)      # the integrator's norms are finished on the devices (mailboxes over NVLink)
    peaks, peaks_src = measured_peaks()
    sampler = ClockSampler(local_rank) if rank == 0 else None

    def make_grid(model, nx, ny):
        js, je = crd.decomp_phi(ny, world, rank)
        g = crd.Grid(ctx, crd.make_params(model, nx, ny, js=js, je=je, arith=arith))
        if use_dist:
            cdist.ring_connect(g, rank, world, cdist.exchange_handles(g.halo_handle(), gloo))
        return g, js, je

    grid, js, je = make_grid(model, nx, ny)
    y, ydot = grid.new_vector(), grid.new_vector()
    grid.fill_synthetic(y)
    points = nx * nyl

    # ---- device-resident throughput: the headline ------------------------------------------------------
    for _ in range(args.warmup):
        grid.f(T_EVAL, y, ydot)
    ctx.sync()
    barrier()
    launches0 = ctx.launches
    t_wall0 = time.time()
    ms = time_rhs(grid, ctx, y, ydot, args.steps)
    barrier()
    t_wall1 = time.time()
    launches = ctx.launches - launches0
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    value = world * points * args.steps / (ms * 1e-3)

    # ---- the same launches for >= 2 s (the headline region of a short run is a burst; under the power cap the clocks settle
    #      lower): every rank runs the same count, computed from the max-over-ranks time above
    n_sust = int(min(20000, max(args.steps, 2200.0 / max(ms_per_step, 1e-3))))
    barrier()
    ts0 = time.time()
    ms_s = max_over_ranks(time_rhs(grid, ctx, y, ydot, n_sust))
    barrier()
    ts1 = time.time()
    sustained = {"steps": n_sust, "seconds": ms_s * 1e-3, "ms_per_step": ms_s / n_sust, "value": world * points * n_sust / (ms_s * 1e-3),
                 "unit": UNIT, "achieved_GBs_per_gpu": BYTES_PER_POINT * points * n_sust / (ms_s * 1e-3) / 1e9,
                 "frac": BYTES_PER_POINT * points * n_sust / (ms_s * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    clocks = None
    if sampler:
        clocks = sampler.window(t_wall0, t_wall1)
        if clocks["samples"] < 3:    # a 25 ms region falls between two 100 ms samples: the sustained run right behind it is the same load
            clocks = sampler.window(t_wall0, ts1)
            clocks["sampled_over"] = "timed region + the sustained run of the same launches right behind it"
        else:
            clocks["sampled_over"] = "timed region"
        sustained["clocks"] = sampler.window(ts0, ts1)

    # ---- parity of what was just timed, row by row against the reference's f() (untimed) ----------------
    try:
        parity = merge_parity(parity_rows(crd, ctx, grid, y, ydot, model, nx, ny, js, je, args.arith, rank))
    except Exception as e:
        parity = {"ok": False, "error": "%s: %s" % (type(e).__name__, str(e)[:160])}
        if use_dist:       # keep the collective calls of merge_parity matched on this rank
            for op in ("sum", "min", "min", "max"):
                reduce_ranks(0.0, op)

    # ---- end to end: host buffers through crd_rhs_host, H2D + D2H inside the timed region ----------------
    nbytes = 16 * points
    lib = crd.lib()
    hy = lib.crd_malloc_host(nbytes)
    hd = lib.crd_malloc_host(nbytes)
    if not hy or not hd:
        raise crd.CrdError("cannot allocate pinned host buffers")
    crd._lib.check(lib.crd_memcpy_d2h(ctx._h, hy, y.device_ptr, nbytes), "stage host state")
    grid.f_host(T_EVAL, hy, hd)      # warm-up (allocates the staging buffers)
    barrier()
    t0 = time.time()
    for _ in range(args.e2e_steps):
        grid.f_host(T_EVAL, hy, hd)
    ctx.sync()
    e2e_s = max_over_ranks(time.time() - t0)
    e2e_value = world * points * args.e2e_steps / e2e_s
    # the host result of the last e2e step must be the device result (same kernel, same bits)
    import ctypes as C
    probe = np.ctypeslib.as_array(C.cast(hd, C.POINTER(C.c_double)), shape=(2 * nx * min(nyl, 64),))
    ref = np.empty_like(probe)
    grid.f(T_EVAL, y, ydot)
    crd._lib.check(lib.crd_memcpy_d2h(ctx._h, ref.ctypes.data, ydot.device_ptr, ref.nbytes), "read back")
    e2e_ok = bool(probe.tobytes() == ref.tobytes())
    del probe
    # the same bytes, from / into the same host buffers, with no kernel in between, all ranks copying at once: the ceiling of
    # this box's PCIe / host-memory path
    ceiling = None
    try:
        barrier()
        c_s = max_over_ranks(copy_ceiling(torch, nbytes, max(1, args.e2e_steps), hy, hd))
        ceiling = {"ms_per_step": 1e3 * c_s, "GBs_each_way_per_gpu": nbytes / c_s / 1e9,
                   "what": "one cudaMemcpyAsync H2D + one D2H of the step's bytes from / to the e2e leg's own page-locked buffers, concurrently, on all %d rank(s) at once; "
                           "mean over %d repetitions after a warm-up, max over ranks (timed like the e2e leg)" % (world, max(1, args.e2e_steps))}
    except Exception as e:
        ceiling = {"error": str(e)[:160]}
    lib.crd_free_host(hy); lib.crd_free_host(hd)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
           "steps": args.e2e_steps, "ms_per_step": 1e3 * e2e_s / args.e2e_steps, "matches_device_result": e2e_ok,
           "api": "crd_rhs_host (C ABI, pinned host buffers, chunked 3-stream pipeline)", "copy_ceiling": ceiling, "numa": numa}
    if ceiling and "ms_per_step" in ceiling:
        e2e["frac_of_copy_ceiling"] = ceiling["ms_per_step"] / e2e["ms_per_step"]

    extras = not args.no_extras
    # ---- integrator steps/s (fused N_Vector ops + RHS), reported alongside -------------------------------
    def integrator_block(g, label):
        """50 steps of one ARK_NORMAL call (the loop as the drivers run it), then 10 ARK_ONE_STEP calls (each also copies the
        state into the caller's vector) on grid g from the synthetic state."""
        g.fill_synthetic(y)
        solver = crd.ARKodeSolver(g, y, t0=T_EVAL, fused="full", max_steps=50)
        solver.set_init_step(1e-9)
        flag, _ = solver.ARKode(T_EVAL + 1.0, crd.ARK_ONE_STEP)   # set-up + first step
        ctx.sync(); barrier()
        n0 = solver.stats()
        t0 = time.time()
        flag_n, _ = solver.ARKode(T_EVAL + 1.0, crd.ARK_NORMAL)    # returns ARK_TOO_MUCH_WORK (-1) after exactly 50 steps
        ctx.sync()
        dt_n = max_over_ranks(time.time() - t0)
        n1 = solver.stats()
        t0 = time.time()
        for _ in range(10):
            flag, tcur = solver.ARKode(T_EVAL + 1.0, crd.ARK_ONE_STEP)
            if flag < 0:
                break
        ctx.sync()
        dt_1 = max_over_ranks(time.time() - t0)
        n2 = solver.stats()
        solver.free()
        g.fill_synthetic(y)
        att = max(1, n1["nst_attempts"] - n0["nst_attempts"])
        # f(tn, ynew) and the next step's second stage are one pass (crd_rhs_pair): 240 B per point
        bpp = 240
        return {"steps_per_s": (n1["nst"] - n0["nst"]) / dt_n, "step_attempts_per_s": att / dt_n, "ms_per_attempt": 1e3 * dt_n / att,
                "nst": n1["nst"] - n0["nst"], "nfe": n1["nfe"] - n0["nfe"], "netf": n1["netf"] - n0["netf"],
                "rhs_per_attempt": (n1["nfe"] - n0["nfe"]) / att, "flag": flag_n, "mode": "ARK_NORMAL, 50-step limit (flag -1 = the limit, as intended)",
                "bytes_per_point_per_attempt": bpp, "floor_ms_at_measured_peak": bpp * points / (peaks["hbm_gbs"] * 1e6),
                "bytes_note": ("f(tn, ynew) and the next second stage in one pass (1 read + 2 written) + 2 x (2 vectors read + 1 written) + (5 read + 1 written), "
                               "16 B per point and vector" if bpp == 240 else
                               "3 x (2 vectors read + 1 written) + (5 read + 1 written) + f(tn, ynew) (1 read + 1 written), 16 B per point and vector"),
                "one_step_mode": {"steps_per_s": (n2["nst"] - n1["nst"]) / dt_1, "nst": n2["nst"] - n1["nst"], "flag": flag,
                                  "note": "ARK_ONE_STEP returns the state in the caller's vector: one more 32 B/point copy per call"},
                "arith": label}

    integ = integ_fast = None
    if extras and not args.no_integrator:
        method = ("Zonneveld 5-3-4 explicit RK, rtol 1e-5 atol 1e-10, host-driven loop (mesh beyond L2: the resident loop does not apply), stage "
                  "assembly fused into the RHS kernels, last stage fused with the step finish, f(tn, yn) of the previous step reused as stage 1 "
                  "(bit-identical to re-evaluating it as ARKode 1.x does: 5 instead of 6 evaluations per step), norms exchanged between the "
                  "GPUs on the device")
        try:
            integ = integrator_block(grid, args.arith + (": every fused kernel reproduces the bits of the op-by-op N_Vector sequence and the error norm is an "
                                                          "exactly rounded sum; the trajectory is bit-identical to the CPU run's" if args.arith == "exact" else ""))
            integ["method"] = method
        except Exception as e:  # the headline metric does not depend on this block
            integ = {"error": str(e)[:200]}
        if args.arith == "exact":
            # the same loop in FAST arithmetic (fma chains, approximate reciprocal weights; trajectory within rtol |y| + atol)
            try:
                gf = crd.Grid(ctx, crd.make_params(model, nx, ny, js=js, je=je, arith=crd.ARITH_FAST))
                if use_dist:
                    cdist.ring_connect(gf, rank, world, cdist.exchange_handles(gf.halo_handle(), gloo))
                integ_fast = integrator_block(gf, "fast: fused multiply-add chains, <= 1e-12 per evaluation, trajectory within the integrator tolerance")
                gf.close()
            except Exception as e:
                integ_fast = {"error": str(e)[:200]}

    fingerprint = None
    if extras and not args.no_integrator and world == 1 and args.workload == "cfg4" and args.arith == "exact" and nx == NX and nyl == ROWS_PER_GPU:
        try:
            fingerprint = integrator_fingerprint(crd, ctx, grid, y, 1, None)
        except Exception as e:
            fingerprint = {"error": str(e)[:200]}

    # ---- the reference's own default meshes (BASELINE configs[1], [2]): whole adaptive integrations ----
    integ_small = None
    if extras and not args.no_integrator and world == 1:
        try:
            integ_small = default_mesh_integrations(crd, ctx)
        except Exception as e:
            integ_small = {"error": str(e)[:200]}

    # ---- the other kernels of one large-mesh integrator step, each timed alone (CUDA events, back-to-back launches) ----
    stage_kernels = None
    if extras and not args.no_integrator and world == 1:
        try:
            stage_kernels = stage_kernel_times(crd, ctx, grid, y, ydot, model, nx, nyl)
            for k in stage_kernels:
                if "GBs" in k:
                    k["frac_of_hbm_peak"] = k["GBs"] / peaks["hbm_gbs"]
        except Exception as e:
            stage_kernels = {"error": str(e)[:200]}

    # ---- the N_Vector operations of the integrator, each timed alone on the headline mesh's vectors ----
    nvec_ops = None
    if extras and args.workload == "cfg4" and not args.strong:
        try:
            grid.fill_synthetic(y)
            nvec_ops = nvector_op_times(crd, ctx, grid, y, ydot, peaks)
        except Exception as e:
            nvec_ops = {"error": str(e)[:200]}

    y.destroy(); ydot.destroy(); grid.close()

    # ---- the reference's two flat programs (one GPU) ----
    flat = None
    if extras and world == 1 and args.workload == "cfg4":
        try:
            flat = flat_models_block(crd, ctx, peaks, arith, args.arith, args.steps, args.warmup)
        except Exception as e:
            flat = {"error": str(e)[:200]}

    # ---- BASELINE configs[4] beside the headline, at every N: Goldbeter torus theta 8192 x 4096 phi rows per GPU --------
    cfg5 = None
    if extras and args.workload == "cfg4" and not args.strong:
        try:
            nx5, nyl5 = 8192, 4096
            g5, js5, je5 = make_grid("gb_torus", nx5, nyl5 * world)
            y5, d5 = g5.new_vector(), g5.new_vector()
            g5.fill_synthetic(y5)
            for _ in range(max(args.warmup, 5)):
                g5.f(T_EVAL, y5, d5)
            ctx.sync(); barrier()
            k5 = max(args.steps, 200)
            tw0 = time.time()
            ms5 = max_over_ranks(time_rhs(g5, ctx, y5, d5, k5))
            barrier()
            tw1 = time.time()
            n5 = int(min(50000, 2200.0 / max(ms5 / k5, 1e-3)))
            ms5s = max_over_ranks(time_rhs(g5, ctx, y5, d5, n5))
            barrier()
            tw2 = time.time()
            try:
                par5 = merge_parity(parity_rows(crd, ctx, g5, y5, d5, "gb_torus", nx5, nyl5 * world, js5, je5, args.arith, rank))
            except Exception as e:
                par5 = {"ok": False, "error": str(e)[:160]}
                if use_dist:
                    for op in ("sum", "min", "min", "max"):
                        reduce_ranks(0.0, op)
            pts5 = nx5 * nyl5
            cfg5 = {"metric": "Goldbeter-torus grid-point RHS evals/sec (fp64)", "value": world * pts5 * k5 / (ms5 * 1e-3), "unit": UNIT,
                    "steps": k5, "ms_per_step": ms5 / k5, "config": workload_config("gb_torus", nx5, nyl5, nyl5 * world, world, args.arith),
                    "roofline": {"bound": "hbm", "achieved": BYTES_PER_POINT * pts5 * k5 / (ms5 * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": BYTES_PER_POINT * pts5 * k5 / (ms5 * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                    "sustained": {"steps": n5, "seconds": ms5s * 1e-3, "frac": BYTES_PER_POINT * pts5 * n5 / (ms5s * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                    "parity": par5, "clocks": sampler.window(tw0, tw2) if sampler else None}
            y5.destroy(); d5.destroy(); g5.close()
        except Exception as e:
            cfg5 = {"error": str(e)[:200]}
    # ---- strong scaling of the literal BASELINE configs[3] mesh (16384 x 16384 in total, 1/N of its rows per GPU), N > 1 ----
    strong = None
    if extras and world > 1 and args.workload == "cfg4" and not args.strong:
        try:
            gs, jss, jes = make_grid("fhn_torus", NX, ROWS_PER_GPU)
            ys, ds = gs.new_vector(), gs.new_vector()
            gs.fill_synthetic(ys)
            for _ in range(max(args.warmup, 10)):
                gs.f(T_EVAL, ys, ds)
            ctx.sync(); barrier()
            ks = max(args.steps, 200)
            mss = max_over_ranks(time_rhs(gs, ctx, ys, ds, ks))
            barrier()
            nss = int(min(100000, 2200.0 / max(mss / ks, 1e-3)))
            msss = max_over_ranks(time_rhs(gs, ctx, ys, ds, nss))
            barrier()
            try:
                pars = merge_parity(parity_rows(crd, ctx, gs, ys, ds, "fhn_torus", NX, ROWS_PER_GPU, jss, jes, args.arith, rank, n_interior=4))
            except Exception as e:
                pars = {"ok": False, "error": str(e)[:160]}
                if use_dist:
                    for op in ("sum", "min", "min", "max"):
                        reduce_ranks(0.0, op)
            if not args.no_integrator and args.arith == "exact":
                try:
                    fingerprint = integrator_fingerprint(crd, ctx, gs, ys, world, gloo)
                except Exception as e:
                    fingerprint = {"error": str(e)[:200]}
            ptss = NX * ROWS_PER_GPU
            strong = {"what": "the literal BASELINE configs[3] mesh, theta 16384 x phi 16384 in total, %d rows per GPU" % (jes - jss + 1),
                      "value": ptss * ks / (mss * 1e-3), "unit": UNIT, "steps": ks, "ms_per_step": mss / ks,
                      "one_gpu_equivalent_ms": world * mss / ks,
                      "frac_of_hbm_peak_per_gpu": BYTES_PER_POINT * ptss / world * ks / (mss * 1e-3) / 1e9 / peaks["hbm_gbs"],
                      "sustained": {"steps": nss, "seconds": msss * 1e-3, "ms_per_step": msss / nss,
                                    "frac_of_hbm_peak_per_gpu": BYTES_PER_POINT * ptss / world * nss / (msss * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                      "parity": pars,
                      "note": "compare one_gpu_equivalent_ms with ms_per_step of the N = 1 line (same steps count there: --steps)"}
            ys.destroy(); ds.destroy(); gs.close()
        except Exception as e:
            strong = {"error": str(e)[:200]}
    if sampler:
        sampler.stop()

    if rank == 0:
        achieved = BYTES_PER_POINT * points / (ms_per_step * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("rhs_kernel_%s_%s_%dx%d" % (model, args.arith, nx, nyl))
            except Exception:
                traffic = None
        metric = METRIC if model == "fhn_torus" else "Goldbeter-torus grid-point RHS evals/sec (fp64)"
        line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": workload_config(model, nx, nyl, ny, world, args.arith),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peaks_src,
                             "kernel": "rhs_tile_kernel<%s,%s,TX=256,TY=16> (TMA bulk-copy tiles)" % (model.upper(), args.arith), "bytes_per_point": BYTES_PER_POINT,
                             "points_per_launch": points},
                "parity": parity, "sustained": sustained, "e2e": e2e,
                "gpu_launches": int(launches), "clocks": clocks, "integrator": integ, "integrator_fast": integ_fast,
                "integrator_stage_kernels": stage_kernels,
                "integrator_default_meshes": integ_small, "cfg5": cfg5, "strong": strong, "integrator_fingerprint": fingerprint,
                "nvector_ops": nvec_ops, "flat_models": flat}
        if world == 1 and extras and not args.no_cpu_baseline:
            try:
                os.sched_setaffinity(0, all_cpus)
                line["cpu_baseline"] = cpu_baseline(model, nx)
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %s" % str(e)[:120]}
        emit(line)

    ctx.close()
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
