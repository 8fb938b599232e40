"""The per-element arithmetic of the fused integrator operations (crdmodel_b200/csrc/crd_fused.cuh), executed on the host.

tests/fused_harness.cpp compiles the product's own source for the CPU with shims for the intrinsics.  In EXACT ("SEQ")
arithmetic the fused stage assembly and step finish must have the bits of the op-by-op sequence of N_Vector operations the RK
driver issues when nothing is fused (crd_ark.cpp assemble() / compute_solution(); restated here with numpy, whose element-wise
float64 operations are separately rounded); the double-double sums must round to the exact sum (math.fsum) for any grouping of
the terms; the branch-free reciprocal must equal the IEEE quotient for every seed of the hardware's accuracy.  No GPU needed:
an arithmetic regression of the integrator's kernels shows up here first."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D, I, L, P = C.c_double, C.c_int, C.c_long, C.c_void_p


@pytest.fixture(scope="module")
def fh(tmp_path_factory):
    so = tmp_path_factory.mktemp("fused") / "fused_harness.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I/usr/local/cuda/include",
                    "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "crdmodel_b200", "csrc"),
                    os.path.join(ROOT, "tests", "fused_harness.cpp"), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    lib.fh_lc.restype = D
    lib.fh_lc.argtypes = [I, I, P, P]
    lib.fh_finish.argtypes = [I, I, P, P, D, D, L, P, P, P, P, I]
    lib.fh_rcp.restype = D
    lib.fh_rcp.argtypes = [D, D]
    lib.fh_rcp_in_range.argtypes = [D]
    lib.fh_stream_seg_rows.argtypes = [L, L, L]
    lib.fh_dd_sum.restype = D
    lib.fh_dd_sum.argtypes = [P, L, I]
    return lib


def test_stage_combination_has_the_bits_of_the_op_by_op_assembly(fh):
    """assemble() without fused operations: sdata = 0; sdata = c_j F_j + sdata (N_VLinearSum: RN(RN(c x) + y)); z = yn + sdata."""
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 4, 5):
        for trial in range(2000):
            v = rng.uniform(-2, 2, n) * 10.0 ** rng.integers(-3, 3, n)
            c = np.concatenate([[1.0], rng.uniform(-1e-2, 1e-2, n - 1)])
            if trial % 7 == 0:
                v[rng.integers(0, n)] = rng.choice([0.0, -0.0])
            sdata = np.float64(0.0)
            for j in range(1, n):
                sdata = np.float64(np.float64(c[j] * v[j]) + sdata)
            want = np.float64(v[0] + sdata) if n > 1 else np.float64(1.0 * v[0])
            got = fh.fh_lc(1, n, c.ctypes.data, v.ctypes.data)
            assert got == want and math.copysign(1, got) == math.copysign(1, want), (n, trial)
            # the fma chain differs from it by rounding only
            fast = fh.fh_lc(0, n, c.ctypes.data, v.ctypes.data)
            assert abs(fast - want) <= 4e-16 * (abs(want) + np.abs(c * v).sum())


def test_step_finish_has_the_bits_of_the_op_by_op_sequence_and_an_exactly_rounded_error_sum(fh):
    """compute_solution() + ewt_set() + N_VWrmsNorm without fused operations, element by element."""
    rng = np.random.default_rng(2)
    n, S = 20000, 5
    h = 3.7e-4
    b = np.array([1 / 6, 1 / 3, 1 / 3, 1 / 6, 0.0]); b2 = np.array([-0.5, 7 / 3, 7 / 3, 13 / 6, -16 / 3])
    hb, hd = h * b, h * (b - b2)
    rtol, atol = 1e-5, 1e-10
    yn = rng.uniform(-2, 2, n)
    yn[::97] = 0.0
    F = rng.uniform(-50, 50, (S, n))
    # op by op (numpy: every element-wise operation separately rounded)
    ycur = yn.copy()
    tempv = np.zeros(n)
    for j in range(S):
        if hb[j] != 0.0:
            ycur = hb[j] * F[j] + ycur
        tempv = hd[j] * F[j] + tempv
    ewt = 1.0 / (rtol * np.abs(yn) + atol)
    terms = (tempv * ewt) ** 2
    ynew = np.empty(n); sums = np.empty(3)
    fh.fh_finish(1, S, hb.ctypes.data, hd.ctypes.data, rtol, atol, n, yn.ctypes.data, F.ctypes.data, ynew.ctypes.data, sums.ctypes.data, 1)
    assert ynew.tobytes() == ycur.tobytes()
    assert sums[0] + sums[2] == math.fsum(terms)           # double-double sum rounds to the exact sum
    # fused multiply-add arithmetic: same quantities to rounding
    ynew_f = np.empty(n); sums_f = np.empty(3)
    fh.fh_finish(0, S, hb.ctypes.data, hd.ctypes.data, rtol, atol, n, yn.ctypes.data, F.ctypes.data, ynew_f.ctypes.data, sums_f.ctypes.data, 1)
    assert np.abs(ynew_f - ycur).max() <= 1e-15 * (1 + np.abs(ycur).max())
    assert abs(sums_f[0] - math.fsum(terms)) <= 1e-9 * math.fsum(terms)
    assert abs(sums_f[1] - sums[1]) <= 1e-12 * sums[1]


def test_double_double_sum_is_the_exact_sum_for_any_grouping(fh):
    rng = np.random.default_rng(3)
    for n in (1, 2, 1000, 200000):
        q = (rng.standard_normal(n) * 10.0 ** rng.integers(-12, 4, n)) ** 2       # non-negative, 32 orders of magnitude
        want = math.fsum(q)
        for lanes in (1, 2, 32, 256, 1000):
            assert fh.fh_dd_sum(q.ctypes.data, n, lanes) == want, (n, lanes)
        perm = rng.permutation(n)
        assert fh.fh_dd_sum(np.ascontiguousarray(q[perm]).ctypes.data, n, 64) == want


def test_branch_free_reciprocal_equals_the_ieee_quotient(fh):
    """x = rtol |y| + atol in practice; checked over the whole guarded range, with seeds at both ends of the hardware's error
    band, plus the significands where a reciprocal iteration is known to be fragile (all ones, 1 + ulp, powers of two)."""
    rng = np.random.default_rng(4)
    xs = np.concatenate([rng.uniform(1, 2, 200000) * 2.0 ** rng.integers(-400, 400, 200000),
                         1e-5 * np.abs(rng.standard_normal(100000)) + 1e-10,
                         np.nextafter(2.0 ** np.arange(-40, 40), 0), np.nextafter(2.0 ** np.arange(-40, 40), 4e300),
                         2.0 ** np.arange(-40, 40), [1e-10, 1.00001e-10, 3.0, 1 / 3]])
    bad = 0
    for x in xs:
        all_ones = float(x).hex().startswith("0x1.fffffffffffffp")      # Markstein's exception: handed to the IEEE division
        assert fh.fh_rcp_in_range(float(x)) == (0 if all_ones else 1)
        for err in (0.0, 2.0 ** -21, -2.0 ** -21, 2.0 ** -23):
            if fh.fh_rcp(float(x), err) != 1.0 / float(x):
                bad += 1
    assert bad == 0
    # outside the guarded range the IEEE division itself is used
    for x in (1e-300, 1e300, 5e-324, 0.0, float("inf")):
        assert not fh.fh_rcp_in_range(x)


def test_streaming_kernel_segments_fill_whole_waves(fh):
    """stream_seg_rows (crd_grid.cuh): the persistent CTAs take the units (strips x row segments) round-robin, so a launch lasts
    ceil(units / ctas) unit times.  The chosen segment length must keep that within a few per cent of the ideal share of rows per
    CTA (plus the two extra rows a segment fetches) for meshes from one million points up, with segments of at least 16 rows."""
    for ctas in (296, 444):
        for nx, nyl in ((1000, 4000), (1024, 4096), (1400, 5600), (2048, 8192), (4096, 16384), (16384, 2048), (16384, 16384),
                        (8192, 4096), (8192, 32768), (520, 2100), (320, 3400), (16384, 512), (777, 12345)):
            strips = (nx + 255) // 256
            rows = fh.fh_stream_seg_rows(nyl, strips, ctas)
            assert 16 <= rows <= 512
            nseg = -(-nyl // rows)
            waves = -(-strips * nseg // ctas)
            cost = waves * (rows + 2)
            ideal = strips * nyl / ctas
            assert cost <= 1.12 * ideal + 24, (ctas, nx, nyl, rows, cost, ideal)
            fixed = -(-strips * -(-nyl // 128) // ctas) * 130          # the fixed 128-row segments this replaces
            assert cost <= fixed
    assert fh.fh_stream_seg_rows(5, 2, 444) == 5 and fh.fh_stream_seg_rows(100, 1, 444) == 16
