"""Output snapshots (crd_snapshot_*, what the drivers' writer consumes instead of reading N_VGetArrayPointer after every
ARKode call, src/FHNmodel_torus.cpp:438-455): the gathered per-variable arrays are the state's values; an output is ENQUEUED —
the time loop's stream only pays the gather kernel (< 1 ms for a 4096 x 4096 output), the device-to-host copy runs beside the
next steps; slots are handed out and returned in order, and the state may change as soon as begin() has returned."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_snapshot_values_and_slot_protocol(crd, ctx):
    n = 100003
    y = np.random.default_rng(0).standard_normal(2 * n)
    Y = crd.NVector.from_numpy(ctx, y)
    for nvars in (1, 2):
        snap = crd.Snapshot(ctx, n, nvars=nvars, nslots=2)
        a = snap.begin(Y)
        crd.N_VScale(2.0, Y, Y)               # the time loop moves on: the snapshot holds the state as it was at begin()
        b = snap.begin(Y)
        assert (a, b) == (0, 1) and snap.begin(Y) == -2          # both slots held
        v0, v1 = snap.wait(a)
        assert v0.tobytes() == y[0::2].tobytes() and ((v1 is None) if nvars == 1 else (v1.tobytes() == y[1::2].tobytes()))
        w0, w1 = snap.wait(b)
        assert w0.tobytes() == (2.0 * y[0::2]).tobytes()
        snap.release(a)
        assert snap.begin(Y) == a              # a released slot is reused
        snap.release(a); snap.release(b)
        crd.N_VScale(0.5, Y, Y)
        snap.close()


def test_a_4096_squared_output_costs_the_time_loop_under_a_millisecond(crd, ctx):
    nx = ny = 4096
    g = crd.Grid(ctx, crd.make_params("fhn_torus", nx, ny))
    y = g.new_vector()
    g.fill_synthetic(y)
    snap = crd.Snapshot(ctx, nx * ny, nvars=1, nslots=3)
    snap.release(snap.begin(y)); ctx.sync()          # warm-up
    ctx.sync()
    ctx.timer_start()
    t0 = time.perf_counter()
    slot = snap.begin(y)
    host_ms = 1e3 * (time.perf_counter() - t0)
    stream_ms = ctx.timer_stop()                     # what the integrator's stream was busy with: the gather kernel
    v0, _ = snap.wait(slot)
    ref = y.to_numpy()
    assert v0.tobytes() == ref[0::2].tobytes()
    snap.release(slot)
    print("\n4096 x 4096 output: %.3f ms on the integrator's stream, begin() returned after %.3f ms" % (stream_ms, host_ms))
    assert stream_ms < 1.0 and host_ms < 20.0      # (host: generous, the box's cores are shared; measured 0.012 ms)
    snap.close(); g.close()
