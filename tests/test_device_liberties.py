"""CPU checks (numpy float32) of the three places where the device code of the UNPINNED models departs from the literal
specification that the oracle follows, and of the SASS operand parser the CPU-side loop budget relies on.
The GPU suite holds the kernels themselves to the oracle; these tests pin down how large the departures can be."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


def test_state_recurrence_equals_the_two_cumulative_sums():
    """q += v dt + a dt^2/2; v += a dt   vs   V = qd0 + cumsum(a dt), dq = V_prev dt + a dt^2/2, Q = q0 + cumsum(dq)
    (S/sampling/standard_normal_noise.py:32-50), both in float32, T = 256, sigma = 0.1 rad/s^2 around the home pose."""
    rng = np.random.default_rng(3)
    T, n, dt = 256, 4096, f32(0.01)
    dt2h = f32(0.5) * f32(np.float64(dt) * np.float64(dt))
    a = (rng.standard_normal((T, n)) * 0.1 + rng.standard_normal(n) * 0.5).astype(f32)
    q0 = rng.uniform(-6.0, 6.0, n).astype(f32)
    qd0 = rng.uniform(-0.5, 0.5, n).astype(f32)
    # reference formulation
    cum_v = np.zeros(n, f32); cum_q = np.zeros(n, f32); Q_ref = np.empty((T, n), f32)
    for t in range(T):
        vprev = cum_v + qd0
        dq = vprev * dt + a[t] * dt2h
        cum_v = a[t] * dt + cum_v
        cum_q = cum_q + dq
        Q_ref[t] = cum_q + q0
    # device formulation (each line one fused multiply-add there; numpy rounds the product separately, which only adds noise)
    q = q0.copy(); v = qd0.copy(); Q_dev = np.empty((T, n), f32)
    for t in range(T):
        q = a[t] * dt2h + (v * dt + q)
        v = a[t] * dt + v
        Q_dev[t] = q
    # float64 truth of the same discrete dynamics
    q64 = np.float64(q0).copy(); v64 = np.float64(qd0).copy(); Q64 = np.empty((T, n))
    for t in range(T):
        q64 = q64 + v64 * np.float64(dt) + np.float64(a[t]) * np.float64(dt2h)
        v64 = v64 + np.float64(a[t]) * np.float64(dt)
        Q64[t] = q64
    err_ref = np.abs(Q_ref - Q64).max()
    err_dev = np.abs(Q_dev - Q64).max()
    assert np.abs(Q_dev - Q_ref).max() < 2e-5            # rad, after 256 steps; 64 steps: ~5e-6
    assert np.abs(Q_dev[:64] - Q_ref[:64]).max() < 6e-6
    assert err_dev < 2e-5 and err_ref < 2e-5             # both are roundings of the same dynamics, neither is privileged


def test_whole_turn_offset_keeps_sin_cos():
    """sin / cos of q0 + x against sin / cos of (q0 - 2 pi rint(q0 / 2 pi)) + x, the argument the device feeds the MUFU unit."""
    rng = np.random.default_rng(4)
    q0 = rng.uniform(-60.0, 60.0, 20000).astype(f32)
    x = rng.uniform(-1.0, 1.0, 20000).astype(f32)
    k = np.rint(q0 * f32(0.15915494)).astype(f32)
    fma = lambda a, b, c: (np.float64(a) * np.float64(b) + np.float64(c)).astype(f32)      # noqa: E731 -- one rounding, as on the device
    q0t = fma(k, f32(-6.2831855), q0)                               # first Cody-Waite term
    q0t = fma(k, f32(1.7484555e-7), q0t)                            # second
    assert np.abs(q0t).max() <= np.pi * 1.0001
    arg_ref = (q0 + x).astype(f32)                                   # what the oracle takes sinf / cosf of
    arg_dev = (q0t + x).astype(f32)
    truth = np.float64(q0) + np.float64(x)
    # the device argument is at least as close to the true angle (mod 2 pi) as the wound-up float is
    d_ref = np.abs(np.angle(np.exp(1j * (np.float64(arg_ref) - truth))))
    d_dev = np.abs(np.angle(np.exp(1j * (np.float64(arg_dev) - truth))))
    assert d_dev.max() < 8e-7 and d_ref.max() < 4e-6
    assert np.abs(np.sin(np.float64(arg_dev)) - np.sin(np.float64(arg_ref))).max() < 5e-6


def test_unwrapped_euler_angles_are_the_same_rotation():
    """wrap(a) = a - 2 pi rint(a / 2 pi) changes neither sin nor cos: leaving the angles unwrapped inside the horizon
    only matters through the accuracy of sin / cos at |a| of a few radians."""
    a = np.linspace(-9.0, 9.0, 7001).astype(f32)
    w = (a - f32(6.2831855) * np.rint(a * f32(0.15915494)).astype(f32)).astype(f32)
    assert np.abs(np.sin(np.float64(w)) - np.sin(np.float64(a))).max() < 1e-6
    assert np.abs(np.cos(np.float64(w)) - np.cos(np.float64(a))).max() < 1e-6


def test_sass_operand_parser():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_operand_model as M
    words = lambda t: sum(x[2] for x in M.source_words(t))      # noqa: E731
    assert words("FFMA R11, R15, R11, R2") == 3
    assert words("FFMA R29, R29, UR5, R0.reuse") == 2
    assert words("FFMA2 R8, R8.F32x2.HI_LO, R10.F32x2.HI_LO, R14.F32x2.HI_LO") == 6
    assert words("FFMA2 R16, R16.F32x2.HI_LO, UR5.F32, R0.reuse.F32") == 3
    assert words("FMUL2 R66, R66.F32x2.HI_LO, UR28.F32x2.HI_LO") == 2
    assert words("FADD2 R50, -R90.F32x2.HI_LO, 1.5707963705062866211") == 2
    assert words("MUFU.SIN R72, R104") == 1
    assert words("FMUL.RZ R104, R70, 0.15915493667125701904") == 1
    assert words("LOP3.LUT R51, R80, 0x80000000, R103, 0xb8, !PT") == 2
    assert words("FSETP.GT.AND P1, PT, |R107|, 0.5, PT") == 1
    assert M.opcode_key("IMAD.WIDE.U32") == "IMAD.WIDE" and M.opcode_key("FFMA2") == "FFMA2"
    assert M.source_words("FFMA R29, R29, UR5, R0.reuse")[-1][3] is True          # the reuse flag is seen
