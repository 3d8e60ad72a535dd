"""ctypes front-end of the CPU oracle (`oracle/mppi_oracle.c`).

TEST INFRASTRUCTURE ONLY -- see the header of mppi_oracle.c.  Imported by tests/,
`__graft_entry__.smoke()` and the cpu_baseline / `--impl reference` legs of bench.py;
never by the product package.

Everything numeric happens in C; this module only marshals numpy buffers and composes
the per-stage functions into whole control steps the way
`mppi_solver/mppi.py:122-169` and `mppi_solver/drone_mppi.py:140-176` do.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmppi_oracle.so")
_lib = None

f32 = np.float32
_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile the C restatement with the system gcc (see oracle/Makefile)."""
    src = os.path.join(_HERE, "mppi_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        r = subprocess.run(["make", "-C", _HERE, "-B", "libmppi_oracle.so"], capture_output=True, text=True)
        if r.returncode != 0:   # no libgomp on this host: single-threaded oracle
            r = subprocess.run(["make", "-C", _HERE, "-B", "OMP=0", "libmppi_oracle.so"],
                               capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_set_threads.restype = C.c_int
    return _lib


def set_threads(n: int) -> int:
    return lib().oracle_set_threads(int(n))


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_fp)


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


# --------------------------------------------------------------------------- chain
class Chain:
    """Joint chain in the order `URDFparser._get_joint_chain` yields it (urdfparser.py:110-120)."""

    def __init__(self, jtype, qidx, xyz, rpy, axis):
        self.jtype = np.asarray(jtype, np.int32)
        self.qidx = np.asarray(qidx, np.int32)
        self.xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
        self.rpy = np.asarray(rpy, np.float32).reshape(-1, 3)
        self.axis = np.asarray(axis, np.float32).reshape(-1, 3)
        self.n = len(self.jtype)


_PI = math.pi
_H = math.pi / 2
# world -> j2s7s300_link_7 of aerial_manipulation/urdf/aerial_manipulator_gpu.urdf
# (joint_base :67-74 fixed; joint_1..7 :100-106,143-149,186-192,229-235,272-278,315-321,358-364).
# root_link "base" does not exist in that URDF, so the walk ends at the absolute root "world"
# (urdfparser.py:65) and the fixed joint_base is part of the chain.  No end-effector offset:
# the end link is link_7 (mppi.py:84-88).
KINOVA_CHAIN = Chain(
    jtype=[0, 1, 1, 1, 1, 1, 1, 1],
    qidx=[-1, 0, 1, 2, 3, 4, 5, 6],
    xyz=[[0, 0, 0], [0, 0, 0.15675], [0, 0.0016, -0.11875], [0, -0.205, 0], [0, 0, -0.205],
         [0, 0.2073, -0.0114], [0, 0, -0.10375], [0, 0.10375, 0]],
    rpy=[[_PI, 0, 0], [0, _PI, 0], [-_H, 0, _PI], [-_H, 0, 0], [_H, 0, _PI], [-_H, 0, _PI],
         [_H, 0, _PI], [-_H, 0, _PI]],
    axis=[[0, 0, 0]] + [[0, 0, 1]] * 7,
)

ARM_WEIGHTS = (50.0, 30.0, 40.0, 30.0)          # cost/cost_manager.py:30-33
DRONE_WEIGHTS = (100.0, 20.0)                   # mppi_solver/drone_mppi.py:93,105
ARM_TARGET_POS = (0.1029, 0.4055, 1.6498)       # mppi_solver/mppi.py:71
ARM_TARGET_QUAT = (-0.5, -0.5, 0.5, -0.5)       # mppi_solver/mppi.py:72 (xyzw)
DRONE_TARGET = (1.0, 2.0, 3.4)                  # mppi_solver/drone_mppi.py:141
Q_HOME = (1.57, 1.7, 0.0, 4.4, 0.0, 4.71, 0.0)  # kinova.py:135
# quad parameters: mass, 1/Ixx, 1/Iyy, 1/Izz, k_d, g_z
# (aerial_manipulation/src/controller.cpp:159-161,488-490; k_d is undefined in the draft -> 0)
QUAD_PARAMS = (14.7, 1.0 / 1.57, 1.0 / 3.93, 1.0 / 2.59, 0.0, -9.81)
WB_PARAMS = (14.7 + 5.5, 1.0 / 1.57, 1.0 / 3.93, 1.0 / 2.59, 0.0, -9.81)


def fk(q, chain: Chain = KINOVA_CHAIN) -> np.ndarray:
    q, qp = _f(q)
    T = np.zeros(16, f32)
    lib().oracle_fk(chain.n, _i(chain.jtype)[1], _i(chain.qidx)[1], _f(chain.xyz)[1], _f(chain.rpy)[1],
                    _f(chain.axis)[1], qp, int(q.size), T.ctypes.data_as(_fp))
    return T.reshape(4, 4)


def lib_make_transform(xyz, rpy) -> np.ndarray:
    """URDF origin -> 4x4 (robot/transformation_matrix.py:4-35)."""
    T = np.zeros(16, f32)
    lib().oracle_make_transform(_f(xyz)[1], _f(rpy)[1], T.ctypes.data_as(_fp))
    return T.reshape(4, 4)


def xyzquat_to_matrix(b) -> np.ndarray:
    T = np.zeros(16, f32)
    lib().oracle_xyzquat_to_matrix(_f(b)[1], T.ctypes.data_as(_fp))
    return T.reshape(4, 4)


def quaternion_to_matrix(q) -> np.ndarray:
    R = np.zeros(9, f32)
    lib().oracle_quaternion_to_matrix(_f(q)[1], R.ctypes.data_as(_fp))
    return R.reshape(3, 3)


def matrix_to_euler_zyx(M) -> np.ndarray:
    e = np.zeros(3, f32)
    lib().oracle_matrix_to_euler_zyx(_f(M)[1], e.ctypes.data_as(_fp))
    return e


# --------------------------------------------------------------------------- costs
def arm_costs(noise_tkn, u_nom, q0, qd0, base, target_pos=ARM_TARGET_POS, target_quat=ARM_TARGET_QUAT,
              dt=0.01, weights=ARM_WEIGHTS, chain: Chain = KINOVA_CHAIN, state_f64=False) -> np.ndarray:
    noise, npt = _f(noise_tkn)
    T, K, nu = noise.shape
    assert nu == 7
    S = np.zeros(K, f32)
    if not state_f64:   # float32 state tensors: round the state first
        q0 = np.asarray(q0, f32)
        qd0 = np.asarray(qd0, f32)
    lib().oracle_arm_costs(K, T, npt, _f(u_nom)[1], _d(q0)[1], _d(qd0)[1], _f(base)[1],
                           chain.n, _i(chain.jtype)[1], _i(chain.qidx)[1], _f(chain.xyz)[1],
                           _f(chain.rpy)[1], _f(chain.axis)[1], _f(target_pos)[1], _f(target_quat)[1],
                           C.c_float(dt), _f(weights)[1], int(bool(state_f64)), S.ctypes.data_as(_fp))
    return S


Q_CENTER = (0.0, 0.0, 0.0, (-3.0718 - 0.0698) / 2, 0.0, (3.7525 - 0.0175) / 2, 0.0)      # cost/joint_space_cost.py:13
Q_LOWER = (-6.2832, 0.8203, -6.2832, 0.5236, -6.2832, 1.1345, -6.2832)                   # :61
Q_UPPER = (6.2832, 5.4629, 6.2832, 5.7596, 6.2832, 5.1487, 6.2832)                       # :62
COST_COVAR, COST_CENTERING, COST_JOINT_TRAJ, COST_ACTION, COST_JOINT_LIMIT = 1, 2, 4, 8, 16


def arm_extra_costs(noise_tkn, u_nom, q0, qd0, flags, dt=0.01, lam=0.1, gamma=0.98, covar_weight=0.1, alpha=0.1,
                    action_weight=0.01, centering_weight=1.0, joint_traj_weight=1.0, limit_penalty=1e10,
                    sigma=0.1, q_traj=None) -> np.ndarray:
    """Sum of the reference's optional cost terms selected by `flags` (cost_manager.py:83-87)."""
    noise, npt = _f(noise_tkn)
    T, K, nu = noise.shape
    assert nu == 7
    S = np.zeros(K, f32)
    ext = np.array([gamma, covar_weight, lam, alpha, action_weight, centering_weight, joint_traj_weight, limit_penalty], f32)
    sig = np.broadcast_to(np.asarray(sigma, f32), (7,)).copy()
    lib().oracle_arm_extra_costs(K, T, npt, _f(u_nom)[1], _f(q0)[1], _f(qd0)[1], C.c_float(dt), int(flags), _f(ext)[1],
                                 _f(sig)[1], _f(Q_CENTER)[1], _f(Q_LOWER)[1], _f(Q_UPPER)[1],
                                 None if q_traj is None else _f(q_traj)[1], S.ctypes.data_as(_fp))
    return S


def drone_costs(noise_tkn, u_nom, x0, v0, target=DRONE_TARGET, dt=0.01, weights=DRONE_WEIGHTS) -> np.ndarray:
    noise, npt = _f(noise_tkn)
    T, K, nu = noise.shape
    assert nu == 3
    S = np.zeros(K, f32)
    lib().oracle_drone_costs(K, T, npt, _f(u_nom)[1], _f(x0)[1], _f(v0)[1], _f(target)[1],
                             C.c_float(dt), _f(weights)[1], S.ctypes.data_as(_fp))
    return S


def quad_costs(noise_tkn, u_nom, state12, target=DRONE_TARGET, dt=0.01, params=QUAD_PARAMS,
               weights=DRONE_WEIGHTS) -> np.ndarray:
    noise, npt = _f(noise_tkn)
    T, K, nu = noise.shape
    assert nu == 4
    S = np.zeros(K, f32)
    lib().oracle_quad_costs(K, T, npt, _f(u_nom)[1], _f(state12)[1], _f(target)[1], C.c_float(dt),
                            _f(params)[1], _f(weights)[1], S.ctypes.data_as(_fp))
    return S


def wb_costs(noise_tkn, u_nom, state12, q0, qd0, target_pos=ARM_TARGET_POS, target_quat=ARM_TARGET_QUAT,
             drone_target=DRONE_TARGET, dt=0.01, params=WB_PARAMS,
             weights=ARM_WEIGHTS + DRONE_WEIGHTS, chain: Chain = KINOVA_CHAIN) -> np.ndarray:
    noise, npt = _f(noise_tkn)
    T, K, nu = noise.shape
    assert nu == 11
    S = np.zeros(K, f32)
    lib().oracle_wb_costs(K, T, npt, _f(u_nom)[1], _f(state12)[1], _f(q0)[1], _f(qd0)[1],
                          chain.n, _i(chain.jtype)[1], _i(chain.qidx)[1], _f(chain.xyz)[1],
                          _f(chain.rpy)[1], _f(chain.axis)[1], _f(target_pos)[1], _f(target_quat)[1],
                          _f(drone_target)[1], C.c_float(dt), _f(params)[1], _f(weights)[1],
                          S.ctypes.data_as(_fp))
    return S


# --------------------------------------------------------------------------- weighting
def weights(S, lam=0.1):
    S, sp = _f(S)
    w = np.zeros_like(S)
    rho, eta = C.c_float(), C.c_float()
    lib().oracle_weights(int(S.size), sp, C.c_float(lam), w.ctypes.data_as(_fp), C.byref(rho), C.byref(eta))
    return w, rho.value, eta.value


def weighted_noise(noise_tkn, w) -> np.ndarray:
    noise, npt = _f(noise_tkn)
    T, K, nu = noise.shape
    out = np.zeros((T, nu), f32)
    lib().oracle_weighted_noise(K, T, nu, npt, _f(w)[1], out.ctypes.data_as(_fp))
    return out


def savgol_taps(window, polyorder=2) -> np.ndarray:
    taps = np.zeros(window, f32)
    rc = lib().oracle_savgol_taps(window, polyorder, taps.ctypes.data_as(_fp))
    if rc:
        raise ValueError(f"savgol_taps rc={rc}")
    return taps


def savgol(seq, window, polyorder=2) -> np.ndarray:
    seq, sp = _f(seq)
    T, nu = seq.shape
    out = np.zeros_like(seq)
    rc = lib().oracle_savgol(T, nu, sp, window, polyorder, out.ctypes.data_as(_fp))
    if rc:
        raise ValueError(f"savgol rc={rc}")
    return out


# --------------------------------------------------------------------------- noise
def philox4x32_10(ctr, key) -> np.ndarray:
    ctr = np.ascontiguousarray(ctr, np.uint32)
    key = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    u32p = C.POINTER(C.c_uint32)
    lib().oracle_philox4x32_10(ctr.ctypes.data_as(u32p), key.ctypes.data_as(u32p), out.ctypes.data_as(u32p))
    return out


def philox4x32(ctr, key, rounds=10) -> np.ndarray:
    ctr = np.ascontiguousarray(ctr, np.uint32)
    key = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    u32p = C.POINTER(C.c_uint32)
    lib().oracle_philox4x32_r(ctr.ctypes.data_as(u32p), key.ctypes.data_as(u32p), int(rounds), out.ctypes.data_as(u32p))
    return out


def philox_noise(K, T, nu, sigma, seed=0, step=0, k_offset=0, rounds=10) -> np.ndarray:
    noise = np.zeros((T, K, nu), f32)
    sig = np.broadcast_to(np.asarray(sigma, f32), (nu,)).copy()
    lib().oracle_philox_noise_r(K, T, nu, C.c_longlong(k_offset), C.c_uint64(seed), C.c_uint64(step), int(rounds),
                                _f(sig)[1], noise.ctypes.data_as(_fp))
    return noise


# --------------------------------------------------------------------------- whole steps
def _update(S, noise_tkn, u_prev, lam, window):
    """mppi.py:143-153 / drone_mppi.py:156-166: weights, weighted noise, Sav-Gol, u += w_eps."""
    w, rho, eta = weights(S, lam)
    w_eps_raw = weighted_noise(noise_tkn, w)
    w_eps = savgol(w_eps_raw, window, 2)
    u_new = (np.asarray(u_prev, f32) + w_eps).astype(f32)
    return dict(S=S, w=w, rho=rho, eta=eta, w_eps_raw=w_eps_raw, w_eps=w_eps, u_new=u_new)


def arm_step(noise_tkn, u_prev, q, qdot, base, lam=0.1, dt=0.01, state_f64=False, **kw):
    """mppi.py:122-162.  qdes keeps the `_qddot * dt` quirk (F11)."""
    S = arm_costs(noise_tkn, u_prev, q, qdot, base, dt=dt, state_f64=state_f64, **kw)
    out = _update(S, noise_tkn, u_prev, lam, 9)
    st = np.float64 if state_f64 else f32
    qddot_prev = np.asarray(u_prev, f32)[0]
    u0 = out["u_new"][0]
    dtf = st(dt) if state_f64 else f32(dt)
    qv, qd = np.asarray(q, st), np.asarray(qdot, st)
    out["vdes"] = (qd + u0 * dtf).astype(st)
    out["qdes"] = (qv + qddot_prev * dtf + f32(0.5) * u0 * f32(dt) * f32(dt)).astype(st)
    return out


def drone_step(noise_tkn, u_prev, x0, v0, lam=0.1, dt=0.01, **kw):
    """drone_mppi.py:140-176."""
    S = drone_costs(noise_tkn, u_prev, x0, v0, dt=dt, **kw)
    out = _update(S, noise_tkn, u_prev, lam, 5)
    u0 = out["u_new"][0]
    x0, v0 = np.asarray(x0, f32), np.asarray(v0, f32)
    dt2 = f32(dt * dt)
    out["v"] = (v0 + f32(dt) * u0).astype(f32)
    out["x"] = (x0 + v0 * f32(dt) + f32(0.5) * u0 * dt2).astype(f32)
    return out


def quad_step(noise_tkn, u_prev, state12, lam=0.1, dt=0.01, window=5, **kw):
    S = quad_costs(noise_tkn, u_prev, state12, dt=dt, **kw)
    return _update(S, noise_tkn, u_prev, lam, window)


def wb_step(noise_tkn, u_prev, state12, q, qdot, lam=0.1, dt=0.01, window=9, **kw):
    S = wb_costs(noise_tkn, u_prev, state12, q, qdot, dt=dt, **kw)
    out = _update(S, noise_tkn, u_prev, lam, window)
    qddot_prev = np.asarray(u_prev, f32)[0, 4:]
    u0 = out["u_new"][0, 4:]
    qv, qd = np.asarray(q, f32), np.asarray(qdot, f32)
    out["vdes"] = (qd + u0 * f32(dt)).astype(f32)
    out["qdes"] = (qv + qddot_prev * f32(dt) + f32(0.5) * u0 * f32(dt) * f32(dt)).astype(f32)
    return out
