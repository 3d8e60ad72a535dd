"""Static FLOP counter over the restated maths of ONE rollout-step (one sample, one horizon step).

TEST / MEASUREMENT INFRASTRUCTURE: re-derives the per-step work "with a static counter over the
restatement" instead of quoting the survey's estimate.  The arithmetic of `oracle/mppi_oracle.c` (arm_sample_cost,
quad_advance, oracle_wb_costs, oracle_drone_costs, oracle_pose_terms, fk_chain) is replayed on a counting scalar type:

  * every add / sub / mul between two run-time values counts 1 FLOP (an FMA is therefore 2);
  * transcendental and special-function evaluations (sin, cos, atan2, asin, sqrt, reciprocal / divide) are counted
    SEPARATELY and are not part of the FLOP figure (SURVEY 8(d) convention);
  * negation, abs, copies, comparisons, min / max and clamps are free.

Two figures per model:
  dense       -- "algorithmic": the URDF constants are treated as GENERAL numbers, i.e. every joint costs a full
                 3x3 . 3x3 constant product + translation, the target rotation is a full matrix.  This is the
                 figure `roofline.achieved` uses (mppi_algorithmic_flops_per_rollout_step).
  structural  -- the same maths with multiplications by constants that are exactly 0 or +-1 and additions of exact
                 zeros skipped: what a kernel with the chain's constants unrolled has to execute
                 (mppi_executed_flops_per_rollout_step; the ncu-measured thread-level FLOP count is reported next to it).

    python -m oracle.flop_count            # prints the table that is frozen in csrc/mppi_b200.cu
"""
from __future__ import annotations

import math

from . import oracle as orc

COUNT = {"add": 0, "mul": 0, "sincos": 0, "atan2": 0, "asin": 0, "sqrt": 0, "rcp": 0}
SPARSE = False            # structural mode: constants 0 / +-1 are free


def reset():
    for k in COUNT:
        COUNT[k] = 0


class V:
    """A run-time value (something that differs per sample or per step)."""
    __slots__ = ("x",)

    def __init__(self, x):
        self.x = float(x)

    # ---- helpers
    @staticmethod
    def _is_const(o):
        return not isinstance(o, V)

    def __add__(self, o):
        if self._is_const(o):
            if SPARSE and o == 0.0:
                return self
            COUNT["add"] += 1
            return V(self.x + o)
        COUNT["add"] += 1
        return V(self.x + o.x)
    __radd__ = __add__

    def __sub__(self, o):
        if self._is_const(o):
            if SPARSE and o == 0.0:
                return self
            COUNT["add"] += 1
            return V(self.x - o)
        COUNT["add"] += 1
        return V(self.x - o.x)

    def __rsub__(self, o):
        if SPARSE and o == 0.0:
            return V(-self.x)
        COUNT["add"] += 1
        return V(o - self.x)

    def __mul__(self, o):
        if self._is_const(o):
            if SPARSE and o == 0.0:
                return 0.0                     # a constant zero: later additions of it are free too
            if SPARSE and abs(o) == 1.0:
                return V(self.x * o)
            COUNT["mul"] += 1
            return V(self.x * o)
        COUNT["mul"] += 1
        return V(self.x * o.x)
    __rmul__ = __mul__

    def __neg__(self):
        return V(-self.x)


def sincos(a: V):
    COUNT["sincos"] += 1
    return V(math.sin(a.x)), V(math.cos(a.x))


def sqrt(a: V):
    COUNT["sqrt"] += 1
    return V(math.sqrt(max(a.x, 0.0)))


def rcp(a: V):
    COUNT["rcp"] += 1
    return V(1.0 / a.x)


def atan2(y: V, x: V):
    COUNT["atan2"] += 1
    return V(math.atan2(y.x, x.x))


def asin(a: V):
    COUNT["asin"] += 1
    return V(math.asin(max(-1.0, min(1.0, a.x))))


def _snap(c):
    """URDF constants: right-angle origins leave ~4e-8 residues in float32; structurally they are 0/+-1."""
    r = round(c)
    return float(r) if abs(c - r) < 1e-6 else float(c)


def folded_chain():
    """The j2s7s300 chain folded to C0 Rz(q1) C1 ... Rz(q7) C7 (the same fold mppi_set_chain performs), as plain
    python constants (R 3x3 row-major, t 3)."""
    import numpy as np
    ch = orc.KINOVA_CHAIN
    C = np.eye(4)
    out = []
    for j in range(ch.n):
        C = C @ orc.lib_make_transform(ch.xyz[j], ch.rpy[j]).astype(np.float64)
        if ch.jtype[j] == 0:
            continue
        out.append(C.copy())
        C = np.eye(4)
    out.append(C.copy())
    return [([[ _snap(c) for c in row[:3]] for row in M[:3]], [_snap(M[r][3]) for r in range(3)]) for M in out]


def compose_const(R, p, Cr, Ct):
    """(R, p) <- (R Cr, p + R Ct) with constant Cr, Ct."""
    newp = [p[r] + (R[r][0] * Ct[0] + R[r][1] * Ct[1] + R[r][2] * Ct[2]) for r in range(3)]
    newR = [[R[r][0] * Cr[0][c] + R[r][1] * Cr[1][c] + R[r][2] * Cr[2][c] for c in range(3)] for r in range(3)]
    return newR, newp


def rotate_z(R, c, s):
    """R <- R Rz(q): columns 0 and 1 mix."""
    return [[R[r][0] * c + R[r][1] * s, R[r][1] * c - R[r][0] * s, R[r][2]] for r in range(3)]


def _fix(x):
    """Sums of constant zeros stay python floats in structural mode; lift to V where a V is required."""
    return x if isinstance(x, V) else V(x)


def pose_terms(R, p, tgt_p, tgt_R):
    """oracle_pose_terms: ||p - p*||, ||euler_ZYX(R^T R*)||; only the five needed entries of D = R^T R*."""
    d = [_fix(p[i]) - tgt_p[i] for i in range(3)]
    pos = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])

    def D(i, j):      # (R^T R*)_ij = sum_k R[k][i] R*[k][j]
        return _fix(_fix(R[0][i]) * tgt_R[0][j] + _fix(R[1][i]) * tgt_R[1][j] + _fix(R[2][i]) * tgt_R[2][j])
    e0 = atan2(D(1, 0), D(0, 0))
    e1 = asin(-D(2, 0))
    e2 = atan2(D(2, 1), D(2, 2))
    ori = sqrt(e0 * e0 + e1 * e1 + e2 * e2)
    return pos, ori


def integrate_joint(a, st, dt, dt2):
    """standard_normal_noise.py:32-50 for one input; st = dict(cum_v, cum_q, vprev), q0/qd0 run-time values."""
    dq = st["vprev"] * dt + (a * 0.5) * dt2
    st["cum_v"] = st["cum_v"] + a * dt
    st["vprev"] = st["cum_v"] + st["qd0"]
    st["cum_q"] = st["cum_q"] + dq
    return st["cum_q"] + st["q0"]


def control(u_nom, z, sigma):
    """v = u + sigma * z (mppi.py:130 on the in-kernel noise)."""
    return u_nom + z * sigma


def quad_advance(s, u, dt, inv_m, Iinv, kd, gz):
    """oracle quad_advance; sin/cos of the CURRENT attitude are carried in s (computed at the end of the previous step)."""
    sphi, cphi, sth, cth, spsi, cpsi = s["sc"]
    inv_cth = rcp(cth)
    tth = sth * inv_cth
    r02 = cpsi * sth * cphi + spsi * sphi
    r12 = spsi * sth * cphi - cpsi * sphi
    r22 = cth * cphi
    w = [s["w"][i] + (u[1 + i] * Iinv[i]) * dt for i in range(3)]
    dphi = w[0] + sphi * tth * w[1] + cphi * tth * w[2]
    dth = cphi * w[1] - sphi * w[2]
    dpsi = (sphi * inv_cth) * w[1] + (cphi * inv_cth) * w[2]
    F = u[0]
    ax = (r02 * F - s["v"][0] * kd) * inv_m
    ay = (r12 * F - s["v"][1] * kd) * inv_m
    az = (r22 * F - s["v"][2] * kd) * inv_m + gz
    s["w"] = w
    s["rpy"] = [s["rpy"][0] + dphi * dt, s["rpy"][1] + dth * dt, s["rpy"][2] + dpsi * dt]      # wrap: free (range reduction)
    s["v"] = [s["v"][0] + ax * dt, s["v"][1] + ay * dt, s["v"][2] + az * dt]
    s["p"] = [s["p"][i] + s["v"][i] * dt for i in range(3)]
    sc = []
    for i in range(3):
        si, ci = sincos(s["rpy"][i])
        sc += [si, ci]
    s["sc"] = sc


def rpy_matrix(sc):
    sr, cr, sp, cp, sy, cy = sc
    return [[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr]]


def drone_pos_cost(p, target):
    e = [p[i] - target[i] for i in range(3)]
    return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]


def _arm_state():
    return [dict(cum_v=V(0.01 * i), cum_q=V(0.02 * i), vprev=V(0.1), q0=V(orc.Q_HOME[i]), qd0=V(0.05)) for i in range(7)]


def count_step(model: str, sparse: bool) -> dict:
    """FLOPs and special-function counts of one rollout-step of `model` ("drone3", "quad4", "arm7", "wb11")."""
    global SPARSE
    SPARSE = sparse
    dt, dt2 = 0.01, 0.0001
    tgt_R = [[0.3, 0.9, 0.1], [0.2, 0.1, -0.9], [-0.9, 0.3, 0.2]] if not sparse else \
        [[_snap(c) for c in row] for row in orc.quaternion_to_matrix(orc.ARM_TARGET_QUAT).tolist()]
    chain = folded_chain()
    if not sparse:
        # general constants: nothing is 0 or +-1
        chain = [([[0.37 + 0.01 * (3 * r + c) for c in range(3)] for r in range(3)], [0.11, 0.23, 0.31]) for _ in chain]
    # ---- state set up outside the counted region
    quad = dict(p=[V(0.0), V(0.0), V(2.1)], rpy=[V(0.01), V(0.02), V(0.03)], v=[V(0.1)] * 3, w=[V(0.01)] * 3,
                sc=[V(0.01), V(1.0), V(0.02), V(1.0), V(0.03), V(1.0)])
    arm = _arm_state()
    drone = [dict(cum_v=V(0.0), cum_q=V(0.0), vprev=V(0.0), q0=V(0.0), qd0=V(0.0)) for _ in range(3)]
    baseR = [[V(1.0), V(0.0), V(0.0)], [V(0.0), V(1.0), V(0.0)], [V(0.0), V(0.0), V(1.0)]]      # arm: B C0, loop-invariant
    basep = [V(0.0), V(0.0), V(2.1)]
    S = V(0.0)
    reset()
    nu = {"drone3": 3, "quad4": 4, "arm7": 7, "wb11": 11}[model]
    u = [control(V(0.1), V(0.3), 0.1 + 0.01 * i) for i in range(nu)]
    if model == "drone3":
        x = [integrate_joint(u[i], drone[i], dt, dt2) for i in range(3)]
        S = S + drone_pos_cost(x, orc.DRONE_TARGET if sparse else (1.1, 2.1, 3.3))
    if model in ("quad4", "wb11"):
        quad_advance(quad, u[:4], dt, 1.0 / 14.7, (0.63, 0.25, 0.38), 0.05 if not sparse else 0.0, -9.81)
        S = S + drone_pos_cost(quad["p"], (1.1, 2.1, 3.3))
    if model in ("arm7", "wb11"):
        a0 = 4 if model == "wb11" else 0
        q = [integrate_joint(u[a0 + i], arm[i], dt, dt2) for i in range(7)]
        if model == "wb11":
            R, p = compose_const(rpy_matrix(quad["sc"]), quad["p"], *chain[0])      # T(p_t, rpy_t) C0, every step
        else:
            R, p = baseR, basep
        for j in range(7):
            sj, cj = sincos(q[j])
            R = rotate_z(R, cj, sj)
            R, p = compose_const(R, p, *chain[j + 1])
        pos, ori = pose_terms(R, p, (0.1029, 0.4055, 1.6498), tgt_R)
        S = S + (pos * 50.0 + ori * 30.0)
    out = dict(COUNT)
    out["flop"] = out["add"] + out["mul"]
    return out


def table() -> dict:
    return {m: {"dense": count_step(m, False), "structural": count_step(m, True)} for m in ("drone3", "quad4", "arm7", "wb11")}


if __name__ == "__main__":
    import json
    print(json.dumps(table(), indent=1))
