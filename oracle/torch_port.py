"""Eager-PyTorch (CPU) restatement of the reference MPPI step: the SECOND oracle.

TEST INFRASTRUCTURE ONLY (same rules as mppi_oracle.c): used by tests/ and by the `cpu_baseline_torch`
leg of bench.py, never by the product package.

Where `mppi_oracle.c` restates the arithmetic sample by sample, this module replays the reference's
*tensor-op sequence* -- cumsum double integrator, batched 4x4 matmul chain, `torch.linalg.inv` of the
rotation block, Euler extraction, `conv1d` Savitzky-Golay -- so that (a) it reproduces the reference's own
float32 rounding almost bit for bit (pinned against tests/golden), and (b) timing it on the GPU box's host
cores gives "the reference's CPU PyTorch path" (BASELINE.json north_star) for a machine the reference
itself cannot travel to.  The unpinned quad / whole-body models follow DESIGN.md section 6 with batched ops
over K inside a Python loop over the horizon.

Reference paths are relative to src/mav_mppi/scripts/.  Noise is [K][T][nu] here (the reference's layout,
sampling/standard_normal_noise.py:24).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

PI, H = math.pi, math.pi / 2
KINOVA = dict(   # aerial_manipulator_gpu.urdf :67-74 (fixed joint_base) and joint_1..7
    fixed=[True] + [False] * 7,
    xyz=[[0, 0, 0], [0, 0, 0.15675], [0, 0.0016, -0.11875], [0, -0.205, 0], [0, 0, -0.205], [0, 0.2073, -0.0114],
         [0, 0, -0.10375], [0, 0.10375, 0]],
    rpy=[[PI, 0, 0], [0, PI, 0], [-H, 0, PI], [-H, 0, 0], [H, 0, PI], [-H, 0, PI], [H, 0, PI], [-H, 0, PI]],
)


# ----------------------------------------------------------------------------- rollout
def double_integrator(acc, q0, qd0, dt):
    """sampling/standard_normal_noise.py:32-50 (twin: mppi_solver/drone_mppi.py:46-55)."""
    K = acc.shape[0]
    v0 = qd0.reshape(1, 1, -1).expand(K, 1, -1)
    vel = torch.cumsum(acc * dt, dim=1) + v0
    vel_before = torch.cat([v0, vel[:, :-1]], dim=1)
    step = vel_before * dt + 0.5 * acc * dt ** 2
    return torch.cumsum(step, dim=1) + q0.reshape(1, 1, -1)


# ----------------------------------------------------------------------------- kinematics
def rot_rpy(r, p, y):
    """robot/transformation_matrix.py:4-25 on 0-dim tensors."""
    cr, sr, cp, sp, cy, sy = torch.cos(r), torch.sin(r), torch.cos(p), torch.sin(p), torch.cos(y), torch.sin(y)
    return torch.stack([torch.stack([cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr]),
                        torch.stack([sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr]),
                        torch.stack([-sp, cp * sr, cp * cr])])


def origin_tf(xyz, rpy):
    """robot/transformation_matrix.py:28-35."""
    T = torch.eye(4)
    a = torch.tensor(rpy)
    T[:3, :3] = rot_rpy(a[0], a[1], a[2])
    T[:3, 3] = torch.tensor(xyz)
    return T


def joint_rot_z(q):
    """robot/transformation_matrix.py:58-95 for the unit z axis: Rodrigues -> 4x4, dtype follows the 4x4 (float32)."""
    c, s, z, o = torch.cos(q), torch.sin(q), torch.zeros_like(q), torch.ones_like(q)
    R = torch.stack([torch.stack([c, -s, z], -1), torch.stack([s, c, z], -1), torch.stack([z, z, o], -1)], -2)
    T = torch.eye(4).expand(*q.shape, 4, 4).clone()
    T[..., :3, :3] = R
    return T


def chain_fk(q, start=None):
    """robot/urdfparser.py:122-163: left-to-right product along the chain; `start` = leading base transform."""
    T = torch.eye(4).expand(1, 1, 4, 4).clone() if start is None else start
    j = 0
    for fixed, xyz, rpy in zip(KINOVA["fixed"], KINOVA["xyz"], KINOVA["rpy"]):
        O = origin_tf(xyz, rpy)
        if fixed:
            T = T @ O
        else:
            T = T @ (O @ joint_rot_z(q[..., j]))
            j += 1
    return T


def base_tf_from_quat(b):
    """robot/urdf_fk.py:30-55 (xyzw, not normalised, float32)."""
    T = torch.eye(4, dtype=torch.float32)
    T[:3, 3] = b[:3]
    x, y, z, w = (b[3 + i] for i in range(4))
    T[:3, :3] = torch.tensor([[1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w],
                              [2 * x * y + 2 * z * w, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * x * w],
                              [2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x * x - 2 * y * y]], dtype=torch.float32)
    return T


def base_tf_from_rpy(x6):
    """robot/transformation_matrix.py:148-187 on [K, T, 6]."""
    r, p, y = x6[..., 3], x6[..., 4], x6[..., 5]
    cr, sr, cp, sp, cy, sy = torch.cos(r), torch.sin(r), torch.cos(p), torch.sin(p), torch.cos(y), torch.sin(y)
    R = torch.stack([torch.stack([cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr], -1),
                     torch.stack([sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr], -1),
                     torch.stack([-sp, cp * sr, cp * cr], -1)], -2)
    T = torch.eye(4).expand(*r.shape, 4, 4).clone()
    T[..., :3, :3] = R
    T[..., :3, 3] = x6[..., :3]
    return T


def quat_to_R(q):
    """utils/rotation_conversions.py:45-75 (xyzw)."""
    i, j, k, r = torch.unbind(q, -1)
    s = 2.0 / (q * q).sum(-1)
    return torch.stack([1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r),
                        s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
                        s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j)], -1).reshape(3, 3)


def euler_zyx(M):
    """utils/rotation_conversions.py:277-319 for "ZYX"."""
    return torch.stack([torch.atan2(M[..., 1, 0], M[..., 0, 0]), torch.asin((-M[..., 2, 0]).clamp(-1.0, 1.0)),
                        torch.atan2(M[..., 2, 1], M[..., 2, 2])], -1)


# ----------------------------------------------------------------------------- costs, weights, filter
def pose_cost(traj, tgt_pos, tgt_quat, w=(50.0, 30.0, 40.0, 30.0)):
    """cost/pose_cost.py:24-63 + cost/cost_manager.py:78-89."""
    Rt = quat_to_R(tgt_quat)

    def terms(Tm):
        dp = Tm[..., :3, 3] - tgt_pos
        ang = euler_zyx(torch.matmul(torch.linalg.inv(Tm[..., :3, :3]), Rt))
        return torch.norm(dp, p=2, dim=-1), torch.norm(ang, p=2, dim=-1)

    ps, os_ = terms(traj[:, :-1])
    pt, ot = terms(traj[:, -1])
    S = torch.zeros(traj.shape[0])
    S += torch.sum(w[0] * ps + w[1] * os_, dim=1)
    S += w[2] * pt + w[3] * ot
    return S


def softmin_weights(S, lam):
    """mppi_solver/mppi.py:173-193."""
    e = torch.exp((-1.0 / lam) * (S - S.min()))
    return e / e.sum()


def savgol(seq, window, order=2):
    """filter/svg_filter.py:13-90."""
    h = window // 2
    x = torch.arange(-h, h + 1, dtype=torch.float32)
    A = torch.stack([x ** i for i in range(order + 1)], dim=1)
    taps = (torch.linalg.inv(A.T @ A) @ A.T)[0]
    cols = []
    for i in range(seq.shape[1]):
        d = seq[:, i]
        padded = torch.cat([d[:h].flip(0), d, d[-h:].flip(0)])
        cols.append(F.conv1d(padded.view(1, 1, -1), taps.flip(0).view(1, 1, -1)).view(-1))
    return torch.stack(cols, dim=1)


def _update(S, noise, u_prev, lam, window):
    w = softmin_weights(S, lam)
    raw = torch.sum(w.view(-1, 1, 1) * noise, dim=0)
    return dict(S=S, w=w, w_eps_raw=raw, u_new=u_prev + savgol(raw, window))


# ----------------------------------------------------------------------------- whole steps
def arm_step(noise, u_prev, q, qdot, base, tgt_pos=(0.1029, 0.4055, 1.6498), tgt_quat=(-0.5, -0.5, 0.5, -0.5),
             lam=0.1, dt=0.01):
    """mppi_solver/mppi.py:122-162 (q, qdot float32 tensors, or float64 as update_joint makes them)."""
    v = u_prev.unsqueeze(0) + noise
    Q = double_integrator(v, q, qdot, dt)
    K, T, _ = Q.shape
    traj = base_tf_from_quat(base).expand(K, T, 4, 4) @ chain_fk(Q)
    out = _update(pose_cost(traj, torch.tensor(tgt_pos), torch.tensor(tgt_quat)), noise, u_prev, lam, 9)
    u0 = out["u_new"][0]
    out["vdes"] = qdot + u0 * dt
    out["qdes"] = q + u_prev[0] * dt + 0.5 * u0 * dt * dt
    return out


def drone_step(noise, u_prev, x0, v0, target=(1.0, 2.0, 3.4), lam=0.1, dt=0.01):
    """mppi_solver/drone_mppi.py:140-176."""
    tgt = torch.tensor(target, dtype=torch.float32)
    X = double_integrator(noise + u_prev, x0, v0, dt)
    S = torch.zeros(X.shape[0])
    S += torch.pow(X[:, :-1] - tgt, 2).sum(-1).sum(-1) * 100
    S += torch.pow(X[:, -1] - tgt, 2).sum(-1) * 20
    out = _update(S, noise, u_prev, lam, 5)
    u0 = out["u_new"][0]
    out["v"] = v0 + dt * u0
    out["x"] = x0 + v0 * dt + 0.5 * u0 * dt ** 2
    return out


def _quad_rollout(u, state, dt, params):
    """DESIGN.md section 6 (QUAD4), batched over K, Python loop over the horizon.  Returns p [K,T,3], rpy [K,T,3]."""
    m, ix, iy, iz, kd, gz = params
    K, T, _ = u.shape
    p, rpy, v, w = (state[3 * i:3 * i + 3].expand(K, 3).clone() for i in range(4))
    Iinv = torch.tensor([ix, iy, iz], dtype=torch.float32)
    ps, rs = [], []
    for t in range(T):
        ph, th, ps_ = rpy[:, 0], rpy[:, 1], rpy[:, 2]
        sph, cph, sth, cth, sps, cps = torch.sin(ph), torch.cos(ph), torch.sin(th), torch.cos(th), torch.sin(ps_), torch.cos(ps_)
        tth = sth / cth
        r3 = torch.stack([cps * sth * cph + sps * sph, sps * sth * cph - cps * sph, cth * cph], -1)
        w = w + dt * (Iinv * u[:, t, 1:4])
        rate = torch.stack([w[:, 0] + sph * tth * w[:, 1] + cph * tth * w[:, 2], cph * w[:, 1] - sph * w[:, 2],
                            (sph / cth) * w[:, 1] + (cph / cth) * w[:, 2]], -1)
        acc = (r3 * u[:, t, 0:1] - kd * v) / m
        acc[:, 2] = acc[:, 2] + gz
        nxt = rpy + dt * rate
        rpy = torch.atan2(torch.sin(nxt), torch.cos(nxt))
        v = v + dt * acc
        p = p + dt * v
        ps.append(p)
        rs.append(rpy)
    return torch.stack(ps, 1), torch.stack(rs, 1)


def quad_step(noise, u_prev, state, target=(1.0, 2.0, 3.4), lam=0.1, dt=0.01,
              params=(14.7, 1 / 1.57, 1 / 3.93, 1 / 2.59, 0.0, -9.81)):
    tgt = torch.tensor(target, dtype=torch.float32)
    P, _ = _quad_rollout(noise + u_prev, state, dt, params)
    d = torch.pow(P - tgt, 2).sum(-1)
    return _update(d[:, :-1].sum(-1) * 100 + d[:, -1] * 20, noise, u_prev, lam, 5)


def wb_step(noise, u_prev, state, q, qdot, tgt_pos=(0.1029, 0.4055, 1.6498), tgt_quat=(-0.5, -0.5, 0.5, -0.5),
            drone_target=(1.0, 2.0, 3.4), lam=0.1, dt=0.01, params=(14.7 + 5.5, 1 / 1.57, 1 / 3.93, 1 / 2.59, 0.0, -9.81)):
    u = noise + u_prev
    P, R = _quad_rollout(u[..., :4], state, dt, params)
    Q = double_integrator(u[..., 4:], q, qdot, dt)
    traj = chain_fk(Q, start=base_tf_from_rpy(torch.cat([P, R], -1)))
    S = pose_cost(traj, torch.tensor(tgt_pos), torch.tensor(tgt_quat))
    d = torch.pow(P - torch.tensor(drone_target, dtype=torch.float32), 2).sum(-1)
    S = S + (d[:, :-1].sum(-1) * 100 + d[:, -1] * 20)
    return _update(S, noise, u_prev, lam, 9)
