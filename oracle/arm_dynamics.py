"""Float64 oracle of the arm node's computed-torque law (TEST INFRASTRUCTURE -- never imported by the product).

Reference: `src/mav_mppi/scripts/kinova.py:126-131,184`

    pin.computeAllTerms(model, data, q, v);  g = data.nle
    torque = data.M[6:, 6:] @ (400 * (qdes - q[7:]) + 40 * (-v[6:])) + g[6:]

with the Pinocchio model of `full_robot_floating2.urdf` (free-flyer base + seven revolute joints, fixed
children merged).  Pinocchio is NOT in this image, so this restates the published algorithms it runs --
recursive Newton-Euler for `nle`, composite-rigid-body for `M` (Featherstone, *Rigid Body Dynamics Algorithms*,
tables 5.1 and 6.2) -- in spatial-vector form, and `tests/test_oracle_dynamics.py` pins it against an
independent Lagrangian derivation (finite differences of the kinetic and potential energy).
PARITY UNPINNED against Pinocchio itself.

Conventions (Pinocchio's): q = [base xyz, base quat xyzw, q1..q7]; v = [base linear velocity, base angular
velocity, both in the BASE frame, qdot1..7]; `nle = rnea(q, v, 0)`; gravity (0, 0, -9.81) in the world frame.
Only the arm rows (6:) are produced: that is all the torque law reads.
"""
from __future__ import annotations

import numpy as np

from . import arm_inertia_gen as gen
from .oracle import KINOVA_CHAIN

GRAVITY = 9.81
KP, KD = 400.0, 40.0                      # kinova.py:184


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def quat_matrix(q):
    x, y, z, w = np.asarray(q, float) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def placements(chain=KINOVA_CHAIN):
    """Joint placements (R_i, p_i, axis_i), i = 1..7: pose of joint i's frame at q_i = 0 in the frame of its parent
    body (the base body for i = 1; fixed joints are folded in) and the unit joint axis in that frame."""
    out, R, p = [], np.eye(3), np.zeros(3)
    for j in range(chain.n):
        Ro, to = _rpy(*chain.rpy[j].astype(float)), chain.xyz[j].astype(float)
        p, R = p + R @ to, R @ Ro
        if chain.jtype[j] == 1:
            ax = chain.axis[j].astype(float)
            out.append((R, p, ax / np.linalg.norm(ax)))
            R, p = np.eye(3), np.zeros(3)
        else:
            assert chain.jtype[j] == 0, "the dynamics oracle handles fixed and revolute joints"
    return out


def _axis_rotation(ax, q):
    """Rodrigues rotation about the unit axis (transformation_matrix.py:68-84)."""
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0.0]])
    return np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * (K @ K)


def _inertia(i):
    ixx, ixy, ixz, iyy, iyz, izz = gen.INERTIA[i]
    return gen.MASS[i], np.array(gen.COM[i]), np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])


def _spatial_inertia_mul(m, c, Ic, w, v):
    """(n, f) = I (w, v) for the spatial inertia of a body with mass m, centre of mass c, inertia Ic about c."""
    h = m * c
    Io_w = Ic @ w + m * (c @ c * w - c * (c @ w))          # inertia about the frame origin times w
    return Io_w + np.cross(h, v), m * v - np.cross(h, w)


def rnea_arm(q_arm, qd_arm, qdd_arm, base_R=np.eye(3), base_twist=np.zeros(6), gravity=True, chain=KINOVA_CHAIN):
    """Arm rows of rnea(q, v, a) with zero base acceleration (Featherstone table 5.1, free-flyer root).

    base_twist = (linear, angular) velocity of the base in the base frame (Pinocchio's v[:6])."""
    pl = placements(chain)
    w_p, v_p = np.asarray(base_twist[3:6], float), np.asarray(base_twist[0:3], float)
    al_p = np.zeros(3)
    a_p = base_R.T @ np.array([0.0, 0.0, GRAVITY]) if gravity else np.zeros(3)      # a_0 = -a_gravity
    Rs, n_l, f_l = [], [], []
    for i in range(7):
        R0, p, z = pl[i]
        R = R0 @ _axis_rotation(z, q_arm[i])                          # child axes in the parent frame
        Rs.append((R, p, z))
        w = R.T @ w_p + z * qd_arm[i]
        v = R.T @ (v_p + np.cross(w_p, p))
        al = R.T @ al_p + z * qdd_arm[i] + np.cross(w, z * qd_arm[i])
        a = R.T @ (a_p + np.cross(al_p, p)) + np.cross(v, z * qd_arm[i])
        m, cm, Ic = _inertia(i)
        hn, hf = _spatial_inertia_mul(m, cm, Ic, w, v)
        an, af = _spatial_inertia_mul(m, cm, Ic, al, a)
        n_l.append(an + np.cross(w, hn) + np.cross(v, hf))
        f_l.append(af + np.cross(w, hf))
        w_p, v_p, al_p, a_p = w, v, al, a
    tau = np.zeros(7)
    for i in range(6, -1, -1):
        R, p, z = Rs[i]
        tau[i] = n_l[i] @ z
        if i > 0:
            fp = R @ f_l[i]
            n_l[i - 1] = n_l[i - 1] + R @ n_l[i] + np.cross(p, fp)
            f_l[i - 1] = f_l[i - 1] + fp
    return tau


def mass_matrix_arm(q_arm, chain=KINOVA_CHAIN):
    """M[6:, 6:]: independent of the base being free (composite inertias of the arm subtrees only)."""
    M = np.zeros((7, 7))
    for j in range(7):
        e = np.zeros(7)
        e[j] = 1.0
        M[:, j] = rnea_arm(q_arm, np.zeros(7), e, gravity=False, chain=chain)
    return M


def nle_arm(q_full, v_full, chain=KINOVA_CHAIN):
    q_full, v_full = np.asarray(q_full, float), np.asarray(v_full, float)
    return rnea_arm(q_full[7:14], v_full[6:13], np.zeros(7), base_R=quat_matrix(q_full[3:7]), base_twist=v_full[:6], chain=chain)


def torque_law(q_full, v_full, qdes, kp=KP, kd=KD, chain=KINOVA_CHAIN):
    """kinova.py:184."""
    q_full, v_full = np.asarray(q_full, float), np.asarray(v_full, float)
    ades = kp * (np.asarray(qdes, float) - q_full[7:14]) + kd * (-v_full[6:13])
    return mass_matrix_arm(q_full[7:14], chain) @ ades + nle_arm(q_full, v_full, chain)


# --------------------------------------------------------------------------- independent Lagrangian check
def link_frames(q_arm, base_R=np.eye(3), base_p=np.zeros(3), chain=KINOVA_CHAIN):
    """World pose (R, p) of every arm link frame and the joint axis in world axes."""
    out, R, p = [], base_R, base_p
    for i, (R0, p0, ax) in enumerate(placements(chain)):
        p = p + R @ p0
        R = R @ R0 @ _axis_rotation(ax, q_arm[i])
        out.append((R, p, R @ ax))
    return out


def lagrangian_mass_matrix(q_arm, chain=KINOVA_CHAIN):
    """M = sum_i m_i Jv_i^T Jv_i + Jw_i^T (R_i Ic_i R_i^T) Jw_i with geometric Jacobians (fixed base)."""
    fr = link_frames(q_arm, chain=chain)
    M = np.zeros((7, 7))
    for i in range(7):
        m, cm, Ic = _inertia(i)
        Ri, pi, _ = fr[i]
        com = pi + Ri @ cm
        Jv, Jw = np.zeros((3, 7)), np.zeros((3, 7))
        for j in range(i + 1):
            zj, pj = fr[j][2], fr[j][1]
            Jw[:, j] = zj
            Jv[:, j] = np.cross(zj, com - pj)
        M += m * Jv.T @ Jv + Jw.T @ (Ri @ Ic @ Ri.T) @ Jw
    return M


def potential_energy(q_arm, base_R=np.eye(3), chain=KINOVA_CHAIN):
    U = 0.0
    for i, (Ri, pi, _) in enumerate(link_frames(q_arm, base_R, chain=chain)):
        m, cm, _ = _inertia(i)
        U += m * GRAVITY * (pi + Ri @ cm)[2]
    return U


def arm_inertia_about_base(q_arm):
    """Rotational inertia tensor of the whole arm about the base origin, in base axes."""
    I = np.zeros((3, 3))
    for i, (Ri, pi, _) in enumerate(link_frames(q_arm)):
        m, cm, Ic = _inertia(i)
        r = pi + Ri @ cm
        I += Ri @ Ic @ Ri.T + m * (r @ r * np.eye(3) - np.outer(r, r))
    return I
