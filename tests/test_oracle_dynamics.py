"""Pins the float64 torque-law oracle (oracle/arm_dynamics.py) without Pinocchio: every term of the recursive
Newton-Euler / composite-rigid-body restatement is compared with an independent Lagrangian derivation
(geometric Jacobians, finite differences of the kinetic and potential energy).  SURVEY 8(f) item 4."""
import numpy as np
import pytest

from oracle import arm_dynamics as dyn

Q = np.array([1.2, 2.0, -0.4, 4.0, 0.7, 4.2, -1.0])
QD = np.array([0.3, -0.5, 0.2, 0.8, -0.6, 0.4, -0.9])


def _grad(f, q, h=1e-6):
    g = np.zeros(7)
    for k in range(7):
        e = np.zeros(7)
        e[k] = h
        g[k] = (f(q + e) - f(q - e)) / (2 * h)
    return g


def test_merged_link7_mass_and_generated_tables_are_current():
    from oracle import arm_inertia_gen as gen
    assert gen.MASS[6] == pytest.approx(0.99 + 1e-4 + 6 * 0.01)            # full_robot_floating2.urdf:364,408,445..630
    assert sum(gen.MASS) == pytest.approx(5.0895, abs=1e-4)


def test_mass_matrix_matches_the_lagrangian_form():
    M = dyn.mass_matrix_arm(Q)
    assert np.allclose(M, M.T, atol=1e-12)
    assert np.linalg.eigvalsh(M).min() > 0
    assert np.allclose(M, dyn.lagrangian_mass_matrix(Q), atol=1e-12)


@pytest.mark.parametrize("base_quat", [(0, 0, 0, 1.0), (0.0499792, -0.0998334, 0.1494381, 0.9824485)])
def test_gravity_term_is_the_gradient_of_the_potential(base_quat):
    R = dyn.quat_matrix(base_quat)
    g = dyn.rnea_arm(Q, np.zeros(7), np.zeros(7), base_R=R)
    assert np.allclose(g, _grad(lambda q: dyn.potential_energy(q, R), Q), atol=1e-7)


def test_coriolis_term_matches_christoffel_symbols():
    c = dyn.rnea_arm(Q, QD, np.zeros(7), gravity=False)
    h = 1e-6
    dM = np.zeros((7, 7, 7))                         # dM[i] = dM / dq_i
    for i in range(7):
        e = np.zeros(7)
        e[i] = h
        dM[i] = (dyn.lagrangian_mass_matrix(Q + e) - dyn.lagrangian_mass_matrix(Q - e)) / (2 * h)
    want = np.einsum("ikj,i,j->k", dM, QD, QD) - 0.5 * np.einsum("kij,i,j->k", dM, QD, QD)
    assert np.allclose(c, want, atol=1e-7)


def test_rotating_base_gives_the_centrifugal_generalised_force():
    w = np.array([0.4, -0.7, 1.1])
    tau = dyn.rnea_arm(Q, np.zeros(7), np.zeros(7), base_twist=np.concatenate([np.zeros(3), w]), gravity=False)
    want = -_grad(lambda q: 0.5 * w @ dyn.arm_inertia_about_base(q) @ w, Q)
    assert np.allclose(tau, want, atol=1e-7)


def test_base_linear_velocity_only_enters_through_omega_cross_v():
    v0, w = np.array([1.5, -0.8, 0.6]), np.array([0.4, -0.7, 1.1])
    none = dyn.rnea_arm(Q, QD, np.zeros(7))
    assert np.allclose(dyn.rnea_arm(Q, QD, np.zeros(7), base_twist=np.concatenate([v0, np.zeros(3)])), none, atol=1e-12)   # Galilean
    both = dyn.rnea_arm(Q, QD, np.zeros(7), base_twist=np.concatenate([v0, w]))
    rot = dyn.rnea_arm(Q, QD, np.zeros(7), base_twist=np.concatenate([np.zeros(3), w]))
    a = np.cross(w, v0)                               # classical acceleration of the base origin, base frame

    def field_potential(q):
        U = 0.0
        for i, (Ri, pi, _) in enumerate(dyn.link_frames(q)):
            m, cm, _ = dyn._inertia(i)
            U += m * a @ (pi + Ri @ cm)
        return U
    assert np.allclose(both - rot, _grad(field_potential, Q), atol=1e-7)


def test_torque_law_composition():
    q_full = np.concatenate([[0.3, -0.2, 1.7, 0.0499792, -0.0998334, 0.1494381, 0.9824485], Q])
    v_full = np.concatenate([[0.2, -0.1, 0.05, 0.1, 0.2, -0.3], QD])
    qdes = Q + 0.01
    tau = dyn.torque_law(q_full, v_full, qdes)
    M = dyn.mass_matrix_arm(Q)
    assert np.allclose(tau, M @ (400 * (qdes - Q) - 40 * QD) + dyn.nle_arm(q_full, v_full))
    assert np.isfinite(tau).all()


def test_general_joint_axes():
    """Chains whose joint axes are not +z (the folded-frame path of the library): same Lagrangian pins."""
    from oracle import oracle as orc
    base = orc.KINOVA_CHAIN
    axis, rpy, xyz = base.axis.copy(), base.rpy.copy(), base.xyz.copy()
    axis[2] = [0, 1, 0]; axis[4] = [1, 0, 0]; axis[6] = [0.6, 0.0, 0.8]
    rpy[3] = [0.3, -0.2, 0.5]; xyz[5] = [0.01, 0.2, -0.03]
    ch = orc.Chain(base.jtype, base.qidx, xyz, rpy, axis)
    M = dyn.mass_matrix_arm(Q, ch)
    assert np.allclose(M, dyn.lagrangian_mass_matrix(Q, ch), atol=1e-12)
    g = dyn.rnea_arm(Q, np.zeros(7), np.zeros(7), chain=ch)
    assert np.allclose(g, _grad(lambda q: dyn.potential_energy(q, chain=ch), Q), atol=1e-7)
