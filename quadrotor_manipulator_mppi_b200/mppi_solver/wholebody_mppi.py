"""Whole-body quadrotor + Kinova arm MPPI (nu=11: thrust, three torques, seven joint accelerations).

PARITY UNPINNED: the reference has no whole-body controller (README.md:33 to-do; hooks only at
`robot/urdfparser.py:36,128-131` and `mppi.py:39`).  Model and cost are specified in DESIGN.md ("wb11"):
quad rigid body moves the arm base, arm double integrator + FK as in the pinned arm path, cost =
arm pose cost on the world end-effector + drone position cost on the base.
State = p[3], rpy[3], v[3], w[3], q[7], qdot[7].
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native
from ..core import NativeSolver
from ..utils.pose import Pose, TensorWatch


class MPPI:
    MODEL = _native.MODEL_WB11

    def __init__(self, *, n_samples: int = 1000, n_horizon: int = 32, dt: float = 0.01, sigma=None,
                 lam: float = 0.1, seed: int = 0, device=None, mass: float = 14.7 + 5.5, k_offset: int = 0,
                 torque_law: bool = False, torque_gains=(400.0, 40.0), philox_rounds=None, fused=None):
        self.n_samples, self.n_horizon, self.dt, self.n_action = int(n_samples), int(n_horizon), float(dt), 11
        self._lambda = float(lam)
        self.mass = float(mass)
        if sigma is None:
            sigma = (30.0 * mass, 1.0, 1.0, 1.0) + (0.1,) * 7
        qp = (mass, 1.0 / 1.57, 1.0 / 3.93, 1.0 / 2.59, 0.0, -9.81)
        self._solver = NativeSolver(self.MODEL, n_samples=n_samples, n_horizon=n_horizon, dt=dt, lam=lam,
                                    sigma=sigma, seed=seed, device=device, quad_params=qp, k_offset=k_offset,
                                    cost_flags=_native.OPT_TORQUE_LAW if torque_law else 0, torque_gains=torque_gains,
                                    philox_rounds=philox_rounds, fused=fused)
        self.torque_law = bool(torque_law)
        self.torque = np.zeros(7)         # joint torques of the arm's computed-torque law (kinova.py:184) when enabled
        self.device = self._solver.device
        self.target_pose = Pose()
        self.target_pose.pose = torch.tensor([0.1029, 0.4055, 1.6498])
        self.target_pose.orientation = torch.tensor([-0.5, -0.5, 0.5, -0.5])
        self.drone_target = torch.tensor([1.0, 2.0, 3.4])
        self._target_watch = TensorWatch()
        self._state = np.zeros(26, np.float32)
        self._solver.set_state(self._state)
        hover = torch.zeros(self.n_horizon, 11)
        hover[:, 0] = mass * 9.81
        self.u_prev = hover

    @property
    def u_prev(self) -> torch.Tensor:
        """Nominal sequence (warm start, not shifted).  A fresh clone: the solver's ping-pong buffers are reused."""
        return self._solver.u_prev.clone()

    @u_prev.setter
    def u_prev(self, value):
        self._solver.u_prev = value

    def set_state(self, p, rpy, v, w, q, qdot):
        self._solver.set_state_parts(p, rpy, v, w, q, qdot)

    def _sync_target(self):
        dt_ = self.drone_target
        if self._target_watch.changed(self.target_pose.pose, self.target_pose.orientation, dt_):
            tgt = self.target_pose.as_floats() + tuple(float(v) for v in torch.as_tensor(dt_).reshape(-1))
            self._solver.set_target(pos=tgt[:3], quat=tgt[3:7], drone_target=tgt[7:])

    def compute_control_input(self, noise=None, noise_layout: str = "tkn"):
        """Returns (qdes[7], vdes[7], next_base_state[12]) as numpy arrays."""
        self._sync_target()
        out = self._solver.step(self._solver.prepare_noise(noise, noise_layout))
        if self.torque_law:
            self.torque = out[_native.MPPI_OUT_TORQUE:_native.MPPI_OUT_TORQUE + 7].astype(np.float64)
        return out[0:7].copy(), out[7:14].copy(), out[_native.MPPI_OUT_BASE:_native.MPPI_OUT_BASE + 12].copy()
