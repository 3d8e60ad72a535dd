"""Arm MPPI controller (Kinova j2s7s300, nu=7): drop-in for the reference class
`mppi_solver/mppi.py:27-200`.

Kept from the reference: the no-arg constructor and its defaults (K=100, T=32, dt=0.01,
sigma=0.1, lambda=0.1; mppi.py:37-42,75, sampling/standard_normal_noise.py:17),
`update_joint(q_full, v_full)` (mppi.py:196-200), `compute_control_input()` returning
`(qdes, vdes)` numpy arrays incl. the `_qddot * dt` term (mppi.py:157-162), the
un-shifted `u_prev` warm start (mppi.py:125,153), `target_pose`, `compute_weights`,
`check_reach`.  New keyword-only constructor arguments and the `noise=` / `return_costs=`
hooks replace the monkey-patching the reference needs for the same purpose.

Everything between "state in" and "controls out" runs in libmppi_b200.so on the GPU.
"""
from __future__ import annotations

import types

import numpy as np
import torch

from .. import _native
from ..core import NativeSolver
from ..utils.pose import Pose, TensorWatch


_U0_NEW = slice(_native.MPPI_OUT_U0_NEW, _native.MPPI_OUT_U0_NEW + 7)
_U0_OLD = slice(_native.MPPI_OUT_U0_OLD, _native.MPPI_OUT_U0_OLD + 7)
_STATS = slice(_native.MPPI_OUT_REACH, _native.MPPI_OUT_ESS + 1)       # reach, rho, eta, ess are consecutive
_HALF = np.float32(0.5)
_TORQUE = slice(_native.MPPI_OUT_TORQUE, _native.MPPI_OUT_TORQUE + 7)


class MPPI:
    MODEL = _native.MODEL_ARM7
    COST_TERMS = {"covar": _native.COST_COVAR, "centering": _native.COST_CENTERING, "joint_traj": _native.COST_JOINT_TRAJ,
                  "action": _native.COST_ACTION, "joint_limit": _native.COST_JOINT_LIMIT}

    def __init__(self, *, n_samples: int = 100, n_horizon: int = 32, dt: float = 0.01, sigma=0.1, lam: float = 0.1,
                 seed: int = 0, device=None, verbose: bool = True, cost_terms=(), torque_law: bool = False,
                 torque_gains=(400.0, 40.0), philox_rounds=None, fused=None, time_parallel=None):
        self.n_action = 7
        self.n_manipulator_dof = 7
        self.n_mobile_dof = 0
        self.n_samples = int(n_samples)
        self.n_horizon = int(n_horizon)
        self.dt = float(dt)
        self._lambda = float(lam)
        # cost_terms: any of the terms the reference constructs but comments out of the sum
        # (cost/cost_manager.py:83-87): "covar", "centering", "joint_traj", "action", "joint_limit"
        flags = 0
        for name in cost_terms:
            flags |= self.COST_TERMS[name]
        # torque_law: also evaluate the arm node's computed-torque law on the device (kinova.py:126-131,184:
        # M[6:,6:] (kp (qdes - q) - kd qdot) + nle[6:]); the result is `self.torque` after each step
        self.torque_law = bool(torque_law)
        if self.torque_law:
            flags |= _native.OPT_TORQUE_LAW
        self._solver = NativeSolver(self.MODEL, n_samples=n_samples, n_horizon=n_horizon, dt=dt, lam=lam, sigma=sigma,
                                    seed=seed, device=device, cost_flags=flags, torque_gains=torque_gains,
                                    philox_rounds=philox_rounds, fused=fused, time_parallel=time_parallel)
        self.torque = np.zeros(7)
        self.device = self._solver.device
        if verbose:
            print(f"[MPPI] Using device: {self.device}")                      # mppi.py:33
        sig = torch.eye(self.n_action, device=self.device) * torch.as_tensor(sigma, dtype=torch.float32, device=self.device)
        self.sample_gen = types.SimpleNamespace(n_sample=self.n_samples, n_horizon=self.n_horizon,
                                                n_action=self.n_action, sigma=sig, device=self.device)
        self.target_pose = Pose()
        self.target_pose.pose = torch.tensor([0.1029, 0.4055, 1.6498])        # mppi.py:71
        self.target_pose.orientation = torch.tensor([-0.5, -0.5, 0.5, -0.5])  # mppi.py:72
        self._target_watch = TensorWatch()
        self.ee_pose = Pose()
        # The measured state is ONE tuple (q, qdot, base xyz+quat, dtype) replaced atomically by update_joint() on the
        # subscriber thread and read once per step: a step never mixes two sensor messages, and the callback never
        # calls into the library.  dtype is float32 until update_joint() feeds numpy doubles.
        self._snap = (np.zeros(7), np.zeros(7), np.array([0, 0, 0, 0, 0, 0, 1.0]), np.float32, np.zeros(6))
        self._state32 = np.zeros(27 if self.torque_law else 21, np.float32)
        self._qdes_np = np.zeros(7, np.float32)
        self._vdes_np = np.zeros(7, np.float32)
        self._dt32 = np.float32(self.dt)
        self.last_costs = None
        self.last_stats = {}
        self.cnt = 0

    # ------------------------------------------------------------------ reference attribute surface
    @property
    def u_prev(self) -> torch.Tensor:
        """Nominal sequence [T][7] (warm start, not shifted: mppi.py:125,153).  A fresh clone, like the reference's
        `u.clone()`: the solver's ping-pong buffers are overwritten two steps later."""
        return self._solver.u_prev.clone()

    @u_prev.setter
    def u_prev(self, value):
        self._solver.u_prev = value

    @property
    def u(self) -> torch.Tensor:
        return self._solver.u_prev[0].clone()

    @property
    def qdes(self) -> torch.Tensor:
        """mppi.py:158 keeps qdes / vdes as tensors; built on demand from the step's numpy results."""
        return torch.from_numpy(self._qdes_np)

    @property
    def vdes(self) -> torch.Tensor:
        return torch.from_numpy(self._vdes_np)

    @property
    def _qddot(self) -> torch.Tensor:
        return self._solver._u[self._solver._cur ^ 1][0].clone()

    @property
    def _q64(self):
        return self._snap[0]

    @property
    def _qdot64(self):
        return self._snap[1]

    @property
    def _base64(self):
        return self._snap[2]

    @property
    def _state_dtype(self):
        return self._snap[3]

    def _replace(self, idx, v):
        snap = list(self._snap)
        snap[idx] = np.asarray(torch.as_tensor(v).detach().cpu().numpy(), np.float64).copy()
        self._snap = tuple(snap)

    @property
    def _q(self) -> torch.Tensor:
        return torch.as_tensor(self._q64.astype(self._state_dtype), device=self.device)

    @_q.setter
    def _q(self, v):
        self._replace(0, v)

    @property
    def _qdot(self) -> torch.Tensor:
        return torch.as_tensor(self._qdot64.astype(self._state_dtype), device=self.device)

    @_qdot.setter
    def _qdot(self, v):
        self._replace(1, v)

    @property
    def base_pose(self) -> torch.Tensor:
        return torch.as_tensor(self._base64.astype(self._state_dtype), device=self.device)

    @base_pose.setter
    def base_pose(self, v):
        self._replace(2, v)

    def update_joint(self, q_full, v_full):
        """mppi.py:196-200.  Safe to call from the subscriber thread while a step is running."""
        q_full = np.asarray(q_full, np.float64)
        v_full = np.asarray(v_full, np.float64)
        if q_full.shape != (14,) or v_full.shape != (13,):
            raise ValueError(f"update_joint expects q[14], v[13]; got {q_full.shape}, {v_full.shape}")
        self._snap = (q_full[7:14].copy(), v_full[6:13].copy(), q_full[:7].copy(), np.float64, v_full[:6].copy())

    # ------------------------------------------------------------------ the control step
    def _sync_target(self):
        if self._target_watch.changed(self.target_pose.pose, self.target_pose.orientation):
            tgt = self.target_pose.as_floats()
            self._solver.set_target(pos=tgt[:3], quat=tgt[3:])

    def compute_control_input(self, noise=None, noise_layout: str = "tkn", return_costs: bool = False):
        """mppi.py:122-169.  `noise`: optional injected noise, [T][K][nu] ("tkn") or the
        reference's [K][T][nu] ("ktn"); default is in-kernel Philox(seed, step counter)."""
        self._sync_target()
        q64, qd64, base64, dtype, twist = self._snap        # one sensor message for the kernel AND the host epilogue
        st = self._state32
        st[0:7] = q64; st[7:14] = qd64; st[14:21] = base64
        if self.torque_law:
            st[21:27] = twist
        q0, qd0 = q64.astype(dtype), qd64.astype(dtype)
        out = self._solver.step(self._solver.prepare_noise(noise, noise_layout), state=st)
        u0 = out[_U0_NEW]                                   # views of the pinned out vector; the arithmetic below copies
        qdd = out[_U0_OLD]
        # mppi.py:157-158 in the reference's own arithmetic and order (float32 control terms, state dtype sum)
        dt = self._dt32
        vdes = qd0 + u0 * dt
        qdes = q0 + qdd * dt + _HALF * u0 * dt * dt
        self._vdes_np, self._qdes_np = vdes, qdes
        if self.torque_law:
            self.torque = out[_TORQUE].astype(np.float64)                       # kinova.py:184
        reach, rho, eta, ess = out[_STATS].tolist()
        self.last_stats = {"rho": rho, "eta": eta, "ess": ess, "reach_err": reach}
        self.cnt += 1
        if reach < 0.005:                                                      # mppi.py:117,165-166
            print("Reach !")
        if return_costs:
            self.last_costs = self._solver.costs.clone()
            return qdes, vdes, self.last_costs
        return qdes, vdes

    def check_reach(self, q_full=None) -> bool:
        """mppi.py:95-120: the position error of FK(base, qdes) is computed by the finalize kernel."""
        return self.last_stats.get("reach_err", float("inf")) < 0.005

    def compute_weights(self, S: torch.Tensor, _lambda) -> torch.Tensor:
        """mppi.py:173-193, kept for API compatibility (the step itself weights on the device)."""
        rho = S.min()
        e = torch.exp((-1.0 / _lambda) * (S - rho))
        return e / e.sum()
