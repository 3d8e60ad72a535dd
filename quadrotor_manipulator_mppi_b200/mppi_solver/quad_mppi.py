"""Rigid-body quadrotor MPPI (nu=4: thrust + three body torques).

PARITY UNPINNED: the reference only holds a commented-out, non-runnable draft of this model
(`mppi_solver/drone_mppi.py:57-83`, helpers `drone.py:114-154,168-185`).  The dynamics run here are
the ones specified in DESIGN.md ("quad4") and restated in oracle/mppi_oracle.c.
State = p[3], rpy[3], v[3], w[3].  Class surface follows the drone controller.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native
from ..core import NativeSolver
from ..utils.pose import TensorWatch


class MPPI:
    MODEL = _native.MODEL_QUAD4

    def __init__(self, *, n_samples: int = 1000, n_timestep: int = 32, dt: float = 0.01, sigma=None,
                 lam: float = 0.1, seed: int = 0, device=None, mass: float = 14.7, philox_rounds=None, fused=None):
        self.n_samples, self.n_timestep, self.dt, self.n_action = int(n_samples), int(n_timestep), float(dt), 4
        self.param_lambda = float(lam)
        self.mass = float(mass)
        if sigma is None:
            sigma = (30.0 * mass, 1.0, 1.0, 1.0)
        qp = (mass, 1.0 / 1.57, 1.0 / 3.93, 1.0 / 2.59, 0.0, -9.81)            # controller.cpp:488-490
        self._solver = NativeSolver(self.MODEL, n_samples=n_samples, n_horizon=n_timestep, dt=dt, lam=lam,
                                    sigma=sigma, seed=seed, device=device, quad_params=qp, philox_rounds=philox_rounds,
                                    fused=fused)
        self.device = self._solver.device
        self.target = torch.tensor([1.0, 2.0, 3.4])
        self._target_watch = TensorWatch()
        self._state = np.zeros(12, np.float32)
        self._solver.set_state(self._state)
        hover = torch.zeros(self.n_timestep, 4)
        hover[:, 0] = mass * 9.81
        self.u_prev = hover                                                     # hover thrust as the nominal

    @property
    def u_prev(self) -> torch.Tensor:
        """Nominal sequence (warm start, not shifted).  A fresh clone: the solver's ping-pong buffers are reused."""
        return self._solver.u_prev.clone()

    @u_prev.setter
    def u_prev(self, value):
        self._solver.u_prev = value

    def set_state(self, p, rpy, v, w):
        self._solver.set_state_parts(p, rpy, v, w)

    def compute_control_input(self, noise=None, noise_layout: str = "tkn"):
        """Returns the one-step-ahead state (p, rpy, v, w) as host tensors (blocking step)."""
        t_ = self.target
        if self._target_watch.changed(t_):                  # re-assigned or edited in place since the last upload
            self._solver.set_target(drone_target=tuple(float(v) for v in torch.as_tensor(t_).reshape(-1)))
        out = torch.from_numpy(self._solver.step(self._solver.prepare_noise(noise, noise_layout))[0:12].copy())
        return out[0:3], out[3:6], out[6:9], out[9:12]
