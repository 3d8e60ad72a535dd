"""Drone MPPI controller (live reference model: nu=3 point mass, "Rotation Fixed"):
drop-in for `mppi_solver/drone_mppi.py:7-183`.

Kept: no-arg constructor and defaults (K=1000, T=32, dt=0.01, sigma=30, lambda=0.1;
drone_mppi.py:16-19,32,34), `set_state(x, v)` (:179-183), `compute_control_input()` returning
torch tensors `(x, v)` (:169-176; host tensors here), `u_prev` warm start without shift (:142,166), the hard-coded
target (1.0, 2.0, 3.4) (:141, now the `target` attribute), `param_lambda`, `n_timestep`.
The reference prints rho every step (:123); here it is in `last_stats`.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native
from ..core import NativeSolver
from ..utils.pose import TensorWatch


_STATS = slice(_native.MPPI_OUT_RHO, _native.MPPI_OUT_ESS + 1)


class MPPI:
    MODEL = _native.MODEL_DRONE3

    def __init__(self, *, n_samples: int = 1000, n_timestep: int = 32, dt: float = 0.01, sigma=30.0,
                 lam: float = 0.1, seed: int = 0, device=None, philox_rounds=None, fused=None, time_parallel=None):
        self.n_samples = int(n_samples)
        self.n_timestep = int(n_timestep)
        self.dt = float(dt)
        self.n_action = 3
        self.param_lambda = float(lam)
        self.param_gamma = self.param_lambda * (1.0 - 0.9)                     # drone_mppi.py:35 (unused there too)
        self._solver = NativeSolver(self.MODEL, n_samples=n_samples, n_horizon=n_timestep, dt=dt, lam=lam,
                                    sigma=sigma, seed=seed, device=device, philox_rounds=philox_rounds, fused=fused,
                                    time_parallel=time_parallel)
        self.device = self._solver.device
        self.sigma = torch.eye(3, device=self.device) * torch.as_tensor(sigma, dtype=torch.float32, device=self.device)
        self.target = torch.tensor([1.0, 2.0, 3.4])                            # drone_mppi.py:141
        self._target_watch = TensorWatch()
        self._state = np.zeros(6, np.float32)
        self._solver.set_state(self._state)
        self.last_costs = None
        self.last_stats = {}
        self._last_out = (float("nan"),) * 3

    @property
    def u_prev(self) -> torch.Tensor:
        """Nominal sequence (warm start, not shifted).  A fresh clone: the solver's ping-pong buffers are reused."""
        return self._solver.u_prev.clone()

    @u_prev.setter
    def u_prev(self, value):
        self._solver.u_prev = value

    @property
    def u(self) -> torch.Tensor:
        return self._solver.u_prev[0].clone()

    def set_state(self, x, v):
        """drone_mppi.py:179-183."""
        self._state[:3] = x
        self._state[3:] = v
        self._solver.set_state(self._state)

    # x_prev / v_prev of the reference (drone_mppi.py:24-25) as lazily built device tensors
    @property
    def x_prev(self) -> torch.Tensor:
        return torch.as_tensor(self._state[:3].copy(), device=self.device)

    @property
    def v_prev(self) -> torch.Tensor:
        return torch.as_tensor(self._state[3:].copy(), device=self.device)

    def compute_control_input(self, noise=None, noise_layout: str = "tkn", return_costs: bool = False):
        """drone_mppi.py:140-176."""
        t_ = self.target
        if self._target_watch.changed(t_):                  # re-assigned or edited in place since the last upload
            self._solver.set_target(drone_target=tuple(float(v) for v in torch.as_tensor(t_).reshape(-1)))
        # blocking step: the out vector arrives in pinned host memory (zero-copy store), so the caller's
        # `xdes.to('cpu').tolist()` (drone.py:240) costs nothing more
        out = self._solver.step(self._solver.prepare_noise(noise, noise_layout))
        xv = torch.from_numpy(out[0:6].copy())
        x, v = xv[0:3], xv[3:6]
        self._last_out = out[_STATS].tolist()
        if return_costs:
            self.last_costs = self._solver.costs.clone()
            return x, v, self.last_costs
        return x, v

    def stats(self) -> dict:
        """rho / eta / effective sample size of the last step (the reference prints rho, drone_mppi.py:123)."""
        rho, eta, ess = self._last_out
        self.last_stats = {"rho": rho, "eta": eta, "ess": ess}
        return self.last_stats

    def compute_weights(self, S: torch.Tensor) -> torch.Tensor:
        """drone_mppi.py:111-130 (API compatibility; the step weights on the device)."""
        rho = S.min()
        e = torch.exp((-1.0 / self.param_lambda) * (S - rho))
        return e / e.sum()

    def apply_constraint(self, u: torch.Tensor) -> torch.Tensor:
        """drone_mppi.py:132-138 (disabled in the reference step, kept as a utility)."""
        return torch.clamp(u, min=-10.0, max=10.0)
