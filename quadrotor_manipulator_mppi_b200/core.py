"""NativeSolver: one MPPI shard on one B200, the object the drop-in classes wrap.

It owns the library handle, the warm-start buffers and the step counter, and it is
where the [K][T][nu] -> [T][K][nu] transpose of injected reference-layout noise lives
(SURVEY F12).  All arithmetic is in libmppi_b200.so.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import torch

from . import _native, ops


class _DevView:
    """Zero-copy torch view of library-owned device memory (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("quadrotor_manipulator_mppi_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"device {dev} is not CUDA; there is no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class NativeSolver:
    def __init__(self, model: int, *, n_samples=None, n_horizon=None, dt=None, lam=None, sigma=None,
                 seed: int = 0, device=None, k_offset: int = 0, savgol_window=None, cost_w=None,
                 quad_params=None, target_pos=None, target_quat=None, drone_target=None, cost_flags: int = 0,
                 torque_gains=None, philox_rounds=None, fused=None, time_parallel=None):
        self.device = _require_cuda(device)
        self._lib = _native.load()
        cfg = _native.default_config(model)
        if n_samples is not None:
            cfg.n_samples = int(n_samples)
        if n_horizon is not None:
            cfg.n_horizon = int(n_horizon)
        if dt is not None:
            cfg.dt = float(dt)
        if lam is not None:
            cfg.lambda_ = float(lam)
        if savgol_window is not None:
            cfg.savgol_window = int(savgol_window)
        self.nu = _native.MODEL_NU[model]
        if sigma is not None:
            sig = np.broadcast_to(np.asarray(sigma, np.float32), (self.nu,))
            for i in range(self.nu):
                cfg.sigma[i] = float(sig[i])
        for name, val in (("cost_w", cost_w), ("quad_params", quad_params), ("target_pos", target_pos),
                          ("target_quat", target_quat), ("drone_target", drone_target)):
            if val is not None:
                arr = getattr(cfg, name)
                for i, v in enumerate(val):
                    arr[i] = float(v)
        cfg.cost_flags = int(cost_flags)
        if torque_gains is not None:
            cfg.torque_kp, cfg.torque_kd = float(torque_gains[0]), float(torque_gains[1])
        cfg.seed = int(seed) & (2 ** 64 - 1)
        cfg.k_offset = int(k_offset)
        cfg.device = self.device.index
        self.cfg = cfg
        self.model = model
        self.K, self.T = cfg.n_samples, cfg.n_horizon
        h = C.c_void_p()
        _native.check(self._lib.mppi_create(C.byref(cfg), C.byref(h)))
        self.handle = h.value
        if philox_rounds is not None:
            self.set_option(_native.OPTION_PHILOX_ROUNDS, int(philox_rounds))
        if fused is not None:
            self.set_option(_native.OPTION_FUSED_STEP, int(bool(fused)))
        if time_parallel is not None:
            self.set_option(_native.OPTION_TIME_PARALLEL, int(time_parallel))
        self.step_counter = 0
        # warm start: two buffers, alternated so a caller holding last step's u_prev keeps valid data
        self._u = [torch.zeros(self.T, self.nu, device=self.device), torch.zeros(self.T, self.nu, device=self.device)]
        self._cur = 0
        self._outs = [torch.zeros(_native.MPPI_OUT_FLOATS, device=self.device) for _ in range(4)]
        self._out_i = 0
        self._u_ptr = [u.data_ptr() for u in self._u]
        self._out_host = torch.zeros(_native.MPPI_OUT_FLOATS)
        self._out_host_np = self._out_host.numpy()
        self._out_host_ptr = self._out_host.data_ptr()
        n_state = _native.MODEL_STATE[model]
        if model == _native.MODEL_ARM7 and (cfg.cost_flags & _native.OPT_TORQUE_LAW):
            n_state += _native.ARM_TWIST_FLOATS                 # + base twist, read by the torque law
        self._state_np = np.zeros(n_state, np.float32)
        self._state_lock = threading.Lock()
        self._state_ptr = _native.fptr(self._state_np)
        self._set_state_c = self._lib.mppi_set_state
        self.costs = torch.as_tensor(_DevView(self._lib.mppi_cost_ptr(self.handle), self.K, "<f4"), device=self.device)
        self.rho_enc = torch.as_tensor(_DevView(self._lib.mppi_rho_ptr(self.handle), 1, "<i4"), device=self.device)
        self.wsum = torch.as_tensor(_DevView(self._lib.mppi_wsum_ptr(self.handle), self._lib.mppi_wsum_count(self.handle),
                                             "<f4"), device=self.device)

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "handle", None):
            self._lib.mppi_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ options / tracing (include/mppi_b200.h MPPI_OPTION_*)
    def set_option(self, option: int, value: int) -> None:
        _native.check(self._lib.mppi_set_option(self.handle, int(option), int(value)), self.handle)

    def get_option(self, option: int) -> int:
        v = C.c_int32()
        _native.check(self._lib.mppi_get_option(self.handle, int(option), C.byref(v)), self.handle)
        return int(v.value)

    @property
    def last_path(self) -> str:
        """Which kernels the most recent step ran: "two_kernels", "fused" (single launch) or "time_parallel"."""
        return _native.PATH_NAMES[self.get_option(_native.OPTION_LAST_PATH)]

    def profile(self, on: bool = True) -> None:
        """CUDA events around the kernels of every step (SURVEY section 5 tracing hook); read with kernel_times()."""
        self.set_option(_native.OPTION_PROFILE, int(bool(on)))

    def kernel_times(self) -> dict:
        """Device time of the most recent profiled step's kernels in microseconds (blocks until that step is done)."""
        us = (C.c_float * 3)()
        _native.check(self._lib.mppi_get_kernel_times(self.handle, us), self.handle)
        return {"rollout_us": float(us[0]), "weighting_finalize_us": float(us[1]), "path": _native.PATH_NAMES[int(us[2])]}

    def trace(self, on: bool = True) -> None:
        """Phase timestamps inside the kernels (device %globaltimer); read with trace_times()."""
        self.set_option(_native.OPTION_TRACE, int(bool(on)))

    def trace_times(self) -> dict:
        """Microseconds from the start of the most recent traced step to each phase boundary it passed."""
        n = len(_native.TRACE_POINTS)
        buf = (C.c_uint64 * n)()
        _native.check(self._lib.mppi_get_trace(self.handle, buf, n), self.handle)
        t0 = buf[0]
        return {name: (buf[i] - t0) / 1e3 for i, name in enumerate(_native.TRACE_POINTS) if buf[i]}

    # ------------------------------------------------------------------ state / warm start
    @property
    def u_prev(self) -> torch.Tensor:
        """The nominal control sequence [T][nu] the next step warm-starts from (not shifted).

        This is a VIEW of one of the solver's two ping-pong device buffers, valid until the second step after it was
        read (the reference hands out a fresh clone every step, mppi.py:153-154); `.clone()` it to keep it longer.
        The drop-in classes' public `u_prev` / `u` attributes return clones."""
        return self._u[self._cur]

    @u_prev.setter
    def u_prev(self, value):
        t = torch.as_tensor(value, dtype=torch.float32).to(self.device)
        if tuple(t.shape) != (self.T, self.nu):
            raise ValueError(f"u_prev must have shape {(self.T, self.nu)}, got {tuple(t.shape)}")
        self._u[self._cur].copy_(t)

    def set_state(self, state) -> None:
        """Thread-safe: may be called from a subscriber thread while another thread steps."""
        with self._state_lock:
            self._state_np[:] = np.asarray(state, dtype=np.float32).reshape(-1)      # raises on a length mismatch
            rc = self._set_state_c(self.handle, self._state_ptr, self._state_np.size)
        if rc:
            _native.check(rc, self.handle)

    def set_state_parts(self, *parts) -> None:
        """Same, from consecutive slices (avoids building a temporary concatenation on the caller side)."""
        with self._state_lock:
            o = 0
            for p in parts:
                n = len(p)
                self._state_np[o:o + n] = p
                o += n
            if o != self._state_np.size:
                raise ValueError(f"state has {o} entries, expected {self._state_np.size}")
            rc = self._set_state_c(self.handle, self._state_ptr, o)
        if rc:
            _native.check(rc, self.handle)

    def set_target(self, pos=None, quat=None, drone_target=None) -> None:
        def p(v, n):
            if v is None:
                return None
            a = np.ascontiguousarray(np.asarray(v, np.float32).reshape(-1))
            assert a.size == n
            return a
        a, b, c = p(pos, 3), p(quat, 4), p(drone_target, 3)
        _native.check(self._lib.mppi_set_target(self.handle, None if a is None else _native.fptr(a),
                                                None if b is None else _native.fptr(b),
                                                None if c is None else _native.fptr(c)), self.handle)

    def update_config(self, **fields) -> None:
        """Change mutable hyper-parameters (sigma, lambda_, dt, cost_w, cost_flags, gamma, ...) between steps."""
        for name, val in fields.items():
            cur = getattr(self.cfg, name)
            if hasattr(cur, "__len__"):
                for i, v in enumerate(val):
                    cur[i] = float(v)
            else:
                setattr(self.cfg, name, val)
        _native.check(self._lib.mppi_update_config(self.handle, C.byref(self.cfg)), self.handle)

    def set_joint_traj(self, traj) -> None:
        """Joint reference trajectory [T][7] for the optional tracking cost (zeros by default)."""
        t = np.ascontiguousarray(traj, np.float32)
        if t.shape != (self.T, 7):
            raise ValueError(f"joint trajectory must be [{self.T}][7]")
        _native.check(self._lib.mppi_set_joint_traj(self.handle, _native.fptr(t)), self.handle)

    def set_arm_inertia(self, mass, com, inertia) -> None:
        """Link masses [7], centres of mass [7][3], inertias [7][6] (xx xy xz yy yz zz about the centre of mass), in the
        URDF link frames with fixed children merged -- the model the torque law (kinova.py:184) runs on."""
        m, c, i = (np.ascontiguousarray(v, np.float32).reshape(-1) for v in (mass, com, inertia))
        if (m.size, c.size, i.size) != (7, 21, 42):
            raise ValueError("expected mass[7], com[7][3], inertia[7][6]")
        _native.check(self._lib.mppi_set_arm_inertia(self.handle, _native.fptr(m), _native.fptr(c), _native.fptr(i)), self.handle)

    def set_chain(self, types, xyz, rpy, axis) -> None:
        t = np.ascontiguousarray(types, np.int32)
        x, r, a = (np.ascontiguousarray(v, np.float32).reshape(-1, 3) for v in (xyz, rpy, axis))
        _native.check(self._lib.mppi_set_chain(self.handle, int(t.size), t.ctypes.data_as(C.POINTER(C.c_int32)),
                                               _native.fptr(x), _native.fptr(r), _native.fptr(a)), self.handle)

    # ------------------------------------------------------------------ noise
    def prepare_noise(self, noise, layout: str = "tkn"):
        """Injected noise -> contiguous float32 [T][K][nu] on the device ("ktn" = reference layout)."""
        if noise is None:
            return None
        n = torch.as_tensor(noise, dtype=torch.float32).to(self.device)
        if layout == "ktn":
            n = n.permute(1, 0, 2)
        elif layout != "tkn":
            raise ValueError("layout must be 'tkn' or 'ktn'")
        if tuple(n.shape) != (self.T, self.K, self.nu):
            raise ValueError(f"noise must be [T={self.T}][K={self.K}][nu={self.nu}], got {tuple(n.shape)}")
        return n.contiguous()

    def generate_noise(self, step_counter=None) -> torch.Tensor:
        """The Philox noise the in-kernel generator uses for `step_counter`, as [T][K][nu]."""
        out = torch.empty(self.T, self.K, self.nu, device=self.device)
        ops.generate_noise(self.handle, self.step_counter if step_counter is None else int(step_counter), out)
        return out

    # ------------------------------------------------------------------ stepping
    # The per-step entry points call the C ABI directly (ctypes, cached device pointers): a control step at small K is
    # tens of microseconds, and a trip through the torch dispatcher costs about as much.  The same calls are also
    # registered as torch custom ops (ops.py) for callers that want them in a traced / scheduled program.
    def _noise_ptr(self, noise):
        if noise is None:
            return None
        if noise.device != self.device:
            raise ValueError(f"noise is on {noise.device}, the solver on {self.device}")
        return ops._chk(noise, "noise", self.T * self.K * self.nu)

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _advance(self, sc: int) -> None:
        self._cur ^= 1
        self.step_counter = sc + 1

    def step_async(self, noise=None, step_counter=None) -> torch.Tensor:
        """One control step, asynchronous on the current stream.  Returns the device `out` vector."""
        sc = self.step_counter if step_counter is None else int(step_counter)
        out = self._outs[self._out_i]
        self._out_i = (self._out_i + 1) & 3
        rc = self._lib.mppi_step(self.handle, self._u_ptr[self._cur], self._noise_ptr(noise), sc, None,
                                 self._u_ptr[self._cur ^ 1], out.data_ptr(), self._stream())
        if rc:
            _native.check(rc, self.handle)
        self._advance(sc)
        return out

    def step(self, noise=None, step_counter=None, state=None, p2p: bool = False) -> np.ndarray:
        """One blocking control step; returns the `out` vector on the host (stored there by the last block of the
        step; the call spins on its sequence word).

        `state`: optional float32 numpy state vector used for exactly this step (the same call stages it and
        snapshots it), so a caller that keeps its own sensor snapshot needs no separate set_state.
        `p2p`: this solver is one K-shard bound to its peers (sharded.enable_p2p); every rank makes the call."""
        sc = self.step_counter if step_counter is None else int(step_counter)
        if state is None:
            sp, sn = None, 0
        else:
            if state.dtype != np.float32 or not state.flags.c_contiguous or state.size != self._state_np.size:
                raise ValueError(f"state must be a contiguous float32 array of {self._state_np.size} entries")
            sp, sn = _native.fptr(state), state.size
        fn = self._lib.mppi_step_p2p_sync if p2p else self._lib.mppi_step_sync
        rc = fn(self.handle, sp, sn, self._u_ptr[self._cur], self._noise_ptr(noise), sc,
                self._u_ptr[self._cur ^ 1], self._out_host_ptr, self._stream())
        if rc:
            _native.check(rc, self.handle)
        self._advance(sc)
        return self._out_host_np

    def step_p2p_async(self, noise=None, step_counter=None) -> torch.Tensor:
        """Sharded step with the NVLink peer exchange fused in (after sharded.enable_p2p)."""
        sc = self.step_counter if step_counter is None else int(step_counter)
        out = self._outs[self._out_i]
        self._out_i = (self._out_i + 1) & 3
        rc = self._lib.mppi_step_p2p(self.handle, self._u_ptr[self._cur], self._noise_ptr(noise), sc, None,
                                     self._u_ptr[self._cur ^ 1], out.data_ptr(), self._stream())
        if rc:
            _native.check(rc, self.handle)
        self._advance(sc)
        return out

    # ---- the three phases, for K-sharded replicas (see sharded.py)
    def rollout(self, noise=None, step_counter=None):
        sc = self.step_counter if step_counter is None else int(step_counter)
        ops.rollout(self.handle, self._u[self._cur], noise, sc, None)

    def weight(self, noise=None, step_counter=None):
        sc = self.step_counter if step_counter is None else int(step_counter)
        ops.weight(self.handle, noise, sc, self.device)

    def finalize(self, step_counter=None) -> torch.Tensor:
        sc = self.step_counter if step_counter is None else int(step_counter)
        nxt = self._cur ^ 1
        out = self._outs[self._out_i]
        self._out_i = (self._out_i + 1) & 3
        ops.finalize(self.handle, self._u[self._cur], sc, self._u[nxt], out)
        self._cur = nxt
        self.step_counter = sc + 1
        return out
