"""torch custom ops `torch.ops.mppi_b200.*`: the thin layer between PyTorch tensors (device
memory, streams) and the C ABI of libmppi_b200.so.

Only the CUDA dispatch key is registered: calling an op with CPU tensors raises
NotImplementedError from the dispatcher -- there is no CPU fallback (contrast the
reference's `cuda if available else cpu`, mppi_solver/mppi.py:32).
"""
from __future__ import annotations

import torch

from . import _native

_lib_def = torch.library.Library("mppi_b200", "DEF")
_lib_def.define("step(int handle, Tensor u_nom, Tensor? noise, int step_counter, Tensor(a!) u_new, "
                "Tensor(b!) out, Tensor(c!)? cost) -> ()")
_lib_def.define("step_sync(int handle, Tensor u_nom, Tensor? noise, int step_counter, Tensor(a!) u_new, "
                "Tensor(b!) out_host) -> ()")
_lib_def.define("step_p2p(int handle, Tensor u_nom, Tensor? noise, int step_counter, Tensor(a!) u_new, "
                "Tensor(b!) out) -> ()")
_lib_def.define("rollout(int handle, Tensor u_nom, Tensor? noise, int step_counter, Tensor(a!)? cost) -> ()")
_lib_def.define("weight(int handle, Tensor? noise, int step_counter) -> ()")
_lib_def.define("finalize(int handle, Tensor u_nom, int step_counter, Tensor(a!) u_new, Tensor(b!) out) -> ()")
_lib_def.define("generate_noise(int handle, int step_counter, Tensor(a!) noise) -> ()")


def _chk(t, name, numel=None):
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous float32 CUDA tensor, got {t.dtype}, contiguous={t.is_contiguous()}")
    if t.data_ptr() % 16:
        raise ValueError(f"{name} must be 16-byte aligned")
    if numel is not None and t.numel() != numel:
        raise ValueError(f"{name} has {t.numel()} elements, expected {numel}")
    return t.data_ptr()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _step_cuda(handle, u_nom, noise, step_counter, u_new, out, cost):
    lib = _native.load()
    _native.check(lib.mppi_step(handle, _chk(u_nom, "u_nom"), None if noise is None else _chk(noise, "noise"),
                                step_counter, None if cost is None else _chk(cost, "cost"),
                                _chk(u_new, "u_new", u_nom.numel()), _chk(out, "out", _native.MPPI_OUT_FLOATS),
                                _stream(u_nom)), handle)


def _step_sync_cuda(handle, u_nom, noise, step_counter, u_new, out_host):
    """Blocking step: out[] lands in `out_host` (a CPU float32 tensor of MPPI_OUT_FLOATS) when this returns."""
    lib = _native.load()
    if out_host.device.type != "cpu" or out_host.dtype != torch.float32 or out_host.numel() != _native.MPPI_OUT_FLOATS:
        raise ValueError("out_host must be a CPU float32 tensor of MPPI_OUT_FLOATS elements")
    _native.check(lib.mppi_step_sync(handle, None, 0, _chk(u_nom, "u_nom"), None if noise is None else _chk(noise, "noise"),
                                     step_counter, _chk(u_new, "u_new", u_nom.numel()), out_host.data_ptr(),
                                     _stream(u_nom)), handle)


def _step_p2p_cuda(handle, u_nom, noise, step_counter, u_new, out):
    """One control step of a K-shard with the peer exchange fused into the weighting kernel (NVLink, no NCCL)."""
    lib = _native.load()
    _native.check(lib.mppi_step_p2p(handle, _chk(u_nom, "u_nom"), None if noise is None else _chk(noise, "noise"),
                                    step_counter, None, _chk(u_new, "u_new", u_nom.numel()),
                                    _chk(out, "out", _native.MPPI_OUT_FLOATS), _stream(u_nom)), handle)


def _rollout_cuda(handle, u_nom, noise, step_counter, cost):
    lib = _native.load()
    _native.check(lib.mppi_rollout(handle, _chk(u_nom, "u_nom"), None if noise is None else _chk(noise, "noise"),
                                   step_counter, None if cost is None else _chk(cost, "cost"), _stream(u_nom)), handle)


def _finalize_cuda(handle, u_nom, step_counter, u_new, out):
    lib = _native.load()
    _native.check(lib.mppi_finalize(handle, _chk(u_nom, "u_nom"), step_counter, _chk(u_new, "u_new", u_nom.numel()),
                                    _chk(out, "out", _native.MPPI_OUT_FLOATS), _stream(u_nom)), handle)


def _generate_noise_cuda(handle, step_counter, noise):
    lib = _native.load()
    _native.check(lib.mppi_generate_noise(handle, step_counter, _chk(noise, "noise"), _stream(noise)), handle)


_lib_def.impl("step", _step_cuda, "CUDA")
_lib_def.impl("step_sync", _step_sync_cuda, "CUDA")
_lib_def.impl("step_p2p", _step_p2p_cuda, "CUDA")
_lib_def.impl("rollout", _rollout_cuda, "CUDA")
_lib_def.impl("finalize", _finalize_cuda, "CUDA")
_lib_def.impl("generate_noise", _generate_noise_cuda, "CUDA")


def weight(handle: int, noise, step_counter: int, device) -> None:
    """`weight` may have no tensor argument (Philox mode), so the dispatcher cannot pick a
    backend from its inputs; it is exposed as a plain function that still only runs on CUDA."""
    lib = _native.load()
    stream = torch.cuda.current_stream(device).cuda_stream
    _native.check(lib.mppi_weight(handle, None if noise is None else _chk(noise, "noise"), step_counter, stream), handle)


step = torch.ops.mppi_b200.step
step_sync = torch.ops.mppi_b200.step_sync
step_p2p = torch.ops.mppi_b200.step_p2p
rollout = torch.ops.mppi_b200.rollout
finalize = torch.ops.mppi_b200.finalize
generate_noise = torch.ops.mppi_b200.generate_noise
