"""K-sharded MPPI across the GPUs of one box: one process per GPU (torchrun), NCCL over NVLink.

Samples are independent until the soft-min, so the K dimension is split into contiguous shards
(rank r owns global samples [k_offset, k_offset + k_local)); Philox counters use the GLOBAL sample
index, so the result does not depend on the number of ranks.  Every rank holds the same state and
nominal controls (inputs are ~1 kB and replicated by the caller: no broadcast on the step path).
Per control step the shards exchange exactly two small messages:

    rollout  ->  allreduce-MIN  (1 x int32: order-preserving encoding of the cost baseline rho)
    weight   ->  allreduce-SUM  (T*nu + 2 floats: weighted-noise sums, eta, sum w^2)
    finalize (replicated, deterministic: every rank ends with the same u_new)

Both collectives are latency-bound (<= 2.8 kB); bench.py reports their time separately.

`exchange="p2p"` replaces the two NCCL calls by ONE exchange fused into the weighting kernel over
NVLink peer memory (CUDA IPC buffers, release/acquire flags): each rank weights with its local
minimum and the rows are combined with exp(-(rho_r - rho)/lambda) -- algebraically the same MIN + SUM,
with no collective launch on the step path (csrc/mppi_kernels.cuh: p2p_exchange).
The reference has no multi-GPU path (it pins CUDA_VISIBLE_DEVICES=0, mppi_solver/mppi.py:30-31).
"""
from __future__ import annotations

import struct

import numpy as np
import torch
import torch.distributed as dist

RHO_INIT = 0x7FFFFFFF


def shard_range(n_samples: int, world: int, rank: int):
    """Contiguous split of K over `world` ranks; the first K % world ranks hold one extra sample."""
    if not (0 <= rank < world) or n_samples < world:
        raise ValueError(f"cannot shard K={n_samples} over world={world} (rank {rank})")
    base, extra = divmod(n_samples, world)
    k_local = base + (1 if rank < extra else 0)
    k_offset = rank * base + min(rank, extra)
    return k_offset, k_local


def encode_ordered(value: float) -> int:
    """float32 -> int32 whose signed order equals the float order (the device's atomicMin key)."""
    i = struct.unpack("<i", struct.pack("<f", float(value)))[0]
    return i if i >= 0 else i ^ 0x7FFFFFFF


def decode_ordered(key: int) -> float:
    i = int(key)
    i = i if i >= 0 else i ^ 0x7FFFFFFF
    return struct.unpack("<f", struct.pack("<i", i))[0]


def enable_p2p(solver, group=None) -> None:
    """Exchange CUDA IPC handles of the shards' exchange buffers and bind them (collective call).

    Either every rank ends up bound or every rank raises: the local outcome is agreed on with a MIN
    all-reduce, so a rank whose cudaIpcOpenMemHandle fails cannot leave its peers waiting."""
    import ctypes as C

    from . import _native
    lib = _native.load()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = (C.c_ubyte * _native.MPPI_IPC_HANDLE_BYTES)()
    err = None
    rc = lib.mppi_p2p_export(solver.handle, world, mine)
    if rc:
        err = lib.mppi_last_error(solver.handle).decode()
    gathered = [None] * world
    dist.all_gather_object(gathered, bytes(mine) if err is None else None, group=group)
    if err is None and all(g is not None for g in gathered):
        rc = lib.mppi_p2p_bind(solver.handle, world, rank, b"".join(gathered))
        if rc:
            err = lib.mppi_last_error(solver.handle).decode()
    elif err is None:
        err = "a peer could not export its exchange buffer"
    ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=solver.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)      # also the barrier: nobody steps before all are bound
    if int(ok.item()) == 0:
        raise RuntimeError(f"peer-to-peer exchange unavailable on rank {rank}: {err or 'a peer failed'}")


class ShardedStepper:
    """Drives one shard.  `solver` is a NativeSolver (or anything with rollout / weight / finalize and
    the `rho_enc` int32[1] / `wsum` float32[T*nu+2] exchange tensors).  exchange = "nccl" (allreduce-MIN
    + allreduce-SUM through torch.distributed) or "p2p" (fused NVLink exchange, CUDA only)."""

    def __init__(self, solver, group=None, exchange: str = "nccl"):
        self.solver = solver
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if exchange not in ("nccl", "p2p"):
            raise ValueError("exchange must be 'nccl' or 'p2p'")
        self.exchange = exchange if self.world > 1 else "nccl"
        self.fallback_reason = None
        if self.exchange == "p2p":
            try:
                enable_p2p(solver, group)
            except RuntimeError as e:       # raised on every rank together: fall back to the NCCL contract path
                self.exchange, self.fallback_reason = "nccl", str(e)

    def step_async(self, noise=None):
        s = self.solver
        if self.exchange == "p2p":
            return s.step_p2p_async(noise)
        s.rollout(noise)
        if self.world > 1:
            dist.all_reduce(s.rho_enc, op=dist.ReduceOp.MIN, group=self.group)
        s.weight(noise)
        if self.world > 1:
            dist.all_reduce(s.wsum, op=dist.ReduceOp.SUM, group=self.group)
        return s.finalize()

    def step(self, noise=None, state=None):
        """Blocking step: returns the out vector as a host numpy array (identical on every rank)."""
        s = self.solver
        if self.exchange == "p2p":
            return s.step(noise, state=state, p2p=True)
        if state is not None:
            s.set_state(state)
        out = self.step_async(noise)
        return out.cpu().numpy() if isinstance(out, torch.Tensor) else np.asarray(out)

    @property
    def u_prev(self):
        return self.solver.u_prev


def make_sharded_solver(model: int, n_samples_total: int, n_horizon: int, *, device=None, group=None,
                        exchange: str = "nccl", **kw):
    """NativeSolver for this rank's shard of a K = n_samples_total problem."""
    from .core import NativeSolver
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    k_offset, k_local = shard_range(n_samples_total, world, rank)
    solver = NativeSolver(model, n_samples=k_local, n_horizon=n_horizon, device=device, k_offset=k_offset, **kw)
    return ShardedStepper(solver, group, exchange)
