"""world_size-2 `gloo` tests of the K-sharding host logic on CPU.

The GPU kernels cannot run here, so each rank's "solver" is a stand-in that computes its shard with
the CPU oracle (test infrastructure) behind the same rollout / weight / finalize + exchange-tensor
interface NativeSolver exposes.  What is under test is the product's host side: `shard_range`, the
order-preserving int32 key, and `ShardedStepper`'s MIN / SUM exchange sequence.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_range_and_ordered_key():
    from quadrotor_manipulator_mppi_b200.sharded import decode_ordered, encode_ordered, shard_range
    for K, W in ((262144, 8), (1000, 3), (7, 7), (1025, 4)):
        spans = [shard_range(K, W, r) for r in range(W)]
        assert spans[0][0] == 0 and sum(n for _, n in spans) == K
        for (o0, n0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + n0 == o1
        assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(3, 4, 0)
    vals = np.array([-3.5e7, -1.0, -1e-30, -0.0, 0.0, 1e-38, 0.1, 1049.03, 3.4e38], np.float32)
    keys = [encode_ordered(v) for v in vals]
    assert keys == sorted(keys) and all(-2 ** 31 <= k < 2 ** 31 for k in keys)
    assert all(np.float32(decode_ordered(k)) == v for k, v in zip(keys, vals))
    assert encode_ordered(float("inf")) < 0x7FFFFFFF            # the re-arm value is above every cost


class OracleShard:
    """NativeSolver stand-in: one shard computed by the CPU oracle."""

    def __init__(self, orc, K_total, T, k_offset, k_local, noise_full, u_prev):
        from quadrotor_manipulator_mppi_b200.sharded import RHO_INIT
        self.orc, self.T, self.nu = orc, T, 7
        self.noise = np.ascontiguousarray(noise_full[:, k_offset:k_offset + k_local])
        self.u = u_prev.copy()
        self.rho_enc = torch.tensor([RHO_INIT], dtype=torch.int32)
        self.wsum = torch.zeros(T * 7 + 2)

    def rollout(self, noise=None):
        from quadrotor_manipulator_mppi_b200.sharded import encode_ordered
        self.S = self.orc.arm_costs(self.noise, self.u, self.orc.Q_HOME, np.zeros(7), [0, 0, 2.1, 0, 0, 0, 1])
        self.rho_enc[0] = min(int(self.rho_enc[0]), encode_ordered(self.S.min()))

    def weight(self, noise=None):
        from quadrotor_manipulator_mppi_b200.sharded import decode_ordered
        rho = np.float32(decode_ordered(int(self.rho_enc[0])))
        w = np.exp(np.float32(-10.0) * (self.S - rho)).astype(np.float32)
        raw = np.einsum("k,tki->ti", w.astype(np.float64), self.noise.astype(np.float64))
        self.wsum[:-2] = torch.from_numpy(raw.reshape(-1).astype(np.float32))
        self.wsum[-2] = float(w.astype(np.float64).sum())
        self.wsum[-1] = float((w.astype(np.float64) ** 2).sum())

    def finalize(self):
        from quadrotor_manipulator_mppi_b200.sharded import RHO_INIT
        ws = self.wsum.numpy()
        raw = (ws[:-2] / ws[-2]).reshape(self.T, 7).astype(np.float32)
        self.u = (self.u + self.orc.savgol(raw, 9)).astype(np.float32)
        self.rho_enc[0] = RHO_INIT
        return self.u


def _worker(rank, world, port, K, T, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        from quadrotor_manipulator_mppi_b200.sharded import ShardedStepper, shard_range
        orc.set_threads(1)
        noise = orc.philox_noise(K, T, 7, 0.1, seed=3, step=0)
        k_off, k_loc = shard_range(K, world, rank)
        u = np.zeros((T, 7), np.float32)
        stepper = ShardedStepper(OracleShard(orc, K, T, k_off, k_loc, noise, u))
        assert stepper.world == world
        for step in range(2):           # two steps: the minimum is re-armed, the warm start carries over
            u = stepper.step_async()
        np.save(os.path.join(out_dir, f"u_{rank}.npy"), u)
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_matches_unsharded(oracle, tmp_path):
    K, T, world = 301, 12, 2            # ragged split: 151 + 150
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, K, T, str(tmp_path)), nprocs=world, join=True)
    u0, u1 = (np.load(tmp_path / f"u_{r}.npy") for r in range(world))
    assert np.array_equal(u0, u1)                       # replicas finish identical
    noise = oracle.philox_noise(K, T, 7, 0.1, seed=3, step=0)
    u = np.zeros((T, 7), np.float32)
    for step in range(2):
        u = oracle.arm_step(noise, u, oracle.Q_HOME, np.zeros(7), [0, 0, 2.1, 0, 0, 0, 1])["u_new"]
    assert np.abs(u0 - u).max() / np.abs(u).max() < 1e-5
