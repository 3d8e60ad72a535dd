#!/usr/bin/env python
"""What the fused NVLink exchange costs: the sharded step at N ranks against the SAME per-GPU shard stepped alone
(no exchange), for a tiny and for the bench shard.  torchrun only.  No L2 flush, 300 steps, p50 of CUDA-event times, max over ranks."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import nominal_controls, synthetic_state  # noqa: E402
from quadrotor_manipulator_mppi_b200 import _native  # noqa: E402
from quadrotor_manipulator_mppi_b200.core import NativeSolver  # noqa: E402
from quadrotor_manipulator_mppi_b200.sharded import ShardedStepper  # noqa: E402

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.current_stream(dev)
qp = (14.7 + 5.5, 1 / 1.57, 1 / 3.93, 1 / 2.59, 0.0, -9.81)
rows = []


def timed(fn, n=300):
    for _ in range(20):
        fn()
    dist.barrier()
    torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(stream); fn(); b.record(stream)
    torch.cuda.synchronize(dev)
    ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(np.percentile(ms.cpu().numpy(), 50)) * 1e3


for k_loc in (1024, 32768):
    s = NativeSolver(_native.MODEL_WB11, n_samples=k_loc, n_horizon=64, seed=0, device=dev, quad_params=qp, k_offset=rank * k_loc)
    s.set_state(synthetic_state("wb"))
    s.u_prev = torch.from_numpy(nominal_controls("wb", 64))
    alone = timed(s.step_async)
    st = ShardedStepper(s, exchange="p2p")
    p2p = timed(st.step_async)
    s2 = NativeSolver(_native.MODEL_WB11, n_samples=k_loc, n_horizon=64, seed=0, device=dev, quad_params=qp, k_offset=rank * k_loc)
    s2.set_state(synthetic_state("wb"))
    s2.u_prev = torch.from_numpy(nominal_controls("wb", 64))
    st2 = ShardedStepper(s2, exchange="nccl")
    nccl = timed(st2.step_async)
    rows.append({"K_per_gpu": k_loc, "n_gpus": world, "alone_us": alone, "p2p_us": p2p, "nccl_us": nccl,
                 "p2p_exchange_cost_us": p2p - alone, "nccl_exchange_cost_us": nccl - alone})
    if rank == 0:
        print(rows[-1], flush=True)
if rank == 0:
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"r02_exchange_cost_n{world}.json"), "w"), indent=1)
dist.destroy_process_group()
