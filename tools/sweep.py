#!/usr/bin/env python
"""Latency / throughput sweep of the MPPI control step over K x T (BASELINE.json configs[4]), 1 to 8 GPUs.

    python tools/sweep.py [--out profiles/r02/sweep_1gpu.json] [--quick]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/sweep.py --out profiles/r02/sweep_Ngpu.json

For every model and (K, T): 5 warm-up steps, then 20 steps timed one by one with CUDA events on the launch stream
(in-kernel Philox noise, warm-started, no L2 flush: the step's working set is S[K] only), max over ranks per step.
Reports p50 / p99 device latency and rollout-steps/s.  With N ranks two columns are measured:
  strong  -- the same total K sharded over the ranks (fused NVLink exchange),
  weak    -- K samples PER RANK (total N*K), the column that shows what N GPUs buy at a fixed per-GPU load.
After the timed steps every configuration is self-checked: the replicas' updated controls must be bit-identical.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import MODELS, nominal_controls, synthetic_state  # noqa: E402
from quadrotor_manipulator_mppi_b200 import _native  # noqa: E402
from quadrotor_manipulator_mppi_b200.core import NativeSolver  # noqa: E402
from quadrotor_manipulator_mppi_b200.sharded import ShardedStepper, shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--medium", action="store_true", help="K in {1k, 16k, 256k, 4M} x T in {16, 64, 256}: the multi-GPU grid")
    ap.add_argument("--philox-rounds", type=int, default=None)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    out_path = a.out or os.path.join(ROOT, "profiles", "r02", f"sweep_{world}gpu.json")
    ids = {"wb": _native.MODEL_WB11, "arm": _native.MODEL_ARM7, "drone": _native.MODEL_DRONE3, "quad": _native.MODEL_QUAD4}
    Ks = [1 << e for e in ((10, 14, 18) if a.quick else (10, 14, 18, 22) if a.medium else (10, 12, 14, 16, 18, 20, 22))]
    Ts = (16, 64) if a.quick else (16, 64, 256) if a.medium else (16, 32, 64, 128, 256)
    stream = torch.cuda.current_stream(dev)
    rows = []

    def run(model, K_total, T):
        k_off, k_loc = shard_range(K_total, world, rank)
        qp = (14.7 + 5.5, 1 / 1.57, 1 / 3.93, 1 / 2.59, 0.0, -9.81) if model == "wb" else None
        s = NativeSolver(ids[model], n_samples=k_loc, n_horizon=T, seed=0, device=dev, quad_params=qp, k_offset=k_off,
                         philox_rounds=a.philox_rounds)
        s.set_state(synthetic_state(model))
        s.u_prev = torch.from_numpy(nominal_controls(model, T))
        st = ShardedStepper(s, exchange="p2p") if world > 1 else None
        step = (lambda: st.step_async()) if world > 1 else s.step_async
        for _ in range(5):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for e0, e1 in ev:
            e0.record(stream)
            step()
            e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1) for e0, e1 in ev], dtype=torch.float64, device=dev)
        identical = True
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            u = s.u_prev.clone()
            g = [torch.empty_like(u) for _ in range(world)]
            dist.all_gather(g, u)
            identical = bool(all(torch.equal(g[0], x) for x in g))
        ms = ms.cpu().numpy()
        finite = bool(torch.isfinite(s.u_prev).all())
        path, exch = s.last_path, (st.exchange if st else None)
        s.close()
        return {"p50_ms": float(np.percentile(ms, 50)), "p99_ms": float(np.percentile(ms, 99)),
                "rollout_steps_per_s": K_total * T / (float(np.median(ms)) * 1e-3), "K_total": K_total, "K_per_gpu": k_loc,
                "path": path, "exchange": exch, "replicas_identical": identical, "finite": finite}

    for model in ("drone", "arm", "quad", "wb"):
        for K in Ks:
            for T in Ts:
                row = {"model": model, "nu": MODELS[model]["nu"], "K": K, "T": T, "n_gpus": world, "strong": run(model, K, T)}
                if world > 1 and K * world <= (1 << 24):
                    row["weak"] = run(model, K * world, T)
                assert row["strong"]["finite"] and row["strong"]["replicas_identical"], row
                rows.append(row)
                if rank == 0:
                    print(row, flush=True)
    if rank == 0:
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        json.dump({"device": torch.cuda.get_device_name(0), "n_gpus": world, "noise": "philox",
                   "philox_rounds": a.philox_rounds or "library default", "rows": rows}, open(out_path, "w"), indent=1)
        md = out_path[:-5] + ".md"
        with open(md, "w") as f:
            f.write(f"# Control-step latency sweep, {world} x B200 (Philox noise, device time, p50 of 20 steps, max over ranks)\n\n")
            f.write("`strong`: the K of the row sharded over the GPUs; `weak`: K samples per GPU (N x K in total).\n\n")
            f.write("| model | K | T | strong p50 ms | strong rollout-steps/s | path |" + (" weak p50 ms (N*K) | weak rollout-steps/s |" if world > 1 else "") + "\n")
            f.write("|---|---|---|---|---|---|" + ("---|---|" if world > 1 else "") + "\n")
            for r in rows:
                line = f"| {r['model']} | {r['K']} | {r['T']} | {r['strong']['p50_ms']:.4f} | {r['strong']['rollout_steps_per_s']:.3e} | {r['strong']['path']} |"
                if world > 1:
                    w = r.get("weak")
                    line += f" {w['p50_ms']:.4f} | {w['rollout_steps_per_s']:.3e} |" if w else " – | – |"
                f.write(line + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
