#!/usr/bin/env python
"""Latency / throughput sweep of the MPPI control step over K x T (BASELINE.json configs[4]).

    python tools/sweep.py [--out profiles/r01/sweep_1gpu.json] [--quick]

For every model and (K, T): 5 warm-up steps, then 20 steps timed one by one with CUDA events on the launch
stream (in-kernel Philox noise, warm-started, no L2 flush: the step's working set is S[K] only).  Reports p50 /
p99 device latency and rollout-steps/s.  Single GPU.
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01", "sweep_1gpu.json"))
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    ids = {"wb": _native.MODEL_WB11, "arm": _native.MODEL_ARM7, "drone": _native.MODEL_DRONE3, "quad": _native.MODEL_QUAD4}
    Ks = [1 << e for e in ((10, 14, 18) if a.quick else (10, 12, 14, 16, 18, 20, 22))]
    Ts = (16, 64) if a.quick else (16, 32, 64, 128, 256)
    stream = torch.cuda.current_stream(dev)
    rows = []
    for model in ("drone", "arm", "quad", "wb"):
        for K in Ks:
            for T in Ts:
                qp = (14.7 + 5.5, 1 / 1.57, 1 / 3.93, 1 / 2.59, 0.0, -9.81) if model == "wb" else None
                s = NativeSolver(ids[model], n_samples=K, n_horizon=T, seed=0, device=dev, quad_params=qp)
                s.set_state(synthetic_state(model))
                s.u_prev = torch.from_numpy(nominal_controls(model, T))
                for _ in range(5):
                    s.step_async()
                torch.cuda.synchronize(dev)
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
                for e0, e1 in ev:
                    e0.record(stream)
                    s.step_async()
                    e1.record(stream)
                torch.cuda.synchronize(dev)
                ms = np.array([e0.elapsed_time(e1) for e0, e1 in ev])
                assert torch.isfinite(s.u_prev).all()
                rows.append({"model": model, "nu": MODELS[model]["nu"], "K": K, "T": T, "p50_ms": float(np.percentile(ms, 50)),
                             "p99_ms": float(np.percentile(ms, 99)), "rollout_steps_per_s": K * T / (float(np.median(ms)) * 1e-3)})
                s.close()
                print(rows[-1], flush=True)
    json.dump({"device": torch.cuda.get_device_name(0), "noise": "philox", "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
