#!/usr/bin/env python
"""Device-timed table of the step / rollout / weighting kernels over the cases VERDICT r01 names.

    python tools/probe_cases.py [--out gpurun_out/probe.json] [--reps 30]

Every case: 5 warm-up steps, then `reps` steps timed one by one with CUDA events on the launch stream (Philox noise,
warm-started), median reported; rollout and weighting phases timed alone the same way.  `lam` is chosen per case:
0.1 = the reference's value (weights collapse for the drone-position cost), "dense" = a lambda at which ESS >= 1e3.
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

IDS = {"wb": _native.MODEL_WB11, "arm": _native.MODEL_ARM7, "drone": _native.MODEL_DRONE3, "quad": _native.MODEL_QUAD4}
CASES = [
    ("wb", 262144, 64, 0.1), ("wb", 262144, 64, "dense"), ("wb", 32768, 64, 0.1), ("wb", 32768, 64, "dense"),
    ("arm", 1 << 20, 32, 0.1), ("arm", 1024, 30, 0.1), ("arm", 100, 32, 0.1), ("arm", 16384, 32, 0.1),
    ("drone", 1024, 30, 0.1), ("drone", 65536, 100, 0.1), ("quad", 65536, 100, 0.1), ("quad", 65536, 100, "dense"),
    ("quad", 1024, 30, 0.1),
]


def timed(fn, reps, stream, dev):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for e0, e1 in ev:
        e0.record(stream)
        fn()
        e1.record(stream)
    torch.cuda.synchronize(dev)
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def dense_lambda(s, target_ess=2000.0):
    """Bisection on lambda for ESS ~ target on the costs of one step."""
    s.step_async()
    S = s.costs.double()
    S = S - S.min()
    lo, hi = 1e-3, 1e9
    for _ in range(60):
        mid = (lo * hi) ** 0.5
        w = torch.exp(-S / mid)
        ess = (w.sum() ** 2 / (w * w).sum()).item()
        if ess < target_ess:
            lo = mid
        else:
            hi = mid
    return hi


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe.json"))
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--only", default=None, help="comma-separated model filter")
    ap.add_argument("--rounds", type=int, default=None)
    ap.add_argument("--fused", type=int, default=None)
    ap.add_argument("--tp", type=int, default=None)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev)
    rows = []
    for model, K, T, lam in CASES:
        if a.only and model not in a.only.split(","):
            continue
        qp = (14.7 + 5.5, 1 / 1.57, 1 / 3.93, 1 / 2.59, 0.0, -9.81) if model == "wb" else None

        def make(lam_):
            s = NativeSolver(IDS[model], n_samples=K, n_horizon=T, seed=0, device=dev, quad_params=qp, lam=lam_,
                             philox_rounds=a.rounds, fused=a.fused, time_parallel=a.tp)
            s.set_state(synthetic_state(model))
            s.u_prev = torch.from_numpy(nominal_controls(model, T))
            return s
        if lam == "dense":
            s0 = make(0.1)
            lam_v = dense_lambda(s0)
            s0.close()
        else:
            lam_v = lam
        s = make(lam_v)
        for _ in range(5):
            out = s.step_async()
        torch.cuda.synchronize(dev)
        step_ms = timed(s.step_async, a.reps, stream, dev)
        out = s.step_async()
        torch.cuda.synchronize(dev)
        o = out.cpu().numpy()
        nz = int((torch.exp(-(s.costs - s.costs.min()) / lam_v) != 0).sum().item())
        roll_ms = timed(lambda: (s.rho_enc.fill_(0x7fffffff), s.rollout())[0], 3, stream, dev)   # warm
        t_r, t_w = [], []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(max(6, a.reps // 3)):
            e0.record(stream); s.rollout(); e1.record(stream); torch.cuda.synchronize(dev)
            t_r.append(e0.elapsed_time(e1))
            e0.record(stream); s.weight(); e1.record(stream); torch.cuda.synchronize(dev)
            t_w.append(e0.elapsed_time(e1))
            s.finalize()
        row = {"model": model, "K": K, "T": T, "lam": lam_v, "dense": lam == "dense", "step_ms": step_ms,
               "rollout_ms": float(np.median(t_r)), "weight_ms": float(np.median(t_w)),
               "ess": float(o[_native.MPPI_OUT_ESS]), "nonzero_weights": nz, "path": s.last_path,
               "rounds": s.get_option(_native.OPTION_PHILOX_ROUNDS)}
        rows.append(row)
        print(row, flush=True)
        s.close()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"device": torch.cuda.get_device_name(0), "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
