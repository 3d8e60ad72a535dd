#!/usr/bin/env python
"""In-kernel phase timestamps (MPPI_OPTION_TRACE) of one control step for a few cases: where the fixed cost of a step goes."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import nominal_controls, synthetic_state  # noqa: E402
from quadrotor_manipulator_mppi_b200 import _native  # noqa: E402
from quadrotor_manipulator_mppi_b200.core import NativeSolver  # noqa: E402

IDS = {"wb": _native.MODEL_WB11, "arm": _native.MODEL_ARM7, "drone": _native.MODEL_DRONE3, "quad": _native.MODEL_QUAD4}
CASES = [("arm", 1024, 30, {}), ("arm", 1024, 30, {"time_parallel": 0}), ("arm", 100, 32, {}), ("drone", 1024, 30, {}),
         ("drone", 1024, 30, {"time_parallel": 0}), ("wb", 32768, 64, {}), ("wb", 32768, 64, {"fused": 1}), ("quad", 1024, 30, {}),
         ("arm", 16384, 32, {}), ("arm", 4096, 32, {})]
rows = []
for model, K, T, kw in CASES:
    qp = (14.7 + 5.5, 1 / 1.57, 1 / 3.93, 1 / 2.59, 0.0, -9.81) if model == "wb" else None
    s = NativeSolver(IDS[model], n_samples=K, n_horizon=T, seed=0, quad_params=qp, **kw)
    s.set_state(synthetic_state(model))
    s.u_prev = torch.from_numpy(nominal_controls(model, T))
    for _ in range(5):
        s.step_async()
    s.trace(True)
    acc = []
    for _ in range(20):
        torch.cuda.synchronize()
        s.step_async()
        acc.append(s.trace_times())
    keys = acc[0].keys()
    med = {k: float(np.median([a[k] for a in acc if k in a])) for k in keys}
    rows.append({"model": model, "K": K, "T": T, "opts": kw, "path": s.last_path, "us_from_start": med})
    print(rows[-1], flush=True)
    s.close()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r02_trace_phases.json"), "w"), indent=1)
