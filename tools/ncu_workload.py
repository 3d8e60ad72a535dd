#!/usr/bin/env python
"""The short, fixed launch sequence tools/ncu_capture.sh profiles (one launch of each kernel of interest):

  1. wb  K=262144 T=64 Philox, lambda 0.1   : rollout_cost_kernel + weight_philox_kernel (collapsed weights)   [headline]
  2. wb  K=262144 T=64 Philox, dense lambda : rollout_cost_kernel + weight_philox_kernel (all noise regenerated)
  3. wb  K=262144 T=64 injected             : rollout (TMA-staged) + weights_kernel + weighted_noise_kernel   [HBM roofline]
  4. wb  K=32768  T=64 Philox               : the per-GPU shard of the 8-GPU configuration
  5. arm K=1024   T=30 Philox               : step_tp_kernel (time-parallel single launch)                     [BASELINE configs[1]]
  6. arm K=1048576 T=32 Philox              : thread-per-sample pair at scale (weights do not collapse)
  7. quad K=65536 T=100 Philox              : BASELINE configs[2]
"""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)

    def mk(model, K, T, lam=0.1, **kw):
        qp = (14.7 + 5.5, 1 / 1.57, 1 / 3.93, 1 / 2.59, 0.0, -9.81) if model == "wb" else None
        s = NativeSolver(IDS[model], n_samples=K, n_horizon=T, seed=0, device=dev, quad_params=qp, lam=lam,
                         philox_rounds=a.rounds, **kw)
        s.set_state(synthetic_state(model))
        s.u_prev = torch.from_numpy(nominal_controls(model, T))
        return s
    s = mk("wb", 262144, 64)
    s.step_async(); torch.cuda.synchronize()
    s.update_config(lambda_=10.5)
    s.step_async(); torch.cuda.synchronize()
    noise = s.generate_noise(0)
    s.update_config(lambda_=0.1)
    s.step_async(noise); torch.cuda.synchronize()
    del noise
    s.close()
    for model, K, T in (("wb", 32768, 64), ("arm", 1024, 30), ("arm", 1 << 20, 32), ("quad", 65536, 100)):
        s = mk(model, K, T)
        s.step_async(); torch.cuda.synchronize()
        print(model, K, T, s.last_path)
        s.close()


if __name__ == "__main__":
    main()
