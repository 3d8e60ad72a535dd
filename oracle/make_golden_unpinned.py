#!/usr/bin/env python
"""Fixtures for the two UNPINNED models (rigid-body quadrotor nu=4, whole body nu=11).

TEST INFRASTRUCTURE.  The reference holds no runnable code for these models, so there is nothing to
generate reference goldens from.  What CAN be pinned is that the two independent restatements of their specification
 agree and stay that way: this script runs `oracle/torch_port.py` -- eager PyTorch, written to replay
the reference's aten op sequence for the parts the reference does have (cumsum integrator, 4x4 chain, linalg.inv pose
cost, conv1d Savitzky-Golay) -- and freezes its outputs.  `tests/test_oracle_golden.py` then holds the C oracle to
these vectors (so it cannot drift unnoticed) and `tests/test_gpu_parity.py` holds the CUDA path to them.

    python -m oracle.make_golden_unpinned          # writes tests/golden/{quad,wb}_*_torchport.npz

The inputs are seeded numpy draws; nothing here reads /root/reference.
"""
import os

import numpy as np
import torch

from . import torch_port as tp

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
Q_HOME = [1.57, 1.7, 0.0, 4.4, 0.0, 4.71, 0.0]


def ktn(noise_tkn):
    return torch.from_numpy(np.ascontiguousarray(noise_tkn.transpose(1, 0, 2)))


def make_quad(K=128, T=40, seed=101, steps=2):
    rng = np.random.default_rng(seed)
    state = np.array([0.1, -0.2, 2.1, 0.05, -0.08, 0.3, 0.2, -0.1, 0.05, 0.1, -0.2, 0.05], np.float32)
    sig = np.array([30 * 14.7, 1, 1, 1], np.float32)
    u = np.zeros((T, 4), np.float32)
    u[:, 0] = 14.7 * 9.81
    out = dict(K=K, T=T, state=state, sigma=sig, lam=np.float32(0.1), window=5, model="quad4")
    for i in range(steps):
        noise = (rng.standard_normal((T, K, 4)) * sig).astype(np.float32)
        o = tp.quad_step(ktn(noise), torch.tensor(u), torch.tensor(state))
        out[f"noise_{i}"] = noise
        out[f"u_prev_{i}"] = u.copy()
        out[f"S_{i}"] = o["S"].numpy()
        out[f"u_new_{i}"] = o["u_new"].numpy()
        u = o["u_new"].numpy().copy()          # warm start carried, not shifted
    return out


def make_wb(K=96, T=20, seed=202, steps=2):
    rng = np.random.default_rng(seed)
    qs = np.array([0.0, 0.0, 2.1, 0.02, -0.03, 0.1, 0.1, 0.0, -0.05, 0.02, 0.01, -0.03], np.float32)
    q = np.array(Q_HOME, np.float32)
    qd = np.array([0.05, -0.1, 0.02, 0.3, -0.2, 0.1, -0.05], np.float32)
    sig = np.array([30 * 20.2, 1, 1, 1] + [0.1] * 7, np.float32)
    u = np.zeros((T, 11), np.float32)
    u[:, 0] = 20.2 * 9.81
    out = dict(K=K, T=T, qstate=qs, q=q, qdot=qd, sigma=sig, lam=np.float32(0.1), window=9, model="wb11")
    for i in range(steps):
        noise = (rng.standard_normal((T, K, 11)) * sig).astype(np.float32)
        o = tp.wb_step(ktn(noise), torch.tensor(u), torch.tensor(qs), torch.tensor(q), torch.tensor(qd))
        out[f"noise_{i}"] = noise
        out[f"u_prev_{i}"] = u.copy()
        out[f"S_{i}"] = o["S"].numpy()
        out[f"u_new_{i}"] = o["u_new"].numpy()
        u = o["u_new"].numpy().copy()
    return out


def main():
    torch.set_num_threads(1)
    torch.manual_seed(0)
    q = make_quad()
    np.savez_compressed(os.path.join(OUT, f"quad_K{q['K']}_T{q['T']}_torchport.npz"), **q)
    w = make_wb()
    np.savez_compressed(os.path.join(OUT, f"wb_K{w['K']}_T{w['T']}_torchport.npz"), **w)
    print("written:", [f for f in sorted(os.listdir(OUT)) if "torchport" in f])


if __name__ == "__main__":
    main()
