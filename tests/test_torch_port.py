"""The eager-PyTorch restatement (oracle/torch_port.py) against the reference's golden vectors and against the
C oracle.  It replays the reference's op sequence, so it lands much closer to the reference's own float32
rounding than the 1e-6 the C oracle is held to."""
import numpy as np
import pytest
import torch

from conftest import fixture_noise, load_golden, rel_inf


def _ktn(noise_tkn):
    return torch.from_numpy(np.ascontiguousarray(noise_tkn.transpose(1, 0, 2)))


@pytest.mark.parametrize("name", ["arm_K64_T32.npz", "arm_K48_T12_tilt.npz", "arm_K64_T32_f64state.npz"])
def test_torch_port_arm_matches_reference(name):
    from oracle import torch_port as tp
    g = load_golden(name)
    dt = torch.float64 if bool(g["f64_state"]) else torch.float32
    q, qd = torch.tensor(g["q"], dtype=dt), torch.tensor(g["qdot"], dtype=dt)
    base = torch.tensor(g["base"], dtype=dt)
    for i in range(len(g["seeds"])):
        noise = _ktn(fixture_noise(g, i, 7))
        o = tp.arm_step(noise, torch.tensor(g[f"u_prev_{i}"]), q, qd, base)
        assert rel_inf(o["S"].numpy(), g[f"S_{i}"]) < 3e-7
        # a single-ulp difference in S moves u_new at the reference's own FP32 noise floor
        e2e = rel_inf(o["u_new"].numpy(), g[f"u_new_{i}"])
        floor = rel_inf(g[f"u_new_{i}"], g[f"u_new_f64_{i}"])
        assert e2e < 1e-4 or rel_inf(o["u_new"].numpy(), g[f"u_new_f64_{i}"]) <= 2 * floor, (e2e, floor)
        assert np.abs(o["qdes"].numpy() - g[f"qdes_{i}"]).max() < 1e-6
        assert o["qdes"].dtype == dt


@pytest.mark.parametrize("name", ["drone_K64_T32.npz"])
def test_torch_port_drone_matches_reference(name):
    from oracle import torch_port as tp
    g = load_golden(name)
    for i in range(len(g["seeds"])):
        noise = _ktn(fixture_noise(g, i, 3))
        x0 = torch.tensor(g["x0"] if i == 0 else g[f"x0_{i}"], dtype=torch.float32)
        v0 = torch.tensor(g["v0"] if i == 0 else g[f"v0_{i}"], dtype=torch.float32)
        o = tp.drone_step(noise, torch.tensor(g[f"u_prev_{i}"]), x0, v0)
        assert rel_inf(o["S"].numpy(), g[f"S_{i}"]) < 3e-7
        assert rel_inf(o["u_new"].numpy(), g[f"u_new_{i}"]) < 1e-5


def test_torch_port_unpinned_models_match_c_oracle(oracle):
    """quad4 / wb11 have no reference: the two independent restatements (C, torch) must agree with each other."""
    from oracle import torch_port as tp
    rng = np.random.default_rng(5)
    K, T = 96, 14
    qs = np.array([0.0, 0.0, 2.1, 0.02, -0.03, 0.1, 0.1, 0.0, -0.05, 0.02, 0.01, -0.03], np.float32)
    q, qd = np.array(oracle.Q_HOME, np.float32), np.array([0.05, -0.1, 0.02, 0.3, -0.2, 0.1, -0.05], np.float32)
    sig = np.array([30 * 20.2, 1, 1, 1] + [0.1] * 7, np.float32)
    noise = (rng.standard_normal((T, K, 11)) * sig).astype(np.float32)
    u = np.zeros((T, 11), np.float32); u[:, 0] = 20.2 * 9.81
    c = oracle.wb_step(noise, u, qs, q, qd)
    t = tp.wb_step(_ktn(noise), torch.tensor(u), torch.tensor(qs), torch.tensor(q), torch.tensor(qd))
    assert rel_inf(t["S"].numpy(), c["S"]) < 2e-6
    nq = (rng.standard_normal((T, K, 4)) * sig[:4] * np.float32(14.7 / 20.2)).astype(np.float32)
    uq = np.zeros((T, 4), np.float32); uq[:, 0] = 14.7 * 9.81
    cq = oracle.quad_step(nq, uq, qs)
    tq = tp.quad_step(_ktn(nq), torch.tensor(uq), torch.tensor(qs))
    assert rel_inf(tq["S"].numpy(), cq["S"]) < 2e-6
