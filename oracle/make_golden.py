"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

    python oracle/make_golden.py          (build container only: needs /root/reference)

TEST INFRASTRUCTURE ONLY.  The reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so parity is pinned by executing the reference classes themselves
through `oracle/ref_harness.py` and committing small input/output fixtures.  Layout
conventions inside the fixtures:

* noise is stored in the NATIVE boundary layout [T][K][nu] (the reference consumed its
  transpose [K][T][nu], standard_normal_noise.py:24);
* big cases store only the torch CPU generator seed plus a float64 checksum of the noise,
  and tests regenerate it with `golden_noise()` below (same code path).

Every array in a fixture is an output of reference code, except `*_f64` entries which
come from the reference run with float64 tensors (two dtype-pinned helper functions
patched, see `_arm_f64`) and serve as the "FP32 noise floor" of SURVEY section 8(c).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

Q_HOME = [1.57, 1.7, 0.0, 4.4, 0.0, 4.71, 0.0]           # kinova.py:135


def golden_noise(seed: int, K: int, T: int, nu: int, sigma: float) -> torch.Tensor:
    """Reference-layout noise [K][T][nu] = randn * sigma from a seeded CPU generator."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(K, T, nu, generator=g, dtype=torch.float32) * sigma


def _tkn(noise_ktn: torch.Tensor) -> np.ndarray:
    return noise_ktn.permute(1, 0, 2).contiguous().numpy()


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


# --------------------------------------------------------------------------- unit pins
def make_unit_pins():
    rh.install_shims()
    with rh.quiet():
        from robot.urdf_fk import URDFFK
        from utils.rotation_conversions import quaternion_to_matrix, matrix_to_euler_angles
        from filter.svg_filter import SavGolFilter
        fk = URDFFK(os.path.join(rh.REF_AERIAL, "urdf", "aerial_manipulator_gpu.urdf"),
                    root_link="base", end_link="j2s7s300_link_7")
    g = torch.Generator().manual_seed(11)
    qs = torch.cat([torch.tensor([Q_HOME, [0.0] * 7]), (torch.rand(14, 7, generator=g) - 0.5) * 12.0])
    bases = torch.tensor([[0, 0, 0, 0, 0, 0, 1.0], [0, 0, 2.1, 0, 0, 0, 1.0],
                          [0.3, -0.2, 1.7, 0.0499792, -0.0998334, 0.1494381, 0.9824485],
                          [-1.0, 0.5, 3.0, 0.5, 0.5, -0.5, 0.5]])
    fk_single = np.stack([np.stack([fk.compute_fk_cpu(b, q) for q in qs]) for b in bases])
    fk.robot._n_samples, fk.robot._n_timestep = 1, 1
    fk_batched = np.stack([_np(fk.compute_fk_gpu(qs.unsqueeze(0), b))[0] for b in bases])
    chain = fk.robot._joint_chain_list
    quats = torch.cat([torch.tensor([[-0.5, -0.5, 0.5, -0.5], [0, 0, 0, 1.0], [0, -0.4871745, 0, -0.8733046]]),
                       torch.randn(8, 4, generator=g)])
    Rq = _np(quaternion_to_matrix(quats))
    eul = _np(matrix_to_euler_angles(torch.tensor(Rq), "ZYX"))
    sg7, sg3 = SavGolFilter(7), SavGolFilter(3)
    seq7 = torch.randn(32, 7, generator=g)
    seq3 = torch.randn(32, 3, generator=g)
    ramp = torch.arange(32, dtype=torch.float32).unsqueeze(1).repeat(1, 7)
    short = torch.randn(6, 3, generator=g)
    np.savez_compressed(
        os.path.join(OUT, "unit_pins.npz"),
        fk_q=_np(qs), fk_base=_np(bases), fk_single=fk_single, fk_batched=fk_batched,
        chain_names=np.array([j.name for j in chain]), chain_types=np.array([j.type for j in chain]),
        chain_xyz=np.array([j.origin.xyz for j in chain], np.float64),
        chain_rpy=np.array([j.origin.rpy for j in chain], np.float64),
        chain_axis=np.array([j.axis for j in chain], np.float64),
        quats=_np(quats), quat_R=Rq, euler_zyx=eul,
        sg_seq7=_np(seq7), sg_out7_w9=_np(sg7.savgol_filter_torch(seq7, 9, 2)),
        sg_seq3=_np(seq3), sg_out3_w5=_np(sg3.savgol_filter_torch(seq3, 5, 2)),
        sg_ramp_w9=_np(sg7.savgol_filter_torch(ramp, 9, 2)),
        sg_short3=_np(short), sg_short3_w5=_np(sg3.savgol_filter_torch(short, 5, 2)),
    )


# --------------------------------------------------------------------------- arm
def _seat_arm(m, q, qdot, base, f64_via_update_joint=False):
    if f64_via_update_joint:   # the ROS path: numpy f64 -> float64 tensors (mppi.py:196-200)
        q_full = np.concatenate([np.asarray(base, np.float64), np.asarray(q, np.float64)])
        v_full = np.concatenate([np.zeros(6), np.asarray(qdot, np.float64)])
        m.update_joint(q_full, v_full)
    else:
        m._q = torch.tensor(q, dtype=torch.float32)
        m._qdot = torch.tensor(qdot, dtype=torch.float32)
        m.base_pose = torch.tensor(base, dtype=torch.float32)


def _arm_f64(K, T, q, qdot, base, noises):
    """Reference arm step with float64 tensors: the FP32 noise floor.

    Two helpers pin float32 explicitly (urdf_fk.py:40,52 and svg_filter.py:52); they are
    replaced by dtype-following equivalents for this run only."""
    m = rh.load_arm(K, T)
    torch.set_default_dtype(torch.float64)
    try:
        m.sample_gen.sigma = m.sample_gen.sigma.double()
        m.u_prev = m.u_prev.double()
        m.target_pose.pose = m.target_pose.pose.double()
        m.target_pose.orientation = m.target_pose.orientation.double()
        m._q, m._qdot = torch.tensor(q).double(), torch.tensor(qdot).double()
        m.base_pose = torch.tensor(base).double()
        m.fk_urdf.robot._tf_fk = torch.eye(4)            # built as f32 in the ctor (urdfparser.py:33)

        def xyzquat(b):
            Tm = torch.eye(4)
            Tm[:3, 3] = b[:3]
            x, y, z, w = b[3:]
            Tm[:3, :3] = torch.tensor([
                [1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w],
                [2 * x * y + 2 * z * w, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * x * w],
                [2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x * x - 2 * y * y]])
            return Tm

        m.fk_urdf.xyzquat_to_matrix = xyzquat

        def savgol64(seq, window_size, polyorder):
            h = window_size // 2
            x = torch.arange(-h, h + 1, dtype=torch.float64)
            A = torch.stack([x ** i for i in range(polyorder + 1)], dim=1)
            taps = (torch.linalg.inv(A.T @ A) @ A.T)[0]
            cols = []
            for i in range(seq.shape[1]):
                d = seq[:, i]
                p = torch.cat([d[:h].flip(0), d, d[-h:].flip(0)])
                cols.append(torch.stack([(p[t:t + window_size] * taps).sum() for t in range(d.numel())]))
            return torch.stack(cols, dim=1)

        outs = []
        for noise in noises:
            cap = {}
            m.sample_gen.sampling = lambda n=noise: n.double()
            orig_w = type(m).compute_weights.__get__(m)

            def cw(S, lam, cap=cap):
                cap["S"] = S.clone()
                return orig_w(S, lam)

            m.compute_weights = cw
            m.svg_filter.savgol_filter_torch = lambda seq, window_size, polyorder: savgol64(seq, window_size, polyorder)
            m.check_reach = lambda *_: False       # side effect only (mppi.py:165-169), f32-pinned inside
            with rh.quiet():
                m.compute_control_input()
            outs.append((_np(cap["S"]), _np(m.u_prev)))
        return outs
    finally:
        torch.set_default_dtype(torch.float32)


def make_arm(name, K, T, seeds, q, qdot, base, store_noise, f64_state=False, with_f64=False):
    m = rh.load_arm(K, T)
    _seat_arm(m, q, qdot, base, f64_state)
    rec = dict(K=K, T=T, q=np.asarray(q, np.float64), qdot=np.asarray(qdot, np.float64),
               base=np.asarray(base, np.float64), seeds=np.asarray(seeds), sigma=0.1, lam=0.1, dt=0.01,
               f64_state=f64_state)
    noises = [golden_noise(s, K, T, 7, 0.1) for s in seeds]
    for i, noise in enumerate(noises):
        u_prev = _np(m.u_prev).copy()
        cap = rh.arm_step(m, noise)
        rec[f"u_prev_{i}"] = u_prev
        if store_noise:
            rec[f"noise_{i}"] = _tkn(noise)
        rec[f"noise_sum_{i}"] = np.float64(noise.double().sum().item())
        rec[f"noise_abs_sum_{i}"] = np.float64(noise.double().abs().sum().item())
        for key in ("S", "w", "w_eps_raw", "w_eps", "u_new", "qdes", "vdes"):
            rec[f"{key}_{i}"] = _np(cap[key])
    if with_f64:
        for i, (S64, u64) in enumerate(_arm_f64(K, T, q, qdot, base, noises)):
            rec[f"S_f64_{i}"], rec[f"u_new_f64_{i}"] = S64, u64
    np.savez_compressed(os.path.join(OUT, name), **rec)


def make_arm_extra(name, K, T, seed, q, qdot, base):
    """Reference arm step with the commented-out cost terms of cost_manager.py:83-87 re-enabled, one at a
    time and all together, by calling the reference's own (constructed but unused) cost objects."""
    rec = dict(K=K, T=T, q=np.asarray(q, np.float64), qdot=np.asarray(qdot, np.float64), base=np.asarray(base, np.float64),
               seeds=np.asarray([seed]), sigma=0.1, lam=0.1, dt=0.01, f64_state=False)
    noise = golden_noise(seed, K, T, 7, 0.1)
    rec["noise_0"] = _tkn(noise)
    g = torch.Generator().manual_seed(seed + 100)
    u_prev = torch.randn(T, 7, generator=g) * 0.2          # non-zero nominal so the covariance term is non-trivial
    rec["u_prev_0"] = _np(u_prev)
    for label, flags in (("covar", 1), ("centering", 2), ("joint_traj", 4), ("action", 8), ("joint_limit", 16), ("all", 31)):
        m = rh.load_arm(K, T)
        _seat_arm(m, q, qdot, base)
        m.u_prev = u_prev.clone()
        cm = m.cost_manager

        def all_cost(cm=cm, flags=flags):
            S = torch.zeros((cm.n_sample), device=cm.device)
            S += cm.pose_cost.compute_stage_cost(cm.eef_trajectories, cm.target)
            S += cm.pose_cost.compute_terminal_cost(cm.eef_trajectories, cm.target)
            if flags & 1:
                S += cm.covar_cost.compute_covar_cost(cm.sigma_matrix, cm.u, cm.v)
            if flags & 2:
                S += cm.joint_cost.compute_centering_cost(cm.qSamples)
            if flags & 4:
                S += cm.joint_cost.compute_jointTraj_cost(cm.qSamples, cm.joint_trajectories)
            if flags & 8:
                S += cm.action_cost.compute_action_cost(cm.uSamples)
            if flags & 16:
                S += cm.joint_cost.compute_joint_limit_cost(cm.qSamples)
            return S

        cm.compute_all_cost = all_cost
        cap = rh.arm_step(m, noise)
        rec[f"S_{label}"] = _np(cap["S"])
        rec[f"u_new_{label}"] = _np(cap["u_new"])
    np.savez_compressed(os.path.join(OUT, name), **rec)


# --------------------------------------------------------------------------- drone
def make_drone(name, K, T, seeds, x0, v0, store_noise):
    m = rh.load_drone(K, T)
    m.set_state(np.asarray(x0, np.float64), np.asarray(v0, np.float64))
    rec = dict(K=K, T=T, x0=np.asarray(x0, np.float64), v0=np.asarray(v0, np.float64),
               seeds=np.asarray(seeds), sigma=30.0, lam=0.1, dt=0.01)
    for i, s in enumerate(seeds):
        noise = golden_noise(s, K, T, 3, 30.0)
        u_prev = _np(m.u_prev).copy()
        cap = rh.drone_step(m, noise)
        rec[f"u_prev_{i}"] = u_prev
        if store_noise:
            rec[f"noise_{i}"] = _tkn(noise)
        rec[f"noise_sum_{i}"] = np.float64(noise.double().sum().item())
        rec[f"noise_abs_sum_{i}"] = np.float64(noise.double().abs().sum().item())
        for key in ("S", "w", "w_eps_raw", "w_eps", "u_new", "x", "v"):
            rec[f"{key}_{i}"] = _np(cap[key])
        # closed loop like drone.py:164-165: the node feeds the measured state back; here the
        # predicted one-step state stands in for the simulator
        m.set_state(_np(cap["x"]).astype(np.float64), _np(cap["v"]).astype(np.float64))
        rec[f"x0_{i + 1}"], rec[f"v0_{i + 1}"] = _np(cap["x"]), _np(cap["v"])
    np.savez_compressed(os.path.join(OUT, name), **rec)


def main():
    assert rh.reference_available(), "reference tree not mounted"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    make_unit_pins()
    base_hover = [0, 0, 2.1, 0, 0, 0, 1.0]
    base_tilt = [0.3, -0.2, 1.7, 0.0499792, -0.0998334, 0.1494381, 0.9824485]
    qd = [0.05, -0.1, 0.02, 0.3, -0.2, 0.1, -0.05]
    make_arm("arm_K64_T32.npz", 64, 32, [1, 2, 3], Q_HOME, [0.0] * 7, base_hover, store_noise=True, with_f64=True)
    make_arm("arm_K48_T12_tilt.npz", 48, 12, [4, 5], [1.2, 2.0, -0.4, 4.0, 0.7, 4.2, -1.0], qd, base_tilt,
             store_noise=True, with_f64=True)
    make_arm("arm_K64_T32_f64state.npz", 64, 32, [1, 2], Q_HOME, qd, base_tilt, store_noise=False, f64_state=True, with_f64=True)
    make_arm("arm_K1024_T30.npz", 1024, 30, [0, 7], Q_HOME, [0.0] * 7, base_hover, store_noise=False, with_f64=True)
    make_arm_extra("arm_extra_costs.npz", 64, 32, 9, [1.2, 0.8513, -0.4, 4.0, 0.7, 4.2, -1.0], qd, base_tilt)
    make_drone("drone_K64_T32.npz", 64, 32, [1, 2, 3], [0, 0, 2.1], [0, 0, 0], store_noise=True)
    make_drone("drone_K1024_T30.npz", 1024, 30, [0, 7], [0, 0, 2.1], [0.2, -0.1, 0.05], store_noise=False)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
