#!/usr/bin/env python
"""bench.py -- rollout-steps/s and control-step latency of the B200 MPPI step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model wb|arm|drone|quad]
                    [--samples K] [--horizon T] [--noise philox|injected] [--impl reference]

One "step" = one MPPI control step (noise + rollout + cost + weighting + update) over the
whole sample batch.  Default workload = BASELINE.json configs[3], the configuration the
target is quoted on: whole-body quadrotor+Kinova, nu=11, K=262144, T=64, in-kernel Philox.
With N>1 (torchrun, one rank per GPU) the K samples are sharded across ranks (strong
scaling of a fixed K) with an allreduce-MIN of the cost baseline and an allreduce-SUM of the
[T*nu+2] weighted-noise buffer per step.

Prints ONE JSON line (rank 0).  `value` = K*T*steps / device time (CUDA events around every
step, max over ranks, L2 flushed between steps outside the events).  `e2e` = the same metric
through the public controller class with HOST state in / HOST controls out every step.
`roofline` is measured live for the dominant kernel (fused rollout+cost) against an FP32 FFMA
probe run in the same process.  `cpu_baseline` = the CPU oracle port (oracle/, C + OpenMP) on a
bounded sample of the same workload on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODELS = {
    "wb": dict(nu=11, K=262144, T=64, desc="whole-body quadrotor+Kinova MPPI nu=11 (BASELINE.json configs[3])"),
    "arm": dict(nu=7, K=1024, T=30, desc="Kinova 7-DOF arm MPPI nu=7 (BASELINE.json configs[1])"),
    "drone": dict(nu=3, K=1024, T=30, desc="point-mass drone MPPI nu=3 (reference live controller, configs[0])"),
    "quad": dict(nu=4, K=65536, T=100, desc="rigid-body quadrotor MPPI nu=4 (BASELINE.json configs[2])"),
}
Q_HOME = [1.57, 1.7, 0.0, 4.4, 0.0, 4.71, 0.0]     # kinova.py:135


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--model", default="wb", choices=list(MODELS))
    ap.add_argument("--samples", type=int, default=None)
    ap.add_argument("--horizon", type=int, default=None)
    ap.add_argument("--noise", default="philox", choices=["philox", "injected"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU shard exchange: fused NVLink peer exchange, or NCCL allreduce-MIN + allreduce-SUM")
    ap.add_argument("--no-flush", action="store_true", help="skip the L2 flush between steps (latency experiments)")
    return ap.parse_args()


def synthetic_state(model: str) -> np.ndarray:
    """SURVEY 8(d) synthetic hover-to-goal state."""
    if model == "drone":
        return np.array([0, 0, 2.1, 0, 0, 0], np.float32)
    if model == "quad":
        return np.array([0, 0, 2.1] + [0] * 9, np.float32)
    if model == "arm":
        return np.array(Q_HOME + [0] * 7 + [0, 0, 2.1, 0, 0, 0, 1], np.float32)
    return np.array([0, 0, 2.1] + [0] * 9 + Q_HOME + [0] * 7, np.float32)


def nominal_controls(model: str, T: int) -> np.ndarray:
    nu = MODELS[model]["nu"]
    u = np.zeros((T, nu), np.float32)
    if model == "quad":
        u[:, 0] = 14.7 * 9.81
    if model == "wb":
        u[:, 0] = (14.7 + 5.5) * 9.81
    return u


# ------------------------------------------------------------------------------------------ CPU oracle legs
def oracle_step_fn(model: str, K: int, T: int):
    """One control step of the CPU oracle port on K samples (noise generation included)."""
    from oracle import oracle as orc
    orc.build()
    nu = MODELS[model]["nu"]
    sigma = {"wb": [30 * 20.2, 1, 1, 1] + [0.1] * 7, "arm": [0.1] * 7, "drone": [30.0] * 3, "quad": [30 * 14.7, 1, 1, 1]}[model]
    st = synthetic_state(model)
    u = nominal_controls(model, T)
    counter = [0]

    def step():
        noise = orc.philox_noise(K, T, nu, sigma, seed=0, step=counter[0])
        counter[0] += 1
        if model == "wb":
            return orc.wb_step(noise, u, st[:12], st[12:19], st[19:26])
        if model == "arm":
            return orc.arm_step(noise, u, st[:7], st[7:14], st[14:21])
        if model == "drone":
            return orc.drone_step(noise, u, st[:3], st[3:6])
        return orc.quad_step(noise, u, st)
    return step, orc


def cpu_baseline(model: str, K: int, T: int, budget_s: float, n_steps: int = 1) -> dict:
    """Time the oracle port on a bounded sample sized to ~budget_s of CPU work in total."""
    probe_K = min(K, 1024)
    step, orc = oracle_step_fn(model, probe_K, T)
    threads = orc.set_threads(len(os.sched_getaffinity(0)))     # all usable host cores, whatever OMP_NUM_THREADS says
    step()
    t0 = time.perf_counter()
    step()
    rate = probe_K * T / max(time.perf_counter() - t0, 1e-6)
    Ks = int(min(K, max(1024, rate * budget_s / (T * max(n_steps, 1)))))
    step, _ = oracle_step_fn(model, Ks, T)
    times = []
    for _ in range(max(n_steps, 1)):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return {"value": Ks * T / med, "unit": "rollout-steps/s", "cores": threads, "kind": "port",
            "sample": f"{Ks} of {K} samples x T={T}, {len(times)} step(s), oracle/mppi_oracle.c with OpenMP "
                      f"({threads} threads of {len(os.sched_getaffinity(0))} usable cores), Philox noise generation included",
            "ms_per_step_sample": med * 1e3}


def cpu_baseline_torch(model: str, K: int, T: int, budget_s: float = 6.0) -> dict:
    """Time oracle/torch_port.py -- the eager-PyTorch restatement that replays the reference's aten op sequence
    (the reference's own CPU PyTorch path, BASELINE.json north_star) -- on a bounded sample, all host cores."""
    import torch
    from oracle import torch_port as tp
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    nu = MODELS[model]["nu"]
    st = torch.from_numpy(synthetic_state(model))
    sig = torch.tensor({"wb": [30 * 20.2, 1, 1, 1] + [0.1] * 7, "arm": [0.1] * 7, "drone": [30.0] * 3, "quad": [30 * 14.7, 1, 1, 1]}[model])

    def make(Ks):
        u = torch.from_numpy(nominal_controls(model, T))

        def step():
            noise = torch.randn(Ks, T, nu) * sig              # standard_normal_noise.py:24
            if model == "wb":
                return tp.wb_step(noise, u, st[:12], st[12:19], st[19:26])
            if model == "arm":
                return tp.arm_step(noise, u, st[:7], st[7:14], st[14:21])
            if model == "drone":
                return tp.drone_step(noise, u, st[:3], st[3:6])
            return tp.quad_step(noise, u, st)
        return step
    probe = min(K, 512)
    step = make(probe)
    step()
    t0 = time.perf_counter()
    step()
    per_sample = (time.perf_counter() - t0) / probe
    Ks = int(min(K, max(probe, budget_s / 2 / max(per_sample, 1e-9))))
    step = make(Ks)
    times = []
    for _ in range(2):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return {"value": Ks * T / med, "unit": "rollout-steps/s", "cores": cores, "kind": "port",
            "sample": f"{Ks} of {K} samples x T={T}, 2 steps, oracle/torch_port.py (eager PyTorch CPU, replays the reference's "
                      f"aten op sequence, torch.randn noise), torch threads = {cores}", "ms_per_step_sample": med * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    spec = MODELS[args.model]
    K = args.samples or spec["K"]
    T = args.horizon or spec["T"]
    total = args.steps + args.warmup
    probe_K = min(K, 1024)
    step, orc = oracle_step_fn(args.model, probe_K, T)
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm must still use every usable host core
    threads = orc.set_threads(len(os.sched_getaffinity(0)))
    step()
    t0 = time.perf_counter()
    step()
    rate = probe_K * T / max(time.perf_counter() - t0, 1e-6)
    Ks = int(min(K, max(256, rate * 120.0 / (T * total))))          # whole run <= ~2 minutes
    step, _ = oracle_step_fn(args.model, Ks, T)
    for _ in range(args.warmup):
        step()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    tot = float(np.sum(times))
    value = Ks * T * args.steps / tot
    sample = (f"{Ks} of {K} samples x T={T} per step, oracle/mppi_oracle.c (CPU restatement of the reference step, "
              f"C + OpenMP, {threads} threads); the reference itself is Python and cannot travel to this box")
    line = {"impl": "reference", "metric": "rollout_steps_per_s", "value": value, "unit": "rollout-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{spec['desc']}, K={K}, T={T}", "noise": "philox", "sample_K": Ks},
            "cpu_baseline": {"value": value, "unit": "rollout-steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "latency_ms": {"p50": float(np.percentile(times, 50) * 1e3), "p99": float(np.percentile(times, 99) * 1e3)},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._thread = threading.Thread(target=self._run, args=(period_s,), daemon=True)
        except Exception as e:  # pragma: no cover
            self._nv, self._err = None, repr(e)

    def _run(self, period):
        nv = self._nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                get = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = get(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(period)

    def __enter__(self):
        if self._thread:
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ native arm
def make_controller(model, K, T, device, k_offset=0, seed=0):
    from quadrotor_manipulator_mppi_b200.mppi_solver import drone_mppi, mppi, quad_mppi, wholebody_mppi
    if model == "wb":
        return wholebody_mppi.MPPI(n_samples=K, n_horizon=T, seed=seed, device=device, k_offset=k_offset)
    if model == "arm":
        return mppi.MPPI(n_samples=K, n_horizon=T, seed=seed, device=device, verbose=False)
    if model == "drone":
        return drone_mppi.MPPI(n_samples=K, n_timestep=T, seed=seed, device=device)
    return quad_mppi.MPPI(n_samples=K, n_timestep=T, seed=seed, device=device)


def sensor_message(model, st):
    """The arguments the ROS node would pass (kinova.py:116 `update_joint(q, v)`, drone.py:164 `set_state(x, v)`)."""
    if model == "wb":
        return (st[:3], st[3:6], st[6:9], st[9:12], st[12:19], st[19:26])
    if model == "arm":
        return (np.concatenate([st[14:21], st[:7]]).astype(np.float64), np.concatenate([np.zeros(6), st[7:14]]).astype(np.float64))
    if model == "drone":
        return (st[:3], st[3:6])
    return (st[:3], st[3:6], st[6:9], st[9:12])


def feed_state(ctrl, model, msg):
    if model == "arm":
        ctrl.update_joint(*msg)
    else:
        ctrl.set_state(*msg)


def run_native(args):
    import torch
    import torch.distributed as dist
    from quadrotor_manipulator_mppi_b200 import _native
    from quadrotor_manipulator_mppi_b200.core import NativeSolver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the native path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    n_gpus = world
    spec = MODELS[args.model]
    K = args.samples or spec["K"]
    T = args.horizon or spec["T"]
    nu = spec["nu"]
    assert K % world == 0, "K must divide across ranks"
    K_loc = K // world
    model_id = {"wb": _native.MODEL_WB11, "arm": _native.MODEL_ARM7, "drone": _native.MODEL_DRONE3, "quad": _native.MODEL_QUAD4}[args.model]
    qp = None
    if args.model == "wb":
        qp = (14.7 + 5.5, 1 / 1.57, 1 / 3.93, 1 / 2.59, 0.0, -9.81)
    solver = NativeSolver(model_id, n_samples=K_loc, n_horizon=T, seed=0, device=device, k_offset=rank * K_loc, quad_params=qp)
    st = synthetic_state(args.model)
    solver.set_state(st)
    solver.u_prev = torch.from_numpy(nominal_controls(args.model, T))
    from quadrotor_manipulator_mppi_b200.sharded import ShardedStepper
    stepper = ShardedStepper(solver, exchange=args.exchange)
    noise = None
    if args.noise == "injected":
        noise = solver.generate_noise(0)        # resident in HBM before the timed region

    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)   # > 126 MB L2
    stream = torch.cuda.current_stream(device)
    rng = np.random.default_rng(1234)

    def one_step():
        if world == 1:
            solver.step_async(noise)
            return 2 if noise is None else 3        # rollout + weighting(+finalize) [+ weights kernel]
        stepper.step_async(noise)
        if stepper.exchange == "p2p":
            return 2 if noise is None else 3
        return 3 if noise is None else 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(max(args.warmup, 3)):
        if flush is not None:
            flush.zero_()
        one_step()
    barrier()

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    with ClockSampler(local_rank) as clocks:
        barrier()
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            # state/goal jitter for latency realism; by-value kernel parameter, no H2D copy
            jit = st.copy()
            jit[:3] += rng.uniform(-0.05, 0.05, 3).astype(np.float32)
            solver.set_state(jit)
            if flush is not None:
                flush.zero_()
            ev[i][0].record(stream)
            launches += one_step()
            ev[i][1].record(stream)
        barrier()
        wall = time.perf_counter() - t_wall0
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)       # per-step max over ranks
    step_ms = step_ms.cpu().numpy()
    total_ms = float(step_ms.sum())
    value = K * T * args.steps / (total_ms * 1e-3)

    # ---- collectives timed on their own (latency-bound, reported separately)
    coll = None
    if world > 1:
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        tm, ts = [], []
        for _ in range(50):
            e0.record(stream)
            dist.all_reduce(solver.rho_enc, op=dist.ReduceOp.MIN)
            e1.record(stream)
            dist.all_reduce(solver.wsum, op=dist.ReduceOp.SUM)
            e2.record(stream)
            torch.cuda.synchronize(device)
            tm.append(e0.elapsed_time(e1)); ts.append(e1.elapsed_time(e2))
        solver.rho_enc.fill_(0x7fffffff)
        coll = {"allreduce_min_ms_p50": float(np.percentile(tm[5:], 50)), "allreduce_sum_ms_p50": float(np.percentile(ts[5:], 50)),
                "allreduce_sum_floats": int(solver.wsum.numel())}

    # ---- roofline of the dominant kernel (fused rollout+cost), timed alone with events on its stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kt, wt = [], []
    for i in range(12):
        if flush is not None:
            flush.zero_()
        e0.record(stream)
        solver.rollout(noise)
        e1.record(stream)
        torch.cuda.synchronize(device)
        kt.append(e0.elapsed_time(e1))
        e0.record(stream)
        solver.weight(noise)
        e1.record(stream)
        torch.cuda.synchronize(device)
        wt.append(e0.elapsed_time(e1))
        solver.finalize()
    k_ms, w_ms = float(np.mean(kt[2:])), float(np.mean(wt[2:]))
    fp32_peak = _native.measure_fp32_peak(local_rank)
    alg_flops = _native.algorithmic_flops(model_id) * K_loc * T
    achieved = alg_flops / (k_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{args.model}_{args.noise}_K{K_loc}_T{T}")
        except Exception:
            traffic = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roofline = {"kernel": "rollout_cost_kernel", "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": achieved / fp32_peak if fp32_peak else None, "traffic": traffic,
                "peak_source": "FFMA probe measured in this run (mppi_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 entry; "
                               "nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s",
                "algorithmic_flop_per_rollout_step": _native.algorithmic_flops(model_id),
                "kernel_ms": k_ms, "share_of_step": k_ms / (k_ms + w_ms),
                "weighting_kernel_ms": w_ms}
    # ---- HBM-bound weighting pass (re-read of a materialised [T][K][nu] noise tensor), timed alone.
    # In Philox mode the step has no HBM-bound kernel, so the injected-noise weighting pass is measured
    # here on the side (same K, T) to report the second roofline the north star names.
    hbm_noise = noise
    if hbm_noise is None and world == 1 and K_loc * T * nu * 4 <= 4 * 1024 ** 3:
        hbm_noise = solver.generate_noise(0)
    if hbm_noise is not None:
        wt2 = []
        for i in range(8):
            if flush is not None:
                flush.zero_()
            solver.rollout(hbm_noise)
            e0.record(stream)
            solver.weight(hbm_noise)
            e1.record(stream)
            torch.cuda.synchronize(device)
            wt2.append(e0.elapsed_time(e1))
            solver.finalize()
        w2_ms = float(np.mean(wt2[2:]))
        gbs = K_loc * T * nu * 4 / (w2_ms * 1e-3) / 1e9
        tkey = f"{args.model}_weighting_K{K_loc}_T{T}"
        wtraffic = None
        if os.path.exists(tpath):
            try:
                wtraffic = json.load(open(tpath)).get(tkey)
            except Exception:
                wtraffic = None
        roofline["weighting_hbm"] = {"kernel": "weights_kernel + weighted_noise_kernel", "bound": "hbm", "achieved": gbs,
                                     "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s", "frac": gbs / peaks.get("hbm_gbs", 6650.0),
                                     "traffic": wtraffic, "kernel_ms": w2_ms,
                                     "algorithmic_bytes_per_rollout_step": nu * 4,
                                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"}
        if noise is None:
            del hbm_noise
            torch.cuda.empty_cache()

    # ---- end to end through the public controller class: host state in, host controls out, every step
    e2e = None
    if world == 1:
        ctrl = make_controller(args.model, K, T, device)
        feed_state(ctrl, args.model, sensor_message(args.model, st))
        n_e2e = max(10, min(args.steps, 100))
        for _ in range(3):
            ctrl.compute_control_input()
        jits = []
        for i in range(n_e2e):                              # synthetic sensor messages, prepared outside the timed region
            jit = st.copy()
            jit[:3] += rng.uniform(-0.05, 0.05, 3).astype(np.float32)
            jits.append(sensor_message(args.model, jit))
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for i in range(n_e2e):
            feed_state(ctrl, args.model, jits[i])           # host state -> kernel parameters
            res = ctrl.compute_control_input()              # host numpy / synchronised outputs
            if args.model in ("drone", "quad"):
                res[0].cpu()                                # drone.py:240 reads xdes on the host
        torch.cuda.synchronize(device)
        e2e_s = time.perf_counter() - t0
        e2e = {"value": K * T * n_e2e / e2e_s, "unit": "rollout-steps/s",
               "h2d_bytes_per_step": int(st.size * 4), "d2h_bytes_per_step": int(_native.MPPI_OUT_FLOATS * 4),
               "ms_per_step": e2e_s / n_e2e * 1e3, "steps": n_e2e,
               "note": "public class compute_control_input(): state passes as a by-value kernel parameter block, u_prev stays device-resident (warm start), the last block of the step stores the out vector into mapped pinned host memory and the call spins on its sequence word"}
    else:
        # multi-rank e2e: the sharded step plus a D2H of the out vector on every rank, wall clock max over ranks
        n_e2e = max(10, min(args.steps, 100))
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            stepper.step(noise, state=st)                  # host state in, host out vector back, on every rank
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_s = float(e2e_s.item())
        e2e = {"value": K * T * n_e2e / e2e_s, "unit": "rollout-steps/s", "h2d_bytes_per_step": int(st.size * 4),
               "d2h_bytes_per_step": int(_native.MPPI_OUT_FLOATS * 4), "ms_per_step": e2e_s / n_e2e * 1e3, "steps": n_e2e}

    cpu = cpu_torch = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.model, K, T, budget_s=12.0, n_steps=2)
        cpu_torch = cpu_baseline_torch(args.model, K, T)

    if rank == 0:
        line = {"metric": "rollout_steps_per_s", "value": value, "unit": "rollout-steps/s", "n_gpus": n_gpus,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{spec['desc']}, K={K}, T={T}", "noise": args.noise,
                           "K_per_gpu": K_loc, "parallelism": f"k-shard x{world}" if world > 1 else "single GPU",
                           "exchange": (stepper.exchange if world > 1 else None),
                           "exchange_fallback": getattr(stepper, "fallback_reason", None),
                           "l2": "no flush" if flush is None else "L2 flushed between steps (256 MiB memset) outside the per-step CUDA events",
                           "timing": "CUDA events around every step on the launch stream, summed, max over ranks"},
                "latency_ms": {"p50": float(np.percentile(step_ms, 50)), "p99": float(np.percentile(step_ms, 99)),
                               "max": float(step_ms.max()), "host_wall_ms_per_step_incl_flush": wall / args.steps * 1e3},
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu, "cpu_baseline_torch": cpu_torch}
        if coll:
            line["collectives"] = coll
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
