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
probe run in the same process; next to the counted FLOP figure it carries the executed one, the
ncu pipe utilisation and the register-operand model of the hot loop (profiles/ncu_metrics.json,
tools/sass_operand_model.py) with the time that model predicts at the sampled SM clock.
`cpu_baseline` = the CPU oracle port (oracle/, C + OpenMP) on a
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
PHILOX_ROUNDS = 7                                  # the library default (MPPI_OPTION_PHILOX_ROUNDS); the CPU legs generate the same noise


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
    ap.add_argument("--lam", type=float, default=None, help="soft-min temperature (default: the reference's 0.1)")
    ap.add_argument("--philox-rounds", type=int, default=None, choices=[7, 10], help="Philox4x32 round count (library default if omitted)")
    ap.add_argument("--fused", type=int, default=-1, choices=[-1, 0, 1], help="single-launch step: -1 library default, 0 off, 1 on")
    ap.add_argument("--time-parallel", type=int, default=-2, choices=[-2, -1, 0, 1], help="warp-per-sample kernel: -2 library default, -1 auto, 0 off, 1 on")
    ap.add_argument("--latency-steps", type=int, default=1000, help="timed steps of the separate latency loop (0 = skip)")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-weights (ESS >= 1e3) timing")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity self-check")
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
        noise = orc.philox_noise(K, T, nu, sigma, seed=0, step=counter[0], rounds=PHILOX_ROUNDS)
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
            "config": {"workload": f"{spec['desc']}, K={K}, T={T}", "noise": "philox", "philox_rounds": PHILOX_ROUNDS, "sample_K": Ks,
                       "l2": "CPU arm: host caches, no flush", "timing": "time.perf_counter around every step, rank 0 only"},
            "cpu_baseline": {"value": value, "unit": "rollout-steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "latency_ms": {"p50": float(np.percentile(times, 50) * 1e3), "p99": float(np.percentile(times, 99) * 1e3)},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.004):
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
def make_controller(model, K, T, device, k_offset=0, seed=0, lam=None, opts=None):
    from quadrotor_manipulator_mppi_b200.mppi_solver import drone_mppi, mppi, quad_mppi, wholebody_mppi
    kw = dict(opts or {})
    if lam is not None:
        kw["lam"] = lam
    if model == "wb":
        kw.pop("time_parallel", None)
        return wholebody_mppi.MPPI(n_samples=K, n_horizon=T, seed=seed, device=device, k_offset=k_offset, **kw)
    if model == "arm":
        return mppi.MPPI(n_samples=K, n_horizon=T, seed=seed, device=device, verbose=False, **kw)
    if model == "drone":
        return drone_mppi.MPPI(n_samples=K, n_timestep=T, seed=seed, device=device, **kw)
    kw.pop("time_parallel", None)
    return quad_mppi.MPPI(n_samples=K, n_timestep=T, seed=seed, device=device, **kw)


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


def dense_lambda(costs, target_ess: float = 2000.0) -> float:
    """A lambda at which the soft-min keeps ESS ~ target on these costs (bisection; used for the dense-weights timing)."""
    import torch
    S = costs.double()
    S = S - S.min()
    lo, hi = 1e-4, 1e10
    for _ in range(70):
        mid = (lo * hi) ** 0.5
        w = torch.exp(-S / mid)
        if (w.sum() ** 2 / (w * w).sum()).item() < target_ess:
            lo = mid
        else:
            hi = mid
    return hi


def _operand_model(om, K_loc, T, k_ms, sm_mhz, sms=148):
    """Predicted kernel time from the static register-operand model of the hot loop (tools/sass_operand_model.py, committed
    in profiles/ncu_metrics.json): every warp instruction costs max(1 issue slot, its FMA-pipe cycles, its register source
    words / 2) on its scheduler; warps of 32 samples are spread over 4 schedulers per SM."""
    if not om or "serial_cost_cycles_per_warp_step" not in om or not sm_mhz:
        return None
    warp_steps_per_scheduler = (K_loc / 32.0) * T / (4 * sms)
    pred_ms = om["serial_cost_cycles_per_warp_step"] * warp_steps_per_scheduler / (sm_mhz * 1e3)
    bound_ms = max(om["issue_slots"], om["fma_pipe_cycles"], om["xu_pipe_cycles"], om["register_source_words"] / 2.0) \
        * warp_steps_per_scheduler / (sm_mhz * 1e3)
    return {**om, "sm_mhz": sm_mhz, "predicted_ms_serial": pred_ms, "measured_over_predicted": k_ms / pred_ms,
            "perfect_overlap_bound_ms": bound_ms, "frac_of_overlap_bound": bound_ms / k_ms,
            "note": "serial = no overlap between instructions of one scheduler; the bound = the busiest of issue, FMA pipe, XU pipe and "
                    "operand bandwidth with perfect overlap"}


def timed_steps(step_fn, n, stream, device, flush, before=None):
    """n steps, each bracketed by CUDA events on the launch stream (L2 flushed outside the events).  Returns ms[n]."""
    import torch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i in range(n):
        if before is not None:
            before(i)
        if flush is not None:
            flush.zero_()
        ev[i][0].record(stream)
        step_fn()
        ev[i][1].record(stream)
    torch.cuda.synchronize(device)
    return np.array([a.elapsed_time(b) for a, b in ev])


def oracle_parity(model, solver_factory, K, T, nu, rank_costs_fn=None, budget_s=20.0):
    """One GPU step (seed 0, step 0, nominal controls, synthetic state) against the CPU oracle on the SAME Philox noise
    (oracle/mppi_oracle.c regenerates it from the same counters): per-sample costs on all K samples, and the updated
    controls stage-isolated (the oracle's weighting on the GPU's costs) -- SURVEY 8(c) protocol.  Checker only."""
    import torch
    from oracle import oracle as orc
    orc.build()
    orc.set_threads(len(os.sched_getaffinity(0)))
    sigma = {"wb": [30 * 20.2, 1, 1, 1] + [0.1] * 7, "arm": [0.1] * 7, "drone": [30.0] * 3, "quad": [30 * 14.7, 1, 1, 1]}[model]
    st = synthetic_state(model)
    u = nominal_controls(model, T)
    s = solver_factory()
    s.set_state(st)
    s.u_prev = torch.from_numpy(u)
    rounds = s.get_option(1)
    out = s.step(None, step_counter=0).copy()
    S_gpu = s.costs.cpu().numpy()
    u_gpu = s.u_prev.cpu().numpy()
    t0 = time.perf_counter()
    noise = orc.philox_noise(K, T, nu, sigma, seed=0, step=0, rounds=rounds)
    if model == "wb":
        S_ref = orc.wb_costs(noise, u, st[:12], st[12:19], st[19:26])
    elif model == "arm":
        S_ref = orc.arm_costs(noise, u, st[:7], st[7:14], st[14:21])
    elif model == "drone":
        S_ref = orc.drone_costs(noise, u, st[:3], st[3:6])
    else:
        S_ref = orc.quad_costs(noise, u, st)
    lam = float(s.cfg.lambda_)
    iso = orc._update(S_gpu, noise, u, lam, int(s.cfg.savgol_window))
    cpu_s = time.perf_counter() - t0
    rel = np.abs(S_gpu.astype(np.float64) - S_ref) / np.maximum(np.abs(S_ref), 1e-30)
    S_err = float(rel.max())
    u_err = float(np.abs(u_gpu.astype(np.float64) - iso["u_new"]).max() / max(np.abs(iso["u_new"]).max(), 1e-30))
    res = {"against": "oracle/mppi_oracle.c on the identical Philox noise (same seed / step / global sample index)",
           "samples_checked": int(K), "S_rel_err_max": S_err, "S_outliers_gt_1e-5": int((rel > 1e-5).sum()),
           "u_new_rel_err_stage_isolated": u_err, "rho_equal_min_S": bool(out[53] == S_gpu.min()),
           "tolerance": {"S": 1e-4, "u_new_stage_isolated": 1e-5}, "oracle_seconds": cpu_s,
           "ok": bool(S_err < 1e-4 and u_err < 1e-5 and out[53] == S_gpu.min())}
    s.close()
    return res


def run_native(args):
    import torch
    import torch.distributed as dist
    from quadrotor_manipulator_mppi_b200 import _native
    from quadrotor_manipulator_mppi_b200.core import NativeSolver
    from quadrotor_manipulator_mppi_b200.sharded import ShardedStepper, shard_range

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
    k_off, K_loc = shard_range(K, world, rank)            # contiguous shards; the first K % world ranks hold one more sample
    model_id = {"wb": _native.MODEL_WB11, "arm": _native.MODEL_ARM7, "drone": _native.MODEL_DRONE3, "quad": _native.MODEL_QUAD4}[args.model]
    qp = (14.7 + 5.5, 1 / 1.57, 1 / 3.93, 1 / 2.59, 0.0, -9.81) if args.model == "wb" else None
    opts = dict(philox_rounds=args.philox_rounds, fused=None if args.fused < 0 else args.fused,
                time_parallel=None if args.time_parallel < -1 else args.time_parallel)

    def make_solver(k_local=K_loc, k_offset=k_off, lam=args.lam):
        return NativeSolver(model_id, n_samples=k_local, n_horizon=T, seed=0, device=device, k_offset=k_offset, quad_params=qp,
                            lam=lam, **opts)

    solver = make_solver()
    st = synthetic_state(args.model)
    solver.set_state(st)
    u_nom0 = torch.from_numpy(nominal_controls(args.model, T))
    solver.u_prev = u_nom0
    stepper = ShardedStepper(solver, exchange=args.exchange)
    noise = None
    if args.noise == "injected":
        noise = solver.generate_noise(0)        # resident in HBM before the timed region

    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)   # > 126 MB L2
    stream = torch.cuda.current_stream(device)
    rng = np.random.default_rng(1234)

    def launches_per_step(s):
        path = s.last_path
        if path in ("fused", "time_parallel"):
            return 1
        base = 2 if noise is None else 3                    # rollout + weighting(+finalize) [+ weights kernel]
        return base if (world == 1 or stepper.exchange == "p2p") else base + 1      # + finalize kernel on the allreduce path

    def one_step():
        if world == 1:
            solver.step_async(noise)
        else:
            stepper.step_async(noise)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def jitter(_i=None):
        # state/goal jitter for latency realism; by-value kernel parameter, no H2D copy.  Every rank draws
        # from the same seeded generator, so the replicas see the same "sensor message" without any communication.
        jit = st.copy()
        jit[:3] += rng.uniform(-0.05, 0.05, 3).astype(np.float32)
        solver.set_state(jit)

    for _ in range(max(args.warmup, 3)):
        if flush is not None:
            flush.zero_()
        one_step()
    barrier()

    with ClockSampler(local_rank) as clocks:
        barrier()
        t_wall0 = time.perf_counter()
        step_ms = timed_steps(one_step, args.steps, stream, device, flush, before=jitter)
        barrier()
        wall = time.perf_counter() - t_wall0
    launches = launches_per_step(solver) * args.steps
    step_ms_t = torch.tensor(step_ms, dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(step_ms_t, op=dist.ReduceOp.MAX)       # per-step max over ranks
    step_ms = step_ms_t.cpu().numpy()
    total_ms = float(step_ms.sum())
    value = K * T * args.steps / (total_ms * 1e-3)
    out_last = solver._outs[(solver._out_i - 1) & 3].cpu().numpy()
    S_dev = solver.costs
    weights_info = {"lambda": float(solver.cfg.lambda_), "ess": float(out_last[_native.MPPI_OUT_ESS]),
                    "nonzero_weights_this_shard": int((torch.exp(-(S_dev - S_dev.min()) / float(solver.cfg.lambda_)) != 0).sum().item()),
                    "path": solver.last_path}

    # ---- latency distribution on its own: >= 1000 timed steps after >= 50 warm-ups, independent of --steps
    n_lat = max(args.latency_steps, 0)
    latency = None
    if n_lat > 0:
        for _ in range(50):
            one_step()
        barrier()
        lat = timed_steps(one_step, n_lat, stream, device, flush, before=jitter)
        lat_t = torch.tensor(lat, dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(lat_t, op=dist.ReduceOp.MAX)
        lat = lat_t.cpu().numpy()
        if world > 1:
            # every one of the n_lat + warm-up steps went through the exchange and u_prev carries over from step to step:
            # a single stale or torn row in any step would leave the replicas' control sequences different for good
            u_now = solver.u_prev.clone()
            gu_ = [torch.empty_like(u_now) for _ in range(world)]
            dist.all_gather(gu_, u_now)
            replicas_identical_after_loop = bool(all(torch.equal(gu_[0], g) for g in gu_)) and bool(torch.isfinite(u_now).all())
        latency = {"n": int(n_lat), "warmup": 50, "p50": float(np.percentile(lat, 50)), "p90": float(np.percentile(lat, 90)),
                   "p99": float(np.percentile(lat, 99)), "max": float(lat.max()), "mean": float(lat.mean()),
                   "host_wall_ms_per_step_incl_flush": wall / args.steps * 1e3}

    if latency is not None and world > 1:
        latency["replicas_bit_identical_after_these_steps"] = replicas_identical_after_loop

    # ---- the same K x T with weights that do NOT collapse (ESS >= 1e3): the weighting pass regenerates all of the noise
    dense = None
    if noise is None and not args.no_dense:
        lam_d = dense_lambda(S_dev, min(2000.0, max(2.0, K_loc / 4.0)))
        if world > 1:
            lt = torch.tensor([lam_d], dtype=torch.float64, device=device)
            dist.broadcast(lt, 0)
            lam_d = float(lt.item())
        solver.update_config(lambda_=lam_d)
        for _ in range(5):
            one_step()
        barrier()
        dms = timed_steps(one_step, max(20, min(args.steps, 100)), stream, device, flush)
        dms_t = torch.tensor(dms, dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(dms_t, op=dist.ReduceOp.MAX)
        dms = dms_t.cpu().numpy()
        od = solver._outs[(solver._out_i - 1) & 3].cpu().numpy()
        dense = {"lambda": lam_d, "ess": float(od[_native.MPPI_OUT_ESS]), "ms_per_step_dense_weights": float(dms.mean()),
                 "p50": float(np.percentile(dms, 50)), "p99": float(np.percentile(dms, 99)), "steps": int(len(dms)),
                 "note": "same K x T; lambda raised until ESS ~ 2e3, so (nearly) every sample keeps a non-zero weight"}
        solver.update_config(lambda_=args.lam if args.lam is not None else 0.1)
        solver.u_prev = u_nom0
        solver.set_state(st)

    # ---- where the time inside a step goes: in-kernel %globaltimer stamps at the phase boundaries (MPPI_OPTION_TRACE)
    phases = None
    try:
        solver.trace(True)
        acc = []
        for _ in range(30):
            barrier()
            one_step()
            torch.cuda.synchronize(device)
            acc.append(solver.trace_times())
        solver.trace(False)
        keys = [k for k in acc[-1] if k != "start"]
        med = {k: float(np.median([x[k] for x in acc if k in x])) for k in keys}
        if world > 1:
            for k in keys:
                t = torch.tensor([med[k]], dtype=torch.float64, device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                med[k] = float(t.item())
        phases = {"us_from_first_block_start": med,
                  "note": "median of 30 steps, max over ranks; 'exchanged' - 'reduced' is the peer-exchange wait (N > 1)"}
        if "exchanged" in med and "reduced" in med:
            phases["peer_exchange_us"] = med["exchanged"] - med["reduced"]
        if "end" in med and "rollout_done" in med:
            phases["tail_after_rollout_us"] = med["end"] - med["rollout_done"]
    except Exception as e:                                   # tracing is diagnostics: never fail the bench on it
        phases = {"error": repr(e)}

    # ---- the same steps with the other Philox round count, for the record (the noise definition is a handle option)
    other_rounds = None
    if noise is None:
        r_now = solver.get_option(_native.OPTION_PHILOX_ROUNDS)
        r_other = 10 if r_now == 7 else 7
        solver.set_option(_native.OPTION_PHILOX_ROUNDS, r_other)
        for _ in range(5):
            one_step()
        barrier()
        oms = timed_steps(one_step, max(20, min(args.steps, 100)), stream, device, flush)
        oms_t = torch.tensor(oms, dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(oms_t, op=dist.ReduceOp.MAX)
        other_rounds = {"philox_rounds": r_other, "ms_per_step": float(oms_t.mean().item()), "steps": int(len(oms))}
        solver.set_option(_native.OPTION_PHILOX_ROUNDS, r_now)
        solver.u_prev = u_nom0
        solver.set_state(st)

    # ---- collectives timed on their own (latency-bound, reported separately) + the contract path (allreduce-MIN / -SUM)
    coll = None
    exchange_nccl = None
    parity = None
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
        # the step through NCCL allreduce-MIN + allreduce-SUM (the exchange BASELINE.json names), same shard, same invocation
        solver_n = make_solver()
        solver_n.set_state(st)
        solver_n.u_prev = u_nom0
        stepper_n = ShardedStepper(solver_n, exchange="nccl")
        for _ in range(5):
            stepper_n.step_async(noise)
        barrier()
        nms = timed_steps(lambda: stepper_n.step_async(noise), max(20, min(args.steps, 100)), stream, device, flush)
        nms_t = torch.tensor(nms, dtype=torch.float64, device=device)
        dist.all_reduce(nms_t, op=dist.ReduceOp.MAX)
        nms = nms_t.cpu().numpy()
        exchange_nccl = {"ms_per_step": float(nms.mean()), "p50": float(np.percentile(nms, 50)), "p99": float(np.percentile(nms, 99)),
                         "steps": int(len(nms)), "value": K * T / (float(nms.mean()) * 1e-3),
                         "path": "rollout -> allreduce-MIN(int32) -> weighting -> allreduce-SUM([T*nu+2] f32) -> finalize"}

        # ---- parity self-check (outside every timed region): N ranks == 1 rank, both exchanges
        checks = {}
        if latency is not None:
            checks["replicas_bit_identical_after_latency_loop"] = replicas_identical_after_loop
        # (a) no peer-exchange timeout during the timed steps: blocking step raises on the sticky failure word
        try:
            o_blk = stepper.step(noise, state=st)
            checks["exchange_timeouts"] = 0 if float(o_blk[_native.MPPI_OUT_STEP]) >= 0 else 1
        except _native.MppiError as e:
            checks["exchange_timeouts"] = 1
            checks["exchange_error"] = str(e)
        # (b) one step from identical inputs through p2p, NCCL and (rank 0) a single full-K solver -- with the reference's
        # lambda (weights may collapse to one sample: then the update is that sample's noise whatever the exchange does)
        # AND with the dense-weights lambda (every shard contributes to every entry of the sums)
        SC = 1000003
        lam_ref = args.lam if args.lam is not None else 0.1
        lam_dense = dense["lambda"] if dense else 50.0
        ok_local = checks["exchange_timeouts"] == 0 and checks.get("replicas_bit_identical_after_latency_loop", True)
        for tag, lam_c in (("", lam_ref), ("dense_", lam_dense)):
            res = {}
            for name, stp, slv in (("used", stepper, solver), ("nccl", stepper_n, solver_n)):
                slv.update_config(lambda_=lam_c)
                slv.set_state(st)
                slv.u_prev = u_nom0
                slv.step_counter = SC
                o = stp.step(noise, state=st)
                res[name] = (slv.u_prev.clone(), torch.from_numpy(np.array(o, copy=True)).to(device), slv.costs.clone())
            u_used, o_used, S_used = res["used"]
            u_nccl, o_nccl, _ = res["nccl"]
            gu = [torch.empty_like(u_used) for _ in range(world)]
            go = [torch.empty_like(o_used) for _ in range(world)]
            dist.all_gather(gu, u_used)
            dist.all_gather(go, o_used)
            checks[tag + "ranks_bit_identical_u_new"] = bool(all(torch.equal(gu[0], g) for g in gu))
            checks[tag + "ranks_bit_identical_out"] = bool(all(torch.equal(go[0], g) for g in go))
            checks[tag + "out_step_nonnegative"] = bool(all(float(g[_native.MPPI_OUT_STEP]) >= 0 for g in go))
            checks[tag + "ess"] = float(go[0][_native.MPPI_OUT_ESS])
            un = u_nccl.double()
            checks[tag + "p2p_vs_nccl_u_new_rel"] = float(((u_used.double() - un).abs().max() / un.abs().max().clamp_min(1e-30)).item())
            # costs of every shard -> rank 0 (ragged shards: pad to the largest)
            kmax = (K + world - 1) // world
            pad = torch.full((kmax,), float("nan"), device=device)
            pad[:K_loc] = S_used
            gs = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(gs, pad)
            if rank == 0:
                full = make_solver(k_local=K, k_offset=0, lam=lam_c)
                full.set_state(st)
                full.u_prev = u_nom0
                full.step(None, step_counter=SC)
                S_all = torch.cat([gs[r][:shard_range(K, world, r)[1]] for r in range(world)])
                checks[tag + "sharded_costs_bitwise_equal_single_rank"] = bool(torch.equal(S_all, full.costs))
                uf = full.u_prev.double()
                checks[tag + "sharded_vs_single_rank_u_new_rel"] = float(((u_used.double() - uf).abs().max() / uf.abs().max().clamp_min(1e-30)).item())
                full.close()
            barrier()            # the other ranks must not enter the next exchange (bounded ~2 s wait) while rank 0 is still busy here
            ok_local = (ok_local and checks[tag + "ranks_bit_identical_u_new"] and checks[tag + "ranks_bit_identical_out"]
                        and checks[tag + "out_step_nonnegative"] and checks[tag + "p2p_vs_nccl_u_new_rel"] <= 1e-6
                        and checks.get(tag + "sharded_costs_bitwise_equal_single_rank", True)
                        and checks.get(tag + "sharded_vs_single_rank_u_new_rel", 0.0) <= 1e-5)
        for slv in (solver, solver_n):
            slv.update_config(lambda_=lam_ref)
        okt = torch.tensor([1 if ok_local else 0], dtype=torch.int32, device=device)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        checks["ok"] = bool(int(okt.item()) == 1)
        checks["exchange_checked"] = [stepper.exchange, "nccl"]
        checks["tolerance"] = {"p2p_vs_nccl": 1e-6, "sharded_vs_single_rank_u_new": 1e-5, "costs": "bitwise"}
        parity = checks
        solver.set_state(st)
        solver.u_prev = u_nom0

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
    alg = _native.algorithmic_flops(model_id)
    achieved = alg * K_loc * T / (k_ms * 1e-3) / 1e12
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_metrics.json"))).get(f"{args.model}_{args.noise}_K{K_loc}_T{T}_r{solver.get_option(1)}", {})
    except Exception:
        prof = {}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    exe = prof.get("executed_flop_per_rollout_step")
    clocks_mhz = clocks.summary().get("sm_mhz")         # median SM clock under load, sampled during the timed steps
    roofline = {"kernel": "rollout_cost_kernel", "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": achieved / fp32_peak if fp32_peak else None, "traffic": prof.get("dram_bytes_per_launch"),
                "peak_source": "FFMA probe measured in this run (mppi_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 entry; "
                               "nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s",
                "algorithmic_flop_per_rollout_step": alg,
                "algorithmic_flop_source": "oracle/flop_count.py: static count over the restated maths, general (dense) URDF constants, "
                                           "FMA = 2, transcendental evaluations excluded (tests/test_flop_count.py); the survey's "
                                           "pre-build estimate for this model was " + str({"wb": 1000, "arm": 840, "quad": 90, "drone": 40}[args.model]),
                "structural_flop_per_rollout_step": _native.structural_flops(model_id),
                "executed_flop_per_rollout_step": exe,
                "frac_executed": (exe * K_loc * T / (k_ms * 1e-3) / 1e12 / fp32_peak) if (exe and fp32_peak) else None,
                "fma_pipe_active_pct": prof.get("fma_pipe_active_pct"), "philox_pipe_share": prof.get("philox_pipe_share"),
                "issue_active_pct": prof.get("issue_active_pct"),
                "executed_source": prof.get("source", "no ncu capture committed for this configuration (profiles/ncu_metrics.json)"),
                "kernel_ms": k_ms, "share_of_step": k_ms / (k_ms + w_ms), "weighting_kernel_ms": w_ms,
                "operand_model": _operand_model(prof.get("operand_model"), K_loc, T, k_ms, clocks_mhz),
                "binding_limit": {"wb": "register operand bandwidth + FMA pipe (FP32 + Philox IMAD.WIDE): 2 register source words per lane per cycle "
                                        "(tools/probe_pipes2.cu), so a 3-register FFMA issues every 1.5 cycles and the 74 TFLOP/s FP32 peak needs "
                                        "constant / uniform operands",
                                  "arm": "register operand bandwidth + FMA pipe (FP32 + Philox IMAD.WIDE)",
                                  "quad": "dependent-issue latency of the per-step chain (Philox IMAD + MUFU Box-Muller + rigid-body step); FP32 is not the binding roofline for this model",
                                  "drone": "dependent-issue latency (Philox IMAD + MUFU); FP32 is not the binding roofline for this model"}[args.model]}
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
        roofline["weighting_hbm"] = {"kernel": "weights_kernel + weighted_noise_kernel", "bound": "hbm", "achieved": gbs,
                                     "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s", "frac": gbs / peaks.get("hbm_gbs", 6650.0),
                                     "traffic": prof.get("weighting_dram_bytes_per_launch"), "kernel_ms": w2_ms,
                                     "algorithmic_bytes_per_rollout_step": nu * 4,
                                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"}
        if noise is None:
            del hbm_noise
            torch.cuda.empty_cache()

    # ---- end to end through the public controller class: host state in, host controls out, every step
    e2e = None
    if world == 1:
        ctrl = make_controller(args.model, K, T, device, lam=args.lam, opts=opts)
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
        # multi-rank e2e: the sharded step plus the host-visible out vector on every rank, wall clock max over ranks
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
        if noise is None and not args.no_parity:
            parity = oracle_parity(args.model, lambda: make_solver(k_local=K, k_offset=0), K, T, nu)

    if rank == 0:
        line = {"metric": "rollout_steps_per_s", "value": value, "unit": "rollout-steps/s", "n_gpus": n_gpus,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{spec['desc']}, K={K}, T={T}", "noise": args.noise,
                           "philox_rounds": solver.get_option(_native.OPTION_PHILOX_ROUNDS),
                           "K_per_gpu": K_loc, "parallelism": f"k-shard x{world}" if world > 1 else "single GPU",
                           "step_path": weights_info["path"],
                           "exchange": (stepper.exchange if world > 1 else None),
                           "exchange_fallback": getattr(stepper, "fallback_reason", None),
                           "l2": "no flush" if flush is None else "L2 flushed between steps (256 MiB memset) outside the per-step CUDA events",
                           "timing": "CUDA events around every step on the launch stream, summed, max over ranks"},
                "latency_ms": latency if latency is not None else {"n": int(args.steps), "p50": float(np.percentile(step_ms, 50)),
                                                                   "p99": float(np.percentile(step_ms, 99)), "max": float(step_ms.max())},
                "weights": weights_info, "dense_weights": dense, "phases": phases, "other_philox_rounds": other_rounds,
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "parity_check": parity, "cpu_baseline": cpu, "cpu_baseline_torch": cpu_torch}
        if coll:
            line["collectives"] = coll
        if exchange_nccl:
            line["exchange_nccl"] = exchange_nccl
        print(json.dumps(line))
    failed = parity is not None and not parity.get("ok", True)
    if world > 1:
        dist.destroy_process_group()
    return 1 if failed else 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
