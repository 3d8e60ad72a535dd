#!/usr/bin/env python
"""A/B device timing of library variants built by tools/build_variant.py (one subprocess per library).

    python tools/ab_variants.py base nb2 pf [--cases wb:262144:64,wb:32768:64] [--reps 40]

Per variant and case: 5 warm-up steps, `reps` steps timed one by one with CUDA events (Philox noise, warm start),
median / min reported, plus a checksum of the per-sample costs of the first step (variants must agree bitwise).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(cases, reps):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from bench import nominal_controls, synthetic_state
    from quadrotor_manipulator_mppi_b200 import _native
    from quadrotor_manipulator_mppi_b200.core import NativeSolver
    ids = {"wb": _native.MODEL_WB11, "arm": _native.MODEL_ARM7, "drone": _native.MODEL_DRONE3, "quad": _native.MODEL_QUAD4}
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev)
    out = {}
    for model, K, T in cases:
        qp = (14.7 + 5.5, 1 / 1.57, 1 / 3.93, 1 / 2.59, 0.0, -9.81) if model == "wb" else None
        s = NativeSolver(ids[model], n_samples=K, n_horizon=T, seed=0, device=dev, quad_params=qp)
        s.set_state(synthetic_state(model))
        s.u_prev = torch.from_numpy(nominal_controls(model, T))
        s.step_async()
        torch.cuda.synchronize(dev)
        chk = float(s.costs.double().sum().item())
        for _ in range(5):
            s.step_async()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for e0, e1 in ev:
            e0.record(stream)
            s.step_async()
            e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = np.array([a.elapsed_time(b) for a, b in ev])
        out[f"{model}:{K}:{T}"] = {"p50_us": round(float(np.median(ms)) * 1e3, 2), "min_us": round(float(ms.min()) * 1e3, 2),
                                   "cost_sum": chk, "path": s.last_path}
        s.close()
    print("RESULT " + json.dumps(out))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    opts = dict(a[2:].split("=", 1) for a in sys.argv[1:] if a.startswith("--") and "=" in a)
    cases = [(m, int(k), int(t)) for m, k, t in (c.split(":") for c in opts.get("cases", "wb:262144:64,wb:32768:64").split(","))]
    reps = int(opts.get("reps", "40"))
    if os.environ.get("MPPI_AB_CHILD"):
        return child(cases, reps)
    table = {}
    for rnd in range(2):                      # two rounds, interleaved, so clock drift does not favour a variant
        for name in args:
            env = dict(os.environ, MPPI_AB_CHILD="1")
            if name != "shipped":
                env["MPPI_B200_LIB"] = os.path.join(ROOT, "variants", f"libmppi_b200_{name}.so")
            r = subprocess.run([sys.executable, os.path.abspath(__file__), *sys.argv[1:]], env=env, capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if not line:
                print(name, "FAILED", r.stdout[-400:], r.stderr[-800:])
                continue
            table.setdefault(name, []).append(json.loads(line[0][7:]))
    for name, runs in table.items():
        for case in runs[0]:
            print(f"{name:10s} {case:18s} p50 {[x[case]['p50_us'] for x in runs]} min {[x[case]['min_us'] for x in runs]} "
                  f"cost_sum {runs[0][case]['cost_sum']:.6e} {runs[0][case]['path']}")
    out = os.path.join(ROOT, "gpurun_out", "ab_variants.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(table, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
