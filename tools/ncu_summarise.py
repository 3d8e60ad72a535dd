#!/usr/bin/env python
"""Turn the ncu captures of tools/ncu_capture.sh into the committed summaries (run here, no GPU):

    python tools/ncu_summarise.py gpurun_out/r02f_r7 7 [gpurun_out/r02f_r10 10]

  profiles/r02/<tag>_launches.csv      every launch: kernel, grid, regs, duration, pipe utilisation, DRAM bytes
  profiles/r02/<tag>_full_summary.csv  the --set full capture's headline rows (if the report exists)
  profiles/ncu_metrics.json            what bench.py's roofline object reads: executed FLOP per rollout-step (from the
                                       SASS page: the FP32 op counters miss the packed FFMA2/FMUL2/FADD2), FMA-pipe / issue
                                       utilisation, Philox share of the FMA-pipe cycles, DRAM bytes per launch
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_flops  # noqa: E402
import sass_operand_model  # noqa: E402

# (kernel regex, grid) -> (bench key stem, rollout-steps per launch)
CASES = [
    (r"rollout_cost_kernel<\(?(?:int\))?3, \(?(?:int\))?0", 2048, "wb_philox_K262144_T64", 262144 * 64),
    (r"rollout_cost_kernel<\(?(?:int\))?3, \(?(?:int\))?0", 256, "wb_philox_K32768_T64", 32768 * 64),
    (r"rollout_cost_kernel<\(?(?:int\))?3, \(?(?:int\))?2", 2048, "wb_injected_K262144_T64", 262144 * 64),
    (r"rollout_cost_kernel<\(?(?:int\))?1, \(?(?:int\))?0", 8192, "arm_philox_K1048576_T32", 1048576 * 32),
    (r"rollout_cost_kernel<\(?(?:int\))?2, \(?(?:int\))?0", 512, "quad_philox_K65536_T100", 65536 * 100),
]


def load_metrics(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    out = {}
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        k = int(d["ID"])
        e = out.setdefault(k, {"kernel": d["Kernel Name"]})
        e[d["Metric Name"]] = d["Metric Value"].replace(",", "")
    return [out[k] for k in sorted(out)]


def f(m, name):
    try:
        return float(m.get(name, "nan"))
    except ValueError:
        return float("nan")


def refresh_static(rounds):
    """--static R: re-derive the figures that need no GPU (operand model, FP32 FLOP of the hot loop) from the library as
    built now, for kernels that changed after the last ncu capture; the ncu-measured fields keep their capture's values
    and `static_refresh` says so."""
    mpath = os.path.join(ROOT, "profiles", "ncu_metrics.json")
    mj = json.load(open(mpath))
    lib = os.path.join(ROOT, "quadrotor_manipulator_mppi_b200", "libmppi_b200.so")
    for stem, model_id, per_iter, nth in (("wb_philox_K262144_T64", 3, 1, 0), ("wb_philox_K32768_T64", 3, 1, 0), ("quad_philox_K65536_T100", 2, 4, 0),
                                          ("arm_philox_K1048576_T32", 1, 2, 1)):
        key = f"{stem}_r{rounds}"
        if key not in mj:
            continue
        om = sass_operand_model.model(lib, f"rollout_cost_kernel<{model_id}, 0, {'true' if model_id in (1, 3) else 'false'}, false, {rounds}>", nth)
        mj[key]["operand_model"] = {"serial_cost_cycles_per_warp_step": om["serial_cost_cycles"] / per_iter, "issue_slots": om["instructions"] / per_iter,
                                    "fma_pipe_cycles": om["pipe"]["fma"] / per_iter, "xu_pipe_cycles": om["pipe"]["xu"] / per_iter,
                                    "register_source_words": om["register_source_words"] / per_iter}
        if "K32768" not in stem:
            mj[key]["executed_flop_per_rollout_step"] = om["fp32_flop_per_thread"] / per_iter
            imad = sum(sass_operand_model.FMA_PIPE[k] * v for k, v in om["counts"].items() if k.startswith("IMAD") and k in sass_operand_model.FMA_PIPE)
            mj[key]["philox_pipe_share"] = imad / om["pipe"]["fma"]
        mj[key]["static_refresh"] = ("operand_model and executed_flop_per_rollout_step re-derived from the SASS of the final library "
                                     "(tools/sass_operand_model.py: hot loop only, no GPU); pipe utilisation, registers and durations are those of the capture named in `source`")
    json.dump(mj, open(mpath, "w"), indent=1, sort_keys=True)


def main():
    args = sys.argv[1:]
    if args and args[0] == "--static":
        return refresh_static(int(args[1]))
    outdir = os.path.join(ROOT, "profiles", "r02")
    os.makedirs(outdir, exist_ok=True)
    mpath = os.path.join(ROOT, "profiles", "ncu_metrics.json")
    metrics_json = json.load(open(mpath)) if os.path.exists(mpath) else {}
    for prefix, rounds in zip(args[0::2], args[1::2]):
        tag = os.path.basename(prefix)
        launches = load_metrics(prefix + "_metrics.csv")
        with open(os.path.join(outdir, f"{tag}_launches.csv"), "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(["kernel", "grid", "block", "regs", "duration_us", "fma_pipe_pct", "fmaheavy_pct", "alu_pct", "xu_inst_pct", "issue_pct",
                        "warps_active_pct", "dram_read_B", "dram_write_B", "warp_inst", "scalar_fp32_flop_counter"])
            for m in launches:
                if f(m, "gpu__time_duration.sum") < 3000:
                    continue
                scal = f(m, "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum") + f(m, "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum") \
                    + 2 * f(m, "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum")
                w.writerow([re.sub(r"\(mppi::.*", "", m["kernel"]).replace("void mppi::", ""), m.get("launch__grid_size"), m.get("launch__block_size"),
                            m.get("launch__registers_per_thread"), round(f(m, "gpu__time_duration.sum") / 1e3, 3),
                            m.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                            m.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active"),
                            m.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                            m.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                            m.get("sm__issue_active.avg.pct_of_peak_sustained_active"),
                            m.get("sm__warps_active.avg.pct_of_peak_sustained_active"), m.get("dram__bytes_read.sum"),
                            m.get("dram__bytes_write.sum"), m.get("smsp__inst_executed.sum"), int(scal)])
        # executed FLOPs from the SASS page of the full report (if captured for this round count)
        flops = {}
        rep = prefix + "_full.ncu-rep"
        if os.path.exists(rep):
            for blk in ncu_flops.kernels(rep, "rollout_cost_kernel"):
                a = ncu_flops.analyse(blk)
                flops.setdefault(re.sub(r"\(mppi::.*", "", blk["name"]), a)
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rr = list(csv.reader(raw.splitlines()))
            keep = ["ID", "Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__waves_per_multiprocessor",
                    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
                    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
                    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
                    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
                    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
                    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
            ix = [rr[0].index(k) for k in keep if k in rr[0]]
            with open(os.path.join(outdir, f"{tag}_full_summary.csv"), "w", newline="") as fh:
                w = csv.writer(fh)
                for r in [rr[0]] + rr[2:]:
                    w.writerow([r[i][:90] for i in ix])
        for rx, grid, stem, steps in CASES:
            hit = [m for m in launches if re.search(rx, m["kernel"]) and int(f(m, "launch__grid_size")) == grid]
            if not hit:
                continue
            m = hit[0]
            entry = {"source": f"profiles/r02/{tag}_launches.csv (ncu --metrics, --clock-control none)"
                               + (f" and the SASS page of {tag}_full.ncu-rep (tools/ncu_flops.py)" if os.path.exists(rep) else ""),
                     "kernel_us_under_ncu": f(m, "gpu__time_duration.sum") / 1e3,
                     "fma_pipe_active_pct": f(m, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                     "issue_active_pct": f(m, "sm__issue_active.avg.pct_of_peak_sustained_active"),
                     "alu_pipe_active_pct": f(m, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                     "dram_bytes_per_launch": f(m, "dram__bytes_read.sum") + f(m, "dram__bytes_write.sum"),
                     "registers": int(f(m, "launch__registers_per_thread")),
                     "scalar_counter_flop_per_rollout_step": (f(m, "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum") + f(m, "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum")
                                                              + 2 * f(m, "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum")) / steps}
            key_kernel = [k for k in flops if re.search(rx, k)]
            if key_kernel:
                a = flops[key_kernel[0]]
                # the first matching launch in the report is the K=262144 / 1M / 65536 one for each template
                entry["executed_flop_per_rollout_step"] = a["executed_fp32_flop"] / steps if "K32768" not in stem else None
                entry["philox_pipe_share"] = a["philox_imad_pipe_share"]
                entry["warp_instructions_per_warp_step"] = a["warp_instructions"] / (steps / 32) if "K32768" not in stem else None
                entry["fma_pipe_cycles_model_per_warp_step"] = a["fma_pipe_cycles_model"] / (steps / 32) if "K32768" not in stem else None
            # static register-operand model of the hot loop of the shipped library (tools/sass_operand_model.py):
            # horizon steps per loop iteration = the noise batch (drone / quad 4) x the unroll factor (arm 2)
            mm = re.search(r"<\\\(\?\(\?:int\\\)\)\?(\d), \\\(\?\(\?:int\\\)\)\?(\d)", rx)
            try:
                model_id, noise_id = int(mm.group(1)), int(mm.group(2))
                if noise_id != 0:
                    raise LookupError("modelled for the Philox kernels only (the TMA-fed loop nests its mbarrier waits)")
                pat = f"rollout_cost_kernel<{model_id}, {noise_id}, {'true' if model_id in (1, 3) else 'false'}, false, " + (f"{rounds}>" if noise_id == 0 else "")
                om = sass_operand_model.model(os.path.join(ROOT, "quadrotor_manipulator_mppi_b200", "libmppi_b200.so"), pat,
                                              1 if (model_id == 1 and noise_id == 0) else 0)
                per_iter = {(1, 0): 2, (2, 0): 4}.get((model_id, noise_id), 1)
                entry["operand_model"] = {"serial_cost_cycles_per_warp_step": om["serial_cost_cycles"] / per_iter,
                                          "issue_slots": om["instructions"] / per_iter, "fma_pipe_cycles": om["pipe"]["fma"] / per_iter,
                                          "xu_pipe_cycles": om["pipe"]["xu"] / per_iter,
                                          "register_source_words": om["register_source_words"] / per_iter}
            except Exception as e:      # noqa: BLE001 -- the summary is still useful without the static model
                entry["operand_model"] = {"error": str(e)}
            if stem.startswith("wb_injected"):
                wn = [x for x in launches if "weighted_noise_kernel" in x["kernel"]]
                if wn:
                    entry["weighting_dram_bytes_per_launch"] = f(wn[0], "dram__bytes_read.sum") + f(wn[0], "dram__bytes_write.sum")
            metrics_json[f"{stem}_r{rounds}"] = entry
        # the Philox bench configuration also reports the HBM-bound weighting traffic (measured on the injected run)
        inj = metrics_json.get(f"wb_injected_K262144_T64_r{rounds}", {})
        if f"wb_philox_K262144_T64_r{rounds}" in metrics_json and "weighting_dram_bytes_per_launch" in inj:
            metrics_json[f"wb_philox_K262144_T64_r{rounds}"]["weighting_dram_bytes_per_launch"] = inj["weighting_dram_bytes_per_launch"]
    json.dump(metrics_json, open(mpath, "w"), indent=1, sort_keys=True)
    print(json.dumps(metrics_json, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
