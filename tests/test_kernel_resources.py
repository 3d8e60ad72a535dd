"""Guards on the compiled kernels' resource usage (cuobjdump, no GPU needed): a register count that creeps over an
occupancy step is a silent performance regression -- the streaming weighting kernel once lost a resident block
(0.130 -> 0.202 ms) when code inlined into its last block grew the kernel from 56 to 80 registers."""
import re
import shutil
import subprocess

import pytest


def _resources():
    from quadrotor_manipulator_mppi_b200 import build
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    r = subprocess.run([exe, "-res-usage", build.build()], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    out = {}
    for name, regs, stack in re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", r.stdout):
        out[name] = (int(regs), int(stack))
    return out


def _one(res, *needles):
    hits = [v for k, v in res.items() if all(n in k for n in needles)]
    assert len(hits) == 1, (needles, [k for k in res if needles[0] in k])
    return hits[0]


def test_register_budgets_of_the_hot_kernels():
    res = _resources()
    for rounds in (10, 7):                                      # Philox round count (MPPI_OPTION_PHILOX_ROUNDS)
        # fused rollout, Philox, baked FK: 4 blocks of 128 threads per SM need <= 128 registers; no stack (no spills, no local arrays)
        for model in (1, 3):                                    # ARM7, WB11
            regs, stack = _one(res, f"rollout_cost_kernelILi{model}ELi0ELb1ELb0ELi{rounds}E")
            assert regs <= 128 and stack == 0
            # the single-launch step inherits the rollout's 4 blocks / SM (its co-residency limit is 592 blocks = 75776 samples)
            assert _one(res, f"step_fused_kernelILi{model}ELb1ELb0ELi{rounds}E")[0] <= 128
        for model in (0, 2):                                    # DRONE3, QUAD4
            regs, stack = _one(res, f"rollout_cost_kernelILi{model}ELi0ELb0ELb0ELi{rounds}E")
            assert regs <= 128 and stack == 0
            assert _one(res, f"step_fused_kernelILi{model}ELb0ELb0ELi{rounds}E")[0] <= 128
        # Philox weighting pass: blocks of up to 1024 threads
        for model in range(4):
            assert _one(res, f"weight_philox_kernelILi{model}ELi{rounds}E")[0] <= 64
        # time-parallel warp-per-sample step: one 512-thread block (16 samples per tile, one published row) per SM
        assert _one(res, f"step_tp_kernelILi1ELi0ELb1ELi1ELi{rounds}E")[0] <= 128
    # streaming weighting pass: 3 blocks of 32*nu threads per SM
    assert _one(res, "weighted_noise_kernelILi3ELi4E")[0] <= 62          # nu = 11: 352 threads
    assert _one(res, "weighted_noise_kernelILi1ELi4E")[0] <= 97          # nu = 7: 224 threads


def test_operand_model_of_the_whole_body_hot_loop():
    """tools/sass_operand_model.py on the shipped library (cuobjdump -sass, no GPU): the whole-body rollout's horizon loop
    priced at max(issue slot, FMA-pipe cycles, register source words / 2) per instruction tracked the measured kernel time
    to +-1 % over four builds (667 / 634 / 622 / 578 cycles per warp-step against 323 / 308 / 300 / 282 us on B200).  A
    change that silently puts work back into the loop (a per-step range reduction is 45 cycles, the cumulative-sum
    integrators 44) shows up here before any GPU run."""
    import os
    import sys
    from quadrotor_manipulator_mppi_b200 import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import sass_operand_model
    if not (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")):
        pytest.skip("cuobjdump unavailable")
    lib = build.build()
    unit = os.path.join(os.path.dirname(lib), "csrc", "_obj", "unit_m3_p0.o")        # the same code, 20x less SASS to dump
    path = unit if os.path.exists(unit) and os.path.getmtime(unit) <= os.path.getmtime(lib) + 1 else lib
    for rounds, cap in ((7, 600), (10, 660)):
        m = sass_operand_model.model(path, f"rollout_cost_kernel<3, 0, true, false, {rounds}>")
        assert 300 < m["instructions"] < 480, m["instructions"]                       # the horizon loop was found, not a helper loop
        assert m["serial_cost_cycles"] <= cap, (rounds, m["serial_cost_cycles"])
        assert m["pipe"]["xu"] <= 8 * 52                                              # MUFU: 24 Box-Muller + 20 sin/cos + 6 pose cost / rigid body
