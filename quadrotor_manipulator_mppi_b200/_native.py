"""ctypes binding of libmppi_b200.so (include/mppi_b200.h).

There is no CPU fallback anywhere in this package: if the shared library is missing it
is built with nvcc; if that is impossible, or if a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

MPPI_MAX_NU = 12
MPPI_STATE_FLOATS = 32
MPPI_OUT_FLOATS = 64
MPPI_IPC_HANDLE_BYTES = 64
MPPI_OUT_BASE, MPPI_OUT_U0_NEW, MPPI_OUT_U0_OLD = 16, 28, 40
MPPI_OUT_REACH, MPPI_OUT_RHO, MPPI_OUT_ETA, MPPI_OUT_ESS, MPPI_OUT_STEP, MPPI_OUT_TORQUE = 52, 53, 54, 55, 56, 57
MODEL_DRONE3, MODEL_ARM7, MODEL_QUAD4, MODEL_WB11 = 0, 1, 2, 3
MODEL_NU = {MODEL_DRONE3: 3, MODEL_ARM7: 7, MODEL_QUAD4: 4, MODEL_WB11: 11}
MODEL_STATE = {MODEL_DRONE3: 6, MODEL_ARM7: 21, MODEL_QUAD4: 12, MODEL_WB11: 26}
ABI_VERSION = 4
COST_COVAR, COST_CENTERING, COST_JOINT_TRAJ, COST_ACTION, COST_JOINT_LIMIT = 1, 2, 4, 8, 16
OPT_TORQUE_LAW = 256            # shares the cost_flags word; ARM7 only
ARM_TWIST_FLOATS = 6            # optional tail of the ARM7 state: v_full[:6] (read by the torque law)

EXPORTS = [
    "mppi_abi_version", "mppi_last_error", "mppi_default_config", "mppi_create", "mppi_destroy",
    "mppi_update_config", "mppi_set_joint_traj", "mppi_set_chain", "mppi_set_arm_inertia", "mppi_set_target", "mppi_set_state", "mppi_step", "mppi_rollout", "mppi_weight",
    "mppi_finalize", "mppi_rho_ptr", "mppi_wsum_ptr", "mppi_wsum_count", "mppi_cost_ptr",
    "mppi_p2p_export", "mppi_p2p_bind", "mppi_step_p2p", "mppi_step_p2p_sync", "mppi_step_sync", "mppi_step_host", "mppi_generate_noise", "mppi_measure_fp32_peak",
    "mppi_algorithmic_flops_per_rollout_step", "mppi_set_option", "mppi_get_option", "mppi_get_kernel_times", "mppi_get_trace", "mppi_reserve_host_noise",
    "mppi_structural_flops_per_rollout_step",
]
OPTION_PHILOX_ROUNDS, OPTION_FUSED_STEP, OPTION_TIME_PARALLEL, OPTION_PROFILE, OPTION_NVTX, OPTION_LAST_PATH = 1, 2, 3, 4, 5, 6
OPTION_TRACE = 7
OPTION_HOST_YIELD = 8
TRACE_POINTS = ("start", "rollout_done", "min_known", "sums_added", "last_block", "reduced", "exchanged", "controls_updated", "end", "weight_start")
PATH_TWO_KERNELS, PATH_FUSED, PATH_TIMEPARALLEL = 1, 2, 3
PATH_NAMES = {0: "none", 1: "two_kernels", 2: "fused", 3: "time_parallel"}
ERR_PEER = 5


class MppiConfig(C.Structure):
    """Mirror of mppi_config_t."""
    _fields_ = [
        ("abi_version", C.c_int32), ("model", C.c_int32), ("n_samples", C.c_int32), ("n_horizon", C.c_int32),
        ("savgol_window", C.c_int32), ("savgol_polyorder", C.c_int32), ("device", C.c_int32), ("n_joints", C.c_int32),
        ("k_offset", C.c_int64), ("seed", C.c_uint64),
        ("dt", C.c_float), ("lambda_", C.c_float),
        ("sigma", C.c_float * MPPI_MAX_NU), ("cost_w", C.c_float * 8), ("quad_params", C.c_float * 6),
        ("target_pos", C.c_float * 3), ("target_quat", C.c_float * 4), ("drone_target", C.c_float * 3),
        ("torque_kp", C.c_float), ("torque_kd", C.c_float), ("reserved", C.c_float * 3),
        ("cost_flags", C.c_int32), ("gamma", C.c_float), ("covar_weight", C.c_float), ("alpha", C.c_float),
        ("action_weight", C.c_float), ("centering_weight", C.c_float), ("joint_traj_weight", C.c_float),
        ("limit_penalty", C.c_float), ("q_center", C.c_float * 7), ("q_lower", C.c_float * 7), ("q_upper", C.c_float * 7),
        ("reserved2", C.c_float),
    ]


_lib = None
_fp = C.POINTER(C.c_float)


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if needed) the shared library and declare its prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()       # no-op when the library matches the sources' content hash; raises if nvcc is unavailable
    lib = C.CDLL(path)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int32
    lib.mppi_abi_version.restype = i32
    lib.mppi_last_error.restype = C.c_char_p
    lib.mppi_last_error.argtypes = [vp]
    lib.mppi_default_config.argtypes = [i32, C.POINTER(MppiConfig)]
    lib.mppi_create.argtypes = [C.POINTER(MppiConfig), C.POINTER(vp)]
    lib.mppi_destroy.argtypes = [vp]
    lib.mppi_update_config.argtypes = [vp, C.POINTER(MppiConfig)]
    lib.mppi_set_joint_traj.argtypes = [vp, _fp]
    lib.mppi_set_chain.argtypes = [vp, i32, C.POINTER(i32), _fp, _fp, _fp]
    lib.mppi_set_target.argtypes = [vp, _fp, _fp, _fp]
    lib.mppi_set_arm_inertia.argtypes = [vp, _fp, _fp, _fp]
    lib.mppi_set_state.argtypes = [vp, _fp, i32]
    lib.mppi_step.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp]
    lib.mppi_rollout.argtypes = [vp, vp, vp, u64, vp, vp]
    lib.mppi_weight.argtypes = [vp, vp, u64, vp]
    lib.mppi_finalize.argtypes = [vp, vp, u64, vp, vp, vp]
    lib.mppi_rho_ptr.restype = vp
    lib.mppi_rho_ptr.argtypes = [vp]
    lib.mppi_wsum_ptr.restype = vp
    lib.mppi_wsum_ptr.argtypes = [vp]
    lib.mppi_wsum_count.restype = i32
    lib.mppi_wsum_count.argtypes = [vp]
    lib.mppi_cost_ptr.restype = vp
    lib.mppi_cost_ptr.argtypes = [vp]
    lib.mppi_p2p_export.argtypes = [vp, i32, vp]
    lib.mppi_p2p_bind.argtypes = [vp, i32, i32, vp]
    lib.mppi_step_p2p.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp]
    lib.mppi_step_sync.argtypes = [vp, _fp, i32, vp, vp, u64, vp, vp, vp]
    lib.mppi_step_p2p_sync.argtypes = [vp, _fp, i32, vp, vp, u64, vp, vp, vp]
    lib.mppi_step_host.argtypes = [vp, _fp, i32, _fp, _fp, u64, _fp, _fp]
    lib.mppi_generate_noise.argtypes = [vp, u64, vp, vp]
    lib.mppi_measure_fp32_peak.argtypes = [i32, _fp]
    lib.mppi_algorithmic_flops_per_rollout_step.restype = C.c_double
    lib.mppi_algorithmic_flops_per_rollout_step.argtypes = [i32]
    lib.mppi_structural_flops_per_rollout_step.restype = C.c_double
    lib.mppi_structural_flops_per_rollout_step.argtypes = [i32]
    lib.mppi_set_option.argtypes = [vp, i32, i32]
    lib.mppi_get_option.argtypes = [vp, i32, C.POINTER(i32)]
    lib.mppi_get_kernel_times.argtypes = [vp, _fp]
    lib.mppi_get_trace.argtypes = [vp, C.POINTER(C.c_uint64), i32]
    lib.mppi_reserve_host_noise.argtypes = [vp]
    for name in EXPORTS:
        if name not in ("mppi_abi_version", "mppi_last_error", "mppi_rho_ptr", "mppi_wsum_ptr", "mppi_wsum_count",
                        "mppi_cost_ptr", "mppi_algorithmic_flops_per_rollout_step", "mppi_structural_flops_per_rollout_step"):
            getattr(lib, name).restype = i32
    if lib.mppi_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libmppi_b200.so ABI {lib.mppi_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


class MppiError(RuntimeError):
    pass


_STATUS = {1: "invalid argument", 2: "wrong architecture (sm_100a only, no CPU fallback)", 3: "CUDA error",
           4: "unsupported", 5: "peer exchange failed"}


def check(rc: int, handle=None):
    if rc != 0:
        msg = load().mppi_last_error(handle)
        raise MppiError(f"libmppi_b200: {_STATUS.get(rc, rc)}: {msg.decode() if msg else ''}")


def default_config(model: int) -> MppiConfig:
    cfg = MppiConfig()
    check(load().mppi_default_config(model, C.byref(cfg)))
    return cfg


def fptr(np_array):
    """float* of a C-contiguous float32 numpy array (kept alive by the caller)."""
    return np_array.ctypes.data_as(_fp)


def measure_fp32_peak(device: int = 0) -> float:
    out = C.c_float()
    check(load().mppi_measure_fp32_peak(device, C.byref(out)))
    return float(out.value)


def algorithmic_flops(model: int) -> float:
    """Algorithmic FLOP per rollout-step: counted over the restated maths with general (dense) URDF
    constants by oracle/flop_count.py; the single figure roofline.achieved uses."""
    return float(load().mppi_algorithmic_flops_per_rollout_step(model))


def structural_flops(model: int) -> float:
    """FLOP per rollout-step once the exact 0 / +-1 constants of the baked chain are skipped (same counter, sparse
    constants; transcendental evaluations excluded)."""
    return float(load().mppi_structural_flops_per_rollout_step(model))
