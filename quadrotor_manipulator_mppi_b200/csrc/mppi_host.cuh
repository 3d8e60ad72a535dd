// mppi_host.cuh -- host-side state of a handle + the per-model launch entry points.
//
// The kernels are templates over (model, noise source, FK variant, cost terms, Philox rounds): ~110 instantiations.
// They are compiled as independent translation units, one per (model, part) -- mppi_model_unit.cu built with
// -DMPPI_UNIT_MODEL=<m> -DMPPI_UNIT_PART=<p> -- so the library builds in parallel (quadrotor_manipulator_mppi_b200/build.py);
// mppi_b200.cu (the C ABI) reaches them through the unit_* functions declared at the end of this file.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges cost nanoseconds unless a profiler is attached

#include "mppi_device.cuh"

struct mppi_ctx {
    mppi_config_t cfg{};
    mppi::StepParams P{};
    mppi::DynBlock dyn{};                 // host copy; passed by value at launch
    std::mutex state_mu;            // guards staged_state (set_state may come from another thread)
    float staged_state[MPPI_STATE_FLOATS]{};
    int nu = 0;
    int num_sms = 148;
    // device scratch (allocated once in mppi_create)
    float *d_cost = nullptr;        // [K]
    int32_t *d_rho = nullptr;       // order-preserving min (signed int32 encoding)
    uint32_t *d_counter = nullptr;  // last-block-done counter
    float *d_part = nullptr;        // [max_parts][T*nu+2]
    float *d_wsum = nullptr;        // [T*nu+2]
    unsigned long long *d_fix = nullptr;   // [T*nu+2] fixed-point accumulators of the Philox weighting pass
    float *d_w = nullptr;           // [K] unnormalised weights (injected-noise path)
    float *d_eta_part = nullptr;    // [<=SMs][2] partial sums of w, w^2
    int wn_resident = 0;            // resident blocks of the streaming weighting kernel (one wave)
    float *d_u = nullptr;           // [T*nu]   (host-buffer API)
    float *d_out = nullptr;         // [MPPI_OUT_FLOATS]
    float *d_noise = nullptr;       // host-buffer API with injected noise, grown on demand
    size_t d_noise_bytes = 0;
    float *h_pinned = nullptr;      // pinned staging for the host-buffer API
    size_t h_pinned_floats = 0;
    float *h_zc = nullptr;          // pinned + mapped: [MPPI_OUT_FLOATS] out vector + 1 sequence word (mppi_step_sync)
    float *d_zc = nullptr;          // the same memory as the device addresses it
    unsigned zc_seq = 0;
    int max_parts = 0;
    // NVLink peer exchange (mppi_p2p_export / mppi_p2p_bind)
    float *p2p_buf = nullptr;       // this rank's exchange buffer (cudaMalloc, exported through CUDA IPC)
    size_t p2p_bytes = 0;
    mppi::P2PParams X{};                  // world == 1 until bound
    mppi::P2PParams X_off{};              // world == 1: exchange disabled
    void *p2p_peer[mppi::kMaxRanks] = {};
    bool baked_fk = false;          // runtime chain == compile-time FkKinova tables
    double align[MPPI_MAX_JOINTS][9] = {};   // A_j: URDF link frame j -> folded link frame (z = joint axis), row-major
    float inertia_raw[7 * 10] = {};  // mass, com[3], inertia[6] per link, in the URDF link frames
    size_t rollout_smem[24] = {};   // tuned dynamic smem per kernel variant and Philox round count (0 = not yet tuned)
    float *d_qtraj = nullptr;       // [T][7] joint reference trajectory (MPPI_COST_JOINT_TRAJ), zeros by default
    cudaStream_t own_stream = nullptr;
    // ---- options (mppi_set_option)
    int philox_rounds = 7;          // 7 (smallest Crush-resistant round count, Salmon et al. SC'11) or 10 (Random123 / cuRAND default)
    int opt_fused = 0;              // single-launch step when the grid is co-resident (measured: the PDL-chained pair is as fast)
    int opt_timepar = -1;           // time-parallel warp-per-sample kernel (ARM7 / DRONE3): -1 auto, 0 off, 1 on when eligible
    int opt_profile = 0;            // record CUDA events around the kernels of every step (mppi_get_kernel_times)
    int opt_nvtx = 1;               // NVTX ranges around the launches
    int opt_host_yield = 0;         // blocking steps: sched_yield() between polls of the result word instead of a pure spin
    unsigned long long *d_trace = nullptr;      // MPPI_OPTION_TRACE: %globaltimer stamps of the step's phases
    // ---- single-launch steps: grid-wide barrier counter + cached launch shapes
    unsigned *d_sync = nullptr;     // monotonic arrival counter
    unsigned sync_target = 0;       // value it reaches after the launches issued so far
    int fused_blocks_max[8] = {-1, -1, -1, -1, -1, -1, -1, -1};   // co-resident capacity of step_fused_kernel per (baked, extra, rounds) variant
    int tp_blocks_max[32] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
    int last_path = 0;              // MPPI_PATH_* of the most recent step
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    bool ev_valid = false;
    unsigned *h_fail = nullptr;     // mapped host word: epoch at which the peer exchange timed out (0 = never)
    unsigned *d_fail = nullptr;
    std::string err;
};

inline thread_local std::string g_create_error;

inline mppi_status_t fail(mppi_handle_t h, mppi_status_t code, const std::string &msg)
{
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}
#define MPPI_CUDA(h, call)                                                                             \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(h, MPPI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

struct NvtxRange {
    bool on;
    NvtxRange(const mppi_ctx *h, const char *name) : on(h->opt_nvtx != 0) { if (on) nvtxRangePushA(name); }
    ~NvtxRange() { if (on) nvtxRangePop(); }
};

#define MPPI_UNIT_DECLS(M)                                                                                                         \
    mppi_status_t unit_rollout_##M(mppi_ctx *, const float *, const float *, float *, cudaStream_t);                                \
    mppi_status_t unit_weight_##M(mppi_ctx *, const float *, bool, const float *, float *, float *, cudaStream_t, const mppi::P2PParams &); \
    mppi_status_t unit_finalize_##M(mppi_ctx *, const float *, float *, float *, cudaStream_t);                                     \
    mppi_status_t unit_fused_##M(mppi_ctx *, const float *, float *, float *, cudaStream_t, const mppi::P2PParams &, bool *);       \
    mppi_status_t unit_tp_##M(mppi_ctx *, const float *, const float *, float *, float *, cudaStream_t, const mppi::P2PParams &, bool *);
MPPI_UNIT_DECLS(0)
MPPI_UNIT_DECLS(1)
MPPI_UNIT_DECLS(2)
MPPI_UNIT_DECLS(3)
