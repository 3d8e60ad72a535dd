// mppi_launch.cuh -- launch logic of the kernels, templated on the model; included by mppi_model_unit.cu only.
#pragma once
#include "mppi_host.cuh"
#include "mppi_kernels.cuh"

namespace mppi {

// The rollout kernel is issue-bound, so its time is (number of waves) x (blocks resident per SM).
// Pick the residency o <= o_max that minimises ceil(blocks / (SMs * o)) * o -- i.e. avoid a nearly
// empty last wave -- and enforce it by padding the dynamic shared memory request.
template <typename KernelT>
size_t tuned_rollout_smem(mppi_ctx *h, KernelT kernel, int threads, int grid, size_t smem_needed)
{
    int omax = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&omax, kernel, threads, smem_needed) != cudaSuccess || omax < 1) {
        cudaGetLastError();
        return smem_needed;
    }
    auto cost = [&](int o) { return (long long)((grid + (long long)h->num_sms * o - 1) / ((long long)h->num_sms * o)) * o; };
    int best_o = omax;
    long long best = cost(omax);
    // only one step below the register-limited residency, and never below 4 blocks (16 warps) per SM:
    // under that the kernel turns latency-bound and the wave model no longer holds
    for (int o = omax - 1; o >= omax - 1 && o >= 4; --o)
        if (cost(o) < best) { best = cost(o); best_o = o; }
    if (best_o == omax) return smem_needed;
    size_t pad = (size_t)(228 * 1024) / best_o - 1024 - 256;      // 1 KB per block is reserved by the driver
    pad &= ~(size_t)255;
    if (pad < smem_needed || pad > 48 * 1024) return smem_needed;
    int got = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&got, kernel, threads, pad) != cudaSuccess || got != best_o) {
        cudaGetLastError();
        return smem_needed;
    }
    return pad;
}


template <int MODEL, int NOISE, bool BAKED, bool EXTRA>
mppi_status_t launch_rollout_variant(mppi_ctx *h, int variant, const float *d_u_nom, const float *d_noise, float *d_cost,
                                     cudaStream_t st)
{
    constexpr int NU = ModelNu<MODEL>::value;
    constexpr int NUP = 4 * ((NU + 3) / 4);
    size_t smem = ((((size_t)h->P.T * NU + 3) & ~(size_t)3) + (size_t)h->P.T * NUP) * sizeof(float);     // u_nom raw + in the pair layout
    if (NOISE == 2) smem += (size_t)kNoiseStages * kRolloutThreads * NU * sizeof(float);
    const int grid = (h->P.K + kRolloutThreads - 1) / kRolloutThreads;
    auto launch = [&](auto kernel, size_t &tuned) -> mppi_status_t {
        if (tuned == 0) {
            if (smem > 48 * 1024)       // long horizons / wide noise tiles: opt in to the large dynamic shared memory carve-out
                MPPI_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            tuned = tuned_rollout_smem(h, kernel, kRolloutThreads, grid, smem);
        }
        kernel<<<grid, kRolloutThreads, tuned, st>>>(h->P, h->dyn, d_u_nom, d_noise, d_cost, h->d_rho, h->d_qtraj);
        MPPI_CUDA(h, cudaGetLastError());
        return MPPI_OK;
    };
    NvtxRange nv(h, "mppi.rollout_cost");
    if (NOISE == 0 && h->philox_rounds == 7)
        return launch(rollout_cost_kernel<MODEL, NOISE, BAKED, EXTRA, (NOISE == 0 ? 7 : 10)>, h->rollout_smem[12 + variant]);
    return launch(rollout_cost_kernel<MODEL, NOISE, BAKED, EXTRA, 10>, h->rollout_smem[variant]);
}

template <int MODEL, int NOISE>
mppi_status_t launch_rollout_noise(mppi_ctx *h, const float *d_u_nom, const float *d_noise, float *d_cost, cudaStream_t st)
{
    constexpr bool HAS_ARM = (MODEL == MPPI_MODEL_ARM7 || MODEL == MPPI_MODEL_WB11);
    const bool baked = HAS_ARM && h->baked_fk;      // FK unrolled from the URDF constants (fk_tables_gen.cuh)
    const bool extra = HAS_ARM && (h->P.cost_flags & MPPI_COST_MASK) != 0;      // optional cost terms: separate, slower instantiation
    const int variant = NOISE * 4 + (baked ? 1 : 0) + (extra ? 2 : 0);
    switch ((baked ? 1 : 0) + (extra ? 2 : 0)) {
        case 0: return launch_rollout_variant<MODEL, NOISE, false, false>(h, variant, d_u_nom, d_noise, d_cost, st);
        case 1: return launch_rollout_variant<MODEL, NOISE, HAS_ARM, false>(h, variant, d_u_nom, d_noise, d_cost, st);
        case 2: return launch_rollout_variant<MODEL, NOISE, false, HAS_ARM>(h, variant, d_u_nom, d_noise, d_cost, st);
        default: return launch_rollout_variant<MODEL, NOISE, HAS_ARM, HAS_ARM>(h, variant, d_u_nom, d_noise, d_cost, st);
    }
}

template <int MODEL>
mppi_status_t launch_rollout(mppi_ctx *h, const float *d_u_nom, const float *d_noise, float *d_cost, cudaStream_t st)
{
    constexpr int NU = ModelNu<MODEL>::value;
    if (!d_noise) return launch_rollout_noise<MODEL, 0>(h, d_u_nom, nullptr, d_cost, st);
    // injected [T][K][nu]: TMA-staged tiles when every 128-sample tile is 16-byte aligned and sized
    const bool tma_ok = ((size_t)h->P.K * NU) % 4 == 0 && (reinterpret_cast<uintptr_t>(d_noise) & 15u) == 0;
    if (tma_ok) return launch_rollout_noise<MODEL, 2>(h, d_u_nom, d_noise, d_cost, st);
    return launch_rollout_noise<MODEL, 1>(h, d_u_nom, d_noise, d_cost, st);
}

// ---- single-launch steps (cooperative: the grid-wide barrier needs every block resident)
template <typename KernelT, typename... Args>
mppi_status_t launch_cooperative(mppi_ctx *h, KernelT kernel, int grid, int threads, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(threads); lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    lc.attrs = at; lc.numAttrs = 1;
    MPPI_CUDA(h, cudaLaunchKernelEx(&lc, kernel, args...));
    return MPPI_OK;
}

template <typename KernelT>
int coresident_blocks(mppi_ctx *h, KernelT kernel, int threads, size_t smem, int &cache)
{
    if (cache < 0) {
        int per_sm = 0;
        if (smem > 48 * 1024 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            cache = 0;
            return 0;
        }
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess) { cudaGetLastError(); per_sm = 0; }
        cache = per_sm * h->num_sms;
    }
    return cache;
}

// rollout + weighting + exchange + finalize in ONE launch (step_fused_kernel); *launched = false when the shard does not
// qualify (injected noise, T > 128, or more blocks than can be co-resident) and the caller takes the two-kernel path.
template <int MODEL, bool BAKED, bool EXTRA, int ROUNDS>
mppi_status_t launch_fused_variant(mppi_ctx *h, int slot, const float *d_u_nom, float *d_u_new, float *d_out, cudaStream_t st,
                                   const P2PParams &X, bool *launched)
{
    constexpr int NU = ModelNu<MODEL>::value;
    constexpr int NUP = 4 * ((NU + 3) / 4);
    const int T = h->P.T, K = h->P.K;
    const int grid = (K + kRolloutThreads - 1) / kRolloutThreads;
    const int R = kRolloutThreads / T;
    size_t floats = (size_t)2 * kRolloutThreads + (size_t)R * T * NUP;            // weights, indices, reduction
    if (floats < (size_t)T * NU + 4 + (size_t)T * NUP) floats = (size_t)T * NU + 4 + (size_t)T * NUP;   // staged nominal sequence, raw + pair layout
    if (floats < (size_t)4 * T * NU + NU + 16) floats = (size_t)4 * T * NU + NU + 16;   // finalize scratch + smem hand-over of sums and u_nom
    const size_t smem = floats * sizeof(float);
    auto kernel = step_fused_kernel<MODEL, BAKED, EXTRA, ROUNDS>;
    if (grid > coresident_blocks(h, kernel, kRolloutThreads, smem, h->fused_blocks_max[slot])) return MPPI_OK;
    NvtxRange nv(h, "mppi.step_fused");
    h->sync_target += (unsigned)grid;
    mppi_status_t rc = launch_cooperative(h, kernel, grid, kRolloutThreads, smem, st, h->P, h->dyn, d_u_nom, h->d_cost, h->d_rho,
                                          (const float *)h->d_qtraj, h->d_fix, h->d_counter, h->d_sync, h->sync_target, h->d_wsum,
                                          d_u_new, d_out, X);
    if (rc != MPPI_OK) { h->sync_target -= (unsigned)grid; return rc; }
    *launched = true;
    return MPPI_OK;
}

template <int MODEL>
mppi_status_t launch_fused(mppi_ctx *h, const float *d_u_nom, float *d_u_new, float *d_out, cudaStream_t st, const P2PParams &X,
                           bool *launched)
{
    constexpr bool HAS_ARM = (MODEL == MPPI_MODEL_ARM7 || MODEL == MPPI_MODEL_WB11);
    *launched = false;
    if (!h->opt_fused || h->P.T > kRolloutThreads) return MPPI_OK;
    const bool baked = HAS_ARM && h->baked_fk;
    const bool extra = HAS_ARM && (h->P.cost_flags & MPPI_COST_MASK) != 0;
    const bool r7 = h->philox_rounds == 7;
    const int slot = (baked ? 1 : 0) + (extra ? 2 : 0) + (r7 ? 4 : 0);
#define MPPI_FUSED_CASE(B, E)                                                                                              \
    return r7 ? launch_fused_variant<MODEL, B, E, 7>(h, slot, d_u_nom, d_u_new, d_out, st, X, launched)                    \
              : launch_fused_variant<MODEL, B, E, 10>(h, slot, d_u_nom, d_u_new, d_out, st, X, launched)
    switch ((baked ? 1 : 0) + (extra ? 2 : 0)) {
        case 0: MPPI_FUSED_CASE(false, false);
        case 1: MPPI_FUSED_CASE(HAS_ARM, false);
        case 2: MPPI_FUSED_CASE(false, HAS_ARM);
        default: MPPI_FUSED_CASE(HAS_ARM, HAS_ARM);
    }
#undef MPPI_FUSED_CASE
}

// Time-parallel warp-per-sample step (step_tp_kernel): ARM7 / DRONE3, default cost terms, T <= 64.
// "auto": beyond these the thread-per-sample kernels fill the machine on their own (profiles/r02/sweep_1gpu.md: the arm
// ties at K = 16384, the much cheaper point-mass step at ~8192)
constexpr int kTpAutoMaxSamplesArm = 8192, kTpAutoMaxSamplesDrone = 4096, kTpAutoMaxSamplesDroneLong = 2048;
template <int MODEL, int NOISE, bool BAKED, int SPL, int ROUNDS>
mppi_status_t launch_tp_variant(mppi_ctx *h, int slot, const float *d_u_nom, const float *d_noise, float *d_u_new, float *d_out,
                                cudaStream_t st, const P2PParams &X, bool *launched)
{
    constexpr int NU = ModelNu<MODEL>::value;
    constexpr int THREADS = tp_threads(SPL), WARPS = THREADS / 32;
    const size_t n = (size_t)h->P.T * NU;
    size_t floats = (1 + (size_t)WARPS) * n;                          // block accumulator + tile contributions of the block's warps
    size_t off_w = 2 * n + NU;                                        // hand-over layout of the last block (see the kernel)
    const size_t rstride = (n + 2 + 3) & ~(size_t)3;
    const size_t red = rstride > (size_t)4 * THREADS ? rstride : (size_t)4 * THREADS;      // combine partials: parts * rstride floats
    if (off_w < n + kTpMaxRows + 4 + red) off_w = n + kTpMaxRows + 4 + red;
    const size_t tail = off_w + 4 + (n + 2) + 2 + n;
    if (floats < tail) floats = tail;
    const size_t smem = floats * sizeof(float);
    auto kernel = step_tp_kernel<MODEL, NOISE, BAKED, SPL, ROUNDS>;
    int &ready = h->tp_blocks_max[slot];
    if (ready < 0) {
        if (smem > 48 * 1024) MPPI_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ready = 1;
    }
    const int n_tiles = (h->P.K + WARPS - 1) / WARPS;
    int cap = h->num_sms < kTpMaxRows ? h->num_sms : kTpMaxRows;      // one row per block, one block per SM
    if (cap > h->max_parts) cap = h->max_parts;
    const int grid = n_tiles < cap ? n_tiles : cap;                  // persistent blocks loop over the remaining tiles
    NvtxRange nv(h, "mppi.step_timeparallel");
    kernel<<<grid, THREADS, smem, st>>>(h->P, h->dyn, d_u_nom, d_noise, h->d_cost, h->d_rho, h->d_part, h->d_eta_part,
                                        h->d_counter, h->d_wsum, d_u_new, d_out, X);
    MPPI_CUDA(h, cudaGetLastError());
    *launched = true;
    return MPPI_OK;
}

template <int MODEL, int NOISE, bool BAKED, int ROUNDS>
mppi_status_t launch_tp_spl(mppi_ctx *h, int slot, int spl, const float *d_u_nom, const float *d_noise, float *d_u_new, float *d_out,
                            cudaStream_t st, const P2PParams &X, bool *launched)
{
    switch (spl) {
        case 1: return launch_tp_variant<MODEL, NOISE, BAKED, 1, ROUNDS>(h, slot, d_u_nom, d_noise, d_u_new, d_out, st, X, launched);
        case 2: return launch_tp_variant<MODEL, NOISE, BAKED, 2, ROUNDS>(h, slot, d_u_nom, d_noise, d_u_new, d_out, st, X, launched);
        case 4: return launch_tp_variant<MODEL, NOISE, BAKED, 4, ROUNDS>(h, slot, d_u_nom, d_noise, d_u_new, d_out, st, X, launched);
        default: return launch_tp_variant<MODEL, NOISE, BAKED, 8, ROUNDS>(h, slot, d_u_nom, d_noise, d_u_new, d_out, st, X, launched);
    }
}

template <int MODEL>
mppi_status_t launch_tp(mppi_ctx *h, const float *d_u_nom, const float *d_noise, float *d_u_new, float *d_out, cudaStream_t st,
                        const P2PParams &X, bool *launched)
{
    *launched = false;
    if constexpr (MODEL == MPPI_MODEL_ARM7 || MODEL == MPPI_MODEL_DRONE3) {
        constexpr bool ARM_MODEL = (MODEL == MPPI_MODEL_ARM7);
        if (h->opt_timepar == 0 || h->P.T > 256 || (h->P.cost_flags & MPPI_COST_MASK) != 0 || h->P.K > (1 << 20)) return MPPI_OK;
        // measured, T = 256 (8 steps per lane): the arm still wins at K = 8192 (115 vs 117 us), the point mass only up to ~2048
        const int auto_max = ARM_MODEL ? kTpAutoMaxSamplesArm : (h->P.T > 128 ? kTpAutoMaxSamplesDroneLong : kTpAutoMaxSamplesDrone);
        if (h->opt_timepar < 0 && h->P.K > auto_max) return MPPI_OK;
        constexpr bool ARM = (MODEL == MPPI_MODEL_ARM7);
        const bool baked = ARM && h->baked_fk;
        const bool r7 = h->philox_rounds == 7 && !d_noise;
        const int spl = h->P.T <= 32 ? 1 : h->P.T <= 64 ? 2 : h->P.T <= 128 ? 4 : 8;       // horizon steps per lane
        const int slot = (baked ? 1 : 0) + (d_noise ? 2 : 0) + (r7 ? 4 : 0) + 8 * (spl == 1 ? 0 : spl == 2 ? 1 : spl == 4 ? 2 : 3);
        if (d_noise) {
            if (baked) return launch_tp_spl<MODEL, 1, ARM, 10>(h, slot, spl, d_u_nom, d_noise, d_u_new, d_out, st, X, launched);
            return launch_tp_spl<MODEL, 1, false, 10>(h, slot, spl, d_u_nom, d_noise, d_u_new, d_out, st, X, launched);
        }
        if (r7) {
            if (baked) return launch_tp_spl<MODEL, 0, ARM, 7>(h, slot, spl, d_u_nom, d_noise, d_u_new, d_out, st, X, launched);
            return launch_tp_spl<MODEL, 0, false, 7>(h, slot, spl, d_u_nom, d_noise, d_u_new, d_out, st, X, launched);
        }
        if (baked) return launch_tp_spl<MODEL, 0, ARM, 10>(h, slot, spl, d_u_nom, d_noise, d_u_new, d_out, st, X, launched);
        return launch_tp_spl<MODEL, 0, false, 10>(h, slot, spl, d_u_nom, d_noise, d_u_new, d_out, st, X, launched);
    }
    return MPPI_OK;
}

template <int MODEL>
mppi_status_t launch_weight(mppi_ctx *h, const float *d_noise, bool fuse, const float *d_u_nom, float *d_u_new,
                            float *d_out, cudaStream_t st, const P2PParams &X)
{
    constexpr int NU = ModelNu<MODEL>::value;
    const int K = h->P.K, T = h->P.T;
    const size_t fin_floats = (size_t)4 * T * NU + NU + 16;      // finalize scratch (2n + nu) + the sums and u_nom handed over in smem
    if (!d_noise) {
        const int TC = T;                       // one thread per horizon step (all Philox calls of the step), R sample sub-ranges
        int R = 512 / TC;
        if (R < 1) R = 1;
        int threads = ((TC * R + 31) / 32) * 32;
        if (threads > 1024) return fail(h, MPPI_ERR_UNSUPPORTED, "horizon too long for the Philox weighting kernel");
        // small K: many short blocks (latency), large K: two per SM.  Fewer, fatter blocks for small K were measured
        // (128 ... 1024 samples per block): no change with collapsed weights, slower with dense ones (K = 32768: 80 -> 97 us
        // at 512 per block)
        int blocks = (K + 31) / 32;
        if (blocks > 2 * h->num_sms) blocks = 2 * h->num_sms;
        if (blocks > h->max_parts) blocks = h->max_parts;
        if (blocks < 1) blocks = 1;
        const int chunk = (K + blocks - 1) / blocks;
        blocks = (K + chunk - 1) / chunk;
        size_t smem_floats = (size_t)2 * kWeightTile + (size_t)R * TC * (4 * ((NU + 3) / 4));
        if (smem_floats < fin_floats) smem_floats = fin_floats;
        // fused steps launch it as a programmatic dependent of the rollout kernel: its launch overlaps the rollout's drain
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(blocks); lc.blockDim = dim3(threads); lc.dynamicSmemBytes = smem_floats * sizeof(float); lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = fuse ? 1 : 0;
        NvtxRange nv(h, "mppi.weight_philox");
        if (h->philox_rounds == 7)
            MPPI_CUDA(h, cudaLaunchKernelEx(&lc, weight_philox_kernel<MODEL, 7>, h->P, h->dyn, (const float *)h->d_cost, h->d_rho, chunk, h->d_fix,
                                            h->d_counter, h->d_wsum, fuse ? 1 : 0, d_u_nom, d_u_new, d_out, X));
        else
            MPPI_CUDA(h, cudaLaunchKernelEx(&lc, weight_philox_kernel<MODEL, 10>, h->P, h->dyn, (const float *)h->d_cost, h->d_rho, chunk, h->d_fix,
                                            h->d_counter, h->d_wsum, fuse ? 1 : 0, d_u_nom, d_u_new, d_out, X));
    } else {
        const bool vec4 = ((size_t)K * NU) % 4 == 0 && (reinterpret_cast<uintptr_t>(d_noise) & 15u) == 0;
        NvtxRange nv(h, "mppi.weights+weighted_noise");
        // weights once, then one resident wave of (G x T) streaming blocks
        int wblocks = (K + 1023) / 1024;
        if (wblocks > h->num_sms) wblocks = h->num_sms;
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(wblocks); lc.blockDim = dim3(256); lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = fuse ? 1 : 0;
        MPPI_CUDA(h, cudaLaunchKernelEx(&lc, weights_kernel, h->P, (const float *)h->d_cost, (const int32_t *)h->d_rho, h->d_w, h->d_eta_part));
        const int threads = 32 * NU;
        size_t smem_floats = (size_t)threads * 4;
        if (smem_floats < fin_floats) smem_floats = fin_floats;
        if (h->wn_resident == 0) {
            int per_sm = 0;
            cudaError_t oe = vec4 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, weighted_noise_kernel<MODEL, 4>, threads, smem_floats * sizeof(float))
                                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, weighted_noise_kernel<MODEL, 1>, threads, smem_floats * sizeof(float));
            if (oe != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 2; }
            h->wn_resident = per_sm * h->num_sms;
        }
        int G = h->wn_resident / T;                       // small problems: one resident wave (latency)
        if (G < 1) G = 1;
        // Large problems: ~300 KB of noise per block and a block count that is a multiple of the SM count, so every SM
        // streams the same number of equal blocks (measured on wb K=262144, T=64: one wave of 11 x 64 blocks 5.27 TB/s,
        // 37 x 64 blocks = 16 per SM 5.81 TB/s, 148 x 64 blocks 4.65 TB/s -- profiles/r01/README.md).
        {
            const double g_target = (double)K / std::fmax(1024.0, 300e3 / (NU * 4.0));
            if (g_target * T >= 8.0 * h->num_sms) {
                int a = h->num_sms, b = T;
                while (b) { const int r = a % b; a = b; b = r; }
                const int g0 = h->num_sms / a;             // smallest G with G*T % num_sms == 0
                const int m = (int)std::floor(g_target / g0 + 0.5);
                G = m >= 1 ? m * g0 : (int)(g_target + 0.5);
            }
        }
        if (G > h->max_parts) G = h->max_parts;
        const int gmax = (K + 127) / 128;
        if (G > gmax) G = gmax;
        int chunk = ((K + G - 1) / G + 127) / 128 * 128;
        G = (K + chunk - 1) / chunk;
        dim3 grid(G, T);
        if (vec4)
            weighted_noise_kernel<MODEL, 4><<<grid, threads, smem_floats * sizeof(float), st>>>(
                h->P, h->dyn, h->d_w, d_noise, h->d_rho, chunk, h->d_part, h->d_eta_part, wblocks, h->d_counter, h->d_wsum,
                fuse ? 1 : 0, d_u_nom, d_u_new, d_out, X);
        else
            weighted_noise_kernel<MODEL, 1><<<grid, threads, smem_floats * sizeof(float), st>>>(
                h->P, h->dyn, h->d_w, d_noise, h->d_rho, chunk, h->d_part, h->d_eta_part, wblocks, h->d_counter, h->d_wsum,
                fuse ? 1 : 0, d_u_nom, d_u_new, d_out, X);
    }
    MPPI_CUDA(h, cudaGetLastError());
    return MPPI_OK;
}

template <int MODEL>
mppi_status_t launch_finalize(mppi_ctx *h, const float *d_u_nom, float *d_u_new, float *d_out, cudaStream_t st)
{
    constexpr int NU = ModelNu<MODEL>::value;
    const size_t smem = ((size_t)2 * h->P.T * NU + NU) * sizeof(float);
    finalize_kernel<MODEL><<<1, 256, smem, st>>>(h->P, h->dyn, h->d_wsum, d_u_nom, d_u_new, d_out, h->d_rho);
    MPPI_CUDA(h, cudaGetLastError());
    return MPPI_OK;
}

}  // namespace mppi
