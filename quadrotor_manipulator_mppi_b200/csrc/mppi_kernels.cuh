// mppi_kernels.cuh -- the kernels of one MPPI control step.
//
//   K2  rollout_cost_kernel   : noise (in-kernel Philox or injected [T][K][nu]) + rollout +
//                               forward kinematics + cost, one thread per sample, state in
//                               registers over the horizon loop; S[K] and the cost minimum
//                               are the only things written.
//   K3  weight_philox_kernel  : w = exp((rho-S)/lambda) and the weighted-noise sums with the noise REGENERATED for the
//                               samples whose weight is not zero (compute-bound; fixed-point integer atomics), or
//       weights_kernel + weighted_noise_kernel : the same with injected noise re-read once from HBM (float4 stream,
//                               per-block partials reduced in a fixed order).
//                               The last block to finish runs K4 in place on a single GPU; with K sharded over GPUs it
//                               first exchanges its row with the peers over NVLink (p2p_exchange).
//   K4  finalize_block        : normalise, Savitzky-Golay, u += w_eps, controller outputs, check_reach, statistics,
//                               the arm node's torque law (mppi_dynamics.cuh), zero-copy result for blocking callers.
#pragma once
#include "mppi_device.cuh"
#include "mppi_dynamics.cuh"

namespace mppi {

#ifndef MPPI_ROLLOUT_THREADS
#define MPPI_ROLLOUT_THREADS 128
#define MPPI_ROLLOUT_MINB 4
#endif
constexpr int kRolloutThreads = MPPI_ROLLOUT_THREADS;
constexpr int kWeightTile = 2048;             // samples whose weights are staged in shared memory at a time
constexpr float kFixScale = 8589934592.0f;   // 2^33: fixed-point scale of the weighting accumulators

// ------------------------------------------------------------------------------------------
// K2: fused noise + rollout + FK + cost.
// Replaces S/mppi_solver/mppi.py:129-140 (sampling, get_sample_joint, compute_fk_gpu,
// CostManager.compute_all_cost) and S/mppi_solver/drone_mppi.py:143-151.
//
// One thread per sample, state in registers across the horizon loop.  The kernel is issue-bound, so
// the FP32 work of ONE sample is packed into Blackwell's FP32x2 instructions (FFMA2 / FMUL2 / FADD2:
// same FLOP rate as scalar FFMA, half the issue slots -- tools/probe_ffma2.cu):
//   * the seven arm joints travel as pairs (0,2) (1,3) (4,6) (5,-) -- the pairing in which the packed
//     Box-Muller emits its normals -- through the double integrator and sin/cos (with the quadrotor's
//     Euler angles filling the spare lanes for the whole-body model);
//   * the FK keeps each rotation column as a (row 0, row 1) pair + a row-2 scalar (Pose3);
//   * the two atan2 of the Euler-angle cost run as one packed evaluation.
// Carrying two samples per thread in f2 was measured too: parity-green but latency-bound at 3 warps per
// scheduler (+3 % only), so packing within a sample at full occupancy is what ships.
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int pairA(int i) { return i == 0 ? 0 : i == 1 ? 1 : i == 2 ? 4 : 5; }     // joint in lane 0 of arm pair i
__host__ __device__ constexpr int pairB(int i) { return i == 0 ? 2 : i == 1 ? 3 : i == 2 ? 6 : -1; }    // joint in lane 1 (-1: spare lane)

// NOISE: 0 = in-kernel Philox, 1 = injected [T][K][nu] read directly (any K, any alignment),
//        2 = injected, staged through shared memory by TMA bulk copies (one 128-sample tile per horizon
//            step, kNoiseStages deep; needs K*nu % 4 == 0 and a 16-byte aligned tensor).
constexpr int kNoiseStages = 4;

// Scheduling shape of the horizon loop, per model.  Each is a compile-time knob so that tools/build_variant.py can build the
// alternatives side by side and tools/ab_variants.py can time them in one GPU call; the defaults are the measured best
// (profiles/r02/README.md, "Experiments measured and not shipped"):
//   *_NB        noise of NB horizon steps generated at a time, then the NB steps run back to back.  Point mass / rigid
//               body: 4 -- their step is a short dependent chain behind one Philox call, so the chains of several steps
//               overlap (2 / 6 / 8 measured slower).  Whole body: 1 (2 measured slower).
//   SMALL_NOBREAK  the batch is ONE basic block: steps past a ragged horizon are computed and not accumulated (-4 %).
//   *_PREFETCH  the noise of step t+1 is generated while step t runs.  Whole body: on (-2 %, -3.5 % at small shards);
//               arm (+1.3 %) and the small models (+7 %): off.
//   *_UNROLL    horizon loop unrolled: arm 2 (-1 %), whole body 1 (no change).
#ifndef MPPI_WB_NB
#define MPPI_WB_NB 1
#endif
#ifndef MPPI_WB_UNROLL
#define MPPI_WB_UNROLL 1
#endif
#ifndef MPPI_ARM_UNROLL
#define MPPI_ARM_UNROLL 2
#endif
#ifndef MPPI_SMALL_NB
#define MPPI_SMALL_NB 4
#endif
#ifndef MPPI_SMALL_NOBREAK
#define MPPI_SMALL_NOBREAK 1
#endif
#ifndef MPPI_SMALL_PREFETCH
#define MPPI_SMALL_PREFETCH 0
#endif
#ifndef MPPI_WB_PREFETCH
#define MPPI_WB_PREFETCH 1
#endif
#ifndef MPPI_ARM_PREFETCH
#define MPPI_ARM_PREFETCH 0
#endif
template <int MODEL, int NOISE> struct NoiseBatch {
    static constexpr int value = (NOISE == 0 && (MODEL == MPPI_MODEL_DRONE3 || MODEL == MPPI_MODEL_QUAD4)) ? MPPI_SMALL_NB
                               : (NOISE == 0 && MODEL == MPPI_MODEL_WB11) ? MPPI_WB_NB : 1;
};

// The body of K2, shared by rollout_cost_kernel and the single-launch step_fused_kernel.  Returns the sample's cost
// (also stored to cost_out[k]); the block minimum has been folded into *rho_enc on return.
template <int MODEL, int NOISE, bool BAKED, bool EXTRA, int ROUNDS>
__device__ __forceinline__ float rollout_body(const StepParams &P, const DynBlock &D,
                    const float *__restrict__ u_nom, const float *__restrict__ noise,
                    float *__restrict__ cost_out, int32_t *__restrict__ rho_enc,
                    const float *__restrict__ q_traj, bool &active_out)
{
    constexpr int NU = ModelNu<MODEL>::value;
    constexpr int NCH = (NU + 3) / 4;
    constexpr bool HAS_ARM = (MODEL == MPPI_MODEL_ARM7 || MODEL == MPPI_MODEL_WB11);
    constexpr bool HAS_QUAD = (MODEL == MPPI_MODEL_QUAD4 || MODEL == MPPI_MODEL_WB11);
    constexpr int ARM0 = (MODEL == MPPI_MODEL_WB11) ? 4 : 0;     // first arm input
    constexpr int QOFF = (MODEL == MPPI_MODEL_WB11) ? 12 : 0;    // arm q in the state vector

    constexpr bool PHILOX = (NOISE == 0);
    constexpr bool FAST_TRIG = (MODEL == MPPI_MODEL_QUAD4 || MODEL == MPPI_MODEL_WB11);       // MUFU sin/cos: unpinned models only
    // Unpinned whole body on the baked chain without the extra cost terms: the joints integrate as the state recurrence
    // q += v dt + a dt^2/2, v += a dt (3 packed FMAs per joint pair) instead of the reference's two cumulative sums + q0
    // (6 packed operations, kept bit for bit for the pinned arm).  Same discrete dynamics, different rounding (1e-7).
    constexpr bool DIRECT = (MODEL == MPPI_MODEL_WB11) && BAKED && !EXTRA;
    extern __shared__ __align__(16) float s_unom[];              // [T][NU] | (NOISE == 2) noise tiles [kNoiseStages][128][NU]
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(8) uint64_t s_full[kNoiseStages];
    __shared__ float s_wmin[kRolloutThreads / 32];

    // ---- stage the nominal control sequence: one TMA bulk copy + scalar tail
    const int n_u = P.T * NU;
    const uint32_t bulk_bytes = (static_cast<uint32_t>(n_u) * 4u) & ~15u;
    const bool bulk_ok = bulk_bytes > 0 && ((reinterpret_cast<uintptr_t>(u_nom) & 15u) == 0);
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        if constexpr (NOISE == 2) {
#pragma unroll
            for (int st = 0; st < kNoiseStages; ++st) mbar_init(&s_full[st], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();
    // ---- injected noise through TMA: the block's samples [k0, k0 + nb) of row t are nb*NU contiguous floats
    const int k0_blk = blockIdx.x * kRolloutThreads;
    const int nb = min(kRolloutThreads, P.K - k0_blk);
    // the same sequence once more in the pair layout the loop consumes: per step and chunk c one float4
    // (u[4c], u[4c+2], u[4c+1], u[4c+3]), zero-padded -- one LDS.128 per chunk instead of nu scalar loads + pair moves
    float *s_upair = s_unom + ((n_u + 3) & ~3);
    float *s_tiles = s_upair + P.T * 4 * NCH;                    // 16-byte aligned behind both copies
    const uint32_t tile_bytes = static_cast<uint32_t>(nb) * NU * 4u;
    if constexpr (NOISE == 2) {
        if (threadIdx.x == 0) {
            for (int st = 0; st < kNoiseStages && st < P.T; ++st) {
                mbar_expect_tx(&s_full[st], tile_bytes);
                tma_bulk_g2s(s_tiles + st * (kRolloutThreads * NU), noise + (static_cast<size_t>(st) * P.K + k0_blk) * NU,
                             tile_bytes, &s_full[st]);
            }
        }
    }
    if (bulk_ok) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&s_bar, bulk_bytes);
            tma_bulk_g2s(s_unom, u_nom, bulk_bytes, &s_bar);
        }
        for (int j = (bulk_bytes >> 2) + threadIdx.x; j < n_u; j += blockDim.x) s_unom[j] = u_nom[j];
        mbar_wait(&s_bar, 0);
    } else {
        for (int j = threadIdx.x; j < n_u; j += blockDim.x) s_unom[j] = u_nom[j];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < P.T * 4 * NCH; j += blockDim.x) {
        const int t = j / (4 * NCH), e = j - t * (4 * NCH);
        const int i = (e & ~3) + ((e & 1) << 1) + ((e >> 1) & 1);          // slot (x, y, z, w) of chunk c <- input 4c + (0, 2, 1, 3)
        s_upair[j] = i < NU ? s_unom[t * NU + i] : 0.f;
    }
    __syncthreads();

    const int k_raw = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = k_raw < P.K;
    const int k = active ? k_raw : P.K - 1;
    const uint32_t kg = static_cast<uint32_t>(P.k_offset + k);
    trace_stamp(D, TR_START, blockIdx.x == 0 && threadIdx.x == 0);

    // ---- per-sample state in registers
    float dcv[3], dcq[3], dvp[3];                  // DRONE3 double integrator
    f2 cum_v[4], cum_q[4];                         // arm joints in pairs (pairA, pairB)
    f2 q0p[4], qd0p[4];                            // measured joint state in the same pairing (uniform)
    f2 q0t[4];                                     // FAST_TRIG: q0 less its whole turns, the base of the sin / cos arguments
    QuadState<float> qs;
    Pose3 base0;                                   // chain root pose composed with C0 (uniform, loop-invariant for ARM7)
    if constexpr (MODEL == MPPI_MODEL_DRONE3) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { dcv[i] = 0.f; dcq[i] = 0.f; dvp[i] = D.state[3 + i]; }
    }
    if constexpr (HAS_ARM) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ja = pairA(i), jb = pairB(i);
            q0p[i] = f2(D.state[QOFF + ja], jb >= 0 ? D.state[QOFF + jb] : 0.f);
            qd0p[i] = f2(D.state[QOFF + 7 + ja], jb >= 0 ? D.state[QOFF + 7 + jb] : 0.f);
            cum_v[i] = f2(0.f); cum_q[i] = f2(0.f);
            q0t[i] = FAST_TRIG ? whole_turns_removed(q0p[i]) : q0p[i];
            if constexpr (DIRECT) { cum_q[i] = q0t[i]; cum_v[i] = qd0p[i]; }     // the state itself: (angle less whole turns, velocity)
        }
    }
    if constexpr (MODEL == MPPI_MODEL_ARM7) {
        float R0[9], p0[3];
        quat_matrix(&D.state[14], R0);                       // base xyz+quat -> B (S/robot/urdf_fk.py:30-55)
        p0[0] = D.state[14]; p0[1] = D.state[15]; p0[2] = D.state[16];
        if constexpr (BAKED) compose_tab<FkKinova, 0>(R0, p0);
        else compose_const(R0, p0, P.chain.R[0], P.chain.t[0]);
        base0 = pose3_from(R0, p0);
    }
    if constexpr (HAS_QUAD) quad_load(qs, D.state);

    float S = 0.f, comp = 0.f;       // Kahan-compensated running cost
    float x_cov = 0.f, x_cen = 0.f, x_trk = 0.f, x_act = 0.f, x_lim = 0.f, gpow = 1.0f;   // EXTRA cost terms, gamma^t
    float Sd = 0.f;                  // squared-distance stage cost (drone / quad part)
    float term_d = 0.f;

    constexpr int NB = NoiseBatch<MODEL, NOISE>::value;
    // controls of horizon step t into the pair layout (a02, a13)
    auto load_controls = [&](const int t, f2 *a02, f2 *a13) {
        // ---- controls of this step: v = u + noise  (S/mppi_solver/mppi.py:130).  Inputs 4c..4c+3 arrive
        // as the pairs (4c, 4c+2) and (4c+1, 4c+3).
        if constexpr (NOISE == 2) mbar_wait(&s_full[t % kNoiseStages], (t / kNoiseStages) & 1);
        float unif[6 * philox_calls(NU)];
        if constexpr (PHILOX)
            philox_step_uniforms<philox_calls(NU), ROUNDS>(kg, static_cast<uint32_t>(t), D.step_lo, D.step_hi, P.rkeys, unif);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int i0 = 4 * c, i1 = 4 * c + 1, i2 = 4 * c + 2, i3 = 4 * c + 3;
            const float4 u4 = *reinterpret_cast<const float4 *>(s_upair + (t * NCH + c) * 4);
            if constexpr (PHILOX) {
                f2 n02, n13;
                normals_quad(unif, c, n02, n13);
                // eps = sigma n is rounded on its own (it is a tensor in the reference, and the rollout on materialised noise
                // must reproduce the in-kernel one bit for bit), then v = u + eps
                a02[c] = vmul(f2(P.sigma[i0], i2 < NU ? P.sigma[i2] : 0.f), n02);
                a13[c] = vmul(f2(i1 < NU ? P.sigma[i1] : 0.f, i3 < NU ? P.sigma[i3] : 0.f), n13);
            } else if constexpr (NOISE == 1) {
                const float *row = noise + (static_cast<size_t>(t) * P.K + k) * NU;
                a02[c] = f2(__ldg(row + i0), i2 < NU ? __ldg(row + i2) : 0.f);
                a13[c] = f2(i1 < NU ? __ldg(row + i1) : 0.f, i3 < NU ? __ldg(row + i3) : 0.f);
            } else {
                // tile of this horizon step, landed in shared memory through TMA (stride NU words per thread:
                // conflict-free for odd NU; NU == 4 reads one float4 per thread)
                const float *row = s_tiles + (t % kNoiseStages) * (kRolloutThreads * NU) + (k - k0_blk) * NU;
                if constexpr (NU == 4) {
                    const float4 v4 = *reinterpret_cast<const float4 *>(row);
                    a02[c] = f2(v4.x, v4.z);
                    a13[c] = f2(v4.y, v4.w);
                } else {
                    a02[c] = f2(row[i0], i2 < NU ? row[i2] : 0.f);
                    a13[c] = f2(i1 < NU ? row[i1] : 0.f, i3 < NU ? row[i3] : 0.f);
                }
            }
            a02[c] = vadd(f2(u4.x, u4.y), a02[c]);
            a13[c] = vadd(f2(u4.z, u4.w), a13[c]);
        }
        if constexpr (NOISE == 2) {
            __syncthreads();                       // every thread has taken its controls out of the stage
            if (threadIdx.x == 0 && t + kNoiseStages < P.T) {
                const int st = t % kNoiseStages;
                mbar_expect_tx(&s_full[st], tile_bytes);
                tma_bulk_g2s(s_tiles + st * (kRolloutThreads * NU),
                             noise + (static_cast<size_t>(t + kNoiseStages) * P.K + k0_blk) * NU, tile_bytes, &s_full[st]);
            }
        }
    };
    // whole body: the noise of step t + 1 is generated while step t runs (independent instruction streams in one
    // basic block: the Box-Muller MUFU work overlaps the FK arithmetic inside every warp, not only across warps)
    constexpr bool PREFETCH = PHILOX && !EXTRA &&
                              ((MODEL == MPPI_MODEL_WB11 && MPPI_WB_PREFETCH != 0) || (MODEL == MPPI_MODEL_ARM7 && MPPI_ARM_PREFETCH != 0) ||
                               ((MODEL == MPPI_MODEL_DRONE3 || MODEL == MPPI_MODEL_QUAD4) && MPPI_SMALL_PREFETCH != 0));
    f2 nx02[NB][NCH], nx13[NB][NCH];
    if constexpr (PREFETCH) {
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) load_controls(min(jb, P.T - 1), nx02[jb], nx13[jb]);
    }
    constexpr int UNROLL_T = (MODEL == MPPI_MODEL_WB11) ? MPPI_WB_UNROLL : (MODEL == MPPI_MODEL_ARM7) ? MPPI_ARM_UNROLL : 1;
#if MPPI_WB_UNROLL > 1 || MPPI_ARM_UNROLL > 1
#pragma unroll UNROLL_T
#endif
    for (int t0 = 0; t0 < P.T; t0 += NB) {
      f2 a02b[NB][NCH], a13b[NB][NCH];
      if constexpr (PREFETCH) {
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) { a02b[jb][c] = nx02[jb][c]; a13b[jb][c] = nx13[jb][c]; }
        }
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) load_controls(min(t0 + NB + jb, P.T - 1), nx02[jb], nx13[jb]);
      } else {
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) load_controls(min(t0 + jb, P.T - 1), a02b[jb], a13b[jb]);
      }
#pragma unroll
      for (int jb = 0; jb < NB; ++jb) {
        const int t = t0 + jb;
        if (NB > 1 && !MPPI_SMALL_NOBREAK && t >= P.T) break;
        const bool valid = (NB == 1) || t < P.T;       // past the horizon (ragged last batch): nothing is accumulated
        const f2 *a02 = a02b[jb], *a13 = a13b[jb];
        // scalar view: input i lives in (i & 1 ? a13 : a02)[i >> 2], lane (i >> 1) & 1
        auto input = [&](int i) -> float { return lane((i & 1) ? a13[i >> 2] : a02[i >> 2], (i >> 1) & 1); };

        const bool last = (t == P.T - 1);
        if constexpr (MODEL == MPPI_MODEL_DRONE3) {
            // double integrator (S/mppi_solver/drone_mppi.py:46-55) + squared distance (:87-107)
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float ai = input(i);
                const float dq = fmaf(dvp[i], P.dt, ai * (0.5f * P.dt2));      // (0.5 a) dt^2 exactly: halving is exact
                dcv[i] = fmaf(ai, P.dt, dcv[i]);
                dvp[i] = dcv[i] + D.state[3 + i];
                dcq[i] += dq;
                const float e = (dcq[i] + D.state[i]) - D.drone_target[i];
                sq = fmaf(e, e, sq);
            }
            if (last) term_d = sq; else if (valid) Sd += sq;
        }
        if constexpr (HAS_QUAD) {
            quad_advance<false, !FAST_TRIG>(qs, input(0), input(1), input(2), input(3), P.dt, P.quad);   // sin/cos refreshed below
            const float ex = qs.p[0] - D.drone_target[0], ey = qs.p[1] - D.drone_target[1], ez = qs.p[2] - D.drone_target[2];
            const float sq = fmaf(ex, ex, fmaf(ey, ey, ez * ez));
            if (last) term_d = sq; else if (valid) Sd += sq;
        }
        if constexpr (MODEL == MPPI_MODEL_QUAD4) {
            f2 s2, c2;
            sincos_sel<FAST_TRIG>(f2(qs.rpy[0], qs.rpy[1]), s2, c2);
            qs.sphi = s2.v.x; qs.cphi = c2.v.x; qs.sth = s2.v.y; qs.cth = c2.v.y;
            sincos_sel<FAST_TRIG>(qs.rpy[2], qs.spsi, qs.cpsi);
        }
        if constexpr (HAS_ARM) {
            // arm accelerations in the joint pairing: chunk c1 = ARM0/4 holds joints 0..3, the next one joints 4..6
            constexpr int C1 = ARM0 / 4;
            const f2 aj[4] = {a02[C1], a13[C1], a02[C1 + 1], a13[C1 + 1]};
            // S/sampling/standard_normal_noise.py:32-50, two joints per instruction
            f2 qp[4], qt[4];                     // joint angles; the same less q0's whole turns (sin / cos arguments)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if constexpr (DIRECT) {
                    cum_q[i] = vfma(aj[i], f2(0.5f * P.dt2), vfma(cum_v[i], f2(P.dt), cum_q[i]));
                    cum_v[i] = vfma(aj[i], f2(P.dt), cum_v[i]);
                    qp[i] = qt[i] = cum_q[i];
                    continue;
                }
                const f2 vprev = vadd(cum_v[i], qd0p[i]);          // V_{t-1} (0 + qd0 at t = 0), rebuilt instead of carried
                const f2 dq = vfma(vprev, f2(P.dt), vmul(aj[i], f2(0.5f * P.dt2)));   // (0.5 a) dt^2 exactly: halving is exact
                cum_v[i] = vfma(aj[i], f2(P.dt), cum_v[i]);
                cum_q[i] = vadd(cum_q[i], dq);
                qp[i] = vadd(cum_q[i], q0p[i]);
                qt[i] = FAST_TRIG ? vadd(cum_q[i], q0t[i]) : qp[i];
            }
            if constexpr (EXTRA) {
                // covar_cost.py:20-25 (u^T Sigma^-1 v), action_cost.py:15-25, joint_space_cost.py:18-77; gamma^t discount
                f2 cov(0.f), act(0.f), cen(0.f), trk(0.f);
                bool out_of_bounds = false;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ja = pairA(i), jb = pairB(i);
                    const float *un = s_unom + t * NU + ARM0;
                    const f2 usig(un[ja] * P.inv_sigma_arm[ja], jb >= 0 ? un[jb] * P.inv_sigma_arm[jb] : 0.f);
                    cov = vfma(usig, aj[i], cov);
                    act = vfma(aj[i], aj[i], act);
                    const f2 dc = vsub(qp[i], f2(P.q_center[ja], jb >= 0 ? P.q_center[jb] : 0.f));
                    cen = vfma(dc, dc, cen);
                    const f2 qt(q_traj ? __ldg(q_traj + t * 7 + ja) : 0.f, (q_traj && jb >= 0) ? __ldg(q_traj + t * 7 + jb) : 0.f);
                    const f2 dk = vsub(qp[i], qt);
                    trk = vfma(dk, dk, trk);
                    out_of_bounds = out_of_bounds || qp[i].v.x < P.q_lower[ja] || qp[i].v.x > P.q_upper[ja];
                    if (jb >= 0) out_of_bounds = out_of_bounds || qp[i].v.y < P.q_lower[jb] || qp[i].v.y > P.q_upper[jb];
                }
                // the spare lane of pair 3 carries zeros in aj / (q - 0) terms: exclude it from the q-dependent sums
                x_cov += cov.v.x + cov.v.y;
                x_act = fmaf(gpow, act.v.x + act.v.y, x_act);
                const float pad_q = qp[3].v.y;        // = 0 + 0 (spare lane): (0 - q_center)^2 would be 0 for both tables
                (void)pad_q;
                x_cen = fmaf(gpow, cen.v.x + cen.v.y, x_cen);
                x_trk = fmaf(gpow, trk.v.x + trk.v.y, x_trk);
                if (out_of_bounds) x_lim = fmaf(gpow, P.limit_penalty, x_lim);
                gpow *= P.gamma;
            }
            // sin / cos of the seven joint angles (and of the new Euler angles for the whole body), packed
            float cq[7], sq[7];
            f2 s2, c2;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                sincos_sel<FAST_TRIG>(qt[i], s2, c2);
                sq[pairA(i)] = s2.v.x; cq[pairA(i)] = c2.v.x; sq[pairB(i)] = s2.v.y; cq[pairB(i)] = c2.v.y;
            }
            Pose3 Tp;
            if constexpr (MODEL == MPPI_MODEL_ARM7) {
                sincos_sel<FAST_TRIG>(qt[3].v.x, sq[5], cq[5]);
                Tp = base0;
            } else {
                sincos_sel<FAST_TRIG>(f2(qt[3].v.x, qs.rpy[0]), s2, c2);
                sq[5] = s2.v.x; cq[5] = c2.v.x; qs.sphi = s2.v.y; qs.cphi = c2.v.y;
                sincos_sel<FAST_TRIG>(f2(qs.rpy[1], qs.rpy[2]), s2, c2);
                qs.sth = s2.v.x; qs.cth = c2.v.x; qs.spsi = s2.v.y; qs.cpsi = c2.v.y;
                // moving base T(p_t, rpy_t) (S/robot/transformation_matrix.py:148-187)
                pose3_rpy(Tp, qs.sphi, qs.cphi, qs.sth, qs.cth, qs.spsi, qs.cpsi);
                Tp.pxy = f2(qs.p[0], qs.p[1]); Tp.pz = qs.p[2];
                if constexpr (BAKED) pose3_compose_tab<FkKinova, 0>(Tp);
                else pose3_compose_const(Tp, P.chain.R[0], P.chain.t[0]);
            }
            if constexpr (BAKED) {
                pose3_fk_tab<FkKinova>(cq, sq, Tp);
            } else {
                float qv[7];
#pragma unroll
                for (int i = 0; i < 4; ++i) { qv[pairA(i)] = qp[i].v.x; if (pairB(i) >= 0) qv[pairB(i)] = qp[i].v.y; }
                pose3_fk_chain<7>(P.chain, qv, cq, sq, Tp);
            }
            float pos, ori;
            pose3_terms(Tp, D, pos, ori);
            // S/cost/cost_manager.py:30-33,78-89
            const float c = last ? fmaf(P.cost_w[2], pos, P.cost_w[3] * ori)
                                 : fmaf(P.cost_w[0], pos, P.cost_w[1] * ori);
            const float y = c - comp;
            const float tS = S + y;
            comp = (tS - S) - y;
            S = tS;
        }
      }
    }
    // programmatic dependent launch: the weighting kernel of this step may start its launch now; it still waits
    // (griddepcontrol.wait) for this whole grid and its memory before it reads S or rho
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if constexpr (MODEL == MPPI_MODEL_DRONE3 || MODEL == MPPI_MODEL_QUAD4) {
        S = fmaf(Sd, P.cost_w[4], term_d * P.cost_w[5]);
    } else if constexpr (MODEL == MPPI_MODEL_WB11) {
        S = S + fmaf(Sd, P.cost_w[4], term_d * P.cost_w[5]);
    }
    if constexpr (EXTRA && HAS_ARM) {
        // same order as the commented-out sum, cost_manager.py:83-87
        if (P.cost_flags & MPPI_COST_COVAR) S += P.covar_scale * x_cov;
        if (P.cost_flags & MPPI_COST_CENTERING) S += P.centering_weight * x_cen;
        if (P.cost_flags & MPPI_COST_JOINT_TRAJ) S += P.joint_traj_weight * x_trk;
        if (P.cost_flags & MPPI_COST_ACTION) S += P.action_weight * x_act;
        if (P.cost_flags & MPPI_COST_JOINT_LIMIT) S += x_lim;
    }

    if (active) cost_out[k] = S;
    // ---- block minimum -> one atomicMin on the order-preserving encoding
    float m = warp_min(active ? S : __int_as_float(0x7f800000));
    if ((threadIdx.x & 31) == 0) s_wmin[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float bm = s_wmin[0];
#pragma unroll
        for (int w = 1; w < kRolloutThreads / 32; ++w) bm = fminf(bm, s_wmin[w]);
        atomicMin(rho_enc, encode_ordered(bm));
    }
    trace_stamp(D, TR_ROLLOUT_DONE, blockIdx.x == gridDim.x - 1 && threadIdx.x == 0);
    active_out = active;
    return S;
}

// 4 blocks/SM (<= 128 registers): measured best -- with a 72-register cap (7 blocks/SM) ptxas cannot interleave the
// Philox multiplies with the FK arithmetic and the FMA pipe stalls more (0.421 -> 0.393 ms on the bench case).
#ifndef MPPI_WB_MINB
#define MPPI_WB_MINB MPPI_ROLLOUT_MINB
#endif
template <int MODEL, int NOISE, bool BAKED, bool EXTRA, int ROUNDS>
__global__ void __launch_bounds__(kRolloutThreads, (MODEL == MPPI_MODEL_WB11 && NOISE == 0 && !EXTRA) ? MPPI_WB_MINB : MPPI_ROLLOUT_MINB)
rollout_cost_kernel(const __grid_constant__ StepParams P, const __grid_constant__ DynBlock D,
                    const float *__restrict__ u_nom, const float *__restrict__ noise,
                    float *__restrict__ cost_out, int32_t *__restrict__ rho_enc,
                    const float *__restrict__ q_traj)
{
    bool active;
    (void)rollout_body<MODEL, NOISE, BAKED, EXTRA, ROUNDS>(P, D, u_nom, noise, cost_out, rho_enc, q_traj, active);
}

// ------------------------------------------------------------------------------------------
// K4: finalize (device function, run by one block).
// S/mppi_solver/mppi.py:148-158, drone_mppi.py:158-170, S/filter/svg_filter.py:13-90.
//   wsum = [T*nu] raw weighted-noise sums, eta, sum w^2     scratch = 2*T*nu + nu floats of smem
// ------------------------------------------------------------------------------------------
// exchange_ok == false (a peer never published its shard row: p2p_exchange gave up): the controls are NOT updated
// (u_new = u_nom), out[MPPI_OUT_STEP] = -1 is part of the published out vector, and the caller raises.
template <int MODEL>
__device__ void finalize_block(const StepParams &P, const DynBlock &D, const float *wsum,
                               const float *u_nom, float *u_new, float *out, int32_t *rho_enc,
                               float *scratch, bool exchange_ok = true)
{
    constexpr int NU = ModelNu<MODEL>::value;
    const int n = P.T * NU;
    float *raw = scratch, *un = scratch + n, *u0_old = scratch + 2 * n;
    const float eta = wsum[n], eta2 = wsum[n + 1];
    const float inv_eta = 1.0f / eta;
    for (int j = threadIdx.x; j < n; j += blockDim.x) raw[j] = wsum[j] * inv_eta;
    if (threadIdx.x < NU) u0_old[threadIdx.x] = u_nom[threadIdx.x];
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int t = j / NU, i = j - t * NU;
        float acc = 0.f;
        for (int jj = -P.sg_half; jj <= P.sg_half; ++jj) {
            int s = t + jj;
            s = (s < 0) ? (-s - 1) : s;                   // data[:h].flip(0)
            s = (s >= P.T) ? (2 * P.T - 1 - s) : s;       // data[-h:].flip(0)
            acc = fmaf(P.taps[jj + P.sg_half], raw[s * NU + i], acc);
        }
        const float v = exchange_ok ? u_nom[j] + acc : u_nom[j];      // u += w_eps
        un[j] = v;
        u_new[j] = v;
    }
    __syncthreads();
    trace_stamp(D, TR_CONTROLS_UPDATED, threadIdx.x == 0);
    // The epilogue is split over three warps so that its serial pieces overlap (this block is the tail of the step):
    //   warp 0 lane 0: controller outputs;  warp 1 lane 0: check_reach FK;  warp 2: u0 / statistics;  warp 3: torque law.
    const float dt = P.dt;
    if (out != nullptr && threadIdx.x == 0) {
        if constexpr (MODEL == MPPI_MODEL_DRONE3) {
            for (int i = 0; i < 3; ++i) {
                out[i] = D.state[i] + D.state[3 + i] * dt + 0.5f * un[i] * P.dt2;      // drone_mppi.py:170
                out[3 + i] = D.state[3 + i] + dt * un[i];                                // :169
            }
        } else if constexpr (MODEL == MPPI_MODEL_ARM7) {
            for (int i = 0; i < 7; ++i) {
                out[i] = D.state[i] + u0_old[i] * dt + 0.5f * un[i] * dt * dt;          // mppi.py:158 (F11)
                out[7 + i] = D.state[7 + i] + un[i] * dt;                                // mppi.py:157
            }
        } else {
            QuadState<float> qs;
            quad_load(qs, D.state);
            quad_advance<false>(qs, un[0], un[1], un[2], un[3], dt, P.quad);
            const int o = (MODEL == MPPI_MODEL_WB11) ? MPPI_OUT_BASE : 0;
            for (int i = 0; i < 3; ++i) { out[o + i] = qs.p[i]; out[o + 3 + i] = qs.rpy[i]; out[o + 6 + i] = qs.v[i]; out[o + 9 + i] = qs.w[i]; }
            if constexpr (MODEL == MPPI_MODEL_WB11) {
                for (int i = 0; i < 7; ++i) {
                    out[i] = D.state[12 + i] + u0_old[4 + i] * dt + 0.5f * un[4 + i] * dt * dt;
                    out[7 + i] = D.state[19 + i] + un[4 + i] * dt;
                }
            }
        }
    }
    if constexpr (MODEL == MPPI_MODEL_ARM7 || MODEL == MPPI_MODEL_WB11) {
        if (out != nullptr && threadIdx.x == 32) {
            // check_reach (mppi.py:95-120): L1 position error of FK(base, qdes) to the target
            constexpr int A0 = (MODEL == MPPI_MODEL_WB11) ? 4 : 0, Q0 = (MODEL == MPPI_MODEL_WB11) ? 12 : 0;
            float qv[7], cq[7], sq[7], R[9], p[3];
            for (int i = 0; i < 7; ++i) {
                qv[i] = D.state[Q0 + i] + u0_old[A0 + i] * dt + 0.5f * un[A0 + i] * dt * dt;
                sincos_pi(qv[i], sq[i], cq[i]);
            }
            if constexpr (MODEL == MPPI_MODEL_ARM7) {
                quat_matrix(&D.state[14], R);
                p[0] = D.state[14]; p[1] = D.state[15]; p[2] = D.state[16];
            } else {
                float sr, cr, sp, cp, sy, cy;
                sincos_pi(D.state[3], sr, cr); sincos_pi(D.state[4], sp, cp); sincos_pi(D.state[5], sy, cy);
                rpy_matrix(sr, cr, sp, cp, sy, cy, R);
                p[0] = D.state[0]; p[1] = D.state[1]; p[2] = D.state[2];
            }
            Pose3 Tp = pose3_from(R, p);
            if (P.chain.baked) {
                pose3_compose_tab<FkKinova, 0>(Tp);
                pose3_fk_tab<FkKinova>(cq, sq, Tp);
            } else {
                pose3_compose_const(Tp, P.chain.R[0], P.chain.t[0]);
                pose3_fk_chain<7>(P.chain, qv, cq, sq, Tp);
            }
            out[MPPI_OUT_REACH] = fabsf(Tp.pxy.v.x - D.target_pos[0]) + fabsf(Tp.pxy.v.y - D.target_pos[1]) + fabsf(Tp.pz - D.target_pos[2]);
        }
    }
    if constexpr (MODEL == MPPI_MODEL_ARM7 || MODEL == MPPI_MODEL_WB11) {
        // computed-torque law of the arm node (kinova.py:184) on a fourth warp: 8 Newton-Euler passes on 8 lanes
        constexpr int A0 = (MODEL == MPPI_MODEL_WB11) ? 4 : 0;
        const int tw = (blockDim.x >= 128) ? 3 : 0;
        if (out != nullptr && (P.cost_flags & MPPI_OPT_TORQUE_LAW) && P.chain.prismatic == 0 && (threadIdx.x >> 5) == tw) {
            float dq[7];                     // qdes - q without cancellation: u0_old dt + u0_new dt^2 / 2 (mppi.py:158)
            for (int i = 0; i < 7; ++i) dq[i] = u0_old[A0 + i] * dt + 0.5f * un[A0 + i] * dt * dt;
            if constexpr (MODEL == MPPI_MODEL_ARM7) arm_torque_from_arm_state(P, D.state, dq, out + MPPI_OUT_TORQUE);
            else arm_torque_from_wb_state(P, D.state, dq, out + MPPI_OUT_TORQUE);
        }
    }
    if (out != nullptr && threadIdx.x >= 64 && threadIdx.x < 96) {
        const int l = threadIdx.x - 64;
        if (l < NU) { out[MPPI_OUT_U0_NEW + l] = un[l]; out[MPPI_OUT_U0_OLD + l] = u0_old[l]; }
        if (l == 16) out[MPPI_OUT_RHO] = decode_ordered(*rho_enc);
        if (l == 17) out[MPPI_OUT_ETA] = eta;
        if (l == 18) out[MPPI_OUT_ESS] = eta * eta / eta2;
        if (l == 19) out[MPPI_OUT_STEP] = exchange_ok ? static_cast<float>(D.step_lo & 0xffffffu) : -1.0f;
    }
    __syncthreads();
    trace_stamp(D, TR_END, threadIdx.x == 0);
    if (threadIdx.x == 0) *rho_enc = kRhoInit;          // re-arm the minimum for the next step
    if (out != nullptr && D.host_out != nullptr) {
        // zero-copy result for the blocking host call: out[] (complete after the barrier above) -> mapped host memory,
        // system-scope fence, then the sequence word the host is spinning on
        if (threadIdx.x < MPPI_OUT_FLOATS) D.host_out[threadIdx.x] = out[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned *>(D.host_out + MPPI_OUT_FLOATS) = D.host_seq;
        }
    }
}

template <int MODEL>
__global__ void __launch_bounds__(256)
finalize_kernel(const __grid_constant__ StepParams P, const __grid_constant__ DynBlock D,
                const float *__restrict__ wsum, const float *u_nom, float *u_new, float *out,
                int32_t *rho_enc)
{
    extern __shared__ __align__(16) float s_fin[];
    finalize_block<MODEL>(P, D, wsum, u_nom, u_new, out, rho_enc, s_fin);
}

// One-shot exchange of the shard results over NVLink peer memory, run by ONE block per rank inside the step's last
// kernel (no NCCL call, no extra launch).  Every rank weights with its LOCAL cost minimum rho_r; rows are combined
// with c_r = exp(-(rho_r - rho)/lambda), which is algebraically the allreduce-MIN + allreduce-SUM of the two-collective
// contract (wsum = sum_r c_r wsum_r).  All ranks sum in rank order, so replicas stay bit-identical.
//
// Low-latency protocol: every element travels as ONE naturally aligned 64-bit store {value, epoch} straight into the
// peers' inboxes, and the receiver polls the element itself until its tag is the current epoch -- one NVLink one-way
// latency, no fence, no separate flag (a 64-bit scalar access is single-copy atomic; this is the scheme of NCCL's LL
// protocol).  The first version (rows, system fence, release flag, acquire spin, then reads) cost two round trips:
// measured 13.3 us per step at 8 GPUs; with the fence folded into the release and the reads batched 9.0 us.
//   inbox slot (parity = epoch & 1, source r, element j); two parities are enough because a rank cannot publish
//   epoch e+2 before every peer has published e+1, i.e. after every peer finished reading e.
__device__ __forceinline__ void st_ll(unsigned long long *p, float v, unsigned tag)
{
    const unsigned long long w = (static_cast<unsigned long long>(tag) << 32) | static_cast<unsigned long long>(__float_as_uint(v));
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_ll(const unsigned long long *p)
{
    unsigned long long w;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ unsigned long long *p2p_ll_slot(float *base, int world, int rowp, int parity, int src)
{
    return reinterpret_cast<unsigned long long *>(base + kMaxRanks * kFlagStrideInts) + (static_cast<size_t>(parity) * world + src) * rowp;
}

template <int MODEL>
__device__ bool p2p_exchange(const StepParams &P, const P2PParams &X, float *wsum, int32_t *rho_enc)
{
    constexpr int NU = ModelNu<MODEL>::value;
    __shared__ float s_scale[kMaxRanks], s_rho[kMaxRanks];
    __shared__ int s_ok;
    const int row = P.T * NU + 2;                 // elements 0 .. row-1 = sums, eta, sum w^2; element `row` = this rank's minimum
    const int parity = static_cast<int>(X.epoch & 1u);
    const int tid = threadIdx.x;
    if (tid == 0) s_ok = 1;
    const float my_rho = decode_ordered(*rho_enc);
    __syncthreads();
    // ---- publish: element j of this rank's row into slot (parity, rank, j) of EVERY rank's inbox (own included)
    for (int j = tid; j <= row; j += blockDim.x) {
        const float v = (j < row) ? wsum[j] : my_rho;
        for (int dst = 0; dst < X.world; ++dst)
            st_ll(p2p_ll_slot(X.base[dst], X.world, X.rowp, parity, X.rank) + j, v, X.epoch);
    }
    // ---- receive: poll the own inbox element by element.  First the sources' minima (element `row`), so that every
    // thread can scale and add its own elements as they arrive, in rank order (same order, same values on every rank:
    // replicas stay bit-identical).
    float *mine = X.base[X.rank];
    const long long t0 = clock64();
    if (tid < X.world) {
        const unsigned long long *slot = p2p_ll_slot(mine, X.world, X.rowp, parity, tid) + row;
        unsigned long long w = ld_ll(slot);
        while (static_cast<unsigned>(w >> 32) != X.epoch) {
            if (clock64() - t0 > 4000000000LL) { s_ok = 0; break; }     // ~2 s: a peer is gone; give up instead of hanging
            w = ld_ll(slot);
        }
        s_rho[tid] = __uint_as_float(static_cast<unsigned>(w));
    }
    __syncthreads();
    float rho = s_rho[0];
    for (int r = 1; r < X.world; ++r) rho = fminf(rho, s_rho[r]);
    if (tid < X.world) s_scale[tid] = expf(-P.inv_lambda * (s_rho[tid] - rho));
    if (tid == 0) *rho_enc = encode_ordered(rho);
    __syncthreads();
    for (int j = tid; j < row; j += blockDim.x) {
        // all sources' loads of this element in flight at once (one L2 latency), then re-poll only the late ones
        unsigned long long w[kMaxRanks];
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r)
            w[r] = (r < X.world) ? ld_ll(p2p_ll_slot(mine, X.world, X.rowp, parity, r) + j) : 0ull;
        bool pending = true;
        while (pending) {
            pending = false;
#pragma unroll
            for (int r = 0; r < kMaxRanks; ++r) {
                if (r < X.world && static_cast<unsigned>(w[r] >> 32) != X.epoch) {
                    w[r] = ld_ll(p2p_ll_slot(mine, X.world, X.rowp, parity, r) + j);
                    pending = true;
                }
            }
            if (pending && clock64() - t0 > 4000000000LL) { s_ok = 0; break; }
        }
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r) {
            if (r < X.world) {
                const float c = s_scale[r];
                acc = fmaf((j == row - 1) ? c * c : c, __uint_as_float(static_cast<unsigned>(w[r])), acc);      // last entry is sum w^2
            }
        }
        wsum[j] = acc;
    }
    __syncthreads();
    if (s_ok == 0 && tid == 0 && X.fail_flag != nullptr) {       // sticky, host-visible: the next call on this handle fails
        *reinterpret_cast<volatile unsigned *>(X.fail_flag) = X.epoch;
        __threadfence_system();
    }
    return s_ok != 0;
}

// Last-block-done: every block publishes its partial row, the last one to arrive sums the rows
// in index order (deterministic) into wsum, exchanges with the peer shards when asked to, and
// optionally finalizes.
template <int MODEL>
__device__ void reduce_partials_and_finalize(const StepParams &P, const DynBlock &D, const float *part,
                                             int n_parts, uint32_t *counter, float *wsum, bool fuse,
                                             const float *u_nom, float *u_new, float *out,
                                             int32_t *rho_enc, float *scratch, const P2PParams &X,
                                             const float *eta_part = nullptr, int n_eta = 0,
                                             unsigned long long *fix = nullptr, bool fix_has_sigma = false)
{
    constexpr int NU = ModelNu<MODEL>::value;
    __shared__ bool s_last;
    const int row = P.T * NU + 2;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t total = gridDim.x * gridDim.y;
        s_last = (atomicAdd(counter, 1u) == total - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    trace_stamp(D, TR_LAST_BLOCK, threadIdx.x == 0);
    // When this block also finalizes, the sums and the nominal controls reach finalize_block through shared memory
    // (behind its own scratch): re-reading what was just stored to global costs an L2 round trip (~0.7 us) each on
    // this serial tail.  The global copy is still written (allreduce contract, observability).
    const int n_ctl = P.T * NU;
    float *s_w = scratch + ((2 * n_ctl + NU + 3) & ~3), *s_u = s_w + ((row + 3) & ~3);
    if (fuse)
        for (int j = threadIdx.x; j < n_ctl; j += blockDim.x) s_u[j] = __ldg(u_nom + j);
    if (fix != nullptr) {
        // fixed-point accumulators -> float sums (sigma applied here), and re-arm them for the next step
        for (int j = threadIdx.x; j < row; j += blockDim.x) {
            const long long q = static_cast<long long>(__ldcg(fix + j));
            fix[j] = 0ull;
            const float scale = (j < row - 2 && !fix_has_sigma) ? P.sigma[j % NU] * (1.0f / kFixScale) : (1.0f / kFixScale);
            const float v = static_cast<float>(q) * scale;
            wsum[j] = v;
            if (fuse) s_w[j] = v;
        }
    } else
    for (int j = threadIdx.x; j < row; j += blockDim.x) {
        // eight loads in flight per thread; rows are always added in the same order (deterministic)
        float acc = 0.f;
        int b = 0;
        for (; b + 8 <= n_parts; b += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(part + static_cast<size_t>(b + u) * row + j);
            acc += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        }
        for (; b < n_parts; ++b) acc += __ldcg(part + static_cast<size_t>(b) * row + j);
        if (eta_part != nullptr && j >= row - 2) {          // eta / sum w^2 come from the weights kernel
            acc = 0.f;
            for (int e = 0; e < n_eta; ++e) acc += __ldcg(eta_part + 2 * e + (j - (row - 2)));
        }
        wsum[j] = acc;
        if (fuse) s_w[j] = acc;
    }
    if (threadIdx.x == 0) *counter = 0u;
    __syncthreads();
    trace_stamp(D, TR_REDUCED, threadIdx.x == 0);
    bool ok = true;
    if (!fuse) {
        if (X.world > 1) ok = p2p_exchange<MODEL>(P, X, wsum, rho_enc);
        return;
    }
    if (X.world > 1) ok = p2p_exchange<MODEL>(P, X, s_w, rho_enc);
    trace_stamp(D, TR_EXCHANGED, threadIdx.x == 0);
    finalize_block<MODEL>(P, D, s_w, s_u, u_new, out, rho_enc, scratch, ok);
}

// ------------------------------------------------------------------------------------------
// K3 (Philox): thread = horizon step t x R sub-ranges of the block's sample chunk; the noise of
// (sample, t) is regenerated from its address exactly as the rollout kernel generates it (all Philox
// calls of the step in one thread, packed Box-Muller on the pairs that are needed: 4 of 6 for nu = 7),
// zero-weight samples are skipped (exact: w == 0.0f contributes 0).
// S/mppi_solver/mppi.py:143-148,173-193.
// ------------------------------------------------------------------------------------------
template <int MODEL, int ROUNDS>
__global__ void __launch_bounds__(1024)
weight_philox_kernel(const __grid_constant__ StepParams P, const __grid_constant__ DynBlock D,
                     const float *__restrict__ S, int32_t *rho_enc, int chunk,
                     unsigned long long *__restrict__ fix, uint32_t *counter, float *wsum, int fuse,
                     const float *u_nom, float *u_new, float *out, const __grid_constant__ P2PParams X)
{
    constexpr int NU = ModelNu<MODEL>::value;
    constexpr int NQ = (NU + 3) / 4;                  // quads of normals per (sample, step), as in the rollout kernel
    constexpr int NUP = 4 * NQ;
    extern __shared__ __align__(16) float s_dyn[];      // [kWeightTile] weights | [kWeightTile] indices | reduction / finalize scratch
    float *s_w = s_dyn;
    int *s_idx = reinterpret_cast<int *>(s_dyn + kWeightTile);
    float *s_red = s_dyn + 2 * kWeightTile;
    __shared__ float s_eta[32], s_eta2[32];
    __shared__ int s_cnt[32];

    const int TC = P.T;
    const int R = blockDim.x / TC;
    const int tid = threadIdx.x;
    const bool worker = tid < R * TC;
    const int r = tid / TC, t = tid - r * TC;
    asm volatile("griddepcontrol.wait;" ::: "memory");      // PDL: launched while the rollout kernel drains (no-op otherwise)
    trace_stamp(D, TR_WEIGHT_START, blockIdx.x == 0 && tid == 0);
    const float rho = decode_ordered(*rho_enc);
    const int k0 = blockIdx.x * chunk;
    const int k1 = min(P.K, k0 + chunk);

    float acc[NUP];
#pragma unroll
    for (int i = 0; i < NUP; ++i) acc[i] = 0.f;
    float eta = 0.f, eta2 = 0.f;

    for (int base = k0; base < k1; base += kWeightTile) {
        const int nt = min(kWeightTile, k1 - base);
        __syncthreads();
        // weights of the tile + an ORDER-PRESERVING compaction of the samples whose weight is not exactly
        // zero (ballot + per-warp offsets), so the noise loop below only visits samples that contribute.
        // With lambda << cost spread (the drone) that is a handful of samples.
        int n_nz = 0;
        for (int pass = 0; pass < nt; pass += blockDim.x) {
            const int kk = pass + tid;
            float w = 0.f;
            if (kk < nt) {
                w = expf(-P.inv_lambda * (S[base + kk] - rho));
                eta += w;
                eta2 = fmaf(w, w, eta2);
            }
            const unsigned ball = __ballot_sync(0xffffffffu, w != 0.f);
            if ((tid & 31) == 0) s_cnt[tid >> 5] = __popc(ball);
            __syncthreads();
            int warp_off = 0, total = 0;
            const int nw = (blockDim.x + 31) >> 5;
            for (int wv = 0; wv < nw; ++wv) {
                const int cw = s_cnt[wv];
                if (wv < (tid >> 5)) warp_off += cw;
                total += cw;
            }
            if (w != 0.f) {
                const int slot = n_nz + warp_off + __popc(ball & ((1u << (tid & 31)) - 1u));
                s_idx[slot] = kk;
                s_w[slot] = w;
            }
            n_nz += total;
            __syncthreads();
        }
        if (worker) {
            for (int j = r; j < n_nz; j += R) {
                const float w = s_w[j];
                float unif[6 * philox_calls(NU)];
                philox_step_uniforms<philox_calls(NU), ROUNDS>(static_cast<uint32_t>(P.k_offset + base + s_idx[j]), static_cast<uint32_t>(t),
                                                               D.step_lo, D.step_hi, P.rkeys, unif);
#pragma unroll
                for (int e = 0; e < NQ; ++e) {         // sigma is applied once, after the reduction
                    f2 n02, n13;
                    normals_quad(unif, e, n02, n13);
                    acc[4 * e] = fmaf(w, n02.v.x, acc[4 * e]);
                    acc[4 * e + 1] = fmaf(w, n13.v.x, acc[4 * e + 1]);
                    acc[4 * e + 2] = fmaf(w, n02.v.y, acc[4 * e + 2]);
                    acc[4 * e + 3] = fmaf(w, n13.v.y, acc[4 * e + 3]);
                }
            }
        }
    }
    // ---- block reductions (fixed order): eta / eta2 over all threads, acc over the R sub-ranges
    eta = warp_sum(eta);
    eta2 = warp_sum(eta2);
    if ((tid & 31) == 0) { s_eta[tid >> 5] = eta; s_eta2[tid >> 5] = eta2; }
    __syncthreads();
    if (worker) {
#pragma unroll
        for (int i = 0; i < NUP; ++i) s_red[(r * TC + t) * NUP + i] = acc[i];
    }
    __syncthreads();
    // Block sums go into 64-bit FIXED-POINT accumulators with integer atomics: integer addition is
    // associative, so the result is bit-reproducible whatever the block order, and the last block has
    // nothing left to reduce (a float row-per-block scheme costs ~15 us of serial L2 round trips here).
    // |sum w n| <= K * 5.7 < 2^29 for K <= 2^26, scale 2^33 -> below 2^62; resolution 1.2e-10.
    const int row = P.T * NU + 2;
    for (int o = tid; o < TC * NU; o += blockDim.x) {          // one output (t, i) per thread
        const int tt = o / NU, i = o - tt * NU;
        float v = 0.f;
        for (int rr = 0; rr < R; ++rr) v += s_red[(rr * TC + tt) * NUP + i];
        if (v != 0.f) atomicAdd(fix + o, static_cast<unsigned long long>(__float2ll_rn(v * kFixScale)));
    }
    if (tid == 0) {
        float e = 0.f, e2 = 0.f;
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nw; ++w) { e += s_eta[w]; e2 += s_eta2[w]; }
        if (e != 0.f) atomicAdd(fix + row - 2, static_cast<unsigned long long>(__float2ll_rn(e * kFixScale)));
        if (e2 != 0.f) atomicAdd(fix + row - 1, static_cast<unsigned long long>(__float2ll_rn(e2 * kFixScale)));
    }
    trace_stamp(D, TR_SUMS_ADDED, blockIdx.x == 0 && tid == 0);
    reduce_partials_and_finalize<MODEL>(P, D, nullptr, 0, counter, wsum, fuse != 0, u_nom, u_new, out,
                                        rho_enc, s_dyn, X, nullptr, 0, fix);
}

// ------------------------------------------------------------------------------------------
// Single-launch control step for grids that are co-resident (cooperative launch; K_local <= 128 x resident blocks,
// i.e. the per-GPU shard of the 8-GPU configuration and every reference-sized problem): rollout, ONE grid-wide
// barrier (the cost minimum is then final), the weighting of the block's own samples with the noise regenerated
// for the survivors, fixed-point atomics, and the exchange + finalize in the last block to arrive.  Replaces the
// rollout -> weighting launch pair (and its ~15 us of fixed cost: second launch, S re-read, 1024-thread blocks
// re-deriving what this block already holds in registers).  S/mppi_solver/mppi.py:122-158.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Monotonic arrival counter; `target` is the value it reaches when every block of THIS launch has arrived (the host
// adds gridDim to it per launch; the signed difference makes the wrap-around harmless).
__device__ __forceinline__ void grid_barrier(unsigned *ctr, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (static_cast<int>(ld_acquire_gpu_u32(ctr) - target) < 0) { }
    }
    __syncthreads();
}

template <int MODEL, bool BAKED, bool EXTRA, int ROUNDS>
__global__ void __launch_bounds__(kRolloutThreads, MPPI_ROLLOUT_MINB)
step_fused_kernel(const __grid_constant__ StepParams P, const __grid_constant__ DynBlock D,
                  const float *__restrict__ u_nom, float *__restrict__ cost_out, int32_t *rho_enc,
                  const float *__restrict__ q_traj, unsigned long long *__restrict__ fix, uint32_t *counter,
                  unsigned *sync_ctr, unsigned sync_target, float *wsum, float *u_new, float *out,
                  const __grid_constant__ P2PParams X)
{
    constexpr int NU = ModelNu<MODEL>::value;
    constexpr int NQ = (NU + 3) / 4;
    constexpr int NUP = 4 * NQ;
    constexpr int NW = kRolloutThreads / 32;
    extern __shared__ __align__(16) float s_dyn[];      // rollout: u_nom | then: weights, indices, reduction, finalize scratch
    __shared__ float s_eta[NW], s_eta2[NW];
    __shared__ int s_cnt[NW];

    bool active;
    const float S = rollout_body<MODEL, 0, BAKED, EXTRA, ROUNDS>(P, D, u_nom, nullptr, cost_out, rho_enc, q_traj, active);
    grid_barrier(sync_ctr, sync_target);                 // every block has folded its minimum into *rho_enc
    const float rho = decode_ordered(__ldcg(rho_enc));
    trace_stamp(D, TR_MIN_KNOWN, blockIdx.x == 0 && threadIdx.x == 0);

    float *s_w = s_dyn;
    int *s_idx = reinterpret_cast<int *>(s_dyn + kRolloutThreads);
    float *s_red = s_dyn + 2 * kRolloutThreads;
    const int tid = threadIdx.x;
    const int TC = P.T;                                  // <= kRolloutThreads (checked by the launcher)
    const int R = kRolloutThreads / TC;
    const bool worker = tid < R * TC;
    const int r = tid / TC, t = tid - r * TC;

    // weights of the block's own samples + order-preserving compaction of the non-zero ones
    const float w = active ? expf(-P.inv_lambda * (S - rho)) : 0.f;
    float eta = warp_sum(w), eta2 = warp_sum(w * w);
    const unsigned ball = __ballot_sync(0xffffffffu, w != 0.f);
    if ((tid & 31) == 0) { s_cnt[tid >> 5] = __popc(ball); s_eta[tid >> 5] = eta; s_eta2[tid >> 5] = eta2; }
    __syncthreads();
    int warp_off = 0, n_nz = 0;
#pragma unroll
    for (int wv = 0; wv < NW; ++wv) {
        if (wv < (tid >> 5)) warp_off += s_cnt[wv];
        n_nz += s_cnt[wv];
    }
    if (w != 0.f) {
        const int slot = warp_off + __popc(ball & ((1u << (tid & 31)) - 1u));
        s_idx[slot] = tid;
        s_w[slot] = w;
    }
    __syncthreads();
    float acc[NUP];
#pragma unroll
    for (int i = 0; i < NUP; ++i) acc[i] = 0.f;
    if (worker) {
        const uint32_t kg0 = static_cast<uint32_t>(P.k_offset + blockIdx.x * kRolloutThreads);
        for (int j = r; j < n_nz; j += R) {
            const float wj = s_w[j];
            float unif[6 * philox_calls(NU)];
            philox_step_uniforms<philox_calls(NU), ROUNDS>(kg0 + static_cast<uint32_t>(s_idx[j]), static_cast<uint32_t>(t),
                                                           D.step_lo, D.step_hi, P.rkeys, unif);
#pragma unroll
            for (int e = 0; e < NQ; ++e) {             // sigma is applied once, after the reduction
                f2 n02, n13;
                normals_quad(unif, e, n02, n13);
                acc[4 * e] = fmaf(wj, n02.v.x, acc[4 * e]);
                acc[4 * e + 1] = fmaf(wj, n13.v.x, acc[4 * e + 1]);
                acc[4 * e + 2] = fmaf(wj, n02.v.y, acc[4 * e + 2]);
                acc[4 * e + 3] = fmaf(wj, n13.v.y, acc[4 * e + 3]);
            }
        }
#pragma unroll
        for (int i = 0; i < NUP; ++i) s_red[(r * TC + t) * NUP + i] = acc[i];
    }
    __syncthreads();
    const int row = P.T * NU + 2;
    if (n_nz > 0) {
        for (int o = tid; o < TC * NU; o += kRolloutThreads) {
            const int tt = o / NU, i = o - tt * NU;
            float v = 0.f;
            for (int rr = 0; rr < R; ++rr) v += s_red[(rr * TC + tt) * NUP + i];
            if (v != 0.f) atomicAdd(fix + o, static_cast<unsigned long long>(__float2ll_rn(v * kFixScale)));
        }
        if (tid == 0) {
            float e = 0.f, e2 = 0.f;
#pragma unroll
            for (int wv = 0; wv < NW; ++wv) { e += s_eta[wv]; e2 += s_eta2[wv]; }
            if (e != 0.f) atomicAdd(fix + row - 2, static_cast<unsigned long long>(__float2ll_rn(e * kFixScale)));
            if (e2 != 0.f) atomicAdd(fix + row - 1, static_cast<unsigned long long>(__float2ll_rn(e2 * kFixScale)));
        }
    }
    trace_stamp(D, TR_SUMS_ADDED, blockIdx.x == 0 && tid == 0);
    reduce_partials_and_finalize<MODEL>(P, D, nullptr, 0, counter, wsum, true, u_nom, u_new, out, rho_enc, s_dyn, X,
                                        nullptr, 0, fix);
}

// ------------------------------------------------------------------------------------------
// Time-parallel control step for the models whose dynamics are a LINEAR double integrator (ARM7, DRONE3 -- the two
// controllers the reference runs): ONE WARP PER SAMPLE, lane l owns the SPL consecutive horizon steps
// [l*SPL, (l+1)*SPL) (SPL = 1, 2, 4 or 8: T <= 256).  The reference integrates with two cumulative sums over the horizon
// (S/sampling/standard_normal_noise.py:37-48, S/mppi_solver/drone_mppi.py:47-54); here they are two warp scans, after
// which every (sample, step) pair evaluates its FK + cost independently -- K*T-way instead of K-way parallelism, which
// is what a reference-sized problem (K = 100 ... 1000, T = 30) needs to fill 148 SMs: the thread-per-sample kernel
// runs it as 8 blocks of serial 30-step chains.  The lanes still hold their step's noise when the sample's cost is
// known, so the weighted-noise sum needs no second pass and no regeneration: each block keeps a running soft-min
// accumulator (block-local minimum, rescaled when a new minimum arrives) over the tiles it loops over, and publishes
// ONE row; the last block to arrive rescales the (<= one per SM) rows to the global minimum, adds them in block order
// (deterministic), exchanges with the peer shards (if any) and finalizes.  One ordinary launch per control step, no
// grid-wide barrier and no atomics on the sums (in-kernel timestamps, tools/trace_phases.py: 256 blocks adding 212
// 64-bit atomics each cost 3.6 us of same-address contention; 64 rows combined by one block cost ~1 us).
// NOISE: 0 = in-kernel Philox (same (k, t, call, step) addressing as every other kernel: bit-identical normals),
//        1 = injected [T][K][nu].
// S/mppi_solver/mppi.py:122-158, S/mppi_solver/drone_mppi.py:140-170.
// ------------------------------------------------------------------------------------------
// Block shape by steps per lane: 512 threads (16 samples per tile) up to four steps per lane; eight steps per lane
// (T <= 256) keep 8 x 15 values per thread across the cost evaluation and run as 256-thread blocks (<= 255 registers).
__host__ __device__ constexpr int tp_threads(int spl) { return spl <= 4 ? 512 : 256; }

__device__ __forceinline__ float warp_excl_scan(float x, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    const float e = __shfl_up_sync(0xffffffffu, x, 1);
    return lane ? e : 0.f;
}

// Sum over the warp with the rounding errors carried along (TwoSum butterfly): the result is the float32 nearest to
// the exact sum of the 32 inputs (to ~2^-45 relative) in every lane.  The arm's cost is ~1e3 with a spread of ~3 against
// lambda = 0.1, so one ulp of S moves a weight by 0.12 %: the serial kernel Kahan-compensates for the same reason.
__device__ __forceinline__ float warp_sum_compensated(float hi, float lo)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float bh = __shfl_xor_sync(0xffffffffu, hi, o), bl = __shfl_xor_sync(0xffffffffu, lo, o);
        const float s = __fadd_rn(hi, bh);
        const float bb = __fsub_rn(s, hi);
        const float e = __fadd_rn(__fsub_rn(hi, __fsub_rn(s, bb)), __fsub_rn(bh, bb));
        hi = s;
        lo = __fadd_rn(__fadd_rn(lo, bl), e);
    }
    return __fadd_rn(hi, lo);
}

constexpr int kTpMaxRows = 160;         // rows the last block combines: one per resident block, <= one block per SM
template <int MODEL, int NOISE, bool BAKED, int SPL, int ROUNDS>
__global__ void __launch_bounds__(tp_threads(SPL))
step_tp_kernel(const __grid_constant__ StepParams P, const __grid_constant__ DynBlock D,
               const float *__restrict__ u_nom, const float *__restrict__ noise, float *__restrict__ cost_out,
               int32_t *rho_enc, float *__restrict__ rows, float *__restrict__ rho_rows, uint32_t *counter,
               float *wsum, float *u_new, float *out, const __grid_constant__ P2PParams X)
{
    static_assert(MODEL == MPPI_MODEL_ARM7 || MODEL == MPPI_MODEL_DRONE3, "linear-integrator models only");
    constexpr int NU = ModelNu<MODEL>::value;
    constexpr int NQ = (NU + 3) / 4;
    constexpr int NUP = 4 * NQ;
    constexpr bool PHILOX = (NOISE == 0);
    constexpr int kTpThreads = tp_threads(SPL), kTpWarps = kTpThreads / 32;
    constexpr int Q0 = 0, QD0 = (MODEL == MPPI_MODEL_ARM7) ? 7 : 3;
    extern __shared__ __align__(16) float s_dyn[];       // [n] block accumulator | [kTpWarps][n] tile contributions (combine / finalize scratch later)
    __shared__ float s_S[kTpWarps], s_W[kTpWarps];
    __shared__ bool s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = P.T * NU;
    float *s_acc = s_dyn, *s_con = s_dyn + n;
    trace_stamp(D, TR_START, blockIdx.x == 0 && tid == 0);
    for (int j = tid; j < n; j += kTpThreads) s_acc[j] = 0.f;
    float rho_blk = __int_as_float(0x7f800000), eta_blk = 0.f, eta2_blk = 0.f;

    Pose3 base0;                                          // arm: base pose composed with C0 (uniform)
    if constexpr (MODEL == MPPI_MODEL_ARM7) {
        float R0[9], p0[3];
        quat_matrix(&D.state[14], R0);
        p0[0] = D.state[14]; p0[1] = D.state[15]; p0[2] = D.state[16];
        if constexpr (BAKED) compose_tab<FkKinova, 0>(R0, p0);
        else compose_const(R0, p0, P.chain.R[0], P.chain.t[0]);
        base0 = pose3_from(R0, p0);
    }
    __syncthreads();

    const int n_tiles = (P.K + kTpWarps - 1) / kTpWarps;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int k_raw = tile * kTpWarps + warp;
        const bool active = k_raw < P.K;                  // warp-uniform
        const int k = active ? k_raw : P.K - 1;
        const uint32_t kg = static_cast<uint32_t>(P.k_offset + k);

        // ---- noise and controls of this lane's steps: a = u_nom + eps  (mppi.py:130)
        float nrm[SPL][NUP];
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            const int t = lane * SPL + s;
            const bool valid = t < P.T;
            const int tc = valid ? t : P.T - 1;
            if constexpr (PHILOX) {
                float unif[6 * philox_calls(NU)];
                philox_step_uniforms<philox_calls(NU), ROUNDS>(kg, static_cast<uint32_t>(tc), D.step_lo, D.step_hi, P.rkeys, unif);
#pragma unroll
                for (int e = 0; e < NQ; ++e) {
                    f2 n02, n13;
                    normals_quad(unif, e, n02, n13);
                    nrm[s][4 * e] = n02.v.x; nrm[s][4 * e + 1] = n13.v.x; nrm[s][4 * e + 2] = n02.v.y; nrm[s][4 * e + 3] = n13.v.y;
                }
            } else {
                const float *row = noise + (static_cast<size_t>(tc) * P.K + k) * NU;
#pragma unroll
                for (int i = 0; i < NUP; ++i) nrm[s][i] = (i < NU) ? __ldg(row + i) : 0.f;
            }
            if (!valid) {
#pragma unroll
                for (int i = 0; i < NUP; ++i) nrm[s][i] = 0.f;
            }
        }
        // ---- double integrator as two scans over the horizon (standard_normal_noise.py:37-48 / drone_mppi.py:47-54)
        float qs[SPL][NU];                                // position-like state after each of this lane's steps
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            // controls of this lane's steps for input i, rebuilt from the normals (only the normals are kept across the
            // cost evaluation: registers scale with SPL)
            float a_i[SPL];
#pragma unroll
            for (int s = 0; s < SPL; ++s) {
                const int t = lane * SPL + s;
                const float eps = PHILOX ? __fmul_rn(P.sigma[i], nrm[s][i]) : nrm[s][i];
                a_i[s] = (t < P.T) ? __fadd_rn(__ldg(u_nom + t * NU + i), eps) : 0.f;
            }
            float lcv[SPL];
            float c = 0.f;
#pragma unroll
            for (int s = 0; s < SPL; ++s) { c = fmaf(a_i[s], P.dt, c); lcv[s] = c; }
            const float ex_v = warp_excl_scan(c, lane);   // sum a dt over all earlier lanes
            const float v0 = D.state[QD0 + i];
            float lcq[SPL];
            float cq = 0.f;
#pragma unroll
            for (int s = 0; s < SPL; ++s) {
                const float vprev = ((s == 0) ? ex_v : ex_v + lcv[s - 1]) + v0;
                cq += fmaf(vprev, P.dt, a_i[s] * (0.5f * P.dt2));
                lcq[s] = cq;
            }
            const float ex_q = warp_excl_scan(cq, lane);
#pragma unroll
            for (int s = 0; s < SPL; ++s) qs[s][i] = (ex_q + lcq[s]) + D.state[Q0 + i];
        }
        // ---- per-step cost, every (sample, step) pair on its own lane
        float S;
        if constexpr (MODEL == MPPI_MODEL_DRONE3) {
            float Sd = 0.f, term = 0.f;
#pragma unroll
            for (int s = 0; s < SPL; ++s) {
                const int t = lane * SPL + s;
                float sq = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) { const float e = qs[s][i] - D.drone_target[i]; sq = fmaf(e, e, sq); }
                if (t < P.T - 1) Sd += sq;
                else if (t == P.T - 1) term = sq;
            }
            Sd = warp_sum(Sd); term = warp_sum(term);
            S = fmaf(Sd, P.cost_w[4], term * P.cost_w[5]);          // drone_mppi.py:87-107
        } else {
            float csum = 0.f, ccomp = 0.f;                // Kahan over this lane's steps, TwoSum across the lanes
#pragma unroll
            for (int s = 0; s < SPL; ++s) {
                const int t = lane * SPL + s;
                float cq[7], sq[7];
                f2 s2, c2;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    sincos_pi(f2(qs[s][2 * i], qs[s][2 * i + 1]), s2, c2);
                    sq[2 * i] = s2.v.x; cq[2 * i] = c2.v.x; sq[2 * i + 1] = s2.v.y; cq[2 * i + 1] = c2.v.y;
                }
                sincos_pi(qs[s][6], sq[6], cq[6]);
                Pose3 Tp = base0;
                if constexpr (BAKED) pose3_fk_tab<FkKinova>(cq, sq, Tp);
                else pose3_fk_chain<7>(P.chain, qs[s], cq, sq, Tp);
                float pos, ori;
                pose3_terms(Tp, D, pos, ori);
                const float c = (t == P.T - 1) ? fmaf(P.cost_w[2], pos, P.cost_w[3] * ori)
                                               : fmaf(P.cost_w[0], pos, P.cost_w[1] * ori);      // cost_manager.py:30-33,78-89
                const float y = __fsub_rn((t < P.T) ? c : 0.f, ccomp);
                const float tS = __fadd_rn(csum, y);
                ccomp = __fsub_rn(__fsub_rn(tS, csum), y);
                csum = tS;
            }
            S = warp_sum_compensated(csum, -ccomp);
        }
        if (lane == 0 && active) cost_out[k] = S;

        // ---- running soft-min accumulation of the tile (fixed order: deterministic)
        if (lane == 0) s_S[warp] = active ? S : __int_as_float(0x7f800000);
        __syncthreads();
        float tmin = s_S[0];
#pragma unroll
        for (int wv = 1; wv < kTpWarps; ++wv) tmin = fminf(tmin, s_S[wv]);
        const float new_rho = fminf(rho_blk, tmin);
        const float c_old = (new_rho < rho_blk) ? expf(-P.inv_lambda * (rho_blk - new_rho)) : 1.0f;    // exp(-inf) = 0 on the first tile
        const float w = active ? expf(-P.inv_lambda * (S - new_rho)) : 0.f;
        if (lane == 0) s_W[warp] = w;
#pragma unroll
        for (int s = 0; s < SPL; ++s) {
            const int t = lane * SPL + s;
            if (t < P.T) {
#pragma unroll
                for (int i = 0; i < NU; ++i) s_con[warp * n + t * NU + i] = w * nrm[s][i];
            }
        }
        __syncthreads();
        for (int j = tid; j < n; j += kTpThreads) {
            float v = s_acc[j] * c_old;
#pragma unroll
            for (int wv = 0; wv < kTpWarps; ++wv) v += s_con[wv * n + j];
            s_acc[j] = v;
        }
        eta_blk *= c_old; eta2_blk *= c_old * c_old;
#pragma unroll
        for (int wv = 0; wv < kTpWarps; ++wv) { const float ww = s_W[wv]; eta_blk += ww; eta2_blk = fmaf(ww, ww, eta2_blk); }
        rho_blk = new_rho;
        __syncthreads();
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    trace_stamp(D, TR_ROLLOUT_DONE, blockIdx.x == 0 && tid == 0);
    // ---- publish this block's row (sums relative to its own minimum); the last block to arrive combines the rows
    const int nout = n + 2;
    const int rstride = (nout + 3) & ~3;                  // rows are read back as float4
    {
        float *my = rows + static_cast<size_t>(blockIdx.x) * rstride;
        for (int j = tid; j < n; j += kTpThreads) my[j] = s_acc[j];
        if (tid == 0) { my[n] = eta_blk; my[n + 1] = eta2_blk; rho_rows[blockIdx.x] = rho_blk; }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();                                  // release: cumulative over the block's stores ordered by the barrier above
        s_last = (atomicAdd(counter, 1u) == gridDim.x - 1u);
        if (s_last) __threadfence();                      // acquire side; the rows are then read with ld.cg (L2)
    }
    __syncthreads();
    trace_stamp(D, TR_SUMS_ADDED, blockIdx.x == 0 && tid == 0);
    if (!s_last) return;
    trace_stamp(D, TR_LAST_BLOCK, tid == 0);
    const int nb = gridDim.x;                             // <= kTpMaxRows
    // shared-memory layout of the hand-over: [0, 2n+nu) finalize scratch | row scales | combine partials | combined sums |
    // nominal controls.  The combined sums and u_nom reach finalize_block through shared memory: every global round
    // trip on this serial tail costs 0.5-1 us.
    float *s_scale = s_con, *s_red = s_dyn + ((n + kTpMaxRows + 3) & ~3);          // float4-aligned
    const int off_w = max(2 * n + NU, n + kTpMaxRows + 4 + max(4 * kTpThreads, rstride)) + 4;      // s_red holds parts * rstride floats
    float *s_w = s_dyn + off_w, *s_u = s_w + nout + 2;
    for (int j = tid; j < n; j += kTpThreads) s_u[j] = __ldg(u_nom + j);            // in flight while the rows are combined
    // global minimum over the rows' minima: one value per thread, warp shuffles, 16 partials
    float rmin = __int_as_float(0x7f800000);
    for (int b = tid; b < nb; b += kTpThreads) { const float r = __ldcg(rho_rows + b); s_scale[b] = r; rmin = fminf(rmin, r); }
    rmin = warp_min(rmin);
    if (lane == 0) s_S[warp] = rmin;
    __syncthreads();
    float rho = s_S[0];
#pragma unroll
    for (int wv = 1; wv < kTpWarps; ++wv) rho = fminf(rho, s_S[wv]);
    for (int b = tid; b < nb; b += kTpThreads) s_scale[b] = expf(-P.inv_lambda * (s_scale[b] - rho));
    __syncthreads();
    // rows are added in block order within each of `parts` interleaved sub-sequences, the sub-sequences in order;
    // a thread owns one float4 column of the rows: the loop is bound by the L2 latency (~0.7 us per dependent batch
    // measured), so it keeps 8 float4 loads in flight (hoisting them above the minimum, 16 deep, measured slower)
    const int c4 = rstride >> 2;
    const int parts = max(1, kTpThreads / c4);
    for (int o = tid; o < parts * c4; o += kTpThreads) {
        const int pt = o / c4, c = o - pt * c4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int e = n + 1 - 4 * c;                                      // component of this column holding sum w^2 (if 0..3)
        for (int b0 = pt; b0 < nb; b0 += 8 * parts) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int b = b0 + u * parts;
                v[u] = (b < nb) ? __ldcg(reinterpret_cast<const float4 *>(rows + static_cast<size_t>(b) * rstride) + c)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int b = b0 + u * parts;
                const float sc = (b < nb) ? s_scale[b] : 0.f;
                const float sq = sc * sc;
                acc.x = fmaf(e == 0 ? sq : sc, v[u].x, acc.x);
                acc.y = fmaf(e == 1 ? sq : sc, v[u].y, acc.y);
                acc.z = fmaf(e == 2 ? sq : sc, v[u].z, acc.z);
                acc.w = fmaf(e == 3 ? sq : sc, v[u].w, acc.w);
            }
        }
        reinterpret_cast<float4 *>(s_red)[o] = acc;
    }
    __syncthreads();
    for (int j = tid; j < nout; j += kTpThreads) {
        float v = 0.f;
        for (int pt = 0; pt < parts; ++pt) v += s_red[pt * rstride + j];
        v = (PHILOX && j < n) ? v * P.sigma[j % NU] : v;                 // unit normals were accumulated: sigma once, here
        s_w[j] = v;
        wsum[j] = v;                                                     // observability / allreduce contract; not re-read here
    }
    if (tid == 0) { *rho_enc = encode_ordered(rho); *counter = 0u; }
    __syncthreads();
    trace_stamp(D, TR_REDUCED, tid == 0);
    bool ok = true;
    if (X.world > 1) ok = p2p_exchange<MODEL>(P, X, s_w, rho_enc);
    trace_stamp(D, TR_EXCHANGED, tid == 0);
    finalize_block<MODEL>(P, D, s_w, s_u, u_new, out, rho_enc, s_dyn, ok);
}

// ------------------------------------------------------------------------------------------
// K3a (injected-noise path): w[k] = exp((rho - S[k]) / lambda) once, plus deterministic partial
// sums of w and w^2.  S/mppi_solver/mppi.py:173-193.
// ------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
weights_kernel(const __grid_constant__ StepParams P, const float *__restrict__ S, const int32_t *__restrict__ rho_enc,
               float *__restrict__ w, float *__restrict__ eta_part)
{
    __shared__ float s_e[8], s_e2[8];
    asm volatile("griddepcontrol.wait;" ::: "memory");      // PDL: launched while the rollout kernel drains (no-op otherwise)
    const float rho = decode_ordered(*rho_enc);
    float eta = 0.f, eta2 = 0.f;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < P.K; k += gridDim.x * blockDim.x) {
        const float v = expf(-P.inv_lambda * (S[k] - rho));
        w[k] = v;
        eta += v;
        eta2 = fmaf(v, v, eta2);
    }
    eta = warp_sum(eta);
    eta2 = warp_sum(eta2);
    if ((threadIdx.x & 31) == 0) { s_e[threadIdx.x >> 5] = eta; s_e2[threadIdx.x >> 5] = eta2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float e = 0.f, e2 = 0.f;
        for (int i = 0; i < 8; ++i) { e += s_e[i]; e2 += s_e2[i]; }
        eta_part[2 * blockIdx.x] = e;
        eta_part[2 * blockIdx.x + 1] = e2;
    }
}

// ------------------------------------------------------------------------------------------
// K3b (injected noise [T][K][nu]): the HBM-bound re-read, a streaming GEMV per horizon step:
//   out[t][i] = sum_k w[k] * noise[t][k][i].
// grid = (G, T) sized to ONE resident wave; block (g, t) streams its contiguous K sub-range of row t
// with 32*nu threads.  One iteration covers 32*VEC samples = blockDim*VEC contiguous floats, so each
// of a thread's VEC lanes keeps the same input index i for the whole loop (coalesced float4 loads,
// four in flight per thread); sums stay in registers until one fixed-order block reduction.
// ------------------------------------------------------------------------------------------
// At least 3 blocks per SM (<= 62 registers at nu = 11): the streaming loop needs ~45; without the bound the finalize code
// inlined into the last block (check_reach FK, torque law) sets the kernel's register count and costs a resident block
// -- measured 0.202 ms vs 0.131 ms for the whole-body re-read (2 vs 3 blocks per SM; 4 and 5 spill: 0.141, 0.153).
#ifndef MPPI_WN_MINB
#define MPPI_WN_MINB 3
#endif
template <int MODEL, int VEC>
__global__ void __launch_bounds__(32 * ModelNu<MODEL>::value, MPPI_WN_MINB)
weighted_noise_kernel(const __grid_constant__ StepParams P, const __grid_constant__ DynBlock D,
                      const float *__restrict__ w, const float *__restrict__ noise, int32_t *rho_enc,
                      int chunk, float *__restrict__ part, const float *__restrict__ eta_part, int n_eta,
                      uint32_t *counter, float *wsum, int fuse, const float *u_nom, float *u_new, float *out,
                      const __grid_constant__ P2PParams X)
{
    constexpr int NU = ModelNu<MODEL>::value;
    constexpr int NT = 32 * NU;
    constexpr int STEP = 32 * VEC;                     // samples per iteration
    extern __shared__ __align__(16) float s_dyn[];     // reduction | finalize scratch

    const int tid = threadIdx.x;
    const int t = blockIdx.y;
    const int k0 = blockIdx.x * chunk;
    const int nk = min(P.K, k0 + chunk) - k0;

    int koff[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) koff[v] = (tid * VEC + v) / NU;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    const float *row = noise + (static_cast<size_t>(t) * P.K + k0) * NU + tid * VEC;
    const float *wk = w + k0;
    int kb = 0;
    if constexpr (VEC == 4) {
        // main loop: 4 iterations in flight
        for (; kb + 3 * STEP + koff[3] < nk; kb += 4 * STEP) {
            float4 x[4];
            float ww[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = __ldcs(reinterpret_cast<const float4 *>(row + static_cast<size_t>(kb + u * STEP) * NU));
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) ww[u][v] = __ldg(wk + kb + u * STEP + koff[v]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc[0] = fmaf(ww[u][0], x[u].x, acc[0]);
                acc[1] = fmaf(ww[u][1], x[u].y, acc[1]);
                acc[2] = fmaf(ww[u][2], x[u].z, acc[2]);
                acc[3] = fmaf(ww[u][3], x[u].w, acc[3]);
            }
        }
        for (; kb < nk; kb += STEP) {
            if (kb + koff[3] < nk) {
                const float4 x = __ldcs(reinterpret_cast<const float4 *>(row + static_cast<size_t>(kb) * NU));
                acc[0] = fmaf(__ldg(wk + kb + koff[0]), x.x, acc[0]);
                acc[1] = fmaf(__ldg(wk + kb + koff[1]), x.y, acc[1]);
                acc[2] = fmaf(__ldg(wk + kb + koff[2]), x.z, acc[2]);
                acc[3] = fmaf(__ldg(wk + kb + koff[3]), x.w, acc[3]);
            } else {
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    if (kb + koff[v] < nk) acc[v] = fmaf(__ldg(wk + kb + koff[v]), row[static_cast<size_t>(kb) * NU + v], acc[v]);
            }
        }
    } else {
        for (; kb < nk; kb += STEP)
            if (kb + koff[0] < nk) acc[0] = fmaf(__ldg(wk + kb + koff[0]), __ldcs(row + static_cast<size_t>(kb) * NU), acc[0]);
    }
    // ---- reduce the NT*VEC per-thread sums to nu outputs, fixed order
    float *s_red = s_dyn;
#pragma unroll
    for (int v = 0; v < VEC; ++v) s_red[tid * VEC + v] = acc[v];
    __syncthreads();
    const int rowlen = P.T * NU + 2;
    float *my = part + static_cast<size_t>(blockIdx.x) * rowlen;
    if (tid < NU) {
        float v = 0.f;
        for (int j = tid; j < NT * VEC; j += NU) v += s_red[j];
        my[t * NU + tid] = v;
    }
    reduce_partials_and_finalize<MODEL>(P, D, part, gridDim.x, counter, wsum, fuse != 0, u_nom, u_new, out,
                                        rho_enc, s_dyn, X, eta_part, n_eta);
}

// Materialise the Philox noise of one step (equivalence checks, HBM-bound experiments).
template <int NU, int ROUNDS>
__global__ void __launch_bounds__(256)
generate_noise_kernel(const __grid_constant__ StepParams P, uint32_t step_lo, uint32_t step_hi,
                      float *__restrict__ noise)
{
    constexpr int NCH = philox_calls(NU);
    const long long n = static_cast<long long>(P.T) * P.K * NCH;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < n;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % NCH);
        const long long tk = idx / NCH;
        const int k = static_cast<int>(tk % P.K);
        const int t = static_cast<int>(tk / P.K);
        float n6[6];
        normal6<ROUNDS>(static_cast<uint32_t>(P.k_offset + k), static_cast<uint32_t>(t * NCH + c), step_lo, step_hi, P.rkeys, n6);
#pragma unroll
        for (int j = 0; j < 6; ++j)
            if (6 * c + j < NU) noise[(static_cast<size_t>(t) * P.K + k) * NU + 6 * c + j] = __fmul_rn(P.sigma[6 * c + j], n6[j]);
    }
}

// FP32 FFMA throughput probe: 8 independent chains per thread.
static __global__ void __launch_bounds__(256) ffma_probe_kernel(float *out, int iters, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == 12345.678f) out[0] = x0;
}

}  // namespace mppi
