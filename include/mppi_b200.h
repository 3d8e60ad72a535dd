/*
 * mppi_b200.h -- C ABI of libmppi_b200.so, the B200 (sm_100a) MPPI control step.
 *
 * The reference (cold-deuu/Quadrotor_Manipulator_MPPI) has no FFI layer: its hot path is
 * the Python classes `mppi_solver/mppi.py:MPPI` (arm) and `mppi_solver/drone_mppi.py:MPPI`
 * (drone) that the ROS nodes call (kinova.py:182, drone.py:164-165).  This header is the
 * native boundary a maintainer binds instead (ctypes stub in INTEGRATION.md); the Python
 * drop-in classes in quadrotor_manipulator_mppi_b200/mppi_solver/ sit directly on it.
 * Each entry point names the reference code it replaces.  Paths are relative to
 * src/mav_mppi/scripts/ of the reference.
 *
 * Conventions
 *   - plain C types only; "device" pointers are CUDA device pointers of the handle's device,
 *     16-byte aligned; "host" pointers are ordinary host memory.
 *   - every call returns mppi_status_t; mppi_last_error(h) gives the text.
 *   - there is NO CPU fallback: a non-sm_100 device is MPPI_ERR_WRONG_ARCH.
 *   - noise layout is [T][K][nu] (the reference's sampler emits [K][T][nu],
 *     sampling/standard_normal_noise.py:24; the caller transposes).
 *   - one stepping thread per handle; mppi_set_state() may be called from another thread
 *     (the rospy subscriber thread, kinova.py:106-116) -- it only touches a mutex-guarded
 *     host staging slot that the next step snapshots.
 */
#ifndef MPPI_B200_H_
#define MPPI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPPI_ABI_VERSION 4
#define MPPI_MAX_NU 12          /* controls per horizon step (whole body = 11)            */
#define MPPI_MAX_HORIZON 256
#define MPPI_MAX_JOINTS 8       /* revolute joints in the arm chain                      */
#define MPPI_STATE_FLOATS 32    /* state vector slot, layout per model below             */
#define MPPI_OUT_FLOATS 64      /* per-step outputs + statistics, layout below           */
#define MPPI_MAX_SAVGOL 31

typedef enum {
    MPPI_OK = 0,
    MPPI_ERR_INVALID_ARG = 1,
    MPPI_ERR_WRONG_ARCH = 2,    /* device is not compute capability 10.x                 */
    MPPI_ERR_CUDA = 3,
    MPPI_ERR_UNSUPPORTED = 4,
    MPPI_ERR_PEER = 5           /* K-sharded step: a peer shard never published its row (exchange timed out);
                                   the controls of that step were not updated                      */
} mppi_status_t;

/* Models.  state[] / out[] layouts:
 *  DRONE3 (drone_mppi.py, live nu=3 point mass): state = x[3], v[3]
 *          out = x_des[3], v_des[3]                              (drone_mppi.py:169-170)
 *  ARM7   (mppi.py, Kinova j2s7s300 nu=7):       state = q[7], qdot[7], base[7] (xyz + quat xyzw)
 *          [, base twist[6] = v_full[:6] of kinova.py:106-116 (linear, angular; base frame) -- 27 floats; only
 *          the torque law reads it, 21 floats = zero twist]
 *          out = qdes[7], vdes[7]                                (mppi.py:157-158, incl. the
 *                                                                 `_qddot * dt` term)
 *          out[57..63] = joint torques of the computed-torque law (kinova.py:184) when MPPI_OPT_TORQUE_LAW is set
 *  QUAD4  (rigid body nu=4; restated from the dead draft drone_mppi.py:57-83, PARITY UNPINNED):
 *          state = p[3], rpy[3], v[3], w[3];  u = F, tau_xyz;  out = next state[12]
 *  WB11   (whole body nu=11, not in the reference, PARITY UNPINNED):
 *          state = p, rpy, v, w (12), q[7], qdot[7]; u = F, tau_xyz, qdd[7]
 *          out = qdes[7], vdes[7]; out[16..27] = next base state[12]
 *  Common tail of out[]: [28..39] u_new[0][:] (first updated control), [40..51] u_nom[0][:] (the
 *  previous first control, i.e. the reference's `_qddot`), then the MPPI_OUT_* scalars below.  */
typedef enum {
    MPPI_MODEL_DRONE3 = 0,
    MPPI_MODEL_ARM7 = 1,
    MPPI_MODEL_QUAD4 = 2,
    MPPI_MODEL_WB11 = 3
} mppi_model_t;

#define MPPI_OUT_BASE 16        /* WB11: next base state                                 */
#define MPPI_OUT_U0_NEW 28
#define MPPI_OUT_U0_OLD 40
#define MPPI_OUT_REACH 52       /* ARM7/WB11: L1 position error of FK(qdes) to the target (check_reach, mppi.py:95-120) */
#define MPPI_OUT_RHO 53         /* minimum cost                                          */
#define MPPI_OUT_ETA 54         /* sum of unnormalised weights                           */
#define MPPI_OUT_ESS 55         /* effective sample size (sum w)^2 / sum w^2             */
#define MPPI_OUT_STEP 56        /* low 24 bits of the step counter used                  */
#define MPPI_OUT_TORQUE 57      /* ARM7 / WB11 + MPPI_OPT_TORQUE_LAW: torque[7] = M[6:,6:] (kp (qdes - q) - kd qdot) + nle[6:] (kinova.py:184) */

/* Replaces the hard-coded constructor constants of mppi.py:37-42,75,
 * sampling/standard_normal_noise.py:17, drone_mppi.py:16-19,32,34 and
 * cost/cost_manager.py:30-33.  mppi_default_config() fills the reference's values.      */
typedef struct mppi_config {
    int32_t abi_version;        /* MPPI_ABI_VERSION                                      */
    int32_t model;              /* mppi_model_t                                          */
    int32_t n_samples;          /* K held by THIS handle (the local shard)               */
    int32_t n_horizon;          /* T                                                     */
    int32_t savgol_window;      /* odd; 9 arm/whole body, 5 drone/quad (mppi.py:149, drone_mppi.py:160) */
    int32_t savgol_polyorder;   /* 2                                                     */
    int32_t device;             /* CUDA device ordinal                                   */
    int32_t n_joints;           /* revolute joints of the arm chain (7), see mppi_set_chain */
    int64_t k_offset;           /* global index of this shard's first sample (Philox addressing) */
    uint64_t seed;              /* Philox key                                            */
    float dt;                   /* 0.01                                                  */
    float lambda_;              /* 0.1                                                   */
    float sigma[MPPI_MAX_NU];   /* per-input noise std (the reference's Sigma = sigma*I multiplies z: std = sigma) */
    /* cost weights: ARM7/WB11 [0..3] = stage pos, stage ori, terminal pos, terminal ori
     *               DRONE3/QUAD4 [4..5] = stage, terminal (squared distance); WB11 uses [0..5] */
    float cost_w[8];
    float quad_params[6];       /* mass, 1/Ixx, 1/Iyy, 1/Izz, k_d, g_z                   */
    float target_pos[3];        /* mppi.py:71                                            */
    float target_quat[4];       /* xyzw, mppi.py:72                                      */
    float drone_target[3];      /* drone_mppi.py:141                                     */
    float torque_kp;            /* 400   kinova.py:184                                                       */
    float torque_kd;            /* 40    kinova.py:184                                                       */
    float reserved[3];
    /* The cost terms the reference constructs but leaves commented out of the sum
     * (cost/cost_manager.py:83-87); enabled per bit, ARM7 / WB11 only.  Defaults are the
     * reference's weights: covar_cost.py:14,20-25, action_cost.py:15-25, joint_space_cost.py:13,18-77. */
    int32_t cost_flags;         /* MPPI_COST_* bits; 0 = the reference's live cost (pose terms only)        */
    float gamma;                /* 0.98  discount (cost_manager.py:26)                                       */
    float covar_weight;         /* 0.1   (cost_manager.py:36)                                                */
    float alpha;                /* 0.1   (cost_manager.py:25): param_gamma = lambda * (1 - alpha)            */
    float action_weight;        /* 0.01  (cost_manager.py:39)                                                */
    float centering_weight;     /* 1.0   (cost_manager.py:42)                                                */
    float joint_traj_weight;    /* 1.0   (cost_manager.py:43)                                                */
    float limit_penalty;        /* 1e10  (joint_space_cost.py:70)                                            */
    float q_center[7];          /* joint_space_cost.py:13                                                    */
    float q_lower[7];           /* joint_space_cost.py:61                                                    */
    float q_upper[7];           /* joint_space_cost.py:62                                                    */
    float reserved2;
} mppi_config_t;

#define MPPI_COST_COVAR 1        /* covar_weight * lambda(1-alpha) * sum_t u_t^T Sigma^-1 v_t   (Sigma = sigma*I as in the reference) */
#define MPPI_COST_CENTERING 2    /* centering_weight * sum_t gamma^t |q_t - q_center|^2                       */
#define MPPI_COST_JOINT_TRAJ 4   /* joint_traj_weight * sum_t gamma^t |q_t - q_traj_t|^2 (q_traj = 0 unless mppi_set_joint_traj) */
#define MPPI_COST_ACTION 8       /* action_weight * sum_t gamma^t |v_t|^2                                     */
#define MPPI_COST_JOINT_LIMIT 16 /* sum_t gamma^t * limit_penalty * [any joint outside q_lower..q_upper]      */
/* Options that share the cost_flags word (they do not select the extra-cost kernels): */
#define MPPI_COST_MASK 31
#define MPPI_OPT_TORQUE_LAW 256  /* ARM7 / WB11: the finalize block also evaluates the arm node's torque law (kinova.py:126-131,184);
                                    WB11 takes the base attitude and twist from its own state (rpy, R^T v, w) */

typedef struct mppi_ctx *mppi_handle_t;

/* Reference defaults for `model` (K, T, sigma, lambda, dt, weights, targets).           */
mppi_status_t mppi_default_config(int32_t model, mppi_config_t *cfg);

/* MPPI.__init__ (mppi.py:28-93 / drone_mppi.py:8-37): allocates every scratch buffer;
 * nothing is allocated on the step path.                                                */
mppi_status_t mppi_create(const mppi_config_t *cfg, mppi_handle_t *out);
mppi_status_t mppi_destroy(mppi_handle_t h);
const char *mppi_last_error(mppi_handle_t h);   /* h may be NULL: last create() error     */
int32_t mppi_abi_version(void);

/* URDFparser chain (robot/urdfparser.py:110-163) as constants: for each joint of the chain
 * from the absolute root to the end link, in order: type (0 fixed, 1 revolute/continuous,
 * 2 prismatic -- robot/transformation_matrix.py:38-55), origin xyz[3], origin rpy[3], axis[3].
 * The library folds it to  C0 J1(q1) C1 ... Jn(qn) Cn  with Ji = Rz(qi) or Trans(0,0,qi); the
 * chain may have 1 to 7 actuated joints (the arm models carry 7 inputs: the slots beyond the chain's
 * joints become null joints whose controls are sampled but cannot influence the cost; more than 7 is
 * MPPI_ERR_UNSUPPORTED; the torque law needs exactly 7).
 * mppi_create() pre-loads the j2s7s300 chain of aerial_manipulator_gpu.urdf.            */
mppi_status_t mppi_set_chain(mppi_handle_t h, int32_t n_chain_joints, const int32_t *types,
                             const float *xyz, const float *rpy, const float *axis);

/* Rigid-body parameters of the seven arm links for the torque law, in the URDF link frames: mass[7],
 * com[7][3], inertia[7][6] = (ixx ixy ixz iyy iyz izz) about the centre of mass, fixed children already merged
 * (what a URDF importer builds; tools/gen_arm_inertia.py).  mppi_create() pre-loads the j2s7s300 values of
 * aerial_manipulation/urdf/full_robot_floating2.urdf (the model kinova.py:55-70 loads into Pinocchio).
 * Prismatic chains are not supported by the torque law.                                                    */
mppi_status_t mppi_set_arm_inertia(mppi_handle_t h, const float *mass, const float *com, const float *inertia);

/* Re-reads the mutable hyper-parameters of cfg (sigma, lambda_, dt, cost_w, quad_params, cost_flags and
 * the extra-cost weights); model,
 * sizes, device, seed and k_offset are fixed at creation.                                */
mppi_status_t mppi_update_config(mppi_handle_t h, const mppi_config_t *cfg);

/* Joint reference trajectory [T][7] for MPPI_COST_JOINT_TRAJ (host floats; the reference passes zeros,
 * mppi.py:137-138).                                                                       */
mppi_status_t mppi_set_joint_traj(mppi_handle_t h, const float *traj_host);

/* target_pose / hard-coded drone target (mppi.py:70-72, drone_mppi.py:141).             */
mppi_status_t mppi_set_target(mppi_handle_t h, const float *target_pos, const float *target_quat_xyzw,
                              const float *drone_target);

/* update_joint (mppi.py:196-200) / set_state (drone_mppi.py:179-183): thread-safe staging of
 * the measured state (host floats, layout per model); snapshotted at the next step.     */
mppi_status_t mppi_set_state(mppi_handle_t h, const float *state_host, int32_t n);

/* compute_control_input (mppi.py:122-169 / drone_mppi.py:140-176), device-pointer form.
 *   d_u_nom  [T][nu]   nominal controls = u_prev (warm start, NOT shifted: SURVEY F4)
 *   d_noise  [T][K][nu] injected noise, or NULL -> in-kernel Philox4x32-R(seed, step_counter), R = MPPI_OPTION_PHILOX_ROUNDS
 *   d_cost_out [K] or NULL  per-sample costs S
 *   d_u_new  [T][nu]   updated controls (may alias d_u_nom)
 *   d_out    [MPPI_OUT_FLOATS] or NULL
 * Asynchronous on `stream` (a cudaStream_t passed as void*).                             */
mppi_status_t mppi_step(mppi_handle_t h, const float *d_u_nom, const float *d_noise,
                        uint64_t step_counter, float *d_cost_out, float *d_u_new, float *d_out,
                        void *stream);

/* The same step split at the two points where K-sharded replicas exchange data:
 *   mppi_rollout   : fused noise + rollout + cost -> S[K], local cost minimum
 *   (allreduce-MIN over mppi_rho_ptr: one int32, order-preserving encoding of the float)
 *   mppi_weight    : exp((rho-S)/lambda) weights, weighted-noise sums -> mppi_wsum_ptr
 *   (allreduce-SUM over mppi_wsum_ptr: T*nu + 2 floats: sums, eta, sum of squared weights)
 *   mppi_finalize  : normalise, Savitzky-Golay, u += w_eps, outputs                     */
mppi_status_t mppi_rollout(mppi_handle_t h, const float *d_u_nom, const float *d_noise,
                           uint64_t step_counter, float *d_cost_out, void *stream);
mppi_status_t mppi_weight(mppi_handle_t h, const float *d_noise, uint64_t step_counter, void *stream);
mppi_status_t mppi_finalize(mppi_handle_t h, const float *d_u_nom, uint64_t step_counter,
                            float *d_u_new, float *d_out, void *stream);
int32_t *mppi_rho_ptr(mppi_handle_t h);         /* device, 1 x int32                     */
float *mppi_wsum_ptr(mppi_handle_t h);          /* device, mppi_wsum_count() floats      */
int32_t mppi_wsum_count(mppi_handle_t h);
float *mppi_cost_ptr(mppi_handle_t h);          /* device, S[K] of the last rollout      */

/* K-sharded replicas without NCCL on the step path: every rank exports a small exchange buffer
 * through CUDA IPC (mppi_p2p_export -> 64-byte handle), the caller all-gathers the handles (any
 * transport) and binds them (mppi_p2p_bind).  mppi_step_p2p is then mppi_step with the shard
 * exchange fused into the weighting kernel: each rank weights with its local cost minimum,
 * publishes (sums, eta, sum w^2, rho) into every peer's inbox over NVLink, and combines the rows
 * with exp(-(rho_r - rho)/lambda) in rank order -- algebraically the allreduce-MIN + allreduce-SUM.
 * Every rank must call it once per control step with the same step_counter (new; no reference
 * counterpart: the reference is single-GPU, mppi_solver/mppi.py:30-31).                         */
#define MPPI_IPC_HANDLE_BYTES 64
mppi_status_t mppi_p2p_export(mppi_handle_t h, int32_t world, void *ipc_handle_out);
mppi_status_t mppi_p2p_bind(mppi_handle_t h, int32_t world, int32_t rank, const void *all_handles);
mppi_status_t mppi_step_p2p(mppi_handle_t h, const float *d_u_nom, const float *d_noise,
                            uint64_t step_counter, float *d_cost_out, float *d_u_new, float *d_out,
                            void *stream);

/* compute_control_input as ONE blocking call for a caller that keeps u_prev on the device (the Python
 * drop-in classes): stages state_host (may be NULL) and runs mppi_step on `stream`; the last block of the
 * step stores out[] straight into mapped pinned host memory and publishes a sequence word the call spins
 * on (no D2H copy, no stream synchronisation; the stream is polled so a failed kernel returns an error).
 * Returns once out_host is complete; d_u_new is complete in stream order.
 * h2d = the state block (kernel parameter), d2h = out[] (zero-copy store).                          */
mppi_status_t mppi_step_sync(mppi_handle_t h, const float *state_host, int32_t n_state,
                             const float *d_u_nom, const float *d_noise, uint64_t step_counter,
                             float *d_u_new, float *out_host, void *stream);

/* The same blocking form of mppi_step_p2p (every rank calls it; each gets the identical out[]).      */
mppi_status_t mppi_step_p2p_sync(mppi_handle_t h, const float *state_host, int32_t n_state,
                                 const float *d_u_nom, const float *d_noise, uint64_t step_counter,
                                 float *d_u_new, float *out_host, void *stream);

/* Host-buffer form (what a non-torch caller uses; also the end-to-end timing path):
 * copies state (if given) and u_inout to the device, steps, copies u_new / out / costs back
 * and synchronises.  noise_host may be NULL (Philox).  With host noise the [T][K][nu] device staging
 * buffer is allocated on first use -- call mppi_reserve_host_noise() once after mppi_create to keep
 * the allocation off the step path.                                                     */
mppi_status_t mppi_reserve_host_noise(mppi_handle_t h);
mppi_status_t mppi_step_host(mppi_handle_t h, const float *state_host, int32_t n_state,
                             float *u_inout_host, const float *noise_host, uint64_t step_counter,
                             float *cost_out_host, float *out_host);

/* Run-time options (new; the reference has no configuration system).
 *   PHILOX_ROUNDS  7 (default: the smallest round count Salmon et al., SC'11, report as Crush-resistant; -30 % multiplies)
 *                  or 10 (Random123 / cuRAND default).  Part of the noise definition: every shard of a sharded solve must agree.
 *   FUSED_STEP     0 (default) / 1: a Philox step whose rollout grid is co-resident (K_local <= 128 x resident blocks,
 *                  T <= 128) runs as ONE cooperative launch (rollout, grid barrier, weighting, exchange, finalize).
 *                  Measured no faster than the dependent-launch pair, hence opt-in.
 *   TIME_PARALLEL  ARM7 / DRONE3 (linear double integrators), default costs (any horizon up to 256): one WARP per sample, the two
 *                  cumulative sums of the reference as warp scans, every (sample, step) evaluates FK + cost on its own
 *                  lane, weighted-noise sums from the registers.  -1 (default) = when K_local <= 8192 (arm) / 4096 (drone),
 *                  0 = never, 1 = whenever eligible
 *   PROFILE        1: CUDA events around the kernels of every step -> mppi_get_kernel_times (SURVEY section 5 tracing hook)
 *   NVTX           1 (default): NVTX ranges "mppi.*" around the launches
 *   LAST_PATH      (read-only) MPPI_PATH_* taken by the most recent step                                              */
#define MPPI_OPTION_PHILOX_ROUNDS 1
#define MPPI_OPTION_FUSED_STEP 2
#define MPPI_OPTION_TIME_PARALLEL 3
#define MPPI_OPTION_PROFILE 4
#define MPPI_OPTION_NVTX 5
#define MPPI_OPTION_LAST_PATH 6
#define MPPI_OPTION_TRACE 7     /* 1: the kernels stamp %globaltimer (ns) at their phase boundaries -> mppi_get_trace */
#define MPPI_OPTION_HOST_YIELD 8 /* 1: blocking steps sched_yield() between polls of the result word (default 0: spin) */
#define MPPI_PATH_TWO_KERNELS 1
#define MPPI_PATH_FUSED 2
#define MPPI_PATH_TIMEPARALLEL 3
mppi_status_t mppi_set_option(mppi_handle_t h, int32_t option, int32_t value);
mppi_status_t mppi_get_option(mppi_handle_t h, int32_t option, int32_t *value);
/* With MPPI_OPTION_PROFILE: device time of the most recent step's kernels, microseconds:
 * us3[0] = rollout kernel (or the whole single-launch step), us3[1] = weighting + finalize (0 for a single launch),
 * us3[2] = MPPI_PATH_*.  Blocks until that step has finished.                                                        */
mppi_status_t mppi_get_kernel_times(mppi_handle_t h, float *us3);
/* With MPPI_OPTION_TRACE: device timestamps (ns, %globaltimer; 0 = point not passed) of the most recent step:
 * [0] first block starts, [1] rollout done (a late block), [2] cost minimum known (single-launch steps),
 * [3] weighted sums added, [4] last block takes over, [5] sums reduced, [6] peer exchange done,
 * [7] controls updated (after Savitzky-Golay), [8] outputs written, [9] weighting kernel starts.
 * Synchronises the device.                                                                                         */
mppi_status_t mppi_get_trace(mppi_handle_t h, uint64_t *stamps_ns, int32_t n);

/* Writes the Philox noise of (seed, step_counter) for this shard to d_noise [T][K][nu]:
 * the exact values the in-kernel generator uses (equivalence checks).                   */
mppi_status_t mppi_generate_noise(mppi_handle_t h, uint64_t step_counter, float *d_noise, void *stream);

/* Measured FP32 FFMA throughput of the device (TFLOP/s), the roofline denominator of the
 * fused rollout kernel (MEASURED_PEAKS.json has no FP32 entry).                         */
mppi_status_t mppi_measure_fp32_peak(int32_t device, float *tflops_out);

/* Algorithmic FLOP per rollout-step (one sample, one horizon step) used for roofline.achieved: counted by
 * oracle/flop_count.py over the restated maths with general (dense) URDF constants, FMA = 2, transcendental
 * evaluations excluded.  The structural figure is the same count with the exact 0 / +-1
 * constants of the j2s7s300 chain skipped (what an unrolled kernel executes, again without transcendentals).  */
double mppi_algorithmic_flops_per_rollout_step(int32_t model);
double mppi_structural_flops_per_rollout_step(int32_t model);

#ifdef __cplusplus
}
#endif
#endif /* MPPI_B200_H_ */
