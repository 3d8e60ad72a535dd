// mppi_device.cuh -- device-side building blocks of the B200 MPPI step (sm_100a).
//
// Reference paths ("S/...") are relative to src/mav_mppi/scripts/ of
// cold-deuu/Quadrotor_Manipulator_MPPI; they name the behaviour each block reproduces,
// not code that was translated: the reference is eager PyTorch over [K,T,4,4] tensors,
// this is a register-resident per-sample recurrence.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mppi_b200.h"
#include "fk_tables_gen.cuh"
#include "mppi_vec.cuh"

namespace mppi {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kInvTwoPi = 0.15915494309189533577f;

// Arm chain folded on the host to  C0 Rz(q1) C1 Rz(q2) ... Rz(qn) Cn  (S/robot/urdfparser.py:122-163).
struct ChainDev {
    float R[MPPI_MAX_JOINTS + 1][9];
    float t[MPPI_MAX_JOINTS + 1][3];
    int n;               // revolute joints
    int last_identity;   // Cn == I (true for the j2s7s300_link_7 end link)
    int baked;           // equals the compile-time FkKinova tables (fk_tables_gen.cuh)
    int prismatic;       // bit j: joint j slides along its (folded) z axis instead of rotating about it
    int null_mask;       // bit j: input slot j is not a joint of this (shorter) chain: no motion
};

// Rigid-body parameters of the seven arm links in the folded link frames + the gains of the computed-torque law
// (S/kinova.py:184); used by mppi_dynamics.cuh when MPPI_OPT_TORQUE_LAW is set.
struct ArmInertiaDev {
    float mass[7];
    float com[7][3];
    float inertia[7][6];         // xx xy xz yy yz zz about the centre of mass
    float kp, kd, gravity;
};

// Everything that is fixed for a handle; passed by value as a kernel parameter (constant bank).
struct StepParams {
    int K, T, nu, nch;           // nch = ceil(nu/4) Philox calls per (sample, step)
    long long k_offset;          // global index of local sample 0
    unsigned seed_lo, seed_hi;
    unsigned rkeys[20];          // Philox round keys: seed + r * (0x9E3779B9, 0xBB67AE85)
    float dt, dt2, inv_lambda;
    int sg_window, sg_half;
    float sigma[MPPI_MAX_NU];
    float cost_w[8];
    float quad[6];               // mass, 1/Ixx, 1/Iyy, 1/Izz, kd, gz
    float taps[MPPI_MAX_SAVGOL];
    ChainDev chain;
    // optional cost terms (cost/cost_manager.py:83-87), used by the EXTRA kernel variants only
    int cost_flags;
    float gamma, covar_scale /* covar_weight * lambda * (1 - alpha) */, action_weight, centering_weight,
        joint_traj_weight, limit_penalty;
    float inv_sigma_arm[7];      // Sigma^-1 of the reference is 1/sigma (Sigma = sigma * I)
    float q_center[7], q_lower[7], q_upper[7];
    ArmInertiaDev arm_inertia;
};

// Everything that changes per control step; also passed by value (192 B), so a step needs
// no host->device copy at all.
struct DynBlock {
    float state[MPPI_STATE_FLOATS];
    float target_pos[3];
    float target_R[9];           // quaternion_to_matrix(target xyzw), S/utils/rotation_conversions.py:45-75
    float drone_target[3];
    unsigned step_lo, step_hi;   // Philox counter words 2,3
    // Blocking steps (mppi_step_sync): the finalize block also stores out[] straight into mapped pinned host memory
    // and then publishes host_seq at host_out[MPPI_OUT_FLOATS]; the host spins on that word instead of paying a
    // D2H copy + stream synchronisation.  nullptr = off.
    float *host_out;
    unsigned host_seq;
    // MPPI_OPTION_TRACE: device buffer of kTracePoints %globaltimer stamps (ns) written at the phase boundaries of the
    // step's kernels by one thread of the block that passes them (nullptr = off).
    unsigned long long *trace;
};
constexpr int kTracePoints = 16;
enum TracePoint { TR_START = 0, TR_ROLLOUT_DONE = 1, TR_MIN_KNOWN = 2, TR_SUMS_ADDED = 3, TR_LAST_BLOCK = 4, TR_REDUCED = 5,
                  TR_EXCHANGED = 6, TR_CONTROLS_UPDATED = 7, TR_END = 8, TR_WEIGHT_START = 9 };
__device__ __forceinline__ void trace_stamp(const DynBlock &D, int point, bool who)
{
    if (D.trace != nullptr && who) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        D.trace[point] = t;
    }
}

// Peer-to-peer exchange over NVLink (K-sharded replicas, one process per GPU; buffers are
// cudaMalloc'ed and mapped into every rank with CUDA IPC).  Rank r's buffer holds, behind a 1 KB reserved header,
//   inbox [2 parity][world][rowp] 64-bit words {value, epoch}: slot (parity, s, j) = element j of the row published
//   by source rank s at an epoch of that parity (low-latency protocol, see p2p_exchange in mppi_kernels.cuh).
// world == 1 disables the exchange.
constexpr int kMaxRanks = 8;
constexpr int kFlagStrideInts = 32;          // header size in units of kMaxRanks * 4 bytes
struct P2PParams {
    int world, rank;
    unsigned epoch;              // starts at 1, +1 per control step, same on every rank; the tag of every published element
    int rowp;                    // elements per inbox row (T*nu + 4, padded to a multiple of 4)
    float *base[kMaxRanks];      // base[r] = rank r's exchange buffer as mapped in THIS process
    unsigned *fail_flag;         // mapped host word: set to the epoch at which a peer never published (sticky)
};

template <int MODEL> struct ModelNu;
template <> struct ModelNu<MPPI_MODEL_DRONE3> { static constexpr int value = 3; };
template <> struct ModelNu<MPPI_MODEL_ARM7>   { static constexpr int value = 7; };
template <> struct ModelNu<MPPI_MODEL_QUAD4>  { static constexpr int value = 4; };
template <> struct ModelNu<MPPI_MODEL_WB11>   { static constexpr int value = 11; };

// ------------------------------------------------------------------------------------------
// Counter-based noise: Philox4x32-10 + Box-Muller.  New in this build (the reference calls
// torch.randn, S/sampling/standard_normal_noise.py:24).  Both the rollout pass and the
// weighting pass derive the normals from the same (sample, step, call) address, so the noise
// never has to exist in HBM; explicit _rn intrinsics keep the two call sites bit-identical
// regardless of how the surrounding code is contracted into FMAs.
// ------------------------------------------------------------------------------------------
// The ten round keys (key + r * Weyl constants) depend only on the seed: the host expands them
// once into StepParams::rkeys, so a round is 2 IMAD.WIDE + 2 LOP3 with constant-bank operands.
// ROUNDS = 10 is the Random123 / cuRAND default; 7 is the smallest round count Salmon et al. (SC'11, table 2)
// report as Crush-resistant (passes BigCrush) -- selectable per handle (MPPI_OPTION_PHILOX_ROUNDS).
template <int ROUNDS = 10>
__device__ __forceinline__ uint4 philox4x32(uint4 c, const uint32_t *rk)
{
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ rk[2 * r], lo1, hi0 ^ c.w ^ rk[2 * r + 1], lo0);
    }
    return c;
}

// Box-Muller: u1 = 2 - [1,2) in (0,1];  theta = ([1,2) - 1.5) * 2pi in [-pi,pi);  MUFU lg2 / sqrt / sin / cos.
// r = sqrt(-2 ln u1) = sqrt(lg2(u1) * (-2 ln 2)); sqrt.approx maps 0 -> 0 (u1 == 1).
// One Philox4x32-10 call yields 128 bits = SIX 21-bit uniforms (126 bits used) = three Box-Muller pairs =
// six normals, so nu = 11 needs two calls per (sample, step) instead of three.  Uniform i is bits
// [21 i, 21 i + 21) of the little-endian 128-bit word, placed in the top of a float mantissa: f in [1, 2)
// with 2^-21 steps (radius up to sqrt(2*21*ln 2) = 5.4 sigma, angle step 3e-6 rad).
constexpr uint32_t kU21 = 0x1FFFFFu;
__device__ __forceinline__ float mantissa_field(uint32_t x)          // (x & (kU21 << 2)) | 0x3f800000 as ONE LOP3 (LUT 0xEA = (a & b) | c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(x), "r"(kU21 << 2), "r"(0x3f800000u));
    return __uint_as_float(d);
}
__device__ __forceinline__ void philox_uniforms6(uint4 r, float f[6])
{
    // the 21-bit field is shifted straight to mantissa bits [2, 23), then masked and given its exponent in one LOP3:
    // one shift (or funnel shift) + one LOP3 per uniform
    f[0] = mantissa_field(r.x << 2);                                // bits [  0,  21)
    f[1] = mantissa_field(__funnelshift_r(r.x, r.y, 19));           // bits [ 21,  42)
    f[2] = mantissa_field(r.y >> 8);                                // bits [ 42,  63)
    f[3] = mantissa_field(__funnelshift_r(r.y, r.z, 29));           // bits [ 63,  84)
    f[4] = mantissa_field(__funnelshift_r(r.z, r.w, 18));           // bits [ 84, 105)
    f[5] = mantissa_field(r.w >> 7);                                // bits [105, 126)
}
// Number of Philox calls per (sample, step): pairs = ceil(nu / 2), three pairs per call.
__host__ __device__ constexpr int philox_calls(int nu) { return ((nu + 1) / 2 + 2) / 3; }

// Two Box-Muller pairs at once in packed FP32x2.  fu = radius uniforms, ft = angle uniforms, both in [1, 2):
//   u1 = 2 - fu in (0, 1],  r = sqrt(-2 ln u1) = sqrt(lg2(u1) * (-2 ln 2))  (sqrt.approx maps 0 -> 0),
//   theta = (ft - 1.5) * 2 pi in [-pi, pi);  rc = r cos(theta), rs = r sin(theta)       (MUFU lg2 / sqrt / sin / cos)
// The MUFU evaluations are issued as the .ftz PTX forms: u1 >= 2^-21 and |theta| <= pi are never subnormal, and the
// non-ftz __log2f / __sincosf carry a subnormal-input guard (FSETP + FMUL + FADD per call) this path does not need.
__device__ __forceinline__ float lg2_ftz(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void sincos_ftz(float x, float &s, float &c)
{
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(x));
}
__device__ __forceinline__ void box_muller2(f2 fu, f2 ft, f2 &rc, f2 &rs)
{
    const f2 u1 = vadd(f2(2.0f), vneg(fu));
    const f2 th = vfma(ft, f2(kTwoPi), f2(-9.42477796076937971538f));      // (ft - 1.5) * 2 pi as one FMA
    const f2 l2(lg2_ftz(u1.v.x), lg2_ftz(u1.v.y));
    const f2 a = vmul(l2, f2(-1.3862943611198906f));
    const f2 r(sqrt_approx(a.v.x), sqrt_approx(a.v.y));
    float s0, c0, s1, c1;
    sincos_ftz(th.v.x, s0, c0);
    sincos_ftz(th.v.y, s1, c1);
    rc = vmul(r, f2(c0, c1));
    rs = vmul(r, f2(s0, s1));
}

// The 12 uniforms of up to two calls for (global sample kg, horizon step t): call j has
// counter = (kg, t * ncalls + j, step_lo, step_hi), key = seed.  Pair p = uniforms (2p, 2p+1) -> normals (2p, 2p+1).
template <int NCALLS, int ROUNDS = 10>
__device__ __forceinline__ void philox_step_uniforms(uint32_t kg, uint32_t t, uint32_t step_lo, uint32_t step_hi,
                                                     const uint32_t *rkeys, float f[6 * NCALLS])
{
#pragma unroll
    for (int j = 0; j < NCALLS; ++j)
        philox_uniforms6(philox4x32<ROUNDS>(make_uint4(kg, t * NCALLS + j, step_lo, step_hi), rkeys), f + 6 * j);
}
// Normals for inputs 4e .. 4e+3 (pairs 2e and 2e+1) as rc = (n[4e], n[4e+2]), rs = (n[4e+1], n[4e+3]).
__device__ __forceinline__ void normals_quad(const float *f, int e, f2 &rc, f2 &rs)
{
    box_muller2(f2(f[4 * e], f[4 * e + 2]), f2(f[4 * e + 1], f[4 * e + 3]), rc, rs);
}
// The six normals of ONE call (weighting pass / noise materialisation): n[2p], n[2p+1] from pair p.
template <int ROUNDS = 10>
__device__ __forceinline__ void normal6(uint32_t kg, uint32_t tcall, uint32_t step_lo, uint32_t step_hi,
                                        const uint32_t *rkeys, float n[6])
{
    float f[6];
    philox_uniforms6(philox4x32<ROUNDS>(make_uint4(kg, tcall, step_lo, step_hi), rkeys), f);
    f2 rc, rs;
    box_muller2(f2(f[0], f[2]), f2(f[1], f[3]), rc, rs);
    n[0] = rc.v.x; n[1] = rs.v.x; n[2] = rc.v.y; n[3] = rs.v.y;
    box_muller2(f2(f[4], f[4]), f2(f[5], f[5]), rc, rs);
    n[4] = rc.v.x; n[5] = rs.v.x;
}

// ------------------------------------------------------------------------------------------
// Order-preserving float <-> int32 map (signed order), so the cost minimum is one atomicMin
// on the device and one MIN all-reduce on an int32 across shards.
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int32_t encode_ordered(float f)
{
#ifdef __CUDA_ARCH__
    const int32_t i = __float_as_int(f);
#else
    int32_t i; memcpy(&i, &f, 4);
#endif
    return i >= 0 ? i : (i ^ 0x7fffffff);
}
__host__ __device__ __forceinline__ float decode_ordered(int32_t e)
{
    const int32_t i = e >= 0 ? e : (e ^ 0x7fffffff);
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}
constexpr int32_t kRhoInit = 0x7fffffff;

// ------------------------------------------------------------------------------------------
// Rigid-body helpers (row-major 3x3), generic over the value type V (float: one sample per thread,
// f2: two samples per thread in packed FP32x2 -- see mppi_vec.cuh).
// ------------------------------------------------------------------------------------------
// Rz(yaw) Ry(pitch) Rx(roll)  (S/robot/transformation_matrix.py:4-25, :148-187; S/drone.py:126-154)
template <class V>
__device__ __forceinline__ void rpy_matrix(V sr, V cr, V sp, V cp, V sy, V cy, V R[9])
{
    const V cysp = vmul(cy, sp), sysp = vmul(sy, sp);
    R[0] = vmul(cy, cp); R[1] = vfma(cysp, sr, vneg(vmul(sy, cr))); R[2] = vfma(cysp, cr, vmul(sy, sr));
    R[3] = vmul(sy, cp); R[4] = vfma(sysp, sr, vmul(cy, cr));       R[5] = vfma(sysp, cr, vneg(vmul(cy, sr)));
    R[6] = vneg(sp);     R[7] = vmul(cp, sr);                       R[8] = vmul(cp, cr);
}

// xyz + quaternion xyzw, NOT normalised (S/robot/urdf_fk.py:30-55).  The base pose is uniform.
__device__ __forceinline__ void quat_matrix(const float *b, float R[9])
{
    const float qx = b[3], qy = b[4], qz = b[5], qw = b[6];
    R[0] = 1.0f - 2.0f * qy * qy - 2.0f * qz * qz; R[1] = 2.0f * qx * qy - 2.0f * qz * qw; R[2] = 2.0f * qx * qz + 2.0f * qy * qw;
    R[3] = 2.0f * qx * qy + 2.0f * qz * qw; R[4] = 1.0f - 2.0f * qx * qx - 2.0f * qz * qz; R[5] = 2.0f * qy * qz - 2.0f * qx * qw;
    R[6] = 2.0f * qx * qz - 2.0f * qy * qw; R[7] = 2.0f * qy * qz + 2.0f * qx * qw; R[8] = 1.0f - 2.0f * qx * qx - 2.0f * qy * qy;
}

// (R,p) <- (R Cr, p + R Ct), constants from the kernel parameter block (any chain)
template <class V>
__device__ __forceinline__ void compose_const(V R[9], V p[3], const float *Cr, const float *Ct)
{
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const V a = R[3 * r], b = R[3 * r + 1], c = R[3 * r + 2];
        p[r] = vfma(a, V(Ct[0]), vfma(b, V(Ct[1]), vfma(c, V(Ct[2]), p[r])));
        R[3 * r]     = vfma(a, V(Cr[0]), vfma(b, V(Cr[3]), vmul(c, V(Cr[6]))));
        R[3 * r + 1] = vfma(a, V(Cr[1]), vfma(b, V(Cr[4]), vmul(c, V(Cr[7]))));
        R[3 * r + 2] = vfma(a, V(Cr[2]), vfma(b, V(Cr[5]), vmul(c, V(Cr[8]))));
    }
}

// sin and cos of x to ~1 ulp (max abs error 1.4e-7) without the libm slow path: reduce by pi
// (2-term Cody-Waite under FMA) to r in [-pi/2, pi/2], near-minimax polynomials in r^2, one sign
// flip.  Valid for |x| < ~1e4 (joint and Euler angles here stay within a few turns).
template <class V>
__device__ __forceinline__ void sincos_pi(V x, V &s, V &c)
{
    const V kf = vfma(x, V(0.318309886183790672f), V(12582912.0f));      // 1.5 * 2^23: rint in the low bits
    const auto n = vbits(kf);
    const V k = vadd(kf, V(-12582912.0f));
    V r = vfma(k, V(-3.14159274101257324f), x);
    r = vfma(k, V(8.742278000372485e-8f), r);
    const V z = vmul(r, r);
    V ps = vfma(V(2.634697921166662e-06f), z, V(-0.00019822725153062493f));
    ps = vfma(ps, z, V(0.008333242498338223f));
    ps = vfma(ps, z, V(-0.1666666567325592f));
    V pc = vfma(V(-2.62966580066859e-07f), z, V(2.4774544726824388e-05f));
    pc = vfma(pc, z, V(-0.0013888651737943292f));
    pc = vfma(pc, z, V(0.0416666604578495f));
    pc = vfma(pc, z, V(-0.5f));
    const V sr = vfma(ps, vmul(z, r), r);
    const V cr = vfma(pc, z, V(1.0f));
    const auto flip = vshl31(n);                                          // odd multiple of pi: negate both
    s = vxor(sr, flip);
    c = vxor(cr, flip);
}

// sin / cos on the MUFU unit for the two UNPINNED models (QUAD4, WB11): sin.approx / cos.approx (max abs error 2^-20.9 ~ 5e-7
// on [-pi, pi] against 1.4e-7 for sincos_pi).  The FMA pipe / register operand bandwidth bound the rollout kernel and the
// XU pipe overlaps them (tools/probe_pipes2.cu): ten polynomial evaluations per whole-body step were 150 of its ~600
// FMA-pipe cycles.  NO range reduction here: the instruction's own x / 2 pi step loses |x| * 2^-23 turns, so the callers
// keep their arguments within a turn or so -- the Euler angles are wrapped to [-pi, pi] by quad_advance every step, and the
// joint angles are measured from whole_turns_removed(q0) (one reduction per rollout instead of one per horizon step).
// The pinned models (ARM7, DRONE3) keep the polynomial: their costs are held to 2e-6 of the reference's and the soft-min
// amplifies cost errors by 1/lambda.
template <class V>
__device__ __forceinline__ void sincos_mufu(V x, V &s, V &c)
{
    s = vmap(x, [](float a) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a)); return y; });
    c = vmap(x, [](float a) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a)); return y; });
}
// x - 2 pi rint(x / 2 pi), two-term Cody-Waite: the trig argument base of a joint whose angle is q0 + (a small excursion)
template <class V>
__device__ __forceinline__ V whole_turns_removed(V x)
{
    const V k = vadd(vfma(x, V(kInvTwoPi), V(12582912.0f)), V(-12582912.0f));      // rint(x / 2 pi)
    V r = vfma(k, V(-6.28318548202514648f), x);
    return vfma(k, V(1.74845553146951715e-7f), r);
}
template <bool FAST, class V>
__device__ __forceinline__ void sincos_sel(V x, V &s, V &c)
{
    if constexpr (FAST) sincos_mufu(x, s, c);
    else sincos_pi(x, s, c);
}

// Branch-free atan2 / asin for the pose cost.  Accuracy (max abs error vs double): atan2 1.5e-7 plus the
// 2-ulp MUFU quotient, asin 1.6e-7; rcp / sqrt are the MUFU approximations (<= 2 ulp).  The reference's
// own float32 atan2/asin carry ~1e-7; these terms enter the cost multiplied by 30 against a float32
// ulp of S of 1.2e-4.
template <class V>
__device__ __forceinline__ V atan2_poly(V y, V x)
{
    const V ax = vabs(x), ay = vabs(y);
    const V mx = vmax(vmax(ax, ay), V(1e-37f)), mn = vmin(ax, ay);         // atan2(0, 0) -> 0
    const V t = vmul(mn, vrcp(mx));
    const V z = vmul(t, t);
    V p = vfma(V(0.003962172195315361f), z, V(-0.02036452107131481f));
    p = vfma(p, z, V(0.049385011196136475f));
    p = vfma(p, z, V(-0.0804230123758316f));
    p = vfma(p, z, V(0.1087741032242775f));
    p = vfma(p, z, V(-0.14259037375450134f));
    p = vfma(p, z, V(0.19998809695243835f));
    p = vfma(p, z, V(-0.33333325386047363f));
    V a = vfma(vmul(p, z), t, t);
    a = vsel(vgt(ay, ax), vsub(V(1.57079632679489662f), a), a);
    a = vsel(vlt(x, V(0.0f)), vsub(V(3.14159265358979324f), a), a);
    return vcopysign(a, y);
}
template <class V>
__device__ __forceinline__ V asin_poly(V x)
{
    const V a = vabs(x);
    const auto big = vgt(a, V(0.5f));
    const V z = vsel(big, vfma(a, V(-0.5f), V(0.5f)), vmul(a, a));
    const V s = vsel(big, vsqrt(z), a);
    V p = vfma(V(0.038206253200769424f), z, V(0.02649438939988613f));
    p = vfma(p, z, V(0.045010678470134735f));
    p = vfma(p, z, V(0.07498808950185776f));
    p = vfma(p, z, V(0.16666673123836517f));
    V r = vfma(vmul(p, z), s, s);
    r = vsel(big, vfma(V(-2.0f), r, V(1.57079632679489662f)), r);
    return vcopysign(r, x);
}

// ---- FK with compile-time constants (tables from tools/gen_fk_tables.py) -------------------
// acc (+)= x * k where k is a constant expression: zero terms vanish, +-1 become add / sub.
#define MPPI_CTERM(acc, have, x, k)                                                                              \
    if constexpr ((k) != 0.0f) {                                                                                 \
        if constexpr (!(have)) acc = ((k) == 1.0f) ? (x) : ((k) == -1.0f) ? vneg(x) : vmul((x), V(k));           \
        else acc = ((k) == 1.0f) ? vadd(acc, (x)) : ((k) == -1.0f) ? vsub(acc, (x)) : vfma((x), V(k), acc);      \
    }

template <class Tab, int J, int COL, class V>
__device__ __forceinline__ V tab_rot_col(V a, V b, V c)
{
    constexpr float k0 = Tab::R[J][COL], k1 = Tab::R[J][3 + COL], k2 = Tab::R[J][6 + COL];
    constexpr bool h1 = (k0 != 0.0f), h2 = h1 || (k1 != 0.0f);
    V acc = V(0.0f);
    MPPI_CTERM(acc, false, a, k0)
    MPPI_CTERM(acc, h1, b, k1)
    MPPI_CTERM(acc, h2, c, k2)
    return acc;
}
template <class Tab, int J, class V>
__device__ __forceinline__ V tab_trans(V a, V b, V c, V p)
{
    constexpr float k0 = Tab::t[J][0], k1 = Tab::t[J][1], k2 = Tab::t[J][2];
    V acc = p;
    MPPI_CTERM(acc, true, a, k0)
    MPPI_CTERM(acc, true, b, k1)
    MPPI_CTERM(acc, true, c, k2)
    return acc;
}
// (R,p) <- (R C_J.R, p + R C_J.t) with C_J from the table
template <class Tab, int J, class V>
__device__ __forceinline__ void compose_tab(V R[9], V p[3])
{
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const V a = R[3 * r], b = R[3 * r + 1], c = R[3 * r + 2];
        p[r] = tab_trans<Tab, J>(a, b, c, p[r]);
        R[3 * r] = tab_rot_col<Tab, J, 0>(a, b, c);
        R[3 * r + 1] = tab_rot_col<Tab, J, 1>(a, b, c);
        R[3 * r + 2] = tab_rot_col<Tab, J, 2>(a, b, c);
    }
}
// ---- Pose in packed layout ----------------------------------------------------------------
// One sample per thread, but the FK still packs: each rotation column is held as the pair
// (row 0, row 1) plus a row-2 scalar, the position as (x, y) + z.  Rotating by a joint angle is then
// 4 packed + 4 scalar ops instead of 12, with (cos, sin) as broadcast scalar operands.
struct Pose3 {
    f2 c0, c1, c2;       // columns, rows 0..1
    float r0, r1, r2;    // row 2 of columns 0..2
    f2 pxy;
    float pz;
};
__device__ __forceinline__ Pose3 pose3_from(const float R[9], const float p[3])
{
    Pose3 T;
    T.c0 = f2(R[0], R[3]); T.c1 = f2(R[1], R[4]); T.c2 = f2(R[2], R[5]);
    T.r0 = R[6]; T.r1 = R[7]; T.r2 = R[8];
    T.pxy = f2(p[0], p[1]); T.pz = p[2];
    return T;
}
// Rz(yaw) Ry(pitch) Rx(roll) with e = (cy, sy), e_perp = (-sy, cy):
//   col0 = e cp,  col1 = e (sp sr) + e_perp cr,  col2 = e (sp cr) - e_perp sr   (transformation_matrix.py:148-187)
__device__ __forceinline__ void pose3_rpy(Pose3 &T, float sr, float cr, float sp, float cp, float sy, float cy)
{
    const f2 e(cy, sy), ep(-sy, cy);
    T.c0 = vmul(e, f2(cp));
    T.c1 = vfma(e, f2(sp * sr), vmul(ep, f2(cr)));
    T.c2 = vfma(e, f2(sp * cr), vneg(vmul(ep, f2(sr))));
    T.r0 = -sp; T.r1 = cp * sr; T.r2 = cp * cr;
}
__device__ __forceinline__ void pose3_rotate_z(Pose3 &T, float c, float s)
{
    const f2 a = T.c0, b = T.c1;
    T.c0 = vfma(a, f2(c), vmul(b, f2(s)));
    T.c1 = vfma(b, f2(c), vneg(vmul(a, f2(s))));
    const float u = T.r0, w = T.r1;
    T.r0 = fmaf(c, u, s * w);
    T.r1 = fmaf(c, w, -s * u);
}
template <class Tab, int J>
__device__ __forceinline__ void pose3_compose_tab(Pose3 &T)
{
    const f2 a = T.c0, b = T.c1, c = T.c2;
    const float u = T.r0, v = T.r1, w = T.r2;
    T.pxy = tab_trans<Tab, J>(a, b, c, T.pxy);
    T.pz = tab_trans<Tab, J>(u, v, w, T.pz);
    T.c0 = tab_rot_col<Tab, J, 0>(a, b, c); T.c1 = tab_rot_col<Tab, J, 1>(a, b, c); T.c2 = tab_rot_col<Tab, J, 2>(a, b, c);
    T.r0 = tab_rot_col<Tab, J, 0>(u, v, w); T.r1 = tab_rot_col<Tab, J, 1>(u, v, w); T.r2 = tab_rot_col<Tab, J, 2>(u, v, w);
}
__device__ __forceinline__ void pose3_compose_const(Pose3 &T, const float *Cr, const float *Ct)
{
    const f2 a = T.c0, b = T.c1, c = T.c2;
    const float u = T.r0, v = T.r1, w = T.r2;
    T.pxy = vfma(a, f2(Ct[0]), vfma(b, f2(Ct[1]), vfma(c, f2(Ct[2]), T.pxy)));
    T.pz = fmaf(u, Ct[0], fmaf(v, Ct[1], fmaf(w, Ct[2], T.pz)));
    T.c0 = vfma(a, f2(Cr[0]), vfma(b, f2(Cr[3]), vmul(c, f2(Cr[6]))));
    T.c1 = vfma(a, f2(Cr[1]), vfma(b, f2(Cr[4]), vmul(c, f2(Cr[7]))));
    T.c2 = vfma(a, f2(Cr[2]), vfma(b, f2(Cr[5]), vmul(c, f2(Cr[8]))));
    T.r0 = fmaf(u, Cr[0], fmaf(v, Cr[3], w * Cr[6]));
    T.r1 = fmaf(u, Cr[1], fmaf(v, Cr[4], w * Cr[7]));
    T.r2 = fmaf(u, Cr[2], fmaf(v, Cr[5], w * Cr[8]));
}
template <class Tab, int J = 0>
__device__ __forceinline__ void pose3_fk_tab(const float *cq, const float *sq, Pose3 &T)
{
    if constexpr (J < Tab::kJoints) {
        pose3_rotate_z(T, cq[J], sq[J]);
        pose3_compose_tab<Tab, J + 1>(T);
        pose3_fk_tab<Tab, J + 1>(cq, sq, T);
    }
}
template <int NJ>
__device__ __forceinline__ void pose3_fk_chain(const ChainDev &ch, const float *qv, const float *cq, const float *sq, Pose3 &T)
{
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        if ((ch.null_mask >> j) & 1) continue;      // padding slot of a chain with fewer than NJ actuated joints (uniform branch)
        if ((ch.prismatic >> j) & 1) {     // S/robot/transformation_matrix.py:38-55: translate by q along the joint axis (uniform branch)
            T.pxy = vfma(T.c2, f2(qv[j]), T.pxy);
            T.pz = fmaf(T.r2, qv[j], T.pz);
        } else {
            pose3_rotate_z(T, cq[j], sq[j]);
        }
        if (j + 1 < NJ || !ch.last_identity) pose3_compose_const(T, ch.R[j + 1], ch.t[j + 1]);
    }
}
// ||p - p*||_2 and ||euler_ZYX(R^T R*)||_2 on a Pose3 (S/cost/pose_cost.py:24-63, S/utils/rotation_conversions.py:277-319;
// inv(R) of a rotation is its transpose); the two atan2 of the ZYX Euler extraction run as one packed evaluation.
__device__ __forceinline__ void pose3_terms(const Pose3 &T, const DynBlock &D, float &pos, float &ori)
{
    const float *Tg = D.target_R;
    const f2 t0(Tg[0], Tg[3]), t1(Tg[1], Tg[4]), t2(Tg[2], Tg[5]);
    const f2 dxy = vadd(T.pxy, f2(-D.target_pos[0], -D.target_pos[1]));
    const float dz = T.pz - D.target_pos[2];
    const f2 dd = vmul(dxy, dxy);
    pos = sqrt_approx(fmaf(dz, dz, dd.v.x + dd.v.y));
    const f2 m00 = vmul(T.c0, t0), m10 = vmul(T.c1, t0), m20 = vmul(T.c2, t0), m21 = vmul(T.c2, t1), m22 = vmul(T.c2, t2);
    const float d00 = fmaf(T.r0, Tg[6], m00.v.x + m00.v.y);
    const float d10 = fmaf(T.r1, Tg[6], m10.v.x + m10.v.y);
    const float d20 = fmaf(T.r2, Tg[6], m20.v.x + m20.v.y);
    const float d21 = fmaf(T.r2, Tg[7], m21.v.x + m21.v.y);
    const float d22 = fmaf(T.r2, Tg[8], m22.v.x + m22.v.y);
    const f2 e02 = atan2_poly(f2(d10, d21), f2(d00, d22));
    const float e1 = asin_poly(fminf(fmaxf(-d20, -1.0f), 1.0f));
    const f2 ee = vmul(e02, e02);
    ori = sqrt_approx(fmaf(e1, e1, ee.v.x + ee.v.y));
}

// Quadrotor rigid body, one step (restated from the dead draft S/mppi_solver/drone_mppi.py:57-83;
// rules fixed in DESIGN.md).  (s*, c*) are sin/cos of the CURRENT attitude on entry and of the
// NEW attitude on exit, so the whole-body FK reuses them.
template <class V>
struct QuadState {
    V p[3], rpy[3], v[3], w[3];
    V sphi, cphi, sth, cth, spsi, cpsi;
};

// a - 2 pi rint(a / 2 pi): exact identity for |a| < pi, so it is applied unconditionally
// (the draft wraps with atan2(sin, cos), drone_mppi.py:77).
template <class V>
__device__ __forceinline__ V wrap_pi(V a)
{
    const V k = vadd(vfma(a, V(kInvTwoPi), V(12582912.0f)), V(-12582912.0f));
    return vfma(V(-kTwoPi), k, a);
}

// WRAP = false (the MUFU sin / cos of the unpinned models): the angles only ever enter through sin / cos, whose own
// range step makes the per-step wrap redundant within a horizon; quad_load removes the whole turns of the measured angles.
template <bool REFRESH = true, bool WRAP = true, class V>
__device__ __forceinline__ void quad_advance(QuadState<V> &s, V F, V tx, V ty, V tz, float dt, const float *qp)
{
    const float inv_m = rcp_approx(qp[0]), kd = qp[4], gz = qp[5];
    const V inv_cth = vrcp(s.cth);
    const V tth = vmul(s.sth, inv_cth);
    const V cs = vmul(s.cpsi, s.sth), ss = vmul(s.spsi, s.sth);
    const V r02 = vfma(cs, s.cphi, vmul(s.spsi, s.sphi));
    const V r12 = vfma(ss, s.cphi, vneg(vmul(s.cpsi, s.sphi)));
    const V r22 = vmul(s.cth, s.cphi);
    s.w[0] = vfma(V(dt * qp[1]), tx, s.w[0]);
    s.w[1] = vfma(V(dt * qp[2]), ty, s.w[1]);
    s.w[2] = vfma(V(dt * qp[3]), tz, s.w[2]);
    const V dphi = vfma(vmul(s.cphi, tth), s.w[2], vfma(vmul(s.sphi, tth), s.w[1], s.w[0]));
    const V dth = vfma(s.cphi, s.w[1], vneg(vmul(s.sphi, s.w[2])));
    const V dpsi = vfma(vmul(s.cphi, inv_cth), s.w[2], vmul(vmul(s.sphi, inv_cth), s.w[1]));
    const V ax = vmul(vfma(r02, F, vneg(vmul(V(kd), s.v[0]))), V(inv_m));
    const V ay = vmul(vfma(r12, F, vneg(vmul(V(kd), s.v[1]))), V(inv_m));
    const V az = vfma(vfma(r22, F, vneg(vmul(V(kd), s.v[2]))), V(inv_m), V(gz));
    s.rpy[0] = vfma(V(dt), dphi, s.rpy[0]);
    s.rpy[1] = vfma(V(dt), dth, s.rpy[1]);
    s.rpy[2] = vfma(V(dt), dpsi, s.rpy[2]);
    if constexpr (WRAP) { s.rpy[0] = wrap_pi(s.rpy[0]); s.rpy[1] = wrap_pi(s.rpy[1]); s.rpy[2] = wrap_pi(s.rpy[2]); }
    s.v[0] = vfma(V(dt), ax, s.v[0]); s.v[1] = vfma(V(dt), ay, s.v[1]); s.v[2] = vfma(V(dt), az, s.v[2]);
    s.p[0] = vfma(V(dt), s.v[0], s.p[0]); s.p[1] = vfma(V(dt), s.v[1], s.p[1]); s.p[2] = vfma(V(dt), s.v[2], s.p[2]);
    if constexpr (REFRESH) {          // callers that batch the sin/cos of several angles pass REFRESH = false
        sincos_pi(s.rpy[0], s.sphi, s.cphi);
        sincos_pi(s.rpy[1], s.sth, s.cth);
        sincos_pi(s.rpy[2], s.spsi, s.cpsi);
    }
}

template <class V>
__device__ __forceinline__ void quad_load(QuadState<V> &s, const float *st)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) { s.p[i] = V(st[i]); s.rpy[i] = wrap_pi(V(st[3 + i])); s.v[i] = V(st[6 + i]); s.w[i] = V(st[9 + i]); }
    sincos_pi(s.rpy[0], s.sphi, s.cphi);
    sincos_pi(s.rpy[1], s.sth, s.cth);
    sincos_pi(s.rpy[2], s.spsi, s.cpsi);
}

// ------------------------------------------------------------------------------------------
// Warp / block reductions.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// TMA 1-D bulk copy global -> shared with an mbarrier (SASS: UBLKCP).  Used to stage the
// nominal control sequence and the injected-noise tiles.  bytes % 16 == 0, both addresses
// 16-byte aligned.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace mppi
