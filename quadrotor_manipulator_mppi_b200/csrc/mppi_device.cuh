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
};

// Everything that changes per control step; also passed by value (192 B), so a step needs
// no host->device copy at all.
struct DynBlock {
    float state[MPPI_STATE_FLOATS];
    float target_pos[3];
    float target_R[9];           // quaternion_to_matrix(target xyzw), S/utils/rotation_conversions.py:45-75
    float drone_target[3];
    unsigned step_lo, step_hi;   // Philox counter words 2,3
};

// Peer-to-peer exchange over NVLink (K-sharded replicas, one process per GPU; buffers are
// cudaMalloc'ed and mapped into every rank with CUDA IPC).  Rank r's buffer holds
//   flags [kMaxRanks] (128-byte stride): flags[s] = last epoch published by source rank s
//   inbox [2 parity][world][rowp] floats: slot (parity, s) = row published by source rank s
// world == 1 disables the exchange.
constexpr int kMaxRanks = 8;
constexpr int kFlagStrideInts = 32;
struct P2PParams {
    int world, rank;
    unsigned epoch;              // starts at 1, +1 per control step, same on every rank
    int rowp;                    // floats per inbox row (T*nu + 4, padded to a multiple of 4)
    float *base[kMaxRanks];      // base[r] = rank r's exchange buffer as mapped in THIS process
};
__host__ __device__ __forceinline__ int *p2p_flags(float *base) { return reinterpret_cast<int *>(base); }
__host__ __device__ __forceinline__ float *p2p_inbox(float *base, int world, int rowp, int parity, int src)
{
    return base + kMaxRanks * kFlagStrideInts + (static_cast<size_t>(parity) * world + src) * rowp;
}

template <int MODEL> struct ModelNu;
template <> struct ModelNu<MPPI_MODEL_DRONE3> { static constexpr int value = 3; };
template <> struct ModelNu<MPPI_MODEL_ARM7>   { static constexpr int value = 7; };
template <> struct ModelNu<MPPI_MODEL_QUAD4>  { static constexpr int value = 4; };
template <> struct ModelNu<MPPI_MODEL_WB11>   { static constexpr int value = 11; };

// ------------------------------------------------------------------------------------------
// Counter-based noise: Philox4x32-10 + Box-Muller.  New in this build (the reference calls
// torch.randn, S/sampling/standard_normal_noise.py:24).  Both the rollout pass and the
// weighting pass call normal4() with the same (sample, step, chunk) address, so the noise
// never has to exist in HBM; explicit _rn intrinsics keep the two call sites bit-identical
// regardless of how the surrounding code is contracted into FMAs.
// ------------------------------------------------------------------------------------------
// The ten round keys (key + r * Weyl constants) depend only on the seed: the host expands them
// once into StepParams::rkeys, so a round is 2 IMAD.WIDE + 2 LOP3 with constant-bank operands.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const uint32_t *rk)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ rk[2 * r], lo1, hi0 ^ c.w ^ rk[2 * r + 1], lo0);
    }
    return c;
}

// u1 = 2 - [1,2) in (0,1];  theta = ([1,2) - 1.5) * 2pi in [-pi,pi);  MUFU lg2 / sqrt / sin / cos.
// r = sqrt(-2 ln u1) = sqrt(lg2(u1) * (-2 ln 2)); sqrt.approx maps 0 -> 0 (u1 == 1).
__device__ __forceinline__ float sqrt_approx(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void box_muller(uint32_t x, uint32_t y, float &n0, float &n1)
{
    const float u1 = __fsub_rn(2.0f, __uint_as_float(0x3f800000u | (x >> 9)));
    const float th = __fmul_rn(__fsub_rn(__uint_as_float(0x3f800000u | (y >> 9)), 1.5f), kTwoPi);
    const float r = sqrt_approx(__fmul_rn(__log2f(u1), -1.3862943611198906f));
    float s, c;
    __sincosf(th, &s, &c);
    n0 = __fmul_rn(r, c);
    n1 = __fmul_rn(r, s);
}

// Four standard normals for inputs 4*chunk..4*chunk+3 of (global sample kg, horizon step t):
// counter = (kg, t*nch + chunk, step_lo, step_hi), key = seed.
__device__ __forceinline__ void normal4(uint32_t kg, uint32_t tc, uint32_t step_lo, uint32_t step_hi,
                                        const uint32_t *rkeys, float n[4])
{
    const uint4 r = philox4x32_10(make_uint4(kg, tc, step_lo, step_hi), rkeys);
    box_muller(r.x, r.y, n[0], n[1]);
    box_muller(r.z, r.w, n[2], n[3]);
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
// Small rigid-body helpers (row-major 3x3).
// ------------------------------------------------------------------------------------------
// Rz(yaw) Ry(pitch) Rx(roll)  (S/robot/transformation_matrix.py:4-25, :148-187; S/drone.py:126-154)
__device__ __forceinline__ void rpy_matrix(float sr, float cr, float sp, float cp, float sy, float cy, float R[9])
{
    R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

// xyz + quaternion xyzw, NOT normalised (S/robot/urdf_fk.py:30-55).
__device__ __forceinline__ void quat_matrix(const float *b, float R[9])
{
    const float qx = b[3], qy = b[4], qz = b[5], qw = b[6];
    R[0] = 1.0f - 2.0f * qy * qy - 2.0f * qz * qz; R[1] = 2.0f * qx * qy - 2.0f * qz * qw; R[2] = 2.0f * qx * qz + 2.0f * qy * qw;
    R[3] = 2.0f * qx * qy + 2.0f * qz * qw; R[4] = 1.0f - 2.0f * qx * qx - 2.0f * qz * qz; R[5] = 2.0f * qy * qz - 2.0f * qx * qw;
    R[6] = 2.0f * qx * qz - 2.0f * qy * qw; R[7] = 2.0f * qy * qz + 2.0f * qx * qw; R[8] = 1.0f - 2.0f * qx * qx - 2.0f * qy * qy;
}

// (R,p) <- (R Cr, p + R Ct)
__device__ __forceinline__ void compose_const(float R[9], float p[3], const float *Cr, const float *Ct)
{
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float a = R[3 * r], b = R[3 * r + 1], c = R[3 * r + 2];
        p[r] = fmaf(a, Ct[0], fmaf(b, Ct[1], fmaf(c, Ct[2], p[r])));
        R[3 * r]     = fmaf(a, Cr[0], fmaf(b, Cr[3], c * Cr[6]));
        R[3 * r + 1] = fmaf(a, Cr[1], fmaf(b, Cr[4], c * Cr[7]));
        R[3 * r + 2] = fmaf(a, Cr[2], fmaf(b, Cr[5], c * Cr[8]));
    }
}

// sin and cos of x to ~1 ulp (max abs error 1.4e-7) without the libm slow path: reduce by pi
// (2-term Cody-Waite under FMA) to r in [-pi/2, pi/2], near-minimax polynomials in r^2, one sign
// flip.  Valid for |x| < ~1e4 (joint and Euler angles here stay within a few turns).
__device__ __forceinline__ void sincos_pi(float x, float &s, float &c)
{
    const float kf = fmaf(x, 0.318309886183790672f, 12582912.0f);       // 1.5 * 2^23: rint in the low bits
    const int n = __float_as_int(kf);
    const float k = kf - 12582912.0f;
    float r = fmaf(k, -3.14159274101257324f, x);
    r = fmaf(k, 8.742278000372485e-8f, r);
    const float z = r * r;
    float ps = fmaf(2.634697921166662e-06f, z, -0.00019822725153062493f);
    ps = fmaf(ps, z, 0.008333242498338223f);
    ps = fmaf(ps, z, -0.1666666567325592f);
    float pc = fmaf(-2.62966580066859e-07f, z, 2.4774544726824388e-05f);
    pc = fmaf(pc, z, -0.0013888651737943292f);
    pc = fmaf(pc, z, 0.0416666604578495f);
    pc = fmaf(pc, z, -0.5f);
    const float sr = fmaf(ps, z * r, r);
    const float cr = fmaf(pc, z, 1.0f);
    const int flip = n << 31;                                             // odd multiple of pi: negate both
    s = __int_as_float(__float_as_int(sr) ^ flip);
    c = __int_as_float(__float_as_int(cr) ^ flip);
}

// Branch-free atan2 / asin / approximate reciprocal and square root for the pose cost.  Accuracy
// (max abs error vs double, measured in tools/fit_math.py): atan2 1.5e-7 + 2-ulp quotient, asin 1.6e-7;
// rcp / sqrt are the MUFU approximations (<= 2 ulp).  The reference's own float32 atan2/asin carry
// ~1e-7; these terms enter the cost multiplied by 30 against a float32 ulp of S of 1.2e-4.
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float atan2_poly(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(fmaxf(ax, ay), 1e-37f), mn = fminf(ax, ay);      // atan2(0, 0) -> 0
    const float t = mn * rcp_approx(mx);
    const float z = t * t;
    float p = fmaf(0.003962172195315361f, z, -0.02036452107131481f);
    p = fmaf(p, z, 0.049385011196136475f);
    p = fmaf(p, z, -0.0804230123758316f);
    p = fmaf(p, z, 0.1087741032242775f);
    p = fmaf(p, z, -0.14259037375450134f);
    p = fmaf(p, z, 0.19998809695243835f);
    p = fmaf(p, z, -0.33333325386047363f);
    float a = fmaf(p * z, t, t);
    a = (ay > ax) ? 1.57079632679489662f - a : a;
    a = (x < 0.0f) ? 3.14159265358979324f - a : a;
    return copysignf(a, y);
}
__device__ __forceinline__ float asin_poly(float x)
{
    const float a = fabsf(x);
    const bool big = a > 0.5f;
    const float z = big ? fmaf(a, -0.5f, 0.5f) : a * a;
    const float s = big ? sqrt_approx(z) : a;
    float p = fmaf(0.038206253200769424f, z, 0.02649438939988613f);
    p = fmaf(p, z, 0.045010678470134735f);
    p = fmaf(p, z, 0.07498808950185776f);
    p = fmaf(p, z, 0.16666673123836517f);
    float r = fmaf(p * z, s, s);
    r = big ? fmaf(-2.0f, r, 1.57079632679489662f) : r;
    return copysignf(r, x);
}

// ---- FK with compile-time constants (tables from tools/gen_fk_tables.py) -------------------
// acc (+)= x * k where k is a constant expression: zero terms vanish, +-1 become add / sub.
#define MPPI_CTERM(acc, have, x, k)                                                                   \
    if constexpr ((k) != 0.0f) {                                                                      \
        if constexpr (!(have)) acc = ((k) == 1.0f) ? (x) : ((k) == -1.0f) ? -(x) : (x) * (k);         \
        else acc = ((k) == 1.0f) ? acc + (x) : ((k) == -1.0f) ? acc - (x) : fmaf((x), (k), acc);      \
    }

template <class Tab, int J, int COL>
__device__ __forceinline__ float tab_rot_col(float a, float b, float c)
{
    constexpr float k0 = Tab::R[J][COL], k1 = Tab::R[J][3 + COL], k2 = Tab::R[J][6 + COL];
    constexpr bool h1 = (k0 != 0.0f), h2 = h1 || (k1 != 0.0f);
    float acc = 0.0f;
    MPPI_CTERM(acc, false, a, k0)
    MPPI_CTERM(acc, h1, b, k1)
    MPPI_CTERM(acc, h2, c, k2)
    return acc;
}
template <class Tab, int J>
__device__ __forceinline__ float tab_trans(float a, float b, float c, float p)
{
    constexpr float k0 = Tab::t[J][0], k1 = Tab::t[J][1], k2 = Tab::t[J][2];
    float acc = p;
    MPPI_CTERM(acc, true, a, k0)
    MPPI_CTERM(acc, true, b, k1)
    MPPI_CTERM(acc, true, c, k2)
    return acc;
}
// (R,p) <- (R C_J.R, p + R C_J.t) with C_J from the table
template <class Tab, int J>
__device__ __forceinline__ void compose_tab(float R[9], float p[3])
{
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float a = R[3 * r], b = R[3 * r + 1], c = R[3 * r + 2];
        p[r] = tab_trans<Tab, J>(a, b, c, p[r]);
        R[3 * r] = tab_rot_col<Tab, J, 0>(a, b, c);
        R[3 * r + 1] = tab_rot_col<Tab, J, 1>(a, b, c);
        R[3 * r + 2] = tab_rot_col<Tab, J, 2>(a, b, c);
    }
}
template <class Tab, int J = 0>
__device__ __forceinline__ void fk_tab(const float *cq, const float *sq, float R[9], float p[3])
{
    if constexpr (J < Tab::kJoints) {
        const float c = cq[J], s = sq[J];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float a = R[3 * r], b = R[3 * r + 1];
            R[3 * r] = fmaf(c, a, s * b);
            R[3 * r + 1] = fmaf(c, b, -s * a);
        }
        compose_tab<Tab, J + 1>(R, p);
        fk_tab<Tab, J + 1>(cq, sq, R, p);
    }
}

// Forward kinematics of the folded chain; on entry (R,p) = world pose of the chain root
// already composed with C0.  S/robot/urdfparser.py:133-161 + transformation_matrix.py:58-95
// collapse to "rotate columns 0/1 by q, then apply the next constant transform".
template <int NJ>
__device__ __forceinline__ void fk_chain(const ChainDev &ch, const float *cq, const float *sq, float R[9], float p[3])
{
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const float c = cq[j], s = sq[j];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float a = R[3 * r], b = R[3 * r + 1];
            R[3 * r]     = fmaf(c, a, s * b);
            R[3 * r + 1] = fmaf(c, b, -s * a);
        }
        if (j + 1 < NJ || !ch.last_identity) compose_const(R, p, ch.R[j + 1], ch.t[j + 1]);
    }
}

// ||p - p*||_2 and ||euler_ZYX(R^T R*)||_2  (S/cost/pose_cost.py:24-63,
// S/utils/rotation_conversions.py:277-319; inv(R) of a rotation is its transpose).
__device__ __forceinline__ void pose_terms(const float R[9], const float p[3], const DynBlock &D, float &pos, float &ori)
{
    const float *Tg = D.target_R;
    const float dx = p[0] - D.target_pos[0], dy = p[1] - D.target_pos[1], dz = p[2] - D.target_pos[2];
    pos = sqrt_approx(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    const float d00 = fmaf(R[0], Tg[0], fmaf(R[3], Tg[3], R[6] * Tg[6]));
    const float d10 = fmaf(R[1], Tg[0], fmaf(R[4], Tg[3], R[7] * Tg[6]));
    const float d20 = fmaf(R[2], Tg[0], fmaf(R[5], Tg[3], R[8] * Tg[6]));
    const float d21 = fmaf(R[2], Tg[1], fmaf(R[5], Tg[4], R[8] * Tg[7]));
    const float d22 = fmaf(R[2], Tg[2], fmaf(R[5], Tg[5], R[8] * Tg[8]));
    const float e0 = atan2_poly(d10, d00);
    const float e1 = asin_poly(fminf(fmaxf(-d20, -1.0f), 1.0f));
    const float e2 = atan2_poly(d21, d22);
    ori = sqrt_approx(fmaf(e0, e0, fmaf(e1, e1, e2 * e2)));
}

// Quadrotor rigid body, one step (restated from the dead draft S/mppi_solver/drone_mppi.py:57-83;
// rules fixed in DESIGN.md).  (s*, c*) are sin/cos of the CURRENT attitude on entry and of the
// NEW attitude on exit, so the whole-body FK reuses them.
struct QuadState {
    float p[3], rpy[3], v[3], w[3];
    float sphi, cphi, sth, cth, spsi, cpsi;
};

__device__ __forceinline__ float wrap_pi(float a)
{
    return (fabsf(a) > kPi) ? fmaf(-kTwoPi, rintf(a * kInvTwoPi), a) : a;
}

__device__ __forceinline__ void quad_advance(QuadState &s, float F, float tx, float ty, float tz,
                                             float dt, const float *qp)
{
    const float inv_m = rcp_approx(qp[0]), kd = qp[4], gz = qp[5];
    const float inv_cth = rcp_approx(s.cth);
    const float tth = s.sth * inv_cth;
    const float r02 = s.cpsi * s.sth * s.cphi + s.spsi * s.sphi;
    const float r12 = s.spsi * s.sth * s.cphi - s.cpsi * s.sphi;
    const float r22 = s.cth * s.cphi;
    s.w[0] = fmaf(dt, qp[1] * tx, s.w[0]);
    s.w[1] = fmaf(dt, qp[2] * ty, s.w[1]);
    s.w[2] = fmaf(dt, qp[3] * tz, s.w[2]);
    const float dphi = s.w[0] + s.sphi * tth * s.w[1] + s.cphi * tth * s.w[2];
    const float dth = s.cphi * s.w[1] - s.sphi * s.w[2];
    const float dpsi = (s.sphi * inv_cth) * s.w[1] + (s.cphi * inv_cth) * s.w[2];
    const float ax = (r02 * F - kd * s.v[0]) * inv_m;
    const float ay = (r12 * F - kd * s.v[1]) * inv_m;
    const float az = gz + (r22 * F - kd * s.v[2]) * inv_m;
    s.rpy[0] = wrap_pi(fmaf(dt, dphi, s.rpy[0]));
    s.rpy[1] = wrap_pi(fmaf(dt, dth, s.rpy[1]));
    s.rpy[2] = wrap_pi(fmaf(dt, dpsi, s.rpy[2]));
    s.v[0] = fmaf(dt, ax, s.v[0]); s.v[1] = fmaf(dt, ay, s.v[1]); s.v[2] = fmaf(dt, az, s.v[2]);
    s.p[0] = fmaf(dt, s.v[0], s.p[0]); s.p[1] = fmaf(dt, s.v[1], s.p[1]); s.p[2] = fmaf(dt, s.v[2], s.p[2]);
    sincos_pi(s.rpy[0], s.sphi, s.cphi);
    sincos_pi(s.rpy[1], s.sth, s.cth);
    sincos_pi(s.rpy[2], s.spsi, s.cpsi);
}

__device__ __forceinline__ void quad_load(QuadState &s, const float *st)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) { s.p[i] = st[i]; s.rpy[i] = st[3 + i]; s.v[i] = st[6 + i]; s.w[i] = st[9 + i]; }
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
