// mppi_dynamics.cuh -- the arm node's computed-torque law on the device.
//
// Replaces, for the arm rows the node reads (S/kinova.py:126-131,184):
//     pin.computeAllTerms(model, data, q, v);  g = data.nle
//     torque = data.M[6:, 6:] @ (400 * (qdes - q[7:]) + 40 * (-v[6:])) + g[6:]
// on the Pinocchio model of aerial_manipulation/urdf/full_robot_floating2.urdf (free-flyer base + seven
// revolute joints).  nle[6:] is one recursive Newton-Euler pass (zero base acceleration, gravity as an
// upward base acceleration, base twist from v[:6]); M[6:, 6:] does not depend on the base being free, and
// column j is one more Newton-Euler pass with unit acceleration on joint j.  There is no data parallelism
// beyond that, so the eight passes run on eight lanes of one warp of the finalize block, next to the
// other epilogue warps; M (400 (qdes - q) - 40 qdot) + nle is then a butterfly sum over those lanes.
#pragma once
#include "mppi_device.cuh"

namespace mppi {

__device__ __forceinline__ void cross3(const float a[3], const float b[3], float r[3])
{
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
// r = R^T x with R = R0 Rz(c, s) (child axes in the parent frame), R0 row-major
__device__ __forceinline__ void rot_t(const float *R0, float c, float s, const float x[3], float r[3])
{
    const float u0 = R0[0] * x[0] + R0[3] * x[1] + R0[6] * x[2];        // R0^T x
    const float u1 = R0[1] * x[0] + R0[4] * x[1] + R0[7] * x[2];
    r[2] = R0[2] * x[0] + R0[5] * x[1] + R0[8] * x[2];
    r[0] = c * u0 + s * u1;
    r[1] = c * u1 - s * u0;
}
// r = R x
__device__ __forceinline__ void rot(const float *R0, float c, float s, const float x[3], float r[3])
{
    const float u0 = c * x[0] - s * x[1], u1 = s * x[0] + c * x[1];      // Rz x
    r[0] = R0[0] * u0 + R0[1] * u1 + R0[2] * x[2];
    r[1] = R0[3] * u0 + R0[4] * u1 + R0[5] * x[2];
    r[2] = R0[6] * u0 + R0[7] * u1 + R0[8] * x[2];
}
// (n, f) = I (w, v): spatial inertia of mass m, centre of mass cm, inertia Ic = (xx xy xz yy yz zz) about cm
__device__ __forceinline__ void inertia_mul(float m, const float *cm, const float *Ic, const float w[3], const float v[3],
                                            float n[3], float f[3])
{
    const float h[3] = {m * cm[0], m * cm[1], m * cm[2]};
    const float cc = cm[0] * cm[0] + cm[1] * cm[1] + cm[2] * cm[2], cw = cm[0] * w[0] + cm[1] * w[1] + cm[2] * w[2];
    float hv[3], hw[3];
    cross3(h, v, hv);
    cross3(h, w, hw);
    n[0] = Ic[0] * w[0] + Ic[1] * w[1] + Ic[2] * w[2] + m * (cc * w[0] - cm[0] * cw) + hv[0];
    n[1] = Ic[1] * w[0] + Ic[3] * w[1] + Ic[4] * w[2] + m * (cc * w[1] - cm[1] * cw) + hv[1];
    n[2] = Ic[2] * w[0] + Ic[4] * w[1] + Ic[5] * w[2] + m * (cc * w[2] - cm[2] * cw) + hv[2];
    f[0] = m * v[0] - hw[0]; f[1] = m * v[1] - hw[1]; f[2] = m * v[2] - hw[2];
}

// One recursive Newton-Euler pass over the seven arm joints (Featherstone, RBDA table 5.1; joint axes +z of the
// folded chain): joint torques for joint rates qd, joint accelerations qdd, base twist (w0, v0) and base
// acceleration a0 (all in the base frame).
static __device__ __noinline__ void rnea_arm7(const ChainDev &ch, const ArmInertiaDev &in, const float *cq, const float *sq,
                                       const float *qd, const float *qdd, const float *w0, const float *v0, const float *a0,
                                       float *tau)
{
    float w[3] = {w0[0], w0[1], w0[2]}, v[3] = {v0[0], v0[1], v0[2]}, al[3] = {0.f, 0.f, 0.f}, a[3] = {a0[0], a0[1], a0[2]};
    float n[7][3], f[7][3];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const float *R0 = ch.R[i], *p = ch.t[i];
        float t1[3], t2[3], wn[3], vn[3], aln[3], an[3];
        rot_t(R0, cq[i], sq[i], w, wn);
        wn[2] += qd[i];
        cross3(w, p, t1);
        t1[0] += v[0]; t1[1] += v[1]; t1[2] += v[2];
        rot_t(R0, cq[i], sq[i], t1, vn);
        rot_t(R0, cq[i], sq[i], al, aln);
        aln[0] += wn[1] * qd[i]; aln[1] -= wn[0] * qd[i]; aln[2] += qdd[i];         // + z qdd + w x (z qd)
        cross3(al, p, t2);
        t2[0] += a[0]; t2[1] += a[1]; t2[2] += a[2];
        rot_t(R0, cq[i], sq[i], t2, an);
        an[0] += vn[1] * qd[i]; an[1] -= vn[0] * qd[i];                               // + v x (z qd)
        float hn[3], hf[3], In[3], If[3], x1[3], x2[3], x3[3];
        inertia_mul(in.mass[i], in.com[i], in.inertia[i], wn, vn, hn, hf);
        inertia_mul(in.mass[i], in.com[i], in.inertia[i], aln, an, In, If);
        cross3(wn, hn, x1); cross3(vn, hf, x2); cross3(wn, hf, x3);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            n[i][k] = In[k] + x1[k] + x2[k];
            f[i][k] = If[k] + x3[k];
            w[k] = wn[k]; v[k] = vn[k]; al[k] = aln[k]; a[k] = an[k];
        }
    }
#pragma unroll
    for (int i = 6; i >= 0; --i) {
        tau[i] = n[i][2];
        if (i > 0) {
            float fp[3], np_[3], pf[3];
            rot(ch.R[i], cq[i], sq[i], f[i], fp);
            rot(ch.R[i], cq[i], sq[i], n[i], np_);
            cross3(ch.t[i], fp, pf);
#pragma unroll
            for (int k = 0; k < 3; ++k) { n[i - 1][k] += np_[k] + pf[k]; f[i - 1][k] += fp[k]; }
        }
    }
}

// Called by ONE full warp.  q / qd: measured joint state; a0: gravity as an upward base acceleration in the base frame;
// v0, w0: base twist (linear, angular; base frame); dq = qdes - q.  Writes torque[7] = M (kp dq - kd qdot) + nle.
__device__ __forceinline__ void arm_torque_warp(const StepParams &P, const float *q, const float *qdot, const float *a0_in,
                                                const float *v0_in, const float *w0_in, const float *dq, float *torque)
{
    const int lane = threadIdx.x & 31;
    float cq[7], sq[7], qd[7], qdd[7], w0[3], v0[3], a0[3];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        sincos_pi(q[i], sq[i], cq[i]);
        qd[i] = (lane == 0) ? qdot[i] : 0.f;                  // lane 0: nle = rnea(q, v, 0); lanes 1..7: columns of M
        qdd[i] = (lane == i + 1) ? 1.f : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a0[k] = (lane == 0) ? a0_in[k] : 0.f;
        v0[k] = (lane == 0) ? v0_in[k] : 0.f;
        w0[k] = (lane == 0) ? w0_in[k] : 0.f;
    }
    float tau[7];
    rnea_arm7(P.chain, P.arm_inertia, cq, sq, qd, qdd, w0, v0, a0, tau);
    // lane 0 holds nle, lane j holds column j - 1 of M: scale the columns by the desired acceleration and sum lanes 0..7
    const int j = (lane >= 1 && lane <= 7) ? lane - 1 : 0;
    const float ades = P.arm_inertia.kp * dq[j] - P.arm_inertia.kd * qdot[j];
    const float scale = (lane == 0) ? 1.0f : (lane <= 7 ? ades : 0.0f);
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        float t = tau[i] * scale;
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        if (lane == 0) torque[i] = t;
    }
}

// ARM7 state: q[7], qdot[7], base xyz + quat xyzw, base twist (linear, angular; base frame -- Pinocchio's v[:6]).
__device__ __forceinline__ void arm_torque_from_arm_state(const StepParams &P, const float *state, const float *dq, float *torque)
{
    // R_base^T (0, 0, g) = g * (third row of R); R from the normalised quaternion
    const float x = state[17], y = state[18], z = state[19], wq = state[20];
    const float s2 = 2.0f / fmaxf(x * x + y * y + z * z + wq * wq, 1e-30f), g = P.arm_inertia.gravity;
    const float a0[3] = {g * s2 * (x * z - y * wq), g * s2 * (y * z + x * wq), g * (1.0f - s2 * (x * x + y * y))};
    arm_torque_warp(P, state, state + 7, a0, state + 21, state + 24, dq, torque);
}

// WB11 state: p[3], rpy[3], v[3] (world), w[3] (body rates), q[7], qdot[7]: attitude R = Rz(yaw) Ry(pitch) Rx(roll),
// base twist = (R^T v, w).
__device__ __forceinline__ void arm_torque_from_wb_state(const StepParams &P, const float *state, const float *dq, float *torque)
{
    float sr, cr, sp, cp, sy, cy, R[9];
    sincos_pi(state[3], sr, cr); sincos_pi(state[4], sp, cp); sincos_pi(state[5], sy, cy);
    rpy_matrix(sr, cr, sp, cp, sy, cy, R);
    const float g = P.arm_inertia.gravity;
    const float a0[3] = {g * R[6], g * R[7], g * R[8]};
    const float *v = state + 6;
    const float v0[3] = {R[0] * v[0] + R[3] * v[1] + R[6] * v[2], R[1] * v[0] + R[4] * v[1] + R[7] * v[2],
                         R[2] * v[0] + R[5] * v[1] + R[8] * v[2]};
    arm_torque_warp(P, state + 12, state + 19, a0, v0, state + 9, dq, torque);
}

}  // namespace mppi
