/*
 * mppi_oracle.c -- CPU restatement of the reference MPPI control step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may call, link or
 * import this file.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py use it, and only as the checker / the CPU
 * baseline -- never as the thing shipped.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - arm  (nu=7)  : PINNED  against tests/golden/arm_*.npz  produced by running the
 *                    unmodified reference classes (oracle/make_golden.py).
 *   - drone(nu=3)  : PINNED  against tests/golden/drone_*.npz, same way.
 *   - quad (nu=4)  : PARITY UNPINNED -- the reference only holds a commented-out,
 *                    non-runnable draft (mppi_solver/drone_mppi.py:57-83); this
 *                    restatement *is* the specification.
 *   - wb   (nu=11) : PARITY UNPINNED -- not implemented in the reference at all
 *                    (README.md:33); composed here from the pinned arm pieces, the
 *                    reference's base_movement hook (robot/urdfparser.py:128-131) and
 *                    the quad draft.
 *
 * All paths "S/..." below are relative to
 *   /root/reference/src/mav_mppi/scripts/
 * Arithmetic is IEEE float32 evaluated in the reference's operation order wherever
 * that order is defined by the reference (element-wise torch ops); compile with
 * -ffp-contract=off so gcc does not fuse what torch keeps separate.  Reductions
 * whose order the reference leaves to the library (torch.sum, matmul inner products
 * inside BLAS, the LU inverse) are evaluated here in double and rounded once, i.e.
 * the oracle is the correctly-rounded value the reference approximates.
 *
 * Noise layout at this boundary is [T][K][nu] (the native boundary layout); the
 * reference's [K][T][nu] (S/sampling/standard_normal_noise.py:24) is transposed by
 * the caller.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAX_JOINTS 16
#define ORACLE_MAX_NU 16

/* ------------------------------------------------------------------ threading */
int oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ------------------------------------------------------------------ 4x4 helpers */
static void mat4_identity(float *T)
{
    memset(T, 0, 16 * sizeof(float));
    T[0] = T[5] = T[10] = T[15] = 1.0f;
}

/* C = A @ B for 4x4 (torch.matmul; inner-product order is BLAS-defined -> double). */
static void mat4_mul(const float *A, const float *B, float *C)
{
    float out[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += (double)A[4 * i + k] * (double)B[4 * k + j];
            out[4 * i + j] = (float)acc;
        }
    memcpy(C, out, sizeof(out));
}

/* S/robot/transformation_matrix.py:4-25  rotation_matrix_rpy: R = Rz(yaw) Ry(pitch) Rx(roll),
 * every product left-to-right in float32. */
static void rpy_to_R(float roll, float pitch, float yaw, float R[9])
{
    float cr = cosf(roll), sr = sinf(roll);
    float cp = cosf(pitch), sp = sinf(pitch);
    float cy = cosf(yaw), sy = sinf(yaw);
    R[0] = cy * cp;
    R[1] = cy * sp * sr - sy * cr;
    R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp;
    R[4] = sy * sp * sr + cy * cr;
    R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;
    R[7] = cp * sr;
    R[8] = cp * cr;
}

/* S/robot/transformation_matrix.py:28-35  make_transform_matrix */
void oracle_make_transform(const float xyz[3], const float rpy[3], float T[16])
{
    float R[9];
    mat4_identity(T);
    rpy_to_R(rpy[0], rpy[1], rpy[2], R);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
        T[4 * i + 3] = xyz[i];
    }
}

/* S/robot/transformation_matrix.py:58-95  revolute_transform: tf_origin @ tf_rot with the
 * Rodrigues rotation about the normalised axis; (c, s) are supplied by the caller so that
 * the float64-state variant (sin/cos evaluated in f64, stored as f32) can
 * share this code. */
static void revolute_local(const float xyz[3], const float rpy[3], const float axis_in[3],
                           float c, float s, float T[16])
{
    float O[16], Rt[16];
    float ax[3] = {axis_in[0], axis_in[1], axis_in[2]};
    float n = sqrtf(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    if (n < 1e-12f) { ax[0] = 1.0f; ax[1] = 0.0f; ax[2] = 0.0f; }
    else { ax[0] /= n; ax[1] /= n; ax[2] /= n; }
    float omc = 1.0f - c;
    float vx = ax[0], vy = ax[1], vz = ax[2];
    oracle_make_transform(xyz, rpy, O);
    mat4_identity(Rt);
    Rt[0] = c + vx * vx * omc;       Rt[1] = vx * vy * omc - vz * s;  Rt[2] = vx * vz * omc + vy * s;
    Rt[4] = vy * vx * omc + vz * s;  Rt[5] = c + vy * vy * omc;       Rt[6] = vy * vz * omc - vx * s;
    Rt[8] = vz * vx * omc - vy * s;  Rt[9] = vz * vy * omc + vx * s;  Rt[10] = c + vz * vz * omc;
    mat4_mul(O, Rt, T);
}

/* S/robot/transformation_matrix.py:38-55  prismatic_transform */
static void prismatic_local(const float xyz[3], const float rpy[3], const float axis_in[3],
                            float q, float T[16])
{
    float O[16], Sl[16];
    float ax[3] = {axis_in[0], axis_in[1], axis_in[2]};
    float n = sqrtf(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    if (n < 1e-12f) { ax[0] = 1.0f; ax[1] = 0.0f; ax[2] = 0.0f; }
    else { ax[0] /= n; ax[1] /= n; ax[2] /= n; }
    oracle_make_transform(xyz, rpy, O);
    mat4_identity(Sl);
    Sl[3] = ax[0] * q; Sl[7] = ax[1] * q; Sl[11] = ax[2] * q;
    mat4_mul(O, Sl, T);
}

/* Chain description, flat arrays so ctypes can pass numpy buffers:
 *   jtype[j] : 0 fixed, 1 revolute/continuous, 2 prismatic
 *   qidx[j]  : index into the joint vector, or -1
 *   xyz,rpy,axis : [n_joints][3]                                                       */
typedef struct {
    int n_joints;
    const int *jtype;
    const int *qidx;
    const float *xyz, *rpy, *axis;
} chain_t;

/* S/robot/urdfparser.py:122-163  forward_kinematics (base_movement handled by callers):
 * tf_fk <- tf_fk @ tf_local, left to right along the chain.  cq/sq hold cos/sin of the
 * joint values already rounded to f32; q itself is only used for prismatic joints. */
static void fk_chain(const chain_t *ch, const float *q, const float *cq, const float *sq,
                     const float *T0, float T[16])
{
    float acc[16], loc[16];
    memcpy(acc, T0, sizeof(acc));
    for (int j = 0; j < ch->n_joints; ++j) {
        const float *xyz = ch->xyz + 3 * j, *rpy = ch->rpy + 3 * j, *axis = ch->axis + 3 * j;
        int qi = ch->qidx[j];
        if (ch->jtype[j] == 1) {
            float c = qi >= 0 ? cq[qi] : 1.0f, s = qi >= 0 ? sq[qi] : 0.0f;
            revolute_local(xyz, rpy, axis, c, s, loc);
        } else if (ch->jtype[j] == 2) {
            prismatic_local(xyz, rpy, axis, qi >= 0 ? q[qi] : 0.0f, loc);
        } else {
            oracle_make_transform(xyz, rpy, loc);
        }
        mat4_mul(acc, loc, acc);
    }
    memcpy(T, acc, sizeof(acc));
}

/* Public single-pose FK (float32 joint values), used by the FK known-answer tests. */
void oracle_fk(int n_joints, const int *jtype, const int *qidx, const float *xyz,
               const float *rpy, const float *axis, const float *q, int nq, float T[16])
{
    chain_t ch = {n_joints, jtype, qidx, xyz, rpy, axis};
    float cq[ORACLE_MAX_NU], sq[ORACLE_MAX_NU], I[16];
    for (int i = 0; i < nq; ++i) { cq[i] = cosf(q[i]); sq[i] = sinf(q[i]); }
    mat4_identity(I);
    fk_chain(&ch, q, cq, sq, I, T);
}

/* S/robot/urdf_fk.py:30-55  xyzquat_to_matrix -- quaternion xyzw, NOT normalised here. */
void oracle_xyzquat_to_matrix(const float b[7], float T[16])
{
    float qx = b[3], qy = b[4], qz = b[5], qw = b[6];
    mat4_identity(T);
    T[3] = b[0]; T[7] = b[1]; T[11] = b[2];
    /* the reference evaluates these in Python on 0-dim tensors: float32 scalars */
    T[0] = 1.0f - 2.0f * (qy * qy) - 2.0f * (qz * qz);
    T[1] = 2.0f * qx * qy - 2.0f * qz * qw;
    T[2] = 2.0f * qx * qz + 2.0f * qy * qw;
    T[4] = 2.0f * qx * qy + 2.0f * qz * qw;
    T[5] = 1.0f - 2.0f * (qx * qx) - 2.0f * (qz * qz);
    T[6] = 2.0f * qy * qz - 2.0f * qx * qw;
    T[8] = 2.0f * qx * qz - 2.0f * qy * qw;
    T[9] = 2.0f * qy * qz + 2.0f * qx * qw;
    T[10] = 1.0f - 2.0f * (qx * qx) - 2.0f * (qy * qy);
}

/* S/robot/transformation_matrix.py:148-187  transformation_matrix_from_xyzrpy
 * (the base_movement hook, urdfparser.py:128-131). */
void oracle_xyzrpy_to_matrix(const float x[6], float T[16])
{
    float R[9];
    mat4_identity(T);
    rpy_to_R(x[3], x[4], x[5], R);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
        T[4 * i + 3] = x[i];
    }
}

/* S/utils/rotation_conversions.py:45-75  quaternion_to_matrix, patched to xyzw (:57),
 * normalising through two_s = 2/|q|^2. */
void oracle_quaternion_to_matrix(const float q[4], float R[9])
{
    float i = q[0], j = q[1], k = q[2], r = q[3];
    float two_s = 2.0f / (i * i + j * j + k * k + r * r);
    R[0] = 1.0f - two_s * (j * j + k * k);
    R[1] = two_s * (i * j - k * r);
    R[2] = two_s * (i * k + j * r);
    R[3] = two_s * (i * j + k * r);
    R[4] = 1.0f - two_s * (i * i + k * k);
    R[5] = two_s * (j * k - i * r);
    R[6] = two_s * (i * k - j * r);
    R[7] = two_s * (j * k + i * r);
    R[8] = 1.0f - two_s * (i * i + j * j);
}

/* S/utils/rotation_conversions.py:277-319  matrix_to_euler_angles(M, "ZYX"):
 * (atan2(M10, M00), asin(clamp(-M20, -1, 1)), atan2(M21, M22)). */
void oracle_matrix_to_euler_zyx(const float M[9], float e[3])
{
    float v = -M[6];
    if (v < -1.0f) v = -1.0f;
    if (v > 1.0f) v = 1.0f;
    e[0] = atan2f(M[3], M[0]);
    e[1] = asinf(v);
    e[2] = atan2f(M[7], M[8]);
}

/* torch.linalg.inv on a 3x3 (S/cost/pose_cost.py:32,54): the exact inverse, evaluated in
 * double by cofactors and rounded once (the reference's LU is a library-ordered
 * approximation of the same value). */
static void inv3(const float *A, float *Ai)
{
    double a = A[0], b = A[1], c = A[2], d = A[3], e = A[4], f = A[5], g = A[6], h = A[7], i = A[8];
    double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    double r = 1.0 / det;
    Ai[0] = (float)((e * i - f * h) * r); Ai[1] = (float)((c * h - b * i) * r); Ai[2] = (float)((b * f - c * e) * r);
    Ai[3] = (float)((f * g - d * i) * r); Ai[4] = (float)((a * i - c * g) * r); Ai[5] = (float)((c * d - a * f) * r);
    Ai[6] = (float)((d * h - e * g) * r); Ai[7] = (float)((b * g - a * h) * r); Ai[8] = (float)((a * e - b * d) * r);
}

/* S/cost/pose_cost.py:24-43 (stage) and :46-63 (terminal) share this per-pose term:
 *   ||p - p*||_2  and  ||euler_ZYX(inv(R) @ R*)||_2     (L2 norms, not squared). */
void oracle_pose_terms(const float T[16], const float tgt_p[3], const float tgt_R[9],
                       float *pos_norm, float *ori_norm)
{
    float R[9], Ri[9], D[9], e[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = T[4 * i + j];
    inv3(R, Ri);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 3; ++k) acc += (double)Ri[3 * i + k] * (double)tgt_R[3 * k + j];
            D[3 * i + j] = (float)acc;
        }
    oracle_matrix_to_euler_zyx(D, e);
    float dx = T[3] - tgt_p[0], dy = T[7] - tgt_p[1], dz = T[11] - tgt_p[2];
    *pos_norm = (float)sqrt((double)dx * dx + (double)dy * dy + (double)dz * dz);
    *ori_norm = (float)sqrt((double)e[0] * e[0] + (double)e[1] * e[1] + (double)e[2] * e[2]);
}

/* ------------------------------------------------------------------ arm (nu = 7) */
/* One sample of S/mppi_solver/mppi.py:129-140:
 *   v = u + noise                                   (mppi.py:130)
 *   double integrator                               (S/sampling/standard_normal_noise.py:32-50)
 *       V_t  = cumsum(a*dt)_t + qd0
 *       dq_t = V_{t-1}*dt + 0.5*a_t*dt^2            (V_{-1} = qd0)
 *       Q_t  = cumsum(dq)_t + q0
 *   FK world = B(base) @ chain(Q_t)                 (S/robot/urdf_fk.py:79-108)
 *   S = sum_{t<T-1} (w0*pos + w1*ori) + (w2*pos + w3*ori)_{T-1}   (S/cost/cost_manager.py:78-89)
 * state_f64 != 0 reproduces update_joint's float64 state tensors (mppi.py:196-200, F8):
 * the integration and sin/cos run in double and are rounded to f32 when written into
 * the 4x4 (transformation_matrix.py:92-93).                                            */
static float arm_sample_cost(int T, int K, int k, int nu, const float *noise, const float *u_nom,
                             const double *q0, const double *qd0, const float *Bm,
                             const chain_t *ch, const float *tgt_p, const float *tgt_R,
                             float dt, const float *wts, int state_f64)
{
    const float dt2 = (float)((double)dt * (double)dt); /* python: dt**2 on a float */
    double S = 0.0;
    float I4[16];
    mat4_identity(I4);
    if (!state_f64) {
        /* torch.cumsum on CPU accumulates float32 inputs in double and rounds every
         * output element to float32 (acc_type<float,false> = double); verified by probe. */
        double cum_v[ORACLE_MAX_NU], cum_q[ORACLE_MAX_NU];
        float vprev[ORACLE_MAX_NU];
        float q[ORACLE_MAX_NU], cq[ORACLE_MAX_NU], sq[ORACLE_MAX_NU];
        for (int i = 0; i < nu; ++i) { cum_v[i] = 0.0; cum_q[i] = 0.0; vprev[i] = (float)qd0[i]; }
        for (int t = 0; t < T; ++t) {
            const float *eps = noise + ((size_t)t * K + k) * nu;
            for (int i = 0; i < nu; ++i) {
                float a = u_nom[t * nu + i] + eps[i];
                float dq = vprev[i] * dt + 0.5f * a * dt2;
                cum_v[i] += (double)(a * dt);
                vprev[i] = (float)cum_v[i] + (float)qd0[i];
                cum_q[i] += (double)dq;
                q[i] = (float)cum_q[i] + (float)q0[i];
                cq[i] = cosf(q[i]); sq[i] = sinf(q[i]);
            }
            float Tr[16], Tw[16], pn, on;
            fk_chain(ch, q, cq, sq, I4, Tr);
            mat4_mul(Bm, Tr, Tw);
            oracle_pose_terms(Tw, tgt_p, tgt_R, &pn, &on);
            float c = (t < T - 1) ? (wts[0] * pn + wts[1] * on) : (wts[2] * pn + wts[3] * on);
            S += (double)c;
        }
    } else {
        double cum_v[ORACLE_MAX_NU], cum_q[ORACLE_MAX_NU], vprev[ORACLE_MAX_NU];
        float q[ORACLE_MAX_NU], cq[ORACLE_MAX_NU], sq[ORACLE_MAX_NU];
        const double dtd = (double)dt; /* samples(f32) * dt promotes through the f64 state */
        for (int i = 0; i < nu; ++i) { cum_v[i] = 0.0; cum_q[i] = 0.0; vprev[i] = qd0[i]; }
        for (int t = 0; t < T; ++t) {
            const float *eps = noise + ((size_t)t * K + k) * nu;
            for (int i = 0; i < nu; ++i) {
                /* samples stay f32: cumsum(samples*dt) is an f32 cumsum, promoted by "+ qdot0" */
                float a = u_nom[t * nu + i] + eps[i];
                double dq = vprev[i] * dtd + (double)(0.5f * a * dt2);
                cum_v[i] += (double)(a * dt);
                vprev[i] = (double)(float)cum_v[i] + qd0[i];
                cum_q[i] = cum_q[i] + dq;
                double qq = cum_q[i] + q0[i];
                q[i] = (float)qq; cq[i] = (float)cos(qq); sq[i] = (float)sin(qq);
            }
            float Tr[16], Tw[16], pn, on;
            fk_chain(ch, q, cq, sq, I4, Tr);
            mat4_mul(Bm, Tr, Tw);
            oracle_pose_terms(Tw, tgt_p, tgt_R, &pn, &on);
            float c = (t < T - 1) ? (wts[0] * pn + wts[1] * on) : (wts[2] * pn + wts[3] * on);
            S += (double)c;
        }
    }
    return (float)S;
}

void oracle_arm_costs(int K, int T, const float *noise /*[T][K][7]*/, const float *u_nom /*[T][7]*/,
                      const double *q0, const double *qd0, const float *base /*[7] xyz+quat xyzw*/,
                      int n_joints, const int *jtype, const int *qidx, const float *xyz,
                      const float *rpy, const float *axis,
                      const float *tgt_p, const float *tgt_quat, float dt,
                      const float *wts /*[4] stage pos, stage ori, term pos, term ori*/,
                      int state_f64, float *S_out /*[K]*/)
{
    chain_t ch = {n_joints, jtype, qidx, xyz, rpy, axis};
    float Bm[16], tgt_R[9];
    oracle_xyzquat_to_matrix(base, Bm);
    oracle_quaternion_to_matrix(tgt_quat, tgt_R);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; ++k)
        S_out[k] = arm_sample_cost(T, K, k, 7, noise, u_nom, q0, qd0, Bm, &ch, tgt_p, tgt_R, dt,
                                   wts, state_f64);
}

/* The cost terms the reference constructs but leaves commented out of the sum
 * (S/cost/cost_manager.py:83-87), evaluated on the same rollout.  flags: 1 covar
 * (S/cost/covar_cost.py:20-25), 2 centering, 4 joint tracking, 16 joint limit
 * (S/cost/joint_space_cost.py:18-77), 8 action (S/cost/action_cost.py:15-25).
 * ext = [gamma, covar_weight, lambda, alpha, action_w, centering_w, joint_traj_w, limit_penalty].
 * PINNED by tests/golden/arm_extra_costs.npz (reference run with those lines re-enabled). */
void oracle_arm_extra_costs(int K, int T, const float *noise /*[T][K][7]*/, const float *u_nom /*[T][7]*/,
                            const float *q0, const float *qd0, float dt, int flags, const float *ext,
                            const float *sigma /*[7]*/, const float *q_center, const float *q_lower,
                            const float *q_upper, const float *q_traj /*[T][7] or NULL*/, float *S_out /*[K]*/)
{
    const float dt2 = (float)((double)dt * (double)dt);
    const float gamma = ext[0];
    const double covar_scale = (double)ext[1] * ((double)ext[2] * (1.0 - (double)ext[3]));   /* python floats */
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; ++k) {
        double cum_v[7] = {0}, cum_q[7] = {0};
        float vprev[7];
        for (int i = 0; i < 7; ++i) vprev[i] = qd0[i];
        double cov = 0, cen = 0, trk = 0, act = 0, lim = 0;
        for (int t = 0; t < T; ++t) {
            const float *eps = noise + ((size_t)t * K + k) * 7;
            const float g = powf(gamma, (float)t);               /* gamma ** torch.arange(T) */
            double c1 = 0, a1 = 0, ce = 0, tr = 0;
            int oob = 0;
            for (int i = 0; i < 7; ++i) {
                float a = u_nom[t * 7 + i] + eps[i];
                float dq = vprev[i] * dt + 0.5f * a * dt2;
                cum_v[i] += (double)(a * dt);
                vprev[i] = (float)cum_v[i] + qd0[i];
                cum_q[i] += (double)dq;
                float q = (float)cum_q[i] + q0[i];
                c1 += (double)u_nom[t * 7 + i] * (double)((1.0f / sigma[i]) * a);
                a1 += (double)(a * a);
                float dc = q - q_center[i];
                ce += (double)(dc * dc);
                float dk = q - (q_traj ? q_traj[t * 7 + i] : 0.0f);
                tr += (double)(dk * dk);
                if (q < q_lower[i] || q > q_upper[i]) oob = 1;
            }
            cov += (double)(float)c1;
            act += (double)((ext[4] * (float)a1) * g);
            cen += (double)((ext[5] * (float)ce) * g);
            trk += (double)((ext[6] * (float)tr) * g);
            if (oob) lim += (double)(ext[7] * g);
        }
        float S = 0.0f;
        if (flags & 1) S += (float)(covar_scale * (double)(float)cov);
        if (flags & 2) S += (float)cen;
        if (flags & 4) S += (float)trk;
        if (flags & 8) S += (float)act;
        if (flags & 16) S += (float)lim;
        S_out[k] = S;
    }
}

/* ------------------------------------------------------------------ drone (nu = 3) */
/* S/mppi_solver/drone_mppi.py:46-55 (predict_trajectory) and :87-107 (costs):
 *   S = 100 * sum_{t<T-1} |x_t - x*|^2 + 20 * |x_{T-1} - x*|^2                          */
void oracle_drone_costs(int K, int T, const float *noise /*[T][K][3]*/, const float *u_nom,
                        const float *x0, const float *v0, const float *target, float dt,
                        const float *wts /*[2] stage, terminal*/, float *S_out)
{
    const float dt2 = (float)((double)dt * (double)dt);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; ++k) {
        double cum_v[3] = {0, 0, 0}, cum_q[3] = {0, 0, 0};   /* CPU cumsum: double accumulator */
        float vprev[3] = {v0[0], v0[1], v0[2]};
        double stage = 0.0;
        float term = 0.0f;
        for (int t = 0; t < T; ++t) {
            const float *eps = noise + ((size_t)t * K + k) * 3;
            double sq = 0.0;
            for (int i = 0; i < 3; ++i) {
                float a = eps[i] + u_nom[t * 3 + i];           /* v = noise + u  (:144) */
                float dq = vprev[i] * dt + 0.5f * a * dt2;     /* (:53) */
                cum_v[i] += (double)(a * dt);                  /* (:49) */
                vprev[i] = (float)cum_v[i] + v0[i];
                cum_q[i] += (double)dq;                        /* (:54) */
                float x = (float)cum_q[i] + x0[i];
                float e = x - target[i];
                sq += (double)(e * e);                         /* torch.pow(err,2).sum */
            }
            if (t < T - 1) stage += (double)(float)sq;
            else term = (float)sq;
        }
        S_out[k] = (float)stage * wts[0] + term * wts[1];
    }
}

/* ------------------------------------------------------------------ quad (nu = 4), UNPINNED */
/* Rigid-body quadrotor rollout restated from the commented-out draft
 * S/mppi_solver/drone_mppi.py:57-83 (helpers S/drone.py:114-154), with the draft's
 * inconsistencies resolved as documented in DESIGN.md:
 *   u_t = (F, tau_x, tau_y, tau_z)
 *   w_t   = w_{t-1} + dt * Iinv * tau_t                               (:72)
 *   rpy_t = wrap(rpy_{t-1} + dt * J(rpy_{t-1}) w_t)                   (:73-77; J = Euler-rate map
 *                                                                     S/drone.py:114-124)
 *   v_t   = v_{t-1} + dt * (g + (R(rpy_{t-1}) [0,0,F_t] - kd v_{t-1}) / m)   (:75,78)
 *   p_t   = p_{t-1} + dt * v_t                                        (:79)
 * with "t-1 = -1" meaning the measured state, i.e. the loop rule is used for t = 0 too.
 * state = [p(3), rpy(3), v(3), w(3)]                                                   */
typedef struct { float p[3], rpy[3], v[3], w[3]; } quad_state_t;

static float wrap_pi(float a) { return atan2f(sinf(a), cosf(a)); }   /* (:77) */

static void quad_advance(quad_state_t *s, const float u[4], float dt, float mass,
                         const float Iinv[3], float kd, float gz)
{
    float phi = s->rpy[0], th = s->rpy[1], psi = s->rpy[2];
    float sphi = sinf(phi), cphi = cosf(phi), sth = sinf(th), cth = cosf(th);
    float spsi = sinf(psi), cpsi = cosf(psi);
    float tth = sth / cth;
    /* third column of R_zyx(rpy_{t-1}) (S/drone.py:126-154) times F */
    float r02 = cpsi * sth * cphi + spsi * sphi;
    float r12 = spsi * sth * cphi - cpsi * sphi;
    float r22 = cth * cphi;
    float w0 = s->w[0] + dt * (Iinv[0] * u[1]);
    float w1 = s->w[1] + dt * (Iinv[1] * u[2]);
    float w2 = s->w[2] + dt * (Iinv[2] * u[3]);
    /* J(phi,theta) @ w  (S/drone.py:114-124) */
    float dphi = w0 + sphi * tth * w1 + cphi * tth * w2;
    float dth = cphi * w1 - sphi * w2;
    float dpsi = (sphi / cth) * w1 + (cphi / cth) * w2;
    float F = u[0];
    float ax = (r02 * F - kd * s->v[0]) / mass;
    float ay = (r12 * F - kd * s->v[1]) / mass;
    float az = gz + (r22 * F - kd * s->v[2]) / mass;
    s->w[0] = w0; s->w[1] = w1; s->w[2] = w2;
    s->rpy[0] = wrap_pi(phi + dt * dphi);
    s->rpy[1] = wrap_pi(th + dt * dth);
    s->rpy[2] = wrap_pi(psi + dt * dpsi);
    s->v[0] = s->v[0] + dt * ax;
    s->v[1] = s->v[1] + dt * ay;
    s->v[2] = s->v[2] + dt * az;
    s->p[0] = s->p[0] + dt * s->v[0];
    s->p[1] = s->p[1] + dt * s->v[1];
    s->p[2] = s->p[2] + dt * s->v[2];
}

/* params = [mass, Iinv_x, Iinv_y, Iinv_z, kd, g_z]; cost = the live drone cost
 * (drone_mppi.py:87-107) on the position. */
void oracle_quad_costs(int K, int T, const float *noise /*[T][K][4]*/, const float *u_nom,
                       const float *state /*[12] p,rpy,v,w*/, const float *target, float dt,
                       const float *params, const float *wts /*[2]*/, float *S_out)
{
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; ++k) {
        quad_state_t s;
        memcpy(&s, state, sizeof(s));
        double stage = 0.0;
        float term = 0.0f;
        for (int t = 0; t < T; ++t) {
            const float *eps = noise + ((size_t)t * K + k) * 4;
            float u[4];
            for (int i = 0; i < 4; ++i) u[i] = u_nom[t * 4 + i] + eps[i];
            quad_advance(&s, u, dt, params[0], params + 1, params[4], params[5]);
            float ex = s.p[0] - target[0], ey = s.p[1] - target[1], ez = s.p[2] - target[2];
            float sq = (float)((double)(ex * ex) + (double)(ey * ey) + (double)(ez * ez));
            if (t < T - 1) stage += (double)sq;
            else term = sq;
        }
        S_out[k] = (float)stage * wts[0] + term * wts[1];
    }
}

/* ------------------------------------------------------------------ whole body (nu = 11), UNPINNED */
/* u_t = (F, tau_xyz, qdd_1..7).  Quad rollout as above gives the moving base (p_t, rpy_t);
 * the arm integrates as in the pinned arm path; the end-effector is
 *   T_world = T(p_t, rpy_t) @ chain(Q_t)   (base_movement semantics, urdfparser.py:128-131,
 *                                           transformation_matrix.py:148-187)
 * and S = arm pose cost (cost_manager.py:30-33 weights) + drone position cost
 * (drone_mppi.py:93,105 weights) on p_t.  wts = [stage pos, stage ori, term pos, term ori,
 * drone stage, drone terminal].                                                        */
void oracle_wb_costs(int K, int T, const float *noise /*[T][K][11]*/, const float *u_nom,
                     const float *qstate /*[12]*/, const float *q0f, const float *qd0f,
                     int n_joints, const int *jtype, const int *qidx, const float *xyz,
                     const float *rpy, const float *axis,
                     const float *tgt_p, const float *tgt_quat, const float *drone_target,
                     float dt, const float *params, const float *wts, float *S_out)
{
    chain_t ch = {n_joints, jtype, qidx, xyz, rpy, axis};
    float tgt_R[9];
    const float dt2 = (float)((double)dt * (double)dt);
    oracle_quaternion_to_matrix(tgt_quat, tgt_R);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; ++k) {
        quad_state_t s;
        memcpy(&s, qstate, sizeof(s));
        double cum_v[7] = {0}, cum_q[7] = {0};
        float vprev[7], q[7], cq[7], sq[7];
        for (int i = 0; i < 7; ++i) vprev[i] = qd0f[i];
        double S = 0.0, dstage = 0.0;
        float dterm = 0.0f;
        for (int t = 0; t < T; ++t) {
            const float *eps = noise + ((size_t)t * K + k) * 11;
            float u[11];
            for (int i = 0; i < 11; ++i) u[i] = u_nom[t * 11 + i] + eps[i];
            quad_advance(&s, u, dt, params[0], params + 1, params[4], params[5]);
            for (int i = 0; i < 7; ++i) {
                float a = u[4 + i];
                float dq = vprev[i] * dt + 0.5f * a * dt2;
                cum_v[i] += (double)(a * dt);
                vprev[i] = (float)cum_v[i] + qd0f[i];
                cum_q[i] += (double)dq;
                q[i] = (float)cum_q[i] + q0f[i];
                cq[i] = cosf(q[i]); sq[i] = sinf(q[i]);
            }
            float x6[6] = {s.p[0], s.p[1], s.p[2], s.rpy[0], s.rpy[1], s.rpy[2]};
            float Bm[16], Tw[16], pn, on;
            oracle_xyzrpy_to_matrix(x6, Bm);
            fk_chain(&ch, q, cq, sq, Bm, Tw);
            oracle_pose_terms(Tw, tgt_p, tgt_R, &pn, &on);
            float c = (t < T - 1) ? (wts[0] * pn + wts[1] * on) : (wts[2] * pn + wts[3] * on);
            S += (double)c;
            float ex = s.p[0] - drone_target[0], ey = s.p[1] - drone_target[1], ez = s.p[2] - drone_target[2];
            float sqd = (float)((double)(ex * ex) + (double)(ey * ey) + (double)(ez * ez));
            if (t < T - 1) dstage += (double)sqd;
            else dterm = sqd;
        }
        S_out[k] = (float)(S + (double)((float)dstage * wts[4]) + (double)(dterm * wts[5]));
    }
}

/* ------------------------------------------------------------------ weighting */
/* S/mppi_solver/mppi.py:173-193 (twin: drone_mppi.py:111-130):
 *   rho = min S;  w = exp((-1/lambda) * (S - rho)) / sum(exp(...))                      */
void oracle_weights(int K, const float *S, float lambda, float *w_out, float *rho_out, float *eta_out)
{
    float rho = S[0];
    for (int k = 1; k < K; ++k) if (S[k] < rho) rho = S[k];
    const float scale = (float)(-1.0 / (double)lambda);
    double eta = 0.0;
    for (int k = 0; k < K; ++k) { w_out[k] = expf(scale * (S[k] - rho)); eta += (double)w_out[k]; }
    float etaf = (float)eta;
    for (int k = 0; k < K; ++k) w_out[k] = w_out[k] / etaf;
    *rho_out = rho; *eta_out = etaf;
}

/* S/mppi_solver/mppi.py:148 / drone_mppi.py:158:  w_eps[t][i] = sum_k w_k * noise[k][t][i]
 * (library-ordered f32 reduction in the reference; exact-in-double here).              */
void oracle_weighted_noise(int K, int T, int nu, const float *noise /*[T][K][nu]*/, const float *w,
                           float *out /*[T][nu]*/)
{
#pragma omp parallel for schedule(static)
    for (int t = 0; t < T; ++t) {
        double acc[ORACLE_MAX_NU] = {0};
        for (int k = 0; k < K; ++k) {
            const float *eps = noise + ((size_t)t * K + k) * nu;
            for (int i = 0; i < nu; ++i) acc[i] += (double)(w[k] * eps[i]);
        }
        for (int i = 0; i < nu; ++i) out[t * nu + i] = (float)acc[i];
    }
}

/* S/filter/svg_filter.py:13-90: per column, least-squares smoothing taps (row 0 of
 * (A^T A)^-1 A^T for the Vandermonde A on x = -h..h), symmetric edge padding that
 * repeats the edge sample (:58), valid correlation with the taps.                      */
int oracle_savgol_taps(int window, int polyorder, float *taps)
{
    int h = window / 2, n = polyorder + 1;
    if (window % 2 != 1 || polyorder >= window || n > 8) return -1;
    double M[8][16];   /* [A^T A | I] */
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
            double acc = 0.0;
            for (int x = -h; x <= h; ++x) acc += pow((double)x, r) * pow((double)x, c);
            M[r][c] = acc; M[r][n + c] = (r == c) ? 1.0 : 0.0;
        }
    for (int p = 0; p < n; ++p) {
        int best = p;
        for (int r = p + 1; r < n; ++r) if (fabs(M[r][p]) > fabs(M[best][p])) best = r;
        if (fabs(M[best][p]) < 1e-300) return -2;
        if (best != p) for (int c = 0; c < 2 * n; ++c) { double tmp = M[p][c]; M[p][c] = M[best][c]; M[best][c] = tmp; }
        double piv = M[p][p];
        for (int c = 0; c < 2 * n; ++c) M[p][c] /= piv;
        for (int r = 0; r < n; ++r) if (r != p) {
            double f = M[r][p];
            for (int c = 0; c < 2 * n; ++c) M[r][c] -= f * M[p][c];
        }
    }
    for (int x = -h; x <= h; ++x) {
        double acc = 0.0;
        for (int c = 0; c < n; ++c) acc += M[0][n + c] * pow((double)x, c);
        taps[x + h] = (float)acc;
    }
    return 0;
}

int oracle_savgol(int T, int nu, const float *seq /*[T][nu]*/, int window, int polyorder, float *out)
{
    float taps[64];
    int h = window / 2;
    if (window > 63 || T <= h) return -3;     /* (:47-48) raises for short data */
    int rc = oracle_savgol_taps(window, polyorder, taps);
    if (rc) return rc;
    for (int i = 0; i < nu; ++i)
        for (int t = 0; t < T; ++t) {
            double acc = 0.0;
            for (int j = -h; j <= h; ++j) {
                int s = t + j;
                if (s < 0) s = -s - 1;                 /* data[:h].flip(0) */
                if (s >= T) s = 2 * T - 1 - s;         /* data[-h:].flip(0) */
                acc += (double)(taps[j + h] * seq[s * nu + i]);
            }
            out[t * nu + i] = (float)acc;
        }
    return 0;
}

/* ------------------------------------------------------------------ counter-based noise */
/* Philox4x32-10 (Salmon et al., SC'11; Random123 v1.x constants).  New in this build --
 * the reference samples with torch.randn (standard_normal_noise.py:24).  Addressing:
 *   counter = (k_global, t * n_calls + call, step_lo, step_hi), key = (seed_lo, seed_hi)
 * yields the six normals for inputs 6*call .. 6*call+5 of sample k at horizon step t. */
/* `rounds` = 10 is the Random123 / cuRAND default (known-answer vectors in tests/test_oracle_golden.py); 7 is the
 * smallest round count the paper reports as Crush-resistant (table 2) -- same round function, fewer rounds.        */
void oracle_philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    oracle_philox4x32_r(ctr, key, 10, out);
}

/* One call = 128 bits = six 21-bit uniforms (bits [21 i, 21 i + 21) of the little-endian word, placed in the top
 * of a float mantissa: f in [1,2)) = three Box-Muller pairs, same bit recipe as the device code:
 *   u1 = 2 - f_radius in (0,1];  r = sqrt(-2 ln u1);  th = (f_angle - 1.5) * 2pi in [-pi,pi);  (r cos th, r sin th) */
static void uniforms6(const uint32_t r[4], float f[6])
{
    const uint32_t M = 0x1FFFFFu;
    uint32_t u[6];
    u[0] = r[0] & M;
    u[1] = ((r[0] >> 21) | (r[1] << 11)) & M;
    u[2] = (r[1] >> 10) & M;
    u[3] = ((r[1] >> 31) | (r[2] << 1)) & M;
    u[4] = ((r[2] >> 20) | (r[3] << 12)) & M;
    u[5] = (r[3] >> 9) & M;
    for (int i = 0; i < 6; ++i) {
        union { uint32_t u; float f; } b;
        b.u = 0x3f800000u | (u[i] << 2);
        f[i] = b.f;
    }
}
static void box_muller(float f_radius, float f_angle, float *n0, float *n1)
{
    float u1 = 2.0f - f_radius;
    float r = sqrtf(-2.0f * logf(u1));
    float th = fmaf(f_angle, 6.28318530717958647692f, -9.42477796076937971538f);   /* (f - 1.5) * 2 pi as one FMA, like the device */
    *n0 = r * cosf(th);
    *n1 = r * sinf(th);
}

void oracle_philox_noise_r(int K, int T, int nu, long long k_offset, uint64_t seed, uint64_t step, int rounds,
                           const float *sigma /*[nu]*/, float *noise /*[T][K][nu]*/)
{
    int nch = ((nu + 1) / 2 + 2) / 3;      /* Philox calls per (sample, step): three pairs per call */
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma omp parallel for schedule(static)
    for (int t = 0; t < T; ++t)
        for (int k = 0; k < K; ++k)
            for (int c = 0; c < nch; ++c) {
                uint32_t ctr[4] = {(uint32_t)(k_offset + k), (uint32_t)(t * nch + c),
                                   (uint32_t)step, (uint32_t)(step >> 32)};
                uint32_t r[4];
                float f[6], n[6];
                oracle_philox4x32_r(ctr, key, rounds, r);
                uniforms6(r, f);
                for (int p = 0; p < 3; ++p) box_muller(f[2 * p], f[2 * p + 1], &n[2 * p], &n[2 * p + 1]);
                for (int j = 0; j < 6; ++j) {
                    int i = 6 * c + j;
                    if (i < nu) noise[((size_t)t * K + k) * nu + i] = sigma[i] * n[j];
                }
            }
}

void oracle_philox_noise(int K, int T, int nu, long long k_offset, uint64_t seed, uint64_t step,
                         const float *sigma /*[nu]*/, float *noise /*[T][K][nu]*/)
{
    oracle_philox_noise_r(K, T, nu, k_offset, seed, step, 10, sigma, noise);
}
