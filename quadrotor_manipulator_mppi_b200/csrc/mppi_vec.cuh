// mppi_vec.cuh -- one- and two-lane FP32 value types for the rollout kernels.
//
// Blackwell (sm_100) adds packed FP32x2 arithmetic (FFMA2 / FADD2 / FMUL2 on 64-bit register
// pairs, with 32-bit immediate or broadcast-scalar operands).  It has the same FLOP rate as scalar
// FFMA but needs half the issue slots (tools/probe_ffma2.cu, tools/probe_pipes.cu).  The rollout
// kernel packs the work of ONE sample into these instructions (joint pairs, rotation-column pairs,
// paired atan2 / Box-Muller), so the scalar math helpers are written once over a value type V and
// instantiated with V = float and V = f2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mppi {

struct f2 {
    float2 v;
    __device__ __forceinline__ f2() {}
    __device__ __forceinline__ f2(float a) : v(make_float2(a, a)) {}          // broadcast (free: .F32 / immediate operand)
    __device__ __forceinline__ f2(float a, float b) : v(make_float2(a, b)) {}
};
struct b2 { bool x, y; };
struct i2 { int x, y; };

template <class V> struct Lanes;
template <> struct Lanes<float> { static constexpr int n = 1; using mask = bool; using ivec = int; };
template <> struct Lanes<f2>    { static constexpr int n = 2; using mask = b2;   using ivec = i2; };

// ---- lane access
__device__ __forceinline__ float lane(float v, int) { return v; }
__device__ __forceinline__ float lane(const f2 &v, int i) { return i ? v.v.y : v.v.x; }

// ---- arithmetic (explicit names: fusion and rounding are part of the contract)
__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float vmul(float a, float b) { return a * b; }
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float vsub(float a, float b) { return a - b; }
__device__ __forceinline__ float vneg(float a) { return -a; }

__device__ __forceinline__ f2 vfma(f2 a, f2 b, f2 c) { f2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
__device__ __forceinline__ f2 vmul(f2 a, f2 b) { f2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 vadd(f2 a, f2 b) { f2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 vneg(f2 a) { return f2(-a.v.x, -a.v.y); }
__device__ __forceinline__ f2 vsub(f2 a, f2 b) { return vadd(a, vneg(b)); }

__device__ __forceinline__ f2 operator+(f2 a, f2 b) { return vadd(a, b); }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { return vsub(a, b); }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { return vmul(a, b); }
__device__ __forceinline__ f2 operator-(f2 a) { return vneg(a); }
__device__ __forceinline__ f2 &operator+=(f2 &a, f2 b) { a = vadd(a, b); return a; }

// ---- per-lane scalar maps (MUFU and friends are not packed)
template <class F> __device__ __forceinline__ float vmap(float x, F f) { return f(x); }
template <class F> __device__ __forceinline__ f2 vmap(f2 x, F f) { return f2(f(x.v.x), f(x.v.y)); }

__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <class V> __device__ __forceinline__ V vrcp(V x) { return vmap(x, [](float a) { return rcp_approx(a); }); }
template <class V> __device__ __forceinline__ V vsqrt(V x) { return vmap(x, [](float a) { return sqrt_approx(a); }); }
template <class V> __device__ __forceinline__ V vabs(V x) { return vmap(x, [](float a) { return fabsf(a); }); }

__device__ __forceinline__ float vmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float vmin(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ f2 vmax(f2 a, f2 b) { return f2(fmaxf(a.v.x, b.v.x), fmaxf(a.v.y, b.v.y)); }
__device__ __forceinline__ f2 vmin(f2 a, f2 b) { return f2(fminf(a.v.x, b.v.x), fminf(a.v.y, b.v.y)); }
__device__ __forceinline__ float vcopysign(float m, float s) { return copysignf(m, s); }
__device__ __forceinline__ f2 vcopysign(f2 m, f2 s) { return f2(copysignf(m.v.x, s.v.x), copysignf(m.v.y, s.v.y)); }

// ---- comparisons / selects (per lane)
__device__ __forceinline__ bool vgt(float a, float b) { return a > b; }
__device__ __forceinline__ bool vlt(float a, float b) { return a < b; }
__device__ __forceinline__ b2 vgt(f2 a, f2 b) { return b2{a.v.x > b.v.x, a.v.y > b.v.y}; }
__device__ __forceinline__ b2 vlt(f2 a, f2 b) { return b2{a.v.x < b.v.x, a.v.y < b.v.y}; }
__device__ __forceinline__ float vsel(bool m, float a, float b) { return m ? a : b; }
__device__ __forceinline__ f2 vsel(b2 m, f2 a, f2 b) { return f2(m.x ? a.v.x : b.v.x, m.y ? a.v.y : b.v.y); }

// ---- bit tricks
__device__ __forceinline__ int vbits(float a) { return __float_as_int(a); }
__device__ __forceinline__ i2 vbits(f2 a) { return i2{__float_as_int(a.v.x), __float_as_int(a.v.y)}; }
__device__ __forceinline__ int vshl31(int n) { return n << 31; }
__device__ __forceinline__ i2 vshl31(i2 n) { return i2{n.x << 31, n.y << 31}; }
__device__ __forceinline__ float vxor(float a, int m) { return __int_as_float(__float_as_int(a) ^ m); }
__device__ __forceinline__ f2 vxor(f2 a, i2 m) { return f2(__int_as_float(__float_as_int(a.v.x) ^ m.x), __int_as_float(__float_as_int(a.v.y) ^ m.y)); }

}  // namespace mppi
