// Operand-form and cross-pipe questions behind the rollout kernel's FMA-pipe model (cycles per warp-slot per scheduler).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_pipes2 tools/probe_pipes2.cu && tools/probe_pipes2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sinf_a(float x) { float y; asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(256) probe(float *out, int iters, float a, float b)
{
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(1.0f + 1e-3f * (threadIdx.x + i), 1.0f - 1e-3f * (threadIdx.x - i));
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = (i + 1) & 7, k = (i + 3) & 7;
                if (MODE == 0) x[i] = __ffma2_rn(x[i], a2, b2);                                             // FFMA2, two operands shared
                if (MODE == 1) x[i] = __ffma2_rn(x[i], x[j], x[k]);                                         // FFMA2, three distinct register pairs
                if (MODE == 2) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }                // 2 FFMA, constant-bank operands
                if (MODE == 3) { x[i].x = fmaf(x[i].x, x[j].y, x[k].x); x[i].y = fmaf(x[i].y, x[j].x, x[k].y); }   // 2 FFMA, 3 registers
                if (MODE == 4) { x[i].x = x[i].x * x[j].y; x[i].y = x[i].y * x[j].x; }                      // 2 FMUL, 2 registers
                if (MODE == 5) { x[i].x = x[i].x + x[j].y; x[i].y = x[i].y + x[j].x; }                      // 2 FADD, 2 registers
                if (MODE == 6) { x[i].x = fmaf(x[i].x, x[j].y, b); x[i].y = fmaf(x[i].y, x[j].x, b); }      // 2 FFMA, 2 registers + constant
                if (MODE == 7) { x[i].x = ex2f(x[i].x); }                                                   // 1 MUFU
                if (MODE == 8) { x[i].x = ex2f(x[i].x); x[j] = __ffma2_rn(x[j], a2, b2); }                  // MUFU + FFMA2
                if (MODE == 9) { x[i].x = ex2f(x[i].x); x[j] = __ffma2_rn(x[j], a2, b2); x[k] = __ffma2_rn(x[k], a2, b2);
                                 x[(i + 5) & 7] = __ffma2_rn(x[(i + 5) & 7], a2, b2); x[(i + 6) & 7] = __ffma2_rn(x[(i + 6) & 7], a2, b2); }   // MUFU + 4 FFMA2
                if (MODE == 10) { x[i].x = sinf_a(x[i].x); }                                                // FMUL (range) + MUFU.SIN
                if (MODE == 11) { x[i].x = fmaf(x[i].x, x[j].y, x[k].x); x[i].y = x[i].y * x[j].x; }        // FFMA 3-reg + FMUL
                if (MODE == 12) { x[i] = __fmul2_rn(x[i], x[j]); }                                          // FMUL2 two distinct pairs
                if (MODE == 13) { x[i].x = fmaf(x[i].x, x[i].y, x[k].x); x[i].y = fmaf(x[i].y, x[i].x, x[k].y); }   // 2 FFMA, 2 distinct registers + self
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    if (s == 1234.5f) out[0] = s;
}

template <int MODE>
float run(int sms, int iters)
{
    float *d; cudaMalloc(&d, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        probe<MODE><<<sms * 8, 256>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaFree(d);
    return best;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 1024;
    const char *names[] = {"FFMA2 shared operands", "FFMA2 3 distinct pairs", "2x FFMA const operands", "2x FFMA 3-reg", "2x FMUL 2-reg", "2x FADD 2-reg",
                           "2x FFMA 2-reg + const", "MUFU.EX2", "MUFU + FFMA2", "MUFU + 4 FFMA2", "sin.approx (FMUL+MUFU)", "FFMA 3-reg + FMUL", "FMUL2 distinct",
                           "2x FFMA 2 distinct regs"};
    float t[14] = {run<0>(sms, iters), run<1>(sms, iters), run<2>(sms, iters), run<3>(sms, iters), run<4>(sms, iters), run<5>(sms, iters), run<6>(sms, iters),
                   run<7>(sms, iters), run<8>(sms, iters), run<9>(sms, iters), run<10>(sms, iters), run<11>(sms, iters), run<12>(sms, iters), run<13>(sms, iters)};
    for (int i = 0; i < 14; ++i) {
        const double cyc = t[i] * 1e-3 * 1.965e9;                        // assumes 1965 MHz
        const double per_slot = cyc / (double)iters / 64.0 / 16.0;      // cycles per warp-slot per scheduler (16 warps each)
        printf("%-26s %.3f ms  %.2f cycles per slot per scheduler\n", names[i], t[i], per_slot);
    }
    return 0;
}
