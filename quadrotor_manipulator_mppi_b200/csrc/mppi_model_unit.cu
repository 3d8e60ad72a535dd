// mppi_model_unit.cu -- one translation unit per (model, part): instantiates the kernels of that model.
//   nvcc ... -DMPPI_UNIT_MODEL=<0..3> -DMPPI_UNIT_PART=<0|1> -c mppi_model_unit.cu
//   part 0: rollout_cost_kernel, the weighting kernels, finalize_kernel        (the two-kernel step and the phase API)
//   part 1: step_fused_kernel, step_tp_kernel                                  (the single-launch steps)
#include "mppi_launch.cuh"

#ifndef MPPI_UNIT_MODEL
#error "build with -DMPPI_UNIT_MODEL=<model> -DMPPI_UNIT_PART=<part>"
#endif
#define MPPI_CAT_(a, b) a##b
#define MPPI_CAT(a, b) MPPI_CAT_(a, b)
#define MPPI_UNIT_FN(name) MPPI_CAT(name, MPPI_UNIT_MODEL)

using namespace mppi;

#if MPPI_UNIT_PART == 0
mppi_status_t MPPI_UNIT_FN(unit_rollout_)(mppi_ctx *h, const float *u, const float *n, float *c, cudaStream_t st)
{
    return launch_rollout<MPPI_UNIT_MODEL>(h, u, n, c, st);
}
mppi_status_t MPPI_UNIT_FN(unit_weight_)(mppi_ctx *h, const float *n, bool fuse, const float *u, float *un, float *o, cudaStream_t st,
                                         const P2PParams &X)
{
    return launch_weight<MPPI_UNIT_MODEL>(h, n, fuse, u, un, o, st, X);
}
mppi_status_t MPPI_UNIT_FN(unit_finalize_)(mppi_ctx *h, const float *u, float *un, float *o, cudaStream_t st)
{
    return launch_finalize<MPPI_UNIT_MODEL>(h, u, un, o, st);
}
#else
mppi_status_t MPPI_UNIT_FN(unit_fused_)(mppi_ctx *h, const float *u, float *un, float *o, cudaStream_t st, const P2PParams &X, bool *launched)
{
    return launch_fused<MPPI_UNIT_MODEL>(h, u, un, o, st, X, launched);
}
mppi_status_t MPPI_UNIT_FN(unit_tp_)(mppi_ctx *h, const float *u, const float *n, float *un, float *o, cudaStream_t st, const P2PParams &X,
                                     bool *launched)
{
    return launch_tp<MPPI_UNIT_MODEL>(h, u, n, un, o, st, X, launched);
}
#endif
