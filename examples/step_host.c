/* Minimal C caller of the C ABI (no torch, no Python): one arm controller stepping from host buffers.
 *
 *   gcc -std=c99 -Iinclude examples/step_host.c -Lquadrotor_manipulator_mppi_b200 -lmppi_b200 \
 *       -Wl,-rpath,$PWD/quadrotor_manipulator_mppi_b200 -o step_host && ./step_host
 *
 * Mirrors what src/mav_mppi/scripts/kinova.py does with its MPPI object: update_joint (:116) -> the state vector,
 * compute_control_input (:182) -> mppi_step_host, torque law (:184) -> out[MPPI_OUT_TORQUE..].                     */
#include <stdio.h>
#include <string.h>

#include "mppi_b200.h"

int main(void)
{
    mppi_config_t cfg;
    mppi_handle_t h = NULL;
    if (mppi_default_config(MPPI_MODEL_ARM7, &cfg) != MPPI_OK) return 1;
    cfg.n_samples = 1024;
    cfg.n_horizon = 30;
    cfg.cost_flags |= MPPI_OPT_TORQUE_LAW;
    if (mppi_create(&cfg, &h) != MPPI_OK) {
        fprintf(stderr, "mppi_create: %s\n", mppi_last_error(NULL));          /* e.g. no sm_100 device: there is no CPU fallback */
        return 2;
    }
    /* q[7], qdot[7], base xyz + quat xyzw, base twist (linear, angular) */
    float state[27] = {1.57f, 1.7f, 0.f, 4.4f, 0.f, 4.71f, 0.f, 0, 0, 0, 0, 0, 0, 0, 0.f, 0.f, 2.1f, 0.f, 0.f, 0.f, 1.f, 0, 0, 0, 0, 0, 0};
    static float u[30 * 7];                                                    /* warm start: zeros, carried between steps */
    float out[MPPI_OUT_FLOATS];
    memset(u, 0, sizeof(u));
    for (unsigned long long step = 0; step < 5; ++step) {
        if (mppi_step_host(h, state, 27, u, NULL, step, NULL, out) != MPPI_OK) {
            fprintf(stderr, "mppi_step_host: %s\n", mppi_last_error(h));
            mppi_destroy(h);
            return 3;
        }
        printf("step %llu  qdes[0]=%.6f  reach=%.4f  rho=%.3f  ess=%.2f  tau[1]=%.3f\n", step, out[0], out[MPPI_OUT_REACH],
               out[MPPI_OUT_RHO], out[MPPI_OUT_ESS], out[MPPI_OUT_TORQUE + 1]);
    }
    mppi_destroy(h);
    return 0;
}
