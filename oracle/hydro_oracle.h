/* hydro_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, fp64, serial) of the reference's per-time-step hot
 * path, used as the checker for the CUDA path and as the "port" CPU baseline.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it; the product (hydro_b200/, libhydro_gpu.so) never
 * does.  Parity pinning: tests/test_oracle_vs_golden.py checks it against
 * fixtures produced by the real reference (oracle/_ref/ref_dump, built from
 * /root/reference by oracle/Makefile) and against the reference's own cavity
 * sample (examples/cavity/sample).
 *
 * The interface mirrors include/hydro_gpu.h one to one (ho_* <-> hg_*), so a
 * parity test issues the same calls against both.
 */
#ifndef HYDRO_ORACLE_H_
#define HYDRO_ORACLE_H_

#include "../include/hydro_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ho_state* ho_handle;

void ho_config_defaults(hg_config* cfg);
int ho_create(const hg_config* cfg, ho_handle* out);
int ho_destroy(ho_handle h);
const char* ho_last_error(ho_handle h);
size_t ho_num_cells(ho_handle h);
size_t ho_num_faces(ho_handle h);
int ho_set_field(ho_handle h, int field, const double* src, size_t n);
int ho_get_field(ho_handle h, int field, double* dst, size_t n);
int ho_step(ho_handle h, hg_step_stats* stats);
int ho_fluid_start_step(ho_handle h);
int ho_fluid_make_iteration(ho_handle h);
int ho_fluid_convergence_indicator(ho_handle h, double* out);
int ho_fluid_is_converged(ho_handle h, int* out);
int ho_fluid_finish_step(ho_handle h);
int ho_fluid_auto_time_step(ho_handle h, double* out);
int ho_set_time_step(ho_handle h, double dt_fluid, double dt_advection);
int ho_advection_step(ho_handle h);
int ho_heat_step(ho_handle h);
int ho_update_properties(ho_handle h);
int ho_calc_stat(ho_handle h, hg_step_stats* stats);
int ho_interp_grad(ho_handle h, const double* u, int cond, int comp,
                   double* grad_x, double* grad_y, double* grad_z);
int ho_linear_solve(ho_handle h, int solver, const double* const coeffs[7],
                    const double* rhs, double* x, double tolerance,
                    int num_iters_limit, double relaxation_factor,
                    int* out_iters, double* out_diff);
int ho_smooth_field(ho_handle h, const double* u, int repeat, double* out);
/* residual (convergence indicator) of every SIMPLE iteration of the last ho_step */
int ho_last_residuals(ho_handle h, double* out, int cap, int* n);

#ifdef __cplusplus
}
#endif
#endif
