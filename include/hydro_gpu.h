/* hydro_gpu.h -- C ABI of the B200-native hot path of divfree/hydro.
 *
 * One opaque handle (`hg_handle`) = the device-resident state of ONE reference
 * "experiment" (reference: one `hydro<Mesh>` module instance,
 * source/hydro2dmpi/hydro2d.hpp:48-192).  Everything `hydro<Mesh>::step()`
 * (hydro2d.hpp:1531-1621) does per time step runs on the GPU behind this
 * interface; host buffers are only touched by hg_set_field / hg_get_field.
 *
 * Conventions
 *  - plain C, no C++/torch types; every entry returns 0 on success, a negative
 *    hg_status otherwise, and never throws or aborts.  hg_last_error(h) gives
 *    the message; the C++ shim (hydro_b200/host/hydro_gpu.hpp) turns non-zero
 *    into `throw std::string`, the reference's own error convention
 *    (fluid.hpp:795-797, control/module.cpp:103-107).
 *  - arrays are fp64 in the reference's raw-index order: cells
 *    raw = i + nx*(j + ny*k) (mesh.hpp:552-570), faces direction-major x,y,z,
 *    each block flattened x-fastest over the cell block enlarged by one in its
 *    own direction (mesh.hpp:640-720).  Vector cell fields are passed as
 *    separate components (SoA), not as the reference's AoS Vect.
 *  - handle-scoped state only, no globals: the reference runs every experiment
 *    on its own std::thread (control/console.cpp:162-178); every entry point
 *    calls cudaSetDevice for the calling thread.
 *  - there is NO CPU fallback: hg_create fails with HG_ERR_NO_DEVICE when no
 *    CUDA device is present.
 */
#ifndef HYDRO_GPU_H_
#define HYDRO_GPU_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HG_MAX_PHASES 3

/* linear solver ids = the names accepted by hydro<Mesh>::GetLinearSolverFactory
 * (hydro2d.hpp:194-248) */
enum hg_linear_solver {
  HG_LS_LU = 0,           /* "lu"           linear.hpp:527-567 */
  HG_LS_LU_RELAXED = 1,   /* "lu_relaxed"   linear.hpp:578-651 */
  HG_LS_GAUSS_SEIDEL = 2, /* "gauss_seidel" linear.hpp:671-716 */
  HG_LS_JACOBI = 3        /* "jacobi"       linear.hpp:736-783 */
};

/* fluid boundary condition kinds = solver::Parse (fluid.hpp:392-417) */
enum hg_bc_kind { HG_BC_WALL = 0, HG_BC_INLET = 1, HG_BC_OUTLET = 2 };

/* domain sides in the order of hydro2d.hpp:283-311 */
enum hg_side {
  HG_SIDE_LEFT = 0, HG_SIDE_RIGHT = 1,   /* x-, x+ : condition_left/right  */
  HG_SIDE_BOTTOM = 2, HG_SIDE_TOP = 3,   /* y-, y+ : condition_bottom/top  */
  HG_SIDE_CLOSE = 4, HG_SIDE_FAR = 5     /* z-, z+ : condition_close/far   */
};

enum hg_status {
  HG_OK = 0,
  HG_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  HG_ERR_NO_DEVICE = -2,   /* no CUDA device: there is no CPU fallback  */
  HG_ERR_CUDA = -3,        /* CUDA runtime error                        */
  HG_ERR_NAN = -4,         /* reference: throw std::string("NaN ...")   */
  HG_ERR_PEER = -5         /* slab decomposition: a wait for a neighbouring rank timed out */
};

/* Field ids for hg_set_field / hg_get_field.  Cell fields have n = nx*ny*nz
 * entries, face fields n = number of faces (all three/two blocks). */
enum hg_field {
  HG_F_VELOCITY_X = 0, HG_F_VELOCITY_Y = 1, HG_F_VELOCITY_Z = 2, /* FluidSolver::GetVelocity()   fluid.hpp:244 */
  HG_F_PRESSURE = 3,                                             /* FluidSolver::GetPressure()   fluid.hpp:246 */
  HG_F_VOLUME_FLUX = 4,                                          /* FluidSolver::GetVolumeFlux() fluid.hpp:248 (face) */
  HG_F_PARTIAL_DENSITY_0 = 5, HG_F_PARTIAL_DENSITY_1 = 6, HG_F_PARTIAL_DENSITY_2 = 7, /* AdvectionSolverMulti::GetField(i) advection.hpp:83 */
  HG_F_TEMPERATURE = 8,                                          /* HeatSolver::GetTemperature() heat.hpp:88 */
  HG_F_DENSITY = 9,        /* fc_density_smooth   hydro2d.hpp:1410 */
  HG_F_VISCOSITY = 10,     /* fc_viscosity_smooth hydro2d.hpp:1413 */
  HG_F_FORCE_X = 11, HG_F_FORCE_Y = 12, HG_F_FORCE_Z = 13,       /* fc_force hydro2d.hpp:1311 */
  HG_F_VOLUME_FRACTION_0 = 14, HG_F_VOLUME_FRACTION_1 = 15, HG_F_VOLUME_FRACTION_2 = 16,
  HG_F_STFORCE_X = 17, HG_F_STFORCE_Y = 18, HG_F_STFORCE_Z = 19, /* fc_stforce hydro2d.hpp:1316 */
  HG_F_VELOCITY_PREV_X = 20, HG_F_VELOCITY_PREV_Y = 21, HG_F_VELOCITY_PREV_Z = 22, /* Layers::time_prev */
  HG_F_PRESSURE_PREV = 23,
  HG_F_VOLUME_FLUX_PREV = 24,
  HG_F_EXCLUDED = 25,      /* 1.0 where MeshStructured::IsExcluded(cell)  (read only) */
  HG_F_CONDUCTIVITY = 26,
  HG_F_COUNT = 27
};

/* Parameters: same names and meaning as the reference's P_int/P_double/P_bool/
 * P_string/P_vect keys (examples/general.hydroconf). */
typedef struct hg_config {
  int dim;                 /* 2 = MODULE hydro2d, 3 = MODULE hydro3d */
  int Nx, Ny, Nz;          /* Nz ignored (1) for dim == 2 */
  double A[3], B[3];       /* domain corners */
  double box_A[3], box_B[3];       /* rigid box (excluded cells), hydro2d.hpp:306-308 */
  int condition_kind[6];           /* hg_bc_kind per hg_side */
  double condition_velocity[6][3]; /* "wall vx vy vz" / "inlet vx vy vz" */
  int pressure_fixed_enable;       /* P_vect("pressure_fixed_point") exists */
  double pressure_fixed_point[3];
  double pressure_fixed_value;
  /* initial state (hydro2d.hpp:310-368, 479-550) */
  double initial_velocity[3];
  int initial_pois;
  int initial_sin_enable;          /* P_vect.exist("initial_sin_n") */
  double initial_sin_n[3], initial_sin_lambda, initial_sin_phase;
  double A1[3], B1[3], A2[3], B2[3];
  double IC[3], IR, IC2[3], IR2;
  double initial_volume_fraction[HG_MAX_PHASES];
  int initial_volume_fraction_smooth_times;
  /* time */
  double dt;
  int dt_auto;
  double cfl, cfl_advection;
  /* physical parameters */
  int num_phases;
  double density[HG_MAX_PHASES], viscosity[HG_MAX_PHASES], conductivity[HG_MAX_PHASES];
  double gravity[3], force[3], sigma;
  /* SIMPLE (hydro2d.hpp:449-463) */
  int fluid_enable, advection_enable;
  double convergence_tolerance;
  int num_iterations_limit;
  double velocity_relaxation_factor, pressure_relaxation_factor, rhie_chow_factor;
  int time_second_order, simpler, force_geometric_average;
  double guess_extrapolation;
  double meshvel[3];
  int meshvel_output;              /* CalcStat adds meshpos += meshvel*dt (hydro2d.hpp:1526-1528) */
  int linear_solver_velocity, linear_solver_pressure, linear_solver_heat; /* hg_linear_solver */
  double lu_relaxed_tolerance;
  int lu_relaxed_num_iters_limit;
  double lu_relaxed_relaxation_factor;
  int density_smooth_times, viscosity_smooth_times, force_smooth_times;
  /* advection (hydro2d.hpp:591-600) */
  double advection_dt_factor;
  int tvd_split;
  double sharp;
  /* heat (hydro2d.hpp:651-687) */
  int heat_enable;
  double temperature_initial;
  double heat_box_lb[3], heat_box_rt[3], heat_box_temperature, heat_relaxation_factor;
  int time_second_order_heat;
  /* multi-GPU z-slab decomposition (no reference counterpart: its MPI is a stub,
   * source/main.cpp:26-28).  world_size 1 = single GPU.  For world_size > 1 the ranks
   * are linked after hg_create with hg_ipc_export/hg_ipc_import (one process per GPU)
   * or hg_link_local (one process); the data plane is NVLink peer memory. */
  int world_size, rank, device;
  /* execution options */
  int pressure_sweeps_per_check;   /* 0 = default */
  int solver_ctas;                 /* 0 = one CTA per SM; >0 limits the persistent solver grids (ranks sharing a device) */
  /* phase slip (Stokes settling relative to the mixture, hydro<Mesh>::CalcPhaseVelocitySlip, hydro2d.hpp:1030-1122):
   * enable_settling_<i>, bubble_radius_<i>; velocity_is_carrier 1 (a volume source) is not on the GPU path */
  int enable_settling[HG_MAX_PHASES];
  int velocity_is_carrier;
  double bubble_radius[HG_MAX_PHASES];
  /* automatic mesh velocity (hydro<Mesh>::CalcStat, hydro2d.hpp:1510-1524; `set string meshvel_auto vx|vcx`): after the
   * statistics of a step the mesh velocity is relaxed towards (v, 0, 0) with weight meshvel_weight, v = stat_vx_1 (1: mean
   * x velocity of phase 1) or stat_vcx_1 (2: velocity of its centre); 0 = off.  Needs num_phases >= 2, one GPU. */
  int meshvel_auto;
  int reserved0;
  double meshvel_weight;
  int reserved[2];
} hg_config;

/* statistics of one hg_step, reference: P_int["s"], CalcStat (hydro2d.hpp:1432-1529) */
typedef struct hg_step_stats {
  int simple_iterations;         /* FluidSolver::GetIterationCount()            */
  double convergence_indicator;  /* GetConvergenceIndicator() after the last it */
  int pressure_sweeps_total;     /* sum of `iter+1` over the step's pressure solves (linear.hpp:712) */
  int advection_substeps;
  double dt, time;
  double pressure_last_diff;     /* `diff` printed by the last pressure solve   */
  double volume[HG_MAX_PHASES], mass[HG_MAX_PHASES];
  double pd_min[HG_MAX_PHASES], pd_max[HG_MAX_PHASES];
  double center[HG_MAX_PHASES][3], velocity[HG_MAX_PHASES][3];
} hg_step_stats;

typedef struct hg_state* hg_handle;

/* Fills `cfg` with the defaults of examples/general.hydroconf. */
void hg_config_defaults(hg_config* cfg);

/* lifetime: replaces hydro<Mesh>::hydro(TExperiment*) (hydro2d.hpp:928-972): mesh,
 * boundary conditions, solvers, initial fields, first UpdateFluidProperties + CalcStat. */
int hg_create(const hg_config* cfg, hg_handle* out);
int hg_destroy(hg_handle h);
const char* hg_last_error(hg_handle h); /* h may be NULL: error of the last failed hg_create */

/* Slab decomposition (world_size > 1), one handle per GPU.  hg_create allocates; the ranks are then linked
 * -- collectively, every rank calls one of the two -- and the initial fields are computed:
 *   separate processes: hg_ipc_export fills one record (hg_ipc_record_size() bytes) with CUDA IPC handles of the
 *     buffers the neighbouring ranks touch; the caller all-gathers the records (any transport) and passes the
 *     world_size records, in rank order, to hg_ipc_import;
 *   one process: hg_link_local takes the world_size handles in rank order.
 * The reference has nothing to replace here (its MPI is a stub, source/main.cpp:26-28). */
size_t hg_ipc_record_size(void);
int hg_ipc_export(hg_handle h, void* record, size_t cap);
int hg_ipc_import(hg_handle h, const void* all_records);
int hg_link_local(hg_handle h, const hg_handle* all_handles);

size_t hg_num_cells(hg_handle h);  /* local cells of this rank's slab */
size_t hg_num_faces(hg_handle h);

/* field transfer (host buffers are copied; nothing is retained).  Setting a
 * velocity/pressure/flux/partial-density/temperature field sets the time_curr and
 * time_prev layers, as the solvers' constructors do (conv_diff.hpp:113-114). */
int hg_set_field(hg_handle h, int field, const double* src, size_t n);
int hg_get_field(hg_handle h, int field, double* dst, size_t n);
/* Asynchronous variants for PINNED host buffers (same ids, sizes and semantics; the module-owned property fields of
 * hydro2d.hpp:449-463 go in before a step, results come out after it).  The upload runs on a copy stream into a staging
 * buffer and is applied when the compute stream reaches the call's position; the download copies a snapshot taken at the
 * call's position, concurrently with later steps.  The host buffer may be reused / read after hg_device_synchronize. */
int hg_set_field_async(hg_handle h, int field, const double* src, size_t n);
int hg_get_field_async(hg_handle h, int field, double* dst, size_t n);

/* one whole time step = hydro<Mesh>::step() (hydro2d.hpp:1531-1621) */
int hg_step(hg_handle h, hg_step_stats* stats /* may be NULL */);
/* the same step in two halves (hg_step = hg_step_begin + hg_step_end): hg_step_begin queues the step on the device and
 * returns, hg_step_end waits for its status block (NaN flags, solver status, statistics) and fills `stats`.  Between the two a
 * caller can queue the transfers of the NEXT step (hg_set_field_async: the upload runs on the copy stream while the step
 * computes and is applied in stream order after it) -- the module-owned property fields of hydro<Mesh> (hydro2d.hpp:449-463)
 * no longer cost their PCIe time.  hg_step_begin still waits where the reference's control flow needs a device result (see
 * hg_run).  Errors of the step are reported by hg_step_end. */
int hg_step_begin(hg_handle h);
int hg_step_end(hg_handle h, hg_step_stats* stats /* may be NULL */);
/* n steps back to back.  The host only waits for the device where the reference's control flow depends on
 * device results: once per step (NaN flags, solver status words, statistics, all in one pinned status block),
 * plus once per SIMPLE iteration when convergence_tolerance > 0 and once per chunk of pressure sweeps when
 * lu_relaxed_tolerance > 0 (the stop tests of solver.hpp:733-736 and linear.hpp:707-710). */
int hg_run(hg_handle h, int nsteps, hg_step_stats* last_stats /* may be NULL */);

/* fine-grained protocol = solver::UnsteadyIterativeSolver (solver.hpp:710-751)
 * on FluidSimple (fluid.hpp:793-1169) */
int hg_fluid_start_step(hg_handle h);
int hg_fluid_make_iteration(hg_handle h);
int hg_fluid_convergence_indicator(hg_handle h, double* out);
int hg_fluid_is_converged(hg_handle h, int* out);
/* convergence indicator of every SIMPLE iteration of the current/last step (the `Rs=` values the
 * reference logs, hydro2d.hpp:1585-1586); at most `cap`, count in *n */
int hg_last_residuals(hg_handle h, double* out, int cap, int* n);
int hg_fluid_finish_step(hg_handle h);
int hg_fluid_auto_time_step(hg_handle h, double* out);      /* FluidSimple::GetAutoTimeStep fluid.hpp:1191 */
int hg_set_time_step(hg_handle h, double dt_fluid, double dt_advection);
int hg_advection_step(hg_handle h);                          /* Start/CalcStep/Finish, advection.hpp:417-545 */
int hg_heat_step(hg_handle h);                               /* heat.hpp:69-84 */
int hg_update_properties(hg_handle h);                       /* hydro2d.hpp:1404-1430 */
int hg_calc_stat(hg_handle h, hg_step_stats* stats);         /* hydro2d.hpp:1432-1529, incl. meshpos += meshvel*dt (1526-1528) */
int hg_get_stats(hg_handle h, hg_step_stats* stats);         /* the statistics of the last CalcStat, nothing is recomputed or advanced */

/* kernel-level entries (parity tests and micro-benchmarks, cf. test/benchmark/main.cpp) */

/* Gradient(Interpolate(u, cond)) (solver.hpp:392-470, 658-677).  cond: 0 = zero
 * derivative, 1 = extrapolation, 2 = Dirichlet with the fluid wall velocity
 * component `comp`.  u: n cells; grad_*: n cells each (grad_z may be NULL for 2-D) */
int hg_interp_grad(hg_handle h, const double* u, int cond, int comp,
                   double* grad_x, double* grad_y, double* grad_z);

/* LinearSolver::Solve (linear.hpp:302-310) for a 7/5-point system in SoA form.
 * coeffs: 7 arrays of n doubles in the order z-,y-,x-,diag,x+,y+,z+ (2-D: the z
 * entries are ignored and may be NULL); rhs = the Expression constants; x = result.
 * out_iters/out_diff: the values the reference prints (linear.hpp:712). */
int hg_linear_solve(hg_handle h, int solver, const double* const coeffs[7],
                    const double* rhs, double* x, double tolerance,
                    int num_iters_limit, double relaxation_factor,
                    int* out_iters, double* out_diff);

/* GetSmoothField (solver.hpp:636-656) */
int hg_smooth_field(hg_handle h, const double* u, int repeat, double* out);

/* phase timers, same keys as the reference's MultiTimer (fluid.hpp:603..1061,
 * hydro2d.hpp:1533-1620).  names: array of `cap` char[64]; returns count in *n */
int hg_timers(hg_handle h, char (*names)[64], double* seconds, int cap, int* n);
int hg_timers_enable(hg_handle h, int enable);

/* counts kernels launched by this handle since creation (bench.py "gpu_launches") */
long long hg_launch_count(hg_handle h);
/* name of the kernel the handle runs for (which: 0 = pressure sweeps, 1 = lu); bench.py labels its roofline with it */
const char* hg_solver_kernel_name(hg_handle h, int which);

/* device/bench helpers: CUDA events recorded on the handle's own stream (8 slots), and per-launch
 * timing of the two persistent solver kernels (which: 0 = pressure sweeps, 1 = lu); reading clears */
int hg_event_record(hg_handle h, int slot);
int hg_event_elapsed_ms(hg_handle h, int slot_a, int slot_b, double* ms);
int hg_profile_enable(hg_handle h, int enable);
int hg_profile_read(hg_handle h, int which, int* count, double* total_ms);
int hg_device_synchronize(hg_handle h);
/* cycle counters of the sweep kernel's warp roles (only in builds with -DGT_CLOCK and HYDRO_GT_CLOCK=1; zeros otherwise); reading clears */
int hg_profile_read_clocks(hg_handle h, unsigned long long out[16]);

#ifdef __cplusplus
}
#endif
#endif /* HYDRO_GPU_H_ */
