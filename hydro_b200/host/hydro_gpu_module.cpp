// hydro_gpu_module.cpp -- the reference-side plugin: registers the GPU path as two new MODULE aliases,
// `hydro3d_gpu` and `hydro2d_gpu`, in the reference's own module registry
// (ModuleRegistrator / TExperiment::module_map_, control/experiment.hpp:136-160), next to the untouched
// CPU modules `hydro3d` / `hydro2d` (hydro2dmpi/hydro3d.cpp:15-22).  A .hydroconf script switches to the
// GPU with one line:   set string MODULE hydro3d_gpu
//
// The class reads the SAME parameter names as hydro<Mesh> (hydro2d.hpp:250-687, examples/general.hydroconf),
// builds an hg_config, and forwards step() to hg_step() through the C ABI.  It is compiled against the
// reference's headers (never copied) and linked with the reference's objects and libhydro_gpu.so
// (hydro_b200/host/Makefile); it is the only reference-facing file a maintainer has to add.
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "control/experiment.hpp"
#include "hydro_gpu.hpp"

namespace hydro_gpu_module {

template <int DIM>
class hydro_gpu : public TModule {
  hg::Handle h_;
  hg_step_stats st_{};

  void vec(const char* name, double out[3]) {
    for (int d = 0; d < 3; ++d) out[d] = 0.;
    if (P_vect.exist(name)) {
      column<double>& v = P_vect[name];
      for (int d = 0; d < 3 && d < v.N; ++d) out[d] = v[d];
    }
  }
  static int linear_id(const std::string& s) {   // hydro<Mesh>::GetLinearSolverFactory, hydro2d.hpp:194-248
    if (s == "lu") return HG_LS_LU;
    if (s == "lu_relaxed") return HG_LS_LU_RELAXED;
    if (s == "gauss_seidel") return HG_LS_GAUSS_SEIDEL;
    if (s == "jacobi") return HG_LS_JACOBI;
    throw std::runtime_error("Unknown linear solver '" + s + "'");
  }
  void condition(const char* name, int side, hg_config& c) {   // solver::Parse, fluid.hpp:392-417
    std::stringstream arg(P_string[name]);
    std::string kind; arg >> kind;
    if (kind == "wall") c.condition_kind[side] = HG_BC_WALL;
    else if (kind == "inlet") c.condition_kind[side] = HG_BC_INLET;
    else if (kind == "outlet") c.condition_kind[side] = HG_BC_OUTLET;
    else throw std::runtime_error("Parse: Unknown boundary condition type");
    for (int d = 0; d < DIM; ++d) arg >> c.condition_velocity[side][d];
  }
  hg_config config() {
    hg_config c;
    hg_config_defaults(&c);
    c.dim = DIM;
    c.Nx = P_int["Nx"]; c.Ny = P_int["Ny"]; c.Nz = DIM == 3 ? P_int["Nz"] : 1;
    vec("A", c.A); vec("B", c.B); vec("box_A", c.box_A); vec("box_B", c.box_B);
    condition("condition_left", HG_SIDE_LEFT, c); condition("condition_right", HG_SIDE_RIGHT, c);
    condition("condition_bottom", HG_SIDE_BOTTOM, c); condition("condition_top", HG_SIDE_TOP, c);
    condition("condition_close", HG_SIDE_CLOSE, c); condition("condition_far", HG_SIDE_FAR, c);
    c.pressure_fixed_enable = P_vect.exist("pressure_fixed_point") ? 1 : 0;
    vec("pressure_fixed_point", c.pressure_fixed_point);
    c.pressure_fixed_value = ecast(P_double("pressure_fixed_value"));
    vec("initial_velocity", c.initial_velocity);
    c.initial_pois = flag("initial_pois") ? 1 : 0;
    c.initial_sin_enable = P_vect.exist("initial_sin_n") ? 1 : 0;
    if (c.initial_sin_enable) { vec("initial_sin_n", c.initial_sin_n); c.initial_sin_lambda = P_double["initial_sin_lambda"]; c.initial_sin_phase = P_double["initial_sin_phase"]; }
    vec("A1", c.A1); vec("B1", c.B1); vec("A2", c.A2); vec("B2", c.B2); vec("IC", c.IC); vec("IC2", c.IC2);
    c.IR = P_double["IR"]; c.IR2 = P_double["IR2"];
    c.num_phases = P_int["num_phases"];
    for (int i = 0; i < c.num_phases && i < HG_MAX_PHASES; ++i) {
      const std::string k = IntToStr(i);
      c.density[i] = P_double["density_" + k]; c.viscosity[i] = P_double["viscosity_" + k];
      c.conductivity[i] = P_double["conductivity_" + k];
      c.initial_volume_fraction[i] = ecast(P_double("initial_volume_fraction_" + k));
    }
    c.initial_volume_fraction_smooth_times = P_int["initial_volume_fraction_smooth_times"];
    c.dt = dt; c.dt_auto = flag("dt_auto") ? 1 : 0; c.cfl = P_double["cfl"]; c.cfl_advection = P_double["cfl_advection"];
    vec("gravity", c.gravity); vec("force", c.force); c.sigma = P_double["sigma"];
    c.fluid_enable = P_bool["fluid_enable"]; c.advection_enable = P_bool["advection_enable"];
    c.convergence_tolerance = P_double["convergence_tolerance"]; c.num_iterations_limit = P_int["num_iterations_limit"];
    c.velocity_relaxation_factor = P_double["velocity_relaxation_factor"];
    c.pressure_relaxation_factor = P_double["pressure_relaxation_factor"];
    c.rhie_chow_factor = P_double["rhie_chow_factor"];
    c.time_second_order = P_bool["time_second_order"]; c.simpler = P_bool["simpler"];
    c.force_geometric_average = P_bool["force_geometric_average"];
    c.guess_extrapolation = P_double["guess_extrapolation"];
    vec("meshvel", c.meshvel); c.meshvel_output = flag("meshvel_output") ? 1 : 0;
    c.linear_solver_velocity = linear_id(P_string["linear_solver_velocity"]);
    c.linear_solver_pressure = linear_id(P_string["linear_solver_pressure"]);
    c.linear_solver_heat = linear_id(P_string["linear_solver_heat"]);
    c.lu_relaxed_tolerance = P_double["lu_relaxed_tolerance"];
    c.lu_relaxed_num_iters_limit = P_int["lu_relaxed_num_iters_limit"];
    c.lu_relaxed_relaxation_factor = P_double["lu_relaxed_relaxation_factor"];
    c.density_smooth_times = P_int["density_smooth_times"]; c.viscosity_smooth_times = P_int["viscosity_smooth_times"];
    c.force_smooth_times = P_int["force_smooth_times"];
    if (P_string["advection_solver"] != "tvd") throw std::runtime_error("hydro_gpu: only advection_solver tvd");
    // options that change the reference's results and are not on the GPU path fail loudly (never silently dropped)
    for (const char* k : {"compressible_enable", "deforming_velocity", "radiation_enable"})
      if (flag(k)) throw std::runtime_error(std::string("hydro_gpu: ") + k + " 1 is not on the GPU path");
    if ((P_string.exist("chemistry") && P_string["chemistry"] != "steady") || (P_double.exist("chem_intensity") && P_double["chem_intensity"] != 0.))
      throw std::runtime_error("hydro_gpu: chemistry other than 'steady' with chem_intensity 0 is not on the GPU path");
    for (const char* k : {"meshvel_auto", "imgu_init", "imgv_init", "img_init"})
      if (P_string.exist(k)) throw std::runtime_error(std::string("hydro_gpu: ") + k + " is not on the GPU path");
    for (int i = 0; i < c.num_phases; ++i)
      if (flag("enable_settling_" + IntToStr(i))) throw std::runtime_error("hydro_gpu: phase slip (enable_settling) is not on the GPU path");
    c.advection_dt_factor = P_double["advection_dt_factor"]; c.tvd_split = P_bool["tvd_split"]; c.sharp = P_double["sharp"];
    c.heat_enable = P_bool["heat_enable"]; c.temperature_initial = P_double["temperature_initial"];
    vec("heat_box_lb", c.heat_box_lb); vec("heat_box_rt", c.heat_box_rt);
    c.heat_box_temperature = P_double["heat_box_temperature"]; c.heat_relaxation_factor = P_double["heat_relaxation_factor"];
    c.time_second_order_heat = P_bool["time_second_order_heat"];
    return c;
  }
  void publish_stat() {   // the P_double keys hydro<Mesh>::CalcStat sets (hydro2d.hpp:1449-1466)
    for (int i = 0; i < h_.config().num_phases; ++i) {
      const std::string k = IntToStr(i);
      P_double.set("stat_volume_" + k, st_.volume[i]); P_double.set("stat_mass_" + k, st_.mass[i]);
      P_double.set("stat_pd_min_" + k, st_.pd_min[i]); P_double.set("stat_pd_max_" + k, st_.pd_max[i]);
      P_double.set("stat_cx_" + k, st_.center[i][0]); P_double.set("stat_vx_" + k, st_.velocity[i][0]);
      P_double.set("stat_cy_" + k, st_.center[i][1]); P_double.set("stat_vy_" + k, st_.velocity[i][1]);
      if (DIM > 2) { P_double.set("stat_cz_" + k, st_.center[i][2]); P_double.set("stat_vz_" + k, st_.velocity[i][2]); }
    }
  }

 public:
  explicit hydro_gpu(TExperiment* _ex) : TExperiment_ref(_ex), TModule(_ex) {
    P_int.set("last_s", 0); P_double.set("last_R", 0); P_double.set("last_Rn", 0);
    P_int.set("s_sum", 0); P_int.set("s_max", 0); P_int.set("s", 0);
    h_.Create(config());                       // throws std::string on failure, like the reference
    P_int.set("cells_number", static_cast<int>(h_.NumCells()));
    h_.Check(hg_get_stats(h_.get(), &st_));   // the CalcStat of the constructor (hydro2d.hpp:962) ran inside hg_create
    publish_stat();
  }
  void step() override {
    ex->timer_.Push("step");
    h_.Step(&st_);
    ex->timer_.Pop();
    dt = st_.dt; P_double["dt"] = dt;
    P_int["s"] = st_.simple_iterations; P_int["s_sum"] += st_.simple_iterations;
    {   // the per-iteration lines hydro<Mesh>::step() logs (hydro2d.hpp:1585-1586)
      std::vector<double> rs(st_.simple_iterations > 0 ? st_.simple_iterations : 1);
      int n = 0;
      h_.Check(hg_last_residuals(h_.get(), rs.data(), static_cast<int>(rs.size()), &n));
      for (int k = 0; k < n; ++k) logger() << ".....s=" << (k + 1) << ", Rs=" << rs[k];
    }
    publish_stat();
  }
  void write_results(bool force = false) override {
    // mesh frames (.vts) are the next scope row; the scalar series is logged
    if (force && !ecast(P_bool("no_output")))
      logger() << "gpu final: t=" << st_.time << " volume_0=" << st_.volume[0] << " pressure sweeps=" << st_.pressure_sweeps_total;
  }
};

namespace registrators {
ModuleRegistrator<hydro_gpu<3>> reg_3d_gpu({"hydro3d_gpu", "hydro3D_uniform_GPU"});
ModuleRegistrator<hydro_gpu<2>> reg_2d_gpu({"hydro2d_gpu", "hydro2D_uniform_GPU"});
}  // namespace registrators

}  // namespace hydro_gpu_module
