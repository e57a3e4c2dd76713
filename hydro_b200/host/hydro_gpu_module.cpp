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
#include <fstream>
#include <memory>
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
  // output (hydro<Mesh>::InitOutput / write_results, hydro2d.hpp:689-927, 1623-1652): the reference's ParaView session
  // (output_paraview.hpp:34-173) and plain scalar session (output.hpp:230-264) restated over fields fetched from the device at
  // frame times only; same files, same schedule, values printed by the same stream formatting
  bool no_output_ = false;
  std::unique_ptr<std::ofstream> pvd_, scalar_;
  std::string filename_field_;
  size_t timestep_ = 0;
  double last_frame_time_ = 0., last_frame_scalar_time_ = 0.;
  std::vector<std::string> scalar_names_;

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
    for (const char* k : {"imgu_init", "imgv_init", "img_init"})
      if (P_string.exist(k)) throw std::runtime_error(std::string("hydro_gpu: ") + k + " is not on the GPU path");
    if (std::string* ma = P_string("meshvel_auto")) {   // hydro2d.hpp:1510-1524
      if (*ma == "vx") c.meshvel_auto = 1; else if (*ma == "vcx") c.meshvel_auto = 2;
      else throw std::runtime_error("hydro_gpu: Unknown meshvel_auto=" + *ma);
      c.meshvel_weight = P_double["meshvel_weight"];
    }
    for (int i = 0; i < c.num_phases; ++i)
      if (flag("enable_settling_" + IntToStr(i))) { c.enable_settling[i] = 1; c.bubble_radius[i] = P_double["bubble_radius_" + IntToStr(i)]; }
    if (flag("velocity_is_carrier")) throw std::runtime_error("hydro_gpu: velocity_is_carrier 1 is not on the GPU path");
    if (P_double.exist("antidiffusion_factor") && P_double["antidiffusion_factor"] != 0.) throw std::runtime_error("hydro_gpu: antidiffusion_factor is not on the GPU path");
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

  bool out_on(const std::string& name) { const std::string k = "output_" + name; return P_bool.exist(k) && P_bool[k]; }
  void init_output() {
    no_output_ = P_bool["no_output"];
    if (no_output_) return;
    for (const char* k : {"output_factor_x", "output_factor_y", "output_factor_z"})
      if (P_int.exist(k) && P_int[k] != 1 && (DIM == 3 || std::string(k) != "output_factor_z"))
        throw std::runtime_error("hydro_gpu: output_factor_* other than 1 is not on the GPU path");
    for (const char* k : {"radiation", "volume_source", "divergence", "curvature"})
      if (out_on(k)) throw std::runtime_error(std::string("hydro_gpu: output_") + k + " is not on the GPU path");
    for (int i = 0; i < h_.config().num_phases; ++i)
      for (const char* k : {"mass_source_", "mass_fraction_", "density_", "target_density_"})
        if (out_on(k + IntToStr(i))) throw std::runtime_error(std::string("hydro_gpu: output_") + k + IntToStr(i) + " is not on the GPU path");
    if (P_string["field_output_format"] != "paraview") throw std::runtime_error("hydro_gpu: field_output_format must be paraview");
    // scalar content (hydro2d.hpp:824-893); the in / out flux statistics (hydro2d.hpp:1473-1507) are zero for closed walls
    // and are not accumulated here
    scalar_names_ = {"time", "num_iters", "iter_diff_velocity"};
    const int np = h_.config().num_phases;
    for (int i = 0; i < np; ++i) for (const char* k : {"mass_", "mass_in_", "mass_out_"}) scalar_names_.push_back(k + IntToStr(i));
    for (int i = 0; i < np; ++i) for (const char* k : {"volume_", "volume_in_", "volume_out_"}) scalar_names_.push_back(k + IntToStr(i));
    for (int i = 0; i < np; ++i) for (const char* k : {"pd_min_", "pd_max_"}) scalar_names_.push_back(k + IntToStr(i));
    for (int i = 0; i < np; ++i) {
      scalar_names_.push_back("cx_" + IntToStr(i)); scalar_names_.push_back("cy_" + IntToStr(i));
      if (DIM > 2) scalar_names_.push_back("cz_" + IntToStr(i));
      scalar_names_.push_back("vx_" + IntToStr(i)); scalar_names_.push_back("vy_" + IntToStr(i));
      if (DIM > 2) scalar_names_.push_back("vz_" + IntToStr(i));
    }
    if (!P_string.exist(_plt_title)) P_string.set(_plt_title, P_string[_exp_name]);
    if (!P_string.exist("filename_field")) P_string.set("filename_field", P_string[_exp_name] + ".field");
    if (!P_string.exist("filename_scalar")) P_string.set("filename_scalar", P_string[_exp_name] + ".scalar");
    filename_field_ = P_string["filename_field"];
    pvd_.reset(new std::ofstream(filename_field_ + ".pvd"));
    *pvd_ << "<?xml version=\"1.0\"?>\n<VTKFile type=\"Collection\" version=\"0.1\" byte_order=\"LittleEndian\">\n  <Collection>\n";
    scalar_.reset(new std::ofstream(P_string["filename_scalar"] + ".dat"));
    for (auto& n : scalar_names_) *scalar_ << n << " ";
    *scalar_ << std::endl;
    for (int i = 0; i < np; ++i)
      for (const char* k : {"stat_mass_in_", "stat_mass_out_", "stat_volume_in_", "stat_volume_out_"})
        if (!P_double.exist(k + IntToStr(i))) P_double.set(k + IntToStr(i), 0.);
  }
  void write_scalar() {   // SessionPlainScalar::Write
    double rs = 1.;
    h_.Check(hg_fluid_convergence_indicator(h_.get(), &rs));
    for (auto& n : scalar_names_) {
      if (n == "time") *scalar_ << P_double["t"] << " ";
      else if (n == "num_iters") *scalar_ << static_cast<double>(P_int["s"]) << " ";
      else if (n == "iter_diff_velocity") *scalar_ << rs << " ";
      else *scalar_ << P_double["stat_" + n] << " ";
    }
    *scalar_ << std::endl;
  }
  void data_array(std::ostream& out, const std::string& name, int ncomp) {
    out << "        <DataArray Name=\"" << name << "\" NumberOfComponents=\"" << ncomp << "\" type=\"Float32\" format=\"ascii\">\n";
  }
  void write_frame(double time) {   // SessionParaviewStructured::Write
    const std::string datafile = filename_field_ + "." + IntToStr(static_cast<int>(timestep_)) + ".vts";
    *pvd_ << "    <DataSet timestep=\"" << time << "\" group=\"\" part=\"0\" file=\"" << datafile << "\"/>\n";
    pvd_->flush();
    ++timestep_;
    const hg_config& c = h_.config();
    const int n[3] = {c.Nx, c.Ny, DIM == 3 ? c.Nz : 0};
    std::ofstream out(datafile);
    out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"StructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n";
    for (const char* tag : {"  <StructuredGrid WholeExtent=\"", "    <Piece Extent=\""})
      out << tag << "0 " << n[0] << " 0 " << n[1] << " 0 " << n[2] << "\">\n";
    // node coordinates: InitUniformMesh (mesh.hpp:722-736), lb + (midx / mesh_size) * (rt - lb), x fastest
    auto node = [&](int d, int i) { return d < DIM ? c.A[d] + (static_cast<double>(i) / n[d]) * (c.B[d] - c.A[d]) : 0.; };
    const int nn[3] = {n[0] + 1, n[1] + 1, DIM == 3 ? n[2] + 1 : 1};
    out << "      <PointData>\n";
    for (int d = 0; d < 3; ++d) {
      const std::string nm(1, "xyz"[d]);
      if (!out_on(nm)) continue;
      data_array(out, nm, 1);
      for (int k = 0; k < nn[2]; ++k) for (int j = 0; j < nn[1]; ++j) for (int i = 0; i < nn[0]; ++i) {
        const int idx[3] = {i, j, k};
        out << node(d, idx[d]) << " ";
      }
      out << "\n        </DataArray>\n";
    }
    out << "      </PointData>\n      <CellData>\n";
    std::vector<double> buf;
    auto cell = [&](const std::string& nm, int field) {
      if (!out_on(nm)) return;
      data_array(out, nm, 1);
      if (field < 0) buf.assign(h_.NumCells(), 0.); else h_.Get(field, buf);
      for (double v : buf) out << v << " ";
      out << "\n        </DataArray>\n";
    };
    // content pool order (hydro2d.hpp:711-818)
    cell("velocity_x", HG_F_VELOCITY_X); cell("velocity_y", HG_F_VELOCITY_Y); cell("velocity_z", DIM == 3 ? HG_F_VELOCITY_Z : -1);
    cell("pressure", HG_F_PRESSURE); cell("density", HG_F_DENSITY); cell("viscosity", HG_F_VISCOSITY);
    if (c.heat_enable) cell("temperature", HG_F_TEMPERATURE); else cell("temperature", -1);
    cell("excluded", HG_F_EXCLUDED);
    for (int i = 0; i < c.num_phases; ++i) cell("partial_density_" + IntToStr(i), HG_F_PARTIAL_DENSITY_0 + i);
    for (int i = 0; i < c.num_phases; ++i) cell("volume_fraction_" + IntToStr(i), HG_F_VOLUME_FRACTION_0 + i);
    out << "      </CellData>\n      <Points>\n";
    data_array(out, "mesh", 3);
    for (int k = 0; k < nn[2]; ++k) for (int j = 0; j < nn[1]; ++j) for (int i = 0; i < nn[0]; ++i)
      out << node(0, i) << " " << node(1, j) << " " << node(2, k) << " ";
    out << "        </DataArray>\n      </Points>\n    </Piece>\n  </StructuredGrid>\n</VTKFile>\n";
  }

 public:
  ~hydro_gpu() { if (pvd_) *pvd_ << "  </Collection>\n</VTKFile>\n"; }
  explicit hydro_gpu(TExperiment* _ex) : TExperiment_ref(_ex), TModule(_ex) {
    P_int.set("last_s", 0); P_double.set("last_R", 0); P_double.set("last_Rn", 0);
    P_int.set("s_sum", 0); P_int.set("s_max", 0); P_int.set("s", 0);
    P_int.set("current_frame", 0); P_int.set("current_frame_scalar", 0);   // hydro2d.hpp:952-953
    h_.Create(config());                       // throws std::string on failure, like the reference
    P_int.set("cells_number", static_cast<int>(h_.NumCells()));
    h_.Check(hg_get_stats(h_.get(), &st_));   // the CalcStat of the constructor (hydro2d.hpp:962) ran inside hg_create
    publish_stat();
    init_output();
    if (!no_output_) {   // hydro2d.hpp:966-971
      if (P_int["max_frame_index"] > 0) write_frame(0.);
      write_scalar();
    }
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
  void write_results(bool force = false) override {   // hydro<Mesh>::write_results, hydro2d.hpp:1623-1652
    if (no_output_) return;
    const double time = st_.time;
    const double total_time = P_double["T"];
    const size_t max_frame_index = P_int["max_frame_index"];
    const double frame_duration = total_time / max_frame_index;
    if (force || (!ecast(P_bool("no_mesh_output")) && time >= last_frame_time_ + frame_duration)) {
      last_frame_time_ += frame_duration;
      write_frame(time);
      logger() << "Frame " << (P_int["current_frame"]++) << ": t=" << time;
    }
    const size_t max_frame_scalar_index = P_int["max_frame_scalar_index"];
    const double frame_scalar_duration = total_time / max_frame_scalar_index;
    if (force || time >= last_frame_scalar_time_ + frame_scalar_duration) {
      last_frame_scalar_time_ = time;
      write_scalar();
      logger() << "Frame_scalar " << (P_int["current_frame_scalar"]++) << ": t=" << time;
    }
  }
};

namespace registrators {
ModuleRegistrator<hydro_gpu<3>> reg_3d_gpu({"hydro3d_gpu", "hydro3D_uniform_GPU"});
ModuleRegistrator<hydro_gpu<2>> reg_2d_gpu({"hydro2d_gpu", "hydro2D_uniform_GPU"});
}  // namespace registrators

}  // namespace hydro_gpu_module
