// hydro_gpu.hpp -- C++ host shim over the C ABI (include/hydro_gpu.h).
//
// Mirrors the reference's solver-side interfaces so reference host code can drive the GPU path:
//   hg::Handle        RAII owner of an hg_handle; non-zero status -> throw std::string, the
//                     reference's error convention (fluid.hpp:795-797, control/module.cpp:103-107)
//   hg::FluidSolver   the solver::UnsteadyIterativeSolver protocol of FluidSimple
//                     (solver.hpp:710-751, fluid.hpp:202-253): StartStep / MakeIteration / IsConverged /
//                     FinishStep / GetConvergenceIndicator / GetIterationCount / GetTime / SetTimeStep /
//                     GetAutoTimeStep, and GetVelocity / GetPressure / GetVolumeFlux returning const
//                     references to lazily refreshed HOST mirrors (device -> host only when dirty)
//   hg::AdvectionSolver, hg::HeatSolver   the same for advection.hpp:62-84 and heat.hpp:19-93
// Header only; link with libhydro_gpu.so.  No CPU fallback: construction throws without a device.
#pragma once
#include <string>
#include <vector>

#include "../../include/hydro_gpu.h"

namespace hg {

class Handle {
 public:
  Handle() = default;
  explicit Handle(const hg_config& cfg) { Create(cfg); }
  Handle(const Handle&) = delete;
  Handle& operator=(const Handle&) = delete;
  ~Handle() { if (h_) hg_destroy(h_); }
  void Create(const hg_config& cfg) {
    if (h_) { hg_destroy(h_); h_ = nullptr; }
    cfg_ = cfg;
    if (int rc = hg_create(&cfg, &h_)) throw std::string("hydro_gpu: ") + hg_last_error(nullptr) + " (status " + std::to_string(rc) + ")";
  }
  hg_handle get() const { return h_; }
  const hg_config& config() const { return cfg_; }
  size_t NumCells() const { return hg_num_cells(h_); }
  size_t NumFaces() const { return hg_num_faces(h_); }
  void Check(int rc) const { if (rc) throw std::string(hg_last_error(h_)); }
  void Step(hg_step_stats* st = nullptr) { Check(hg_step(h_, st)); }
  void Get(int field, std::vector<double>& out) const {
    out.resize(field == HG_F_VOLUME_FLUX || field == HG_F_VOLUME_FLUX_PREV ? NumFaces() : NumCells());
    Check(hg_get_field(h_, field, out.data(), out.size()));
  }
  void Set(int field, const std::vector<double>& in) { Check(hg_set_field(h_, field, in.data(), in.size())); }

 private:
  hg_handle h_ = nullptr;
  hg_config cfg_{};
};

// Host mirror refreshed on demand: getter => D2H only if a device-side call invalidated it.
class Mirror {
 public:
  Mirror(const Handle* h, int field) : h_(h), field_(field) {}
  void Invalidate() { dirty_ = true; }
  const std::vector<double>& Get() const {
    if (dirty_) { h_->Get(field_, data_); dirty_ = false; }
    return data_;
  }
 private:
  const Handle* h_;
  int field_;
  mutable std::vector<double> data_;
  mutable bool dirty_ = true;
};

// solver::FluidSolver<Mesh> (fluid.hpp:202-253) on the device.  Vector fields are exposed per component
// (SoA) instead of the reference's FieldCell<Vect>.
class FluidSolver {
 public:
  explicit FluidSolver(Handle* h)
      : h_(h), u_{Mirror(h, HG_F_VELOCITY_X), Mirror(h, HG_F_VELOCITY_Y), Mirror(h, HG_F_VELOCITY_Z)},
        p_(h, HG_F_PRESSURE), flux_(h, HG_F_VOLUME_FLUX), time_(0.), dt_(h->config().dt) {}
  void StartStep() { iters_ = 0; h_->Check(hg_fluid_start_step(h_->get())); }
  void MakeIteration() { h_->Check(hg_fluid_make_iteration(h_->get())); ++iters_; }
  bool IsConverged() const { int c = 0; h_->Check(hg_fluid_is_converged(h_->get(), &c)); return c != 0; }
  void FinishStep() { h_->Check(hg_fluid_finish_step(h_->get())); time_ += dt_; Dirty(); }
  void CalcStep() { while (!IsConverged()) MakeIteration(); }   // solver.hpp:746-750
  double GetConvergenceIndicator() const { double r = 1.; h_->Check(hg_fluid_convergence_indicator(h_->get(), &r)); return r; }
  size_t GetIterationCount() const { return iters_; }
  double GetTime() const { return time_; }
  double GetTimeStep() const { return dt_; }
  void SetTimeStep(double dt, double dt_advection) { dt_ = dt; h_->Check(hg_set_time_step(h_->get(), dt, dt_advection)); }
  double GetAutoTimeStep() const { double v = 0.; h_->Check(hg_fluid_auto_time_step(h_->get(), &v)); return v; }
  const std::vector<double>& GetVelocity(int comp) const { return u_[comp].Get(); }
  const std::vector<double>& GetPressure() const { return p_.Get(); }
  const std::vector<double>& GetVolumeFlux() const { return flux_.Get(); }
 private:
  void Dirty() { for (auto& m : u_) m.Invalidate(); p_.Invalidate(); flux_.Invalidate(); }
  Handle* h_;
  Mirror u_[3], p_, flux_;
  double time_, dt_;
  size_t iters_ = 0;
};

// solver::AdvectionSolverMulti<Mesh, FieldFace<Scal>> (advection.hpp:62-84)
class AdvectionSolver {
 public:
  explicit AdvectionSolver(Handle* h)
      : h_(h), pd_{Mirror(h, HG_F_PARTIAL_DENSITY_0), Mirror(h, HG_F_PARTIAL_DENSITY_1), Mirror(h, HG_F_PARTIAL_DENSITY_2)} {}
  void Step() { h_->Check(hg_advection_step(h_->get())); for (auto& m : pd_) m.Invalidate(); }   // Start/CalcStep/Finish
  const std::vector<double>& GetField(size_t i) const { return pd_[i].Get(); }
 private:
  Handle* h_;
  Mirror pd_[3];
};

// solver::HeatSolver<Mesh> (heat.hpp:19-93)
class HeatSolver {
 public:
  explicit HeatSolver(Handle* h) : h_(h), t_(h, HG_F_TEMPERATURE) {}
  void Step() { h_->Check(hg_heat_step(h_->get())); t_.Invalidate(); }
  const std::vector<double>& GetTemperature() const { return t_.Get(); }
 private:
  Handle* h_;
  Mirror t_;
};

}  // namespace hg
