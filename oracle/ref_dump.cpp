// oracle/ref_dump.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference (headers and .cpp files compiled where they
// lie under /root/reference, see oracle/Makefile) and dumps raw fp64 fields so
// that (i) the C restatement in oracle/hydro_oracle.c can be pinned against the
// real reference and (ii) golden fixtures for tests/golden/ can be generated
// (oracle/make_golden.py).  No reference source is copied: this file only
// includes the reference's headers and calls its public/console API.
//
// usage: ref_dump <script.hydroconf> <nsteps> <outdir> [--iters]
//   <script> must `ae` an experiment, set parameters and call `init`
//   (no `start`).  Without --iters each step is the reference's own
//   hydro<Mesh>::step() (hydro2d.hpp:1531-1621).  With --iters the same
//   sequence of solver calls is issued from here so that the convergence
//   indicator of every SIMPLE iteration can be recorded at full precision.

#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

// The module keeps its solvers private; the dump needs to read them, so this
// one translation unit is compiled with g++ -fno-access-control (oracle/Makefile).
#include "hydro2dmpi/hydro2d.hpp"
#include "control/console.hpp"

namespace {

void WriteNpy(const std::string& path, const std::vector<double>& v) {
  // NumPy .npy v1.0, little-endian float64, 1-D
  std::ostringstream h;
  h << "{'descr': '<f8', 'fortran_order': False, 'shape': (" << v.size()
    << ",), }";
  std::string hs = h.str();
  size_t total = 10 + hs.size() + 1;
  size_t pad = (64 - total % 64) % 64;
  hs += std::string(pad, ' ');
  hs += "\n";
  std::ofstream f(path, std::ios::binary);
  const char magic[] = "\x93NUMPY";
  f.write(magic, 6);
  char ver[2] = {1, 0};
  f.write(ver, 2);
  uint16_t hl = static_cast<uint16_t>(hs.size());
  f.write(reinterpret_cast<const char*>(&hl), 2);
  f.write(hs.data(), hs.size());
  f.write(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(double));
}

template <class Field>
std::vector<double> Flat(const Field& f) {
  std::vector<double> r;
  r.reserve(f.size());
  for (auto idx : f.GetRange()) {
    r.push_back(static_cast<double>(f[idx]));
  }
  return r;
}

template <class Field>
std::vector<double> FlatComp(const Field& f, size_t n) {
  std::vector<double> r;
  r.reserve(f.size());
  for (auto idx : f.GetRange()) {
    r.push_back(f[idx][n]);
  }
  return r;
}

template <class Mesh>
int Run(hydro2D_uniform_MPI::hydro<Mesh>* mod, TExperiment* ex, int nsteps,
        const std::string& outdir, bool iters) {
  using solver::Layers;
  constexpr size_t dim = Mesh::dim;
  std::vector<double> rs_hist, niter_hist, dt_hist, nadv_hist;

  for (int n = 0; n < nsteps; ++n) {
    if (!iters) {
      mod->step();
      niter_hist.push_back(
          static_cast<double>(mod->fluid_solver->GetIterationCount()));
    } else {
      // same call sequence as hydro<Mesh>::step(), hydro2d.hpp:1531-1621
      auto& P_double = ex->P_double;
      auto& P_bool = ex->P_bool;
      if (ex->flag("dt_auto")) {
        double cfl = P_double["cfl"];
        double cfla = P_double["cfl_advection"];
        double dtm = mod->fluid_solver->GetAutoTimeStep();
        mod->dt = dtm * cfl;
        P_double["dt"] = mod->dt;
        mod->fluid_solver->SetTimeStep(mod->dt);
        mod->advection_solver->SetTimeStep(dtm * cfla);
      }
      mod->fluid_solver->StartStep();
      if (P_bool["fluid_enable"]) {
        while (!mod->fluid_solver->IsConverged()) {
          mod->fluid_solver->MakeIteration();
          rs_hist.push_back(mod->fluid_solver->GetConvergenceIndicator());
        }
      }
      niter_hist.push_back(
          static_cast<double>(mod->fluid_solver->GetIterationCount()));
      mod->fluid_solver->FinishStep();
      int nadv = 0;
      if (P_bool["advection_enable"]) {
        while (mod->advection_solver->GetTime() <
               mod->fluid_solver->GetTime() -
                   0.5 * mod->advection_solver->GetTimeStep()) {
          mod->advection_solver->StartStep();
          mod->advection_solver->CalcStep();
          mod->advection_solver->FinishStep();
          ++nadv;
        }
      }
      nadv_hist.push_back(nadv);
      if (P_bool["heat_enable"]) {
        mod->heat_solver->StartStep();
        mod->heat_solver->CalcStep();
        mod->heat_solver->FinishStep();
      }
      mod->UpdateFluidProperties();
      mod->CalcStat();
    }
    dt_hist.push_back(mod->dt);
    mod->increase_time();
  }

  auto& mesh = mod->mesh;
  auto out = [&outdir](const std::string& name, const std::vector<double>& v) {
    WriteNpy(outdir + "/" + name + ".npy", v);
  };
  const auto& vel = mod->fluid_solver->GetVelocity();
  for (size_t d = 0; d < dim; ++d) {
    out("u" + std::to_string(d), FlatComp(vel, d));
    out("force" + std::to_string(d), FlatComp(mod->fc_force, d));
    out("stforce" + std::to_string(d), FlatComp(mod->fc_stforce, d));
  }
  out("p", Flat(mod->fluid_solver->GetPressure()));
  out("flux", Flat(mod->fluid_solver->GetVolumeFlux()));
  out("rho", Flat(mod->fc_density_smooth));
  out("mu", Flat(mod->fc_viscosity_smooth));
  for (size_t i = 0; i < mod->num_phases; ++i) {
    out("pd" + std::to_string(i), Flat(mod->advection_solver->GetField(i)));
    out("vf" + std::to_string(i), Flat(mod->v_fc_volume_fraction[i]));
  }
  if (ex->flag("heat_enable")) {
    out("temp", Flat(mod->heat_solver->GetTemperature()));
  }
  {
    std::vector<double> excl;
    for (auto c : mesh.Cells()) {
      excl.push_back(mesh.IsExcluded(c) ? 1. : 0.);
    }
    out("excluded", excl);
  }
  out("rs", rs_hist);
  out("niter", niter_hist);
  out("dt", dt_hist);
  out("nadv", nadv_hist);
  // per-phase statistics of the last CalcStat (hydro2d.hpp:1432-1529)
  {
    std::vector<double> st;
    for (size_t i = 0; i < mod->num_phases; ++i) {
      std::string s = IntToStr(i);
      st.push_back(ex->P_double["stat_volume_" + s]);
      st.push_back(ex->P_double["stat_mass_" + s]);
      st.push_back(ex->P_double["stat_pd_min_" + s]);
      st.push_back(ex->P_double["stat_pd_max_" + s]);
      st.push_back(ex->P_double["stat_cx_" + s]);
      st.push_back(ex->P_double["stat_cy_" + s]);
      st.push_back(dim > 2 ? ex->P_double["stat_cz_" + s] : 0.);
      st.push_back(ex->P_double["stat_vx_" + s]);
      st.push_back(ex->P_double["stat_vy_" + s]);
      st.push_back(dim > 2 ? ex->P_double["stat_vz_" + s] : 0.);
    }
    out("stat", st);
  }
  return 0;
}

}  // namespace

int main(int argc, char* argv[]) {
  if (argc < 4) {
    std::cerr << "usage: ref_dump <script> <nsteps> <outdir> [--iters]\n";
    return 2;
  }
  std::string script = argv[1];
  int nsteps = std::atoi(argv[2]);
  std::string outdir = argv[3];
  bool iters = (argc > 4 && std::string(argv[4]) == "--iters");
  try {
    TConsole console;
    console.cmd_run(script);
    TExperiment* ex = console.check_cur_exp();
    if (!ex->st_init || !ex->module) {
      std::cerr << "script did not init an experiment\n";
      return 3;
    }
    using M2 = geom::geom2d::MeshStructured<double>;
    using M3 = geom::geom3d::MeshStructured<double>;
    int rc = 4;
    if (auto m3 = dynamic_cast<hydro2D_uniform_MPI::hydro<M3>*>(
            ex->module.get())) {
      rc = Run(m3, ex, nsteps, outdir, iters);
    } else if (auto m2 = dynamic_cast<hydro2D_uniform_MPI::hydro<M2>*>(
                   ex->module.get())) {
      rc = Run(m2, ex, nsteps, outdir, iters);
    } else {
      std::cerr << "unknown module type\n";
    }
    std::cout.flush();
    std::_Exit(rc);  // skip console/scheduler teardown
  } catch (std::string msg) {
    std::cerr << "ERROR: " << msg << std::endl;
    return 5;
  } catch (std::exception& e) {
    std::cerr << "ERROR: " << e.what() << std::endl;
    return 5;
  }
}
