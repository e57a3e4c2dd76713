// shim_check.cpp -- drives the GPU path through the C++ solver shim (hydro_gpu.hpp) the way reference host code
// drives solver::FluidSolver / AdvectionSolverMulti (hydro<Mesh>::step, hydro2d.hpp:1531-1621: StartStep, `while
// (!IsConverged()) MakeIteration()`, FinishStep, advection step, properties, statistics) and dumps the host mirrors
// the getters return.  tests/test_dropin_module.py compares the dump with hg_step() on the same configuration.
//   shim_check <hg_config bytes file> <nsteps> <output file>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "hydro_gpu.hpp"

static void dump(FILE* f, const std::vector<double>& v) {
  const unsigned long long n = v.size();
  fwrite(&n, sizeof n, 1, f);
  fwrite(v.data(), sizeof(double), v.size(), f);
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: shim_check config.bin nsteps out.bin\n"); return 2; }
  hg_config cfg;
  FILE* fc = fopen(argv[1], "rb");
  if (!fc || fread(&cfg, 1, sizeof cfg, fc) != sizeof cfg) { fprintf(stderr, "bad config file (struct size %zu)\n", sizeof cfg); return 2; }
  fclose(fc);
  const int nsteps = atoi(argv[2]);
  try {
    hg::Handle h(cfg);
    hg::FluidSolver fluid(&h);
    hg::AdvectionSolver adv(&h);
    size_t iters = 0;
    for (int s = 0; s < nsteps; ++s) {
      fluid.StartStep();
      while (!fluid.IsConverged()) fluid.MakeIteration();
      iters += fluid.GetIterationCount();
      fluid.FinishStep();
      if (cfg.advection_enable) adv.Step();   // advection_dt_factor 1: one sub-step per time step
      h.Check(hg_update_properties(h.get()));
      hg_step_stats st;
      h.Check(hg_calc_stat(h.get(), &st));
    }
    FILE* fo = fopen(argv[3], "wb");
    if (!fo) return 2;
    for (int d = 0; d < cfg.dim; ++d) dump(fo, fluid.GetVelocity(d));
    dump(fo, fluid.GetPressure());
    dump(fo, fluid.GetVolumeFlux());
    for (int p = 0; p < cfg.num_phases; ++p) dump(fo, adv.GetField(p));
    fclose(fo);
    printf("iterations %zu time %.17g indicator %.17g\n", iters, fluid.GetTime(), fluid.GetConvergenceIndicator());
  } catch (const std::string& e) {   // the reference's error convention
    fprintf(stderr, "error: %s\n", e.c_str());
    return 1;
  }
  return 0;
}
