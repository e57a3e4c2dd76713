// hg_kernels.cuh -- CUDA kernels of the per-time-step path (sm_100a, fp64, -fmad=false).
//
// One thread per cell, x fastest (coalesced 8-byte loads along i); face quantities are
// recomputed from cell values instead of being stored as face fields (the reference
// materialises a FieldFace for every Interpolate call).  Each kernel cites the reference
// code it replaces; the arithmetic order follows oracle/hydro_oracle.c.
#pragma once
#include <type_traits>
#include <cooperative_groups.h>
#include "hg_device.cuh"

namespace cg = cooperative_groups;

// thread -> cell (i, j, k), raw index c (mesh.hpp:552-561).  32-bit divisions when the mesh allows it (64-bit
// division / modulo costs a few hundred instructions per thread, comparable to a whole stencil kernel body).
#define CELL_LOOP_PROLOG(g)                                                   \
  long long c_ = (long long)blockIdx.x * blockDim.x + threadIdx.x;            \
  long long nc_ = (long long)(g).n[0] * (g).n[1] * (g).n[2];                  \
  if ((g).cells) { if (c_ >= (g).ncells) return; c_ = (g).cells[c_]; }        \
  if (c_ >= nc_) return;                                                      \
  int i, j, k;                                                                \
  if (nc_ < (1LL << 31)) {                                                    \
    const unsigned c32_ = (unsigned)c_, nx_ = (unsigned)(g).n[0], nxy_ = nx_ * (unsigned)(g).n[1]; \
    const unsigned k_ = c32_ / nxy_, r_ = c32_ - k_ * nxy_, j_ = r_ / nx_;    \
    k = (int)k_; j = (int)j_; i = (int)(r_ - j_ * nx_);                       \
  } else {                                                                    \
    i = (int)(c_ % (g).n[0]);                                                 \
    j = (int)((c_ / (g).n[0]) % (g).n[1]);                                    \
    k = (int)(c_ / ((long long)(g).n[0] * (g).n[1]));                         \
  }                                                                           \
  const long long c = c_;

struct P3 { double* p[3]; };
struct CP3 { const double* p[3]; };
struct P9 { double* p[9]; };
struct P7 { double* p[7]; };
// hyperplane-major row arrays of k_gs_tiled (Co5 / gt_co5_index in hg_gs_tiled.cuh), array 0: what a kernel of this file needs
struct Co5Fwd { long long plane; int nxp, pad; };
HD long long co5fwd_index(const Co5Fwd& c, int i, int j, int k) { return (long long)(i + j + k + 1 + c.pad) * c.plane + (long long)j * c.nxp + i; }

// ---------------------------------------------------------------- small utilities
__global__ void k_fill(double* a, double v, long long n) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) a[t] = v;
}
// StartStep: iter_curr = time_curr + (time_curr - time_prev) * guess_extrapolation
// (fluid.hpp:800-811, conv_diff.hpp:123-128)
// nanflag != nullptr: also the IsNan scan of time_curr (solver.hpp:17-30, fluid.hpp:795-799) over the first n_scan entries
__global__ void k_start_layer(double* __restrict__ ic, const double* __restrict__ tc,
                              const double* __restrict__ tp, double ge, long long n, int* nanflag = nullptr, long long n_scan = 0) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    const double v = tc[t];
    ic[t] = v + (v - tp[t]) * ge;
    if (nanflag && t < n_scan && !(v * 0. == 0.)) *nanflag = 1;
  }
}
// IsNan scans of up to four arrays in one pass
struct Nan4 { const double* a[4]; int* flag[4]; int n; };
__global__ void k_nan_flag4(Nan4 q, long long n) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
#pragma unroll
  for (int m = 0; m < 4; ++m) if (m < q.n && !(q.a[m][t] * 0. == 0.)) *q.flag[m] = 1;
}
// CalcDiff over a cell list (the shell of the interior kernels)
__global__ void k_resid_list(Geo g, CP3 ic, CP3 ip, double* out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.;
  if (t < g.ncells) {
    const long long c = g.cells[t];
    double sq = 0.;
#pragma unroll
    for (int d = 0; d < 3; ++d) { const double e = ip.p[d][c] - ic.p[d][c]; sq += e * e; }
    v = sqrt(sq);
    if (!(v == v)) v = 0.;
  }
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0 && v > 0.) atomic_max_nonneg(out, v);
}
__global__ void k_flag_to_double(const int* flag, double* out) { *out = *flag ? 1. : 0.; }
// IsNan scan (solver.hpp:17-30): sets *flag when !(a*0 == 0)
__global__ void k_nan_flag(const double* __restrict__ a, long long n, int* flag) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && !(a[t] * 0. == 0.)) *flag = 1;
}

// `iter` and `diff` of a finished Gauss-Seidel solve from its per-sweep norms: the loop
// `do { sweep } while (diff > tol && iter++ < limit)` (linear.hpp:688-710) stops at the first sweep whose norm is not
// above the tolerance.  out = {iter, diff}.  With tol == 0 the sweeps after that one are fixed points (all corrections
// are exactly zero), so running all limit+1 sweeps leaves the same solution and only the reported count differs.
__global__ void k_sor_result(const double* __restrict__ diffs, int max_total, double tol, double* __restrict__ out) {
  int stop = -1;
  for (int k = 0; k < max_total; ++k) if (!(diffs[k] > tol)) { stop = k; break; }
  out[0] = stop >= 0 ? (double)stop : (double)max_total;   // `iter++ < limit` increments even when it fails
  out[1] = stop >= 0 ? diffs[stop] : diffs[max_total - 1];
}
// End of a time step: everything the host wants to know, as doubles behind the statistics (one transfer, one wait):
// out[0..7] NaN flags, [8] / [9] abort words of the sweep / lu dataflow kernels, [10] convergence indicator of the last
// SIMPLE iteration, [11] sum of iter+1 over the step's deferred pressure solves, [12] diff of the last one
__global__ void k_status_pack(const int* __restrict__ nanflags, const int* gt_ctl, const int* lt_ctl, const double* __restrict__ resid_last,
                              const double* __restrict__ sorres, int nsolves, double* __restrict__ out) {
  if (threadIdx.x < 8) out[threadIdx.x] = nanflags[threadIdx.x] ? 1. : 0.;
  if (threadIdx.x == 8) out[8] = gt_ctl ? (double)gt_ctl[1] : 0.;
  if (threadIdx.x == 9) out[9] = lt_ctl ? (double)lt_ctl[1] : 0.;
  if (threadIdx.x == 10) out[10] = resid_last ? *resid_last : 1.;
  if (threadIdx.x == 11) {
    double sum = 0., last = 0.;
    for (int m = 0; m < nsolves; ++m) { sum += sorres[2 * m] + 1.; last = sorres[2 * m + 1]; }
    out[11] = sum; out[12] = last;
  }
}

// ---------------------------------------------------------------- GetSmoothField
// One repeat of Average(Interpolate(u, zero-derivative)) (solver.hpp:621-656)
template <int DIM>
__global__ void k_smooth(Geo g, const double* __restrict__ u, double* __restrict__ out) {
  CELL_LOOP_PROLOG(g)
  double sum = 0.;
#pragma unroll
  for (int q = 0; q < 2 * DIM; ++q) {
    int d = q >> 1, o = q & 1;
    sum += face_value<DIM, K_NEUMANN0>(g, u, d, i + (d == 0 ? o : 0), j + (d == 1 ? o : 0), k + (d == 2 ? o : 0), 0);
  }
  out[c] = sum / (double)(2 * DIM);
}

// Gradient(Interpolate(u, cond)) for hg_interp_grad
template <int DIM, int KIND>
__global__ void k_interp_grad(Geo g, const double* __restrict__ u, int aux, P3 out) {
  CELL_LOOP_PROLOG(g)
#pragma unroll
  for (int d = 0; d < DIM; ++d) out.p[d][c] = cell_grad<DIM, KIND>(g, u, d, i, j, k, aux);
}

// ---------------------------------------------------------------- fluid properties
// CalcPhasesVolumeFraction + GetVolumeAveraged (hydro2d.hpp:981-997, 1235-1246)
struct PropArgs {
  int np; double density[3], viscosity[3], conductivity[3];
  const double* pd[3]; double* vf[3]; double* rho_raw; double* mu_raw; double* kc;
};
__global__ void k_volfrac(PropArgs a, long long n) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double v[3]; double sum = 0.;
  for (int p = 0; p < a.np; ++p) { v[p] = a.pd[p][c] / a.density[p]; sum += v[p]; }
  double r = 0., m = 0., kk = 0.;
  for (int p = 0; p < a.np; ++p) {
    v[p] /= sum; a.vf[p][c] = v[p];
    r += a.density[p] * v[p]; m += a.viscosity[p] * v[p]; kk += a.conductivity[p] * v[p];
  }
  a.rho_raw[c] = r; a.mu_raw[c] = m; a.kc[c] = kk;
}
// fc_force = gravity * fc_density + force (hydro2d.hpp:1308-1314)
__global__ void k_force(int dim, const double* __restrict__ rho_raw, double gx, double gy, double gz,
                        double fx, double fy, double fz, P3 out, long long n) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  out.p[0][c] = fx + gx * rho_raw[c];
  out.p[1][c] = fy + gy * rho_raw[c];
  if (dim > 2) out.p[2][c] = fz + gz * rho_raw[c];
}
// surface tension force (hydro2d.hpp:1318-1370); gs = Gradient(Interpolate(vf1, pd cond of phase 0))
template <int DIM>
__global__ void k_stforce(Geo g, CP3 gs, double sigma, P3 out) {
  CELL_LOOP_PROLOG(g)
  double fv[3] = {0., 0., 0.};
#pragma unroll
  for (int q = 0; q < 2 * DIM; ++q) {
    int qd = q >> 1, o = q & 1;
    int fi = i + (qd == 0 ? o : 0), fj = j + (qd == 1 ? o : 0), fk = k + (qd == 2 ? o : 0);
    double gv[3] = {0., 0., 0.}, nn[3] = {0., 0., 0.};
    double sq = 0.;
    for (int d = 0; d < DIM; ++d) { gv[d] = face_value<DIM, K_NEUMANN0>(g, gs.p[d], qd, fi, fj, fk, 0); sq += gv[d] * gv[d]; }
    double nrm = sqrt(sq);
    for (int d = 0; d < DIM; ++d) nn[d] = gv[d] / (nrm + 1e-6);
    double so[3] = {0., 0., 0.}; so[qd] = g.area[qd] * (o ? 1. : -1.);
    double sdn = 0.; for (int d = 0; d < DIM; ++d) sdn += so[d] * nn[d];
    for (int d = 0; d < DIM; ++d) { fv[d] += gv[d] * sdn; fv[d] -= so[d] * nrm; }
  }
  for (int d = 0; d < DIM; ++d) { fv[d] /= g.vol; out.p[d][c] = fv[d] * sigma; }
}
template <int DIM>
__global__ void k_grad_pd(Geo g, const double* __restrict__ u, const double* __restrict__ pdinit, P3 out) {
  CELL_LOOP_PROLOG(g)
#pragma unroll
  for (int d = 0; d < DIM; ++d) out.p[d][c] = cell_grad<DIM, K_PD>(g, u, d, i, j, k, 0, pdinit);
}

// ---------------------------------------------------------------- SIMPLE iteration kernels
// K_pre: restored external force (CalcExtForce, fluid.hpp:602-631) and pressure gradient
// Gradient(Interpolate(p_prev, extrapolation)) (fluid.hpp:827-829)
template <int DIM>
__global__ void __launch_bounds__(256, 8) k_pre(Geo g, CP3 force, const double* __restrict__ pprev, P3 fcr, P3 gp) {
  CELL_LOOP_PROLOG(g)
  auto body = [&](auto int_) {   // int_: interior cell, no boundary / excluded-cell tests (same arithmetic)
    constexpr bool INT = decltype(int_)::value;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double sum = 0.;
      double fm = face_value<DIM, K_NEUMANN0, INT>(g, force.p[d], d, i, j, k, 0);
      double fp = face_value<DIM, K_NEUMANN0, INT>(g, force.p[d], d, i + (d == 0), j + (d == 1), k + (d == 2), 0);
      sum += (g.area[d] * fm) * (0.5 * g.h[d]);
      sum += (g.area[d] * fp) * (0.5 * g.h[d]);
      fcr.p[d][c] = sum / g.vol;
      gp.p[d][c] = cell_grad<DIM, K_EXTRAP, INT>(g, pprev, d, i, j, k, 0);
    }
  };
  if (cell_interior<DIM>(g, i, j, k, 1)) body(std::true_type{}); else body(std::false_type{});
}

// K_velgrad: G[n*DIM+d] = d-component of Gradient(Interpolate(u_n, wall velocity)) -- used by the
// explicit viscous term (fluid.hpp:838-843) and by the deferred upwind correction (conv_diff.hpp:135)
template <int DIM>
__global__ void k_velgrad(Geo g, CP3 u, P9 G) {
  CELL_LOOP_PROLOG(g)
  auto body = [&](auto int_) {
    constexpr bool INT = decltype(int_)::value;
#pragma unroll
    for (int n = 0; n < DIM; ++n)
#pragma unroll
      for (int d = 0; d < DIM; ++d) G.p[n * DIM + d][c] = cell_grad<DIM, K_VEL, INT>(g, u.p[n], d, i, j, k, n);
  };
  if (cell_interior<DIM>(g, i, j, k, 1)) body(std::true_type{}); else body(std::false_type{});
}

struct P9c { const double* p[9]; };

// K_source: momentum source = explicit viscous term + (-grad p + restored force + surface tension)
// (fluid.hpp:835-870)
template <int DIM>
__global__ void __launch_bounds__(256, 8) k_source(Geo g, P9c G, const double* __restrict__ mu, CP3 gp, CP3 fcr, CP3 stf, int use_stf, P3 fs) {
  CELL_LOOP_PROLOG(g)
  double acc[3] = {0., 0., 0.};
  auto body = [&](auto int_) {
  constexpr bool INT = decltype(int_)::value;
#pragma unroll
  for (int n = 0; n < DIM; ++n) {
    int mi = i, mj = j, mk = k;
    int pi = i + (n == 0), pj = j + (n == 1), pk = k + (n == 2);
    double mum = face_value<DIM, K_NEUMANN0, INT>(g, mu, n, mi, mj, mk, 0);
    double mup = face_value<DIM, K_NEUMANN0, INT>(g, mu, n, pi, pj, pk, 0);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double gm = face_value<DIM, K_NEUMANN0, INT>(g, G.p[n * DIM + d], n, mi, mj, mk, 0);
      double gq = face_value<DIM, K_NEUMANN0, INT>(g, G.p[n * DIM + d], n, pi, pj, pk, 0);
      double sum = 0.;
      sum += gm * (mum * (g.area[n] * -1.));
      sum += gq * (mup * (g.area[n] * 1.));
      acc[d] += sum / g.vol;
    }
  }
  };
  if (cell_interior<DIM>(g, i, j, k, 1)) body(std::true_type{}); else body(std::false_type{});
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    double st = use_stf ? stf.p[d][c] : 0.;
    double t = ((gp.p[d][c] * (-1.) + fcr.p[d][c]) + st) + 0.;  // + u*(rho*q_vol - q_mass), sources are zero
    fs.p[d][c] = acc[d] + t;
  }
}

// K_assemble: ConvectionDiffusionScalarImplicit::MakeIteration assembly (conv_diff.hpp:149-227) for
// NCOMP scalars sharing one coefficient matrix (the dim velocity components, or the temperature).
// Writes the 7-coefficient rows and the delta-form constants in the hyperplane-major layout used by
// the ordered sweeps, and Expression::CoeffSum (fluid.hpp:876-883).
struct AsmArgs {
  const double* prev[3];   // iter_prev = field the iteration starts from
  const double* tc[3];     // time_curr
  const double* tp[3];     // time_prev
  const double* src[3];
  const double* grad[9];   // G[n*DIM+d] (only d == face direction is read)
  const double* rho;       // nullptr = 1 (heat.hpp:31)
  const double* mu;        // cell diffusion rate, faces by zero-derivative interpolation
  const double* F;         // volume flux (face field)
  double co[3];            // BDF coefficients (solver.hpp:816-861)
  double relax;
  double* A[7];            // sheared
  double* R[3];            // sheared
  double* coeffsum;        // natural, nullable
  double coeffsum_div;     // dim
  int out_sheared;         // 1: A/R are written in the sheared layout directly; 0: natural (k_shear3 follows)
};
template <int DIM, int KIND, int NCOMP>
__global__ void __launch_bounds__(256, 5) k_assemble(Geo g, AsmArgs a) {
  CELL_LOOP_PROLOG(g)
  const long long cs = a.out_sheared ? shidx(g, i, j, k) : c;
  if (cell_excl(g, i, j, k)) {   // conv_diff.hpp:222-226
#pragma unroll
    for (int t = 0; t < 7; ++t) if (DIM > 2 || (t != CZM && t != CZP)) a.A[t][cs] = (t == CD) ? 1. : 0.;
    for (int n = 0; n < NCOMP; ++n) a.R[n][cs] = 0.;
    if (a.coeffsum) { double s = 0.; for (int n = 0; n < NCOMP; ++n) s += 1.; a.coeffsum[c] = s / a.coeffsum_div; }
    return;
  }
  {
  constexpr bool INT = false;   // an interior instantiation (face_info<DIM, true>) was measured slower here (register pressure)
  const int tmap[6] = {CXM, CXP, CYM, CYP, CZM, CZP};
  const long long off[7] = {-g.sz, -g.sy, -1, 0, 1, g.sy, g.sz};
  double cdiag = 0., ddiag = 0.;
  bool have_c = false, have_d = false;
  double cn[6] = {0, 0, 0, 0, 0, 0}, dn[6] = {0, 0, 0, 0, 0, 0};
  bool present[6] = {false, false, false, false, false, false};
  double cconst[NCOMP], dconst[NCOMP];
#pragma unroll
  for (int n = 0; n < NCOMP; ++n) { cconst[n] = 0.; dconst[n] = 0.; }
#pragma unroll
  for (int q = 0; q < 2 * DIM; ++q) {
    const int d = q >> 1, o = q & 1;
    const double sgn = o ? 1. : -1.;
    const int fi = i + (d == 0 ? o : 0), fj = j + (d == 1 ? o : 0), fk = k + (d == 2 ? o : 0);
    FaceInfo f = face_info<DIM, INT>(g, d, fi, fj, fk);
    if (f.type == FT_EXCL) continue;
    const double Ff = a.F[fidx(g, d, fi, fj, fk)];
    if (f.type == FT_INNER) {
      const double muf = a.mu[f.cm] * (1. - 0.5) + a.mu[f.cp] * 0.5;
      double vm, vp;
      int up;  // 0: cm upwind, 1: cp upwind, 2: central  (solver.hpp:223-242, threshold 1e-10)
      if (Ff > 1e-10) { vm = 1.; vp = 0.; up = 0; }
      else if (Ff < -1e-10) { vm = 0.; vp = 1.; up = 1; }
      else { vm = 0.5; vp = 0.5; up = 2; }
      const double alpha = 1. / g.h[d];
      const double dm = ((-alpha) * (-muf)) * g.area[d];
      const double dp = ((alpha) * (-muf)) * g.area[d];
      double cself, cnb, dself, dnb;
      if (o) { cself = vm * Ff; cnb = vp * Ff; dself = dm; dnb = dp; }
      else { cself = vp * Ff; cnb = vm * Ff; dself = dp; dnb = dm; }
      cself *= sgn; cnb *= sgn; dself *= sgn; dnb *= sgn;
      cdiag = have_c ? cdiag + cself : cself; have_c = true;
      ddiag = have_d ? ddiag + dself : dself; have_d = true;
      cn[q] = cnb; dn[q] = dnb; present[q] = true;
#pragma unroll
      for (int n = 0; n < NCOMP; ++n) {
        double vc = 0.;
        if (up == 0) vc = -(a.grad[n * DIM + d][f.cm] * (-0.5 * g.h[d]));
        else if (up == 1) vc = -(a.grad[n * DIM + d][f.cp] * (0.5 * g.h[d]));
        cconst[n] += (vc * Ff) * sgn;
        dconst[n] += ((0. * (-muf)) * g.area[d]) * sgn;
      }
    } else {
      // boundary face (solver.hpp:258-280, 318-340)
      const double muf = a.mu[c];
      const double factor = o ? 1. : -1.;   // id == 0 for a plus face
      bool dirichlet = false;
      if (KIND == K_VEL) dirichlet = true;
      else if (KIND == K_TEMP) dirichlet = face_temp_dirichlet<DIM>(g, d, fi, fj, fk);
      if (dirichlet) {
        const double alpha = 1. / (0.5 * g.h[d]) * factor;
        const double dself = (((-alpha) * (-muf)) * g.area[d]) * sgn;
        ddiag = have_d ? ddiag + dself : dself; have_d = true;
#pragma unroll
        for (int n = 0; n < NCOMP; ++n) {
          const double val = (KIND == K_VEL) ? bc_velocity(g, f.side, n, d, fi, fj, fk) : g.heat_T;
          cconst[n] += (val * Ff) * sgn;
          dconst[n] += (((alpha * val) * (-muf)) * g.area[d]) * sgn;
        }
      } else {
        const double alpha = (0.5 * g.h[d]) * factor;
        const double cself = (1. * Ff) * sgn;
        cdiag = have_c ? cdiag + cself : cself; have_c = true;
#pragma unroll
        for (int n = 0; n < NCOMP; ++n) {
          cconst[n] += ((alpha * 0.) * Ff) * sgn;
          dconst[n] += ((0. * (-muf)) * g.area[d]) * sgn;
        }
      }
    }
  }
  const double r = a.rho ? a.rho[c] : 1.;
  const HgDiv dvol = hg_div_prepare(g.vol);   // ~20 divisions by the cell volume per cell
  double coef[7] = {0, 0, 0, 0, 0, 0, 0};
  coef[CD] = ((have_c ? hg_div(cdiag, dvol) : 0.) + a.co[2]) * r + (have_d ? hg_div(ddiag, dvol) : 0.);
#pragma unroll
  for (int q = 0; q < 2 * DIM; ++q) if (present[q]) coef[tmap[q]] = hg_div(cn[q], dvol) * r + hg_div(dn[q], dvol);
  // delta form: constant := eqn.Evaluate(prev) in ascending index order (conv_diff.hpp:218)
#pragma unroll
  for (int n = 0; n < NCOMP; ++n) {
    const double uconst = a.co[0] * a.tp[n][c] + a.co[1] * a.tc[n][c];
    double ev = ((hg_div(cconst[n], dvol) + uconst) * r + hg_div(dconst[n], dvol)) - a.src[n][c];
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      if (DIM == 2 && (t == CZM || t == CZP)) continue;
      if (t == CD) { ev += a.prev[n][c] * coef[CD]; continue; }
      const int q = (t == CXM) ? 0 : (t == CXP) ? 1 : (t == CYM) ? 2 : (t == CYP) ? 3 : (t == CZM) ? 4 : 5;
      if (present[q]) ev += a.prev[n][c + off[t]] * coef[t];
    }
    a.R[n][cs] = ev;
  }
  coef[CD] /= a.relax;   // conv_diff.hpp:221
  if (a.coeffsum) {
    double csum = 0.;
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      if (DIM == 2 && (t == CZM || t == CZP)) continue;
      if (t == CD) { csum += coef[CD]; continue; }
      const int q = (t == CXM) ? 0 : (t == CXP) ? 1 : (t == CYM) ? 2 : (t == CYP) ? 3 : (t == CZM) ? 4 : 5;
      if (present[q]) csum += coef[t];
    }
    double s = 0.;
    for (int n = 0; n < NCOMP; ++n) s += csum;   // fluid.hpp:878-882
    a.coeffsum[c] = s / a.coeffsum_div;
  }
#pragma unroll
  for (int t = 0; t < 7; ++t) if (DIM > 2 || (t != CZM && t != CZP)) a.A[t][cs] = coef[t];
  }
}

// u_curr = u_prev + corr (conv_diff.hpp:246-248); corr is in the sheared layout
template <int DIM, int NCOMP>
__global__ void k_apply_corr(Geo g, CP3 prev, CP3 X, P3 curr) {
  CELL_LOOP_PROLOG(g)
  const long long cs = shidx(g, i, j, k);
#pragma unroll
  for (int n = 0; n < NCOMP; ++n) curr.p[n][c] = prev.p[n][c] + X.p[n][cs];
}

// K_fstar: Rhie-Chow volume flux (fluid.hpp:903-940).  One thread per cell computes the cell's
// minus faces, plus the plus face where the cell is the last one in that direction.
struct FstarArgs {
  const double* us[3]; const double* gp[3]; const double* fcr[3]; const double* force[3];
  const double* pprev; const double* dc; double rc; double meshvel[3]; double* Fs;
};
template <int DIM>
DV double fstar_face(const Geo& g, const FstarArgs& a, int d, int fi, int fj, int fk) {
  FaceInfo f = face_info<DIM>(g, d, fi, fj, fk);
  const double A = g.area[d];
  const double mv = a.meshvel[d] * A;
  if (f.type == FT_INNER) {
    const double ffu = a.us[d][f.cm] * (1. - 0.5) + a.us[d][f.cp] * 0.5;
    const double vfi = ffu * A;
    const double fgp = a.gp[d][f.cm] * (1. - 0.5) + a.gp[d][f.cp] * 0.5;
    const double ffr = a.fcr[d][f.cm] * (1. - 0.5) + a.fcr[d][f.cp] * 0.5;
    const double ffe = a.force[d][f.cm] * (1. - 0.5) + a.force[d][f.cp] * 0.5;
    const double dfc = a.dc[f.cm] * (1. - 0.5) + a.dc[f.cp] * 0.5;
    const double wide = (fgp - ffr) * A;
    const double compact = (a.pprev[f.cp] - a.pprev[f.cm]) / g.h[d] * A - ffe * A;
    return (vfi + a.rc * (wide - compact) / dfc + 0) - mv;
  }
  const double ffu = (f.type == FT_BOUND) ? bc_velocity(g, f.side, d, d, fi, fj, fk) : 0.;
  return ffu * A - mv;
}
template <int DIM>
__global__ void __launch_bounds__(256, 6) k_fstar(Geo g, FstarArgs a) {
  CELL_LOOP_PROLOG(g)
  (void)c;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    a.Fs[fidx(g, d, i, j, k)] = fstar_face<DIM>(g, a, d, i, j, k);
    const int x = d == 0 ? i : (d == 1 ? j : k);
    if (x == g.n[d] - 1) {
      const int fi = i + (d == 0), fj = j + (d == 1), fk = k + (d == 2);
      a.Fs[fidx(g, d, fi, fj, fk)] = fstar_face<DIM>(g, a, d, fi, fj, fk);
    }
  }
}

// face coefficient c_f = A / (h * d_f) of the flux-correction expression (fluid.hpp:957-964)
template <int DIM>
DV double face_coeff(const Geo& g, const double* __restrict__ dc, int d, const FaceInfo& f) {
  const double dfc = dc[f.cm] * (1. - 0.5) + dc[f.cp] * 0.5;
  const double coeff = -g.area[d] / (g.h[d] * dfc);
  return -coeff;
}

// K_prhs: constants of the pressure-correction rows (fluid.hpp:972-1014) and the diagonal field in
// the sheared layout; the sweep kernels regenerate the off-diagonals A/(h d_f) from it.
// One cell of K_prhs: rp = constant, cf[d] = plus-face coefficients, dg = explicit diagonal (only when want_dg)
template <int DIM>
DV void prhs_cell(const Geo& g, const double* __restrict__ Fs, const double* __restrict__ dc, int i, int j, int k, long long c,
                  bool want_dg, double& rp, double (&cf)[3], double& dg) {
  // face coefficients of the cell's plus faces for the sweep kernel (0 when the face is not inner)
  cf[0] = cf[1] = cf[2] = 0.; dg = 0.;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    FaceInfo f = face_info<DIM>(g, d, i + (d == 0), j + (d == 1), k + (d == 2));
    if (f.type == FT_INNER) cf[d] = face_coeff<DIM>(g, dc, d, f);
  }
  if (want_dg) {
    // k_gs_tiled: explicit diagonal = ordered sum over the faces x-,x+,y-,y+,z-,z+ (fluid.hpp:979-984; absent
    // faces add 0), 1 for identity rows; the coefficients of the faces of the fixed-pressure cell are stored
    // as 0 so that the terms removed by SetKnownValue (fluid.hpp:1010) vanish from the sums
    double cm[3] = {0., 0., 0.};
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      FaceInfo f = face_info<DIM>(g, d, i, j, k);
      if (f.type == FT_INNER) cm[d] = face_coeff<DIM>(g, dc, d, f);
    }
    double diag = cm[0] + cf[0]; diag = diag + cm[1]; diag = diag + cf[1];
    if (DIM > 2) { diag = diag + cm[2]; diag = diag + cf[2]; }
    const bool ident = c == g.pfix || cell_excl(g, i, j, k);
    dg = ident ? 1. : diag;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const long long nb = cidx(g, i + (d == 0), j + (d == 1), k + (d == 2));
      if (c == g.pfix || nb == g.pfix) cf[d] = 0.;
    }
  }
  if (cell_excl(g, i, j, k)) { rp = 0.; return; }
  if (c == g.pfix) { rp = -g.pfix_value; return; }
  double cst = 0.;
  double extra = 0.; bool has_extra = false;
#pragma unroll
  for (int q = 0; q < 2 * DIM; ++q) {
    const int d = q >> 1, o = q & 1;
    const int fi = i + (d == 0 ? o : 0), fj = j + (d == 1 ? o : 0), fk = k + (d == 2 ? o : 0);
    cst += Fs[fidx(g, d, fi, fj, fk)] * (o ? 1. : -1.);
  }
  double rhs = cst + -(0. * g.vol);
  if (g.pfix != HG_NO_CELL) {
    // SetKnownValue (linear.hpp:238-250): rows coupling to the fixed cell get value*coeff added, in
    // ascending order of the fixed cell's neighbours (fluid.hpp:1002-1011 loops over all cells once)
#pragma unroll
    for (int q = 0; q < 2 * DIM; ++q) {
      const int d = q >> 1, o = q & 1;
      const int ni = i + (d == 0 ? (o ? 1 : -1) : 0), nj = j + (d == 1 ? (o ? 1 : -1) : 0), nk = k + (d == 2 ? (o ? 1 : -1) : 0);
      if (!cell_in(g, ni, nj, nk) || cidx(g, ni, nj, nk) != g.pfix) continue;
      const int fi = i + (d == 0 ? o : 0), fj = j + (d == 1 ? o : 0), fk = k + (d == 2 ? o : 0);
      FaceInfo f = face_info<DIM>(g, d, fi, fj, fk);
      if (f.type != FT_INNER) continue;
      const double cfn = face_coeff<DIM>(g, dc, d, f);
      extra = g.pfix_value * (-cfn); has_extra = true;
    }
  }
  rp = (g.pfix != HG_NO_CELL && has_extra) ? rhs + extra : rhs;
}
template <int DIM>
__global__ void __launch_bounds__(256, 6) k_prhs(Geo g, const double* __restrict__ Fs, const double* __restrict__ dc, int out_sheared,
                       double* __restrict__ RP, double* __restrict__ CX, double* __restrict__ CY, double* __restrict__ CZ,
                       double* __restrict__ DG = nullptr) {
  CELL_LOOP_PROLOG(g)
  const long long cs = out_sheared ? shidx(g, i, j, k) : c;
  double rp, cf[3], dg;
  prhs_cell<DIM>(g, Fs, dc, i, j, k, c, DG != nullptr, rp, cf, dg);
  if (DG) DG[cs] = dg;
  CX[cs] = cf[0]; CY[cs] = cf[1]; if (DIM > 2) CZ[cs] = cf[2];
  RP[cs] = rp;
}

// K_prows: explicit rows of the pressure-correction system (fluid.hpp:972-1014) for solvers that need the
// matrix (lu_relaxed): diagonal = ordered sum of the face coefficients, off-diagonals -c_f, excluded and
// fixed-pressure cells as identity rows, terms toward the fixed-pressure cell removed (SetKnownValue)
template <int DIM>
__global__ void k_prows(Geo g, const double* __restrict__ dc, int out_sheared, P7 A) {
  CELL_LOOP_PROLOG(g)
  const long long cs = out_sheared ? shidx(g, i, j, k) : c;
  double coef[7] = {0, 0, 0, 0, 0, 0, 0};
  if (cell_excl(g, i, j, k) || c == g.pfix) coef[CD] = 1.;
  else {
    const int tmap[6] = {CXM, CXP, CYM, CYP, CZM, CZP};
    double diag = 0.; bool have = false;
#pragma unroll
    for (int q = 0; q < 2 * DIM; ++q) {
      const int d = q >> 1, o = q & 1;
      FaceInfo f = face_info<DIM>(g, d, i + (d == 0 ? o : 0), j + (d == 1 ? o : 0), k + (d == 2 ? o : 0));
      if (f.type != FT_INNER) continue;
      const double cf = face_coeff<DIM>(g, dc, d, f);
      diag = have ? diag + cf : cf; have = true;
      const long long nb = o ? f.cp : f.cm;
      coef[tmap[q]] = (nb == g.pfix) ? 0. : -cf;
    }
    coef[CD] = diag;
  }
#pragma unroll
  for (int t = 0; t < 7; ++t) if (DIM > 2 || (t != CZM && t != CZP)) A.p[t][cs] = coef[t];
}

// K_pcorr: p' back to the natural layout, p_curr = p_prev + alpha_p p' (fluid.hpp:1035-1038)
template <int DIM>
__global__ void k_pcorr(Geo g, const double* __restrict__ PP, const double* __restrict__ pprev, double alpha,
                        double* __restrict__ pc, double* __restrict__ pcurr) {
  CELL_LOOP_PROLOG(g)
  const double v = PP[shidx(g, i, j, k)];
  pc[c] = v;
  pcurr[c] = pprev[c] + alpha * v;
}

// K_correct: velocity correction u += -grad p' / d_c (fluid.hpp:1040-1050) and the divergence-free
// fluxes F = F* + c_f (p'_m - p'_p) (fluid.hpp:1053-1056)
struct CorrArgs { const double* pc; const double* dc; const double* Fs; double* u[3]; double* F;
                  const double* uprev[3]; double* resid; };   // interior kernel: max_c |u_prev - u_new| (CalcDiff), nullable
template <int DIM, bool INT = false>
DV double fcorr_face(const Geo& g, const CorrArgs& a, int d, int fi, int fj, int fk) {
  const long long fx = fidx(g, d, fi, fj, fk);
  double r = a.Fs[fx];
  FaceInfo f = face_info<DIM, INT>(g, d, fi, fj, fk);
  if (f.type == FT_INNER) {
    const double cf = face_coeff<DIM>(g, a.dc, d, f);
    r += a.pc[f.cm] * cf;
    r += a.pc[f.cp] * (-cf);
  }
  return r;
}
template <int DIM>
__global__ void __launch_bounds__(256, 6) k_correct(Geo g, CorrArgs a) {
  CELL_LOOP_PROLOG(g)
  auto body = [&](auto int_) {
    constexpr bool INT = decltype(int_)::value;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const double gpc = cell_grad<DIM, K_EXTRAP, INT>(g, a.pc, d, i, j, k, 0);
      a.u[d][c] += gpc / (-a.dc[c]);
      a.F[fidx(g, d, i, j, k)] = fcorr_face<DIM, INT>(g, a, d, i, j, k);
      const int x = d == 0 ? i : (d == 1 ? j : k);
      if (!INT && x == g.n[d] - 1) {
        const int fi = i + (d == 0), fj = j + (d == 1), fk = k + (d == 2);
        a.F[fidx(g, d, fi, fj, fk)] = fcorr_face<DIM>(g, a, d, fi, fj, fk);
      }
    }
  };
  body(std::false_type{});   // the interior instantiation was measured slightly slower for this kernel
}

// K_resid: CalcDiff of vector fields, max_c ||a - b||_2 (solver.hpp:804-813)
template <int DIM>
__global__ void k_resid(CP3 ic, CP3 ip, long long n, double* out) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.;
  if (c < n) {
    double sq = 0.;
#pragma unroll
    for (int d = 0; d < DIM; ++d) { double e = ip.p[d][c] - ic.p[d][c]; sq += e * e; }
    v = sqrt(sq);
    if (!(v == v)) v = 0.;   // std::max(res, NaN) keeps res
  }
  v = warp_max(v);
  __shared__ double sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.;
    v = warp_max(v);
    if (threadIdx.x == 0) atomic_max_nonneg(out, v);
  }
}

// GetAutoTimeStep: min over non-excluded cells and their faces of |V / F|, F != 0 (fluid.hpp:1191-1206)
template <int DIM>
__global__ void k_auto_dt(Geo g, const double* __restrict__ F, double* out) {
  long long c_ = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nc_ = (long long)g.n[0] * g.n[1] * g.n[2];
  double v = 1e10;
  if (c_ < nc_) {
    int i = (int)(c_ % g.n[0]); int j = (int)((c_ / g.n[0]) % g.n[1]); int k = (int)(c_ / ((long long)g.n[0] * g.n[1]));
    if (!cell_excl(g, i, j, k)) {
#pragma unroll
      for (int q = 0; q < 2 * DIM; ++q) {
        const int d = q >> 1, o = q & 1;
        const double fl = F[fidx(g, d, i + (d == 0 ? o : 0), j + (d == 1 ? o : 0), k + (d == 2 ? o : 0))];
        if (fl != 0.) v = smin(v, fabs(g.vol / fl));
      }
    }
  }
  v = warp_min(v);
  __shared__ double sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 1e10;
    v = warp_min(v);
    if (threadIdx.x == 0) atomic_min_nonneg(out, v);
  }
}

// ---------------------------------------------------------------- advection
// AdvectionSolverMultiExplicit::MakeIteration for one phase and one stage (advection.hpp:440-480):
// Superbee face values from Gradient(Interpolate(u)) (solver.hpp:560-619), explicit update.
template <int DIM>
DV double adv_face_value(const Geo& g, const double* __restrict__ u, const double* __restrict__ pdinit,
                         const double* __restrict__ F, int d, int fi, int fj, int fk) {
  FaceInfo f = face_info<DIM>(g, d, fi, fj, fk);
  if (f.type == FT_EXCL) return 0.;
  if (f.type == FT_BOUND) return face_value<DIM, K_PD>(g, u, d, fi, fj, fk, 0, pdinit);
  const double Ff = F[fidx(g, d, fi, fj, fk)];
  const double uP = u[f.cm], uE = u[f.cp];
  const double du = uE - uP;
  if (Ff > 1e-8) {
    const double gP = cell_grad<DIM, K_PD>(g, u, d, fi - (d == 0), fj - (d == 1), fk - (d == 2), 0, pdinit);
    const double pq = -4. * (gP * (-0.5 * g.h[d])) - du;
    return uP + 0.5 * superbee(du, pq);
  } else if (Ff < -1e-8) {
    const double gE = cell_grad<DIM, K_PD>(g, u, d, fi, fj, fk, 0, pdinit);
    const double pq = 4. * (gE * (0.5 * g.h[d])) - du;
    return uE - 0.5 * superbee(du, pq);
  }
  return 0.5 * (uP + uE);
}
template <int DIM>
__global__ void __launch_bounds__(256, 8) k_advect(Geo g, const double* __restrict__ u, const double* __restrict__ pdinit,
                         const double* __restrict__ F, double dt, int num_stages, int stage, double* __restrict__ out) {
  CELL_LOOP_PROLOG(g)
  double fsum = 0.;
#pragma unroll
  for (int q = 0; q < 2 * DIM; ++q) {
    if ((q / 2) % num_stages != stage) continue;
    const int d = q >> 1, o = q & 1;
    const int fi = i + (d == 0 ? o : 0), fj = j + (d == 1 ? o : 0), fk = k + (d == 2 ? o : 0);
    const double fu = adv_face_value<DIM>(g, u, pdinit, F, d, fi, fj, fk);
    fsum += fu * F[fidx(g, d, fi, fj, fk)] * (o ? 1. : -1.);
  }
  out[c] = u[c] + -dt / g.vol * fsum;
}

// Interface sharpening (advection.hpp:479-529), once per field after the advection stages: gc = Gradient(Interpolate(u)) is in
// gcv (k_grad_pd); a face carries ff = |F nf| nf (sharp A |gf| - af (1 - af / am)) with gf = Interpolate(gc, zero derivative),
// n = gf / (|gf| + 1e-6), nf = n . normal, where |gf| >= 1; a cell gets u + dt 0 (sources) + dt sum(outward ff) / V
template <int DIM>
DV double sharp_face(const Geo& g, const double* __restrict__ u, const double* __restrict__ pdinit, CP3 gcv,
                     const double* __restrict__ F, double sharp, double am, int d, int fi, int fj, int fk) {
  double n[3] = {0., 0., 0.}, sq = 0.;
#pragma unroll
  for (int c = 0; c < DIM; ++c) { n[c] = face_value<DIM, K_NEUMANN0>(g, gcv.p[c], d, fi, fj, fk, 0); sq += n[c] * n[c]; }
  const double nrm = sqrt(sq);
  if (nrm < 1.) return 0.;
  double nf = 0.;
#pragma unroll
  for (int c = 0; c < DIM; ++c) { n[c] /= (nrm + 1e-6); nf += n[c] * (c == d ? 1. : 0.); }
  const double af = face_value<DIM, K_PD>(g, u, d, fi, fj, fk, 0, pdinit);
  const double uf = F[fidx(g, d, fi, fj, fk)];
  const double epsh = sharp * g.area[d];
  return fabs(uf * nf) * nf * (epsh * nrm - af * (1. - af / am));
}
template <int DIM>
__global__ void __launch_bounds__(256, 4) k_sharpen(Geo g, const double* __restrict__ u, const double* __restrict__ pdinit, CP3 gcv,
                                                   const double* __restrict__ F, double dt, double sharp, double am, double* __restrict__ out) {
  CELL_LOOP_PROLOG(g)
  double sh = 0.;
#pragma unroll
  for (int q = 0; q < 2 * DIM; ++q) {
    const int d = q >> 1, o = q & 1;
    const int fi = i + (d == 0 ? o : 0), fj = j + (d == 1 ? o : 0), fk = k + (d == 2 ? o : 0);
    sh += (o ? 1. : -1.) * sharp_face<DIM>(g, u, pdinit, gcv, F, sharp, am, d, fi, fj, fk);
  }
  double v = u[c];
  v += dt * 0.;
  v += dt * sh / g.vol;
  out[c] = v;
}

// ---------------------------------------------------------------- SIMPLER: second pressure solve (fluid.hpp:1060-1155)
// fc_evaluated = momentum equations (rows A, constants R of this iteration, hyperplane-major) applied to the velocity change of
// the iteration + restored force - pressure gradient (:1073-1094)
struct SimplerArgs {
  const double* A[7]; const double* R[3]; const double* uc[3]; const double* up[3]; const double* fcr[3]; const double* gp[3];
  const double* force[3]; const double* dc; const double* F; const double* p; double rc;
  double* fev[3];
  double* RP; double* CO; Co5Fwd co5; int out_mode;   // constants: 1 = hyperplane-major RP, 2 = array 0 of the CO5 rows
};
template <int DIM>
__global__ void __launch_bounds__(256, 4) k_simpler_eval(Geo g, SimplerArgs a) {
  CELL_LOOP_PROLOG(g)
  const long long cs = shidx(g, i, j, k);
  const long long off[7] = {-g.sz, -g.sy, -1, 0, 1, g.sy, g.sz};
  const int di[7] = {0, 0, -1, 0, 1, 0, 0}, dj[7] = {0, -1, 0, 0, 0, 1, 0}, dk[7] = {-1, 0, 0, 0, 0, 0, 1};
#pragma unroll
  for (int n = 0; n < DIM; ++n) {
    double ev = a.R[n][cs];
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      if (DIM == 2 && (t == CZM || t == CZP)) continue;
      const double coef = a.A[t][cs];
      if (t != CD && !(cell_in(g, i + di[t], j + dj[t], k + dk[t]) && coef != 0.)) continue;   // terms of the expression only
      const long long nb = c + off[t];
      ev += (a.uc[n][nb] - a.up[n][nb]) * coef;
    }
    a.fev[n][c] = ev + (a.fcr[n][c] - a.gp[n][c]);
  }
}
// ff_rhs on the cell's faces (:1099-1112), constant = -sum (:1114-1121), fixed-pressure row (:1124-1141), then
// constant := Evaluate(pressure) over the rows of the pressure-correction system (:1143-1146)
template <int DIM>
__global__ void __launch_bounds__(256, 4) k_simpler_rhs(Geo g, SimplerArgs a) {
  CELL_LOOP_PROLOG(g)
  double cfq[6] = {0., 0., 0., 0., 0., 0.};
  bool innerq[6] = {false, false, false, false, false, false};
  long long nbq[6] = {0, 0, 0, 0, 0, 0};
  double sum = 0.;
#pragma unroll
  for (int q = 0; q < 2 * DIM; ++q) {
    const int d = q >> 1, o = q & 1;
    const int fi = i + (d == 0 ? o : 0), fj = j + (d == 1 ? o : 0), fk = k + (d == 2 ? o : 0);
    const FaceInfo f = face_info<DIM>(g, d, fi, fj, fk);
    double frhs = 0.;
    if (f.type == FT_INNER) {
      double dot1 = 0., dot2 = 0.;
#pragma unroll
      for (int cc = 0; cc < DIM; ++cc) {
        const double S = cc == d ? g.area[d] : 0.;
        const double fe = a.fev[cc][f.cm] * (1. - 0.5) + a.fev[cc][f.cp] * 0.5;
        const double ff = a.force[cc][f.cm] * (1. - 0.5) + a.force[cc][f.cp] * 0.5;
        const double fu = a.uc[cc][f.cm] * (1. - 0.5) + a.uc[cc][f.cp] * 0.5;
        dot1 += (fe - ff) * S;
        dot2 += fu * S;
      }
      const double dfc = a.dc[f.cm] * (1. - 0.5) + a.dc[f.cp] * 0.5;
      frhs = dot1 / dfc + (a.F[fidx(g, d, fi, fj, fk)] - dot2) / a.rc;
      cfq[q] = face_coeff<DIM>(g, a.dc, d, f);
      innerq[q] = true;
      nbq[q] = o ? f.cp : f.cm;
    }
    sum += frhs * (o ? 1. : -1.);
  }
  double cst = -sum;
  const bool ident = c == g.pfix || cell_excl(g, i, j, k);
  if (c == g.pfix) cst = -g.pfix_value;
  if (ident) cst += a.p[c] * 1.;
  else {
    double diag = cfq[0] + cfq[1]; diag = diag + cfq[2]; diag = diag + cfq[3];
    if (DIM > 2) { diag = diag + cfq[4]; diag = diag + cfq[5]; }
    // ascending index: z-, y-, x-, diagonal, x+, y+, z+ (terms towards the fixed-pressure cell were removed, fluid.hpp:1010)
    const int order[7] = {4, 2, 0, -1, 1, 3, 5};
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      const int q = order[t];
      if (q < 0) { cst += a.p[c] * diag; continue; }
      if (q >= 2 * DIM || !innerq[q] || nbq[q] == g.pfix) continue;
      cst += a.p[nbq[q]] * (-cfq[q]);
    }
  }
  if (a.out_mode == 2) a.CO[co5fwd_index(a.co5, i, j, k)] = cst;
  else a.RP[shidx(g, i, j, k)] = cst;
}
// p_curr += p'' (:1150-1152), p'' hyperplane-major
template <int DIM>
__global__ void k_simpler_padd(Geo g, const double* __restrict__ PP, double* __restrict__ pcurr) {
  CELL_LOOP_PROLOG(g)
  pcurr[c] += PP[shidx(g, i, j, k)];
}

// ---------------------------------------------------------------- phase slip (CalcPhaseVelocitySlip, hydro2d.hpp:1030-1122)
// velocity_is_carrier 0: Stokes settling velocity of every enabled phase relative to the carrier, made relative to the mixture
// (cell vectors slipv[phase][component]); slip flux on the inner faces, corrected to zero volume-weighted average.
struct SlipArgs {
  int np; int enable[3]; double radius[3], density[3], gravity[3];
  const double* vf[3]; const double* rho_raw; const double* mu;
  double* slipv[3][3];       // cell
  double* fslip[3];          // face
};
template <int DIM>
__global__ void __launch_bounds__(256) k_slip_cell(Geo g, SlipArgs a) {
  CELL_LOOP_PROLOG(g)
  (void)i; (void)j; (void)k;
  double rel[3][3];
#pragma unroll
  for (int ph = 0; ph < 3; ++ph)
#pragma unroll
    for (int d = 0; d < 3; ++d) rel[ph][d] = 0.;
  const double md = a.rho_raw[c], mv = a.mu[c];
#pragma unroll
  for (int ph = 0; ph < 3; ++ph) {
    if (ph < a.np && a.enable[ph]) {
      const double pc = a.vf[ph][c];
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        rel[ph][d] = (pc < 0.01 || pc > 0.99) ? 0. : a.gravity[d] * (a.density[ph] - md) * (a.radius[ph] * a.radius[ph]) / (18. * mv);
    }
  }
  double carrier[3] = {0., 0., 0.};
#pragma unroll
  for (int ph = 0; ph < 3; ++ph) if (ph < a.np) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) carrier[d] += rel[ph][d] * a.vf[ph][c];
  }
#pragma unroll
  for (int ph = 0; ph < 3; ++ph) if (ph < a.np) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) a.slipv[ph][d][c] = rel[ph][d] - carrier[d];
  }
}
template <int DIM>
DV void slip_face(const Geo& g, const SlipArgs& a, int d, int fi, int fj, int fk) {
  const FaceInfo f = face_info<DIM>(g, d, fi, fj, fk);
  const long long fx = fidx(g, d, fi, fj, fk);
  if (f.type != FT_INNER) {
    for (int ph = 0; ph < a.np; ++ph) a.fslip[ph][fx] = 0.;
    return;
  }
  double fs[3] = {0., 0., 0.};
#pragma unroll
  for (int ph = 0; ph < 3; ++ph) if (ph < a.np) {
    double dot = 0.;
#pragma unroll
    for (int c = 0; c < DIM; ++c) dot += (a.slipv[ph][c][f.cm] * (1. - 0.5) + a.slipv[ph][c][f.cp] * 0.5) * (c == d ? g.area[d] : 0.);
    fs[ph] = dot;
  }
  double aver = 0.;
#pragma unroll
  for (int ph = 0; ph < 3; ++ph) if (ph < a.np) aver += (a.vf[ph][f.cm] * (1. - 0.5) + a.vf[ph][f.cp] * 0.5) * fs[ph];
#pragma unroll
  for (int ph = 0; ph < 3; ++ph) if (ph < a.np) a.fslip[ph][fx] = fs[ph] - aver;
}
template <int DIM>
__global__ void __launch_bounds__(256) k_slip_face(Geo g, SlipArgs a) {
  CELL_LOOP_PROLOG(g)
  (void)c;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    slip_face<DIM>(g, a, d, i, j, k);
    const int x = d == 0 ? i : (d == 1 ? j : k);
    if (x == g.n[d] - 1) slip_face<DIM>(g, a, d, i + (d == 0), j + (d == 1), k + (d == 2));
  }
}
// flux of a phase's advection = mixture flux + slip flux (advection.hpp:449-454)
__global__ void k_face_add(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, long long n) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = a[t] + b[t];
}

// ---------------------------------------------------------------- outlet conditions (fluid.hpp:542-600)
// Pass A (one thread per boundary face of the domain sides): every outlet face takes the velocity of its cell; the face's
// contributions to the outlet flux, the outlet area and the inlet flux are stored at the face's position in the reference's
// summation order (ascending face index: direction, then k, j, i).  Pass B adds them in exactly that order (one thread adds,
// the block stages the terms in shared memory), so the balance has the reference's bits; directions without an inlet or
// outlet side are skipped.  Pass C applies the additive normal correction.
struct OutletArgs { const double* u[3]; double* outvel; long long outplane; double* part; double* corr; };
template <int DIM>
DV bool outlet_face(const Geo& g, long long t, int& side, int& d, int& fi, int& fj, int& fk, long long& pidx) {
  // thread -> (side, face of the side): sides 0 .. 2 DIM - 1, g.outplane slots each, of which na * nb are faces
  side = (int)(t / g.outplane); pidx = t - (long long)side * g.outplane;
  if (side >= 2 * DIM) return false;
  d = side >> 1;
  const int na = d == 0 ? g.n[1] : g.n[0], nb = d == 2 ? g.n[1] : g.n[2];   // (n[2] = 1 in 2-D)
  if (pidx >= (long long)na * nb) return false;
  const int a = (int)(pidx % na), b = (int)(pidx / na);
  const int x = (side & 1) ? g.n[d] : 0;
  fi = d == 0 ? x : a; fj = d == 1 ? x : (d == 0 ? a : b); fk = d == 2 ? x : (d == 0 || d == 1 ? b : 0);
  return true;
}
// position of a domain-side face in the face-index order of its direction, and the first position of a direction
HD long long outlet_seq(const Geo& g, int d, int hi, int fi, int fj, int fk) {
  if (d == 0) return 2LL * (fj + (long long)g.n[1] * fk) + hi;                 // i + (nx+1) (j + ny k): i = 0, nx alternate
  if (d == 1) return fi + (long long)g.n[0] * (hi + 2LL * fk);                 // i + nx (j + (ny+1) k): rows j = 0, ny per k
  return fi + (long long)g.n[0] * (fj + (long long)g.n[1] * hi);               // i + nx (j + ny k): planes k = 0, nz
}
HD long long outlet_base(const Geo& g, int d) {
  const long long c0 = 2LL * g.n[1] * g.n[2], c1 = 2LL * g.n[0] * g.n[2];
  return d == 0 ? 0 : (d == 1 ? c0 : c0 + c1);
}
template <int DIM>
__global__ void __launch_bounds__(256) k_outlet_collect(Geo g, OutletArgs a) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int side, d, fi, fj, fk; long long pidx;
  if (!outlet_face<DIM>(g, t, side, d, fi, fj, fk, pidx)) return;
  double v[3] = {0., 0., 0.};   // outlet flux, outlet area, inlet flux
  const FaceInfo f = face_info<DIM>(g, d, fi, fj, fk);
  if (f.type == FT_BOUND && f.side < 6) {
    const long long cc = f.id == 0 ? f.cm : f.cp;
    if (g.bckind[side] == 2) {
      const double factor = f.id == 0 ? 1. : -1.;
      double dot = 0.;
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
        const double uc = a.u[c][cc];
        a.outvel[((long long)side * 3 + c) * a.outplane + pidx] = uc;
        dot += uc * (c == d ? g.area[d] : 0.);
      }
      v[0] = dot * factor; v[1] = g.area[d];
    } else if (g.bckind[side] == 1) {
      const double factor = f.id == 0 ? -1. : 1.;
      double dot = 0.;
#pragma unroll
      for (int c = 0; c < DIM; ++c) dot += g.bcvel[side][c] * (c == d ? g.area[d] : 0.);
      v[2] = dot * factor;
    }
  }
  const long long q = outlet_base(g, d) + outlet_seq(g, d, side & 1, fi, fj, fk);
  a.part[3 * q] = v[0]; a.part[3 * q + 1] = v[1]; a.part[3 * q + 2] = v[2];
}
template <int DIM>
__global__ void __launch_bounds__(256) k_outlet_correction(Geo g, const double* __restrict__ part, double* corr) {
  __shared__ double sm[3 * 256];
  double s[3] = {0., 0., 0.};
  for (int d = 0; d < DIM; ++d) {
    if (g.bckind[2 * d] == 0 && g.bckind[2 * d + 1] == 0) continue;   // walls on both sides: no term
    const long long base = outlet_base(g, d), cnt = outlet_base(g, d + 1 < 3 ? d + 1 : 2) - base;
    const long long n = d == 2 ? 2LL * g.n[0] * g.n[1] : cnt;
    for (long long q0 = 0; q0 < n; q0 += 256) {
      const int m = (int)(n - q0 < 256 ? n - q0 : 256);
      for (int e = threadIdx.x; e < 3 * m; e += 256) sm[e] = part[3 * (base + q0) + e];
      __syncthreads();
      if (threadIdx.x == 0) for (int e = 0; e < m; ++e) { s[0] += sm[3 * e]; s[1] += sm[3 * e + 1]; s[2] += sm[3 * e + 2]; }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) *corr = (s[2] - s[0]) / s[1];   // (inlet - outlet) / outlet area
}
template <int DIM>
__global__ void __launch_bounds__(256) k_outlet_apply(Geo g, OutletArgs a) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int side, d, fi, fj, fk; long long pidx;
  if (!outlet_face<DIM>(g, t, side, d, fi, fj, fk, pidx) || g.bckind[side] != 2) return;
  const FaceInfo f = face_info<DIM>(g, d, fi, fj, fk);
  if (f.type != FT_BOUND || f.side >= 6) return;
  const double factor = f.id == 0 ? 1. : -1.;
  const double corr = *a.corr;
#pragma unroll
  for (int c = 0; c < DIM; ++c) {
    double* p = &a.outvel[((long long)side * 3 + c) * a.outplane + pidx];
    *p = *p + ((c == d ? g.area[d] / g.area[d] : 0. / g.area[d]) * corr) * factor;
  }
}

// ---------------------------------------------------------------- statistics (CalcStat, hydro2d.hpp:1432-1466)
// out[ph*12 + {0 volume, 1..3 centre sums, 4..6 velocity sums}] (atomicAdd), out[ph*12+7] pd_min, +8 pd_max
struct StatArgs { int np; const double* vf[3]; const double* pd[3]; const double* u[3]; double* out; };
constexpr int STAT_CPT = 8;   // cells per thread: partial sums in registers, one block reduction at the end
template <int DIM>
__global__ void __launch_bounds__(256, 4) k_stat(Geo g, StatArgs a) {
  const long long nc_ = (long long)g.n[0] * g.n[1] * g.n[2];
  const long long base = (long long)blockIdx.x * (256 * STAT_CPT) + threadIdx.x;
  double acc[3][9];
#pragma unroll
  for (int ph = 0; ph < 3; ++ph) {
#pragma unroll
    for (int q = 0; q < 7; ++q) acc[ph][q] = 0.;
    acc[ph][7] = 1e300; acc[ph][8] = -1e300;
  }
#pragma unroll 2
  for (int r = 0; r < STAT_CPT; ++r) {
    const long long c_ = base + 256LL * r;
    if (c_ >= nc_) break;
    int i, j, k;
    if (nc_ < (1LL << 31)) {
      const unsigned c32 = (unsigned)c_, nx = (unsigned)g.n[0], nxy = nx * (unsigned)g.n[1];
      const unsigned k_ = c32 / nxy, r_ = c32 - k_ * nxy, j_ = r_ / nx;
      k = (int)k_; j = (int)j_; i = (int)(r_ - j_ * nx);
    } else { i = (int)(c_ % g.n[0]); j = (int)((c_ / g.n[0]) % g.n[1]); k = (int)(c_ / ((long long)g.n[0] * g.n[1])); }
    double x[3]; cell_center(g, i, j, k, x);
    double uv[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) uv[d] = d < DIM ? a.u[d][c_] : 0.;
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
      if (ph < a.np) {
        const double w = a.vf[ph][c_] * g.vol;
        acc[ph][0] += w;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { acc[ph][1 + d] += x[d] * w; acc[ph][4 + d] += uv[d] * w; }
        const double pd = a.pd[ph][c_];
        acc[ph][7] = pd < acc[ph][7] ? pd : acc[ph][7];
        acc[ph][8] = acc[ph][8] < pd ? pd : acc[ph][8];
      }
    }
  }
  __shared__ double sm[27][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int ph = 0; ph < 3; ++ph) {
    if (ph < a.np) {
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        double v = acc[ph][q];
        if (q < 7) v = warp_sum(v); else if (q == 7) v = warp_min(v); else v = warp_max(v);
        if (lane == 0) sm[ph * 9 + q][w] = v;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 27 && threadIdx.x / 9 < a.np) {
    const int ph = threadIdx.x / 9, q = threadIdx.x % 9;
    double v = sm[threadIdx.x][0];
    for (int t = 1; t < 8; ++t) { const double o = sm[threadIdx.x][t]; v = q < 7 ? v + o : (q == 7 ? (o < v ? o : v) : (v < o ? o : v)); }
    if (q < 7) atomicAdd(&a.out[ph * 12 + q], v);
    else {   // signed doubles: CAS loop
      unsigned long long* ad = (unsigned long long*)&a.out[ph * 12 + q]; unsigned long long old = *ad, assumed;
      do {
        assumed = old;
        const double cur = __longlong_as_double((long long)assumed);
        if (q == 7 ? !(v < cur) : !(cur < v)) break;
        old = atomicCAS(ad, assumed, (unsigned long long)__double_as_longlong(v));
      } while (assumed != old);
    }
  }
}

// ---------------------------------------------------------------- initial fields (hydro2d.hpp:310-368, 506-550)
struct InitArgs {
  double v0[3]; int pois; int sin_on; double sin_n[3], sin_lambda, sin_phase;
  double A1[3], B1[3], A2[3], B2[3], IC[3], IR, IC2[3], IR2;
  int np; double density[3], ivf[3];
  double* u[3]; double* pd[3];
};
DV bool rect_inside(const double* lb, const double* rt, const double* x, int dim) {
  for (int d = 0; d < dim; ++d) if (x[d] < lb[d] || rt[d] < x[d]) return false;
  return true;
}
template <int DIM>
__global__ void k_init_fields(Geo g, InitArgs a) {
  CELL_LOOP_PROLOG(g)
  double x[3]; cell_center(g, i, j, k, x);
  double v[3] = {a.v0[0], a.v0[1], DIM > 2 ? a.v0[2] : 0.};
  if (a.pois) v[0] = x[1] * (1. - x[1]) * 4. * a.v0[0];
  if (a.sin_on) {
    const double pi = atan(1.) * 4.;
    double kd = 0.;
    for (int d = 0; d < DIM; ++d) kd += (a.sin_n[d] * (2. * pi / a.sin_lambda)) * x[d];
    const double sn = sin(kd - a.sin_phase);
    for (int d = 0; d < DIM; ++d) v[d] *= sn;
  }
  for (int d = 0; d < DIM; ++d) a.u[d][c] = v[d];
  double pdv[3];
  for (int p = 0; p < a.np; ++p) pdv[p] = a.density[p] * a.ivf[p];
  double d1 = 0., d2 = 0.;
  for (int d = 0; d < DIM; ++d) { double e = a.IC[d] - x[d]; d1 += e * e; e = a.IC2[d] - x[d]; d2 += e * e; }
  if (rect_inside(a.A2, a.B2, x, DIM)) { if (a.np > 2) pdv[2] = a.density[2]; }
  else if (rect_inside(a.A1, a.B1, x, DIM)) { if (a.np > 1) pdv[1] = a.density[1]; }
  else if (sqrt(d1) < a.IR) { if (a.np > 1) pdv[1] = a.density[1]; }
  else if (sqrt(d2) < a.IR2) { if (a.np > 1) pdv[1] = a.density[1]; }
  for (int p = 0; p < a.np; ++p) a.pd[p][c] = pdv[p];
}
// partial density of phase 0 compensates the others (hydro2d.hpp:542-550)
__global__ void k_pd0(int np, double d0, double d1, double d2, const double* pd1, const double* pd2, double* pd0, long long n) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double vs = 0.;
  if (np > 1) vs += pd1[c] / d1;
  if (np > 2) vs += pd2[c] / d2;
  pd0[c] = (1. - vs) * d0;
}
// initial volume flux (FluidSimple ctor, fluid.hpp:770-785): F = Interpolate(u, wall).S - meshvel.S
template <int DIM>
__global__ void k_init_flux(Geo g, CP3 u, double mv0, double mv1, double mv2, double* __restrict__ F) {
  CELL_LOOP_PROLOG(g)
  (void)c;
  const double mv[3] = {mv0, mv1, mv2};
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    F[fidx(g, d, i, j, k)] = face_value<DIM, K_VEL>(g, u.p[d], d, i, j, k, d) * g.area[d] - mv[d] * g.area[d];
    const int x = d == 0 ? i : (d == 1 ? j : k);
    if (x == g.n[d] - 1) {
      const int fi = i + (d == 0), fj = j + (d == 1), fk = k + (d == 2);
      F[fidx(g, d, fi, fj, fk)] = face_value<DIM, K_VEL>(g, u.p[d], d, fi, fj, fk, d) * g.area[d] - mv[d] * g.area[d];
    }
  }
}
__global__ void k_excl_mask(Geo g, double bx0, double bx1, double bx2, double by0, double by1, double by2,
                            unsigned char* excl, int* any) {
  // covers the halo planes too (global coordinates make them consistent without an exchange)
  long long c_ = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nxy = (long long)g.n[0] * g.n[1];
  if (c_ >= nxy * (g.n[2] + g.zlo + g.zhi)) return;
  const int i = (int)(c_ % g.n[0]), j = (int)((c_ / g.n[0]) % g.n[1]), k = (int)(c_ / nxy) - g.zlo;
  const long long c = cidx(g, i, j, k);
  double x[3]; cell_center(g, i, j, k, x);
  const double lb[3] = {bx0, bx1, bx2}, rt[3] = {by0, by1, by2};
  const bool in = rect_inside(lb, rt, x, g.dim);
  excl[c] = in ? 1 : 0;
  if (in) *any = 1;
}
__global__ void k_mask_to_double(const unsigned char* m, double* out, long long n) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) out[c] = m ? (m[c] ? 1. : 0.) : 0.;
}
// layout conversion natural <-> sheared for the kernel-level linear-solve entry
template <int DIM>
__global__ void k_to_sheared(Geo g, const double* __restrict__ in, double* __restrict__ out) {
  CELL_LOOP_PROLOG(g)
  out[shidx(g, i, j, k)] = in[c];
}
template <int DIM>
__global__ void k_from_sheared(Geo g, const double* __restrict__ in, double* __restrict__ out) {
  CELL_LOOP_PROLOG(g)
  out[c] = in[shidx(g, i, j, k)];
}

// natural -> sheared layout for up to 10 arrays through a 32x32 (i,k) shared-memory tile at fixed j: a
// diagonal i+k = const of the tile is a contiguous run of the sheared array, written by one warp.
struct ShearArgs { const double* in[10]; double* out[10]; int narr; };
__global__ void __launch_bounds__(256) k_shear3(Geo g, ShearArgs a) {
  __shared__ double tile[32][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i0 = blockIdx.x * 32, j = blockIdx.y, k0 = blockIdx.z * 32;
  for (int q = 0; q < a.narr; ++q) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kl = ty + 8 * r;
      if (i0 + tx < g.n[0] && k0 + kl < g.n[2]) tile[kl][tx] = a.in[q][cidx(g, i0 + tx, j, k0 + kl)];
    }
    __syncthreads();
    for (int d = ty; d < 63; d += 8) {
      const int il = (d > 31 ? d - 31 : 0) + tx;
      const int kl = d - il;
      if (il <= 31 && kl >= 0 && kl <= 31 && i0 + il < g.n[0] && k0 + kl < g.n[2])
        a.out[q][shidx(g, i0 + il, j, k0 + kl)] = tile[kl][il];
    }
    __syncthreads();
  }
}
