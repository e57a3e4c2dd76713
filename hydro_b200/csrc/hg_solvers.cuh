// hg_solvers.cuh -- the reference's order-dependent linear solvers on the GPU.
//
// Lexicographic Gauss-Seidel / SOR (linear.hpp:685-715) and the one-forward-one-backward "lu"
// sweep (linear.hpp:533-566) read already-updated values of the three lower neighbours
// (i-1,j,k), (i,j-1,k), (i,j,k-1).  Every cell of a hyperplane i+j+k = k' therefore only depends on
// plane k'-1 (new values) and plane k'+1 (old values): processing planes in ascending k' reproduces
// the serial result exactly, whatever the order inside a plane.  Sweep s+1 may process plane k' once
// sweep s has finished plane k'+1, so all sweeps are pipelined: global step T handles plane
// k' = T - 2 s of every sweep s in flight.
//
// Data layout: hyperplane-major ("sheared"), P[k'][j][i], so a plane is contiguous and accesses are
// coalesced along i.  One persistent cooperative kernel per solve; a grid barrier separates steps.
#pragma once
#include <cooperative_groups.h>
#include "hg_device.cuh"

namespace cg = cooperative_groups;

constexpr int SOLVER_BX = 64;   // threads along i
constexpr int SOLVER_BY = 4;    // rows (j) per tile
constexpr int SOLVER_THREADS = SOLVER_BX * SOLVER_BY;

struct PlaneTiling {
  int rg;   // row groups per plane = ceil(ny / BY)
  int ch;   // i-chunks per row group
  int tiles;
};
__host__ __device__ inline PlaneTiling plane_tiling(const Geo& g) {
  PlaneTiling t;
  t.rg = (g.n[1] + SOLVER_BY - 1) / SOLVER_BY;
  int width = g.n[2] + SOLVER_BY - 1;              // i-extent of the valid band over BY rows
  if (width > g.n[0]) width = g.n[0];
  t.ch = (width + SOLVER_BX - 1) / SOLVER_BX;
  t.tiles = t.rg * t.ch;
  return t;
}
// cell handled by this thread for tile `local` of plane kp; returns false if none
DV bool tile_cell(const Geo& g, const PlaneTiling& pt, int kp, int local, int& i, int& j, int& k) {
  const int rgi = local / pt.ch, chi = local % pt.ch;
  const int j0 = rgi * SOLVER_BY;
  int ilo = kp - (j0 + SOLVER_BY - 1) - (g.n[2] - 1);
  if (ilo < 0) ilo = 0;
  j = j0 + (threadIdx.x / SOLVER_BX);
  i = ilo + chi * SOLVER_BX + (threadIdx.x % SOLVER_BX);
  k = kp - i - j;
  return j < g.n[1] && i < g.n[0] && k >= 0 && k < g.n[2];
}

// ---------------------------------------------------------------- pressure: Gauss-Seidel / SOR
struct GsArgs {
  const double* D;     // diagonal field d_c (fc_diag_coeff_), sheared
  const double* RP;    // row constants, sheared
  double* PP;          // solution, sheared; must be zero on entry for sweep 0 (linear.hpp:686)
  double* diff;        // per-sweep max |value - x| (linear.hpp:707), indexed by absolute sweep number
  int s_begin, s_end;  // sweeps [s_begin, s_end) are run by this launch
  double omega;
};

// off-diagonal coupling c_f = A/(h d_f) toward neighbour (ni,nj,nk) through face direction d; 0 when
// the face is not an inner face or either cell is the fixed-pressure cell (fluid.hpp:997-1014)
template <int DIM>
DV double gs_coeff(const Geo& g, const double* __restrict__ D, double d0, int d, int ni, int nj, int nk, bool& inner) {
  inner = cell_ok(g, ni, nj, nk);
  if (!inner) return 0.;
  const double dn = D[shidx(g, ni, nj, nk)];
  // d_f = d[cm]*0.5 + d[cp]*0.5 (commutative), coeff = -A/(h d_f), c_f = -coeff
  const double dfc = dn * (1. - 0.5) + d0 * 0.5;
  const double coeff = -g.area[d] / (g.h[d] * dfc);
  return -coeff;
}

template <int DIM>
__global__ void __launch_bounds__(SOLVER_THREADS) k_gs_persistent(Geo g, GsArgs a) {
  cg::grid_group grid = cg::this_grid();
  const PlaneTiling pt = plane_tiling(g);
  const int S = a.s_end - a.s_begin;
  const int Tmax = (g.np - 1) + 2 * (S - 1);
  __shared__ double sm[SOLVER_THREADS / 32];
  for (int T = 0; T <= Tmax; ++T) {
    // sweeps (relative) with 0 <= T - 2 s <= np-1
    int smin = (T - (g.np - 1) + 1) / 2; if (T - (g.np - 1) <= 0) smin = 0;
    int smax = T / 2; if (smax > S - 1) smax = S - 1;
    const int nact = smax - smin + 1;
    const long long ntiles = (long long)nact * pt.tiles;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int s = smin + (int)(tile / pt.tiles);
      const int local = (int)(tile % pt.tiles);
      const int kp = T - 2 * s;
      int i, j, k;
      double ac = 0.;
      if (tile_cell(g, pt, kp, local, i, j, k)) {
        const long long cs = shidx(g, i, j, k);
        const double rhs = a.RP[cs];
        const double xold = __ldcg(&a.PP[cs]);
        double diag, sum = 0.;
        const bool ident = cell_excl(g, i, j, k) || cidx(g, i, j, k) == g.pfix;
        if (ident) { diag = 1.; }
        else {
          const double d0 = a.D[cs];
          bool in_xm, in_xp, in_ym, in_yp, in_zm = false, in_zp = false;
          const double cxm = gs_coeff<DIM>(g, a.D, d0, 0, i - 1, j, k, in_xm);
          const double cxp = gs_coeff<DIM>(g, a.D, d0, 0, i + 1, j, k, in_xp);
          const double cym = gs_coeff<DIM>(g, a.D, d0, 1, i, j - 1, k, in_ym);
          const double cyp = gs_coeff<DIM>(g, a.D, d0, 1, i, j + 1, k, in_yp);
          double czm = 0., czp = 0.;
          if (DIM > 2) { czm = gs_coeff<DIM>(g, a.D, d0, 2, i, j, k - 1, in_zm); czp = gs_coeff<DIM>(g, a.D, d0, 2, i, j, k + 1, in_zp); }
          // diagonal: contributions merged in face order x-,x+,y-,y+,z-,z+ (fluid.hpp:979-984)
          bool have = false; diag = 0.;
          if (in_xm) { diag = have ? diag + cxm : cxm; have = true; }
          if (in_xp) { diag = have ? diag + cxp : cxp; have = true; }
          if (in_ym) { diag = have ? diag + cym : cym; have = true; }
          if (in_yp) { diag = have ? diag + cyp : cyp; have = true; }
          if (DIM > 2) {
            if (in_zm) { diag = have ? diag + czm : czm; have = true; }
            if (in_zp) { diag = have ? diag + czp : czp; have = true; }
          }
          // off-diagonal terms in ascending index order z-,y-,x-,x+,y+,z+ (linear.hpp:694-701);
          // terms toward the fixed-pressure cell were removed by SetKnownValue
          const long long pf = g.pfix;
          if (DIM > 2 && in_zm && cidx(g, i, j, k - 1) != pf) sum += (-czm) * __ldcg(&a.PP[shidx(g, i, j, k - 1)]);
          if (in_ym && cidx(g, i, j - 1, k) != pf) sum += (-cym) * __ldcg(&a.PP[shidx(g, i, j - 1, k)]);
          if (in_xm && cidx(g, i - 1, j, k) != pf) sum += (-cxm) * __ldcg(&a.PP[shidx(g, i - 1, j, k)]);
          if (in_xp && cidx(g, i + 1, j, k) != pf) sum += (-cxp) * __ldcg(&a.PP[shidx(g, i + 1, j, k)]);
          if (in_yp && cidx(g, i, j + 1, k) != pf) sum += (-cyp) * __ldcg(&a.PP[shidx(g, i, j + 1, k)]);
          if (DIM > 2 && in_zp && cidx(g, i, j, k + 1) != pf) sum += (-czp) * __ldcg(&a.PP[shidx(g, i, j, k + 1)]);
        }
        const double value = -(rhs + sum) / diag;
        const double corr = value - xold;
        a.PP[cs] = xold + corr * a.omega;
        ac = fabs(corr);
        if (!(ac == ac)) ac = 0.;
      }
      // per-sweep max-norm (linear.hpp:707)
      ac = warp_max(ac);
      if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = ac;
      __syncthreads();
      if (threadIdx.x < 32) {
        double v = threadIdx.x < SOLVER_THREADS / 32 ? sm[threadIdx.x] : 0.;
        v = warp_max(v);
        if (threadIdx.x == 0 && v > 0.) atomic_max_nonneg(&a.diff[a.s_begin + s], v);
      }
      __syncthreads();
    }
    grid.sync();
  }
}

// ---------------------------------------------------------------- "lu": one forward + one backward sweep
struct LuArgs {
  const double* A[7];   // sheared rows
  const double* R[3];   // sheared constants
  double* X[3];         // sheared result
  int ncomp;
};
template <int DIM>
__global__ void __launch_bounds__(SOLVER_THREADS) k_lu_persistent(Geo g, LuArgs a) {
  cg::grid_group grid = cg::this_grid();
  const PlaneTiling pt = plane_tiling(g);
  // forward step (linear.hpp:537-548)
  for (int kp = 0; kp < g.np; ++kp) {
    for (int local = blockIdx.x; local < pt.tiles; local += gridDim.x) {
      int i, j, k;
      if (!tile_cell(g, pt, kp, local, i, j, k)) continue;
      const long long cs = shidx(g, i, j, k);
      const bool zm = DIM > 2 && k > 0, ym = j > 0, xm = i > 0;
      const double azm = zm ? a.A[CZM][cs] : 0., aym = ym ? a.A[CYM][cs] : 0., axm = xm ? a.A[CXM][cs] : 0.;
      const double diag = a.A[CD][cs];
      const long long nzm = zm ? shidx(g, i, j, k - 1) : 0, nym = ym ? shidx(g, i, j - 1, k) : 0, nxm = xm ? shidx(g, i - 1, j, k) : 0;
      for (int n = 0; n < a.ncomp; ++n) {
        double sum = 0.;
        if (zm) sum += azm * __ldcg(&a.X[n][nzm]);
        if (ym) sum += aym * __ldcg(&a.X[n][nym]);
        if (xm) sum += axm * __ldcg(&a.X[n][nxm]);
        a.X[n][cs] = (-a.R[n][cs] - sum) / diag;
      }
    }
    grid.sync();
  }
  // backward step (linear.hpp:551-563)
  for (int kp = g.np - 1; kp >= 0; --kp) {
    for (int local = blockIdx.x; local < pt.tiles; local += gridDim.x) {
      int i, j, k;
      if (!tile_cell(g, pt, kp, local, i, j, k)) continue;
      const long long cs = shidx(g, i, j, k);
      const bool zp = DIM > 2 && k + 1 < g.n[2], yp = j + 1 < g.n[1], xp = i + 1 < g.n[0];
      const double azp = zp ? a.A[CZP][cs] : 0., ayp = yp ? a.A[CYP][cs] : 0., axp = xp ? a.A[CXP][cs] : 0.;
      const double diag = a.A[CD][cs];
      const long long nzp = zp ? shidx(g, i, j, k + 1) : 0, nyp = yp ? shidx(g, i, j + 1, k) : 0, nxp = xp ? shidx(g, i + 1, j, k) : 0;
      for (int n = 0; n < a.ncomp; ++n) {
        double sum = 0.;
        if (zp) sum += azp * __ldcg(&a.X[n][nzp]);
        if (yp) sum += ayp * __ldcg(&a.X[n][nyp]);
        if (xp) sum += axp * __ldcg(&a.X[n][nxp]);
        a.X[n][cs] = __ldcg(&a.X[n][cs]) - sum / diag;
      }
    }
    grid.sync();
  }
}

// ---------------------------------------------------------------- generic-matrix sweeps (hg_linear_solve and
// pressure systems given explicitly): SOR with stored rows, same pipelining as k_gs_persistent
struct SorArgs {
  const double* A[7]; const double* R; double* X; double* diff; int s_begin, s_end; double omega;
};
template <int DIM>
__global__ void __launch_bounds__(SOLVER_THREADS) k_sor_matrix_persistent(Geo g, SorArgs a) {
  cg::grid_group grid = cg::this_grid();
  const PlaneTiling pt = plane_tiling(g);
  const int S = a.s_end - a.s_begin;
  const int Tmax = (g.np - 1) + 2 * (S - 1);
  __shared__ double sm[SOLVER_THREADS / 32];
  for (int T = 0; T <= Tmax; ++T) {
    int smin = (T - (g.np - 1) + 1) / 2; if (T - (g.np - 1) <= 0) smin = 0;
    int smax = T / 2; if (smax > S - 1) smax = S - 1;
    const long long ntiles = (long long)(smax - smin + 1) * pt.tiles;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int s = smin + (int)(tile / pt.tiles);
      const int local = (int)(tile % pt.tiles);
      const int kp = T - 2 * s;
      int i, j, k;
      double ac = 0.;
      if (tile_cell(g, pt, kp, local, i, j, k)) {
        const long long cs = shidx(g, i, j, k);
        double sum = 0.;
        if (DIM > 2 && k > 0) sum += a.A[CZM][cs] * __ldcg(&a.X[shidx(g, i, j, k - 1)]);
        if (j > 0) sum += a.A[CYM][cs] * __ldcg(&a.X[shidx(g, i, j - 1, k)]);
        if (i > 0) sum += a.A[CXM][cs] * __ldcg(&a.X[shidx(g, i - 1, j, k)]);
        if (i + 1 < g.n[0]) sum += a.A[CXP][cs] * __ldcg(&a.X[shidx(g, i + 1, j, k)]);
        if (j + 1 < g.n[1]) sum += a.A[CYP][cs] * __ldcg(&a.X[shidx(g, i, j + 1, k)]);
        if (DIM > 2 && k + 1 < g.n[2]) sum += a.A[CZP][cs] * __ldcg(&a.X[shidx(g, i, j, k + 1)]);
        const double xold = __ldcg(&a.X[cs]);
        const double value = -(a.R[cs] + sum) / a.A[CD][cs];
        const double corr = value - xold;
        a.X[cs] = xold + corr * a.omega;
        ac = fabs(corr);
        if (!(ac == ac)) ac = 0.;
      }
      ac = warp_max(ac);
      if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = ac;
      __syncthreads();
      if (threadIdx.x < 32) {
        double v = threadIdx.x < SOLVER_THREADS / 32 ? sm[threadIdx.x] : 0.;
        v = warp_max(v);
        if (threadIdx.x == 0 && v > 0.) atomic_max_nonneg(&a.diff[a.s_begin + s], v);
      }
      __syncthreads();
    }
    grid.sync();
  }
}

// ---------------------------------------------------------------- Jacobi (linear.hpp:750-782): one launch per sweep
// on the natural layout; pressure rows regenerated from d_c like the Gauss-Seidel kernel.
struct JacArgs {
  const double* A[7];  // natural layout rows, or all nullptr -> regenerate from D
  const double* D;     // natural d_c (only when A is null)
  const double* R; const double* xin; double* xout; double* diff; double omega;
};
template <int DIM>
DV double jac_coeff(const Geo& g, const double* __restrict__ D, double d0, int d, int ni, int nj, int nk, bool& inner) {
  inner = cell_ok(g, ni, nj, nk);
  if (!inner) return 0.;
  const double dn = D[cidx(g, ni, nj, nk)];
  const double dfc = dn * (1. - 0.5) + d0 * 0.5;
  const double coeff = -g.area[d] / (g.h[d] * dfc);
  return -coeff;
}
template <int DIM>
__global__ void k_jacobi_sweep(Geo g, JacArgs a) {
  long long c_ = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nc_ = (long long)g.n[0] * g.n[1] * g.n[2];
  double ac = 0.;
  if (c_ < nc_) {
    const long long c = c_;
    int i = (int)(c_ % g.n[0]); int j = (int)((c_ / g.n[0]) % g.n[1]); int k = (int)(c_ / ((long long)g.n[0] * g.n[1]));
    double sum = 0., diag;
    if (a.A[CD]) {
      if (DIM > 2 && k > 0) sum += a.A[CZM][c] * a.xin[c - g.sz];
      if (j > 0) sum += a.A[CYM][c] * a.xin[c - g.sy];
      if (i > 0) sum += a.A[CXM][c] * a.xin[c - 1];
      if (i + 1 < g.n[0]) sum += a.A[CXP][c] * a.xin[c + 1];
      if (j + 1 < g.n[1]) sum += a.A[CYP][c] * a.xin[c + g.sy];
      if (DIM > 2 && k + 1 < g.n[2]) sum += a.A[CZP][c] * a.xin[c + g.sz];
      diag = a.A[CD][c];
    } else if (cell_excl(g, i, j, k) || c == g.pfix) {
      diag = 1.;
    } else {
      const double d0 = a.D[c];
      bool in_xm, in_xp, in_ym, in_yp, in_zm = false, in_zp = false;
      const double cxm = jac_coeff<DIM>(g, a.D, d0, 0, i - 1, j, k, in_xm);
      const double cxp = jac_coeff<DIM>(g, a.D, d0, 0, i + 1, j, k, in_xp);
      const double cym = jac_coeff<DIM>(g, a.D, d0, 1, i, j - 1, k, in_ym);
      const double cyp = jac_coeff<DIM>(g, a.D, d0, 1, i, j + 1, k, in_yp);
      double czm = 0., czp = 0.;
      if (DIM > 2) { czm = jac_coeff<DIM>(g, a.D, d0, 2, i, j, k - 1, in_zm); czp = jac_coeff<DIM>(g, a.D, d0, 2, i, j, k + 1, in_zp); }
      bool have = false; diag = 0.;
      if (in_xm) { diag = have ? diag + cxm : cxm; have = true; }
      if (in_xp) { diag = have ? diag + cxp : cxp; have = true; }
      if (in_ym) { diag = have ? diag + cym : cym; have = true; }
      if (in_yp) { diag = have ? diag + cyp : cyp; have = true; }
      if (DIM > 2) {
        if (in_zm) { diag = have ? diag + czm : czm; have = true; }
        if (in_zp) { diag = have ? diag + czp : czp; have = true; }
      }
      const long long pf = g.pfix;
      if (DIM > 2 && in_zm && c - g.sz != pf) sum += (-czm) * a.xin[c - g.sz];
      if (in_ym && c - g.sy != pf) sum += (-cym) * a.xin[c - g.sy];
      if (in_xm && c - 1 != pf) sum += (-cxm) * a.xin[c - 1];
      if (in_xp && c + 1 != pf) sum += (-cxp) * a.xin[c + 1];
      if (in_yp && c + g.sy != pf) sum += (-cyp) * a.xin[c + g.sy];
      if (DIM > 2 && in_zp && c + g.sz != pf) sum += (-czp) * a.xin[c + g.sz];
    }
    const double xold = a.xin[c];
    const double value = -(a.R[c] + sum) / diag;
    const double corr = value - xold;
    a.xout[c] = xold + corr * a.omega;
    ac = fabs(corr);
    if (!(ac == ac)) ac = 0.;
  }
  ac = warp_max(ac);
  __shared__ double sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = ac;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.;
    v = warp_max(v);
    if (threadIdx.x == 0 && v > 0.) atomic_max_nonneg(a.diff, v);
  }
}
