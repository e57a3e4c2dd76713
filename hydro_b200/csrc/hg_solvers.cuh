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
#include "hg_slab.cuh"

namespace cg = cooperative_groups;

constexpr int SOLVER_BX = 64;   // threads along i
constexpr int SOLVER_BY = 4;    // rows (j) per tile
constexpr int SOLVER_TILE = SOLVER_BX * SOLVER_BY;   // cells per tile = threads per tile slot
constexpr int SOLVER_SLOTS = 4;                       // tile slots per CTA
constexpr int SOLVER_THREADS = SOLVER_TILE * SOLVER_SLOTS;   // 1024: one CTA per SM, 148 CTAs -> cheap grid barrier
constexpr int SOLVER_SC = 1024;                       // max sweeps in flight per launch (shared table)

// Grid-wide barrier for the persistent (cooperatively launched, hence co-resident) solver kernels:
// one atomic arrival per CTA plus a generation word.  Cheaper than cg::grid_group::sync(), which also
// invalidates L1 on every poll; data exchanged between SMs across the barrier (the solution vector)
// is therefore always read with ld.global.cg (__ldcg).
// The host zeroes the arrival counter before every launch; barrier number `epoch` (0,1,2,...) is complete
// when the counter reaches (epoch+1)*nblocks.  Arrival is a release atomic, the poll an acquire load, so the
// critical path is one atomic plus one load round trip through L2.
DV void grid_barrier(unsigned long long* bar, unsigned int nblocks, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long target = (unsigned long long)(epoch + 1u) * nblocks;
    unsigned long long old;
    asm volatile("atom.add.release.gpu.global.u64 %0, [%1], 1;" : "=l"(old) : "l"(bar) : "memory");
    if (old + 1ull < target) {
      unsigned long long cur;
      do {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(cur) : "l"(bar) : "memory");
      } while (cur < target);
    } else {
      __threadfence();
    }
  }
  ++epoch;
  __syncthreads();
}
// Hyperplane tiles.  A tile is SOLVER_BY consecutive rows j x SOLVER_BX consecutive i of one plane
// k' = i+j+k; only tiles that contain at least one cell are listed (built once per mesh on the host).
struct TileTable {
  const int* tile_j0;    // first row of the tile
  const int* tile_i0;    // first i of the tile
  const int* tileoff;    // [np+1] first tile of plane k'
  const int* cum2;       // [np] tiles of planes k', k'-2, k'-4, ... (same-parity running sum)
  unsigned long long* bar;   // grid barrier arrival counter (zeroed by the host before each launch)
};
// cell handled by this thread in tile e of plane kp
DV bool tile_cell(const Geo& g, const TileTable& tt, int e, int kp, int& i, int& j, int& k) {
  const int lt = threadIdx.x & (SOLVER_TILE - 1);
  j = tt.tile_j0[e] + (lt / SOLVER_BX);
  i = tt.tile_i0[e] + (lt % SOLVER_BX);
  k = kp - i - j;
  return j < g.n[1] && i < g.n[0] && k >= 0 && k < g.n[2];
}
// Tiles of step T for sweeps [smin, smax] (planes T-2s): flat id -> (plane, tile) by binary search in cum2
struct StepTiles { int kp_lo, kp_hi, base, total; };
// T is the LOCAL step (global step minus the slab's first plane k0); it may lie outside the local range
DV StepTiles step_tiles(const Geo& g, const TileTable& tt, int T, int S) {
  StepTiles st; st.kp_lo = st.kp_hi = 0; st.base = 0; st.total = 0;
  if (T < 0) return st;
  int smin = 0; if (T - (g.np - 1) > 0) smin = (T - (g.np - 1) + 1) / 2;
  int smax = T / 2; if (smax > S - 1) smax = S - 1;
  if (smax < smin) return st;
  st.kp_hi = T - 2 * smin; st.kp_lo = T - 2 * smax;
  st.base = st.kp_lo >= 2 ? tt.cum2[st.kp_lo - 2] : 0;
  st.total = tt.cum2[st.kp_hi] - st.base;
  return st;
}
// Per step: sc[m] = number of tiles of the m+1 lowest active planes (inclusive running count), kept in
// shared memory; a slot walks its ids in ascending order, so (plane, tile) follow by a two-pointer scan.
DV void fill_step_table(const TileTable& tt, const StepTiles& st, int* sc) {
  const int nact = st.total > 0 ? ((st.kp_hi - st.kp_lo) >> 1) + 1 : 0;
  for (int t = threadIdx.x; t < nact; t += blockDim.x) sc[t] = tt.cum2[st.kp_lo + 2 * t] - st.base;
  __syncthreads();
}

// ---------------------------------------------------------------- pressure: Gauss-Seidel / SOR
struct GsArgs {
  const double* CX;    // c_f = A/(h d_f) of the x+ face of each cell (0 when that face is not inner), sheared
  const double* CY;    // same for the y+ face
  const double* CZ;    // same for the z+ face
  const double* RP;    // row constants, sheared
  double* PP;          // solution, sheared; must be zero on entry for sweep 0 (linear.hpp:686)
  double* diff;        // per-sweep max |value - x| (linear.hpp:707), indexed by absolute sweep number
  int s_begin, s_end;  // sweeps [s_begin, s_end) are run by this launch
  double omega;
  TileTable tt;
  SlabLink link;       // neighbouring slabs (multi-GPU), link.on == 0 on a single GPU
};

// Rows of the pressure-correction system (fluid.hpp:972-1014) are rebuilt on the fly from the three face
// coefficient fields: diagonal = ordered sum of the six face coefficients (absent faces store 0, and
// x + 0 == x, so the reference's "merge only existing terms" gives the same bits), off-diagonals = -c_f.
template <int DIM, bool EXCL, bool LINK = false>
__global__ void __launch_bounds__(SOLVER_THREADS, 1) k_gs_persistent(Geo g, GsArgs a) {
  __shared__ int sc[SOLVER_SC];
  const int S = a.s_end - a.s_begin;
  // slab decomposition: every rank walks its own hyperplanes; the interface values arrive tagged with their sweep
  // (ll_wait), which is all the coupling the lexicographic order needs
  const SlabLink L = a.link;
  const int Tmax = (g.np - 1) + 2 * (S - 1);
  const long long PS = (long long)g.n[1] * g.n[0];   // plane stride of the sheared layout
  const int nx = g.n[0];
  const int gslot = blockIdx.x * SOLVER_SLOTS + (threadIdx.x / SOLVER_TILE), nslots = gridDim.x * SOLVER_SLOTS;
  unsigned int epoch = 0;
  for (int T = 0; T <= Tmax; ++T) {
    const StepTiles st = step_tiles(g, a.tt, T, S);
    fill_step_table(a.tt, st, sc);
    // contiguous id range per slot (neighbouring tiles share rows -> cache reuse); first plane by binary search
    const int chunk = (st.total + nslots - 1) / nslots;
    const int id0 = gslot * chunk, id1 = min(id0 + chunk, st.total);
    int m = 0;
    if (id0 < id1) {
      int lo = 0, hi = (st.kp_hi - st.kp_lo) >> 1;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (sc[mid] > id0) hi = mid; else lo = mid + 1; }
      m = lo;
    }
    for (int id = id0; id < id1; ++id) {
      while (sc[m] <= id) ++m;
      const int kp = st.kp_lo + 2 * m;
      const int e = a.tt.tileoff[kp] + id - (m > 0 ? sc[m - 1] : 0);
      const int s = (T - kp) >> 1;
      int i, j, k;
      double ac = 0.;
      if (tile_cell(g, a.tt, e, kp, i, j, k)) {
        const long long cs = ((long long)(kp + 1) * g.n[1] + j) * nx + i;
        const long long c = i + g.sy * j + g.sz * k;
        // neighbours in the sheared layout: x-: cs-PS-1, y-: cs-PS-nx, z-: cs-PS, x+: cs+PS+1, y+: cs+PS+nx, z+: cs+PS
        const bool in_xm = i > 0, in_xp = i + 1 < nx, in_ym = j > 0, in_yp = j + 1 < g.n[1];
        const bool in_zm = DIM > 2 && (k > 0 || g.zlo > 0), in_zp = DIM > 2 && (k + 1 < g.n[2] || g.zhi > 0);
        bool ident = c == g.pfix;
        if (EXCL) ident = ident || g.excl[c] != 0;
        // all loads up front (independent addresses)
        const double rhs = a.RP[cs];
        const double xold = __ldcg(&a.PP[cs]);
        const double cxp = a.CX[cs], cyp = a.CY[cs], czp = DIM > 2 ? a.CZ[cs] : 0.;
        const double cxm = in_xm ? a.CX[cs - PS - 1] : 0., cym = in_ym ? a.CY[cs - PS - nx] : 0.;
        const double czm = in_zm ? a.CZ[cs - PS] : 0.;
        const double pxm = in_xm ? __ldcg(&a.PP[cs - PS - 1]) : 0., pxp = in_xp ? __ldcg(&a.PP[cs + PS + 1]) : 0.;
        const double pym = in_ym ? __ldcg(&a.PP[cs - PS - nx]) : 0., pyp = in_yp ? __ldcg(&a.PP[cs + PS + nx]) : 0.;
        // slab interfaces: the neighbour's value arrives tagged with its sweep; z- needs this sweep, z+ the previous one
        const long long c2 = (long long)j * nx + i;
        const unsigned tg = L.tag0 + (unsigned)(a.s_begin + s);
        double pzm = 0., pzp = 0.;
        if (in_zm) pzm = (k > 0 || !LINK) ? __ldcg(&a.PP[cs - PS]) : ll_wait(L.from_lo + c2, tg + 1u, L.err);
        if (in_zp) pzp = (k + 1 < g.n[2] || !LINK) ? __ldcg(&a.PP[cs + PS]) : ll_wait(L.from_hi + c2, tg, L.err);
        double diag = 1., sum = 0.;
        if (!ident) {
          // diagonal: face order x-,x+,y-,y+,z-,z+ (fluid.hpp:979-984)
          diag = cxm + cxp; diag = diag + cym; diag = diag + cyp;
          if (DIM > 2) { diag = diag + czm; diag = diag + czp; }
          // off-diagonal terms in ascending index order z-,y-,x-,x+,y+,z+ (linear.hpp:694-701);
          // terms toward the fixed-pressure cell were removed by SetKnownValue (fluid.hpp:1010)
          const long long pf = g.pfix;
          if (in_zm && c - g.sz != pf) sum += (-czm) * pzm;
          if (in_ym && c - g.sy != pf) sum += (-cym) * pym;
          if (in_xm && c - 1 != pf) sum += (-cxm) * pxm;
          if (in_xp && c + 1 != pf) sum += (-cxp) * pxp;
          if (in_yp && c + g.sy != pf) sum += (-cyp) * pyp;
          if (in_zp && c + g.sz != pf) sum += (-czp) * pzp;
        }
        const double value = -(rhs + sum) / diag;
        const double corr = value - xold;
        const double xnew = xold + corr * a.omega;
        a.PP[cs] = xnew;
        // interface cells: the new value is also the neighbour slab's halo value (peer store; made visible by the
        // system-scope fence + flag of the step barrier, which is cumulative over the CTAs' release arrivals)
        if (LINK && DIM > 2) {
          if (k == g.n[2] - 1 && L.has_hi) ll_store(L.to_hi + c2, xnew, tg + 1u);
          if (k == 0 && L.has_lo) ll_store(L.to_lo + c2, xnew, tg + 1u);
        }
        ac = fabs(corr);
        if (!(ac == ac)) ac = 0.;
      }
      // per-sweep max-norm (linear.hpp:707): one reduction atomic per warp
      ac = warp_max(ac);
      if ((threadIdx.x & 31) == 0 && ac > 0.) atomic_max_nonneg(&a.diff[a.s_begin + s], ac);
    }
    grid_barrier(a.tt.bar, gridDim.x, epoch);
  }
}

// ---------------------------------------------------------------- "lu": one forward + one backward sweep
struct LuArgs {
  const double* A[7];   // sheared rows
  const double* R[3];   // sheared constants
  double* X[3];         // sheared result
  int ncomp;
  TileTable tt;
  SlabLink link;       // lu: planes [n] of from_lo/to_hi and from_hi/to_lo belong to component n (stride link_stride)
  long long link_stride;
};
template <int DIM, bool LINK = false>
__global__ void __launch_bounds__(SOLVER_THREADS, 1) k_lu_persistent(Geo g, LuArgs a) {
  const long long PS = (long long)g.n[1] * g.n[0];
  const int nx = g.n[0];
  const int gslot = blockIdx.x * SOLVER_SLOTS + (threadIdx.x / SOLVER_TILE), nslots = gridDim.x * SOLVER_SLOTS;
  const SlabLink L = a.link;
  unsigned int epoch = 0;
  // forward step (linear.hpp:537-548), planes in ascending order
  for (int kp = 0; kp < g.np; ++kp) {
    for (int e = a.tt.tileoff[kp] + gslot; e < a.tt.tileoff[kp + 1]; e += nslots) {
      int i, j, k;
      if (!tile_cell(g, a.tt, e, kp, i, j, k)) continue;
      const long long cs = ((long long)(kp + 1) * g.n[1] + j) * nx + i;
      const bool zm = DIM > 2 && (k > 0 || g.zlo > 0), ym = j > 0, xm = i > 0;
      const double azm = zm ? a.A[CZM][cs] : 0., aym = ym ? a.A[CYM][cs] : 0., axm = xm ? a.A[CXM][cs] : 0.;
      const double diag = a.A[CD][cs];
      for (int n = 0; n < a.ncomp; ++n) {
        double sum = 0.;
        if (zm) sum += azm * ((k > 0 || !LINK) ? __ldcg(&a.X[n][cs - PS]) : ll_wait(L.from_lo + n * a.link_stride + (long long)j * nx + i, L.tag0 + 1u, L.err));
        if (ym) sum += aym * __ldcg(&a.X[n][cs - PS - nx]);
        if (xm) sum += axm * __ldcg(&a.X[n][cs - PS - 1]);
        const double xv = (-a.R[n][cs] - sum) / diag;
        a.X[n][cs] = xv;
        if (LINK && DIM > 2 && k == g.n[2] - 1 && L.has_hi) ll_store(L.to_hi + n * a.link_stride + (long long)j * nx + i, xv, L.tag0 + 1u);
      }
    }
    grid_barrier(a.tt.bar, gridDim.x, epoch);
  }
  // backward step (linear.hpp:551-563)
  for (int kp = g.np - 1; kp >= 0; --kp) {
    for (int e = a.tt.tileoff[kp] + gslot; e < a.tt.tileoff[kp + 1]; e += nslots) {
      int i, j, k;
      if (!tile_cell(g, a.tt, e, kp, i, j, k)) continue;
      const long long cs = ((long long)(kp + 1) * g.n[1] + j) * nx + i;
      const bool zp = DIM > 2 && (k + 1 < g.n[2] || g.zhi > 0), yp = j + 1 < g.n[1], xp = i + 1 < nx;
      const double azp = zp ? a.A[CZP][cs] : 0., ayp = yp ? a.A[CYP][cs] : 0., axp = xp ? a.A[CXP][cs] : 0.;
      const double diag = a.A[CD][cs];
      for (int n = 0; n < a.ncomp; ++n) {
        double sum = 0.;
        if (zp) sum += azp * ((k + 1 < g.n[2] || !LINK) ? __ldcg(&a.X[n][cs + PS]) : ll_wait(L.from_hi + n * a.link_stride + (long long)j * nx + i, L.tag0 + 2u, L.err));
        if (yp) sum += ayp * __ldcg(&a.X[n][cs + PS + nx]);
        if (xp) sum += axp * __ldcg(&a.X[n][cs + PS + 1]);
        const double xv = __ldcg(&a.X[n][cs]) - sum / diag;
        a.X[n][cs] = xv;
        if (LINK && DIM > 2 && k == 0 && L.has_lo) ll_store(L.to_lo + n * a.link_stride + (long long)j * nx + i, xv, L.tag0 + 2u);
      }
    }
    grid_barrier(a.tt.bar, gridDim.x, epoch);
  }
}

// ---------------------------------------------------------------- "lu_relaxed" (linear.hpp:592-650)
// forward step: corr_i = (-f_i - sum_{j<i} a_ij corr_j) / (a_ii + relax), ordered -> hyperplane wavefront
struct LurArgs { const double* A[7]; const double* F; double* corr; double relax; TileTable tt; };
template <int DIM>
__global__ void __launch_bounds__(SOLVER_THREADS, 1) k_lur_forward(Geo g, LurArgs a) {
  const long long PS = (long long)g.n[1] * g.n[0];
  const int nx = g.n[0];
  const int gslot = blockIdx.x * SOLVER_SLOTS + (threadIdx.x / SOLVER_TILE), nslots = gridDim.x * SOLVER_SLOTS;
  unsigned int epoch = 0;
  for (int kp = 0; kp < g.np; ++kp) {
    for (int e = a.tt.tileoff[kp] + gslot; e < a.tt.tileoff[kp + 1]; e += nslots) {
      int i, j, k;
      if (!tile_cell(g, a.tt, e, kp, i, j, k)) continue;
      const long long cs = ((long long)(kp + 1) * g.n[1] + j) * nx + i;
      double sum = 0.;
      if (DIM > 2 && k > 0) sum += a.A[CZM][cs] * __ldcg(&a.corr[cs - PS]);
      if (j > 0) sum += a.A[CYM][cs] * __ldcg(&a.corr[cs - PS - nx]);
      if (i > 0) sum += a.A[CXM][cs] * __ldcg(&a.corr[cs - PS - 1]);
      a.corr[cs] = (-a.F[cs] - sum) / (a.A[CD][cs] + a.relax);
    }
    grid_barrier(a.tt.bar, gridDim.x, epoch);
  }
}
// backward step: corr_i -= (sum_{j>i} a_ij res_j) / (a_ii + relax) -- the reference reads `res`, not `corr`
// (linear.hpp:632), so this step has no recurrence
template <int DIM>
__global__ void k_lur_backward(Geo g, LurArgs a, const double* __restrict__ res) {
  long long c_ = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nc_ = (long long)g.n[0] * g.n[1] * g.n[2];
  if (c_ >= nc_) return;
  const int i = (int)(c_ % g.n[0]), j = (int)((c_ / g.n[0]) % g.n[1]), k = (int)(c_ / ((long long)g.n[0] * g.n[1]));
  const long long cs = shidx(g, i, j, k);
  const long long PS = (long long)g.n[1] * g.n[0];
  double sum = 0.;
  if (DIM > 2 && k + 1 < g.n[2]) sum += a.A[CZP][cs] * res[cs + PS];
  if (j + 1 < g.n[1]) sum += a.A[CYP][cs] * res[cs + PS + g.n[0]];
  if (i + 1 < g.n[0]) sum += a.A[CXP][cs] * res[cs + PS + 1];
  a.corr[cs] -= sum / (a.A[CD][cs] + a.relax);
}
// res += corr, diff = max |corr| (linear.hpp:641-644)
template <int DIM>
__global__ void k_lur_update(Geo g, const double* __restrict__ corr, double* __restrict__ res, double* diff) {
  long long c_ = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nc_ = (long long)g.n[0] * g.n[1] * g.n[2];
  double ac = 0.;
  if (c_ < nc_) {
    const int i = (int)(c_ % g.n[0]), j = (int)((c_ / g.n[0]) % g.n[1]), k = (int)(c_ / ((long long)g.n[0] * g.n[1]));
    const long long cs = shidx(g, i, j, k);
    const double cr = corr[cs];
    res[cs] += cr;
    ac = fabs(cr);
    if (!(ac == ac)) ac = 0.;
  }
  ac = warp_max(ac);
  if ((threadIdx.x & 31) == 0 && ac > 0.) atomic_max_nonneg(diff, ac);
}
// f = system.Evaluate(res): constant + sum over terms in ascending index order (linear.hpp:645-647)
template <int DIM>
__global__ void k_lur_residual(Geo g, LurArgs a, const double* __restrict__ R, const double* __restrict__ res, double* __restrict__ f) {
  long long c_ = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nc_ = (long long)g.n[0] * g.n[1] * g.n[2];
  if (c_ >= nc_) return;
  const int i = (int)(c_ % g.n[0]), j = (int)((c_ / g.n[0]) % g.n[1]), k = (int)(c_ / ((long long)g.n[0] * g.n[1]));
  const long long cs = shidx(g, i, j, k);
  const long long PS = (long long)g.n[1] * g.n[0];
  const int nx = g.n[0];
  double r = R[cs];
  if (DIM > 2 && k > 0) r += res[cs - PS] * a.A[CZM][cs];
  if (j > 0) r += res[cs - PS - nx] * a.A[CYM][cs];
  if (i > 0) r += res[cs - PS - 1] * a.A[CXM][cs];
  r += res[cs] * a.A[CD][cs];
  if (i + 1 < nx) r += res[cs + PS + 1] * a.A[CXP][cs];
  if (j + 1 < g.n[1]) r += res[cs + PS + nx] * a.A[CYP][cs];
  if (DIM > 2 && k + 1 < g.n[2]) r += res[cs + PS] * a.A[CZP][cs];
  f[cs] = r;
}

// ---------------------------------------------------------------- generic-matrix sweeps (hg_linear_solve and
// pressure systems given explicitly): SOR with stored rows, same pipelining as k_gs_persistent
struct SorArgs {
  const double* A[7]; const double* R; double* X; double* diff; int s_begin, s_end; double omega; TileTable tt;
};
template <int DIM>
__global__ void __launch_bounds__(SOLVER_THREADS, 1) k_sor_matrix_persistent(Geo g, SorArgs a) {
  __shared__ int sc[SOLVER_SC];
  const int S = a.s_end - a.s_begin;
  const int Tmax = (g.np - 1) + 2 * (S - 1);
  const long long PS = (long long)g.n[1] * g.n[0];
  const int nx = g.n[0];
  const int gslot = blockIdx.x * SOLVER_SLOTS + (threadIdx.x / SOLVER_TILE), nslots = gridDim.x * SOLVER_SLOTS;
  unsigned int epoch = 0;
  for (int T = 0; T <= Tmax; ++T) {
    const StepTiles st = step_tiles(g, a.tt, T, S);
    fill_step_table(a.tt, st, sc);
    // contiguous id range per slot (neighbouring tiles share rows -> cache reuse); first plane by binary search
    const int chunk = (st.total + nslots - 1) / nslots;
    const int id0 = gslot * chunk, id1 = min(id0 + chunk, st.total);
    int m = 0;
    if (id0 < id1) {
      int lo = 0, hi = (st.kp_hi - st.kp_lo) >> 1;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (sc[mid] > id0) hi = mid; else lo = mid + 1; }
      m = lo;
    }
    for (int id = id0; id < id1; ++id) {
      while (sc[m] <= id) ++m;
      const int kp = st.kp_lo + 2 * m;
      const int e = a.tt.tileoff[kp] + id - (m > 0 ? sc[m - 1] : 0);
      const int s = (T - kp) >> 1;
      int i, j, k;
      double ac = 0.;
      if (tile_cell(g, a.tt, e, kp, i, j, k)) {
        const long long cs = ((long long)(kp + 1) * g.n[1] + j) * nx + i;
        double sum = 0.;
        if (DIM > 2 && k > 0) sum += a.A[CZM][cs] * __ldcg(&a.X[cs - PS]);
        if (j > 0) sum += a.A[CYM][cs] * __ldcg(&a.X[cs - PS - nx]);
        if (i > 0) sum += a.A[CXM][cs] * __ldcg(&a.X[cs - PS - 1]);
        if (i + 1 < nx) sum += a.A[CXP][cs] * __ldcg(&a.X[cs + PS + 1]);
        if (j + 1 < g.n[1]) sum += a.A[CYP][cs] * __ldcg(&a.X[cs + PS + nx]);
        if (DIM > 2 && k + 1 < g.n[2]) sum += a.A[CZP][cs] * __ldcg(&a.X[cs + PS]);
        const double xold = __ldcg(&a.X[cs]);
        const double value = -(a.R[cs] + sum) / a.A[CD][cs];
        const double corr = value - xold;
        a.X[cs] = xold + corr * a.omega;
        ac = fabs(corr);
        if (!(ac == ac)) ac = 0.;
      }
      ac = warp_max(ac);
      if ((threadIdx.x & 31) == 0 && ac > 0.) atomic_max_nonneg(&a.diff[a.s_begin + s], ac);
    }
    grid_barrier(a.tt.bar, gridDim.x, epoch);
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
