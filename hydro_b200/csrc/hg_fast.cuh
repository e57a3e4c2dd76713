// hg_fast.cuh -- interior kernels of the SIMPLE iteration (3-D): the cells whose whole stencil (radius 1; advection: 2) lies inside the
// mesh and touches no excluded cell and no fixed-pressure cell.  On those cells every face is an inner face, so the
// boundary-condition machinery of hg_kernels.cuh (face classification, excluded-cell mask, extrapolation) folds away:
// the arithmetic below is the generic kernels' arithmetic for inner faces, operation by operation (same association, same
// divisions), so the results are bit-identical.  The remaining cells (a shell one or two cells thick, the neighbourhood of an
// obstacle and of the fixed-pressure cell: 2.3 % / 4.6 % at 256^3) are listed once at creation and keep the generic kernels, which
// are launched over that list (Geo::cells).
//
// Round 1 ran the generic code on every cell: the kernels were instruction bound (k_assemble 2840 instructions per cell,
// 0.25 of the HBM roofline), the rows went through a natural-layout staging copy and a separate transpose, and the
// explicit viscous term, the pressure gradient and the restored force each made their own pass over HBM.  Here:
//   k_fa_grad      = K_velgrad + K_pre                 (fluid.hpp:827-843, 602-631)
//   k_fb_momentum  = K_source + K_assemble + transpose (fluid.hpp:835-892, conv_diff.hpp:149-227); rows and constants leave
//                    through a shared-memory tile straight into the hyperplane-major layout of the lu sweeps
//   k_fc_flux_rows = K_fstar + K_prhs + row packing    (fluid.hpp:903-1014); CO5 rows of k_gs_tiled written through the tile
//   k_fd_correct   = K_correct                         (fluid.hpp:1040-1056)
//   k_fe_advect    = K_advect                          (advection.hpp:440-480)
// plus the layout conversions of the solver results (k_ft_apply_corr, k_ft_pcorr) as tile transposes.
//
// Thread layout: a CTA is a tile of 32 (i) x FT_K (k) cells at fixed j, one warp per k: x neighbours are in the warp's own
// 256-byte row, z neighbours in the CTA (L1), y neighbours come from L2.  A diagonal i + k = const of the tile is a
// contiguous run of the hyperplane-major arrays (index ((i+j+k+1) ny + j) nx + i).
#pragma once
#include "hg_device.cuh"
#include "hg_kernels.cuh"
#include "hg_gs_tiled.cuh"

constexpr int FT_K = 8;
constexpr int FT_THREADS = 32 * FT_K;
constexpr int FT_PITCH = 32;   // tile row pitch (doubles): a diagonal read (il + 1, kl - 1) moves by -31 words -> distinct banks

#define FT_PROLOG(g)                                                                        \
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;                                   \
  const int i0 = blockIdx.x * 32, k0 = blockIdx.y * FT_K, j = blockIdx.z;                   \
  const int i = i0 + tx, k = k0 + ty;                                                       \
  const bool in = i < (g).n[0] && k < (g).n[2];                                             \
  const long long c = cidx(g, i, j, k);

#ifndef FT_PREFETCH
#define FT_PREFETCH 1
#endif
// The interior kernels are gathers with ~100 loads per cell and few warps per SM: their first touch of every input row is
// requested up front (no register is tied up, all rows of a CTA are in flight at once), the loads then hit L2.
DV void ft_prefetch(const void* p) {
#if FT_PREFETCH
  asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
#endif
}
DV double ft_avg(double a, double b) { return a * (1. - 0.5) + b * 0.5; }   // Interpolate on an inner face (solver.hpp:425-426)
// Gradient(Interpolate(u))[d] of a cell with inner faces (solver.hpp:658-677): um, uc, up = u at c - e_d, c, c + e_d
DV double ft_grad(double um, double uc, double up, double aneg, double apos, const HgDiv& dvol) {
  const double fm = ft_avg(um, uc), fp = ft_avg(uc, up);
  double sum = 0.;
  sum += aneg * fm;
  sum += apos * fp;
  return hg_div(sum, dvol);
}

// Marks the cells that keep the generic kernels: some cell of the (2 R + 1)^3 cube around them is outside the (global) mesh
// or excluded, or the fixed-pressure cell is within one cell.  Covers the owned cells.  R = 1 for the kernels that read their
// neighbours' results from memory (all faces of the cell are inner faces), R = 2 for the advection (gradients of the
// neighbouring cells are formed in place).
__global__ void k_fast_mask(Geo g, int R, unsigned char* __restrict__ slow) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nc = (long long)g.n[0] * g.n[1] * g.n[2];
  if (t >= nc) return;
  const int i = (int)(t % g.n[0]), j = (int)((t / g.n[0]) % g.n[1]), k = (int)(t / ((long long)g.n[0] * g.n[1]));
  bool s = false;
  for (int dk = -R; dk <= R; ++dk) for (int dj = -R; dj <= R; ++dj) for (int di = -R; di <= R; ++di) {
    const int kg = k + dk + g.k0;
    if (kg < 0 || kg >= g.nzg) { s = true; continue; }
    if (!cell_ok(g, i + di, j + dj, k + dk)) s = true;
    if (g.pfix != HG_NO_CELL && abs(di) <= 1 && abs(dj) <= 1 && abs(dk) <= 1 && cidx(g, i + di, j + dj, k + dk) == g.pfix) s = true;
  }
  slow[t] = s ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ K_velgrad + K_pre
struct FaArgs {
  const double* u[3]; const double* force[3]; const double* p;
  double* G[9]; double* fcr[3]; double* gp[3];
};
__global__ void __launch_bounds__(FT_THREADS, 6) k_fa_grad(Geo g, const unsigned char* __restrict__ slow, FaArgs a) {
  FT_PROLOG(g)
  if (!in) return;
#pragma unroll
  for (int d = 0; d < 3; ++d) { ft_prefetch(a.u[d] + c); ft_prefetch(a.force[d] + c); }
  ft_prefetch(a.p + c);
  if (slow[c]) return;
  const HgDiv dvol = hg_div_prepare(g.vol);
  const long long off[3] = {1, g.sy, g.sz};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double aneg = g.area[d] * -1., apos = g.area[d] * 1.;
    // restored force (CalcExtForce, fluid.hpp:602-631)
    {
      const double* __restrict__ f = a.force[d];
      const double f0 = f[c];
      const double fm = ft_avg(f[c - off[d]], f0), fp = ft_avg(f0, f[c + off[d]]);
      double sum = 0.;
      sum += (g.area[d] * fm) * (0.5 * g.h[d]);
      sum += (g.area[d] * fp) * (0.5 * g.h[d]);
      a.fcr[d][c] = hg_div(sum, dvol);
    }
    // pressure gradient (fluid.hpp:827-829)
    {
      const double* __restrict__ p = a.p;
      a.gp[d][c] = ft_grad(p[c - off[d]], p[c], p[c + off[d]], aneg, apos, dvol);
    }
  }
  // velocity gradients G[n*3+d] (fluid.hpp:838-843)
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    const double* __restrict__ u = a.u[n];
    const double uc = u[c];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      a.G[n * 3 + d][c] = ft_grad(u[c - off[d]], uc, u[c + off[d]], g.area[d] * -1., g.area[d] * 1., dvol);
  }
}

// ------------------------------------------------------------------------------ tile -> hyperplane-major store / load
// tile[a][kl][il] (pitch FT_PITCH) of NA arrays -> out[a][shidx]: every warp takes four diagonals per pass, eight lanes each
// (a diagonal of the 32 x FT_K tile has at most FT_K cells).  act[kl] = ballot of the threads whose cell is written.
template <int NA, class IndexFn>
DV void ft_store_diagonals(const double* tile, const unsigned* act, double* const* out, int nx, int nz, int i0, int k0, IndexFn index) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int p = lane & 7, sub = lane >> 3;
  static_assert(FT_K == 8, "eight lanes per diagonal");
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int D = pass * 32 + w * 4 + sub;          // diagonal il + kl = D, 0 .. 38
    const int il = (D > FT_K - 1 ? D - (FT_K - 1) : 0) + p, kl = D - il;
    const bool ok = D < 32 + FT_K - 1 && il <= 31 && kl >= 0 && i0 + il < nx && k0 + kl < nz && ((act[kl] >> il) & 1u);
    if (ok) {
      const long long cs = index(i0 + il, k0 + kl);
#pragma unroll
      for (int q = 0; q < NA; ++q) out[q][cs] = tile[(q * FT_K + kl) * FT_PITCH + il];
    }
  }
}

// ------------------------------------------------------------------------------------------------ K_source + K_assemble
#ifndef FB_MINB
#define FB_MINB 4
#endif
struct FbArgs {
  const double* prev[3]; const double* tc[3]; const double* tp[3];
  const double* G[9]; const double* rho; const double* mu; const double* F;
  const double* gp[3]; const double* fcr[3]; const double* stf[3]; int use_stf;
  double co[3]; double relax;
  double* out[10];     // A[7] (z-,y-,x-,d,x+,y+,z+) and R[3], hyperplane-major
  double* coeffsum;    // natural
};
__global__ void __launch_bounds__(FT_THREADS, FB_MINB) k_fb_momentum(Geo g, const unsigned char* __restrict__ slow, FbArgs a) {
  __shared__ double tile[10 * FT_K * FT_PITCH];
  __shared__ unsigned act[FT_K];
  FT_PROLOG(g)
  if (in) {
#pragma unroll
    for (int q = 0; q < 9; ++q) ft_prefetch(a.G[q] + c);
#pragma unroll
    for (int n = 0; n < 3; ++n) { ft_prefetch(a.prev[n] + c); ft_prefetch(a.tc[n] + c); ft_prefetch(a.tp[n] + c); ft_prefetch(a.gp[n] + c); ft_prefetch(a.fcr[n] + c);
                                  if (a.use_stf) ft_prefetch(a.stf[n] + c); ft_prefetch(a.F + fidx(g, n, i, j, k)); }
    ft_prefetch(a.rho + c); ft_prefetch(a.mu + c);
  }
  const bool fast = in && !slow[c];
  const unsigned bal = __ballot_sync(0xffffffffu, fast);
  if (tx == 0) act[ty] = bal;
  if (fast) {
    const HgDiv dvol = hg_div_prepare(g.vol);
    const long long off[3] = {1, g.sy, g.sz};
    // face indices (mesh.hpp:698-705): minus face of the cell in direction d, the plus face is one face stride further
    const long long fst[3] = {1, g.n[0], (long long)g.n[0] * g.n[1]};
    const long long fm_[3] = {fidx(g, 0, i, j, k), fidx(g, 1, i, j, k), fidx(g, 2, i, j, k)};
    const double mu_c = a.mu[c];
    double muf[6];
#pragma unroll
    for (int d = 0; d < 3; ++d) { muf[2 * d] = ft_avg(a.mu[c - off[d]], mu_c); muf[2 * d + 1] = ft_avg(mu_c, a.mu[c + off[d]]); }
    // ---- explicit viscous term + (-grad p + restored force + surface tension) (fluid.hpp:835-870)
    double src[3];
    {
      double acc[3] = {0., 0., 0.};
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        const double wm = muf[2 * n] * (g.area[n] * -1.), wp = muf[2 * n + 1] * (g.area[n] * 1.);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double* __restrict__ G = a.G[n * 3 + d];
          const double g0 = G[c];
          const double gm = ft_avg(G[c - off[n]], g0), gq = ft_avg(g0, G[c + off[n]]);
          double sum = 0.;
          sum += gm * wm;
          sum += gq * wp;
          acc[d] += hg_div(sum, dvol);
        }
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double st = a.use_stf ? a.stf[d][c] : 0.;
        const double t = ((a.gp[d][c] * (-1.) + a.fcr[d][c]) + st) + 0.;
        src[d] = acc[d] + t;
      }
    }
    // ---- assembly (conv_diff.hpp:149-227), all six faces inner
    double cdiag = 0., ddiag = 0.;
    double cn[6], dn[6], cconst[3] = {0., 0., 0.};
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const int d = q >> 1, o = q & 1;
      const double sgn = o ? 1. : -1.;
      const double Ff = a.F[fm_[d] + (o ? fst[d] : 0)];
      const long long cm = o ? c : c - off[d], cp = o ? c + off[d] : c;
      double vm, vp; int up;   // solver.hpp:223-242, threshold 1e-10
      if (Ff > 1e-10) { vm = 1.; vp = 0.; up = 0; }
      else if (Ff < -1e-10) { vm = 0.; vp = 1.; up = 1; }
      else { vm = 0.5; vp = 0.5; up = 2; }
      const double alpha = 1. / g.h[d];
      const double dm = ((-alpha) * (-muf[q])) * g.area[d];
      const double dp = ((alpha) * (-muf[q])) * g.area[d];
      double cself, cnb, dself, dnb;
      if (o) { cself = vm * Ff; cnb = vp * Ff; dself = dm; dnb = dp; }
      else { cself = vp * Ff; cnb = vm * Ff; dself = dp; dnb = dm; }
      cself *= sgn; cnb *= sgn; dself *= sgn; dnb *= sgn;
      cdiag = q == 0 ? cself : cdiag + cself;
      ddiag = q == 0 ? dself : ddiag + dself;
      cn[q] = cnb; dn[q] = dnb;
      // deferred second-order upwind correction: gradient of the upwind cell (conv_diff.hpp:135, solver.hpp:223-242)
      const long long cu = up == 0 ? cm : cp;
      const double hs = up == 0 ? -0.5 * g.h[d] : 0.5 * g.h[d];
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        const double gu = a.G[n * 3 + d][cu];
        double vc = -(gu * hs);
        if (up == 2) vc = 0.;
        cconst[n] += (vc * Ff) * sgn;
      }
    }
    const double r = a.rho[c];
    const int tmap[6] = {CXM, CXP, CYM, CYP, CZM, CZP};
    double coef[7];
    coef[CD] = (hg_div(cdiag, dvol) + a.co[2]) * r + hg_div(ddiag, dvol);
#pragma unroll
    for (int q = 0; q < 6; ++q) coef[tmap[q]] = hg_div(cn[q], dvol) * r + hg_div(dn[q], dvol);
    // delta form: constant := eqn.Evaluate(prev) in ascending index order (conv_diff.hpp:218); the diffusive constants are
    // sums of +-0 (no Dirichlet face here): + 0.
    const long long offs[7] = {-g.sz, -g.sy, -1, 0, 1, g.sy, g.sz};
    double ev[3];
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      const double uconst = a.co[0] * a.tp[n][c] + a.co[1] * a.tc[n][c];
      double e = ((hg_div(cconst[n], dvol) + uconst) * r + 0.) - src[n];
      const double* __restrict__ pv = a.prev[n];
#pragma unroll
      for (int t = 0; t < 7; ++t) e += pv[c + offs[t]] * coef[t];
      ev[n] = e;
    }
    coef[CD] /= a.relax;   // conv_diff.hpp:221
    {
      double csum = 0.;
#pragma unroll
      for (int t = 0; t < 7; ++t) csum += coef[t];
      double s3 = 0.;
      for (int n = 0; n < 3; ++n) s3 += csum;   // fluid.hpp:878-882
      a.coeffsum[c] = s3 / 3.;
    }
    double* const tl = tile + ty * FT_PITCH + tx;
#pragma unroll
    for (int t = 0; t < 7; ++t) tl[t * FT_K * FT_PITCH] = coef[t];
#pragma unroll
    for (int n = 0; n < 3; ++n) tl[(7 + n) * FT_K * FT_PITCH] = ev[n];
  }
  __syncthreads();
  ft_store_diagonals<10>(tile, act, a.out, g.n[0], g.n[2], i0, k0, [&](int ii, int kk) { return shidx(g, ii, j, kk); });
}

// ------------------------------------------------------------------------------------- K_fstar + K_prhs + row packing
struct FcArgs {
  const double* us[3]; const double* gp[3]; const double* fcr[3]; const double* force[3];
  const double* pprev; const double* dc; double rc; double meshvel[3];
  double* Fs;          // face field, natural
  double* co[5];       // CO5 arrays: constant, diagonal, x+, y+, z+ coefficient (base pointers of the arrays)
  Co5 co5;
};
// Rhie-Chow flux of an inner face between cm and cp in direction d (fluid.hpp:903-940); dfc = d_c at the face
DV double ft_fstar(const FcArgs& a, const Geo& g, int d, long long cm, long long cp, double dfc, const HgDiv& dh) {
  const double A = g.area[d];
  const double mv = a.meshvel[d] * A;
  const double ffu = ft_avg(a.us[d][cm], a.us[d][cp]);
  const double vfi = ffu * A;
  const double fgp = ft_avg(a.gp[d][cm], a.gp[d][cp]);
  const double ffr = ft_avg(a.fcr[d][cm], a.fcr[d][cp]);
  const double ffe = ft_avg(a.force[d][cm], a.force[d][cp]);
  const double wide = (fgp - ffr) * A;
  const double compact = hg_div(a.pprev[cp] - a.pprev[cm], dh) * A - ffe * A;
  return (vfi + a.rc * (wide - compact) / dfc + 0) - mv;
}
__global__ void __launch_bounds__(FT_THREADS, 4) k_fc_flux_rows(Geo g, const unsigned char* __restrict__ slow, FcArgs a) {
  __shared__ double tile[5 * FT_K * FT_PITCH];
  __shared__ unsigned act[FT_K];
  FT_PROLOG(g)
  if (in) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { ft_prefetch(a.us[d] + c); ft_prefetch(a.gp[d] + c); ft_prefetch(a.fcr[d] + c); ft_prefetch(a.force[d] + c); }
    ft_prefetch(a.pprev + c); ft_prefetch(a.dc + c);
  }
  const bool fast = in && !slow[c];
  const unsigned bal = __ballot_sync(0xffffffffu, fast);
  if (tx == 0) act[ty] = bal;
  if (fast) {
    const long long off[3] = {1, g.sy, g.sz};
    const double dcc = a.dc[c];
    double fl[6], cfm[3], cfp[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const HgDiv dh = hg_div_prepare(g.h[d]);
      const long long cm = c - off[d], cp = c + off[d];
      const double dfm = ft_avg(a.dc[cm], dcc), dfp = ft_avg(dcc, a.dc[cp]);
      fl[2 * d] = ft_fstar(a, g, d, cm, c, dfm, dh);
      fl[2 * d + 1] = ft_fstar(a, g, d, c, cp, dfp, dh);
      // face coefficient c_f = A / (h d_f) (fluid.hpp:957-964)
      { const double coeff = -g.area[d] / (g.h[d] * dfm); cfm[d] = -coeff; }
      { const double coeff = -g.area[d] / (g.h[d] * dfp); cfp[d] = -coeff; }
      a.Fs[fidx(g, d, i, j, k)] = fl[2 * d];
      // top plane of a z-slab below another slab: the cell also owns its plus z-face (as the generic kernel's last cell)
      if (d == 2 && k == g.n[2] - 1) a.Fs[fidx(g, 2, i, j, k + 1)] = fl[5];
    }
    double diag = cfm[0] + cfp[0]; diag = diag + cfm[1]; diag = diag + cfp[1]; diag = diag + cfm[2]; diag = diag + cfp[2];
    double cst = 0.;
#pragma unroll
    for (int q = 0; q < 6; ++q) cst += fl[q] * ((q & 1) ? 1. : -1.);
    const double rhs = cst + -(0. * g.vol);
    double* const tl = tile + ty * FT_PITCH + tx;
    tl[0] = rhs; tl[FT_K * FT_PITCH] = diag;
#pragma unroll
    for (int d = 0; d < 3; ++d) tl[(2 + d) * FT_K * FT_PITCH] = cfp[d];
  }
  __syncthreads();
  const Co5 co = a.co5;
  ft_store_diagonals<5>(tile, act, a.co, g.n[0], g.n[2], i0, k0, [&](int ii, int kk) { return gt_co5_index(co, 0, ii, j, kk); });
}
// the generic K_prhs for the listed cells, rows written in the CO5 layout
__global__ void __launch_bounds__(256, 4) k_prhs_co5(Geo g, const double* __restrict__ Fs, const double* __restrict__ dc, double* __restrict__ CO, Co5 co) {
  CELL_LOOP_PROLOG(g)
  double rp, cf[3], dg;
  prhs_cell<3>(g, Fs, dc, i, j, k, c, true, rp, cf, dg);
  CO[gt_co5_index(co, 0, i, j, k)] = rp;
  CO[gt_co5_index(co, 1, i, j, k)] = dg;
#pragma unroll
  for (int d = 0; d < 3; ++d) CO[gt_co5_index(co, 2 + d, i, j, k)] = cf[d];
}

// ------------------------------------------------------------------------------------------------ K_correct
__global__ void __launch_bounds__(FT_THREADS, 4) k_fd_correct(Geo g, const unsigned char* __restrict__ slow, CorrArgs a) {
  __shared__ double smax[FT_K];
  FT_PROLOG(g)
  if (in) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { ft_prefetch(a.u[d] + c); ft_prefetch(a.Fs + fidx(g, d, i, j, k)); if (a.resid) ft_prefetch(a.uprev[d] + c); }
    ft_prefetch(a.pc + c); ft_prefetch(a.dc + c);
  }
  double rs = 0.;   // |u_prev - u_new| of this cell (CalcDiff, solver.hpp:804-813)
  if (in && !slow[c]) {
  const HgDiv dvol = hg_div_prepare(g.vol);
  const long long off[3] = {1, g.sy, g.sz};
  const double pcc = a.pc[c], dcc = a.dc[c];
  double sq = 0.;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double pm = a.pc[c - off[d]], pp = a.pc[c + off[d]];
    const double gpc = ft_grad(pm, pcc, pp, g.area[d] * -1., g.area[d] * 1., dvol);
    const double un = a.u[d][c] + gpc / (-dcc);
    a.u[d][c] = un;
    if (a.resid) { const double e = a.uprev[d][c] - un; sq += e * e; }
    // minus face: F = F* + c_f (p'_m - p'_p) (fluid.hpp:1053-1056)
    const long long fx = fidx(g, d, i, j, k);
    double r = a.Fs[fx];
    const double dfc = ft_avg(a.dc[c - off[d]], dcc);
    const double coeff = -g.area[d] / (g.h[d] * dfc);
    const double cf = -coeff;
    r += pm * cf;
    r += pcc * (-cf);
    a.F[fx] = r;
    if (d == 2 && k == g.n[2] - 1) {   // top plane of a z-slab below another slab: the plus z-face too
      const long long fp_ = fidx(g, 2, i, j, k + 1);
      double rp = a.Fs[fp_];
      const double dfp = ft_avg(dcc, a.dc[c + off[2]]);
      const double coeffp = -g.area[2] / (g.h[2] * dfp);
      const double cfp = -coeffp;
      rp += pcc * cfp;
      rp += pp * (-cfp);
      a.F[fp_] = rp;
    }
  }
  rs = sqrt(sq);
  if (!(rs == rs)) rs = 0.;   // std::max(res, NaN) keeps res
  }
  if (a.resid) {
    rs = warp_max(rs);
    if (tx == 0) smax[ty] = rs;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m = smax[0];
#pragma unroll
      for (int q = 1; q < FT_K; ++q) m = m < smax[q] ? smax[q] : m;
      if (m > 0.) atomic_max_nonneg(a.resid, m);
    }
  }
}

// ------------------------------------------------------------------------------------------------ K_advect
__global__ void __launch_bounds__(FT_THREADS, 5) k_fe_advect(Geo g, const unsigned char* __restrict__ slow, const double* __restrict__ u,
                                                             const double* __restrict__ F, double dt, int num_stages, int stage,
                                                             double* __restrict__ out) {
  FT_PROLOG(g)
  if (!in || slow[c]) return;
  const HgDiv dvol = hg_div_prepare(g.vol);
  const long long off[3] = {1, g.sy, g.sz};
  const long long fst[3] = {1, g.n[0], (long long)g.n[0] * g.n[1]};
  const double uc = u[c];
  double fsum = 0.;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (d % num_stages != stage) continue;
    const double aneg = g.area[d] * -1., apos = g.area[d] * 1.;
    const double um2 = u[c - 2 * off[d]], um1 = u[c - off[d]], up1 = u[c + off[d]], up2 = u[c + 2 * off[d]];
    const long long fx = fidx(g, d, i, j, k);
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      const double Ff = F[fx + (o ? fst[d] : 0)];
      // cells P (cm) and E (cp) of the face and their outer neighbours
      const double uPm = o ? um1 : um2, uP = o ? uc : um1, uE = o ? up1 : uc, uEp = o ? up2 : up1;
      const double du = uE - uP;
      double fu;
      if (Ff > 1e-8) {
        const double gP = ft_grad(uPm, uP, uE, aneg, apos, dvol);
        const double pq = -4. * (gP * (-0.5 * g.h[d])) - du;
        fu = uP + 0.5 * superbee(du, pq);
      } else if (Ff < -1e-8) {
        const double gE = ft_grad(uP, uE, uEp, aneg, apos, dvol);
        const double pq = 4. * (gE * (0.5 * g.h[d])) - du;
        fu = uE - 0.5 * superbee(du, pq);
      } else fu = 0.5 * (uP + uE);
      fsum += fu * Ff * (o ? 1. : -1.);
    }
  }
  out[c] = uc + -dt / g.vol * fsum;
}

// ------------------------------------------------------------------------ solver results back to the natural layout
// curr = prev + corr (conv_diff.hpp:246-248), corr hyperplane-major: the diagonals of the tile are read as contiguous runs
template <int NCOMP>
__global__ void __launch_bounds__(FT_THREADS, 6) k_ft_apply_corr(Geo g, CP3 prev, CP3 X, P3 curr) {
  __shared__ double tile[NCOMP * FT_K * FT_PITCH];
  FT_PROLOG(g)
  {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, p = lane & 7, sub = lane >> 3;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int D = pass * 32 + w * 4 + sub;
      const int il = (D > FT_K - 1 ? D - (FT_K - 1) : 0) + p, kl = D - il;
      if (D < 32 + FT_K - 1 && il <= 31 && kl >= 0 && i0 + il < g.n[0] && k0 + kl < g.n[2]) {
        const long long cs = shidx(g, i0 + il, j, k0 + kl);
#pragma unroll
        for (int n = 0; n < NCOMP; ++n) tile[(n * FT_K + kl) * FT_PITCH + il] = X.p[n][cs];
      }
    }
  }
  __syncthreads();
  if (!in) return;
#pragma unroll
  for (int n = 0; n < NCOMP; ++n) curr.p[n][c] = prev.p[n][c] + tile[(n * FT_K + ty) * FT_PITCH + tx];
}
// p' back to the natural layout, p_curr = p_prev + alpha_p p' (fluid.hpp:1035-1038)
__global__ void __launch_bounds__(FT_THREADS, 6) k_ft_pcorr(Geo g, const double* __restrict__ PP, const double* __restrict__ pprev, double alpha,
                                                            double* __restrict__ pc, double* __restrict__ pcurr) {
  __shared__ double tile[FT_K * FT_PITCH];
  FT_PROLOG(g)
  {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, p = lane & 7, sub = lane >> 3;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int D = pass * 32 + w * 4 + sub;
      const int il = (D > FT_K - 1 ? D - (FT_K - 1) : 0) + p, kl = D - il;
      if (D < 32 + FT_K - 1 && il <= 31 && kl >= 0 && i0 + il < g.n[0] && k0 + kl < g.n[2])
        tile[kl * FT_PITCH + il] = PP[shidx(g, i0 + il, j, k0 + kl)];
    }
  }
  __syncthreads();
  if (!in) return;
  const double v = tile[ty * FT_PITCH + tx];
  pc[c] = v;
  pcurr[c] = pprev[c] + alpha * v;
}
