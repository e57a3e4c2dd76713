/* hydro_oracle.c -- TEST INFRASTRUCTURE ONLY (see hydro_oracle.h).
 *
 * CPU restatement of divfree/hydro's per-time-step path on a uniform Cartesian
 * mesh.  Plain C99, serial, fp64, SoA arrays in the reference's raw index order.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/source/hydro2dmpi unless noted).  The reference builds every
 * matrix row as a sorted `Expression` (linear.hpp:46-262); this file keeps the
 * same floating-point association (which operand pairs are added first) but
 * stores rows as 7 coefficient arrays in the order z-,y-,x-,diag,x+,y+,z+ (the
 * ascending-raw-index order Expression::Evaluate/CoeffSum iterate in).
 *
 * Geometry is closed form (h = (B-A)/N) instead of the reference's per-cell
 * tables computed from node coordinates (mesh3d.hpp:264-460); the two agree to
 * rounding, and exactly when N is a power of two on a unit box.
 *
 * Not restated (the product rejects the same options): geometric force averaging, the carrier-velocity variant of the
 * phase slip, chemistry/radiation, compressibility, PIC advection, periodic meshes.
 */
#include "hydro_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { L_TC = 0, L_TP = 1, L_IC = 2, L_IP = 3 }; /* solver.hpp:762 Layers */
enum { FT_INNER = 0, FT_BOUND = 1, FT_EXCL = 2 };
enum { K_NONE = 0, K_NEUMANN0, K_EXTRAP, K_VEL, K_TEMP, K_PD };
enum { CZM = 0, CYM = 1, CXM = 2, CD = 3, CXP = 4, CYP = 5, CZP = 6 };

struct ho_state {
  hg_config cfg;
  int dim, n[3];
  size_t nc, nfd[3], foff[3], nf;
  double h[3], vol, area[3], lb[3];
  unsigned char* cexcl;
  unsigned char* ftype;
  signed char* fside;     /* boundary faces: index into bcvel (0..5 sides, 6 box) */
  unsigned char* ftdir;   /* temperature: 1 = Dirichlet (heat box) */
  double bcvel[7][3];
  double* ma[3][7]; double* mr[3];  /* momentum equations of the iteration (GetVelocityEquations), kept for SIMPLER */
  double* fev[3];                   /* fc_evaluated */
  double* slipv[HG_MAX_PHASES][3]; /* v_fc_velocity_slip */
  double* fslip[HG_MAX_PHASES];    /* v_ff_volume_flux_slip */
  int any_slip;
  double* outvel[3];      /* velocity of the outlet faces (OutletAuto, fluid.hpp:318-336), face-field sized, zero at construction */
  int any_outlet;
  int bckind[7];
  long pfix_cell;
  double *u[4][3], *p[4], *F[4];
  double* pd[HG_MAX_PHASES][4];
  double* pd_inlet[HG_MAX_PHASES]; /* Dirichlet value on inlet faces (hydro2d.hpp:567-575) */
  double* T[4];
  double *vf[HG_MAX_PHASES], *rho_raw, *rho, *mu, *kc;
  double *force[3], *stforce[3], *qvol, *qmass, *tsrc;
  /* FluidSimple buffers (fluid.hpp:466-483) */
  double *ffe[3], *fcr[3], *ffr[3];      /* ff_ext_force_, fc_ext_force_restored_, ff_ext_force_restored_ */
  double *muf, *ffp, *gp[3], *fgp[3];    /* ff_kinematic_viscosity_, ff_pressure_, fc_pressure_grad_, ff_pressure_grad_ */
  double *fs[3];                         /* fc_force_ (momentum source) */
  double *dc, *dfc, *ffu[3], *Fs, *cf;   /* fc_diag_coeff_, ff_diag_coeff_, ff_velocity_asterisk_, flux*, face coeff */
  double *pc, *a[7], *rhs, *corr;        /* pressure correction, matrix, constants */
  double *w1, *w2[3], *wf, *wf3[3], *kf; /* scratch */
  double time_fluid, time_adv, time_heat, dt, dt_adv;
  int iter_count, heat_iter;
  int last_sweeps_total; double last_diff;
  int nres; double res_hist[4096];
  hg_step_stats stat;
  double meshpos[3];
  double stat_cx[HG_MAX_PHASES]; int stat_cx_set[HG_MAX_PHASES];   /* P_double["stat_cx_<i>"] of the previous CalcStat */
  char err[256];
};

static char g_err[256];

/* ---------------------------------------------------------------- indexing */
/* cells: BlockGeneric::GetIdx mesh.hpp:552-561; faces: BlockFaces::GetIdx mesh.hpp:698-705 */
static inline size_t cidx(const struct ho_state* s, int i, int j, int k) {
  return (size_t)i + (size_t)s->n[0] * ((size_t)j + (size_t)s->n[1] * (size_t)k);
}
static inline size_t fidx(const struct ho_state* s, int d, int i, int j, int k) {
  size_t ex = (size_t)s->n[0] + (d == 0), ey = (size_t)s->n[1] + (d == 1);
  return s->foff[d] + (size_t)i + ex * ((size_t)j + ey * (size_t)k);
}
/* GetNeighbourFace(cell, q), q = 0..2*dim-1: x-,x+,y-,y+,z-,z+ (mesh3d.hpp:290-296) */
static inline size_t nface(const struct ho_state* s, int i, int j, int k, int q) {
  int d = q >> 1, o = q & 1;
  return fidx(s, d, i + (d == 0 ? o : 0), j + (d == 1 ? o : 0), k + (d == 2 ? o : 0));
}
static inline void face_midx(const struct ho_state* s, size_t f, int* d, int* i, int* j, int* k) {
  int dd = 0;
  while (dd + 1 < s->dim && f >= s->foff[dd + 1]) ++dd;
  size_t r = f - s->foff[dd];
  size_t ex = (size_t)s->n[0] + (dd == 0), ey = (size_t)s->n[1] + (dd == 1);
  *d = dd; *i = (int)(r % ex); r /= ex; *j = (int)(r % ey); *k = (int)(r / ey);
}
/* the two cells of a face: cm = midx - e_d, cp = midx (mesh3d.hpp:344-347); -1 = none */
static inline void face_cells(const struct ho_state* s, int d, int i, int j, int k, long* cm, long* cp) {
  int im = i - (d == 0), jm = j - (d == 1), km = k - (d == 2);
  *cm = (im >= 0 && jm >= 0 && km >= 0) ? (long)cidx(s, im, jm, km) : -1;
  *cp = (i < s->n[0] && j < s->n[1] && k < s->n[2]) ? (long)cidx(s, i, j, k) : -1;
  if (*cm >= 0 && s->cexcl[*cm]) *cm = -1; /* ExcludeCells mesh3d.hpp:183-193 */
  if (*cp >= 0 && s->cexcl[*cp]) *cp = -1;
}
static inline void cell_center(const struct ho_state* s, int i, int j, int k, double x[3]) {
  x[0] = s->lb[0] + (i + 0.5) * s->h[0];
  x[1] = s->lb[1] + (j + 0.5) * s->h[1];
  x[2] = s->dim > 2 ? s->lb[2] + (k + 0.5) * s->h[2] : 0.;
}

static double* dalloc(size_t n) {
  double* p = (double*)calloc(n ? n : 1, sizeof(double));
  return p;
}

/* ------------------------------------------------ Interpolate cell -> face */
/* Dirichlet velocity of a boundary face: wall / inlet value of its side, or the outlet face's own velocity */
static double bc_velocity(const struct ho_state* s, size_t f, int comp) {
  const int side = (int)s->fside[f];
  if (side < 6 && s->bckind[side] == HG_BC_OUTLET) return s->outvel[comp][f];
  return s->bcvel[side][comp];
}

/* solver.hpp:392-470.  kind selects the MapFace of conditions:
 *   K_NONE      empty map (boundary faces stay 0)           fluid.hpp:886
 *   K_NEUMANN0  ConditionFaceDerivativeFixed(0)             fluid.hpp:723-730
 *   K_EXTRAP    ConditionFaceExtrapolation                  fluid.hpp:705,727
 *   K_VEL       ConditionFaceValueFixed(wall velocity[comp]) fluid.hpp:702-716
 *   K_TEMP      heat box Dirichlet / zero derivative        hydro2d.hpp:664-676
 *   K_PD        partial density: inlet value / zero deriv.  hydro2d.hpp:563-583 */
static void interp(const struct ho_state* s, const double* u, int kind, int comp, double* res) {
  const int dim = s->dim;
  memset(res, 0, s->nf * sizeof(double)); /* "Valid value essential for extrapolation" solver.hpp:402 */
  for (int d = 0; d < dim; ++d) {
    int ex = s->n[0] + (d == 0), ey = s->n[1] + (d == 1), ez = s->n[2] + (d == 2);
    for (int k = 0; k < ez; ++k) for (int j = 0; j < ey; ++j) for (int i = 0; i < ex; ++i) {
      size_t f = fidx(s, d, i, j, k);
      if (s->ftype[f] != FT_INNER) continue;
      long cm, cp; face_cells(s, d, i, j, k, &cm, &cp);
      res[f] = u[cm] * (1. - 0.5) + u[cp] * 0.5; /* solver.hpp:425-426 */
    }
  }
  if (kind == K_NONE) return;
  /* boundary faces in ascending face index (std::map order, solver.hpp:431) */
  for (int d = 0; d < dim; ++d) {
    int ex = s->n[0] + (d == 0), ey = s->n[1] + (d == 1), ez = s->n[2] + (d == 2);
    for (int k = 0; k < ez; ++k) for (int j = 0; j < ey; ++j) for (int i = 0; i < ex; ++i) {
      size_t f = fidx(s, d, i, j, k);
      if (s->ftype[f] != FT_BOUND) continue;
      long cm, cp; face_cells(s, d, i, j, k, &cm, &cp);
      int id = (cm >= 0) ? 0 : 1;          /* GetValidNeighbourCellId mesh.hpp:422-429 */
      long cc = id == 0 ? cm : cp;
      int dirichlet = 0; double val = 0.;
      int k2 = kind;
      if (kind == K_VEL) { dirichlet = 1; val = bc_velocity(s, f, comp); }
      else if (kind == K_TEMP) { if (s->ftdir[f]) { dirichlet = 1; val = s->cfg.heat_box_temperature; } else k2 = K_NEUMANN0; }
      else if (kind == K_PD) { if (s->bckind[(int)s->fside[f]] == HG_BC_INLET) { dirichlet = 1; val = s->pd_inlet[comp][f]; } else k2 = K_NEUMANN0; }
      if (dirichlet) { res[f] = val; continue; }
      if (k2 == K_NEUMANN0) {
        double factor = (id == 0 ? 1. : -1.);
        double alpha = (0.5 * s->h[d]) * factor;
        res[f] = u[cc] + 0. * alpha;       /* solver.hpp:441-445 */
        continue;
      }
      /* extrapolation, solver.hpp:446-464 */
      {
        double factor = (id == 0 ? 1. : -1.);
        double normal[3] = {0., 0., 0.};
        normal[d] = (s->area[d] / s->area[d]) * factor;
        double dist = 0.5 * s->h[d];
        double nom = u[cc] / dist;
        double den = 1. / dist;
        int ci = i - (id == 0 && d == 0), cj = j - (id == 0 && d == 1), ck = k - (id == 0 && d == 2);
        for (int q = 0; q < 2 * dim; ++q) {
          size_t nf_ = nface(s, ci, cj, ck, q);
          int qd = q >> 1;
          double so[3] = {0., 0., 0.};
          so[qd] = s->area[qd] * ((q & 1) ? 1. : -1.);
          double dot = 0.;
          for (int c = 0; c < dim; ++c) dot += so[c] * normal[c];
          if (nf_ == f) den -= dot / s->vol;
          else nom += res[nf_] * dot / s->vol;
        }
        res[f] = nom / den;
      }
    }
  }
}

/* Gradient(FieldFace) solver.hpp:658-677 */
static void gradient(const struct ho_state* s, const double* ff, double* g[3]) {
  for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
    size_t c = cidx(s, i, j, k);
    if (s->cexcl[c]) { for (int d = 0; d < s->dim; ++d) g[d][c] = 0.; continue; }
    for (int d = 0; d < s->dim; ++d) {
      double sum = 0.;
      sum += (s->area[d] * -1.) * ff[nface(s, i, j, k, 2 * d)];
      sum += (s->area[d] * 1.) * ff[nface(s, i, j, k, 2 * d + 1)];
      g[d][c] = sum / s->vol;
    }
  }
}

/* Average(FieldFace) solver.hpp:621-634 and GetSmoothField solver.hpp:636-656 */
static void smooth(const struct ho_state* s, const double* u, int repeat, double* out, double* wf, double* wc) {
  memcpy(out, u, s->nc * sizeof(double));
  for (int r = 0; r < repeat; ++r) {
    interp(s, out, K_NEUMANN0, 0, wf);
    for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
      double sum = 0.;
      for (int q = 0; q < 2 * s->dim; ++q) sum += wf[nface(s, i, j, k, q)];
      wc[cidx(s, i, j, k)] = sum / (double)(2 * s->dim);
    }
    memcpy(out, wc, s->nc * sizeof(double));
  }
}

/* ---------------------------------------------------------- linear solvers */
/* neighbour raw offsets in the order z-,y-,x-,(diag),x+,y+,z+ */
static void nb_offsets(const struct ho_state* s, long off[7]) {
  off[CZM] = -(long)s->n[0] * s->n[1]; off[CYM] = -(long)s->n[0]; off[CXM] = -1; off[CD] = 0;
  off[CXP] = 1; off[CYP] = s->n[0]; off[CZP] = (long)s->n[0] * s->n[1];
}
static inline int nb_exists(const struct ho_state* s, int i, int j, int k, int t) {
  switch (t) {
    case CZM: return s->dim > 2 && k > 0;
    case CYM: return j > 0;
    case CXM: return i > 0;
    case CXP: return i + 1 < s->n[0];
    case CYP: return j + 1 < s->n[1];
    case CZP: return s->dim > 2 && k + 1 < s->n[2];
  }
  return 1;
}

/* LuDecomposition::Solve linear.hpp:533-566 */
static void solve_lu(const struct ho_state* s, double* const a[7], const double* rhs, double* x) {
  long off[7]; nb_offsets(s, off);
  for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
    size_t c = cidx(s, i, j, k);
    double sum = 0;
    for (int t = CZM; t <= CXM; ++t) if (nb_exists(s, i, j, k, t)) sum += a[t][c] * x[(long)c + off[t]];
    x[c] = (-rhs[c] - sum) / a[CD][c];
  }
  for (int k = s->n[2] - 1; k >= 0; --k) for (int j = s->n[1] - 1; j >= 0; --j) for (int i = s->n[0] - 1; i >= 0; --i) {
    size_t c = cidx(s, i, j, k);
    double sum = 0;
    for (int t = CZP; t >= CXP; --t) if (nb_exists(s, i, j, k, t)) sum += a[t][c] * x[(long)c + off[t]];
    x[c] -= sum / a[CD][c];
  }
}

/* GaussSeidel::Solve linear.hpp:685-715 / Jacobi::Solve linear.hpp:750-782 */
static void solve_sor(const struct ho_state* s, int jacobi, double* const a[7], const double* rhs, double* x,
                      double tol, int limit, double omega, int* out_iter, double* out_diff) {
  long off[7]; nb_offsets(s, off);
  double* next = jacobi ? dalloc(s->nc) : NULL;
  double* res = x;
  memset(res, 0, s->nc * sizeof(double));
  size_t iter = 0; double diff = 0.;
  do {
    diff = 0.;
    for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
      size_t c = cidx(s, i, j, k);
      double sum = 0.;
      for (int t = 0; t < 7; ++t) if (t != CD && nb_exists(s, i, j, k, t)) sum += a[t][c] * res[(long)c + off[t]];
      double value = -(rhs[c] + sum) / a[CD][c];
      double corr = value - res[c];
      diff = fmax(diff, fabs(corr));
      if (jacobi) next[c] = res[c] + corr * omega;
      else res[c] += corr * omega;
    }
    if (jacobi) { double* t = res; res = next; next = t; }
  } while (diff > tol && iter++ < (size_t)limit);
  if (jacobi) {
    if (res != x) { memcpy(x, res, s->nc * sizeof(double)); free(res); }
    else free(next);
  }
  *out_iter = (int)iter; *out_diff = diff;
}

/* LuDecompositionRelaxed::Solve linear.hpp:592-650 */
static void solve_lu_relaxed(const struct ho_state* s, double* const a[7], const double* rhs, double* x,
                             double tol, int limit, double relax, int* out_iter, double* out_diff) {
  long off[7]; nb_offsets(s, off);
  double* corr = dalloc(s->nc); double* f = dalloc(s->nc);
  memset(x, 0, s->nc * sizeof(double));
  memcpy(f, rhs, s->nc * sizeof(double));
  size_t iter = 0; double diff = 0.;
  do {
    diff = 0.;
    for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
      size_t c = cidx(s, i, j, k);
      double sum = 0;
      for (int t = CZM; t <= CXM; ++t) if (nb_exists(s, i, j, k, t)) sum += a[t][c] * corr[(long)c + off[t]];
      corr[c] = (-f[c] - sum) / (a[CD][c] + relax);
    }
    for (int k = s->n[2] - 1; k >= 0; --k) for (int j = s->n[1] - 1; j >= 0; --j) for (int i = s->n[0] - 1; i >= 0; --i) {
      size_t c = cidx(s, i, j, k);
      double sum = 0;
      /* the reference reads `res`, not `corr`, here (linear.hpp:632) */
      for (int t = CZP; t >= CXP; --t) if (nb_exists(s, i, j, k, t)) sum += a[t][c] * x[(long)c + off[t]];
      corr[c] -= sum / (a[CD][c] + relax);
    }
    for (size_t c = 0; c < s->nc; ++c) { x[c] += corr[c]; diff = fmax(diff, fabs(corr[c])); }
    for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
      size_t c = cidx(s, i, j, k);
      double r = rhs[c];
      for (int t = 0; t < 7; ++t) if (t == CD || nb_exists(s, i, j, k, t)) r += x[(long)c + off[t]] * a[t][c];
      f[c] = r;
    }
  } while (diff > tol && iter++ < (size_t)limit);
  free(corr); free(f);
  *out_iter = (int)iter; *out_diff = diff;
}

static void solve(struct ho_state* s, int solver, double* const a[7], const double* rhs, double* x, int* it, double* df) {
  *it = 0; *df = 0.;
  const hg_config* c = &s->cfg;
  switch (solver) {
    case HG_LS_LU: memset(x, 0, s->nc * sizeof(double)); solve_lu(s, a, rhs, x); break;
    case HG_LS_LU_RELAXED: solve_lu_relaxed(s, a, rhs, x, c->lu_relaxed_tolerance, c->lu_relaxed_num_iters_limit, c->lu_relaxed_relaxation_factor, it, df); break;
    case HG_LS_GAUSS_SEIDEL: solve_sor(s, 0, a, rhs, x, c->lu_relaxed_tolerance, c->lu_relaxed_num_iters_limit, c->lu_relaxed_relaxation_factor, it, df); break;
    default: solve_sor(s, 1, a, rhs, x, c->lu_relaxed_tolerance, c->lu_relaxed_num_iters_limit, c->lu_relaxed_relaxation_factor, it, df); break;
  }
}

/* GetDerivativeApproxCoeffs solver.hpp:816-861 with args {-2dt,-dt,0}, target 0 */
static void bdf_coeffs(double dt, int second_order, double co[3]) {
  double args[3] = {-2. * dt, -dt, 0.};
  int skip = second_order ? 0 : 1, size = 3 - skip;
  co[0] = co[1] = co[2] = 0.;
  for (int i = 0; i < size; ++i) {
    double denom = 1., numer = 0.;
    for (int j = 0; j < size; ++j) if (j != i) {
      denom *= args[skip + i] - args[skip + j];
      double term = 1.;
      for (int k = 0; k < size; ++k) if (k != i && k != j) term *= 0. - args[skip + k];
      numer += term;
    }
    co[skip + i] = numer / denom;
  }
}

/* ------------------------------------- ConvectionDiffusionScalarImplicit */
/* One MakeIteration (conv_diff.hpp:130-251) for a scalar with layers fl[4].
 * kind/comp: boundary condition selector (K_VEL comp n, or K_TEMP).
 * rho == NULL means the scaling field is identically 1 (heat.hpp:31,68). */
static void convdiff_iteration(struct ho_state* s, double* fl[4], int kind, int comp, const double* rho,
                               const double* muf, const double* src, const double* F, double relax,
                               int second_order, double dt, int solver, double* coeffsum) {
  const int dim = s->dim;
  double* prev = fl[L_IP]; double* curr = fl[L_IC];
  memcpy(prev, curr, s->nc * sizeof(double));                       /* :133 */
  interp(s, prev, kind, comp, s->wf);
  double* g[3] = {s->w2[0], s->w2[1], s->w2[2]};
  gradient(s, s->wf, g);                                            /* :135 */
  double co[3]; bdf_coeffs(dt, second_order, co);                   /* :203-204 */
  long off[7]; nb_offsets(s, off);
  static const int tmap[6] = {CXM, CXP, CYM, CYP, CZM, CZP};
  for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
    size_t c = cidx(s, i, j, k);
    for (int t = 0; t < 7; ++t) s->a[t][c] = 0.;
    if (s->cexcl[c]) { s->a[CD][c] = 1.; s->rhs[c] = 0.; if (coeffsum) coeffsum[c] = 1.; continue; } /* :222-226 */
    double cdiag = 0., ddiag = 0., cconst = 0., dconst = 0.;
    int have_c = 0, have_d = 0;
    double cn[6] = {0, 0, 0, 0, 0, 0}, dn[6] = {0, 0, 0, 0, 0, 0};
    int present[6] = {0, 0, 0, 0, 0, 0};
    for (int q = 0; q < 2 * dim; ++q) {
      int d = q >> 1;
      double sgn = (q & 1) ? 1. : -1.;                               /* GetOutwardFactor */
      size_t f = nface(s, i, j, k, q);
      int fi = i + (q == 1), fj = j + (q == 3), fk = k + (q == 5);
      if (s->ftype[f] == FT_EXCL) continue;                          /* :154,171: Expr() */
      double Ff = F[f];
      if (s->ftype[f] == FT_INNER) {
        long cm, cp; face_cells(s, d, fi, fj, fk, &cm, &cp);
        double vm, vp, vc = 0.;                                      /* solver.hpp:223-242 */
        if (Ff > 1e-10) { vm = 1.; vp = 0.; vc = -(g[d][cm] * (-0.5 * s->h[d])); }
        else if (Ff < -1e-10) { vm = 0.; vp = 1.; vc = -(g[d][cp] * (0.5 * s->h[d])); }
        else { vm = 0.5; vp = 0.5; }
        double alpha = 1. / s->h[d];                                 /* solver.hpp:296-300 */
        double dm = ((-alpha) * (-muf[f])) * s->area[d];             /* conv_diff.hpp:178-179 */
        double dp = ((alpha) * (-muf[f])) * s->area[d];
        double cself, cnb, dself, dnb;
        if (q & 1) { cself = vm * Ff; cnb = vp * Ff; dself = dm; dnb = dp; }   /* this cell is cm */
        else { cself = vp * Ff; cnb = vm * Ff; dself = dp; dnb = dm; }         /* this cell is cp */
        cself *= sgn; cnb *= sgn; dself *= sgn; dnb *= sgn;
        cdiag = have_c ? cdiag + cself : cself; have_c = 1;
        ddiag = have_d ? ddiag + dself : dself; have_d = 1;
        cn[q] = cnb; dn[q] = dnb; present[q] = 1;
        cconst += (vc * Ff) * sgn;
        dconst += ((0. * (-muf[f])) * s->area[d]) * sgn;
      } else {
        /* boundary face: solver.hpp:258-280 and 318-340 */
        int id = (q & 1) ? 0 : 1;
        double factor = (id == 0 ? 1. : -1.);
        int dirichlet = 0; double val = 0.;
        if (kind == K_VEL) { dirichlet = 1; val = bc_velocity(s, f, comp); }
        else if (kind == K_TEMP && s->ftdir[f]) { dirichlet = 1; val = s->cfg.heat_box_temperature; }
        if (dirichlet) {
          cconst += (val * Ff) * sgn;
          double alpha = 1. / (0.5 * s->h[d]) * factor;
          double dself = (((-alpha) * (-muf[f])) * s->area[d]) * sgn;
          ddiag = have_d ? ddiag + dself : dself; have_d = 1;
          dconst += (((alpha * val) * (-muf[f])) * s->area[d]) * sgn;
        } else {
          double alpha = (0.5 * s->h[d]) * factor;
          double cself = (1. * Ff) * sgn;
          cdiag = have_c ? cdiag + cself : cself; have_c = 1;
          cconst += ((alpha * 0.) * Ff) * sgn;
          dconst += ((0. * (-muf[f])) * s->area[d]) * sgn;
        }
      }
    }
    double r = rho ? rho[c] : 1.;
    /* eqn = (unsteady + cflux_sum / V) * rho + dflux_sum / V - Expr(source)   :212-215 */
    double uconst = co[0] * fl[L_TP][c] + co[1] * fl[L_TC][c];       /* :208-210 */
    double diag = ((have_c ? cdiag / s->vol : 0.) + co[2]) * r + (have_d ? ddiag / s->vol : 0.);
    double cst = ((cconst / s->vol + uconst) * r + dconst / s->vol) - src[c];
    for (int q = 0; q < 2 * dim; ++q) if (present[q]) s->a[tmap[q]][c] = (cn[q] / s->vol) * r + dn[q] / s->vol;
    s->a[CD][c] = diag;
    /* delta form: constant := eqn.Evaluate(fc_prev)   :218 */
    double ev = cst;
    for (int t = 0; t < 7; ++t) {
      if (t == CD) { ev += prev[c] * s->a[CD][c]; continue; }
      int q = -1; for (int qq = 0; qq < 6; ++qq) if (tmap[qq] == t) q = qq;
      if (q < 2 * dim && present[q]) ev += prev[(long)c + off[t]] * s->a[t][c];
    }
    s->rhs[c] = ev;
    s->a[CD][c] /= relax;                                            /* :221 */
    if (coeffsum) {                                                  /* Expression::CoeffSum linear.hpp:136-142 */
      double cs = 0.;
      for (int t = 0; t < 7; ++t) {
        if (t == CD) { cs += s->a[CD][c]; continue; }
        int q = -1; for (int qq = 0; qq < 6; ++qq) if (tmap[qq] == t) q = qq;
        if (q < 2 * dim && present[q]) cs += s->a[t][c];
      }
      coeffsum[c] = cs;
    }
  }
  if (kind == K_VEL && s->cfg.simpler) {   /* the equations stay available (GetVelocityEquations, conv_diff.hpp) */
    for (int t = 0; t < 7; ++t) memcpy(s->ma[comp][t], s->a[t], s->nc * sizeof(double));
    memcpy(s->mr[comp], s->rhs, s->nc * sizeof(double));
  }
  int it; double df;
  solve(s, solver, s->a, s->rhs, s->corr, &it, &df);                 /* :245 */
  for (size_t c = 0; c < s->nc; ++c) curr[c] = prev[c] + s->corr[c]; /* :246-248 */
}

/* --------------------------------------------------------------- FluidSimple */
static int has_nan(const double* a, size_t n) {
  for (size_t i = 0; i < n; ++i) if (!(a[i] * 0. == 0.)) return 1;   /* IsNan solver.hpp:17-30 */
  return 0;
}

int ho_fluid_start_step(ho_handle s) {                               /* fluid.hpp:793-812 */
  s->iter_count = 0;
  if (has_nan(s->p[L_TC], s->nc)) { snprintf(s->err, sizeof s->err, "NaN initial pressure"); return HG_ERR_NAN; }
  double ge = s->cfg.guess_extrapolation;
  for (int d = 0; d < s->dim; ++d) {                                 /* conv_diff.hpp:118-129 */
    if (has_nan(s->u[L_TC][d], s->nc)) { snprintf(s->err, sizeof s->err, "NaN initial field"); return HG_ERR_NAN; }
    for (size_t c = 0; c < s->nc; ++c)
      s->u[L_IC][d][c] = s->u[L_TC][d][c] + (s->u[L_TC][d][c] - s->u[L_TP][d][c]) * ge;
  }
  for (size_t c = 0; c < s->nc; ++c) s->p[L_IC][c] = s->p[L_TC][c] + (s->p[L_TC][c] - s->p[L_TP][c]) * ge;
  for (size_t f = 0; f < s->nf; ++f) s->F[L_IC][f] = s->F[L_TC][f] + (s->F[L_TC][f] - s->F[L_TP][f]) * ge;
  return 0;
}

/* UpdateOutletBaseConditions, fluid.hpp:542-600: every outlet face takes the velocity of its cell (iter_curr); a uniform
 * normal velocity is added so that the outlet flux equals the inlet flux (boundary faces in ascending face index) */
static void update_outlet(struct ho_state* s) {
  if (!s->any_outlet) return;
  const int dim = s->dim;
  double inlet = 0., outlet = 0., area = 0.;
  for (int pass = 0; pass < 2; ++pass) {
    const double corr = pass ? (inlet - outlet) / area : 0.;
    for (int d = 0; d < dim; ++d) {
      int ex = s->n[0] + (d == 0), ey = s->n[1] + (d == 1), ez = s->n[2] + (d == 2);
      for (int k = 0; k < ez; ++k) for (int j = 0; j < ey; ++j) for (int i = 0; i < ex; ++i) {
        size_t f = fidx(s, d, i, j, k);
        if (s->ftype[f] != FT_BOUND) continue;
        const int side = (int)s->fside[f];
        if (side >= 6) continue;   /* faces of the rigid box are walls */
        long cm, cp; face_cells(s, d, i, j, k, &cm, &cp);
        const int id = (cm >= 0) ? 0 : 1;
        const long cc = id == 0 ? cm : cp;
        if (s->bckind[side] == HG_BC_OUTLET) {
          const double factor = (id == 0 ? 1. : -1.);
          if (!pass) {
            double dot = 0.;
            for (int c = 0; c < dim; ++c) { s->outvel[c][f] = s->u[L_IC][c][cc]; dot += s->outvel[c][f] * (c == d ? s->area[d] : 0.); }
            outlet += dot * factor;
            area += s->area[d];
          } else {
            for (int c = 0; c < dim; ++c) s->outvel[c][f] = s->outvel[c][f] + ((c == d ? s->area[d] / s->area[d] : 0. / s->area[d]) * corr) * factor;
          }
        } else if (s->bckind[side] == HG_BC_INLET && !pass) {
          const double factor = (id == 0 ? -1. : 1.);
          double dot = 0.;
          for (int c = 0; c < dim; ++c) dot += s->bcvel[side][c] * (c == d ? s->area[d] : 0.);
          inlet += dot * factor;
        }
      }
    }
    /* + sum of volume_source * V over the cells: zero sources */
  }
}

int ho_fluid_make_iteration(ho_handle s) {                           /* fluid.hpp:814-1158 */
  const int dim = s->dim;
  const hg_config* cfg = &s->cfg;
  double* pprev = s->p[L_IP]; double* pcurr = s->p[L_IC];
  memcpy(pprev, pcurr, s->nc * sizeof(double));                      /* :815-818 */
  memcpy(s->F[L_IP], s->F[L_IC], s->nf * sizeof(double));
  const double* Fprev = s->F[L_IP];
  update_outlet(s);                                                  /* :820-821 */

  /* CalcExtForce fluid.hpp:602-631 */
  for (int d = 0; d < dim; ++d) interp(s, s->force[d], K_NEUMANN0, 0, s->ffe[d]);
  for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
    size_t c = cidx(s, i, j, k);
    for (int d = 0; d < dim; ++d) {
      double sum = 0.;
      for (int o = 0; o < 2; ++o) {
        size_t f = nface(s, i, j, k, 2 * d + o);
        sum += (s->area[d] * s->ffe[d][f]) * (0.5 * s->h[d]);
      }
      s->fcr[d][c] = sum / s->vol;
    }
  }
  for (int d = 0; d < dim; ++d) interp(s, s->fcr[d], K_NEUMANN0, 0, s->ffr[d]);
  /* CalcKinematicViscosity fluid.hpp:632-641 */
  interp(s, s->mu, K_NEUMANN0, 0, s->muf);
  /* pressure gradient :827-832 */
  interp(s, pprev, K_EXTRAP, 0, s->ffp);
  gradient(s, s->ffp, s->gp);
  for (int d = 0; d < dim; ++d) interp(s, s->gp[d], K_NEUMANN0, 0, s->fgp[d]);

  /* explicit viscous term :835-853 */
  for (int d = 0; d < dim; ++d) memset(s->fs[d], 0, s->nc * sizeof(double));
  for (int n = 0; n < dim; ++n) {
    interp(s, s->u[L_IC][n], K_VEL, n, s->wf);
    double* gc[3] = {s->w2[0], s->w2[1], s->w2[2]};
    gradient(s, s->wf, gc);
    for (int d = 0; d < dim; ++d) interp(s, gc[d], K_NEUMANN0, 0, s->wf3[d]);
    for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
      size_t c = cidx(s, i, j, k);
      size_t fm = nface(s, i, j, k, 2 * n), fp = nface(s, i, j, k, 2 * n + 1);
      for (int d = 0; d < dim; ++d) {
        double sum = 0.;
        sum += s->wf3[d][fm] * (s->muf[fm] * (s->area[n] * -1.));
        sum += s->wf3[d][fp] * (s->muf[fp] * (s->area[n] * 1.));
        s->fs[d][c] += sum / s->vol;
      }
    }
  }
  /* append to force :857-870 */
  for (size_t c = 0; c < s->nc; ++c) {
    double sc = s->rho[c] * s->qvol[c] - s->qmass[c];
    for (int d = 0; d < dim; ++d) {
      double t = ((s->gp[d][c] * (-1.) + s->fcr[d][c]) + s->stforce[d][c]) + s->u[L_IC][d][c] * sc;
      s->fs[d][c] += t;
    }
  }

  /* conv_diff_solver_->MakeIteration :873, fluid.hpp:157-165 */
  for (int n = 0; n < dim; ++n) {
    double* fl[4] = {s->u[L_TC][n], s->u[L_TP][n], s->u[L_IC][n], s->u[L_IP][n]};
    convdiff_iteration(s, fl, K_VEL, n, s->rho, s->muf, s->fs[n], Fprev, cfg->velocity_relaxation_factor,
                       cfg->time_second_order, s->dt, cfg->linear_solver_velocity, s->w1);
    /* diag coefficient :876-883: sum over components of CoeffSum, / dim */
    for (size_t c = 0; c < s->nc; ++c) s->dc[c] = (n == 0 ? 0. : s->dc[c]) + s->w1[c];
  }
  for (size_t c = 0; c < s->nc; ++c) s->dc[c] = s->dc[c] / (double)dim;
  interp(s, s->dc, K_NONE, 0, s->dfc);                               /* :886-887 */
  for (int d = 0; d < dim; ++d) interp(s, s->u[L_IC][d], K_VEL, d, s->ffu[d]); /* :889-892 */

  /* Rhie-Chow :903-940 and flux-correction coefficients :950-969 */
  for (int d = 0; d < dim; ++d) {
    int ex = s->n[0] + (d == 0), ey = s->n[1] + (d == 1), ez = s->n[2] + (d == 2);
    for (int k = 0; k < ez; ++k) for (int j = 0; j < ey; ++j) for (int i = 0; i < ex; ++i) {
      size_t f = fidx(s, d, i, j, k);
      double vfi = s->ffu[d][f] * s->area[d];
      double mv = cfg->meshvel[d] * s->area[d];
      if (s->ftype[f] == FT_INNER) {
        long cm, cp; face_cells(s, d, i, j, k, &cm, &cp);
        double wide = (s->fgp[d][f] - s->ffr[d][f]) * s->area[d];
        double compact = (pprev[cp] - pprev[cm]) / s->h[d] * s->area[d] - s->ffe[d][f] * s->area[d];
        s->Fs[f] = (vfi + cfg->rhie_chow_factor * (wide - compact) / s->dfc[f] + 0) - mv;
        double coeff = -s->area[d] / (s->h[d] * s->dfc[f]);
        s->cf[f] = -coeff;
      } else {
        s->Fs[f] = vfi - mv;
        s->cf[f] = 0.;
      }
    }
  }
  /* pressure-correction system :972-1014 */
  static const int tmap[6] = {CXM, CXP, CYM, CYP, CZM, CZP};
  for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
    size_t c = cidx(s, i, j, k);
    for (int t = 0; t < 7; ++t) s->a[t][c] = 0.;
    if (s->cexcl[c]) { s->a[CD][c] = 1.; s->rhs[c] = 0.; continue; }
    double diag = 0., cst = 0.; int have = 0;
    for (int q = 0; q < 2 * dim; ++q) {
      size_t f = nface(s, i, j, k, q);
      double sgn = (q & 1) ? 1. : -1.;
      if (s->ftype[f] == FT_INNER) {
        /* face expr {+cf on cm, -cf on cp}; this cell is cm for plus faces */
        double self = ((q & 1) ? s->cf[f] : -s->cf[f]) * sgn;
        double nb = ((q & 1) ? -s->cf[f] : s->cf[f]) * sgn;
        diag = have ? diag + self : self; have = 1;
        s->a[tmap[q]][c] = nb;
      }
      cst += s->Fs[f] * sgn;
    }
    s->a[CD][c] = diag;
    s->rhs[c] = cst + -(s->qvol[c] * s->vol);
  }
  if (s->pfix_cell >= 0) {                                           /* :997-1014 */
    long pc = s->pfix_cell; double val = cfg->pressure_fixed_value;
    long off[7]; nb_offsets(s, off);
    int pi = (int)(pc % s->n[0]), pj = (int)((pc / s->n[0]) % s->n[1]), pk = (int)(pc / ((long)s->n[0] * s->n[1]));
    for (int t = 0; t < 7; ++t) {
      if (t == CD || !nb_exists(s, pi, pj, pk, t)) continue;
      long nb = pc + off[t];                                         /* row nb has its term (6-t) on pc */
      if (s->cexcl[nb]) continue;
      s->rhs[nb] += val * s->a[6 - t][nb];                           /* SetKnownValue linear.hpp:238-250 */
      s->a[6 - t][nb] = 0.;
    }
    for (int t = 0; t < 7; ++t) s->a[t][pc] = 0.;
    s->a[CD][pc] = 1.; s->rhs[pc] = -val;
  }
  int it; double df;
  solve(s, cfg->linear_solver_pressure, s->a, s->rhs, s->pc, &it, &df); /* :1030 */
  s->last_sweeps_total += it + 1; s->last_diff = df;

  /* corrections :1033-1056 */
  for (size_t c = 0; c < s->nc; ++c) pcurr[c] = pprev[c] + cfg->pressure_relaxation_factor * s->pc[c];
  interp(s, s->pc, K_EXTRAP, 0, s->wf);
  double* gpc[3] = {s->w2[0], s->w2[1], s->w2[2]};
  gradient(s, s->wf, gpc);
  for (int d = 0; d < dim; ++d)
    for (size_t c = 0; c < s->nc; ++c) s->u[L_IC][d][c] += gpc[d][c] / (-s->dc[c]);
  for (int d = 0; d < dim; ++d) {
    int ex = s->n[0] + (d == 0), ey = s->n[1] + (d == 1), ez = s->n[2] + (d == 2);
    for (int k = 0; k < ez; ++k) for (int j = 0; j < ey; ++j) for (int i = 0; i < ex; ++i) {
      size_t f = fidx(s, d, i, j, k);
      double r = s->Fs[f];
      if (s->ftype[f] == FT_INNER) {
        long cm, cp; face_cells(s, d, i, j, k, &cm, &cp);
        r += s->pc[cm] * s->cf[f];
        r += s->pc[cp] * (-s->cf[f]);
      }
      s->F[L_IC][f] = r;
    }
  }
  if (cfg->simpler) {                                                /* fluid.hpp:1060-1155 */
    long off[7]; nb_offsets(s, off);
    for (int d = 0; d < dim; ++d) interp(s, s->u[L_IC][d], K_VEL, d, s->ffu[d]);   /* ff_velocity :1069-1071 */
    /* fc_evaluated = momentum equations on the velocity change + restored force - pressure gradient :1073-1094 */
    for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
      size_t c = cidx(s, i, j, k);
      for (int n = 0; n < dim; ++n) {
        double ev = s->mr[n][c];
        for (int t = 0; t < 7; ++t) {
          if (t != CD && !(nb_exists(s, i, j, k, t) && s->ma[n][t][c] != 0.)) continue;   /* terms of the expression only */
          long nb = (long)c + off[t];
          ev += (s->u[L_IC][n][nb] - s->u[L_IP][n][nb]) * s->ma[n][t][c];
        }
        s->fev[n][c] = ev + (s->fcr[n][c] - s->gp[n][c]);
      }
    }
    for (int d = 0; d < dim; ++d) interp(s, s->fev[d], K_NONE, 0, s->wf3[d]);     /* ff_evaluated :1096-1097 */
    /* ff_rhs :1099-1112 -> kf */
    double* frhs = s->kf;
    memset(frhs, 0, s->nf * sizeof(double));
    for (int d = 0; d < dim; ++d) {
      int ex = s->n[0] + (d == 0), ey = s->n[1] + (d == 1), ez = s->n[2] + (d == 2);
      for (int k = 0; k < ez; ++k) for (int j = 0; j < ey; ++j) for (int i = 0; i < ex; ++i) {
        size_t f = fidx(s, d, i, j, k);
        if (s->ftype[f] != FT_INNER) continue;
        double dot1 = 0., dot2 = 0.;
        for (int c = 0; c < dim; ++c) {
          const double S = (c == d ? s->area[d] : 0.);
          dot1 += (s->wf3[c][f] - s->ffe[c][f]) * S;
          dot2 += s->ffu[c][f] * S;
        }
        frhs[f] = dot1 / s->dfc[f] + (s->F[L_IC][f] - dot2) / cfg->rhie_chow_factor;
      }
    }
    /* constants :1114-1121, cell conditions :1124-1141 (the terms towards the fixed cell were removed by the first
     * solve's SetKnownValue), then constant := Evaluate(pressure) :1143-1146 */
    for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
      size_t c = cidx(s, i, j, k);
      double sum = 0.;
      for (int q = 0; q < 2 * dim; ++q) sum += frhs[nface(s, i, j, k, q)] * ((q & 1) ? 1. : -1.);
      double cst = -sum;
      if ((long)c == s->pfix_cell) cst = -cfg->pressure_fixed_value;
      for (int t = 0; t < 7; ++t) {
        if (t != CD && !(nb_exists(s, i, j, k, t) && s->a[t][c] != 0.)) continue;
        cst += pcurr[(long)c + off[t]] * s->a[t][c];
      }
      s->rhs[c] = cst;
    }
    int it2; double df2;
    solve(s, cfg->linear_solver_pressure, s->a, s->rhs, s->pc, &it2, &df2);    /* :1148 */
    s->last_sweeps_total += it2 + 1; s->last_diff = df2;
    for (size_t c = 0; c < s->nc; ++c) pcurr[c] += s->pc[c];                   /* :1150-1152 */
  }
  ++s->iter_count;
  return 0;
}

int ho_fluid_convergence_indicator(ho_handle s, double* out) {       /* fluid.hpp:174-179, solver.hpp:804-813 */
  if (s->iter_count == 0) { *out = 1.; return 0; }
  double res = 0.;
  for (size_t c = 0; c < s->nc; ++c) {
    double sq = 0.;
    for (int d = 0; d < s->dim; ++d) { double e = s->u[L_IP][d][c] - s->u[L_IC][d][c]; sq += e * e; }
    res = fmax(res, sqrt(sq));
  }
  *out = res; return 0;
}

int ho_fluid_is_converged(ho_handle s, int* out) {                   /* solver.hpp:733-736 */
  double r; ho_fluid_convergence_indicator(s, &r);
  *out = (s->iter_count >= s->cfg.num_iterations_limit) || (r < s->cfg.convergence_tolerance);
  return 0;
}

int ho_fluid_finish_step(ho_handle s) {                              /* fluid.hpp:1159-1169 */
  memcpy(s->p[L_TP], s->p[L_TC], s->nc * sizeof(double));
  memcpy(s->F[L_TP], s->F[L_TC], s->nf * sizeof(double));
  memcpy(s->p[L_TC], s->p[L_IC], s->nc * sizeof(double));
  memcpy(s->F[L_TC], s->F[L_IC], s->nf * sizeof(double));
  if (has_nan(s->p[L_TC], s->nc)) { snprintf(s->err, sizeof s->err, "NaN pressure"); return HG_ERR_NAN; }
  for (int d = 0; d < s->dim; ++d) {                                 /* conv_diff.hpp:252-259 */
    memcpy(s->u[L_TP][d], s->u[L_TC][d], s->nc * sizeof(double));
    memcpy(s->u[L_TC][d], s->u[L_IC][d], s->nc * sizeof(double));
    if (has_nan(s->u[L_TC][d], s->nc)) { snprintf(s->err, sizeof s->err, "NaN field"); return HG_ERR_NAN; }
  }
  s->time_fluid += s->dt;
  return 0;
}

int ho_fluid_auto_time_step(ho_handle s, double* out) {              /* fluid.hpp:1191-1206 */
  double dt = 1e10;
  for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
    if (s->cexcl[cidx(s, i, j, k)]) continue;
    for (int q = 0; q < 2 * s->dim; ++q) {
      double fl = s->F[L_TC][nface(s, i, j, k, q)];
      if (fl != 0.) dt = fmin(dt, fabs(s->vol / fl));
    }
  }
  *out = dt; return 0;
}

int ho_set_time_step(ho_handle s, double dt_fluid, double dt_adv) {
  s->dt = dt_fluid; s->dt_adv = dt_adv; return 0;
}

/* --------------------------------------------- AdvectionSolverMultiExplicit */
static void advection_iteration(struct ho_state* s) {                /* advection.hpp:440-539 */
  const int dim = s->dim;
  const double* F = s->F[L_IC];                                      /* hydro2d.hpp:596 */
  for (int ph = 0; ph < s->cfg.num_phases; ++ph) {
    double* prev = s->pd[ph][L_IP]; double* curr = s->pd[ph][L_IC];
    /* ff_volume_flux = mixture flux + slip flux of the phase (advection.hpp:449-454) */
    const double* Fa = F;
    if (s->any_slip) { for (size_t f = 0; f < s->nf; ++f) s->cf[f] = F[f] + s->fslip[ph][f]; Fa = s->cf; }
    memcpy(prev, curr, s->nc * sizeof(double));
    int num_stages = s->cfg.tvd_split ? dim : 1;
    for (int stage = 0; stage < num_stages; ++stage) {
      interp(s, curr, K_PD, ph, s->wf);
      double* g[3] = {s->w2[0], s->w2[1], s->w2[2]};
      gradient(s, s->wf, g);
      /* InterpolateSuperbee solver.hpp:560-619 -> wf reused for the result */
      double* fu = s->kf;
      memset(fu, 0, s->nf * sizeof(double));
      for (int d = 0; d < dim; ++d) {
        int ex = s->n[0] + (d == 0), ey = s->n[1] + (d == 1), ez = s->n[2] + (d == 2);
        for (int k = 0; k < ez; ++k) for (int j = 0; j < ey; ++j) for (int i = 0; i < ex; ++i) {
          size_t f = fidx(s, d, i, j, k);
          long P, E; face_cells(s, d, i, j, k, &P, &E);
          if (s->ftype[f] == FT_INNER) {
            double du = curr[E] - curr[P];
            if (Fa[f] > 1e-8) {
              double pq = -4. * (g[d][P] * (-0.5 * s->h[d])) - du;
              double sb = 0.;
              if (du > 0. && pq > 0.) sb = fmax(fmin(2 * du, pq), fmin(du, 2 * pq));
              else if (du < 0. && pq < 0.) sb = -fmax(fmin(-2 * du, -pq), fmin(-du, -2 * pq));
              fu[f] = curr[P] + 0.5 * sb;
            } else if (Fa[f] < -1e-8) {
              double pq = 4. * (g[d][E] * (0.5 * s->h[d])) - du;
              double sb = 0.;
              if (du > 0. && pq > 0.) sb = fmax(fmin(2 * du, pq), fmin(du, 2 * pq));
              else if (du < 0. && pq < 0.) sb = -fmax(fmin(-2 * du, -pq), fmin(-du, -2 * pq));
              fu[f] = curr[E] - 0.5 * sb;
            } else fu[f] = 0.5 * (curr[P] + curr[E]);
          } else if (s->ftype[f] == FT_BOUND) {
            fu[f] = s->wf[f];  /* same Dirichlet / zero-derivative values as Interpolate (solver.hpp:599-617) */
          }
        }
      }
      for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
        size_t c = cidx(s, i, j, k);
        double fsum = 0.;
        for (int q = 0; q < 2 * dim; ++q) {
          if ((q / 2) % num_stages != stage) continue;
          size_t f = nface(s, i, j, k, q);
          fsum += fu[f] * Fa[f] * ((q & 1) ? 1. : -1.);
        }
        curr[c] += -s->dt_adv / s->vol * fsum;
      }
    }
    /* Interface sharpening, advection.hpp:479-529 (once per field, after the stages).  With sharp == 0 every ff is 0 and the
     * update adds dt*0/V (the sources are zero: chem_intensity 0): skipped, it cannot change a value */
    if (fabs(s->cfg.sharp) > 1e-10) {
      double* af = s->wf; double* ff = s->kf;
      double* gc[3] = {s->w2[0], s->w2[1], s->w2[2]};
      interp(s, curr, K_PD, ph, af);                               /* :488 */
      gradient(s, af, gc);                                         /* :489 */
      for (int d = 0; d < dim; ++d) interp(s, gc[d], K_NEUMANN0, 0, s->ffe[d]);   /* :490, zero-derivative for Vect */
      memset(ff, 0, s->nf * sizeof(double));
      const double am = s->cfg.density[ph];                        /* sharp_max_ = v_true_density, hydro2d.hpp:600 */
      for (int d = 0; d < dim; ++d) {
        int ex = s->n[0] + (d == 0), ey = s->n[1] + (d == 1), ez = s->n[2] + (d == 2);
        for (int k = 0; k < ez; ++k) for (int j = 0; j < ey; ++j) for (int i = 0; i < ex; ++i) {
          size_t f = fidx(s, d, i, j, k);
          double n[3] = {0., 0., 0.}, sq = 0.;
          for (int c = 0; c < dim; ++c) { n[c] = s->ffe[c][f]; sq += n[c] * n[c]; }
          const double nrm = sqrt(sq);
          if (nrm < 1.) continue;                                  /* th = 1 */
          double nf = 0.;
          for (int c = 0; c < dim; ++c) { n[c] /= (nrm + 1e-6); nf += n[c] * (c == d ? 1. : 0.); }   /* n.dot(GetNormal) */
          const double uf = Fa[f];
          const double epsh = s->cfg.sharp * s->area[d];
          ff[f] = fabs(uf * nf) * nf * (epsh * nrm - af[f] * (1. - af[f] / am));
        }
      }
      for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
        size_t c = cidx(s, i, j, k);
        curr[c] += s->dt_adv * 0.;                                 /* sources :513-515 */
        double sh = 0.;
        for (int q = 0; q < 2 * dim; ++q) sh += ((q & 1) ? 1. : -1.) * ff[nface(s, i, j, k, q)];
        curr[c] += s->dt_adv * sh / s->vol;                        /* :529 */
      }
    }
  }
}

int ho_advection_step(ho_handle s) {                                 /* advection.hpp:417-423, 540-545 */
  for (int ph = 0; ph < s->cfg.num_phases; ++ph) {
    memcpy(s->pd[ph][L_TP], s->pd[ph][L_TC], s->nc * sizeof(double));
    memcpy(s->pd[ph][L_IC], s->pd[ph][L_TP], s->nc * sizeof(double));
  }
  advection_iteration(s);
  for (int ph = 0; ph < s->cfg.num_phases; ++ph) memcpy(s->pd[ph][L_TC], s->pd[ph][L_IC], s->nc * sizeof(double));
  s->time_adv += s->dt_adv;
  return 0;
}

/* ------------------------------------------------------------------ heat */
int ho_heat_step(ho_handle s) {                                      /* heat.hpp:69-84, hydro2d.hpp:1607-1613 */
  if (has_nan(s->T[L_TC], s->nc)) { snprintf(s->err, sizeof s->err, "NaN initial field"); return HG_ERR_NAN; }
  memcpy(s->T[L_IC], s->T[L_TC], s->nc * sizeof(double));            /* StartStep, guess_extrapolation = 0 */
  /* CalcStep: limit 1, tolerance 0.5 (hydro2d.hpp:685): exactly one iteration since indicator starts at 1 */
  interp(s, s->kc, K_NEUMANN0, 0, s->kf);                            /* heat.hpp:73-75 */
  convdiff_iteration(s, s->T, K_TEMP, 0, NULL, s->kf, s->tsrc, s->F[L_IC], s->cfg.heat_relaxation_factor,
                     s->cfg.time_second_order_heat, s->cfg.dt /* HeatSolver keeps its ctor dt, hydro2d.hpp:684 */,
                     s->cfg.linear_solver_heat, NULL);
  memcpy(s->T[L_TP], s->T[L_TC], s->nc * sizeof(double));            /* FinishStep conv_diff.hpp:252-259 */
  memcpy(s->T[L_TC], s->T[L_IC], s->nc * sizeof(double));
  if (has_nan(s->T[L_TC], s->nc)) { snprintf(s->err, sizeof s->err, "NaN field"); return HG_ERR_NAN; }
  s->time_heat += s->cfg.dt;
  return 0;
}

/* ------------------------------------------------------- fluid properties */
/* CalcPhaseVelocitySlip, hydro2d.hpp:1030-1122 (velocity_is_carrier 0): Stokes settling velocity of every phase relative
 * to the carrier, made relative to the mixture; slip flux on the inner faces, corrected to zero volume-weighted average */
static void calc_slip(struct ho_state* s) {
  if (!s->any_slip) return;
  const hg_config* cfg = &s->cfg;
  const int np = cfg->num_phases, dim = s->dim;
  for (size_t c = 0; c < s->nc; ++c) {
    double rel[HG_MAX_PHASES][3] = {{0.}};
    for (int i = 0; i < np; ++i) {
      if (!cfg->enable_settling[i]) continue;
      const double pc = s->vf[i][c], md = s->rho_raw[c], mv = s->mu[c], pdn = cfg->density[i];
      for (int d = 0; d < dim; ++d)
        rel[i][d] = (pc < 0.01 || pc > 0.99) ? 0. : cfg->gravity[d] * (pdn - md) * (cfg->bubble_radius[i] * cfg->bubble_radius[i]) / (18. * mv);
    }
    double carrier[3] = {0., 0., 0.};
    for (int i = 0; i < np; ++i) for (int d = 0; d < dim; ++d) carrier[d] += rel[i][d] * s->vf[i][c];
    for (int i = 0; i < np; ++i) for (int d = 0; d < dim; ++d) s->slipv[i][d][c] = rel[i][d] - carrier[d];
  }
  for (int d = 0; d < dim; ++d) {
    int ex = s->n[0] + (d == 0), ey = s->n[1] + (d == 1), ez = s->n[2] + (d == 2);
    for (int k = 0; k < ez; ++k) for (int j = 0; j < ey; ++j) for (int i = 0; i < ex; ++i) {
      size_t f = fidx(s, d, i, j, k);
      if (s->ftype[f] != FT_INNER) { for (int ph = 0; ph < np; ++ph) s->fslip[ph][f] = 0.; continue; }
      long cm, cp; face_cells(s, d, i, j, k, &cm, &cp);
      for (int ph = 0; ph < np; ++ph) {
        double dot = 0.;
        for (int c = 0; c < dim; ++c) dot += (s->slipv[ph][c][cm] * (1. - 0.5) + s->slipv[ph][c][cp] * 0.5) * (c == d ? s->area[d] : 0.);
        s->fslip[ph][f] = dot;
      }
      double aver = 0.;
      for (int ph = 0; ph < np; ++ph) aver += (s->vf[ph][cm] * (1. - 0.5) + s->vf[ph][cp] * 0.5) * s->fslip[ph][f];
      for (int ph = 0; ph < np; ++ph) s->fslip[ph][f] -= aver;
    }
  }
}

int ho_update_properties(ho_handle s) {                              /* hydro2d.hpp:1404-1430 */
  const hg_config* cfg = &s->cfg;
  const int np = cfg->num_phases, dim = s->dim;
  for (size_t c = 0; c < s->nc; ++c) {                               /* CalcPhasesVolumeFraction :981-997 */
    double sum = 0.;
    for (int i = 0; i < np; ++i) { s->vf[i][c] = s->pd[i][L_TC][c] / cfg->density[i]; sum += s->vf[i][c]; }
    for (int i = 0; i < np; ++i) s->vf[i][c] /= sum;
  }
  for (size_t c = 0; c < s->nc; ++c) {                               /* GetVolumeAveraged :1235-1246 */
    double r = 0., m = 0., kk = 0.;
    for (int i = 0; i < np; ++i) { r += cfg->density[i] * s->vf[i][c]; m += cfg->viscosity[i] * s->vf[i][c]; kk += cfg->conductivity[i] * s->vf[i][c]; }
    s->rho_raw[c] = r; s->w1[c] = m; s->kc[c] = kk;
    s->qvol[c] = 0.; s->qmass[c] = 0.; s->tsrc[c] = 0.;
  }
  smooth(s, s->rho_raw, cfg->density_smooth_times, s->rho, s->wf, s->corr);
  smooth(s, s->w1, cfg->viscosity_smooth_times, s->mu, s->wf, s->corr);
  /* CalcForce :1308-1376 */
  for (size_t c = 0; c < s->nc; ++c)
    for (int d = 0; d < dim; ++d) { s->force[d][c] = cfg->force[d] + cfg->gravity[d] * s->rho_raw[c]; s->stforce[d][c] = 0.; }
  if (np >= 2) {
    /* surface tension :1318-1370: a = volume fraction of phase 1, conditions of phase 0 */
    interp(s, s->vf[1], K_PD, 0, s->wf);
    double* gsc[3] = {s->w2[0], s->w2[1], s->w2[2]};
    gradient(s, s->wf, gsc);
    for (int d = 0; d < dim; ++d) interp(s, gsc[d], K_NEUMANN0, 0, s->wf3[d]);
    for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
      size_t c = cidx(s, i, j, k);
      double fv[3] = {0., 0., 0.};
      for (int q = 0; q < 2 * dim; ++q) {
        size_t f = nface(s, i, j, k, q);
        int qd = q >> 1;
        double g[3] = {0, 0, 0}, nn[3] = {0, 0, 0};
        double sq = 0.;
        for (int d = 0; d < dim; ++d) { g[d] = s->wf3[d][f]; sq += g[d] * g[d]; }
        double nrm = sqrt(sq);
        for (int d = 0; d < dim; ++d) nn[d] = g[d] / (nrm + 1e-6);
        double so[3] = {0., 0., 0.}; so[qd] = s->area[qd] * ((q & 1) ? 1. : -1.);
        double sdn = 0.; for (int d = 0; d < dim; ++d) sdn += so[d] * nn[d];
        for (int d = 0; d < dim; ++d) { fv[d] += g[d] * sdn; fv[d] -= so[d] * nrm; }
      }
      for (int d = 0; d < dim; ++d) { fv[d] /= s->vol; s->stforce[d][c] = fv[d] * cfg->sigma; }
    }
  }
  for (int d = 0; d < dim; ++d) {
    smooth(s, s->force[d], cfg->force_smooth_times, s->w1, s->wf, s->corr);
    memcpy(s->force[d], s->w1, s->nc * sizeof(double));
  }
  calc_slip(s);                                                      /* :1422 */
  return 0;
}

int ho_calc_stat(ho_handle s, hg_step_stats* st) {                   /* hydro2d.hpp:1432-1529 */
  for (int i = 0; i < s->cfg.num_phases; ++i) {
    double volume = 0., pmin = 1e10, pmax = -1e10, cen[3] = {0, 0, 0}, vel[3] = {0, 0, 0};
    for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int ii = 0; ii < s->n[0]; ++ii) {
      size_t c = cidx(s, ii, j, k);
      double x[3]; cell_center(s, ii, j, k, x);
      double cc = s->vf[i][c];
      volume += cc * s->vol;
      double pd = s->pd[i][L_TC][c];
      pmin = fmin(pmin, pd); pmax = fmax(pmax, pd);
      for (int d = 0; d < s->dim; ++d) { cen[d] += x[d] * (cc * s->vol); vel[d] += s->u[L_TC][d][c] * (cc * s->vol); }
    }
    st->volume[i] = volume; st->mass[i] = volume * s->cfg.density[i];
    st->pd_min[i] = pmin; st->pd_max[i] = pmax;
    for (int d = 0; d < 3; ++d) { st->center[i][d] = d < s->dim ? cen[d] / volume + s->meshpos[d] : 0.; st->velocity[i][d] = d < s->dim ? vel[d] / volume : 0.; }
  }
  /* stat_vcx_<i> = (cx - previous cx) / dt, 0 at the first call (:1459-1463); ADHOC mesh velocity towards phase 1 (:1510-1524) */
  double vcx[HG_MAX_PHASES];
  for (int i = 0; i < s->cfg.num_phases; ++i) {
    const double prev = s->stat_cx_set[i] ? s->stat_cx[i] : st->center[i][0];
    vcx[i] = (st->center[i][0] - prev) / s->dt;
    s->stat_cx[i] = st->center[i][0]; s->stat_cx_set[i] = 1;
  }
  if (s->cfg.meshvel_auto) {
    const double v0 = s->cfg.meshvel_auto == 1 ? st->velocity[1][0] : vcx[1], w = s->cfg.meshvel_weight;
    for (int d = 0; d < 3; ++d) s->cfg.meshvel[d] = (d == 0 ? v0 : 0.) * w + s->cfg.meshvel[d] * (1. - w);
  }
  if (s->cfg.meshvel_output) for (int d = 0; d < s->dim; ++d) s->meshpos[d] += s->cfg.meshvel[d] * s->dt; /* :1526-1528 */
  return 0;
}

/* ------------------------------------------------------------------ step */
int ho_step(ho_handle s, hg_step_stats* stats) {                     /* hydro2d.hpp:1531-1621 */
  const hg_config* cfg = &s->cfg;
  int rc;
  s->nres = 0; s->last_sweeps_total = 0;
  if (cfg->dt_auto) {
    double dtm; ho_fluid_auto_time_step(s, &dtm);
    s->dt = dtm * cfg->cfl; s->dt_adv = dtm * cfg->cfl_advection;
  }
  if ((rc = ho_fluid_start_step(s))) return rc;
  if (cfg->fluid_enable) {
    int conv; ho_fluid_is_converged(s, &conv);
    while (!conv) {
      if ((rc = ho_fluid_make_iteration(s))) return rc;
      double r; ho_fluid_convergence_indicator(s, &r);
      if (s->nres < 4096) s->res_hist[s->nres++] = r;
      ho_fluid_is_converged(s, &conv);
    }
  }
  if ((rc = ho_fluid_finish_step(s))) return rc;
  int nadv = 0;
  if (cfg->advection_enable) {
    while (s->time_adv < s->time_fluid - 0.5 * s->dt_adv) { ho_advection_step(s); ++nadv; }
  }
  if (cfg->heat_enable) { if ((rc = ho_heat_step(s))) return rc; }
  ho_update_properties(s);
  ho_calc_stat(s, &s->stat);
  s->stat.simple_iterations = s->iter_count;
  s->stat.convergence_indicator = s->nres ? s->res_hist[s->nres - 1] : 1.;
  s->stat.pressure_sweeps_total = s->last_sweeps_total;
  s->stat.pressure_last_diff = s->last_diff;
  s->stat.advection_substeps = nadv;
  s->stat.dt = s->dt; s->stat.time = s->time_fluid;
  if (stats) *stats = s->stat;
  return 0;
}

int ho_last_residuals(ho_handle s, double* out, int cap, int* n) {
  int m = s->nres < cap ? s->nres : cap;
  memcpy(out, s->res_hist, m * sizeof(double)); *n = m; return 0;
}

/* ---------------------------------------------------------------- set up */
void ho_config_defaults(hg_config* c) {          /* examples/general.hydroconf */
  memset(c, 0, sizeof *c);
  c->dim = 2; c->Nx = 100; c->Ny = 100; c->Nz = 5;
  c->B[0] = c->B[1] = c->B[2] = 1.;
  c->B1[2] = 1.; c->B2[2] = 1.;
  c->dt = 0.01; c->cfl = 0.5; c->cfl_advection = 0.5;
  c->num_phases = 1;
  for (int i = 0; i < HG_MAX_PHASES; ++i) { c->density[i] = 1.; c->viscosity[i] = 1.; c->conductivity[i] = 1.; }
  c->fluid_enable = 1; c->advection_enable = 1;
  c->advection_dt_factor = 0.1;
  c->convergence_tolerance = 1e-2; c->num_iterations_limit = 10;
  c->velocity_relaxation_factor = 0.8; c->pressure_relaxation_factor = 0.9;
  c->linear_solver_velocity = HG_LS_LU; c->linear_solver_pressure = HG_LS_GAUSS_SEIDEL; c->linear_solver_heat = HG_LS_LU;
  c->lu_relaxed_relaxation_factor = 1.9; c->lu_relaxed_num_iters_limit = 1000; c->lu_relaxed_tolerance = 1e-3;
  c->time_second_order = 1; c->rhie_chow_factor = 1.;
  c->initial_volume_fraction_smooth_times = 2; c->density_smooth_times = 2; c->viscosity_smooth_times = 2;
  c->heat_relaxation_factor = 1.; c->time_second_order_heat = 1; c->meshvel_output = 1;
  c->meshvel_auto = 0; c->meshvel_weight = 0.5;
  c->world_size = 1;
}

static int inside(const double lb[3], const double rt[3], const double x[3], int dim) { /* Rect::IsInside vect.hpp:236-243 */
  for (int d = 0; d < dim; ++d) if (x[d] < lb[d] || rt[d] < x[d]) return 0;
  return 1;
}

int ho_create(const hg_config* cfg, ho_handle* out) {
  if (!cfg || !out) return HG_ERR_INVALID;
  if ((cfg->dim != 2 && cfg->dim != 3) || cfg->num_phases < 1 || cfg->num_phases > HG_MAX_PHASES ||
      cfg->force_geometric_average || cfg->velocity_is_carrier) {
    snprintf(g_err, sizeof g_err, "unsupported configuration"); return HG_ERR_INVALID;
  }
  struct ho_state* s = (struct ho_state*)calloc(1, sizeof *s);
  s->cfg = *cfg;
  const int dim = s->dim = cfg->dim;
  s->n[0] = cfg->Nx; s->n[1] = cfg->Ny; s->n[2] = dim > 2 ? cfg->Nz : 1;
  s->nc = (size_t)s->n[0] * s->n[1] * s->n[2];
  s->vol = 1.;
  for (int d = 0; d < 3; ++d) {
    s->lb[d] = cfg->A[d];
    s->h[d] = d < dim ? (cfg->B[d] - cfg->A[d]) / s->n[d] : 1.;
    if (d < dim) s->vol *= s->h[d];
  }
  for (int d = 0; d < dim; ++d) { s->area[d] = 1.; for (int e = 0; e < dim; ++e) if (e != d) s->area[d] *= s->h[e]; }
  s->nf = 0;
  for (int d = 0; d < 3; ++d) {
    s->foff[d] = s->nf;
    s->nfd[d] = d < dim ? (size_t)(s->n[0] + (d == 0)) * (s->n[1] + (d == 1)) * (s->n[2] + (d == 2)) : 0;
    s->nf += s->nfd[d];
  }
  const size_t nc = s->nc, nf = s->nf;
  s->cexcl = (unsigned char*)calloc(nc, 1); s->ftype = (unsigned char*)calloc(nf, 1);
  s->fside = (signed char*)calloc(nf, 1); s->ftdir = (unsigned char*)calloc(nf, 1);
  for (int sd = 0; sd < 6; ++sd) { s->bckind[sd] = cfg->condition_kind[sd]; for (int d = 0; d < 3; ++d) s->bcvel[sd][d] = cfg->condition_velocity[sd][d]; }
  s->bckind[6] = HG_BC_WALL;
  /* rigid box hydro2d.hpp:306-308, 409-416 */
  for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
    double x[3]; cell_center(s, i, j, k, x);
    if (inside(cfg->box_A, cfg->box_B, x, dim)) s->cexcl[cidx(s, i, j, k)] = 1;
  }
  /* face classification hydro2d.hpp:371-426 + MeshStructured::ExcludeCells */
  for (int d = 0; d < dim; ++d) {
    int ex = s->n[0] + (d == 0), ey = s->n[1] + (d == 1), ez = s->n[2] + (d == 2);
    for (int k = 0; k < ez; ++k) for (int j = 0; j < ey; ++j) for (int i = 0; i < ex; ++i) {
      size_t f = fidx(s, d, i, j, k);
      long cm, cp; face_cells(s, d, i, j, k, &cm, &cp);
      int idx[3] = {i, j, k};
      if (cm >= 0 && cp >= 0) s->ftype[f] = FT_INNER;
      else if (cm < 0 && cp < 0) s->ftype[f] = FT_EXCL;
      else {
        s->ftype[f] = FT_BOUND;
        if (idx[d] == 0) s->fside[f] = (signed char)(2 * d);
        else if (idx[d] == s->n[d]) s->fside[f] = (signed char)(2 * d + 1);
        else s->fside[f] = 6;
        double xf[3]; cell_center(s, i, j, k, xf); xf[d] -= 0.5 * s->h[d];
        if (inside(cfg->heat_box_lb, cfg->heat_box_rt, xf, dim)) s->ftdir[f] = 1; /* hydro2d.hpp:664-676 */
      }
    }
  }
  s->pfix_cell = -1;
  if (cfg->pressure_fixed_enable) {                                  /* FindNearestCell mesh.hpp:411-420 */
    long best = 0; double bd = 0.;
    for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
      double x[3]; cell_center(s, i, j, k, x);
      double sq = 0.; for (int d = 0; d < dim; ++d) { double e = x[d] - cfg->pressure_fixed_point[d]; sq += e * e; }
      double dd = sqrt(sq);
      long c = (long)cidx(s, i, j, k);
      if (c == 0) { bd = dd; best = 0; } else if (dd < bd) { bd = dd; best = c; }
    }
    s->pfix_cell = best;
  }
  for (int l = 0; l < 4; ++l) {
    for (int d = 0; d < 3; ++d) s->u[l][d] = dalloc(nc);
    s->p[l] = dalloc(nc); s->F[l] = dalloc(nf); s->T[l] = dalloc(nc);
    for (int ph = 0; ph < HG_MAX_PHASES; ++ph) s->pd[ph][l] = dalloc(nc);
  }
  for (int ph = 0; ph < HG_MAX_PHASES; ++ph) { s->vf[ph] = dalloc(nc); s->pd_inlet[ph] = dalloc(nf); }
  for (int d = 0; d < 3; ++d) s->outvel[d] = dalloc(nf);
  if (cfg->simpler) for (int n = 0; n < 3; ++n) { s->mr[n] = dalloc(nc); s->fev[n] = dalloc(nc); for (int t = 0; t < 7; ++t) s->ma[n][t] = dalloc(nc); }
  for (int ph = 0; ph < HG_MAX_PHASES; ++ph) { s->fslip[ph] = dalloc(nf); for (int d = 0; d < 3; ++d) s->slipv[ph][d] = dalloc(nc); if (ph < cfg->num_phases && cfg->enable_settling[ph]) s->any_slip = 1; }
  for (int sd = 0; sd < 2 * dim; ++sd) if (cfg->condition_kind[sd] == HG_BC_OUTLET) s->any_outlet = 1;
  s->rho_raw = dalloc(nc); s->rho = dalloc(nc); s->mu = dalloc(nc); s->kc = dalloc(nc);
  s->qvol = dalloc(nc); s->qmass = dalloc(nc); s->tsrc = dalloc(nc);
  s->muf = dalloc(nf); s->ffp = dalloc(nf); s->dc = dalloc(nc); s->dfc = dalloc(nf); s->Fs = dalloc(nf); s->cf = dalloc(nf);
  s->pc = dalloc(nc); s->rhs = dalloc(nc); s->corr = dalloc(nc); s->w1 = dalloc(nc); s->wf = dalloc(nf); s->kf = dalloc(nf);
  for (int t = 0; t < 7; ++t) s->a[t] = dalloc(nc);
  for (int d = 0; d < 3; ++d) {
    s->force[d] = dalloc(nc); s->stforce[d] = dalloc(nc); s->ffe[d] = dalloc(nf); s->fcr[d] = dalloc(nc); s->ffr[d] = dalloc(nf);
    s->gp[d] = dalloc(nc); s->fgp[d] = dalloc(nf); s->fs[d] = dalloc(nc); s->ffu[d] = dalloc(nf); s->w2[d] = dalloc(nc); s->wf3[d] = dalloc(nf);
  }
  s->dt = cfg->dt; s->dt_adv = cfg->dt * cfg->advection_dt_factor;   /* hydro2d.hpp:599 */

  /* initial velocity hydro2d.hpp:310-368 */
  double pi = atan(1.) * 4.;
  for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
    size_t c = cidx(s, i, j, k);
    double x[3]; cell_center(s, i, j, k, x);
    double v[3] = {cfg->initial_velocity[0], cfg->initial_velocity[1], dim > 2 ? cfg->initial_velocity[2] : 0.};
    if (cfg->initial_pois) v[0] = x[1] * (1. - x[1]) * 4. * cfg->initial_velocity[0];
    if (cfg->initial_sin_enable) {
      double kd = 0.;
      for (int d = 0; d < dim; ++d) kd += (cfg->initial_sin_n[d] * (2. * pi / cfg->initial_sin_lambda)) * x[d];
      double sn = sin(kd - cfg->initial_sin_phase);
      for (int d = 0; d < dim; ++d) v[d] *= sn;
    }
    for (int d = 0; d < dim; ++d) s->u[L_TC][d][c] = s->u[L_TP][d][c] = v[d];
  }
  /* FluidSimple ctor: initial volume fluxes fluid.hpp:770-785 */
  for (int d = 0; d < dim; ++d) {
    interp(s, s->u[L_TC][d], K_VEL, d, s->wf);
    for (size_t f = s->foff[d]; f < s->foff[d] + s->nfd[d]; ++f) {
      s->F[L_TC][f] = s->wf[f] * s->area[d] - cfg->meshvel[d] * s->area[d];
      s->F[L_TP][f] = s->F[L_TC][f];
    }
  }
  /* InitAdvectionSolver hydro2d.hpp:479-550 */
  for (int ph = 0; ph < cfg->num_phases; ++ph)
    for (size_t c = 0; c < nc; ++c) s->pd[ph][L_TC][c] = cfg->density[ph] * cfg->initial_volume_fraction[ph];
  for (int k = 0; k < s->n[2]; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
    size_t c = cidx(s, i, j, k);
    double x[3]; cell_center(s, i, j, k, x);
    double d1 = 0., d2 = 0.;
    for (int d = 0; d < dim; ++d) { double e = cfg->IC[d] - x[d]; d1 += e * e; e = cfg->IC2[d] - x[d]; d2 += e * e; }
    if (inside(cfg->A2, cfg->B2, x, dim)) { if (cfg->num_phases > 2) s->pd[2][L_TC][c] = cfg->density[2]; }
    else if (inside(cfg->A1, cfg->B1, x, dim)) { if (cfg->num_phases > 1) s->pd[1][L_TC][c] = cfg->density[1]; }
    else if (sqrt(d1) < cfg->IR) { if (cfg->num_phases > 1) s->pd[1][L_TC][c] = cfg->density[1]; }
    else if (sqrt(d2) < cfg->IR2) { if (cfg->num_phases > 1) s->pd[1][L_TC][c] = cfg->density[1]; }
  }
  for (int ph = 1; ph < cfg->num_phases; ++ph) {
    smooth(s, s->pd[ph][L_TC], cfg->initial_volume_fraction_smooth_times, s->w1, s->wf, s->corr);
    memcpy(s->pd[ph][L_TC], s->w1, nc * sizeof(double));
  }
  for (size_t c = 0; c < nc; ++c) {
    double vs = 0.;
    for (int ph = 1; ph < cfg->num_phases; ++ph) vs += s->pd[ph][L_TC][c] / cfg->density[ph];
    s->pd[0][L_TC][c] = (1. - vs) * cfg->density[0];
  }
  /* inlet values of partial density hydro2d.hpp:563-575 */
  for (int d = 0; d < dim; ++d) {
    int ex = s->n[0] + (d == 0), ey = s->n[1] + (d == 1), ez = s->n[2] + (d == 2);
    for (int k = 0; k < ez; ++k) for (int j = 0; j < ey; ++j) for (int i = 0; i < ex; ++i) {
      size_t f = fidx(s, d, i, j, k);
      if (s->ftype[f] != FT_BOUND || s->bckind[(int)s->fside[f]] != HG_BC_INLET) continue;
      long cm, cp; face_cells(s, d, i, j, k, &cm, &cp);
      long cc = cm >= 0 ? cm : cp;
      for (int ph = 0; ph < cfg->num_phases; ++ph) s->pd_inlet[ph][f] = s->pd[ph][L_TC][cc];
    }
  }
  for (size_t c = 0; c < nc; ++c) s->T[L_TC][c] = s->T[L_TP][c] = cfg->temperature_initial; /* hydro2d.hpp:656-657 */
  ho_update_properties(s);
  ho_calc_stat(s, &s->stat);
  *out = s;
  return 0;
}

int ho_destroy(ho_handle s) {
  if (!s) return 0;
  for (int l = 0; l < 4; ++l) {
    for (int d = 0; d < 3; ++d) free(s->u[l][d]);
    free(s->p[l]); free(s->F[l]); free(s->T[l]);
    for (int ph = 0; ph < HG_MAX_PHASES; ++ph) free(s->pd[ph][l]);
  }
  for (int ph = 0; ph < HG_MAX_PHASES; ++ph) { free(s->vf[ph]); free(s->pd_inlet[ph]); }
  for (int d = 0; d < 3; ++d) free(s->outvel[d]);
  for (int n = 0; n < 3; ++n) { free(s->mr[n]); free(s->fev[n]); for (int t = 0; t < 7; ++t) free(s->ma[n][t]); }
  for (int ph = 0; ph < HG_MAX_PHASES; ++ph) { free(s->fslip[ph]); for (int d = 0; d < 3; ++d) free(s->slipv[ph][d]); }
  free(s->rho_raw); free(s->rho); free(s->mu); free(s->kc); free(s->qvol); free(s->qmass); free(s->tsrc);
  free(s->muf); free(s->ffp); free(s->dc); free(s->dfc); free(s->Fs); free(s->cf);
  free(s->pc); free(s->rhs); free(s->corr); free(s->w1); free(s->wf); free(s->kf);
  for (int t = 0; t < 7; ++t) free(s->a[t]);
  for (int d = 0; d < 3; ++d) {
    free(s->force[d]); free(s->stforce[d]); free(s->ffe[d]); free(s->fcr[d]); free(s->ffr[d]);
    free(s->gp[d]); free(s->fgp[d]); free(s->fs[d]); free(s->ffu[d]); free(s->w2[d]); free(s->wf3[d]);
  }
  free(s->cexcl); free(s->ftype); free(s->fside); free(s->ftdir);
  free(s);
  return 0;
}

const char* ho_last_error(ho_handle s) { return s ? s->err : g_err; }
size_t ho_num_cells(ho_handle s) { return s->nc; }
size_t ho_num_faces(ho_handle s) { return s->nf; }

static double* field_ptr(struct ho_state* s, int field, int layer, size_t* n) {
  *n = s->nc;
  if (field >= HG_F_VELOCITY_X && field <= HG_F_VELOCITY_Z) return field - HG_F_VELOCITY_X < s->dim ? s->u[layer][field - HG_F_VELOCITY_X] : NULL;
  if (field >= HG_F_VELOCITY_PREV_X && field <= HG_F_VELOCITY_PREV_Z) return field - HG_F_VELOCITY_PREV_X < s->dim ? s->u[L_TP][field - HG_F_VELOCITY_PREV_X] : NULL;
  if (field == HG_F_PRESSURE) return s->p[layer];
  if (field == HG_F_PRESSURE_PREV) return s->p[L_TP];
  if (field == HG_F_VOLUME_FLUX) { *n = s->nf; return s->F[layer]; }
  if (field == HG_F_VOLUME_FLUX_PREV) { *n = s->nf; return s->F[L_TP]; }
  if (field >= HG_F_PARTIAL_DENSITY_0 && field <= HG_F_PARTIAL_DENSITY_2) return field - HG_F_PARTIAL_DENSITY_0 < s->cfg.num_phases ? s->pd[field - HG_F_PARTIAL_DENSITY_0][layer] : NULL;
  if (field == HG_F_TEMPERATURE) return s->T[layer];
  if (field == HG_F_DENSITY) return s->rho;
  if (field == HG_F_VISCOSITY) return s->mu;
  if (field == HG_F_CONDUCTIVITY) return s->kc;
  if (field >= HG_F_FORCE_X && field <= HG_F_FORCE_Z) return field - HG_F_FORCE_X < s->dim ? s->force[field - HG_F_FORCE_X] : NULL;
  if (field >= HG_F_STFORCE_X && field <= HG_F_STFORCE_Z) return field - HG_F_STFORCE_X < s->dim ? s->stforce[field - HG_F_STFORCE_X] : NULL;
  if (field >= HG_F_VOLUME_FRACTION_0 && field <= HG_F_VOLUME_FRACTION_2) return field - HG_F_VOLUME_FRACTION_0 < s->cfg.num_phases ? s->vf[field - HG_F_VOLUME_FRACTION_0] : NULL;
  return NULL;
}

int ho_set_field(ho_handle s, int field, const double* src, size_t n) {
  size_t m; double* p = field_ptr(s, field, L_TC, &m);
  if (!p || n != m || field == HG_F_EXCLUDED) { snprintf(s->err, sizeof s->err, "bad field/size"); return HG_ERR_INVALID; }
  memcpy(p, src, n * sizeof(double));
  int layered = (field <= HG_F_TEMPERATURE);
  if (layered) { double* q = field_ptr(s, field, L_TP, &m); memcpy(q, src, n * sizeof(double)); }
  return 0;
}

int ho_get_field(ho_handle s, int field, double* dst, size_t n) {
  if (field == HG_F_EXCLUDED) {
    if (n != s->nc) return HG_ERR_INVALID;
    for (size_t c = 0; c < n; ++c) dst[c] = s->cexcl[c] ? 1. : 0.;
    return 0;
  }
  size_t m; double* p = field_ptr(s, field, L_TC, &m);
  if (!p || n != m) { snprintf(s->err, sizeof s->err, "bad field/size"); return HG_ERR_INVALID; }
  memcpy(dst, p, n * sizeof(double));
  return 0;
}

/* ------------------------------------------------- kernel-level entries */
int ho_interp_grad(ho_handle s, const double* u, int cond, int comp, double* gx, double* gy, double* gz) {
  int kind = cond == 0 ? K_NEUMANN0 : cond == 1 ? K_EXTRAP : K_VEL;
  interp(s, u, kind, comp, s->wf);
  double* g[3] = {s->w2[0], s->w2[1], s->w2[2]};
  gradient(s, s->wf, g);
  memcpy(gx, g[0], s->nc * sizeof(double)); memcpy(gy, g[1], s->nc * sizeof(double));
  if (s->dim > 2 && gz) memcpy(gz, g[2], s->nc * sizeof(double));
  return 0;
}

int ho_linear_solve(ho_handle s, int solver, const double* const coeffs[7], const double* rhs, double* x,
                    double tol, int limit, double relax, int* out_iters, double* out_diff) {
  double* a[7];
  for (int t = 0; t < 7; ++t) {
    a[t] = dalloc(s->nc);
    if (coeffs[t] && (s->dim > 2 || (t != CZM && t != CZP))) memcpy(a[t], coeffs[t], s->nc * sizeof(double));
  }
  hg_config save = s->cfg;
  s->cfg.lu_relaxed_tolerance = tol; s->cfg.lu_relaxed_num_iters_limit = limit; s->cfg.lu_relaxed_relaxation_factor = relax;
  int it; double df;
  solve(s, solver, a, rhs, x, &it, &df);
  s->cfg = save;
  if (out_iters) *out_iters = it;
  if (out_diff) *out_diff = df;
  for (int t = 0; t < 7; ++t) free(a[t]);
  return 0;
}

int ho_smooth_field(ho_handle s, const double* u, int repeat, double* out) {
  double* wc = dalloc(s->nc);
  smooth(s, u, repeat, out, s->wf, wc);
  free(wc);
  return 0;
}
