// hydro_gpu.cu -- C ABI (include/hydro_gpu.h) over the CUDA kernels: device-resident state of one
// reference experiment and the orchestration of hydro<Mesh>::step() (hydro2d.hpp:1531-1621).
// No CPU fallback: every compute entry needs a CUDA device.
#include "../../include/hydro_gpu.h"

#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <initializer_list>
#include <algorithm>

#include "hg_kernels.cuh"
#include "hg_solvers.cuh"
#include "hg_slab.cuh"
#include "hg_gs_tiled.cuh"
#include "hg_lu_tiled.cuh"
#include "hg_fast.cuh"

enum { L_TC = 0, L_TP = 1, L_IC = 2, L_IP = 3 };

struct Timer { cudaEvent_t a, b; double total = 0.; bool open = false; };

struct hg_state {
  hg_config cfg;
  int dev = 0;
  cudaStream_t st = nullptr;
  // asynchronous field transfers (hg_set_field_async / hg_get_field_async): copy streams, staging / snapshot buffers per field
  cudaStream_t st_h2d = nullptr, st_d2h = nullptr;
  std::map<int, double*> xfer_stage, xfer_snap;
  std::map<int, cudaEvent_t> xfer_ev, xfer_stage_ev;
  cudaEvent_t xfer_ev_tmp[2] = {nullptr, nullptr};   // ordering events of the asynchronous transfers (created once)
  int dim = 0, n[3] = {1, 1, 1};
  long long nc = 0, nf = 0, nsh = 0;
  Geo geo;
  unsigned char* excl = nullptr;
  bool any_excl = false;
  double *u[4][3] = {}, *p[4] = {}, *F[4] = {}, *T[4] = {};
  double* pd[HG_MAX_PHASES][4] = {};
  double* pd_init[HG_MAX_PHASES] = {};
  double *vf[HG_MAX_PHASES] = {}, *rho_raw = nullptr, *mu_raw = nullptr, *rho = nullptr, *mu = nullptr, *kc = nullptr;
  double *force[3] = {}, *stforce[3] = {};
  double *gp[3] = {}, *fcr[3] = {}, *G[9] = {}, *fs[3] = {}, *dc = nullptr, *Fs = nullptr, *pc = nullptr;
  double *w1 = nullptr, *w2 = nullptr, *zero = nullptr;
  double *An[10] = {};   // natural-layout staging of rows/constants before the shear transpose (3-D)
  double *A[7] = {}, *R[3] = {}, *X[3] = {}, *D = nullptr, *CYs = nullptr, *CZs = nullptr, *RP = nullptr, *PP = nullptr, *PPsave = nullptr;
  double* mailbox = nullptr; unsigned long long* xflags = nullptr;   // peer-written scratch (multi-GPU)
  // time-skewed tile sweeps (hg_gs_tiled.cuh): explicit diagonal, task lists per number of sweeps in a launch
  bool gs_tiled = false;
  double* DGs = nullptr;
  double* CO = nullptr; Co5 co5;               // rows of k_gs_tiled: five hyperplane-major arrays (gt_co5_index)
  CUtensorMap tmco;
  // task lists of k_gs_tiled per number of sweeps in a launch: GT_PLAN_SLOTS device buffers sized for the largest launch,
  // allocated at creation (no allocation on the stepping path), reused least-recently-used
  struct GtPlan { int S = -1, ntasks = 0; GtTask* tasks = nullptr; int* progress = nullptr; GtTask* stage = nullptr; cudaEvent_t staged = nullptr; long long used = 0; };
  static constexpr int GT_PLAN_SLOTS = 4;
  GtPlan gt_plans[GT_PLAN_SLOTS];
  int gt_plan_cap = 0; long long gt_plan_clock = 0;
  int* gt_ctl = nullptr;
  std::vector<int> sor_pred = std::vector<int>(256, -1);   // stopping sweep of the pressure solve of SIMPLE iteration q in the previous step
  std::vector<int> sor_old = std::vector<int>(256, -1);    // ... and in the step before that (trend of the prediction)
  int gt_gbase = 0;                           // sweep groups of the current solve launched so far (parity of the slabs' "down" planes)
  unsigned long long* gt_clk = nullptr;
  // lu as a dataflow of column boxes (hg_lu_tiled.cuh)
  bool lu_tiled = false; int2* lt_boxes = nullptr; int lt_nboxes = 0, lt_nbi = 0; int* lt_progress = nullptr; int* lt_ctl = nullptr;
  int num_sms = 0;
  // interior kernels (hg_fast.cuh): byte mask of the cells that keep the generic kernels, their list, Geo of the list launches
  // [0]: radius-1 stencils (SIMPLE iteration), [1]: radius 2 (advection)
  bool fast = false; unsigned char* slow = nullptr; int* slow_list = nullptr; int nslow = 0;
  unsigned char* slow2 = nullptr; int* slow2_list = nullptr; int nslow2 = 0;
  bool any_slip = false; double* slipv[HG_MAX_PHASES][3] = {}; double* fslip[HG_MAX_PHASES] = {};   // phase slip
  double* outvel = nullptr; double* outpart = nullptr; bool any_outlet = false; int out_blocks = 0; long long out_terms = 0;   // outlet conditions
  double* resid = nullptr;    // per-iteration convergence indicators of the current step (device, 4096)
  double* scal = nullptr;     // device scalars: [0] resid, [1] auto dt, [2..] stat (36), then diffs
  int* flag = nullptr;        // [0] scratch NaN flag (immediate checks), [1] any-excluded flag, [4..11] deferred NaN flags of a step
  bool step_pending = false; int step_nadv = 0;   // between hg_step_begin and hg_step_end
  bool defer = false;         // inside hg_step: NaN flags, solver status words, sweep counts and statistics are read once, at the end
  double* sorres = nullptr;   // per pressure solve of the step: {iter, diff} computed on the device (deferred mode)
  int nsolves = 0;
  double* hscal = nullptr;    // pinned host mirror
  double* hinit = nullptr;    // pinned: initial values of the statistics accumulators
  int max_sweeps = 0;
  double* diffs = nullptr;    // per-sweep max norms (device)
  double* hdiffs = nullptr;   // pinned
  double time_fluid = 0., time_adv = 0., dt = 0., dt_adv = 0.;
  double meshpos[3] = {0., 0., 0.};
  double stat_cx[HG_MAX_PHASES] = {0., 0., 0.}; bool stat_cx_set[HG_MAX_PHASES] = {false, false, false};   // stat_cx_<i> of the previous CalcStat
  int iter_count = 0;
  double last_resid = 1.;
  int sweeps_total = 0; double last_diff = 0.;
  hg_step_stats stat;
  long long launches = 0;
  int grid_solver = 0, grid_lu = 0;
  TileTable tt;
  bool timers_on = false;
  std::map<std::string, Timer> timers;
  std::vector<std::string> timer_stack;
  std::vector<void*> allocs;
  // z-slab decomposition.  Every array is its own cudaMalloc (one 12 GB arena measured 15 % slower for the
  // sweep kernel on B200); the few buffers a peer GPU touches are exported through CUDA IPC (hg_slab.cuh).
  int world = 1, rank = 0, k0 = 0, k1 = 0, nzg = 1;
  long long nxy = 0, ncg = 0;
  Slab slab;
  bool initialised = false;
  std::vector<void*> ipc_opened;
  bool profile_on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev[2];   // [0] pressure sweeps kernel, [1] lu kernel
  cudaEvent_t user_ev[8] = {};
  std::string err;
};

static thread_local std::string g_create_err;
static const bool g_trace = getenv("HYDRO_SLAB_TRACE") != nullptr;
#define TRACE(...) do { if (g_trace) { fprintf(stderr, __VA_ARGS__); fflush(stderr); } } while (0)

#define CK(call)                                                                   \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      s->err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
      return HG_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

static inline unsigned nblk(long long n, int t = 256) { return (unsigned)((n + t - 1) / t); }

template <class T>
static int dalloc(hg_state* s, T** p, long long n, bool zero = true) {
  void* q = nullptr;
  CK(cudaMalloc(&q, (size_t)(n > 0 ? n : 1) * sizeof(T)));
  if (zero) CK(cudaMemsetAsync(q, 0, (size_t)(n > 0 ? n : 1) * sizeof(T), s->st));
  s->allocs.push_back(q);
  *p = (T*)q;
  return 0;
}

static void tpush(hg_state* s, const char* name) {
  if (!s->timers_on) return;
  Timer& t = s->timers[name];
  if (!t.a) { cudaEventCreate(&t.a); cudaEventCreate(&t.b); }
  cudaEventRecord(t.a, s->st);
  t.open = true;
  s->timer_stack.push_back(name);
}
static void tpop(hg_state* s) {
  if (!s->timers_on || s->timer_stack.empty()) return;
  Timer& t = s->timers[s->timer_stack.back()];
  s->timer_stack.pop_back();
  cudaEventRecord(t.b, s->st);
  cudaEventSynchronize(t.b);
  float ms = 0.f; cudaEventElapsedTime(&ms, t.a, t.b);
  t.total += ms * 1e-3; t.open = false;
}

#define LAUNCH(s, kern, grid, block, ...)                      \
  do { kern<<<grid, block, 0, (s)->st>>>(__VA_ARGS__); ++(s)->launches; } while (0)
#define DIMSEL(s, KERN, grid, block, ...)                                                     \
  do { if ((s)->dim == 3) { KERN<3><<<grid, block, 0, (s)->st>>>(__VA_ARGS__); }             \
       else { KERN<2><<<grid, block, 0, (s)->st>>>(__VA_ARGS__); } ++(s)->launches; } while (0)

static Geo list_geo(const hg_state* s, int R = 1) { Geo g = s->geo; g.cells = R == 1 ? s->slow_list : s->slow2_list; g.ncells = R == 1 ? s->nslow : s->nslow2; return g; }
static dim3 fast_grid(const hg_state* s) { return dim3((s->n[0] + 31) / 32, (s->n[2] + FT_K - 1) / FT_K, s->n[1]); }
static CP3 cp3(double* const a[3]) { CP3 r; for (int d = 0; d < 3; ++d) r.p[d] = a[d]; return r; }
static P3 p3(double* const a[3]) { P3 r; for (int d = 0; d < 3; ++d) r.p[d] = a[d]; return r; }

// ------------------------------------------------------------------ z-slab plumbing (hg_slab.cuh)
static int slab_exchange(hg_state* s, double* const* arrs, int n, int planes) {
  if (s->world <= 1) return 0;
  if (!s->slab.linked) { s->err = "multi-GPU handle used before hg_ipc_import"; return HG_ERR_INVALID; }
  if (n > SLAB_MAX_ARRAYS || planes > HG_HALO) { s->err = "slab_exchange: too many arrays/planes"; return HG_ERR_INVALID; }
  Slab& sl = s->slab;
  const unsigned long long v = ++sl.xseq;
  const int parity = (int)(v & 1ull);
  TRACE("[r%d] exchange %llu arrays %d planes %d\n", s->rank, v, n, planes);
  PackArgs pa; UnpackArgs ua; pa.n = ua.n = n; pa.planes = ua.planes = planes;
  for (int q = 0; q < n; ++q) { pa.src[q] = arrs[q]; ua.dst[q] = arrs[q]; }
  const long long total = (long long)n * planes * s->nxy;
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148 * 8);
  k_slab_pack<<<blocks, 256, 0, s->st>>>(s->geo, pa, sl.xbuf, parity, sl.has_lo, sl.has_hi);
  unsigned long long* f_lo = sl.has_lo ? slab_flags(sl.mail_peer[s->rank - 1], s->world) + SF_X_HI : nullptr;
  unsigned long long* f_hi = sl.has_hi ? slab_flags(sl.mail_peer[s->rank + 1], s->world) + SF_X_LO : nullptr;
  k_slab_signal<<<1, 1, 0, s->st>>>(f_lo, f_hi, v);
  k_slab_wait<<<1, 1, 0, s->st>>>(sl.has_lo, sl.has_hi, slab_flags(sl.mail, s->world), v);
  k_slab_unpack<<<blocks, 256, 0, s->st>>>(s->geo, ua, sl.has_lo ? sl.xbuf_lo : nullptr, sl.has_hi ? sl.xbuf_hi : nullptr, parity);
  s->launches += 4;
  return 0;
}
static int slab_exchange(hg_state* s, std::initializer_list<double*> arrs, int planes) {
  std::vector<double*> v;
  for (double* q : arrs) if (q) v.push_back(q);
  return slab_exchange(s, v.data(), (int)v.size(), planes);
}
// k_gs_tiled on slabs: rows of the first GT_B - 1 planes of every slab -> ghost hyperplanes of the slab below (same handshake
// as slab_exchange: pack, flags, pull)
static int gt_ghost_rows(hg_state* s) {
  if (s->world <= 1 || !s->gs_tiled || GT_GHOST == 0) return 0;
  if (!s->slab.linked) { s->err = "multi-GPU handle used before hg_ipc_import"; return HG_ERR_INVALID; }
  Slab& sl = s->slab;
  const unsigned long long v = ++sl.xseq;
  const int parity = (int)(v & 1ull);
  const unsigned blocks = nblk(5LL * GT_GHOST * s->nxy);
  if (sl.has_lo) { k_gt_ghost_pack<<<blocks, 256, 0, s->st>>>(s->geo, s->CO, s->co5, sl.xbuf, parity); ++s->launches; }
  unsigned long long* f_lo = sl.has_lo ? slab_flags(sl.mail_peer[s->rank - 1], s->world) + SF_X_HI : nullptr;
  unsigned long long* f_hi = sl.has_hi ? slab_flags(sl.mail_peer[s->rank + 1], s->world) + SF_X_LO : nullptr;
  k_slab_signal<<<1, 1, 0, s->st>>>(f_lo, f_hi, v);
  k_slab_wait<<<1, 1, 0, s->st>>>(sl.has_lo, sl.has_hi, slab_flags(sl.mail, s->world), v);
  if (sl.has_hi) { k_gt_ghost_unpack<<<blocks, 256, 0, s->st>>>(s->geo, s->CO, s->co5, sl.xbuf_hi, parity); ++s->launches; }
  s->launches += 2;
  return 0;
}
// all-gather of n device doubles from every rank into host memory out[world][n] (stream-synchronising)
static int slab_gather(hg_state* s, const double* dev_vals, int n, std::vector<double>& out) {
  out.assign((size_t)s->world * n, 0.);
  if (s->world <= 1) {
    CK(cudaMemcpyAsync(out.data(), dev_vals, n * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CK(cudaStreamSynchronize(s->st));
    return 0;
  }
  if (!s->slab.linked) { s->err = "multi-GPU handle used before hg_ipc_import"; return HG_ERR_INVALID; }
  if (n > SLAB_MAIL) { s->err = "slab_gather: too many values"; return HG_ERR_INVALID; }
  Slab& sl = s->slab;
  const unsigned long long v = ++sl.mseq;
  const int parity = (int)(v & 1ull);
  TRACE("[r%d] gather %llu n %d\n", s->rank, v, n);
  MailPeers mp; for (int r = 0; r < s->world; ++r) mp.mail[r] = sl.mail_peer[r];
  k_mail_post<<<1, 1024, 0, s->st>>>(mp, dev_vals, n, parity, s->rank, s->world, v);
  k_mail_wait<<<1, 64, 0, s->st>>>(sl.mail, s->world, v);
  s->launches += 2;
  for (int r = 0; r < s->world; ++r)
    CK(cudaMemcpyAsync(out.data() + (size_t)r * n, sl.mail + ((long long)parity * s->world + r) * SLAB_MAIL, n * sizeof(double),
                       cudaMemcpyDeviceToHost, s->st));
  unsigned long long perr = 0;
  CK(cudaMemcpyAsync(&perr, slab_flags(sl.mail, s->world) + SF_ERR, sizeof perr, cudaMemcpyDeviceToHost, s->st));
  CK(cudaStreamSynchronize(s->st));
  TRACE("[r%d] gather %llu done err %llu\n", s->rank, v, perr);
  if (perr) { s->err = "slab decomposition: a wait for a neighbouring rank timed out"; return HG_ERR_CUDA; }
  return 0;
}
// reduction of n device doubles over the ranks, done on the host in rank order (identical on every rank);
// op: 0 = max, 1 = min, 2 = sum.  Result in host memory `out`.
static int slab_reduce(hg_state* s, const double* dev_vals, int n, int op, double* out) {
  std::vector<double> all;
  if (int rc = slab_gather(s, dev_vals, n, all)) return rc;
  for (int q = 0; q < n; ++q) {
    double v = all[q];
    for (int r = 1; r < s->world; ++r) {
      const double w = all[(size_t)r * n + q];
      v = op == 0 ? (v < w ? w : v) : op == 1 ? (w < v ? w : v) : v + w;
    }
    out[q] = v;
  }
  return 0;
}
// all ranks have executed everything enqueued so far (no-op on one GPU)
static int slab_sync(hg_state* s) {
  if (s->world <= 1) return 0;
  double dummy[1];
  return slab_reduce(s, s->scal + 60, 1, 0, dummy);
}
#define XCH(s, planes, ...) \
  do { if ((s)->world > 1) { if (int rc_ = slab_exchange((s), {__VA_ARGS__}, (planes))) return rc_; } } while (0)
// what a linked solver kernel needs: which=0 Gauss-Seidel planes (tag0 = solve), which=1 lu planes (tag0 = lu solve)
static SlabLink slab_link(hg_state* s, int which) {
  SlabLink L; memset(&L, 0, sizeof L);
  if (s->world <= 1) return L;
  Slab& sl = s->slab;
  const long long nxy = s->nxy;
  L.on = 1; L.has_lo = sl.has_lo; L.has_hi = sl.has_hi; L.k0 = s->k0; L.np_glob = sl.np_glob; L.nz_lo = sl.nz_lo;
  L.from_lo = sl.ll + (which ? 2 : 0) * nxy;
  L.from_hi = sl.ll + (which ? 5 : 1) * nxy;
  L.to_lo = sl.has_lo ? sl.ll_lower + (which ? 5 : 1) * nxy : nullptr;
  L.to_hi = sl.has_hi ? sl.ll_upper + (which ? 2 : 0) * nxy : nullptr;
  L.up_from = sl.ll + SLAB_LL_UP * nxy; L.down_from = sl.ll + SLAB_LL_DOWN * nxy;
  L.up_to = sl.has_hi ? sl.ll_upper + SLAB_LL_UP * nxy : nullptr;
  L.down_to = sl.has_lo ? sl.ll_lower + SLAB_LL_DOWN * nxy : nullptr;
  L.gbase = s->gt_gbase;
  L.tag0 = which ? sl.lu_seq * 4u : sl.solve_seq * 2048u;
  L.err = slab_flags(sl.mail, s->world) + SF_ERR;
  return L;
}

// GetDerivativeApproxCoeffs (solver.hpp:816-861), args {-2dt,-dt,0}, target 0
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

// ------------------------------------------------------------------ solvers (host side)
template <class K, class A>
static int coop_launch(hg_state* s, K kern, int grid, Geo g, A args, int prof_slot = -1) {
  void* params[] = {(void*)&g, (void*)&args};
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  CK(cudaMemsetAsync(s->tt.bar, 0, 4 * sizeof(unsigned long long), s->st));   // barrier arrival counter + release word
  if (s->profile_on && prof_slot >= 0) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, s->st); }
  // Cooperative launch = guaranteed co-residency of the grid (the kernels use their own grid barrier).  Ranks that
  // share a device (solver_ctas > 0, tests) launch normally: two cooperative grids are not run side by side, and
  // the limited grids fit next to each other.
  if (s->cfg.solver_ctas > 0) { CK(cudaLaunchKernel((void*)kern, dim3(grid), dim3(SOLVER_THREADS), params, 0, s->st)); }
  else CK(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(SOLVER_THREADS), params, 0, s->st));
  if (e0) { cudaEventRecord(e1, s->st); s->prof_ev[prof_slot].push_back({e0, e1}); }
  ++s->launches;
  return 0;
}

static int ensure_sweep_capacity(hg_state* s, int nsweeps) {
  if (nsweeps <= s->max_sweeps) return 0;
  int cap = nsweeps + 16;
  CK(cudaMalloc((void**)&s->diffs, cap * sizeof(double)));
  s->allocs.push_back(s->diffs);
  CK(cudaMallocHost((void**)&s->hdiffs, cap * sizeof(double)));
  s->max_sweeps = cap;
  return 0;
}

// Runs `do { sweep } while (diff > tol && iter++ < limit)` (linear.hpp:688-710) with pipelined sweeps.
// launch(s_begin, s_end) must run sweeps [s_begin, s_end) on x; save/restore checkpoint x when a
// chunk overshoots the stopping sweep.  Returns the reference's `iter` and `diff`.
// defer_out != nullptr (one GPU, tol == 0): nothing is read back -- {iter, diff} are written to defer_out on the device
// (k_sor_result) and *out_iter = -1.
template <class LaunchFn>
static int run_sor(hg_state* s, double* x, long long nx_, double tol, int limit, LaunchFn launch, int* out_iter, double* out_diff,
                   double* defer_out = nullptr) {
  const int max_total = limit + 1;
  if (int rc = ensure_sweep_capacity(s, max_total)) return rc;
  CK(cudaMemsetAsync(x, 0, nx_ * sizeof(double), s->st));
  CK(cudaMemsetAsync(s->diffs, 0, max_total * sizeof(double), s->st));
  // slabs: the interface planes start as "value 0 before the first sweep"; a neighbour's first sweep step stores
  // into them, so it must not start before the reset has been executed here
  auto ll_reset = [&]() -> int {
    if (s->world <= 1) return 0;
    ++s->slab.solve_seq;
    k_ll_fill<<<nblk(2 * s->nxy), 256, 0, s->st>>>(s->slab.ll, 2 * s->nxy, s->slab.solve_seq * 2048u);
    // box-dataflow sweeps: the ghost planes of the first group start from the initial guess 0
    k_ll_fill<<<nblk(2LL * SLAB_GB * s->nxy), 256, 0, s->st>>>(s->slab.ll + SLAB_LL_DOWN * s->nxy, 2LL * SLAB_GB * s->nxy, s->slab.solve_seq * 2048u);
    s->gt_gbase = 0;
    s->launches += 2;
    return slab_sync(s);
  };
  if (int rc = ll_reset()) return rc;
  if (!(tol > 0.) && defer_out && s->world == 1) {
    // `diff > tol` only fails for diff == 0: from that sweep on every correction is exactly zero, so running all
    // limit+1 sweeps leaves the same solution; the count the reference would report is derived on the device
    for (int sb = 0; sb < max_total; sb += SOLVER_SC) if (int rc = launch(sb, std::min(sb + SOLVER_SC, max_total))) return rc;
    LAUNCH(s, k_sor_result, 1, 1, s->diffs, max_total, tol, defer_out);
    *out_iter = -1; *out_diff = 0.;
    return 0;
  }
  if (!(tol > 0.)) {
    // `diff > tol` only fails for diff == 0 (or NaN): run all limit+1 sweeps, inspect the history once
    for (int sb = 0; sb < max_total; sb += SOLVER_SC) {
      const int se = std::min(sb + SOLVER_SC, max_total);
      if (int rc = launch(sb, se)) return rc;
      if (int rc = slab_reduce(s, s->diffs + sb, se - sb, 0, s->hdiffs + sb)) return rc;
    }
    int stop = -1;
    for (int k = 0; k < max_total; ++k) if (!(s->hdiffs[k] > tol)) { stop = k; break; }
    if (stop >= 0 && stop < max_total - 1) {   // stopped early: redo with the exact sweep count
      const double dstop = s->hdiffs[stop];
      CK(cudaMemsetAsync(x, 0, nx_ * sizeof(double), s->st));
      CK(cudaMemsetAsync(s->diffs, 0, max_total * sizeof(double), s->st));
      if (int rc = ll_reset()) return rc;
      for (int sb = 0; sb < stop + 1; sb += SOLVER_SC) if (int rc = launch(sb, std::min(sb + SOLVER_SC, stop + 1))) return rc;
      *out_iter = stop; *out_diff = dstop;
      return 0;
    }
    *out_iter = stop >= 0 ? stop : limit + 1;   // `iter++ < limit` increments even when it fails
    *out_diff = s->hdiffs[max_total - 1];
    return 0;
  }
  int chunk = s->cfg.pressure_sweeps_per_check > 0 ? s->cfg.pressure_sweeps_per_check : 128;
  if (chunk > SOLVER_SC) chunk = SOLVER_SC;
  // The stopping sweep is only known after the fact, and a chunk that runs past it is replayed from its checkpoint with the
  // exact count: fixed chunks of 128 sweeps execute up to twice the sweeps the reference does.  Prediction: the stopping sweep of
  // the same SIMPLE iteration of the previous time step, scaled by the trend seen since (iteration q-1 of this step against
  // iteration q-1 of the previous step; for the first iteration the previous step against the one before) -- during the start-up
  // of a flow the counts fall by a third from step to step, and a prediction that is too high costs a whole chunk (the stop lies
  // inside it: replay), one that is too low only a few short chunks.  So phase 1 runs to a fraction of the prediction in chunks of
  // up to `chunk` sweeps, phase 2 continues in short chunks of constant size.  Without a prediction: `chunk`.
  const int pidx = std::min(std::max(s->iter_count, 0), (int)s->sor_pred.size() - 1);
  const int pred = s->sor_pred[pidx];
  static const int margin = getenv("HYDRO_SOR_MARGIN") ? std::max(0, atoi(getenv("HYDRO_SOR_MARGIN"))) : 4;
  static const int small = getenv("HYDRO_SOR_SMALL") ? std::max(1, atoi(getenv("HYDRO_SOR_SMALL"))) : 16;
  static const double frac = getenv("HYDRO_SOR_FRAC") ? atof(getenv("HYDRO_SOR_FRAC")) : 0.7;
  static const int probe = getenv("HYDRO_SOR_PROBE") ? std::max(4, atoi(getenv("HYDRO_SOR_PROBE"))) : 32;
  static const bool sor_trace = getenv("HYDRO_SOR_TRACE") != nullptr;   // diagnostics: prediction, stopping sweep, sweeps executed
  // first chunk: a probe of at most 32 sweeps (shorter when the history with its trend says so: the counts can drop abruptly
  // from one step to the next, and a first chunk that contains the stop is replayed whole); every further chunk from the decay of
  // the norm inside this solve -- the last 16 sweeps extrapolated to the tolerance, a fraction of it, at least `small` sweeps
  int n_next = std::min(chunk, probe);
  if (pred >= 0) {
    const int tq = pidx > 0 ? pidx - 1 : 0;   // where the trend is read
    if (s->sor_pred[tq] >= 0 && s->sor_old[tq] > 0) {
      const double trend = std::min(1.25, std::max(0.25, (double)(s->sor_pred[tq] + 1) / (s->sor_old[tq] + 1)));
      const int hist = std::max(small, std::max(0, (int)(frac * trend * (pred + 1)) - margin) & ~(GT_B - 1));   // whole sweep groups
      n_next = std::min(n_next, hist);
    }
  }
  int done = 0, executed = 0;
  while (done < max_total) {
    int n = std::min(std::min(n_next, chunk), max_total - done);
    CK(cudaMemcpyAsync(s->PPsave, x, nx_ * sizeof(double), cudaMemcpyDeviceToDevice, s->st));
    const int gbase_save = s->gt_gbase;
    if (s->world > 1) {
      CK(cudaMemcpyAsync(s->slab.ll + 8 * s->nxy, s->slab.ll, 2 * s->nxy * sizeof(uint4), cudaMemcpyDeviceToDevice, s->st));
      CK(cudaMemcpyAsync(s->slab.ll + SLAB_LL_DOWN_SAVE * s->nxy, s->slab.ll + SLAB_LL_DOWN * s->nxy, 2LL * SLAB_GB * s->nxy * sizeof(uint4), cudaMemcpyDeviceToDevice, s->st));
    }
    if (int rc = launch(done, done + n)) return rc;
    executed += n;
    if (int rc = slab_reduce(s, s->diffs + done, n, 0, s->hdiffs + done)) return rc;
    int stop = -1;
    for (int k = done; k < done + n; ++k) if (!(s->hdiffs[k] > tol)) { stop = k; break; }
    if (stop >= 0) {
      if (stop != done + n - 1) {
        CK(cudaMemcpyAsync(x, s->PPsave, nx_ * sizeof(double), cudaMemcpyDeviceToDevice, s->st));
        if (s->world > 1) {
          CK(cudaMemcpyAsync(s->slab.ll, s->slab.ll + 8 * s->nxy, 2 * s->nxy * sizeof(uint4), cudaMemcpyDeviceToDevice, s->st));
          CK(cudaMemcpyAsync(s->slab.ll + SLAB_LL_DOWN * s->nxy, s->slab.ll + SLAB_LL_DOWN_SAVE * s->nxy, 2LL * SLAB_GB * s->nxy * sizeof(uint4), cudaMemcpyDeviceToDevice, s->st));
          s->gt_gbase = gbase_save;
        }
        CK(cudaMemsetAsync(s->diffs + done, 0, n * sizeof(double), s->st));
        if (int rc = slab_sync(s)) return rc;
        if (int rc = launch(done, stop + 1)) return rc;
        executed += stop + 1 - done;
      }
      if (sor_trace) fprintf(stderr, "sor: iteration %d predicted %d stopped at %d, %d sweeps executed for %d\n", s->iter_count, pred, stop, executed, stop + 1);
      *out_iter = stop; *out_diff = s->hdiffs[stop];
      s->sor_old[pidx] = s->sor_pred[pidx]; s->sor_pred[pidx] = stop;
      return 0;
    }
    done += n;
    n_next = small;
    { const int k1 = done - 1, k0 = std::max(0, k1 - 16);
      const double d1 = s->hdiffs[k1], d0 = s->hdiffs[k0];
      if (k1 > k0 && d1 > tol && d0 > d1) {
        const double rem = std::log(tol / d1) / (std::log(d1 / d0) / (k1 - k0));   // sweeps still to go at this decay
        if (rem > 0. && rem < 1e6) n_next = std::max(small, (int)(frac * rem) & ~(GT_B - 1));
      } }
  }
  *out_iter = limit + 1; *out_diff = s->hdiffs[max_total - 1];
  s->sor_old[pidx] = s->sor_pred[pidx]; s->sor_pred[pidx] = max_total - 1;
  return 0;
}

// Jacobi::Solve (linear.hpp:750-782): independent sweeps on the natural layout; the stop test needs the
// per-sweep norm, fetched every `check` sweeps; the state is recomputed from a checkpoint when the
// stopping sweep falls inside a batch.  rows == nullptr: pressure rows regenerated from d_c.  Result in s->pc.
static int run_jacobi(hg_state* s, double* const* rows, const double* D, const double* R, double tol, int limit, double omega,
                      int* out_iter, double* out_diff) {
  const int check = 16;
  if (int rc = ensure_sweep_capacity(s, check + 2)) return rc;
  double* xc = s->pc; double* xo = s->w2;
  CK(cudaMemsetAsync(xc, 0, s->nc * sizeof(double), s->st));
  auto sweeps = [&](int nsw) -> int {
    double* a_ = xc; double* b_ = xo;
    for (int q = 0; q < nsw; ++q) {
      JacArgs a; for (int t = 0; t < 7; ++t) a.A[t] = rows ? rows[t] : nullptr;
      a.D = D; a.R = R; a.xin = a_; a.xout = b_; a.diff = s->diffs + q; a.omega = omega;
      if (s->world > 1) { if (int rc = slab_exchange(s, {a_}, 1)) return rc; }   // slabs: the iterate's values across the interfaces
      DIMSEL(s, k_jacobi_sweep, nblk(s->nc), 256, s->geo, a);
      std::swap(a_, b_);
    }
    return 0;
  };
  int done = 0;
  for (;;) {
    const int nb = std::min(check, limit + 1 - done);
    CK(cudaMemcpyAsync(s->PPsave, xc, s->nc * sizeof(double), cudaMemcpyDeviceToDevice, s->st));
    CK(cudaMemsetAsync(s->diffs, 0, nb * sizeof(double), s->st));
    if (int rc = sweeps(nb)) return rc;
    if (int rc = slab_reduce(s, s->diffs, nb, 0, s->hdiffs)) return rc;   // (one GPU: a plain read-back)
    int nsw = -1;
    for (int q = 0; q < nb; ++q) {
      const int k = done + q;
      if (!(s->hdiffs[q] > tol)) { *out_iter = k; *out_diff = s->hdiffs[q]; nsw = q + 1; break; }
      if (!(k < limit)) { *out_iter = k + 1; *out_diff = s->hdiffs[q]; nsw = q + 1; break; }
    }
    if (nsw < 0) { done += nb; if (nb & 1) std::swap(xc, xo); continue; }
    if (nsw != nb) {
      CK(cudaMemcpyAsync(xc, s->PPsave, s->nc * sizeof(double), cudaMemcpyDeviceToDevice, s->st));
      CK(cudaMemsetAsync(s->diffs, 0, nb * sizeof(double), s->st));
      if (int rc = sweeps(nsw)) return rc;
    }
    if (nsw & 1) std::swap(xc, xo);
    break;
  }
  if (xc != s->pc) CK(cudaMemcpyAsync(s->pc, xc, s->nc * sizeof(double), cudaMemcpyDeviceToDevice, s->st));
  return 0;
}

static int shear_arrays(hg_state* s, double* const* in, double* const* out, int n);
// LuDecompositionRelaxed::Solve (linear.hpp:592-650) on sheared rows A, constants R; result in `res` (sheared).
// Scratch: corr, f (sheared).  One host round trip per outer iteration (the stop test needs max|corr|).
static int run_lu_relaxed(hg_state* s, const double* R, double* res, double* corr, double* f, double tol, int limit, double relax,
                          int* out_iter, double* out_diff) {
  if (int rc = ensure_sweep_capacity(s, 4)) return rc;
  CK(cudaMemsetAsync(res, 0, s->nsh * sizeof(double), s->st));
  CK(cudaMemsetAsync(corr, 0, s->nsh * sizeof(double), s->st));
  CK(cudaMemcpyAsync(f, R, s->nsh * sizeof(double), cudaMemcpyDeviceToDevice, s->st));
  LurArgs a; for (int t = 0; t < 7; ++t) a.A[t] = s->A[t];
  a.F = f; a.corr = corr; a.relax = relax; a.tt = s->tt;
  const unsigned gb = nblk(s->nc);
  size_t iter = 0; double diff = 0.;
  do {
    if (s->dim == 3) { if (int rc = coop_launch(s, k_lur_forward<3>, s->grid_lu, s->geo, a)) return rc; }
    else { if (int rc = coop_launch(s, k_lur_forward<2>, s->grid_lu, s->geo, a)) return rc; }
    DIMSEL(s, k_lur_backward, gb, 256, s->geo, a, res);
    CK(cudaMemsetAsync(s->diffs, 0, sizeof(double), s->st));
    DIMSEL(s, k_lur_update, gb, 256, s->geo, corr, res, s->diffs);
    DIMSEL(s, k_lur_residual, gb, 256, s->geo, a, R, res, f);
    CK(cudaMemcpyAsync(s->hdiffs, s->diffs, sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CK(cudaStreamSynchronize(s->st));
    diff = s->hdiffs[0];
  } while (diff > tol && iter++ < (size_t)limit);
  *out_iter = (int)iter; *out_diff = diff;
  return 0;
}

// Task list of k_gs_tiled for a launch of S sweeps: boxes (I, J, group) sorted by the step at which they can start,
// TX I + TY J + (TX + TY + 2B + 2) group (a box starts TX / TY steps after its left / lower neighbour and 2B+1 steps after the box
// (I+1, J+1) of the previous group), which also puts every dependency of a box before it (own group: (I-1,J), (I,J-1), (I-1,J-1); previous group: (I..I+1, J..J+1)).
static int gt_num_tasks(const hg_state* s, int S) {
  const int nx = s->n[0], ny = s->n[1];
  const int NG = (S + GT_B - 1) / GT_B, NI = (nx + GT_B - 1 + GT_TX - 1) / GT_TX, NJ = (ny + GT_B - 1 + GT_TY - 1) / GT_TY;
  return NG * NI * NJ;   // upper bound (boxes without cells are dropped)
}
static int gt_build_tasks(hg_state* s, int S, GtTask* tasks, int* ntasks) {
  const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
  const int NG = (S + GT_B - 1) / GT_B, NI = (nx + GT_B - 1 + GT_TX - 1) / GT_TX, NJ = (ny + GT_B - 1 + GT_TY - 1) / GT_TY;
  struct Key { int w, gI, J, I; };
  std::vector<Key> keys;
  auto nsw_of = [&](int gI) { return std::min(GT_B, S - gI * GT_B); };
  auto exists = [&](int I, int J, int gI) {
    if (I < 0 || J < 0 || gI < 0 || I >= NI || J >= NJ || gI >= NG) return false;
    const int nsw = nsw_of(gI);
    return I * GT_TX - (nsw - 1) < nx && J * GT_TY - (nsw - 1) < ny;
  };
  // Claim order: by a weighted start step WI I + WJ J + WG group.  The natural weights (TX, TY, TX + TY + 2B + 2) claim the
  // boxes in the order in which they can start.  Any weights with WG > WI + WJ put every dependency of a box before it.
  // A smaller WJ brings the boxes (I, J) and (I, J+1), whose (TY + B)-row footprints share B rows, closer together in
  // time, so that the shared rows are still in L2 when the second box asks for them.
  static const int WJ = getenv("HYDRO_GT_WJ") ? std::max(0, atoi(getenv("HYDRO_GT_WJ"))) : GT_TY;
  static const int WI = getenv("HYDRO_GT_WI") ? std::max(1, atoi(getenv("HYDRO_GT_WI"))) : GT_TX;
  // WG: the natural weight WI + WJ + 2B + 2 interleaves about ten groups (spatial neighbours of one group are then ~100 steps
  // apart in time and their shared rows have left L2); a large WG claims group by group
  static const int WG0 = getenv("HYDRO_GT_WG") ? atoi(getenv("HYDRO_GT_WG")) : 0;
  const int WG = std::max(WG0, WI + WJ + 2 * GT_B + 2);
  for (int gI = 0; gI < NG; ++gI) for (int I = 0; I < NI; ++I) for (int J = 0; J < NJ; ++J)
    if (exists(I, J, gI)) keys.push_back({WI * I + WJ * J + WG * gI, gI, J, I});
  std::stable_sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) { return a.w < b.w; });
  std::vector<int> index((size_t)NG * NJ * NI, -1);
  auto at = [&](int I, int J, int gI) -> int { return exists(I, J, gI) ? index[((size_t)gI * NJ + J) * NI + I] : -1; };
  for (size_t q = 0; q < keys.size(); ++q) index[((size_t)keys[q].gI * NJ + keys[q].J) * NI + keys[q].I] = (int)q;
  for (size_t q = 0; q < keys.size(); ++q) {
    const Key& k = keys[q];
    GtTask& t = tasks[q];
    t.I0 = k.I * GT_TX; t.J0 = k.J * GT_TY; t.s0 = k.gI * GT_B; t.nsw = nsw_of(k.gI);
    t.Tlo = std::max(0, t.I0 - (t.nsw - 1)) + std::max(0, t.J0 - (t.nsw - 1)) - 2;
    t.Thi = std::min(nx - 1, t.I0 + GT_TX - 1) + std::min(ny - 1, t.J0 + GT_TY - 1) + (nz - 1) + 2 * (t.nsw - 1);
    t.dep[0] = at(k.I - 1, k.J, k.gI); t.dep[1] = at(k.I, k.J - 1, k.gI); t.dep[2] = at(k.I - 1, k.J - 1, k.gI);
    t.dep[3] = at(k.I, k.J, k.gI - 1); t.dep[4] = at(k.I + 1, k.J, k.gI - 1);
    t.dep[5] = at(k.I, k.J + 1, k.gI - 1); t.dep[6] = at(k.I + 1, k.J + 1, k.gI - 1);
    for (int d = 0; d < GT_MAXDEP; ++d) if (t.dep[d] >= (int)q) { s->err = "gt_plan: dependency order violated"; return HG_ERR_INVALID; }
  }
  *ntasks = (int)keys.size();
  return 0;
}
// largest number of sweeps a launch of this handle can have (run_sor chunks)
static int gt_max_launch_sweeps(const hg_state* s) {
  const int max_total = s->cfg.lu_relaxed_num_iters_limit + 1;
  int chunk = SOLVER_SC;
  if (s->cfg.lu_relaxed_tolerance > 0.) {
    chunk = s->cfg.pressure_sweeps_per_check > 0 ? s->cfg.pressure_sweeps_per_check : 128;
    if (chunk > SOLVER_SC) chunk = SOLVER_SC;
  }
  return std::max(1, std::min(chunk, max_total));
}
static int gt_plans_allocate(hg_state* s) {   // at creation
  s->gt_plan_cap = gt_num_tasks(s, gt_max_launch_sweeps(s));
  for (auto& pl : s->gt_plans) {
    if (dalloc(s, &pl.tasks, s->gt_plan_cap, false) || dalloc(s, &pl.progress, s->gt_plan_cap, true)) return HG_ERR_CUDA;
    CK(cudaMallocHost((void**)&pl.stage, (size_t)s->gt_plan_cap * sizeof(GtTask)));
    CK(cudaEventCreateWithFlags(&pl.staged, cudaEventDisableTiming));
  }
  return 0;
}
static int gt_plan(hg_state* s, int S, hg_state::GtPlan** out) {
  hg_state::GtPlan* lru = &s->gt_plans[0];
  for (auto& pl : s->gt_plans) {
    if (pl.S == S) { pl.used = ++s->gt_plan_clock; *out = &pl; return 0; }
    if (pl.used < lru->used) lru = &pl;
  }
  if (gt_num_tasks(s, S) > s->gt_plan_cap) { s->err = "gt_plan: launch larger than the preallocated task list"; return HG_ERR_INVALID; }
  // the pinned staging buffer of this slot may still be read by its previous upload
  CK(cudaEventSynchronize(lru->staged));
  if (int rc = gt_build_tasks(s, S, lru->stage, &lru->ntasks)) return rc;
  CK(cudaMemcpyAsync(lru->tasks, lru->stage, (size_t)lru->ntasks * sizeof(GtTask), cudaMemcpyHostToDevice, s->st));
  CK(cudaEventRecord(lru->staged, s->st));
  lru->S = S; lru->used = ++s->gt_plan_clock;
  *out = lru;
  return 0;
}
static int gt_launch(hg_state* s, int sb, int se, double omega) {
  hg_state::GtPlan* pl = nullptr;
  if (int rc = gt_plan(s, se - sb, &pl)) return rc;
  GtArgs a; a.PP = s->PP; a.diff = s->diffs;
  a.s_begin = sb; a.omega = omega; a.tasks = pl->tasks; a.ntasks = pl->ntasks; a.progress = pl->progress; a.ctl = s->gt_ctl;
  a.lag_prev = 2 * GT_B + 1;
  a.PS8 = 8LL * s->n[0] * s->n[1]; a.DSH8 = 8LL * (2LL * s->n[0] * s->n[1] + s->n[0] + 1);
  a.clk = s->gt_clk;
  CK(cudaMemsetAsync(pl->progress, 0, pl->ntasks * sizeof(int), s->st));
  CK(cudaMemsetAsync(s->gt_ctl, 0, sizeof(int), s->st));   // next-task counter; the abort flag [1] is sticky
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (s->profile_on) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, s->st); }
  int grid = std::min(pl->ntasks, s->num_sms * GT_CTAS_PER_SM);
  if (s->cfg.solver_ctas > 0) grid = std::min(grid, s->cfg.solver_ctas);   // ranks sharing a device (tests)
  a.link = slab_link(s, 0);
  static const bool force_link = getenv("HYDRO_GT_FORCE_LINK") != nullptr;   // diagnostics: the slab instantiation on one GPU
  if (s->world > 1 || force_link) k_gs_tiled<true><<<grid, GT_BLOCK, GT_SMEM_BYTES, s->st>>>(s->geo, a, s->tmco);
  else k_gs_tiled<false><<<grid, GT_BLOCK, GT_SMEM_BYTES, s->st>>>(s->geo, a, s->tmco);
  CK(cudaGetLastError());
  if (e0) { cudaEventRecord(e1, s->st); s->prof_ev[0].push_back({e0, e1}); }
  ++s->launches;
  s->gt_gbase += (se - sb + GT_B - 1) / GT_B;
  return 0;
}
static int gt_check(hg_state* s) {   // after the sweeps: did a dependency wait time out?
  int h = 0;
  CK(cudaMemcpyAsync(&h, s->gt_ctl + 1, sizeof(int), cudaMemcpyDeviceToHost, s->st));
  CK(cudaStreamSynchronize(s->st));
  if (h) { s->err = "k_gs_tiled: dependency wait timed out"; return HG_ERR_CUDA; }
  return 0;
}

static int lt_check(hg_state* s);
static int solve_system(hg_state* s, int solver, double* R, double* X, double* t1, double* t2, int* it, double* df);
static void launch_pcorr(hg_state* s) {   // p' back to the natural layout, p_curr = p_prev + alpha p'
  const double alpha = s->cfg.pressure_relaxation_factor;
  if (s->dim == 3) k_ft_pcorr<<<fast_grid(s), FT_THREADS, 0, s->st>>>(s->geo, s->PP, s->p[L_IP], alpha, s->pc, s->p[L_IC]);
  else k_pcorr<2><<<nblk(s->nc), 256, 0, s->st>>>(s->geo, s->PP, s->p[L_IP], alpha, s->pc, s->p[L_IC]);
  ++s->launches;
}
// second = the SIMPLER solve (fluid.hpp:1148-1152): same rows, new constants, the result is added to the pressure unrelaxed
static void launch_padd(hg_state* s) {
  DIMSEL(s, k_simpler_padd, nblk(s->nc), 256, s->geo, s->PP, s->p[L_IC]);
}
static int solve_pressure(hg_state* s, bool second = false) {
  const hg_config& c = s->cfg;
  int it = 0; double df = 0.;
  if (c.linear_solver_pressure == HG_LS_GAUSS_SEIDEL) {
    auto launch = [&](int sb, int se) -> int {
      if (s->gs_tiled) return gt_launch(s, sb, se, c.lu_relaxed_relaxation_factor);
      GsArgs a{}; a.CX = s->D; a.CY = s->CYs; a.CZ = s->CZs; a.RP = s->RP; a.PP = s->PP; a.diff = s->diffs; a.s_begin = sb; a.s_end = se;
      a.omega = c.lu_relaxed_relaxation_factor; a.tt = s->tt;
      a.link = slab_link(s, 0);
      if (s->dim == 3 && s->world > 1) return s->any_excl ? coop_launch(s, k_gs_persistent<3, true, true>, s->grid_solver, s->geo, a, 0)
                                                          : coop_launch(s, k_gs_persistent<3, false, true>, s->grid_solver, s->geo, a, 0);
      if (s->dim == 3) return s->any_excl ? coop_launch(s, k_gs_persistent<3, true>, s->grid_solver, s->geo, a, 0)
                                          : coop_launch(s, k_gs_persistent<3, false>, s->grid_solver, s->geo, a, 0);
      return s->any_excl ? coop_launch(s, k_gs_persistent<2, true>, s->grid_solver, s->geo, a, 0)
                         : coop_launch(s, k_gs_persistent<2, false>, s->grid_solver, s->geo, a, 0);
    };
    double* defer_out = (s->defer && s->world == 1 && s->nsolves < 4096) ? s->sorres + 2 * s->nsolves : nullptr;
    if (int rc = run_sor(s, s->PP, s->nsh, c.lu_relaxed_tolerance, c.lu_relaxed_num_iters_limit, launch, &it, &df, defer_out)) return rc;
    if (!s->defer && s->gs_tiled) if (int rc = gt_check(s)) return rc;   // inside hg_step: part of the step's status block
    if (second) launch_padd(s); else launch_pcorr(s);
  } else if (c.linear_solver_pressure == HG_LS_JACOBI) {
    // natural layout: constants back from the sheared array, rows regenerated from d_c
    DIMSEL(s, k_from_sheared, nblk(s->nc), 256, s->geo, s->RP, s->w1);
    if (int rc = run_jacobi(s, nullptr, s->dc, s->w1, c.lu_relaxed_tolerance, c.lu_relaxed_num_iters_limit,
                            c.lu_relaxed_relaxation_factor, &it, &df)) return rc;
    // p_curr = p_prev + alpha p'
    DIMSEL(s, k_to_sheared, nblk(s->nc), 256, s->geo, s->pc, s->PP);
    if (second) launch_padd(s); else launch_pcorr(s);
  } else if (c.linear_solver_pressure == HG_LS_LU_RELAXED) {
    P7 rows;
    if (s->dim == 3) {
      for (int t = 0; t < 7; ++t) rows.p[t] = s->An[t];
      DIMSEL(s, k_prows, nblk(s->nc), 256, s->geo, s->dc, 0, rows);
      shear_arrays(s, s->An, s->A, 7);
    } else {
      for (int t = 0; t < 7; ++t) rows.p[t] = s->A[t];
      DIMSEL(s, k_prows, nblk(s->nc), 256, s->geo, s->dc, 1, rows);
    }
    if (int rc = run_lu_relaxed(s, s->RP, s->PP, s->X[0], s->X[1], c.lu_relaxed_tolerance, c.lu_relaxed_num_iters_limit,
                                c.lu_relaxed_relaxation_factor, &it, &df)) return rc;
    if (second) launch_padd(s); else launch_pcorr(s);
  } else {   // lu: one forward + backward sweep over the explicit rows (linear.hpp:533-566); the factory reports no sweeps
    P7 rows;
    if (s->dim == 3) {
      for (int t = 0; t < 7; ++t) rows.p[t] = s->An[t];
      DIMSEL(s, k_prows, nblk(s->nc), 256, s->geo, s->dc, 0, rows);
      shear_arrays(s, s->An, s->A, 7);
    } else {
      for (int t = 0; t < 7; ++t) rows.p[t] = s->A[t];
      DIMSEL(s, k_prows, nblk(s->nc), 256, s->geo, s->dc, 1, rows);
    }
    if (int rc = solve_system(s, HG_LS_LU, s->RP, s->PP, nullptr, nullptr, &it, &df)) return rc;
    if (second) launch_padd(s); else launch_pcorr(s);
  }
  if (!s->defer && s->lu_tiled) if (int rc = lt_check(s)) return rc;   // the momentum solve's dataflow kernel
  if (it < 0) ++s->nsolves;   // deferred: counted on the device, added at the end of the step
  else { s->sweeps_total += it + 1; s->last_diff = df; }
  return 0;
}

// natural -> sheared copies of n arrays (3-D: tiled transpose; the 2-D kernels write sheared directly)
static int shear_arrays(hg_state* s, double* const* in, double* const* out, int n) {
  ShearArgs a; a.narr = n;
  for (int q = 0; q < n; ++q) { a.in[q] = in[q]; a.out[q] = out[q]; }
  dim3 grid((s->n[0] + 31) / 32, s->n[1], (s->n[2] + 31) / 32);
  k_shear3<<<grid, 256, 0, s->st>>>(s->geo, a);
  ++s->launches;
  return 0;
}

static int solve_lu_tiled(hg_state* s, int ncomp, double* const* Rs = nullptr, double* const* Xs = nullptr) {
  LtArgs a{};
  for (int t = 0; t < 7; ++t) a.A[t] = s->A[t];
  for (int n = 0; n < 3; ++n) { a.R[n] = Rs && n < ncomp ? Rs[n] : s->R[n]; a.X[n] = Xs && n < ncomp ? Xs[n] : s->X[n]; }
  a.ncomp = ncomp; a.boxes = s->lt_boxes; a.nboxes = s->lt_nboxes; a.nbi = s->lt_nbi; a.progress = s->lt_progress; a.ctl = s->lt_ctl;
  if (s->world > 1) ++s->slab.lu_seq;
  a.link = slab_link(s, 1); a.link_stride = s->nxy;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (s->profile_on) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, s->st); }
  int grid = std::min(s->lt_nboxes, s->num_sms * LT_CTAS_PER_SM);   // all boxes resident when they fit
  if (s->cfg.solver_ctas > 0) grid = std::min(grid, s->cfg.solver_ctas);           // ranks sharing a device (tests)
#ifdef LT_TRACE
  static unsigned long long* trace_dev = nullptr; static int trace_calls = 0;
  if (!trace_dev) cudaMalloc((void**)&trace_dev, (size_t)s->lt_nboxes * 4 * sizeof(unsigned long long));
  a.trace = trace_dev;
  static unsigned long long* trace2_dev = nullptr;
  if (!trace2_dev) { cudaMalloc((void**)&trace2_dev, (size_t)s->lt_nboxes * 512 * 8); cudaMemset(trace2_dev, 0, (size_t)s->lt_nboxes * 512 * 8); }
  a.trace2 = trace2_dev;
#endif
  for (int dir = 0; dir < 2; ++dir) {
    CK(cudaMemsetAsync(s->lt_progress, 0, s->lt_nboxes * sizeof(int), s->st));
    CK(cudaMemsetAsync(s->lt_ctl, 0, sizeof(int), s->st));   // next-box counter; the abort flag [1] is sticky
    if (s->world > 1) {
      if (dir == 0) k_lu_tiled<0, true><<<grid, LT_THREADS, LT_RING_BYTES, s->st>>>(s->geo, a);
      else k_lu_tiled<1, true><<<grid, LT_THREADS, LT_RING_BYTES, s->st>>>(s->geo, a);
    } else {
      if (dir == 0) k_lu_tiled<0><<<grid, LT_THREADS, LT_RING_BYTES, s->st>>>(s->geo, a);
      else k_lu_tiled<1><<<grid, LT_THREADS, LT_RING_BYTES, s->st>>>(s->geo, a);
    }
    CK(cudaGetLastError());
    ++s->launches;
#ifdef LT_TRACE
    if (getenv("HYDRO_LT_TRACE") && ++trace_calls == 25 + dir) {   // one forward and one backward launch of a warmed-up step
      CK(cudaStreamSynchronize(s->st));
      std::vector<unsigned long long> h((size_t)s->lt_nboxes * 4);
      cudaMemcpy(h.data(), trace_dev, h.size() * 8, cudaMemcpyDeviceToHost);
      FILE* f = fopen((std::string(getenv("HYDRO_LT_TRACE")) + (dir ? ".bwd" : ".fwd")).c_str(), "w");
      if (f) { fprintf(f, "%d %d\n", s->lt_nboxes, s->lt_nbi);
               for (int q = 0; q < s->lt_nboxes; ++q) fprintf(f, "%d %llu %llu %llu %llu\n", q, h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
               fclose(f); }
      if (dir == 0) {
        std::vector<unsigned long long> h2((size_t)s->lt_nboxes * 512);
        cudaMemcpy(h2.data(), trace2_dev, h2.size() * 8, cudaMemcpyDeviceToHost);
        FILE* f2 = fopen((std::string(getenv("HYDRO_LT_TRACE")) + ".macro").c_str(), "wb");
        if (f2) { fwrite(h2.data(), 8, h2.size(), f2); fclose(f2); }
      }
    }
#endif
  }
  if (e0) { cudaEventRecord(e1, s->st); s->prof_ev[1].push_back({e0, e1}); }
  return 0;
}
static int lt_check(hg_state* s) {   // did a dependency wait time out?
  int h = 0;
  CK(cudaMemcpyAsync(&h, s->lt_ctl + 1, sizeof(int), cudaMemcpyDeviceToHost, s->st));
  CK(cudaStreamSynchronize(s->st));
  if (h) { s->err = "k_lu_tiled: dependency wait timed out"; return HG_ERR_CUDA; }
  return 0;
}

static int solve_lu(hg_state* s, int ncomp, double* const* Rs = nullptr, double* const* Xs = nullptr) {
  if (s->lu_tiled) return solve_lu_tiled(s, ncomp, Rs, Xs);
  LuArgs a{};
  for (int t = 0; t < 7; ++t) a.A[t] = s->A[t];
  for (int n = 0; n < 3; ++n) { a.R[n] = Rs && n < ncomp ? Rs[n] : s->R[n]; a.X[n] = Xs && n < ncomp ? Xs[n] : s->X[n]; }
  a.ncomp = ncomp; a.tt = s->tt;
  if (s->world > 1) ++s->slab.lu_seq;
  a.link = slab_link(s, 1); a.link_stride = s->nxy;
  if (s->dim == 3 && s->world > 1) return coop_launch(s, k_lu_persistent<3, true>, s->grid_lu, s->geo, a, 1);
  if (s->dim == 3) return coop_launch(s, k_lu_persistent<3>, s->grid_lu, s->geo, a, 1);
  return coop_launch(s, k_lu_persistent<2>, s->grid_lu, s->geo, a, 1);
}

// One system of the reference's linear-solver factory (hydro2d.hpp:194-218, linear.hpp:533-782): rows in s->A (hyperplane-
// major), constants R, result X (both hyperplane-major).  lu runs the box dataflow; the iterative solvers are the matrix
// kernels of hg_solvers.cuh with the shared lu_relaxed_* parameters.  Scratch: t1, t2 (hyperplane-major, lu_relaxed), the
// natural-layout staging arrays An, w1, w2, pc (Jacobi).  One GPU (the slab kernels exist for lu and the pressure sweeps only).
static int solve_system(hg_state* s, int solver, double* R, double* X, double* t1, double* t2, int* it, double* df) {
  const hg_config& c = s->cfg;
  const unsigned gb = nblk(s->nc);
  *it = 0; *df = 0.;
  if (solver == HG_LS_LU) { double* Rs[1] = {R}; double* Xs[1] = {X}; return solve_lu(s, 1, Rs, Xs); }
  if (s->world > 1) { s->err = "multi-GPU slabs: only lu (and gauss_seidel for the pressure) are decomposed"; return HG_ERR_INVALID; }
  if (solver == HG_LS_GAUSS_SEIDEL) {
    auto launch = [&](int sb, int se) -> int {
      SorArgs a; for (int t = 0; t < 7; ++t) a.A[t] = s->A[t];
      a.R = R; a.X = X; a.diff = s->diffs; a.s_begin = sb; a.s_end = se; a.omega = c.lu_relaxed_relaxation_factor; a.tt = s->tt;
      if (s->dim == 3) return coop_launch(s, k_sor_matrix_persistent<3>, s->grid_solver, s->geo, a);
      return coop_launch(s, k_sor_matrix_persistent<2>, s->grid_solver, s->geo, a);
    };
    return run_sor(s, X, s->nsh, c.lu_relaxed_tolerance, c.lu_relaxed_num_iters_limit, launch, it, df);
  }
  if (solver == HG_LS_LU_RELAXED)
    return run_lu_relaxed(s, R, X, t1, t2, c.lu_relaxed_tolerance, c.lu_relaxed_num_iters_limit, c.lu_relaxed_relaxation_factor, it, df);
  if (solver == HG_LS_JACOBI) {
    for (int t = 0; t < 7; ++t) {
      if (s->dim == 2 && (t == CZM || t == CZP)) continue;
      DIMSEL(s, k_from_sheared, gb, 256, s->geo, s->A[t], s->An[t]);
    }
    DIMSEL(s, k_from_sheared, gb, 256, s->geo, R, s->w1);
    if (int rc = run_jacobi(s, s->An, nullptr, s->w1, c.lu_relaxed_tolerance, c.lu_relaxed_num_iters_limit, c.lu_relaxed_relaxation_factor, it, df)) return rc;
    DIMSEL(s, k_to_sheared, gb, 256, s->geo, s->pc, X);
    return 0;
  }
  s->err = "Unknown linear solver";
  return HG_ERR_INVALID;
}

// ------------------------------------------------------------------ properties / statistics
static int smooth_field(hg_state* s, const double* in, int repeat, double* out) {
  // GetSmoothField (solver.hpp:636-656); ping-pong between out and w2
  if (repeat <= 0) { if (in != out) CK(cudaMemcpyAsync(out, in, s->nc * sizeof(double), cudaMemcpyDeviceToDevice, s->st)); return 0; }
  const double* src = in;
  double* bufs[2] = {(repeat % 2) ? out : s->w2, (repeat % 2) ? s->w2 : out};
  for (int r = 0; r < repeat; ++r) {
    double* dst = bufs[r % 2];
    XCH(s, 1, const_cast<double*>(src));
    DIMSEL(s, k_smooth, nblk(s->nc), 256, s->geo, src, dst);
    src = dst;
  }
  return 0;
}

extern "C" int hg_update_properties(hg_handle s) {
  if (!s) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  tpush(s, "step.fluid_properties");
  const hg_config& c = s->cfg;
  PropArgs a; a.np = c.num_phases;
  for (int p = 0; p < 3; ++p) {
    a.density[p] = c.density[p]; a.viscosity[p] = c.viscosity[p]; a.conductivity[p] = c.conductivity[p];
    a.pd[p] = s->pd[p][L_TC]; a.vf[p] = s->vf[p];
  }
  a.rho_raw = s->rho_raw; a.mu_raw = s->mu_raw; a.kc = s->kc;
  LAUNCH(s, k_volfrac, nblk(s->nc), 256, a, s->nc);
  if (int rc = smooth_field(s, s->rho_raw, c.density_smooth_times, s->rho)) return rc;
  if (int rc = smooth_field(s, s->mu_raw, c.viscosity_smooth_times, s->mu)) return rc;
  LAUNCH(s, k_force, nblk(s->nc), 256, s->dim, s->rho_raw, c.gravity[0], c.gravity[1], c.gravity[2], c.force[0], c.force[1],
         c.force[2], p3(s->force), s->nc);
  if (c.num_phases >= 2 && c.sigma != 0.) {
    XCH(s, 1, s->vf[1]);   // slabs: the volume fraction and then its gradient across the interfaces
    P3 gs; for (int d = 0; d < 3; ++d) gs.p[d] = s->G[d];
    DIMSEL(s, k_grad_pd, nblk(s->nc), 256, s->geo, s->vf[1], s->pd_init[0], gs);
    XCH(s, 1, s->G[0], s->G[1], s->dim > 2 ? s->G[2] : nullptr);
    CP3 gsc; for (int d = 0; d < 3; ++d) gsc.p[d] = s->G[d];
    DIMSEL(s, k_stforce, nblk(s->nc), 256, s->geo, gsc, c.sigma, p3(s->stforce));
  }
  if (c.force_smooth_times > 0) {
    for (int d = 0; d < s->dim; ++d) {
      if (int rc = smooth_field(s, s->force[d], c.force_smooth_times, s->w1)) return rc;
      CK(cudaMemcpyAsync(s->force[d], s->w1, s->nc * sizeof(double), cudaMemcpyDeviceToDevice, s->st));
    }
  }
  if (s->any_slip) {   // CalcPhaseVelocitySlip (hydro2d.hpp:1422)
    SlipArgs q; q.np = c.num_phases;
    for (int ph = 0; ph < 3; ++ph) {
      q.enable[ph] = ph < c.num_phases ? c.enable_settling[ph] : 0; q.radius[ph] = c.bubble_radius[ph]; q.density[ph] = c.density[ph];
      q.vf[ph] = s->vf[ph] ? s->vf[ph] : s->zero; q.fslip[ph] = s->fslip[ph];
      for (int d = 0; d < 3; ++d) q.slipv[ph][d] = s->slipv[ph][d];
    }
    for (int d = 0; d < 3; ++d) q.gravity[d] = c.gravity[d];
    q.rho_raw = s->rho_raw; q.mu = s->mu;
    DIMSEL(s, k_slip_cell, nblk(s->nc), 256, s->geo, q);
    DIMSEL(s, k_slip_face, nblk(s->nc), 256, s->geo, q);
  }
  // slabs: face values of density, viscosity and force are taken across the slab interfaces
  XCH(s, 1, s->rho, s->mu, s->force[0], s->force[1], s->dim > 2 ? s->force[2] : nullptr, c.heat_enable ? s->kc : nullptr);
  tpop(s);
  return 0;
}

// CalcStat (hydro2d.hpp:1432-1529).  with_status (end of hg_step): the step's status block (NaN flags, abort words of the
// dataflow kernels, last convergence indicator, deferred sweep counts) travels behind the 36 statistics values in the
// same transfer / the same rank-ordered reduction, so a step waits for the device once.
enum { ST_STAT = 2, ST_STATUS = 38, ST_TOTAL = 51 };   // offsets in s->scal / s->hscal
static int calc_stat_finish(hg_state* s, hg_step_stats* st);
// enqueue: statistics kernel + status block + transfer into the pinned mirror (slabs: the rank-ordered reduction, which waits);
// calc_stat_finish waits for the stream and decodes
static int calc_stat_enqueue(hg_state* s, bool with_status) {
  const hg_config& c = s->cfg;
  double init[36];
  for (int p = 0; p < 3; ++p) { for (int q = 0; q < 12; ++q) init[p * 12 + q] = 0.; init[p * 12 + 7] = 1e10; init[p * 12 + 8] = -1e10; }
  memcpy(s->hinit, init, sizeof init);   // pinned: the async copy reads it when the stream gets there
  CK(cudaMemcpyAsync(s->scal + ST_STAT, s->hinit, sizeof(init), cudaMemcpyHostToDevice, s->st));
  StatArgs a; a.np = c.num_phases;
  for (int p = 0; p < 3; ++p) { a.vf[p] = s->vf[p]; a.pd[p] = s->pd[p][L_TC]; }
  for (int d = 0; d < 3; ++d) a.u[d] = s->u[L_TC][d] ? s->u[L_TC][d] : s->zero;
  a.out = s->scal + ST_STAT;
  DIMSEL(s, k_stat, nblk(s->nc, 256 * STAT_CPT), 256, s->geo, a);
  int nvals = 36;
  if (with_status) {
    const double* rl = s->iter_count > 0 ? s->resid + (s->iter_count - 1 < 4095 ? s->iter_count - 1 : 4095) : nullptr;
    LAUNCH(s, k_status_pack, 1, 32, s->flag + 4, s->gs_tiled ? s->gt_ctl : nullptr, s->lu_tiled ? s->lt_ctl : nullptr, rl, s->sorres,
           s->nsolves, s->scal + ST_STATUS);
    nvals = ST_TOTAL - ST_STAT;
  }
  if (s->world > 1) {   // sums in rank order; minima / maxima
    std::vector<double> all;
    if (int rc = slab_gather(s, s->scal + ST_STAT, nvals, all)) return rc;
    for (int q = 0; q < nvals; ++q) {
      double v = all[q];
      for (int r = 1; r < s->world; ++r) {
        const double w = all[(size_t)r * nvals + q];
        if (q >= 36) v = v < w ? w : v;   // status words: maximum
        else v = (q % 12 == 7) ? (w < v ? w : v) : (q % 12 == 8) ? (v < w ? w : v) : v + w;
      }
      s->hscal[ST_STAT + q] = v;
    }
  } else {
    CK(cudaMemcpyAsync(s->hscal + ST_STAT, s->scal + ST_STAT, nvals * sizeof(double), cudaMemcpyDeviceToHost, s->st));
  }
  return 0;
}
static int calc_stat_finish(hg_state* s, hg_step_stats* st) {
  const hg_config& c = s->cfg;
  if (s->world == 1) CK(cudaStreamSynchronize(s->st));
  hg_step_stats& o = s->stat;
  for (int p = 0; p < c.num_phases; ++p) {
    const double* r = s->hscal + ST_STAT + p * 12;
    o.volume[p] = r[0]; o.mass[p] = r[0] * c.density[p]; o.pd_min[p] = r[7]; o.pd_max[p] = r[8];
    for (int d = 0; d < 3; ++d) {
      o.center[p][d] = d < s->dim ? r[1 + d] / r[0] + s->meshpos[d] : 0.;
      o.velocity[p][d] = d < s->dim ? r[4 + d] / r[0] : 0.;
    }
  }
  // stat_vcx_<i> = (cx - previous cx) / dt, 0 at the first call (hydro2d.hpp:1459-1463); the ADHOC mesh velocity follows
  // phase 1 (hydro2d.hpp:1510-1524): meshvel = (v, 0, 0) w + meshvel (1 - w), taken by the next step's fluxes
  double vcx[HG_MAX_PHASES] = {0., 0., 0.};
  for (int p = 0; p < c.num_phases; ++p) {
    const double prev = s->stat_cx_set[p] ? s->stat_cx[p] : o.center[p][0];
    vcx[p] = (o.center[p][0] - prev) / s->dt;
    s->stat_cx[p] = o.center[p][0]; s->stat_cx_set[p] = true;
  }
  if (c.meshvel_auto) {
    const double v0 = c.meshvel_auto == 1 ? o.velocity[1][0] : vcx[1], w = c.meshvel_weight;
    for (int d = 0; d < 3; ++d) s->cfg.meshvel[d] = (d == 0 ? v0 : 0.) * w + s->cfg.meshvel[d] * (1. - w);
  }
  if (c.meshvel_output) for (int d = 0; d < s->dim; ++d) s->meshpos[d] += c.meshvel[d] * s->dt;   // hydro2d.hpp:1526-1528
  if (st) {
    for (int p = 0; p < HG_MAX_PHASES; ++p) {
      st->volume[p] = o.volume[p]; st->mass[p] = o.mass[p]; st->pd_min[p] = o.pd_min[p]; st->pd_max[p] = o.pd_max[p];
      for (int d = 0; d < 3; ++d) { st->center[p][d] = o.center[p][d]; st->velocity[p][d] = o.velocity[p][d]; }
    }
  }
  return 0;
}
static int calc_stat(hg_state* s, hg_step_stats* st, bool with_status) {
  if (int rc = calc_stat_enqueue(s, with_status)) return rc;
  return calc_stat_finish(s, st);
}
extern "C" int hg_calc_stat(hg_handle s, hg_step_stats* st) {
  if (!s) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  return calc_stat(s, st, false);
}
extern "C" int hg_get_stats(hg_handle s, hg_step_stats* st) {
  if (!s || !st) return HG_ERR_INVALID;
  *st = s->stat;
  return 0;
}

// ------------------------------------------------------------------ fluid solver protocol
enum { NF_INIT_P = 0, NF_INIT_U = 1, NF_FIN_P = 2, NF_FIN_U = 3, NF_HEAT_INIT = 4, NF_HEAT_FIN = 5, NF_COUNT = 8 };
static const char* const nan_msgs[NF_COUNT] = {"NaN initial pressure", "NaN initial field", "NaN pressure", "NaN field",
                                               "NaN initial field", "NaN field", "", ""};
// IsNan scan of a field (solver.hpp:17-30).  Inside hg_step the flag is only raised (slot `which` of the step's status
// block, read once at the end of the step); called through the fine-grained entries it is checked at once.
static int check_nan(hg_state* s, const double* a, long long n, int which) {
  if (s->defer) { LAUNCH(s, k_nan_flag, nblk(n), 256, a, n, s->flag + 4 + which); return 0; }
  CK(cudaMemsetAsync(s->flag, 0, sizeof(int), s->st));
  LAUNCH(s, k_nan_flag, nblk(n), 256, a, n, s->flag);
  int h = 0;
  if (s->world > 1) {   // every rank must take the same branch
    LAUNCH(s, k_flag_to_double, 1, 1, s->flag, s->scal + 61);
    double any = 0.;
    if (int rc = slab_reduce(s, s->scal + 61, 1, 0, &any)) return rc;
    h = any != 0.;
  } else {
    CK(cudaMemcpyAsync(&h, s->flag, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CK(cudaStreamSynchronize(s->st));
  }
  if (h) { s->err = nan_msgs[which]; return HG_ERR_NAN; }
  return 0;
}

extern "C" int hg_fluid_start_step(hg_handle s) {   // fluid.hpp:793-812
  if (!s) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  s->iter_count = 0;
  CK(cudaMemsetAsync(s->resid, 0, 4096 * sizeof(double), s->st));
  // slabs: the caller may have set density / viscosity / force since the last hg_update_properties
  XCH(s, 1, s->rho, s->mu, s->force[0], s->force[1], s->dim > 2 ? s->force[2] : nullptr);
  const double ge = s->cfg.guess_extrapolation;
  if (s->defer) {
    // inside hg_step the IsNan scans of time_curr ride on the pass that reads it anyway (flags read at the end of the step)
    for (int d = 0; d < s->dim; ++d)
      LAUNCH(s, k_start_layer, nblk(s->nc), 256, s->u[L_IC][d], s->u[L_TC][d], s->u[L_TP][d], ge, s->nc, s->flag + 4 + NF_INIT_U, s->nc);
    LAUNCH(s, k_start_layer, nblk(s->nc), 256, s->p[L_IC], s->p[L_TC], s->p[L_TP], ge, s->nc, s->flag + 4 + NF_INIT_P, s->nc);
  } else {
    if (int rc = check_nan(s, s->p[L_TC], s->nc, NF_INIT_P)) return rc;
    for (int d = 0; d < s->dim; ++d) {
      if (int rc = check_nan(s, s->u[L_TC][d], s->nc, NF_INIT_U)) return rc;
      LAUNCH(s, k_start_layer, nblk(s->nc), 256, s->u[L_IC][d], s->u[L_TC][d], s->u[L_TP][d], ge, s->nc);
    }
    LAUNCH(s, k_start_layer, nblk(s->nc), 256, s->p[L_IC], s->p[L_TC], s->p[L_TP], ge, s->nc);
  }
  LAUNCH(s, k_start_layer, nblk(s->nf), 256, s->F[L_IC], s->F[L_TC], s->F[L_TP], ge, s->nf);
  return 0;
}

// the dim component systems share their matrix (conv_diff.hpp:149-259): lu solves them in one pass, the iterative solvers
// of the factory one after the other (pressure arrays PP / RP are free at this point: scratch of lu_relaxed)
static int solve_momentum(hg_state* s) {
  if (s->cfg.linear_solver_velocity == HG_LS_LU) return solve_lu(s, s->dim);
  for (int n = 0; n < s->dim; ++n) {
    int it; double df;
    if (int rc = solve_system(s, s->cfg.linear_solver_velocity, s->R[n], s->X[n], s->PP, s->RP, &it, &df)) return rc;
  }
  return 0;
}

extern "C" int hg_fluid_make_iteration(hg_handle s) {   // fluid.hpp:814-1158
  if (!s) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  const hg_config& c = s->cfg;
  const unsigned gb = nblk(s->nc);
  // iter_prev <- iter_curr by pointer rotation: every kernel below writes all of the new iter_curr
  std::swap(s->p[L_IP], s->p[L_IC]);
  std::swap(s->F[L_IP], s->F[L_IC]);
  for (int d = 0; d < s->dim; ++d) std::swap(s->u[L_IP][d], s->u[L_IC][d]);
  // from here on: *_IP hold the state the iteration starts from

  XCH(s, 1, s->u[L_IP][0], s->u[L_IP][1], s->u[L_IP][2], s->p[L_IP]);
  if (s->any_outlet) {   // UpdateOutletBaseConditions + UpdateDerivedConditions (fluid.hpp:820-821)
    OutletArgs a; for (int d = 0; d < 3; ++d) a.u[d] = s->u[L_IP][d] ? s->u[L_IP][d] : s->zero;
    a.outvel = s->outvel; a.outplane = s->geo.outplane; a.part = s->outpart; a.corr = s->outpart + 3 * s->out_terms;
    DIMSEL(s, k_outlet_collect, s->out_blocks, 256, s->geo, a);
    DIMSEL(s, k_outlet_correction, 1, 256, s->geo, s->outpart, a.corr);
    DIMSEL(s, k_outlet_apply, s->out_blocks, 256, s->geo, a);
  }
  const bool fast = s->fast;
  const Geo gl = list_geo(s);
  const unsigned gbl = nblk(s->nslow);
  const dim3 fg = fast_grid(s);
  if (fast) {
    // interior cells: gradients + restored force in one pass, then source + assembly + transpose in one pass (hg_fast.cuh);
    // the listed cells (near walls / obstacle / fixed-pressure cell) keep the generic kernels
    tpush(s, "fluid.0-1.gradients");
    { FaArgs a;
      for (int d = 0; d < 3; ++d) { a.u[d] = s->u[L_IP][d]; a.force[d] = s->force[d]; a.fcr[d] = s->fcr[d]; a.gp[d] = s->gp[d]; }
      a.p = s->p[L_IP];
      for (int q = 0; q < 9; ++q) a.G[q] = s->G[q];
      k_fa_grad<<<fg, FT_THREADS, 0, s->st>>>(s->geo, s->slow, a); ++s->launches; }
    P9 G; for (int q = 0; q < 9; ++q) G.p[q] = s->G[q];
    if (s->nslow) {
      k_pre<3><<<gbl, 256, 0, s->st>>>(gl, cp3(s->force), s->p[L_IP], p3(s->fcr), p3(s->gp)); ++s->launches;
      k_velgrad<3><<<gbl, 256, 0, s->st>>>(gl, cp3(s->u[L_IP]), G); ++s->launches;
    }
    tpop(s);
    if (s->world > 1) if (int rc = slab_exchange(s, s->G, 9, 1)) return rc;
    tpush(s, "fluid.2.convection-diffusion");
    const int use_stf = (c.num_phases >= 2 && c.sigma != 0.) ? 1 : 0;
    if (s->nslow) {
      P9c Gc; for (int q = 0; q < 9; ++q) Gc.p[q] = s->G[q];
      k_source<3><<<gbl, 256, 0, s->st>>>(gl, Gc, s->mu, cp3(s->gp), cp3(s->fcr), cp3(s->stforce), use_stf, p3(s->fs)); ++s->launches;
      AsmArgs a;
      for (int n = 0; n < 3; ++n) { a.prev[n] = s->u[L_IP][n]; a.tc[n] = s->u[L_TC][n]; a.tp[n] = s->u[L_TP][n]; a.src[n] = s->fs[n]; a.R[n] = s->R[n]; }
      for (int q = 0; q < 9; ++q) a.grad[q] = s->G[q];
      a.rho = s->rho; a.mu = s->mu; a.F = s->F[L_IP];
      bdf_coeffs(s->dt, c.time_second_order, a.co);
      a.relax = c.velocity_relaxation_factor;
      a.coeffsum = s->dc; a.coeffsum_div = 3.; a.out_sheared = 1;
      for (int t = 0; t < 7; ++t) a.A[t] = s->A[t];
      k_assemble<3, K_VEL, 3><<<gbl, 256, 0, s->st>>>(gl, a); ++s->launches;
    }
    { FbArgs a;
      for (int n = 0; n < 3; ++n) { a.prev[n] = s->u[L_IP][n]; a.tc[n] = s->u[L_TC][n]; a.tp[n] = s->u[L_TP][n];
                                    a.gp[n] = s->gp[n]; a.fcr[n] = s->fcr[n]; a.stf[n] = s->stforce[n]; a.out[7 + n] = s->R[n]; }
      for (int q = 0; q < 9; ++q) a.G[q] = s->G[q];
      for (int t = 0; t < 7; ++t) a.out[t] = s->A[t];
      a.rho = s->rho; a.mu = s->mu; a.F = s->F[L_IP]; a.use_stf = use_stf;
      bdf_coeffs(s->dt, c.time_second_order, a.co);
      a.relax = c.velocity_relaxation_factor; a.coeffsum = s->dc;
      k_fb_momentum<<<fg, FT_THREADS, 0, s->st>>>(s->geo, s->slow, a); ++s->launches; }
    if (int rc = solve_momentum(s)) return rc;
    k_ft_apply_corr<3><<<fg, FT_THREADS, 0, s->st>>>(s->geo, cp3(s->u[L_IP]), cp3(s->X), p3(s->u[L_IC])); ++s->launches;
    tpop(s);
  } else {
  tpush(s, "fluid.0.pressure-gradient");
  DIMSEL(s, k_pre, gb, 256, s->geo, cp3(s->force), s->p[L_IP], p3(s->fcr), p3(s->gp));
  tpop(s);
  tpush(s, "fluid.1a.explicit-viscosity");
  { P9 G; for (int q = 0; q < 9; ++q) G.p[q] = s->G[q];
    DIMSEL(s, k_velgrad, gb, 256, s->geo, cp3(s->u[L_IP]), G);
    if (s->world > 1) if (int rc = slab_exchange(s, s->G, s->dim * s->dim, 1)) return rc;
    P9c Gc; for (int q = 0; q < 9; ++q) Gc.p[q] = s->G[q];
    const int use_stf = (c.num_phases >= 2 && c.sigma != 0.) ? 1 : 0;
    DIMSEL(s, k_source, gb, 256, s->geo, Gc, s->mu, cp3(s->gp), cp3(s->fcr), cp3(s->stforce), use_stf, p3(s->fs)); }
  tpop(s);
  tpush(s, "fluid.2.convection-diffusion");
  { AsmArgs a;
    for (int n = 0; n < 3; ++n) { a.prev[n] = s->u[L_IP][n]; a.tc[n] = s->u[L_TC][n]; a.tp[n] = s->u[L_TP][n]; a.src[n] = s->fs[n]; a.R[n] = s->R[n]; }
    for (int q = 0; q < 9; ++q) a.grad[q] = s->G[q];
    a.rho = s->rho; a.mu = s->mu; a.F = s->F[L_IP];
    bdf_coeffs(s->dt, c.time_second_order, a.co);
    a.relax = c.velocity_relaxation_factor;
    a.coeffsum = s->dc; a.coeffsum_div = (double)s->dim;
    if (s->dim == 3) {
      a.out_sheared = 0;
      for (int t = 0; t < 7; ++t) a.A[t] = s->An[t];
      for (int n = 0; n < 3; ++n) a.R[n] = s->An[7 + n];
      k_assemble<3, K_VEL, 3><<<gb, 256, 0, s->st>>>(s->geo, a);
      ++s->launches;
      double* outs[10]; for (int t = 0; t < 7; ++t) outs[t] = s->A[t]; for (int n = 0; n < 3; ++n) outs[7 + n] = s->R[n];
      shear_arrays(s, s->An, outs, 10);
    } else {
      a.out_sheared = 1;
      for (int t = 0; t < 7; ++t) a.A[t] = s->A[t];
      k_assemble<2, K_VEL, 2><<<gb, 256, 0, s->st>>>(s->geo, a);
      ++s->launches;
    }
    if (int rc = solve_momentum(s)) return rc;
    if (s->dim == 3) { k_apply_corr<3, 3><<<gb, 256, 0, s->st>>>(s->geo, cp3(s->u[L_IP]), cp3(s->X), p3(s->u[L_IC])); }
    else { k_apply_corr<2, 2><<<gb, 256, 0, s->st>>>(s->geo, cp3(s->u[L_IP]), cp3(s->X), p3(s->u[L_IC])); }
    ++s->launches; }
  tpop(s);
  }
  XCH(s, 1, s->u[L_IC][0], s->u[L_IC][1], s->u[L_IC][2], s->gp[0], s->gp[1], s->gp[2], s->fcr[0], s->fcr[1], s->fcr[2], s->dc);
  if (fast && s->gs_tiled) {
    // Rhie-Chow fluxes + rows of the pressure-correction system, packed for k_gs_tiled, in one pass over the interior cells
    tpush(s, "fluid.3-5.fluxes+pressure-system");
    { FcArgs a;
      for (int d = 0; d < 3; ++d) { a.us[d] = s->u[L_IC][d]; a.gp[d] = s->gp[d]; a.fcr[d] = s->fcr[d]; a.force[d] = s->force[d]; a.meshvel[d] = c.meshvel[d]; }
      a.pprev = s->p[L_IP]; a.dc = s->dc; a.rc = c.rhie_chow_factor; a.Fs = s->Fs; a.co5 = s->co5;
      for (int q = 0; q < 5; ++q) a.co[q] = s->CO + q * s->co5.arr;
      k_fc_flux_rows<<<fg, FT_THREADS, 0, s->st>>>(s->geo, s->slow, a); ++s->launches; }
    if (s->nslow) {
      FstarArgs a;
      for (int d = 0; d < 3; ++d) { a.us[d] = s->u[L_IC][d]; a.gp[d] = s->gp[d]; a.fcr[d] = s->fcr[d]; a.force[d] = s->force[d]; a.meshvel[d] = c.meshvel[d]; }
      a.pprev = s->p[L_IP]; a.dc = s->dc; a.rc = c.rhie_chow_factor; a.Fs = s->Fs;
      k_fstar<3><<<gbl, 256, 0, s->st>>>(gl, a); ++s->launches;
      k_prhs_co5<<<gbl, 256, 0, s->st>>>(gl, s->Fs, s->dc, s->CO, s->co5); ++s->launches;
    }
    if (s->geo.zlo > 0) { k_gt_cz_halo<<<nblk(s->nxy), 256, 0, s->st>>>(s->geo, s->dc, s->CO, s->co5); ++s->launches; }
    if (int rc = gt_ghost_rows(s)) return rc;
    tpop(s);
  } else {
  tpush(s, "fluid.3.momentum-interpolation");
  { FstarArgs a;
    for (int d = 0; d < 3; ++d) { a.us[d] = s->u[L_IC][d]; a.gp[d] = s->gp[d]; a.fcr[d] = s->fcr[d]; a.force[d] = s->force[d]; a.meshvel[d] = c.meshvel[d]; }
    a.pprev = s->p[L_IP]; a.dc = s->dc; a.rc = c.rhie_chow_factor; a.Fs = s->Fs;
    DIMSEL(s, k_fstar, gb, 256, s->geo, a); }
  tpop(s);
  tpush(s, "fluid.5.pressure-system");
  if (s->dim == 3) {
    DIMSEL(s, k_prhs, gb, 256, s->geo, s->Fs, s->dc, 0, s->An[0], s->An[1], s->An[2], s->An[3], s->gs_tiled ? s->An[4] : nullptr);
    if (s->gs_tiled) {
      // natural -> packed rows of k_gs_tiled in one pass (transpose + packing)
      GtPackArgs pa; pa.in[0] = s->An[0]; pa.in[1] = s->An[4]; pa.in[2] = s->An[1]; pa.in[3] = s->An[2]; pa.in[4] = s->An[3];
      pa.CO = s->CO; pa.dc = s->dc; pa.co = s->co5;
      k_gt_shear_pack<<<dim3((s->n[0] + 31) / 32, s->n[1], (s->n[2] + 31) / 32), 256, 0, s->st>>>(s->geo, pa);
      ++s->launches;
      if (s->geo.zlo > 0) { k_gt_cz_halo<<<nblk(s->nxy), 256, 0, s->st>>>(s->geo, s->dc, s->CO, s->co5); ++s->launches; }
      if (int rc = gt_ghost_rows(s)) return rc;
    } else {
      double* outs[4] = {s->RP, s->D, s->CYs, s->CZs};
      shear_arrays(s, s->An, outs, 4);
    }
    if (s->geo.zlo > 0 && !s->gs_tiled) { k_cz_halo<3><<<nblk(s->nxy), 256, 0, s->st>>>(s->geo, s->dc, s->CZs); ++s->launches; }
  } else {
    DIMSEL(s, k_prhs, gb, 256, s->geo, s->Fs, s->dc, 1, s->RP, s->D, s->CYs, s->CZs, nullptr);
  }
  tpop(s);
  }
  tpush(s, "fluid.6.pressure-solve");
  if (int rc = solve_pressure(s)) return rc;
  tpop(s);
  XCH(s, 1, s->pc);
  tpush(s, "fluid.7.correction");
  // convergence indicator (fluid.hpp:174-179): computed now into resid[iteration], fetched lazily
  double* rdst = s->resid + (s->iter_count < 4096 ? s->iter_count : 4095);
  if (s->iter_count >= 4095) CK(cudaMemsetAsync(rdst, 0, sizeof(double), s->st));
  { CorrArgs a; a.pc = s->pc; a.dc = s->dc; a.Fs = s->Fs; a.F = s->F[L_IC];
    for (int d = 0; d < 3; ++d) { a.u[d] = s->u[L_IC][d]; a.uprev[d] = s->u[L_IP][d]; }
    a.resid = nullptr;
    if (fast) {
      a.resid = rdst;   // interior cells: the maximum is taken where the new velocity is formed; the shell has its own pass
      k_fd_correct<<<fg, FT_THREADS, 0, s->st>>>(s->geo, s->slow, a); ++s->launches;
      if (s->nslow) {
        a.resid = nullptr;
        k_correct<3><<<gbl, 256, 0, s->st>>>(gl, a); ++s->launches;
        k_resid_list<<<gbl, 256, 0, s->st>>>(gl, cp3(s->u[L_IC]), cp3(s->u[L_IP]), rdst); ++s->launches;
      }
    } else DIMSEL(s, k_correct, gb, 256, s->geo, a); }
  tpop(s);
  if (!fast) {
    if (s->dim == 3) { k_resid<3><<<gb, 256, 0, s->st>>>(cp3(s->u[L_IC]), cp3(s->u[L_IP]), s->nc, rdst); }
    else { k_resid<2><<<gb, 256, 0, s->st>>>(cp3(s->u[L_IC]), cp3(s->u[L_IP]), s->nc, rdst); }
    ++s->launches;
  }
  if (c.simpler) {   // SIMPLER (fluid.hpp:1060-1155): pressure from the momentum equations evaluated on the new velocity
    tpush(s, "fluid.8.simpler");
    SimplerArgs a;
    for (int t = 0; t < 7; ++t) a.A[t] = s->A[t];
    for (int n = 0; n < 3; ++n) { a.R[n] = s->R[n]; a.uc[n] = s->u[L_IC][n] ? s->u[L_IC][n] : s->zero; a.up[n] = s->u[L_IP][n] ? s->u[L_IP][n] : s->zero;
                                  a.fcr[n] = s->fcr[n]; a.gp[n] = s->gp[n]; a.force[n] = s->force[n]; a.fev[n] = s->G[n]; }
    a.dc = s->dc; a.F = s->F[L_IC]; a.p = s->p[L_IC]; a.rc = c.rhie_chow_factor;
    a.RP = s->RP; a.CO = s->CO; a.out_mode = s->gs_tiled ? 2 : 1;
    a.co5.plane = s->co5.plane; a.co5.nxp = s->co5.nxp; a.co5.pad = GT_PAD;
    DIMSEL(s, k_simpler_eval, gb, 256, s->geo, a);
    DIMSEL(s, k_simpler_rhs, gb, 256, s->geo, a);
    if (int rc = solve_pressure(s, true)) return rc;
    tpop(s);
  }
  ++s->iter_count;
  s->last_resid = -1.;   // not fetched yet
  return 0;
}

extern "C" int hg_fluid_convergence_indicator(hg_handle s, double* out) {
  if (!s || !out) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  if (s->iter_count == 0) { *out = 1.; return 0; }
  if (s->last_resid < 0.) {
    const int idx = s->iter_count - 1 < 4095 ? s->iter_count - 1 : 4095;
    if (int rc = slab_reduce(s, s->resid + idx, 1, 0, s->hscal)) return rc;
    s->last_resid = s->hscal[0];
  }
  *out = s->last_resid;
  return 0;
}

extern "C" int hg_fluid_is_converged(hg_handle s, int* out) {   // solver.hpp:733-736
  if (!s || !out) return HG_ERR_INVALID;
  if (s->iter_count >= s->cfg.num_iterations_limit) { *out = 1; return 0; }
  if (!(s->cfg.convergence_tolerance > 0.)) { *out = 0; return 0; }   // indicator >= 0 is never < tol
  double r; if (int rc = hg_fluid_convergence_indicator(s, &r)) return rc;
  *out = r < s->cfg.convergence_tolerance;
  return 0;
}

extern "C" int hg_fluid_finish_step(hg_handle s) {   // fluid.hpp:1159-1169, conv_diff.hpp:252-259
  if (!s) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  // time_prev <- time_curr, time_curr <- iter_curr: rotate buffers, iter_curr keeps the old time_prev storage
  auto rot = [](double*& tc, double*& tp, double*& ic) { double* o = tp; tp = tc; tc = ic; ic = o; };
  rot(s->p[L_TC], s->p[L_TP], s->p[L_IC]);
  rot(s->F[L_TC], s->F[L_TP], s->F[L_IC]);
  for (int d = 0; d < s->dim; ++d) rot(s->u[L_TC][d], s->u[L_TP][d], s->u[L_IC][d]);
  if (s->defer) {   // one pass over the four arrays, flags read at the end of the step
    Nan4 q; q.n = 1 + s->dim;
    q.a[0] = s->p[L_TC]; q.flag[0] = s->flag + 4 + NF_FIN_P;
    for (int d = 0; d < 3; ++d) { q.a[1 + d] = d < s->dim ? s->u[L_TC][d] : s->p[L_TC]; q.flag[1 + d] = s->flag + 4 + NF_FIN_U; }
    LAUNCH(s, k_nan_flag4, nblk(s->nc), 256, q, s->nc);
  } else {
    if (int rc = check_nan(s, s->p[L_TC], s->nc, NF_FIN_P)) return rc;
    for (int d = 0; d < s->dim; ++d) if (int rc = check_nan(s, s->u[L_TC][d], s->nc, NF_FIN_U)) return rc;
  }
  s->time_fluid += s->dt;
  return 0;
}

extern "C" int hg_fluid_auto_time_step(hg_handle s, double* out) {
  if (!s || !out) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  double init = 1e10;
  CK(cudaMemcpyAsync(s->scal + 1, &init, sizeof(double), cudaMemcpyHostToDevice, s->st));
  DIMSEL(s, k_auto_dt, nblk(s->nc), 256, s->geo, s->F[L_TC], s->scal + 1);
  if (int rc = slab_reduce(s, s->scal + 1, 1, 1, s->hscal + 1)) return rc;
  *out = s->hscal[1];
  return 0;
}

extern "C" int hg_set_time_step(hg_handle s, double dt_fluid, double dt_adv) {
  if (!s) return HG_ERR_INVALID;
  s->dt = dt_fluid; s->dt_adv = dt_adv;
  return 0;
}

extern "C" int hg_advection_step(hg_handle s) {   // advection.hpp:417-545
  if (!s) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  const hg_config& c = s->cfg;
  const int num_stages = c.tvd_split ? s->dim : 1;
  for (int ph = 0; ph < c.num_phases; ++ph) {
    // StartStep: time_prev = time_curr (values); the update reads time_curr and writes a fresh buffer
    // flux of this phase: mixture flux + its slip flux (advection.hpp:449-454); F* of the SIMPLE iteration is free here
    const double* Fadv = s->F[L_TC];
    if (s->any_slip) { LAUNCH(s, k_face_add, nblk(s->nf), 256, s->F[L_TC], s->fslip[ph], s->Fs, s->nf); Fadv = s->Fs; }
    double* src = s->pd[ph][L_TC];
    CK(cudaMemcpyAsync(s->pd[ph][L_TP], src, s->nc * sizeof(double), cudaMemcpyDeviceToDevice, s->st));
    double* bufs[2] = {s->pd[ph][L_IC], s->pd[ph][L_IP]};
    const double* in = src;
    double* out = nullptr;
    for (int stage = 0; stage < num_stages; ++stage) {
      out = bufs[stage % 2];
      XCH(s, 2, const_cast<double*>(in));
      if (s->fast) {
        k_fe_advect<<<fast_grid(s), FT_THREADS, 0, s->st>>>(s->geo, s->slow2, in, Fadv, s->dt_adv, num_stages, stage, out); ++s->launches;
        if (s->nslow2) { k_advect<3><<<nblk(s->nslow2), 256, 0, s->st>>>(list_geo(s, 2), in, s->pd_init[ph], Fadv, s->dt_adv, num_stages, stage, out); ++s->launches; }
      } else DIMSEL(s, k_advect, nblk(s->nc), 256, s->geo, in, s->pd_init[ph], Fadv, s->dt_adv, num_stages, stage, out);
      in = out;
    }
    if (std::fabs(c.sharp) > 1e-10) {   // interface sharpening, once per field after the stages (advection.hpp:479-529)
      double* const sharpened = bufs[num_stages % 2];   // (the input of the last stage, no longer needed, or the other buffer)
      XCH(s, 2, out);
      P3 gc; CP3 gcc; for (int d = 0; d < 3; ++d) { gc.p[d] = s->G[d]; gcc.p[d] = s->G[d]; }
      DIMSEL(s, k_grad_pd, nblk(s->nc), 256, s->geo, out, s->pd_init[ph], gc);
      XCH(s, 1, s->G[0], s->G[1], s->dim > 2 ? s->G[2] : nullptr);
      DIMSEL(s, k_sharpen, nblk(s->nc), 256, s->geo, out, s->pd_init[ph], gcc, Fadv, s->dt_adv, c.sharp, c.density[ph], sharpened);
      out = sharpened;
    }
    // FinishStep: time_curr = iter_curr
    if (out == s->pd[ph][L_IC]) std::swap(s->pd[ph][L_TC], s->pd[ph][L_IC]);
    else std::swap(s->pd[ph][L_TC], s->pd[ph][L_IP]);
  }
  s->time_adv += s->dt_adv;
  return 0;
}

extern "C" int hg_heat_step(hg_handle s) {   // heat.hpp:69-84 + conv_diff.hpp:118-259, one iteration
  if (!s) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  const hg_config& c = s->cfg;
  if (int rc = check_nan(s, s->T[L_TC], s->nc, NF_HEAT_INIT)) return rc;
  XCH(s, 1, s->T[L_TC]);   // slabs: face values of the temperature across the interfaces
  const unsigned gb = nblk(s->nc);
  // gradient of the temperature for the deferred upwind correction (conv_diff.hpp:135)
  P3 g3; for (int d = 0; d < 3; ++d) g3.p[d] = s->G[d];
  if (s->dim == 3) { k_interp_grad<3, K_TEMP><<<gb, 256, 0, s->st>>>(s->geo, s->T[L_TC], 0, g3); }
  else { k_interp_grad<2, K_TEMP><<<gb, 256, 0, s->st>>>(s->geo, s->T[L_TC], 0, g3); }
  ++s->launches;
  XCH(s, 1, s->G[0], s->G[1], s->dim > 2 ? s->G[2] : nullptr);   // gradients of the upwind cells across the interfaces
  AsmArgs a;
  for (int n = 0; n < 3; ++n) { a.prev[n] = s->T[L_TC]; a.tc[n] = s->T[L_TC]; a.tp[n] = s->T[L_TP]; a.src[n] = s->zero; a.R[n] = s->R[n]; }
  for (int q = 0; q < 9; ++q) a.grad[q] = s->G[q % 3];
  a.rho = nullptr; a.mu = s->kc; a.F = s->F[L_TC];
  bdf_coeffs(c.dt /* HeatSolver keeps the time step of its constructor (hydro2d.hpp:684) */, c.time_second_order_heat, a.co);
  a.relax = c.heat_relaxation_factor;
  for (int t = 0; t < 7; ++t) a.A[t] = s->A[t];
  a.coeffsum = nullptr; a.coeffsum_div = 1.; a.out_sheared = 1;
  if (s->dim == 3) { k_assemble<3, K_TEMP, 1><<<gb, 256, 0, s->st>>>(s->geo, a); }
  else { k_assemble<2, K_TEMP, 1><<<gb, 256, 0, s->st>>>(s->geo, a); }
  ++s->launches;
  { int it; double df;
    if (int rc = solve_system(s, c.linear_solver_heat, s->R[0], s->X[0], s->PP, s->RP, &it, &df)) return rc; }
  CP3 prev, X; P3 curr;
  for (int d = 0; d < 3; ++d) { prev.p[d] = s->T[L_TC]; X.p[d] = s->X[d]; curr.p[d] = s->T[L_IC]; }
  if (s->dim == 3) { k_apply_corr<3, 1><<<gb, 256, 0, s->st>>>(s->geo, prev, X, curr); }
  else { k_apply_corr<2, 1><<<gb, 256, 0, s->st>>>(s->geo, prev, X, curr); }
  ++s->launches;
  // FinishStep: time_prev <- time_curr <- iter_curr
  double* o = s->T[L_TP]; s->T[L_TP] = s->T[L_TC]; s->T[L_TC] = s->T[L_IC]; s->T[L_IC] = o;
  if (int rc = check_nan(s, s->T[L_TC], s->nc, NF_HEAT_FIN)) return rc;
  if (!s->defer && s->lu_tiled) if (int rc = lt_check(s)) return rc;
  return 0;
}

static int step_body(hg_state* s, int* nadv_out) {
  const hg_config& c = s->cfg;
  int rc;
  if (c.dt_auto) {
    double dtm; if ((rc = hg_fluid_auto_time_step(s, &dtm))) return rc;
    s->dt = dtm * c.cfl; s->dt_adv = dtm * c.cfl_advection;
  }
  tpush(s, "step.fluid");
  if ((rc = hg_fluid_start_step(s))) return rc;
  if (c.fluid_enable) {
    int conv; if ((rc = hg_fluid_is_converged(s, &conv))) return rc;
    while (!conv) {
      if ((rc = hg_fluid_make_iteration(s))) return rc;
      if ((rc = hg_fluid_is_converged(s, &conv))) return rc;
    }
  }
  if ((rc = hg_fluid_finish_step(s))) return rc;
  tpop(s);
  int nadv = 0;
  if (c.advection_enable) {
    tpush(s, "step.advection");
    while (s->time_adv < s->time_fluid - 0.5 * s->dt_adv) { if ((rc = hg_advection_step(s))) return rc; ++nadv; }
    tpop(s);
  }
  if (c.heat_enable) { tpush(s, "step.heat"); if ((rc = hg_heat_step(s))) return rc; tpop(s); }
  if ((rc = hg_update_properties(s))) return rc;
  *nadv_out = nadv;
  return calc_stat_enqueue(s, true);
}
// hydro<Mesh>::step() (hydro2d.hpp:1531-1621) in two halves: hg_step_begin enqueues the whole step (it only waits where the
// control flow needs a device result: stop tests with a tolerance, dt_auto, slab reductions), hg_step_end waits for the status
// block and decodes it.  Between the two the caller can queue the transfers of the next step (hg_set_field_async).
extern "C" int hg_step_begin(hg_handle s) {
  if (!s) return HG_ERR_INVALID;
  if (s->step_pending) { s->err = "hg_step_begin: the previous step has not been ended (hg_step_end)"; return HG_ERR_INVALID; }
  cudaSetDevice(s->dev);
  const size_t tdepth = s->timer_stack.size();
  tpush(s, "step");
  s->sweeps_total = 0; s->nsolves = 0;
  // the device is waited for where the control flow needs a result (stop tests with a tolerance, dt_auto) and once at the
  // end: NaN flags, solver status words, sweep counts and statistics come back in one block
  s->defer = true;
  CK(cudaMemsetAsync(s->flag + 4, 0, NF_COUNT * sizeof(int), s->st));
  s->step_nadv = 0;
  const int rc = step_body(s, &s->step_nadv);
  s->defer = false;
  while (s->timer_stack.size() > tdepth) tpop(s);   // also on error paths
  if (rc) return rc;
  s->step_pending = true;
  return 0;
}
extern "C" int hg_step_end(hg_handle s, hg_step_stats* stats) {
  if (!s) return HG_ERR_INVALID;
  if (!s->step_pending) { s->err = "hg_step_end without hg_step_begin"; return HG_ERR_INVALID; }
  cudaSetDevice(s->dev);
  s->step_pending = false;
  if (int rc = calc_stat_finish(s, nullptr)) return rc;
  const double* sb = s->hscal + ST_STATUS;
  if (sb[8] != 0.) { s->err = "k_gs_tiled: dependency wait timed out"; return HG_ERR_CUDA; }
  if (sb[9] != 0.) { s->err = "k_lu_tiled: dependency wait timed out"; return HG_ERR_CUDA; }
  for (int q = 0; q < NF_COUNT; ++q) if (sb[q] != 0.) { s->err = nan_msgs[q]; return HG_ERR_NAN; }
  s->sweeps_total += (int)sb[11];
  if (s->nsolves > 0) s->last_diff = sb[12];
  s->stat.simple_iterations = s->iter_count;
  if (s->iter_count > 0) s->last_resid = sb[10];
  s->stat.convergence_indicator = s->iter_count > 0 ? sb[10] : 1.;
  s->stat.pressure_sweeps_total = s->sweeps_total;
  s->stat.pressure_last_diff = s->last_diff;
  s->stat.advection_substeps = s->step_nadv;
  s->stat.dt = s->dt; s->stat.time = s->time_fluid;
  if (stats) *stats = s->stat;
  return 0;
}
extern "C" int hg_step(hg_handle s, hg_step_stats* stats) {
  if (int rc = hg_step_begin(s)) return rc;
  return hg_step_end(s, stats);
}

extern "C" int hg_run(hg_handle s, int nsteps, hg_step_stats* last) {
  for (int i = 0; i < nsteps; ++i) if (int rc = hg_step(s, i == nsteps - 1 ? last : nullptr)) return rc;
  return 0;
}

// ------------------------------------------------------------------ lifetime
extern "C" void hg_config_defaults(hg_config* c) {   // examples/general.hydroconf
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

static int fail_create(hg_state* s, int code, const std::string& msg) {
  g_create_err = msg;
  if (s) {
    for (void* p : s->allocs) cudaFree(p);
    if (s->hscal) cudaFreeHost(s->hscal);
    if (s->hdiffs) cudaFreeHost(s->hdiffs);
    if (s->st) cudaStreamDestroy(s->st);
    delete s;
  }
  return code;
}

// Initial fields, first UpdateFluidProperties + CalcStat (hydro2d.hpp:928-972).  On several GPUs this needs the
// neighbouring slabs (smoothing, face fluxes), so it runs when the ranks have been linked.
static int init_fields(hg_state* s) {
  const hg_config& cfg = s->cfg;
  Geo& g = s->geo;
  const int dim = s->dim;
  const long long nc = s->nc, nf = s->nf;
  // initial fields
  {
    InitArgs a; memset(&a, 0, sizeof a);
    for (int d = 0; d < 3; ++d) {
      a.v0[d] = cfg.initial_velocity[d]; a.sin_n[d] = cfg.initial_sin_n[d];
      a.A1[d] = cfg.A1[d]; a.B1[d] = cfg.B1[d]; a.A2[d] = cfg.A2[d]; a.B2[d] = cfg.B2[d]; a.IC[d] = cfg.IC[d]; a.IC2[d] = cfg.IC2[d];
      a.u[d] = d < dim ? s->u[L_TC][d] : s->w1;
    }
    a.pois = cfg.initial_pois; a.sin_on = cfg.initial_sin_enable; a.sin_lambda = cfg.initial_sin_lambda; a.sin_phase = cfg.initial_sin_phase;
    a.IR = cfg.IR; a.IR2 = cfg.IR2; a.np = cfg.num_phases;
    for (int p = 0; p < 3; ++p) { a.density[p] = cfg.density[p]; a.ivf[p] = cfg.initial_volume_fraction[p]; a.pd[p] = p < cfg.num_phases ? s->pd[p][L_TC] : s->w1; }
    DIMSEL(s, k_init_fields, nblk(nc), 256, g, a);
    for (int ph = 1; ph < cfg.num_phases; ++ph) {
      if (cfg.initial_volume_fraction_smooth_times > 0) {
        if (int rc = smooth_field(s, s->pd[ph][L_TC], cfg.initial_volume_fraction_smooth_times, s->w1)) return rc;
        cudaMemcpyAsync(s->pd[ph][L_TC], s->w1, nc * sizeof(double), cudaMemcpyDeviceToDevice, s->st);
      }
    }
    LAUNCH(s, k_pd0, nblk(nc), 256, cfg.num_phases, cfg.density[0], cfg.density[1], cfg.density[2],
           cfg.num_phases > 1 ? s->pd[1][L_TC] : s->zero, cfg.num_phases > 2 ? s->pd[2][L_TC] : s->zero, s->pd[0][L_TC], nc);
    for (int ph = 0; ph < cfg.num_phases; ++ph) {
      cudaMemcpyAsync(s->pd_init[ph], s->pd[ph][L_TC], nc * sizeof(double), cudaMemcpyDeviceToDevice, s->st);
      cudaMemcpyAsync(s->pd[ph][L_TP], s->pd[ph][L_TC], nc * sizeof(double), cudaMemcpyDeviceToDevice, s->st);
    }
    for (int d = 0; d < dim; ++d) cudaMemcpyAsync(s->u[L_TP][d], s->u[L_TC][d], nc * sizeof(double), cudaMemcpyDeviceToDevice, s->st);
    CP3 uu; for (int d = 0; d < 3; ++d) uu.p[d] = d < dim ? s->u[L_TC][d] : s->zero;
    XCH(s, 1, s->u[L_TC][0], s->u[L_TC][1], s->u[L_TC][2]);
    DIMSEL(s, k_init_flux, nblk(nc), 256, g, uu, cfg.meshvel[0], cfg.meshvel[1], cfg.meshvel[2], s->F[L_TC]);
    cudaMemcpyAsync(s->F[L_TP], s->F[L_TC], nf * sizeof(double), cudaMemcpyDeviceToDevice, s->st);
    if (cfg.heat_enable) {
      LAUNCH(s, k_fill, nblk(nc), 256, s->T[L_TC], cfg.temperature_initial, nc);
      LAUNCH(s, k_fill, nblk(nc), 256, s->T[L_TP], cfg.temperature_initial, nc);
    }
  }
  if (int rc = hg_update_properties(s)) return rc;
  if (int rc = hg_calc_stat(s, nullptr)) return rc;
  CK(cudaStreamSynchronize(s->st));
  CK(cudaGetLastError());
  s->initialised = true;
  return 0;
}

extern "C" int hg_create(const hg_config* cfg, hg_handle* out) {
  if (!cfg || !out) { g_create_err = "null argument"; return HG_ERR_INVALID; }
  *out = nullptr;
  if ((cfg->dim != 2 && cfg->dim != 3) || cfg->Nx < 1 || cfg->Ny < 1 || (cfg->dim == 3 && cfg->Nz < 1))
    return fail_create(nullptr, HG_ERR_INVALID, "bad mesh size / dim");
  if (cfg->num_phases < 1 || cfg->num_phases > HG_MAX_PHASES) return fail_create(nullptr, HG_ERR_INVALID, "num_phases must be 1..3");
  if (cfg->simpler && (cfg->world_size > 1 || (cfg->linear_solver_pressure != HG_LS_GAUSS_SEIDEL && cfg->linear_solver_pressure != HG_LS_JACOBI)))
    return fail_create(nullptr, HG_ERR_INVALID, "simpler 1 runs on one GPU with gauss_seidel or jacobi for the pressure system");
  if (cfg->velocity_is_carrier) return fail_create(nullptr, HG_ERR_INVALID, "velocity_is_carrier 1 is not on the GPU path");
  for (int ph = 0; ph < cfg->num_phases; ++ph)
    if (cfg->enable_settling[ph] && cfg->world_size > 1) return fail_create(nullptr, HG_ERR_INVALID, "multi-GPU slabs: phase slip is not decomposed");
  if (cfg->force_geometric_average) return fail_create(nullptr, HG_ERR_INVALID, "force_geometric_average 1 is not on the GPU path");
  if (cfg->meshvel_auto < 0 || cfg->meshvel_auto > 2) return fail_create(nullptr, HG_ERR_INVALID, "Unknown meshvel_auto");
  if (cfg->meshvel_auto && cfg->num_phases < 2) return fail_create(nullptr, HG_ERR_INVALID, "meshvel_auto follows phase 1: num_phases >= 2");
  if (cfg->meshvel_auto && cfg->world_size > 1) return fail_create(nullptr, HG_ERR_INVALID, "multi-GPU slabs: meshvel_auto is not decomposed");
  for (int sd = 0; sd < 2 * cfg->dim; ++sd)
    if (cfg->condition_kind[sd] == HG_BC_OUTLET && cfg->world_size > 1) return fail_create(nullptr, HG_ERR_INVALID, "multi-GPU slabs: outlet conditions are not decomposed");
  for (int id : {cfg->linear_solver_velocity, cfg->linear_solver_pressure, cfg->linear_solver_heat})
    if (id < HG_LS_LU || id > HG_LS_JACOBI) return fail_create(nullptr, HG_ERR_INVALID, "Unknown linear solver");
  if (cfg->world_size > 1 && (cfg->linear_solver_velocity != HG_LS_LU || (cfg->heat_enable && cfg->linear_solver_heat != HG_LS_LU)))
    return fail_create(nullptr, HG_ERR_INVALID, "multi-GPU slabs support lu for the momentum and temperature systems");
  if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size || cfg->world_size > 64)
    return fail_create(nullptr, HG_ERR_INVALID, "bad world_size / rank");
  if (cfg->world_size > 1 && (cfg->dim != 3 || cfg->Nz < 2 * cfg->world_size))
    return fail_create(nullptr, HG_ERR_INVALID, "z-slab decomposition needs dim 3 and at least 2 planes per rank");
  if (cfg->world_size > 1 && cfg->linear_solver_pressure != HG_LS_GAUSS_SEIDEL && cfg->linear_solver_pressure != HG_LS_JACOBI)
    return fail_create(nullptr, HG_ERR_INVALID, "multi-GPU slabs support gauss_seidel and jacobi for the pressure system");
  if (cfg->world_size > SLAB_MAX_WORLD) return fail_create(nullptr, HG_ERR_INVALID, "world_size too large");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail_create(nullptr, HG_ERR_NO_DEVICE, "no CUDA device: the GPU path has no CPU fallback");
  hg_state* s = new hg_state();
  s->cfg = *cfg;
  s->dev = cfg->device;
  if (s->dev < 0 || s->dev >= ndev) return fail_create(s, HG_ERR_INVALID, "bad device index");
  if (cudaSetDevice(s->dev) != cudaSuccess) return fail_create(s, HG_ERR_CUDA, "cudaSetDevice failed");
  int coop = 0; cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s->dev);
  if (!coop) return fail_create(s, HG_ERR_CUDA, "device lacks cooperative launch");
  if (cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) != cudaSuccess) return fail_create(s, HG_ERR_CUDA, "stream create failed");
  const int dim = s->dim = cfg->dim;
  // z-slab owned by this rank (hydro_b200/parallel.py: slab_range)
  s->world = cfg->world_size; s->rank = cfg->rank;
  s->nzg = dim > 2 ? cfg->Nz : 1;
  { const int base = s->nzg / s->world, rem = s->nzg % s->world;
    s->k0 = s->rank * base + std::min(s->rank, rem); s->k1 = s->k0 + base + (s->rank < rem ? 1 : 0); }
  s->n[0] = cfg->Nx; s->n[1] = cfg->Ny; s->n[2] = s->k1 - s->k0;
  s->nxy = (long long)s->n[0] * s->n[1];
  s->nc = s->nxy * s->n[2];
  s->ncg = s->nxy * s->nzg;
  Geo& g = s->geo;
  memset(&g, 0, sizeof g);
  g.dim = dim; g.sy = s->n[0]; g.sz = (long long)s->n[0] * s->n[1];
  g.k0 = s->k0; g.nzg = s->nzg;
  g.zlo = s->rank > 0 ? HG_HALO : 0; g.zhi = s->rank + 1 < s->world ? HG_HALO : 0;
  g.vol = 1.;
  const int nglob[3] = {cfg->Nx, cfg->Ny, s->nzg};
  for (int d = 0; d < 3; ++d) {
    g.n[d] = s->n[d]; g.lb[d] = cfg->A[d];
    g.h[d] = d < dim ? (cfg->B[d] - cfg->A[d]) / nglob[d] : 1.;
    if (d < dim) g.vol *= g.h[d];
  }
  for (int d = 0; d < 3; ++d) { g.area[d] = 1.; if (d < dim) for (int e = 0; e < dim; ++e) if (e != d) g.area[d] *= g.h[e]; }
  s->nf = 0;
  for (int d = 0; d < 3; ++d) {
    g.foff[d] = s->nf;
    if (d < dim) s->nf += (long long)(s->n[0] + (d == 0)) * (s->n[1] + (d == 1)) * (s->n[2] + (d == 2));
  }
  g.np = s->n[0] + s->n[1] + s->n[2] - 2;
  s->nsh = (long long)(g.np + 2) * s->n[1] * s->n[0];   // one extra plane at each end (slab halo cells)
  for (int sd = 0; sd < 6; ++sd) { g.bckind[sd] = cfg->condition_kind[sd]; for (int d = 0; d < 3; ++d) g.bcvel[sd][d] = cfg->condition_velocity[sd][d]; }
  g.bckind[6] = HG_BC_WALL;
  for (int d = 0; d < 3; ++d) { g.heat_lb[d] = cfg->heat_box_lb[d]; g.heat_rt[d] = cfg->heat_box_rt[d]; }
  g.heat_T = cfg->heat_box_temperature;
  g.pfix = HG_NO_CELL; g.pfix_value = cfg->pressure_fixed_value;
  g.excl = nullptr;

  // ---- allocation: cell arrays carry HG_HALO planes on both sides (pointer = first owned cell)
  const long long nf = s->nf;
  {
    bool okA = true;
    const long long ncell = s->nxy * (s->n[2] + 2 * HG_HALO);
    auto take = [&](long long n) -> double* { double* q = nullptr; if (okA) okA = dalloc(s, &q, n) == 0; return q; };
    auto cells = [&]() -> double* { double* q = take(ncell); return q ? q + HG_HALO * s->nxy : nullptr; };
    hg_state& T = *s;
    for (int l = 0; l < 4; ++l) {
      for (int d = 0; d < dim; ++d) T.u[l][d] = cells();
      T.p[l] = cells(); T.F[l] = take(nf);
      if (cfg->heat_enable && l < 3) T.T[l] = cells();
      for (int ph = 0; ph < cfg->num_phases; ++ph) T.pd[ph][l] = cells();
    }
    for (int ph = 0; ph < cfg->num_phases; ++ph) { T.vf[ph] = cells(); T.pd_init[ph] = cells(); }
    T.rho_raw = cells(); T.mu_raw = cells(); T.rho = cells(); T.mu = cells(); T.kc = cells();
    T.dc = cells(); T.Fs = take(nf); T.pc = cells(); T.w1 = cells(); T.w2 = cells(); T.zero = cells();
    for (int d = 0; d < dim; ++d) { T.force[d] = cells(); T.stforce[d] = cells(); T.gp[d] = cells(); T.fcr[d] = cells(); T.fs[d] = cells(); }
    for (int q = 0; q < dim * dim; ++q) T.G[q] = cells();
    for (int q = 0; q < 7; ++q) T.A[q] = take(s->nsh);
    for (int q = 0; q < (dim == 3 ? 10 : 7); ++q) T.An[q] = cells();   // (2-D: natural-layout rows of the Jacobi solver)
    for (int n = 0; n < dim; ++n) { T.R[n] = take(s->nsh); T.X[n] = take(s->nsh); }
    // the tile sweeps read a few hyperplanes beyond both ends of the solution and of the x+/y+ coefficient arrays:
    // GT_PAD zero hyperplanes around them (never written)
    const long long padn = dim == 3 ? (long long)GT_PAD * s->nxy : 0;
    auto take_padded = [&](long long n) -> double* { double* q = take(n + 2 * padn); return q ? q + padn : nullptr; };
    T.D = take_padded(s->nsh); T.CYs = take_padded(s->nsh); T.CZs = take(dim > 2 ? s->nsh : 1); T.RP = take(s->nsh);
    T.PP = take_padded(s->nsh); T.PPsave = take(s->nsh);
    T.scal = take(64); T.resid = take(4096); T.sorres = take(2 * 4096);
    // pressure sweeps: the box dataflow (k_gs_tiled) in 3-D, on one GPU and on every z-slab of a decomposed run;
    // HYDRO_GS_KERNEL=hyperplane / HYDRO_GS_SLAB_KERNEL=hyperplane keep the pipelined hyperplane kernel (also used in 2-D)
    { const char* e = getenv("HYDRO_GS_KERNEL");
      const char* es = getenv("HYDRO_GS_SLAB_KERNEL");
      // (a slab runs up to GT_B - 1 planes of the slab above as ghost planes and sends its bottom GT_B planes down)
      const bool slab_ok = s->world == 1 || (!(es && !strcmp(es, "hyperplane")) && s->nzg / s->world >= GT_B);
      T.gs_tiled = dim == 3 && slab_ok && cfg->linear_solver_pressure == HG_LS_GAUSS_SEIDEL && !(e && !strcmp(e, "hyperplane")) &&
                   8LL * (GT_PAD + 2) * s->nxy < (1LL << 31);   // 32-bit byte offsets inside k_gs_tiled
    }
    if (T.gs_tiled) {
      T.co5 = gt_co5(s->n[0], s->n[1], s->geo.np);
      T.CO = take(5 * T.co5.arr);   // zeroed
    }
    // buffers peers read or write: exchange staging (2 parities x 2 directions x SLAB_MAX_ARRAYS x HG_HALO planes),
    // mailbox (2 parities x world x SLAB_MAIL doubles) and flag words
    if (s->world > 1) {
      T.slab.xbuf = take(2LL * 2 * SLAB_MAX_ARRAYS * HG_HALO * s->nxy);
      T.slab.mail = take(2LL * s->world * SLAB_MAIL + 64);
      T.slab.ll = (uint4*)take(2LL * SLAB_LL_PLANES * s->nxy);   // 16-byte entries
    }
    if (!okA) return fail_create(s, HG_ERR_CUDA, "allocation failed: " + s->err);
  }
  bool ok = true;
  for (int sd = 0; sd < 2 * dim; ++sd) if (cfg->condition_kind[sd] == HG_BC_OUTLET) s->any_outlet = true;
  for (int ph = 0; ph < cfg->num_phases; ++ph) if (cfg->enable_settling[ph]) s->any_slip = true;
  if (s->any_slip) {
    for (int ph = 0; ph < cfg->num_phases && ok; ++ph) {
      ok = dalloc(s, &s->fslip[ph], s->nf) == 0;
      for (int d = 0; d < dim && ok; ++d) { double* q = nullptr; ok = dalloc(s, &q, s->nxy * (s->n[2] + 2 * HG_HALO)) == 0; s->slipv[ph][d] = q ? q + HG_HALO * s->nxy : nullptr; }
    }
  }
  if (s->any_outlet) {   // velocities of the outlet faces: [side][component][face of the side], zero at construction (OutletAuto)
    const long long pl = std::max({(long long)s->n[1] * s->n[2], (long long)s->n[0] * s->n[2], dim > 2 ? (long long)s->n[0] * s->n[1] : 0LL});
    g.outplane = pl;
    s->out_blocks = (int)nblk(2LL * dim * pl);
    const long long nterms = 2LL * s->n[1] * s->n[2] + 2LL * s->n[0] * s->n[2] + (dim > 2 ? 2LL * s->n[0] * s->n[1] : 0LL);
    s->out_terms = nterms;
    ok = dalloc(s, &s->outvel, 18 * pl) == 0 && dalloc(s, &s->outpart, 3 * nterms + 8) == 0;
    g.outvel = s->outvel;
  }
  if (ok) { int* fp = nullptr; ok = dalloc(s, &fp, 16) == 0; s->flag = fp; }
  if (ok) { unsigned char* ep = nullptr; ok = dalloc(s, &ep, s->nxy * (s->n[2] + 2 * HG_HALO)) == 0; s->excl = ep ? ep + HG_HALO * s->nxy : nullptr; }
  if (!ok || cudaMallocHost((void**)&s->hscal, 64 * 128 * sizeof(double)) != cudaSuccess ||
      cudaMallocHost((void**)&s->hinit, 64 * sizeof(double)) != cudaSuccess) return fail_create(s, HG_ERR_CUDA, "allocation failed: " + s->err);
  // unused component slots alias a zero array so structs of 3 pointers are always valid
  for (int d = dim; d < 3; ++d) {
    for (int l = 0; l < 4; ++l) s->u[l][d] = nullptr;
    s->force[d] = s->stforce[d] = s->gp[d] = s->fcr[d] = s->fs[d] = s->zero;
    s->R[d] = s->X[d] = s->PPsave;
  }
  for (int q = dim * dim; q < 9; ++q) s->G[q] = s->zero;

  // solver grids: persistent cooperative kernels, all CTAs co-resident
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, s->dev);
  s->num_sms = prop.multiProcessorCount;
  if (s->gs_tiled) {
    if (dalloc(s, &s->gt_ctl, 4, true)) return fail_create(s, HG_ERR_CUDA, "allocation failed: " + s->err);
    // entries of the row arrays without a cell: zero coefficients (set by the allocation), unit diagonal
    k_gt_co_fill<<<nblk(s->co5.arr), 256, 0, s->st>>>(s->CO + s->co5.arr, s->co5.arr);
    if (cudaStreamSynchronize(s->st) != cudaSuccess) return fail_create(s, HG_ERR_CUDA, "k_gt_co_fill failed");
    if (gt_plans_allocate(s)) return fail_create(s, HG_ERR_CUDA, "allocation failed: " + s->err);
    if (getenv("HYDRO_GT_CLOCK")) { if (dalloc(s, &s->gt_clk, 16, true)) return fail_create(s, HG_ERR_CUDA, "allocation failed: " + s->err); }
    { // TMA descriptor of the row arrays: doubles [5][nhp][ny][nxp], box = one hyperplane under the footprint of a task
      typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
      void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn)
        return fail_create(s, HG_ERR_CUDA, "cuTensorMapEncodeTiled not available");
      const Co5& co = s->co5;
      const cuuint64_t dims[4] = {(cuuint64_t)s->n[0], (cuuint64_t)s->n[1], (cuuint64_t)co.nhp, 5};
      const cuuint64_t strides[3] = {(cuuint64_t)co.nxp * 8, (cuuint64_t)co.plane * 8, (cuuint64_t)co.arr * 8};
      const cuuint32_t box[4] = {GT_CW, GT_CH, 1, 5}, estr[4] = {1, 1, 1, 1};
      const CUresult r = ((EncodeFn)fn)(&s->tmco, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, s->CO, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail_create(s, HG_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); }
    if (cudaFuncSetAttribute(k_gs_tiled<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_gs_tiled<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM_BYTES) != cudaSuccess)
      return fail_create(s, HG_ERR_CUDA, "k_gs_tiled: shared memory request rejected");
  }
  { const char* e = getenv("HYDRO_LU_KERNEL");
    s->lu_tiled = dim == 3 && !(e && !strcmp(e, "hyperplane")); }
  if (s->lu_tiled) {
    if (cudaFuncSetAttribute(k_lu_tiled<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_RING_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_lu_tiled<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_RING_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_lu_tiled<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_RING_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_lu_tiled<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_RING_BYTES) != cudaSuccess)
      return fail_create(s, HG_ERR_CUDA, "k_lu_tiled: shared memory request rejected");
    const int nbi = (s->n[0] + LT_TX - 1) / LT_TX, nbj = (s->n[1] + LT_TY - 1) / LT_TY;
    std::vector<int2> boxes;
    for (int J = 0; J < nbj; ++J) for (int I = 0; I < nbi; ++I) boxes.push_back(make_int2(I, J));
    std::stable_sort(boxes.begin(), boxes.end(), [](const int2& p, const int2& q) { return p.x * LT_TX + p.y * LT_TY < q.x * LT_TX + q.y * LT_TY; });
    s->lt_nboxes = (int)boxes.size(); s->lt_nbi = nbi;
    if (dalloc(s, &s->lt_boxes, s->lt_nboxes, false) || dalloc(s, &s->lt_progress, s->lt_nboxes, true) || dalloc(s, &s->lt_ctl, 4, true))
      return fail_create(s, HG_ERR_CUDA, "allocation failed: " + s->err);
    cudaMemcpyAsync(s->lt_boxes, boxes.data(), boxes.size() * sizeof(int2), cudaMemcpyHostToDevice, s->st);
    cudaStreamSynchronize(s->st);
  }
  int occ_gs = 0, occ_lu = 0;
  if (dim == 3) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_gs, k_gs_persistent<3, true>, SOLVER_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_lu, k_lu_persistent<3>, SOLVER_THREADS, 0);
  } else {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_gs, k_gs_persistent<2, true>, SOLVER_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_lu, k_lu_persistent<2>, SOLVER_THREADS, 0);
  }
  if (occ_gs < 1 || occ_lu < 1) return fail_create(s, HG_ERR_CUDA, "solver kernel does not fit on an SM");
  s->grid_solver = prop.multiProcessorCount * std::min(occ_gs, 1);
  s->grid_lu = prop.multiProcessorCount * std::min(occ_lu, 1);
  if (cfg->solver_ctas > 0) {   // several ranks sharing one device (tests): their persistent kernels must be co-resident
    s->grid_solver = std::min(s->grid_solver, cfg->solver_ctas); s->grid_lu = std::min(s->grid_lu, cfg->solver_ctas);
  }

  // hyperplane tile table for the ordered sweeps (hg_solvers.cuh)
  {
    std::vector<int> tj, ti, toff(g.np + 1, 0), cum2(g.np, 0);
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    for (int kp = 0; kp < g.np; ++kp) {
      toff[kp] = (int)tj.size();
      for (int j0 = 0; j0 < ny; j0 += SOLVER_BY) {
        if (j0 > kp) break;
        const int j1 = std::min(j0 + SOLVER_BY - 1, ny - 1);
        const int ihi = std::min(nx - 1, kp - j0);             // largest valid i over the rows
        const int ilo = std::max(0, kp - j1 - (nz - 1));       // smallest valid i over the rows
        if (ilo > ihi) continue;
        for (int i0 = ilo; i0 <= ihi; i0 += SOLVER_BX) { tj.push_back(j0); ti.push_back(i0); }
      }
    }
    toff[g.np] = (int)tj.size();
    for (int kp = 0; kp < g.np; ++kp) cum2[kp] = (toff[kp + 1] - toff[kp]) + (kp >= 2 ? cum2[kp - 2] : 0);
    int *d_tj = nullptr, *d_ti = nullptr, *d_off = nullptr, *d_c2 = nullptr;
    if (dalloc(s, &d_tj, (long long)tj.size(), false) || dalloc(s, &d_ti, (long long)ti.size(), false) ||
        dalloc(s, &d_off, g.np + 1, false) || dalloc(s, &d_c2, g.np, false))
      return fail_create(s, HG_ERR_CUDA, "allocation failed: " + s->err);
    cudaMemcpy(d_tj, tj.data(), tj.size() * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(d_ti, ti.data(), ti.size() * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(d_off, toff.data(), toff.size() * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(d_c2, cum2.data(), cum2.size() * sizeof(int), cudaMemcpyHostToDevice);
    s->tt.tile_j0 = d_tj; s->tt.tile_i0 = d_ti; s->tt.tileoff = d_off; s->tt.cum2 = d_c2;
    unsigned long long* d_bar = nullptr;
    if (dalloc(s, &d_bar, 4, true)) return fail_create(s, HG_ERR_CUDA, "allocation failed: " + s->err);
    s->tt.bar = d_bar;
  }
  // rigid box -> excluded cells (hydro2d.hpp:409-416)
  {
    int* anyflag = s->flag + 1;
    cudaMemsetAsync(anyflag, 0, sizeof(int), s->st);
    k_excl_mask<<<nblk(s->nxy * (s->n[2] + g.zlo + g.zhi)), 256, 0, s->st>>>(g, cfg->box_A[0], cfg->box_A[1], cfg->box_A[2], cfg->box_B[0], cfg->box_B[1], cfg->box_B[2], s->excl, anyflag);
    ++s->launches;
    int h = 0; cudaMemcpyAsync(&h, anyflag, sizeof(int), cudaMemcpyDeviceToHost, s->st); cudaStreamSynchronize(s->st);
    s->any_excl = h != 0;
    g.excl = s->any_excl ? s->excl : nullptr;
  }
  // fixed pressure cell: FindNearestCell (mesh.hpp:411-420), first minimum in (global) raw order
  if (cfg->pressure_fixed_enable) {
    long long best = 0; double bd = 0.;
    for (int k = 0; k < s->nzg; ++k) for (int j = 0; j < s->n[1]; ++j) for (int i = 0; i < s->n[0]; ++i) {
      double x[3] = {g.lb[0] + (i + 0.5) * g.h[0], g.lb[1] + (j + 0.5) * g.h[1], dim > 2 ? g.lb[2] + (k + 0.5) * g.h[2] : 0.};
      double sq = 0.; for (int d = 0; d < dim; ++d) { double e = x[d] - cfg->pressure_fixed_point[d]; sq += e * e; }
      double dd = std::sqrt(sq);
      long long c = i + g.sy * j + g.sz * k;
      if (c == 0) { bd = dd; best = 0; } else if (dd < bd) { bd = dd; best = c; }
    }
    const int kg = (int)(best / g.sz);
    if (kg >= s->k0 - g.zlo && kg < s->k1 + g.zhi) g.pfix = best - (long long)s->k0 * g.sz;   // local index, may be in a halo plane
  }
  // interior / listed cells (hg_fast.cuh): 3-D meshes below 2^31 cells; HYDRO_FAST=0 keeps the generic kernels everywhere
  { const char* e = getenv("HYDRO_FAST");
    s->fast = dim == 3 && s->nc < (1LL << 31) && !(e && !strcmp(e, "0")); }
  if (s->fast) {
    for (int R = 1; R <= 2; ++R) {
      unsigned char*& mask = R == 1 ? s->slow : s->slow2;
      int*& lst = R == 1 ? s->slow_list : s->slow2_list;
      int& cnt = R == 1 ? s->nslow : s->nslow2;
      if (dalloc(s, &mask, s->nc, false)) return fail_create(s, HG_ERR_CUDA, "allocation failed: " + s->err);
      k_fast_mask<<<nblk(s->nc), 256, 0, s->st>>>(g, R, mask);
      std::vector<unsigned char> hm((size_t)s->nc);
      if (cudaMemcpyAsync(hm.data(), mask, (size_t)s->nc, cudaMemcpyDeviceToHost, s->st) != cudaSuccess ||
          cudaStreamSynchronize(s->st) != cudaSuccess) return fail_create(s, HG_ERR_CUDA, "k_fast_mask failed");
      std::vector<int> list;
      for (long long c = 0; c < s->nc; ++c) if (hm[(size_t)c]) list.push_back((int)c);
      cnt = (int)list.size();
      if (dalloc(s, &lst, std::max(cnt, 1), false)) return fail_create(s, HG_ERR_CUDA, "allocation failed: " + s->err);
      if (cnt) cudaMemcpy(lst, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice);
    }
    if (s->nslow == s->nc) s->fast = false;   // no interior cell
  }
  s->dt = cfg->dt; s->dt_adv = cfg->dt * cfg->advection_dt_factor;
  if (s->world > 1) {
    Slab& sl = s->slab;
    sl.world = s->world; sl.rank = s->rank; sl.has_lo = s->rank > 0; sl.has_hi = s->rank + 1 < s->world;
    sl.np_glob = s->n[0] + s->n[1] + s->nzg - 2;
    if (sl.has_lo) { const int base = s->nzg / s->world, rem = s->nzg % s->world; sl.nz_lo = base + (s->rank - 1 < rem ? 1 : 0); }
  }

  if (s->world > 1) {
    // load the kernels that only decomposed runs launch now: a first launch (lazy module loading) synchronises the
    // context, which dead-locks ranks that share a device while one waits for the other inside a kernel
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_slab_pack); cudaFuncGetAttributes(&fa, k_slab_signal); cudaFuncGetAttributes(&fa, k_slab_wait);
    cudaFuncGetAttributes(&fa, k_slab_unpack); cudaFuncGetAttributes(&fa, k_mail_post); cudaFuncGetAttributes(&fa, k_mail_wait);
    cudaFuncGetAttributes(&fa, k_cz_halo<3>); cudaFuncGetAttributes(&fa, k_flag_to_double);
    cudaFuncGetAttributes(&fa, k_gt_ghost_pack); cudaFuncGetAttributes(&fa, k_gt_ghost_unpack); cudaFuncGetAttributes(&fa, k_gt_cz_halo);
    cudaFuncGetAttributes(&fa, k_fa_grad); cudaFuncGetAttributes(&fa, k_fb_momentum); cudaFuncGetAttributes(&fa, k_fc_flux_rows);
    cudaFuncGetAttributes(&fa, k_fd_correct); cudaFuncGetAttributes(&fa, k_fe_advect); cudaFuncGetAttributes(&fa, k_prhs_co5);
    cudaFuncGetAttributes(&fa, k_ft_apply_corr<3>); cudaFuncGetAttributes(&fa, k_ft_pcorr); cudaFuncGetAttributes(&fa, k_ll_fill);
    cudaGetLastError();
  }
  // no allocation after this point on the stepping path: cudaMalloc synchronises the device, which dead-locks
  // ranks that share a device while one of them waits for the other inside a kernel
  if (ensure_sweep_capacity(s, std::max(cfg->lu_relaxed_num_iters_limit + 1, 64))) return fail_create(s, HG_ERR_CUDA, s->err);
  if (s->world == 1) {
    if (init_fields(s)) return fail_create(s, HG_ERR_CUDA, s->err);
  } else if (cudaStreamSynchronize(s->st) != cudaSuccess) {
    return fail_create(s, HG_ERR_CUDA, "initialisation failed");
  }
  *out = s;
  return 0;
}

// ------------------------------------------------------------------ linking the ranks of a slab decomposition
// One process per GPU: every rank exports CUDA IPC handles of the few buffers its neighbours touch
// (hg_ipc_export), the host side all-gathers the records (torch.distributed / MPI / files: plumbing), every rank
// imports them (hg_ipc_import).  Ranks living in one process (tests; several ranks on one device) are linked by
// pointer (hg_link_local).  Linking is collective: it ends with the initial fields, which need the neighbours.
struct hg_ipc_record { cudaIpcMemHandle_t xbuf, mail, ll; };
extern "C" size_t hg_ipc_record_size(void) { return sizeof(hg_ipc_record); }

extern "C" int hg_ipc_export(hg_handle s, void* buf, size_t cap) {
  if (!s || !buf || cap < sizeof(hg_ipc_record)) return HG_ERR_INVALID;
  if (s->world <= 1) { s->err = "hg_ipc_export: single-GPU handle"; return HG_ERR_INVALID; }
  cudaSetDevice(s->dev);
  hg_ipc_record r; memset(&r, 0, sizeof r);
  CK(cudaIpcGetMemHandle(&r.xbuf, s->slab.xbuf));
  CK(cudaIpcGetMemHandle(&r.mail, s->slab.mail));
  CK(cudaIpcGetMemHandle(&r.ll, s->slab.ll));
  memcpy(buf, &r, sizeof r);
  return 0;
}

static int finish_link(hg_state* s) {
  s->slab.mail_peer[s->rank] = s->slab.mail;
  s->slab.linked = true;
  return init_fields(s);
}

extern "C" int hg_ipc_import(hg_handle s, const void* all_records) {
  if (!s || !all_records) return HG_ERR_INVALID;
  if (s->world <= 1) { s->err = "hg_ipc_import: single-GPU handle"; return HG_ERR_INVALID; }
  cudaSetDevice(s->dev);
  const hg_ipc_record* rec = (const hg_ipc_record*)all_records;
  auto open = [&](const cudaIpcMemHandle_t& h, double** out) -> int {
    void* q = nullptr;
    CK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
    s->ipc_opened.push_back(q);
    *out = (double*)q;
    return 0;
  };
  Slab& sl = s->slab;
  for (int r = 0; r < s->world; ++r) if (r != s->rank) if (int rc = open(rec[r].mail, &sl.mail_peer[r])) return rc;
  if (sl.has_lo) {
    double* q = nullptr;
    if (int rc = open(rec[s->rank - 1].xbuf, &q)) return rc;
    sl.xbuf_lo = q;
    if (int rc = open(rec[s->rank - 1].ll, &q)) return rc;
    sl.ll_lower = (uint4*)q;
  }
  if (sl.has_hi) {
    double* q = nullptr;
    if (int rc = open(rec[s->rank + 1].xbuf, &q)) return rc;
    sl.xbuf_hi = q;
    if (int rc = open(rec[s->rank + 1].ll, &q)) return rc;
    sl.ll_upper = (uint4*)q;
  }
  return finish_link(s);
}

extern "C" int hg_link_local(hg_handle s, const hg_handle* all) {
  if (!s || !all) return HG_ERR_INVALID;
  if (s->world <= 1) { s->err = "hg_link_local: single-GPU handle"; return HG_ERR_INVALID; }
  cudaSetDevice(s->dev);
  Slab& sl = s->slab;
  for (int r = 0; r < s->world; ++r) {
    if (!all[r] || all[r]->world != s->world || all[r]->rank != r) { s->err = "hg_link_local: handles must be given in rank order"; return HG_ERR_INVALID; }
    if (all[r]->dev != s->dev) {
      cudaError_t e = cudaDeviceEnablePeerAccess(all[r]->dev, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { s->err = "peer access between the devices is not available"; return HG_ERR_CUDA; }
      cudaGetLastError();
    }
    sl.mail_peer[r] = all[r]->slab.mail;
  }
  if (sl.has_lo) { hg_state* o = all[s->rank - 1]; sl.xbuf_lo = o->slab.xbuf; sl.ll_lower = o->slab.ll; }
  if (sl.has_hi) { hg_state* o = all[s->rank + 1]; sl.xbuf_hi = o->slab.xbuf; sl.ll_upper = o->slab.ll; }
  return finish_link(s);
}

extern "C" int hg_destroy(hg_handle s) {
  if (!s) return 0;
  cudaSetDevice(s->dev);
  cudaStreamSynchronize(s->st);
  for (void* p : s->ipc_opened) cudaIpcCloseMemHandle(p);
  for (void* p : s->allocs) cudaFree(p);
  if (s->hscal) cudaFreeHost(s->hscal);
  if (s->hdiffs) cudaFreeHost(s->hdiffs);
  if (s->hinit) cudaFreeHost(s->hinit);
  for (auto& pl : s->gt_plans) { if (pl.stage) cudaFreeHost(pl.stage); if (pl.staged) cudaEventDestroy(pl.staged); }
  for (cudaEvent_t e : s->user_ev) if (e) cudaEventDestroy(e);
  for (auto& v : s->prof_ev) for (auto& pr : v) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  if (s->xfer_ev_tmp[0]) { cudaEventDestroy(s->xfer_ev_tmp[0]); cudaEventDestroy(s->xfer_ev_tmp[1]); }
  for (auto& kv : s->timers) { if (kv.second.a) { cudaEventDestroy(kv.second.a); cudaEventDestroy(kv.second.b); } }
  cudaStreamDestroy(s->st);
  if (s->st_h2d) cudaStreamDestroy(s->st_h2d);
  if (s->st_d2h) cudaStreamDestroy(s->st_d2h);
  for (auto& e : s->xfer_ev) cudaEventDestroy(e.second);
  for (auto& e : s->xfer_stage_ev) cudaEventDestroy(e.second);
  delete s;
  return 0;
}

extern "C" const char* hg_last_error(hg_handle s) { return s ? s->err.c_str() : g_create_err.c_str(); }
extern "C" size_t hg_num_cells(hg_handle s) { return s ? (size_t)s->nc : 0; }
extern "C" size_t hg_num_faces(hg_handle s) { return s ? (size_t)s->nf : 0; }
extern "C" long long hg_launch_count(hg_handle s) { return s ? s->launches : 0; }
extern "C" const char* hg_solver_kernel_name(hg_handle s, int which) {
  if (!s) return "";
  if (which == 0) {
    if (s->cfg.linear_solver_pressure == HG_LS_JACOBI) return "k_jacobi_sweep";
    if (s->cfg.linear_solver_pressure == HG_LS_LU_RELAXED) return "k_lur_forward";
    if (s->gs_tiled) return s->world > 1 ? "k_gs_tiled<LINK>" : "k_gs_tiled";
    return s->world > 1 ? "k_gs_persistent<LINK>" : "k_gs_persistent";
  }
  if (s->lu_tiled) return s->world > 1 ? "k_lu_tiled<LINK>" : "k_lu_tiled";
  return s->world > 1 ? "k_lu_persistent<LINK>" : "k_lu_persistent";
}
extern "C" int hg_device_synchronize(hg_handle s) {
  if (!s) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  CK(cudaStreamSynchronize(s->st));
  if (s->st_h2d) CK(cudaStreamSynchronize(s->st_h2d));
  if (s->st_d2h) CK(cudaStreamSynchronize(s->st_d2h));
  return 0;
}

static double* field_ptr(hg_state* s, int field, int layer, long long* n) {
  *n = s->nc;
  const int np = s->cfg.num_phases, dim = s->dim;
  if (field >= HG_F_VELOCITY_X && field <= HG_F_VELOCITY_Z) return field - HG_F_VELOCITY_X < dim ? s->u[layer][field - HG_F_VELOCITY_X] : nullptr;
  if (field >= HG_F_VELOCITY_PREV_X && field <= HG_F_VELOCITY_PREV_Z) return field - HG_F_VELOCITY_PREV_X < dim ? s->u[L_TP][field - HG_F_VELOCITY_PREV_X] : nullptr;
  if (field == HG_F_PRESSURE) return s->p[layer];
  if (field == HG_F_PRESSURE_PREV) return s->p[L_TP];
  if (field == HG_F_VOLUME_FLUX) { *n = s->nf; return s->F[layer]; }
  if (field == HG_F_VOLUME_FLUX_PREV) { *n = s->nf; return s->F[L_TP]; }
  if (field >= HG_F_PARTIAL_DENSITY_0 && field <= HG_F_PARTIAL_DENSITY_2) return field - HG_F_PARTIAL_DENSITY_0 < np ? s->pd[field - HG_F_PARTIAL_DENSITY_0][layer] : nullptr;
  if (field == HG_F_TEMPERATURE) return s->cfg.heat_enable ? s->T[layer] : nullptr;
  if (field == HG_F_DENSITY) return s->rho;
  if (field == HG_F_VISCOSITY) return s->mu;
  if (field == HG_F_CONDUCTIVITY) return s->kc;
  if (field >= HG_F_FORCE_X && field <= HG_F_FORCE_Z) return field - HG_F_FORCE_X < dim ? s->force[field - HG_F_FORCE_X] : nullptr;
  if (field >= HG_F_STFORCE_X && field <= HG_F_STFORCE_Z) return field - HG_F_STFORCE_X < dim ? s->stforce[field - HG_F_STFORCE_X] : nullptr;
  if (field >= HG_F_VOLUME_FRACTION_0 && field <= HG_F_VOLUME_FRACTION_2) return field - HG_F_VOLUME_FRACTION_0 < np ? s->vf[field - HG_F_VOLUME_FRACTION_0] : nullptr;
  return nullptr;
}

extern "C" int hg_set_field(hg_handle s, int field, const double* src, size_t n) {
  if (!s || !src) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  long long m; double* p = field_ptr(s, field, L_TC, &m);
  if (!p || (long long)n != m || field == HG_F_EXCLUDED) { s->err = "hg_set_field: bad field id or size"; return HG_ERR_INVALID; }
  CK(cudaMemcpyAsync(p, src, n * sizeof(double), cudaMemcpyHostToDevice, s->st));
  if (field <= HG_F_TEMPERATURE) {
    double* q = field_ptr(s, field, L_TP, &m);
    CK(cudaMemcpyAsync(q, src, n * sizeof(double), cudaMemcpyHostToDevice, s->st));
  }
  CK(cudaStreamSynchronize(s->st));
  return 0;
}

// Asynchronous transfers for callers that keep their buffers in pinned memory and overlap the copies with the
// time step (the module-owned property fields go in before a step, results come out after it, hydro2d.hpp:449-463).
// set: the upload runs on its own stream into a staging buffer; the compute stream copies it into the field (and its
// previous-time layer, like hg_set_field) once it has arrived, so a step that is still running is never disturbed.
// get: the compute stream takes a snapshot of the field when it reaches this point; the download of the snapshot runs
// on its own stream, concurrently with the next step.  hg_device_synchronize waits for all of it.
static int xfer_setup(hg_state* s, int field, long long m, std::map<int, double*>& bufs, double** out) {
  if (!s->st_h2d) {
    CK(cudaStreamCreateWithFlags(&s->st_h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s->st_d2h, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&s->xfer_ev_tmp[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->xfer_ev_tmp[1], cudaEventDisableTiming));
  }
  auto it = bufs.find(field);
  if (it == bufs.end()) {
    double* q = nullptr;
    if (int rc = dalloc(s, &q, m, false)) return rc;
    it = bufs.emplace(field, q).first;
  }
  *out = it->second;
  return 0;
}
extern "C" int hg_set_field_async(hg_handle s, int field, const double* src, size_t n) {
  if (!s || !src) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  long long m; double* p = field_ptr(s, field, L_TC, &m);
  if (!p || (long long)n != m || field == HG_F_EXCLUDED) { s->err = "hg_set_field_async: bad field id or size"; return HG_ERR_INVALID; }
  double* stage = nullptr;
  if (int rc = xfer_setup(s, field, m, s->xfer_stage, &stage)) return rc;
  // the staging buffer is free once the compute stream has applied its previous contents (an event recorded right after that
  // copy, NOT the end of everything queued since: the upload of the next step's fields overlaps the running step)
  auto ev = s->xfer_stage_ev.find(field);
  if (ev == s->xfer_stage_ev.end()) {
    cudaEvent_t e = nullptr; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ev = s->xfer_stage_ev.emplace(field, e).first;
  } else {
    CK(cudaStreamWaitEvent(s->st_h2d, ev->second, 0));
  }
  CK(cudaMemcpyAsync(stage, src, n * sizeof(double), cudaMemcpyHostToDevice, s->st_h2d));
  // (an event can be re-recorded once the wait on its previous record has been enqueued)
  CK(cudaEventRecord(s->xfer_ev_tmp[1], s->st_h2d)); CK(cudaStreamWaitEvent(s->st, s->xfer_ev_tmp[1], 0));
  CK(cudaMemcpyAsync(p, stage, n * sizeof(double), cudaMemcpyDeviceToDevice, s->st));
  if (field <= HG_F_TEMPERATURE) {
    double* q = field_ptr(s, field, L_TP, &m);
    CK(cudaMemcpyAsync(q, stage, n * sizeof(double), cudaMemcpyDeviceToDevice, s->st));
  }
  CK(cudaEventRecord(ev->second, s->st));
  return 0;
}
extern "C" int hg_get_field_async(hg_handle s, int field, double* dst, size_t n) {
  if (!s || !dst) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  long long m; double* p = field_ptr(s, field, L_TC, &m);
  if (!p || (long long)n != m || field == HG_F_EXCLUDED) { s->err = "hg_get_field_async: bad field id or size"; return HG_ERR_INVALID; }
  double* snap = nullptr;
  if (int rc = xfer_setup(s, field, m, s->xfer_snap, &snap)) return rc;
  auto it = s->xfer_ev.find(field);
  if (it == s->xfer_ev.end()) {
    cudaEvent_t e = nullptr; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    it = s->xfer_ev.emplace(field, e).first;
  } else {
    CK(cudaStreamWaitEvent(s->st, it->second, 0));   // the previous download of this snapshot has finished
  }
  CK(cudaMemcpyAsync(snap, p, n * sizeof(double), cudaMemcpyDeviceToDevice, s->st));
  CK(cudaEventRecord(s->xfer_ev_tmp[0], s->st)); CK(cudaStreamWaitEvent(s->st_d2h, s->xfer_ev_tmp[0], 0));
  CK(cudaMemcpyAsync(dst, snap, n * sizeof(double), cudaMemcpyDeviceToHost, s->st_d2h));
  CK(cudaEventRecord(it->second, s->st_d2h));
  return 0;
}

extern "C" int hg_get_field(hg_handle s, int field, double* dst, size_t n) {
  if (!s || !dst) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  if (field == HG_F_EXCLUDED) {
    if ((long long)n != s->nc) { s->err = "hg_get_field: bad size"; return HG_ERR_INVALID; }
    LAUNCH(s, k_mask_to_double, nblk(s->nc), 256, s->any_excl ? s->excl : (const unsigned char*)nullptr, s->w1, s->nc);
    CK(cudaMemcpyAsync(dst, s->w1, n * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CK(cudaStreamSynchronize(s->st));
    return 0;
  }
  long long m; double* p = field_ptr(s, field, L_TC, &m);
  if (!p || (long long)n != m) { s->err = "hg_get_field: bad field id or size"; return HG_ERR_INVALID; }
  CK(cudaMemcpyAsync(dst, p, n * sizeof(double), cudaMemcpyDeviceToHost, s->st));
  CK(cudaStreamSynchronize(s->st));
  return 0;
}

// ------------------------------------------------------------------ kernel-level entries
extern "C" int hg_interp_grad(hg_handle s, const double* u, int cond, int comp, double* gx, double* gy, double* gz) {
  if (!s || !u || !gx || !gy) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  CK(cudaMemcpyAsync(s->w1, u, s->nc * sizeof(double), cudaMemcpyHostToDevice, s->st));
  P3 out; for (int d = 0; d < 3; ++d) out.p[d] = s->G[d];
  const unsigned gb = nblk(s->nc);
  if (s->dim == 3) {
    if (cond == 0) k_interp_grad<3, K_NEUMANN0><<<gb, 256, 0, s->st>>>(s->geo, s->w1, comp, out);
    else if (cond == 1) k_interp_grad<3, K_EXTRAP><<<gb, 256, 0, s->st>>>(s->geo, s->w1, comp, out);
    else k_interp_grad<3, K_VEL><<<gb, 256, 0, s->st>>>(s->geo, s->w1, comp, out);
  } else {
    if (cond == 0) k_interp_grad<2, K_NEUMANN0><<<gb, 256, 0, s->st>>>(s->geo, s->w1, comp, out);
    else if (cond == 1) k_interp_grad<2, K_EXTRAP><<<gb, 256, 0, s->st>>>(s->geo, s->w1, comp, out);
    else k_interp_grad<2, K_VEL><<<gb, 256, 0, s->st>>>(s->geo, s->w1, comp, out);
  }
  ++s->launches;
  double* dst[3] = {gx, gy, gz};
  for (int d = 0; d < s->dim; ++d) if (dst[d]) CK(cudaMemcpyAsync(dst[d], s->G[d], s->nc * sizeof(double), cudaMemcpyDeviceToHost, s->st));
  CK(cudaStreamSynchronize(s->st));
  return 0;
}

extern "C" int hg_smooth_field(hg_handle s, const double* u, int repeat, double* out) {
  if (!s || !u || !out) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  CK(cudaMemcpyAsync(s->w1, u, s->nc * sizeof(double), cudaMemcpyHostToDevice, s->st));
  if (int rc = smooth_field(s, s->w1, repeat, s->pc)) return rc;
  CK(cudaMemcpyAsync(out, s->pc, s->nc * sizeof(double), cudaMemcpyDeviceToHost, s->st));
  CK(cudaStreamSynchronize(s->st));
  return 0;
}

extern "C" int hg_linear_solve(hg_handle s, int solver, const double* const coeffs[7], const double* rhs, double* x,
                               double tol, int limit, double relax, int* out_iters, double* out_diff) {
  if (!s || !coeffs || !rhs || !x) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  const unsigned gb = nblk(s->nc);
  // rows and constants: host natural layout -> device sheared layout
  for (int t = 0; t < 7; ++t) {
    if (s->dim == 2 && (t == CZM || t == CZP)) continue;
    if (!coeffs[t]) { s->err = "hg_linear_solve: missing coefficient array"; return HG_ERR_INVALID; }
    CK(cudaMemcpyAsync(s->w1, coeffs[t], s->nc * sizeof(double), cudaMemcpyHostToDevice, s->st));
    DIMSEL(s, k_to_sheared, gb, 256, s->geo, s->w1, s->A[t]);
  }
  CK(cudaMemcpyAsync(s->w1, rhs, s->nc * sizeof(double), cudaMemcpyHostToDevice, s->st));
  DIMSEL(s, k_to_sheared, gb, 256, s->geo, s->w1, s->R[0]);
  int it = 0; double df = 0.;
  if (solver == HG_LS_LU) {
    if (int rc = solve_lu(s, 1)) return rc;
    if (s->lu_tiled) if (int rc = lt_check(s)) return rc;
    DIMSEL(s, k_from_sheared, gb, 256, s->geo, s->X[0], s->pc);
  } else if (solver == HG_LS_GAUSS_SEIDEL) {
    auto launch = [&](int sb, int se) -> int {
      SorArgs a; for (int t = 0; t < 7; ++t) a.A[t] = s->A[t];
      a.R = s->R[0]; a.X = s->PP; a.diff = s->diffs; a.s_begin = sb; a.s_end = se; a.omega = relax; a.tt = s->tt;
      if (s->dim == 3) return coop_launch(s, k_sor_matrix_persistent<3>, s->grid_solver, s->geo, a);
      return coop_launch(s, k_sor_matrix_persistent<2>, s->grid_solver, s->geo, a);
    };
    const int save = s->cfg.pressure_sweeps_per_check;
    if (int rc = run_sor(s, s->PP, s->nsh, tol, limit, launch, &it, &df)) { s->cfg.pressure_sweeps_per_check = save; return rc; }
    DIMSEL(s, k_from_sheared, gb, 256, s->geo, s->PP, s->pc);
  } else if (solver == HG_LS_LU_RELAXED) {
    if (int rc = run_lu_relaxed(s, s->R[0], s->PP, s->X[0], s->X[1], tol, limit, relax, &it, &df)) return rc;
    DIMSEL(s, k_from_sheared, gb, 256, s->geo, s->PP, s->pc);
  } else if (solver == HG_LS_JACOBI) {
    // natural-layout rows in G[0..6] (dim*dim >= 4; use A-sized scratch via from_sheared)
    double* nat[7];
    double* pool[7] = {s->gp[0], s->gp[1], s->fcr[0], s->fcr[1], s->fs[0], s->fs[1], s->G[0]};
    for (int t = 0; t < 7; ++t) {
      nat[t] = pool[t];
      if (s->dim == 2 && (t == CZM || t == CZP)) continue;
      DIMSEL(s, k_from_sheared, gb, 256, s->geo, s->A[t], nat[t]);
    }
    DIMSEL(s, k_from_sheared, gb, 256, s->geo, s->R[0], s->w1);
    if (int rc = run_jacobi(s, nat, nullptr, s->w1, tol, limit, relax, &it, &df)) return rc;
  } else {
    s->err = "hg_linear_solve: solver not on the GPU path";
    return HG_ERR_INVALID;
  }
  CK(cudaMemcpyAsync(x, s->pc, s->nc * sizeof(double), cudaMemcpyDeviceToHost, s->st));
  CK(cudaStreamSynchronize(s->st));
  if (out_iters) *out_iters = it;
  if (out_diff) *out_diff = df;
  return 0;
}

extern "C" int hg_last_residuals(hg_handle s, double* out, int cap, int* n) {
  if (!s || !out || !n) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  int m = s->iter_count < cap ? s->iter_count : cap;
  if (m > 4096) m = 4096;
  if (m > 0) { CK(cudaMemcpyAsync(out, s->resid, m * sizeof(double), cudaMemcpyDeviceToHost, s->st)); CK(cudaStreamSynchronize(s->st)); }
  *n = m;
  return 0;
}

// ---- instrumentation: CUDA events on the handle's own stream (torch.cuda.Event only sees torch's stream)
extern "C" int hg_event_record(hg_handle s, int slot) {
  if (!s || slot < 0 || slot >= 8) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  if (!s->user_ev[slot]) CK(cudaEventCreate(&s->user_ev[slot]));
  CK(cudaEventRecord(s->user_ev[slot], s->st));
  return 0;
}
extern "C" int hg_event_elapsed_ms(hg_handle s, int slot_a, int slot_b, double* ms) {
  if (!s || !ms || slot_a < 0 || slot_b < 0 || slot_a >= 8 || slot_b >= 8 || !s->user_ev[slot_a] || !s->user_ev[slot_b]) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  CK(cudaEventSynchronize(s->user_ev[slot_b]));
  float f = 0.f; CK(cudaEventElapsedTime(&f, s->user_ev[slot_a], s->user_ev[slot_b]));
  *ms = f;
  return 0;
}
extern "C" int hg_profile_enable(hg_handle s, int enable) {
  if (!s) return HG_ERR_INVALID;
  s->profile_on = enable != 0;
  return 0;
}
extern "C" int hg_profile_read(hg_handle s, int which, int* count, double* total_ms) {
  if (!s || which < 0 || which > 1 || !count || !total_ms) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  CK(cudaStreamSynchronize(s->st));
  double tot = 0.; int n = 0;
  for (auto& pr : s->prof_ev[which]) {
    float f = 0.f; cudaEventElapsedTime(&f, pr.first, pr.second); tot += f; ++n;
    cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
  }
  s->prof_ev[which].clear();
  *count = n; *total_ms = tot;
  return 0;
}

extern "C" int hg_profile_read_clocks(hg_handle s, unsigned long long out[16]) {
  if (!s || !out) return HG_ERR_INVALID;
  cudaSetDevice(s->dev);
  for (int q = 0; q < 16; ++q) out[q] = 0;
  if (!s->gt_clk) return 0;
  CK(cudaMemcpyAsync(out, s->gt_clk, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->st));
  CK(cudaMemsetAsync(s->gt_clk, 0, 16 * sizeof(unsigned long long), s->st));
  CK(cudaStreamSynchronize(s->st));
  return 0;
}

extern "C" int hg_timers_enable(hg_handle s, int enable) {
  if (!s) return HG_ERR_INVALID;
  s->timers_on = enable != 0;
  return 0;
}
extern "C" int hg_timers(hg_handle s, char (*names)[64], double* seconds, int cap, int* n) {
  if (!s || !n) return HG_ERR_INVALID;
  int k = 0;
  for (auto& kv : s->timers) {
    if (k >= cap) break;
    if (names) { strncpy(names[k], kv.first.c_str(), 63); names[k][63] = 0; }
    if (seconds) seconds[k] = kv.second.total;
    ++k;
  }
  *n = k;
  return 0;
}
