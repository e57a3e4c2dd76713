// hg_lu_tiled.cuh -- the "lu" solver (LuDecomposition::Solve, linear.hpp:533-566): ONE lexicographic forward sweep
//   x_i = (-c_i - sum_{j<i} a_ij x_j) / a_ii                 (linear.hpp:537-548)
// and ONE backward sweep
//   x_i -= (sum_{j>i} a_ij x_j) / a_ii                        (linear.hpp:551-563)
// as a dataflow of column boxes instead of one grid barrier per hyperplane (k_lu_persistent, hg_solvers.cuh).
//
// A sweep only reads the three neighbours that come earlier in its order, so all cells of a hyperplane
// S = i'+j'+k' are independent (primed = counted from the corner where the sweep starts: the backward sweep is the
// forward sweep in mirrored indices).  A CTA owns a box of LT_TX x LT_TY columns and all k; thread (a,b) owns one
// column and walks it one cell per step: at step S it updates k' = S - i' - j'.  Its z-neighbour is its own last
// value (a register), the x- and y-neighbours were computed at step S-1 by the threads (a-1,b) and (a,b-1): they
// are exchanged through a double-buffered shared-memory frame with one barrier (of the workers) per step.  Values of the
// boxes to the left and below arrive through the result array in global memory: a box publishes the number of
// completed steps (release store) every LT_M steps, a dependent box polls it (acquire load) and then loads the halo
// values of the next LT_M steps at once, so the critical path has one L2 round trip per LT_M hyperplanes instead
// of a grid barrier per hyperplane.  Boxes are claimed from a list sorted by their first step, so a box only
// waits for boxes claimed before it.
//
// The arithmetic (term order z, y, x; one division per component) is that of k_lu_persistent: bit-identical.
#pragma once
#ifndef LT_SLEEP_NS
#define LT_SLEEP_NS 200   // poll interval of the publisher warp
#endif
#include "hg_device.cuh"
#include "hg_slab.cuh"

#ifndef LT_M_N
#define LT_M_N 4
#endif
#ifndef LT_TY_N
#define LT_TY_N 8
#endif
#ifndef LT_PF_N
#define LT_PF_N 8
#endif
constexpr int LT_TX = 32, LT_TY = LT_TY_N, LT_M = LT_M_N, LT_PF = LT_PF_N;   // operands are requested LT_PF steps ahead
constexpr int LT_WORK = LT_TX * LT_TY;            // worker threads: one per column of the box
constexpr int LT_THREADS = LT_WORK + 64;          // + halo warp + publisher warp
constexpr int LT_FW = LT_TX + 1, LT_FH = LT_TY + 1, LT_FRAME = LT_FW * LT_FH;   // frame: halo column / row at index 0
constexpr int LT_HALO = LT_TX + LT_TY;            // halo entries per frame: column 0 (rows 1..TY), row 0 (columns 1..TX)
constexpr int LT_PBIAS = 1;
#ifndef LT_CTAS_N
#define LT_CTAS_N 1
#endif
constexpr int LT_CTAS_PER_SM = LT_CTAS_N;
// ring of operand slots in (dynamic) shared memory: [LT_PF][7 arrays][LT_WORK] doubles, filled by cp.async
constexpr int LT_RING_BYTES = LT_PF * 7 * LT_TX * LT_TY_N * 8;
// 8-byte copy global -> shared; bytes = 0: nothing is read (the destination is zero-filled)
DV void lt_cp_async8(unsigned dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(src), "r"(bytes) : "memory");
}
DV void lt_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> DV void lt_cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

struct LtArgs {
  const double* A[7];   // sheared rows
  const double* R[3];   // sheared constants (forward sweep)
  double* X[3];         // sheared result: written by the forward sweep, updated in place by the backward sweep
  int ncomp;
  const int2* boxes;    // (I', J') in claim order
  int nboxes, nbi;      // boxes, boxes along i
  int* progress;        // [nboxes] completed steps + LT_PBIAS, indexed J' * nbi + I'; zeroed before the launch
  int* ctl;             // [0] next box, [1] abort flag
  SlabLink link;        // z-slab decomposition: tagged interface planes of the neighbouring slabs (hg_slab.cuh); component n at + n link_stride
  long long link_stride;
#ifdef LT_TRACE
  unsigned long long* trace;   // diagnostics build: per box (claimed, first macro step started, finished) in ns + SM id
  unsigned long long* trace2;  // [box][4][128]: per macro step: workers started it / published / halo warp's poll succeeded / stage stored
#endif
};
#ifdef LT_TRACE
DV unsigned long long lt_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
DV unsigned lt_smid() { unsigned r; asm volatile("mov.u32 %0, %smid;" : "=r"(r)); return r; }
#endif

DV int lt_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
DV void lt_st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
DV int lt_ld_acquire_cta(const int* p) {   // shared memory
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
DV void lt_st_release_cta(int* p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
DV void lt_bar_work() { asm volatile("bar.sync 1, %0;" :: "n"(LT_WORK) : "memory"); }          // the workers, once per step
DV void lt_bar_macro() { asm volatile("bar.sync 2, %0;" :: "n"(LT_WORK + 32) : "memory"); }    // workers + halo warp, once per macro step

// DIR = 0: forward sweep, DIR = 1: backward sweep.  LINK: the mesh is one z-slab of a decomposed run -- the first cell of
// a column takes its z-neighbour from the slab the sweep comes from (a value tagged with the solve, written by that
// slab's thread when it finished the column: the data is its own flag), the last cell hands its value on.  The slabs
// run the same kernel concurrently; a slab's wavefront simply starts when the first columns of its neighbour are done.
//
// Roles (round 2): LT_WORK worker threads walk the columns; the HALO warp polls the neighbouring boxes' progress and loads
// the halo values of the NEXT macro step (LT_M steps) into a double-buffered stage while the workers run the current
// one; the PUBLISHER warp forwards the number of completed steps to global memory (the release at GPU scope costs a few
// thousand cycles: round 1 paid it, a poll and a load round trip inside every macro step of every box).
template <int DIR, bool LINK = false>
__global__ void __launch_bounds__(LT_THREADS, LT_CTAS_PER_SM) k_lu_tiled(Geo g, LtArgs a) {
  extern __shared__ __align__(16) double lt_ring[];  // operand ring (see LT_RING_BYTES)
  __shared__ double fr[2][3][LT_FRAME];            // values of the last two steps
  __shared__ double stage[2][LT_M][3][LT_HALO];    // halo values of the LT_M steps of a macro step, by macro-step parity
  __shared__ int s_box, s_prog;
  const int tid = threadIdx.x, ta = tid & (LT_TX - 1), tb = tid / LT_TX;
  const bool worker = tid < LT_WORK, halo_warp = tid >= LT_WORK && tid < LT_WORK + 32;
  const int nx = g.n[0], ny = g.n[1], nz = g.n[2];
  const long long PS = (long long)nx * ny;
  const long long sgn = DIR ? -1 : 1;              // a step moves the thread by one hyperplane: +PS / -PS entries
  const long long nbx = sgn * (PS + 1), nby = sgn * (PS + nx), nbz = sgn * PS;   // cs - nb* = the earlier neighbour
  // coefficient slots: diagonal + the three earlier neighbours
  const double* const Az = a.A[DIR ? CZP : CZM];
  const double* const Ay = a.A[DIR ? CYP : CYM];
  const double* const Ax = a.A[DIR ? CXP : CXM];
  const double* const Ad = a.A[CD];
  // components beyond ncomp alias component 0 (loaded, never stored): no conditional loads in the step loop
  const double* src[3]; double* dstx[3];
#pragma unroll
  for (int n = 0; n < 3; ++n) { const int q = n < a.ncomp ? n : 0; src[n] = DIR ? a.X[q] : a.R[q]; dstx[n] = a.X[q]; }
  for (;;) {
    __syncthreads();
    if (tid == 0) s_box = atomicAdd(&a.ctl[0], 1);
    __syncthreads();
    const int t = s_box;
    if (t >= a.nboxes || *(volatile int*)&a.ctl[1]) return;
    const int2 bx = a.boxes[t];
    const int I0 = bx.x * LT_TX, J0 = bx.y * LT_TY;
    const int Slo = I0 + J0;
    const int Shi = min(I0 + LT_TX, nx) - 1 + min(J0 + LT_TY, ny) - 1 + nz - 1;
    const int nmacro = (Shi - Slo + LT_M) / LT_M;
    const int me = bx.y * a.nbi + bx.x;
    const int dep_x = bx.x > 0 ? me - 1 : -1, dep_y = bx.y > 0 ? me - a.nbi : -1;
#ifdef LT_TRACE
    if (tid == 0 && a.trace) { a.trace[me * 4] = lt_now(); a.trace[me * 4 + 3] = lt_smid(); }
#endif
    for (int q = tid; q < 2 * 3 * LT_FRAME; q += LT_THREADS) (&fr[0][0][0])[q] = 0.;
    for (int q = tid; q < 2 * LT_M * 3 * LT_HALO; q += LT_THREADS) (&stage[0][0][0][0])[q] = 0.;
    if (tid == 0) s_prog = 0;
    __syncthreads();
    if (!worker && !halo_warp) {
      // ------------------------------------------------------------ publisher warp
      if ((tid & 31) == 0) {
        int last = 0;
        for (;;) {
          const int v = lt_ld_acquire_cta(&s_prog);
          if (v != last) { lt_st_release(&a.progress[me], v); last = v; if (v == 0x7fffffff) break;
#ifdef LT_TRACE
                           { const int mm = (v - LT_PBIAS - Slo) / LT_M; if (a.trace2 && mm >= 0 && mm < 128) a.trace2[((long long)me * 4 + 1) * 128 + mm] = lt_now(); }
#endif
          }
          else { if (LT_SLEEP_NS > 0) __nanosleep(LT_SLEEP_NS); }
        }
      }
      __syncwarp();
      continue;
    }
    if (halo_warp) {
      // ------------------------------------------------------------ halo warp: the stage of macro step m is complete
      // before the macro barrier #m; afterwards the stage of m+1 is loaded while the workers run m
      const int lane = tid & 31;
      constexpr int NE = (LT_M * LT_HALO + 31) / 32;   // (step, entry) pairs per lane
      for (int m = 0; m < nmacro; ++m) {
        const int S0 = Slo + m * LT_M;
        if (dep_x >= 0 || dep_y >= 0) {
          if (lane < 2) {
            const int dep = lane == 0 ? dep_x : dep_y;
            if (dep >= 0) {
              const int need = S0 + LT_M - 1 + LT_PBIAS;
              long long t0 = 0;
              for (unsigned spins = 0; lt_ld_acquire(&a.progress[dep]) < need; ++spins) {
                if ((spins & 0xff) == 0xff) {   // bounded wait: a scheduling bug must not hang the device
                  const long long now = clock64();
                  if (t0 == 0) t0 = now;
                  if (now - t0 > 4000000000LL) atomicExch(&a.ctl[1], 1);
                  if (*(volatile int*)&a.ctl[1]) break;
                }
              }
            }
          }
          __syncwarp();
#ifdef LT_TRACE
          if (lane == 0 && a.trace2 && m < 128) a.trace2[((long long)me * 4 + 2) * 128 + m] = lt_now();
#endif
          // entry e of a frame's halo: e < TY: column 0, row e+1 (x-neighbour of thread (0, e)); else row 0, column e-TY+1
          double hvv[NE][3]; bool hok[NE];
#pragma unroll
          for (int r = 0; r < NE; ++r) {
            const int q = lane + 32 * r;
            const int st = q / LT_HALO, e = q - st * LT_HALO;
            const bool isx = e < LT_TY;
            const int qa = isx ? 0 : e - LT_TY, qb = isx ? e : 0;          // the thread whose neighbour this is
            const int qip = I0 + qa, qjp = J0 + qb;
            const int S = S0 + st, kp = S - qip - qjp;
            const bool v = q < LT_M * LT_HALO && (isx ? dep_x >= 0 : dep_y >= 0) && qip < nx && qjp < ny && kp >= 0 && kp < nz;
            const int qi = DIR ? nx - 1 - qip : qip, qj = DIR ? ny - 1 - qjp : qjp;
            long long c = (DIR ? ((long long)(g.np - S) * ny + qj) * nx + qi : ((long long)(S + 1) * ny + qj) * nx + qi) - (isx ? nbx : nby);
            if (!v) c = 0;
            hok[r] = v;
            // (entries without a cell are not loaded: every box asking for the same spare entry would make it a hot spot in L2)
#pragma unroll
            for (int n = 0; n < 3; ++n) hvv[r][n] = v ? __ldcg(&dstx[n][c]) : 0.;
          }
#pragma unroll
          for (int r = 0; r < NE; ++r) {
            const int q = lane + 32 * r;
            if (q < LT_M * LT_HALO) {
              const int st = q / LT_HALO, e = q - st * LT_HALO;
#pragma unroll
              for (int n = 0; n < 3; ++n) stage[m & 1][st][n][e] = hok[r] ? hvv[r][n] : 0.;
            }
          }
#ifdef LT_TRACE
          if (lane == 0 && a.trace2 && m < 128) a.trace2[((long long)me * 4 + 3) * 128 + m] = lt_now();
#endif
        }
        lt_bar_macro();
      }
      lt_bar_macro();   // the workers have finished the last macro step
      continue;
    }
    // -------------------------------------------------------------- workers
    const int ip = I0 + ta, jp = J0 + tb;                      // primed column
    const bool col = ip < nx && jp < ny;
    const int i = DIR ? nx - 1 - ip : ip, j = DIR ? ny - 1 - jp : jp;
    // sheared index of the thread's cell at step S: plane (S + 1) forward, (np - S) backward (np - 1 = largest i+j+k)
    long long cs = DIR ? ((long long)(g.np - Slo) * ny + j) * nx + i : ((long long)(Slo + 1) * ny + j) * nx + i;
    // neighbour existence as the reference tests it (k_lu_persistent)
    const bool hx = DIR ? i + 1 < nx : i > 0, hy = DIR ? j + 1 < ny : j > 0;
    double xz[3] = {0., 0., 0.};                               // own value of the previous step
    // Operands (three coefficients, diagonal, constants of the components) are requested LT_PF steps ahead with cp.async
    // into the thread's own entries of a ring in shared memory: the rows come from DRAM (the sheared arrays are written
    // once and read once), and a step is much shorter than a DRAM round trip -- round 1 kept them in registers 4 steps
    // ahead and every step waited for memory.  Threads without a cell at that step read entry 0 of the arrays (an unused
    // corner entry) with a source size of zero -- nothing is read, the entry is zero-filled, the result is discarded:
    // unconditional copies, issued back to back.  (Reading the corner entry for real made it a hot spot: while a box fills or
    // drains most of its threads have no cell, and every warp of every box asked L2 for the same seven sectors in every step --
    // the first steps of a box ran at half speed, which is what a dependent box waits for.)
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(lt_ring) + tid * 8;
    auto request_ops = [&](int slot, long long c, int S) {
      const int kp = S - ip - jp;
      const bool v = col && kp >= 0 && kp < nz;
      if (!v) c = 0;
      const unsigned nb = v ? 8u : 0u;
      const unsigned d = ring_s + slot * (7 * LT_WORK * 8);
      lt_cp_async8(d, Az + c, nb); lt_cp_async8(d + LT_WORK * 8, Ay + c, nb); lt_cp_async8(d + 2 * LT_WORK * 8, Ax + c, nb);
      lt_cp_async8(d + 3 * LT_WORK * 8, Ad + c, nb);
#pragma unroll
      for (int n = 0; n < 3; ++n) lt_cp_async8(d + (4 + n) * LT_WORK * 8, src[n] + c, nb);
      lt_cp_commit();
    };
#pragma unroll 1
    for (int q = 0; q < LT_PF; ++q) request_ops(q, cs + q * nbz, Slo + q);
    int slot = 0;
    double* const f0 = &fr[0][0][0] + (tb + 1) * LT_FW + ta + 1;   // own slot in frame 0, component 0
    for (int m = 0; m < nmacro; ++m) {
      const int S0 = Slo + m * LT_M;
      lt_bar_macro();   // macro step m-1 is complete; the stage of macro step m is loaded
#ifdef LT_TRACE
      if (tid == 0 && m == 0 && a.trace) a.trace[me * 4 + 1] = lt_now();
      if (tid == 0 && a.trace2 && m < 128) a.trace2[((long long)me * 4 + 0) * 128 + m] = lt_now();
#endif
      if (tid == 0 && m > 0) lt_st_release_cta(&s_prog, S0 + LT_PBIAS);   // steps < S0 are complete (publisher warp)
      // ---- LT_M steps
#pragma unroll
      for (int st = 0; st < LT_M; ++st) {
        const int S = S0 + st, par = S & 1;
        const int kp = S - ip - jp;
        const bool v = col && kp >= 0 && kp < nz;
        const int k = DIR ? nz - 1 - kp : kp;
        const bool zin = LINK && (DIR ? a.link.has_hi : a.link.has_lo) && kp == 0;      // z-neighbour in the other slab
        const bool hz = (DIR ? k + 1 < nz : k > 0) || zin;
        lt_cp_wait<LT_PF - 1>();   // the copies of this step (the oldest group in flight) have landed
        const double* const ops = lt_ring + slot * (7 * LT_WORK) + tid;
        const double cz = ops[0], cy = ops[LT_WORK], cx = ops[2 * LT_WORK], dg = v ? ops[3 * LT_WORK] : 1.;
        const HgDiv ddg = hg_div_prepare(dg);   // one reciprocal for the components
        double rr[3];
#pragma unroll
        for (int n = 0; n < 3; ++n) rr[n] = ops[(4 + n) * LT_WORK];
        request_ops(slot, cs + LT_PF * nbz, S + LT_PF);
        slot = slot + 1 == LT_PF ? 0 : slot + 1;
        // halo of the frame of step S-1 for this step's edge threads
        if (tid < LT_HALO) {
#pragma unroll
          for (int n = 0; n < 3; ++n) {
            const int e = tid;
            double* const dst = &fr[par ^ 1][n][0] + (e < LT_TY ? (e + 1) * LT_FW : e - LT_TY + 1);
            *dst = stage[m & 1][st][n][e];
          }
        }
        lt_bar_work();
        double* const fprev = f0 + (par ^ 1) * 3 * LT_FRAME;
        double* const fcur = f0 + par * 3 * LT_FRAME;
        if (LINK && v && zin) {
          const uint4* const from = (DIR ? a.link.from_hi : a.link.from_lo) + (long long)j * nx + i;
#pragma unroll
          for (int n = 0; n < 3; ++n) if (n < a.ncomp) xz[n] = ll_wait(from + n * a.link_stride, a.link.tag0 + (DIR ? 2u : 1u), a.link.err);
        }
#pragma unroll
        for (int n = 0; n < 3; ++n) {
          double xv = 0.;
          if (n < a.ncomp) {
            const double vx = fprev[n * LT_FRAME - 1], vy = fprev[n * LT_FRAME - LT_FW];
            double sum = 0.;
            if (hz) sum += cz * xz[n];
            if (hy) sum += cy * vy;
            if (hx) sum += cx * vx;
            xv = DIR ? rr[n] - hg_div(sum, ddg) : hg_div(-rr[n] - sum, ddg);
            if (v) a.X[n][cs] = xv; else xv = 0.;
            if (LINK && v && kp == nz - 1 && (DIR ? a.link.has_lo : a.link.has_hi))
              ll_store((DIR ? a.link.to_lo : a.link.to_hi) + n * a.link_stride + (long long)j * nx + i, xv, a.link.tag0 + (DIR ? 2u : 1u));
          }
          fcur[n * LT_FRAME] = xv;
          xz[n] = xv;
        }
        cs += nbz;
      }
    }
    lt_bar_macro();   // all steps done (the halo warp takes part)
#ifdef LT_TRACE
    if (tid == 0 && a.trace) a.trace[me * 4 + 2] = lt_now();
#endif
    if (tid == 0) lt_st_release_cta(&s_prog, 0x7fffffff);
  }
}
