// hg_slab.cuh -- z-slab decomposition over the GPUs of one node: peer-memory halo exchange, mailbox
// all-gather for the global reductions and the per-step neighbour handshake of the pipelined sweeps.
//
// The reference has no decomposition (its MPI is a stub, source/main.cpp:26-28).  One process per GPU;
// torch.distributed (NCCL) is only the plumbing that carries the CUDA IPC handles.  On the data path every
// rank maps four kinds of peer buffers once (hg_ipc_export / hg_ipc_import):
//   xbuf   exchange staging: a rank PACKS its boundary planes locally, raises a flag in the neighbours'
//          memory, the neighbours PULL the planes over NVLink into their halo planes;
//   mail   mailbox + flag words: every rank pushes its partial scalars (max-norms, CalcStat sums, NaN flags)
//          into every rank's mailbox; the reduction is then done locally in rank order (deterministic and
//          identical on all ranks, so the control flow -- iteration and sweep counts -- stays in lockstep);
//   ll     interface planes of the ordered sweeps (Gauss-Seidel, lu): the thread that updates an interface cell
//          stores {low word, tag, high word, tag} (16 bytes, each half written atomically) into the neighbour's
//          plane; the thread that needs the value -- the z- dependency of the upper slab in the same sweep, the z+
//          dependency of the lower slab in the next one -- polls until both tags carry the sweep it expects.  The
//          data is its own flag: no fence, no handshake, one NVLink latency per hyperplane step (hg_solvers.cuh).
// Staging and mailbox are double-buffered by sequence parity: a rank can never be two exchanges ahead of a
// neighbour, because every exchange needs that neighbour's flag.
#pragma once
#include "hg_device.cuh"

constexpr int SLAB_MAX_ARRAYS = 16;
constexpr int SLAB_MAIL = 1056;      // doubles per rank per mailbox round (per-sweep norms of a 1024-sweep chunk fit)
constexpr int SLAB_MAX_WORLD = 56;
// flag words (unsigned long long) stored after the mailbox doubles
enum { SF_X_LO = 0, SF_X_HI = 1, SF_S_LO = 2, SF_S_HI = 3, SF_ERR = 6, SF_MAIL0 = 8 };

struct Slab {
  int world = 1, rank = 0, has_lo = 0, has_hi = 0;
  int nz_lo = 0, np_glob = 0;
  double* xbuf = nullptr;
  double* mail = nullptr;
  const double* xbuf_lo = nullptr;
  const double* xbuf_hi = nullptr;
  double* mail_peer[SLAB_MAX_WORLD + 8] = {};
  // tagged interface planes: [0] Gauss-Seidel from below, [1] from above, [2+n] lu component n from below,
  // [5+n] from above, [8],[9] checkpoint of [0],[1]; nxy entries each
  uint4* ll = nullptr;          // mine (neighbours write [0], [2..4] / [1], [5..7])
  uint4* ll_lower = nullptr;    // the lower neighbour's block (I write its planes "from above")
  uint4* ll_upper = nullptr;    // the upper neighbour's block (I write its planes "from below")
  unsigned solve_seq = 0, lu_seq = 0;
  unsigned long long xseq = 0, mseq = 0;
  bool linked = false;
};

HD unsigned long long* slab_flags(double* mail, int world) {
  return reinterpret_cast<unsigned long long*>(mail + 2LL * world * SLAB_MAIL);
}

DV unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
DV void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
// Bounded wait for a peer's flag (about 4 s): a rank that died or a protocol bug must not hang the device.  On
// time-out the local error word is raised (the host reports it at the next reduction) and the caller goes on.
DV void slab_wait(const unsigned long long* flag, unsigned long long v, unsigned long long* err) {
  long long t0 = 0;
  for (unsigned spins = 0; ld_acquire_sys(flag) < v; ++spins) {
    if ((spins & 0x3ff) == 0x3ff) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 8000000000LL || *(volatile unsigned long long*)err) { *(volatile unsigned long long*)err = 1ull; return; }
    }
  }
}

struct PackArgs { const double* src[SLAB_MAX_ARRAYS]; int n, planes; };
struct UnpackArgs { double* dst[SLAB_MAX_ARRAYS]; int n, planes; };

HD long long xbuf_index(long long nxy, int parity, int dir, int arr, int plane, long long c2) {
  return (((long long)(parity * 2 + dir) * SLAB_MAX_ARRAYS + arr) * HG_HALO + plane) * nxy + c2;
}

// boundary planes -> local staging: dir 0 = for the lower neighbour (my bottom planes), dir 1 = for the upper
__global__ void k_slab_pack(Geo g, PackArgs a, double* __restrict__ xbuf, int parity, int has_lo, int has_hi) {
  const long long nxy = (long long)g.n[0] * g.n[1];
  const long long per = (long long)a.planes * nxy;
  const long long total = per * a.n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int arr = (int)(t / per);
    const long long r = t % per;
    const int pl = (int)(r / nxy);
    const long long c2 = r % nxy;
    if (has_lo) xbuf[xbuf_index(nxy, parity, 0, arr, pl, c2)] = a.src[arr][(long long)pl * nxy + c2];
    if (has_hi) xbuf[xbuf_index(nxy, parity, 1, arr, pl, c2)] = a.src[arr][(long long)(g.n[2] - a.planes + pl) * nxy + c2];
  }
}

// raise "staging of exchange v is complete" in the neighbours' memory
__global__ void k_slab_signal(unsigned long long* flag_in_lower, unsigned long long* flag_in_upper, unsigned long long v) {
  __threadfence_system();
  if (flag_in_lower) st_release_sys(flag_in_lower, v);
  if (flag_in_upper) st_release_sys(flag_in_upper, v);
}

// wait for the neighbours' flags (one thread: a spinning grid could starve a peer that shares the device) ...
__global__ void k_slab_wait(int has_lo, int has_hi, unsigned long long* myflags, unsigned long long v) {
  if (has_lo) slab_wait(&myflags[SF_X_LO], v, &myflags[SF_ERR]);
  if (has_hi) slab_wait(&myflags[SF_X_HI], v, &myflags[SF_ERR]);
}
// ... then pull their staged planes into my halo planes
__global__ void k_slab_unpack(Geo g, UnpackArgs a, const double* xbuf_lo, const double* xbuf_hi, int parity) {
  const long long nxy = (long long)g.n[0] * g.n[1];
  const long long per = (long long)a.planes * nxy;
  const long long total = per * a.n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int arr = (int)(t / per);
    const long long r = t % per;
    const int pl = (int)(r / nxy);
    const long long c2 = r % nxy;
    // lower neighbour staged its TOP planes in dir 1 -> my planes -planes .. -1
    if (xbuf_lo) a.dst[arr][(long long)(pl - a.planes) * nxy + c2] = __ldcv(&xbuf_lo[xbuf_index(nxy, parity, 1, arr, pl, c2)]);
    // upper neighbour staged its BOTTOM planes in dir 0 -> my planes n[2] .. n[2]+planes-1
    if (xbuf_hi) a.dst[arr][(long long)(g.n[2] + pl) * nxy + c2] = __ldcv(&xbuf_hi[xbuf_index(nxy, parity, 0, arr, pl, c2)]);
  }
}

// mailbox all-gather: push n doubles into every rank's mailbox slot [parity][me], then raise flag[me] there
struct MailPeers { double* mail[SLAB_MAX_WORLD + 8]; };
__global__ void k_mail_post(MailPeers p, const double* __restrict__ vals, int n, int parity, int me, int world, unsigned long long v) {
  const int t = threadIdx.x;
  for (int q = t; q < n; q += blockDim.x) {
    const double x = vals[q];
    for (int r = 0; r < world; ++r) p.mail[r][((long long)parity * world + me) * SLAB_MAIL + q] = x;
  }
  __threadfence_system();
  __syncthreads();
  if (t < world) st_release_sys(slab_flags(p.mail[t], world) + SF_MAIL0 + me, v);
}
__global__ void k_mail_wait(double* mymail, int world, unsigned long long v) {
  const int t = threadIdx.x;
  if (t < world) slab_wait(slab_flags(mymail, world) + SF_MAIL0 + t, v, slab_flags(mymail, world) + SF_ERR);
}

// planes 0..9 as listed at Slab::ll; then the box-dataflow sweeps (k_gs_tiled<LINK>): SLAB_GB "up" planes (top-plane values of
// the lower neighbour, one per sweep of a group), 2 x SLAB_GB "down" planes (the upper neighbour's bottom planes after the last
// sweep of a group, double-buffered by group parity) and 2 x SLAB_GB planes for the checkpoint of the "down" planes
constexpr int SLAB_GB = 8;           // largest sweep group
constexpr int SLAB_LL_UP = 10, SLAB_LL_DOWN = SLAB_LL_UP + SLAB_GB, SLAB_LL_DOWN_SAVE = SLAB_LL_DOWN + 2 * SLAB_GB;
constexpr int SLAB_LL_PLANES = SLAB_LL_DOWN_SAVE + 2 * SLAB_GB;
// tagged 16-byte transport of one double (the scheme of NCCL's LL protocol): each 8-byte half is stored atomically
// and carries the tag, so a reader that sees the expected tag in both halves has the whole value
DV void ll_store(uint4* p, double v, unsigned tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"((unsigned)b), "r"(tag), "r"((unsigned)(b >> 32)), "r"(tag) : "memory");
}
DV double ll_wait(const uint4* p, unsigned tag, unsigned long long* err) {
  unsigned x, t0_, y, t1_;
  long long t0 = 0;
  for (unsigned spins = 0;; ++spins) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(t0_), "=r"(y), "=r"(t1_) : "l"(p) : "memory");
    if (t0_ == tag && t1_ == tag) break;
    if ((spins & 0x3ff) == 0x3ff) {   // bounded (about 4 s): see slab_wait
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 8000000000LL || *(volatile unsigned long long*)err) { *(volatile unsigned long long*)err = 1ull; break; }
    }
  }
  return __longlong_as_double((long long)(((unsigned long long)y << 32) | x));
}
// one (non-blocking) look at an entry: issue several, then check them -- a batch costs one memory latency
DV uint4 ll_load(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
DV bool ll_ok(const uint4& v, unsigned tag) { return v.y == tag && v.w == tag; }
DV double ll_value(const uint4& v) { return __longlong_as_double((long long)(((unsigned long long)v.z << 32) | v.x)); }
__global__ void k_ll_fill(uint4* p, long long n, unsigned tag) {   // value 0 with the given tag
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) p[t] = make_uint4(0u, tag, 0u, tag);
}
// what the ordered-sweep kernels need to talk to the neighbouring slabs
struct SlabLink {
  int on;                         // 0 = single GPU
  int has_lo, has_hi, k0, np_glob, nz_lo;
  const uint4* from_lo;           // my plane written by the lower neighbour (values of its top cells)
  const uint4* from_hi;           // my plane written by the upper neighbour (values of its bottom cells)
  uint4* to_lo;                   // the lower neighbour's "from above" plane
  uint4* to_hi;                   // the upper neighbour's "from below" plane
  unsigned tag0;                  // tag of "before the first sweep of the solve"; sweep s of the solve carries tag0 + s + 1
  // box-dataflow sweeps (k_gs_tiled<LINK>): no value travels downwards inside a sweep group.  A slab also runs the first
  // nsw-1-ds planes of the slab above as ghost planes in sweep ds of a group (the same arithmetic on the same inputs as the
  // owner: identical bits), so its top cell finds its z+ value in its own frames; the owner sends the final values of its
  // bottom planes once per group (down), the lower slab its top-plane value of every sweep (up, one plane per sweep of the
  // group: the owner has consumed a group's values before it produces what the next group's top cells wait for).
  const uint4* up_from;           // mine [sweep of the group][nxy], written by the lower neighbour
  uint4* up_to;                   // the upper neighbour's up_from
  const uint4* down_from;         // mine [group parity][plane][nxy], written by the upper neighbour
  uint4* down_to;                 // the lower neighbour's down_from
  int gbase;                      // groups of this solve launched before this launch (group parity)
  unsigned long long* err;        // local error word (wait timed out)
};

// z+ face coefficients of the lower halo plane (k = -1): the sweep kernel reads them as the z- coupling of the
// bottom owned plane (c_f of the interface face, fluid.hpp:957-964); written straight into the sheared array
template <int DIM>
__global__ void k_cz_halo(Geo g, const double* __restrict__ dc, double* __restrict__ CZ) {
  const long long nxy = (long long)g.n[0] * g.n[1];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nxy) return;
  const int i = (int)(t % g.n[0]), j = (int)(t / g.n[0]);
  const long long cm = cidx(g, i, j, -1), cp = cidx(g, i, j, 0);
  const bool inner = cell_ok(g, i, j, -1) && cell_ok(g, i, j, 0);
  double cf = 0.;
  if (inner) {
    const double dfc = dc[cm] * (1. - 0.5) + dc[cp] * 0.5;
    const double coeff = -g.area[2] / (g.h[2] * dfc);
    cf = -coeff;
  }
  CZ[shidx(g, i, j, -1)] = cf;
}
