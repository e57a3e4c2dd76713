"""z-slab decomposition of the hot path over the GPUs of one node.

The reference has no decomposition (its MPI is a stub, source/main.cpp:26-28), so this is a new design:

* cells are split into contiguous z-slabs (k is the slowest index, mesh.hpp:552-561, so a slab is a
  contiguous range of every cell array); halo depth 2 planes for u, force and the partial densities
  (radius-2 operators: explicit viscous term fluid.hpp:838-853, Superbee advection solver.hpp:560-619),
  1 plane for p, p', d_c, mu; interface z-faces are computed redundantly by both neighbours;
* stencil/assembly/advection kernels need halo exchanges only (`HALO_PLAN`);
* the order-dependent solvers (lexicographic Gauss-Seidel linear.hpp:685-715, `lu` linear.hpp:533-566) stay
  EXACT: the global hyperplane schedule of hg_solvers.cuh (step T handles plane T-2s of sweep s) is kept, every
  rank processes its part of each plane, and after every step the new values of a slab's top plane go to the
  upper neighbour (its z- dependency, same sweep) and of its bottom plane to the lower neighbour (its z+
  dependency, previous sweep).  `slab_sor_reference` below executes exactly this protocol with
  torch.distributed point-to-point messages and is checked bit-for-bit against the serial solver
  (tests/test_parallel_cpu.py, gloo, world_size 2 and 3).  On the device the same messages are peer-memory
  stores over NVLink followed by a system-scope flag (one handshake per step), see DESIGN.md.

The device path: hydro_b200/csrc/hg_slab.cuh + the slab-aware kernels; `run_local_ranks` drives several ranks from one
process (tests), bench.py one rank per process under torchrun.
"""
from dataclasses import dataclass

import numpy as np


def slab_range(nz, world, rank):
    """Contiguous z-range [k0, k1) owned by `rank`; the first nz % world ranks get one extra plane."""
    base, rem = divmod(nz, world)
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


@dataclass
class Slab:
    nx: int
    ny: int
    nz: int       # global
    world: int
    rank: int

    @property
    def k0(self):
        return slab_range(self.nz, self.world, self.rank)[0]

    @property
    def k1(self):
        return slab_range(self.nz, self.world, self.rank)[1]

    @property
    def nzl(self):
        return self.k1 - self.k0

    @property
    def lower(self):
        return self.rank - 1 if self.rank > 0 else None

    @property
    def upper(self):
        return self.rank + 1 if self.rank + 1 < self.world else None

    def planes(self):
        """Number of global hyperplanes i+j+k = const."""
        return self.nx + self.ny + self.nz - 2

    def local_cells_of_plane(self, kp):
        """(i, j, k_global) arrays of the owned cells with i+j+k = kp, in the raw order of the plane."""
        out = []
        for k in range(self.k0, self.k1):
            for j in range(self.ny):
                i = kp - j - k
                if 0 <= i < self.nx:
                    out.append((i, j, k))
        return out


# Which fields must have fresh halo planes before which kernel (stage order of hg_fluid_make_iteration).
# (field, halo planes, consumer)
HALO_PLAN = [
    ("force[3], mu", 1, "k_pre / k_source / k_assemble (once per time step, after hg_update_properties)"),
    ("p (iter_prev)", 1, "k_pre, k_fstar"),
    ("u (iter_prev)[3]", 1, "k_velgrad, k_assemble"),
    ("G[n][z], G[z][d]", 1, "k_source (faces in z), k_assemble (deferred correction on z faces)"),
    ("u* (iter_curr)[3], gp[3], fcr[3], d_c", 1, "k_fstar, k_prhs, k_correct"),
    ("p'", 1, "k_correct (gradient with extrapolation condition)"),
    ("partial densities", 2, "k_advect (Superbee needs the gradient of the upwind neighbour)"),
    ("smoothed fields", 1, "k_smooth, once per repeat"),
]


def step_range(slab, nsweeps):
    """Global steps of the pipelined schedule: T = 0 .. planes-1 + 2 (S-1)."""
    return range(0, slab.planes() + 2 * (nsweeps - 1))


def active_sweeps(slab, T, nsweeps):
    """Sweeps s whose plane T-2s exists."""
    return [s for s in range(nsweeps) if 0 <= T - 2 * s < slab.planes()]


def slab_sor_reference(slab, rows, rhs, nsweeps, omega, dist=None):
    """Slab-distributed lexicographic SOR, exactly linear.hpp:685-715, by the global hyperplane schedule.

    rows: 7 arrays [nzl, ny, nx] (z-,y-,x-,diag,x+,y+,z+) of the OWNED cells, rhs likewise.
    Returns (x [nzl, ny, nx], per-sweep max|corr| over the owned cells).
    Interface values travel after every step: top plane -> upper rank, bottom plane -> lower rank.
    """
    nx, ny, nzl, k0 = slab.nx, slab.ny, slab.nzl, slab.k0
    x = np.zeros((nzl + 2, ny, nx))          # one halo plane below (index 0) and above (index nzl+1)
    diff = np.zeros(nsweeps)
    for T in step_range(slab, nsweeps):
        sweeps = active_sweeps(slab, T, nsweeps)
        send_up, send_dn = [], []
        for s in sweeps:
            kp = T - 2 * s
            for (i, j, k) in slab.local_cells_of_plane(kp):
                kl = k - k0 + 1
                c = (k - k0, j, i)
                acc = 0.0
                if k > 0:
                    acc += rows[0][c] * x[kl - 1, j, i]
                if j > 0:
                    acc += rows[1][c] * x[kl, j - 1, i]
                if i > 0:
                    acc += rows[2][c] * x[kl, j, i - 1]
                if i + 1 < nx:
                    acc += rows[4][c] * x[kl, j, i + 1]
                if j + 1 < ny:
                    acc += rows[5][c] * x[kl, j + 1, i]
                if k + 1 < slab.nz:
                    acc += rows[6][c] * x[kl + 1, j, i]
                value = -(rhs[c] + acc) / rows[3][c]
                corr = value - x[kl, j, i]
                diff[s] = max(diff[s], abs(corr))
                x[kl, j, i] += corr * omega
                if k == slab.k1 - 1 and slab.upper is not None:
                    send_up.append((j, i, x[kl, j, i]))
                if k == k0 and slab.lower is not None:
                    send_dn.append((j, i, x[kl, j, i]))
        if dist is not None and slab.world > 1:
            _exchange(dist, slab, x, send_up, send_dn)
    return x[1:-1], diff


def _exchange(dist, slab, x, send_up, send_dn):
    """One step's interface messages; both neighbours know the cell lists, so only counts + values travel."""
    import torch

    def pack(lst):
        t = torch.zeros(1 + 3 * len(lst), dtype=torch.float64)
        t[0] = len(lst)
        if lst:
            t[1:] = torch.tensor(lst, dtype=torch.float64).reshape(-1)
        return t

    cap = 1 + 3 * slab.nx * slab.ny
    ops, bufs = [], {}
    for peer, lst in ((slab.upper, send_up), (slab.lower, send_dn)):
        if peer is None:
            continue
        out = torch.zeros(cap, dtype=torch.float64)
        p = pack(lst)
        out[:p.numel()] = p
        bufs[peer] = torch.zeros(cap, dtype=torch.float64)
        ops.append(dist.P2POp(dist.isend, out, peer))
        ops.append(dist.P2POp(dist.irecv, bufs[peer], peer))
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    for peer, buf in bufs.items():
        n = int(buf[0].item())
        vals = buf[1:1 + 3 * n].reshape(n, 3).numpy()
        plane = 0 if peer == slab.lower else slab.nzl + 1   # lower rank's top plane -> my bottom halo, and v.v.
        for j, i, v in vals:
            x[plane, int(j), int(i)] = v


# ---------------------------------------------------------------- driving the device path
def join_cells(parts):
    """Cell field of the whole mesh from the ranks' slabs (a slab is a contiguous range of the raw order)."""
    return np.concatenate(parts)


def join_faces(parts, nx, ny, nz, world):
    """Face field (x-, y-, z-blocks, mesh.hpp:698-705) of the whole mesh from the ranks' local face arrays;
    interface z-faces exist on both neighbours (identical values): the lower rank's copy is dropped."""
    xs, ys, zs = [], [], []
    for r, f in enumerate(parts):
        k0, k1 = slab_range(nz, world, r)
        nzl = k1 - k0
        nxf, nyf = (nx + 1) * ny * nzl, nx * (ny + 1) * nzl
        xs.append(f[:nxf])
        ys.append(f[nxf:nxf + nyf])
        z = f[nxf + nyf:].reshape(nzl + 1, ny * nx)
        zs.append(z if r == world - 1 else z[:-1])
    return np.concatenate(xs + ys + [z.reshape(-1) for z in zs])


def run_local_ranks(params, world, nsteps, fields, devices=None, solver_ctas=0, timeout=120.0):
    """Runs `world` slab ranks in this process, one thread per rank (the C calls release the GIL), and returns
    (stats of the last step of rank 0, {field: whole-mesh array}).  devices: one per rank (default: all on 0,
    which needs solver_ctas <= SMs / world so that the ranks' persistent kernels are co-resident)."""
    import threading
    from .capi import Hydro
    from .config import FACE_FIELDS, F
    devices = devices or [0] * world
    hs, errs, out, stats = [None] * world, [], [None] * world, [None] * world
    bar = threading.Barrier(world)

    def work(r):
        try:
            hs[r] = Hydro(params, device=devices[r], world_size=world, rank=r, solver_ctas=solver_ctas)
        except Exception as e:   # noqa: BLE001
            errs.append((r, e))
        try:
            bar.wait(timeout)
            if errs:
                return
            hs[r].link_local(hs)
            st = None
            for _ in range(nsteps):
                st = hs[r].step()
            stats[r] = st
            out[r] = {n: hs[r].get(n) for n in fields}
        except Exception as e:   # noqa: BLE001
            errs.append((r, e))

    th = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout)
    if errs:
        raise RuntimeError("rank %d: %s" % (errs[0][0], errs[0][1]))
    if any(t.is_alive() for t in th):
        raise RuntimeError("slab ranks did not finish")
    p = params
    nx, ny, nz = int(p["Nx"]), int(p["Ny"]), int(p["Nz"])
    whole = {}
    for n in fields:
        parts = [out[r][n] for r in range(world)]
        whole[n] = join_faces(parts, nx, ny, nz, world) if F[n] in FACE_FIELDS else join_cells(parts)
    for h in hs:
        h.close()
    return stats[0], whole
