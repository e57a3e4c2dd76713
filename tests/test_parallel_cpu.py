"""Host-side logic of the z-slab decomposition, on CPU with gloo, world_size 2 and 3:
the slab-pipelined hyperplane schedule with per-step interface messages reproduces the serial
lexicographic Gauss-Seidel of the oracle (linear.hpp:685-715) bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from hydro_b200.parallel import Slab, active_sweeps, slab_range, slab_sor_reference, step_range  # noqa: E402


def test_slab_ranges_cover_domain():
    for nz in (7, 8, 256):
        for world in (1, 2, 3, 8):
            r = [slab_range(nz, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nz
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_schedule_respects_dependencies():
    """Plane k' of sweep s runs at step k'+2s: after plane k'-1 of sweep s and plane k'+1 of sweep s-1."""
    slab = Slab(6, 5, 4, 1, 0)
    when = {}
    for T in step_range(slab, 4):
        for s in active_sweeps(slab, T, 4):
            when[(T - 2 * s, s)] = T
    for (kp, s), T in when.items():
        if kp > 0:
            assert when[(kp - 1, s)] < T
        if s > 0 and kp + 1 < slab.planes():
            assert when[(kp + 1, s - 1)] < T
    assert len(when) == slab.planes() * 4


def make_system(nx, ny, nz, seed):
    rng = np.random.default_rng(seed)
    rows = [-rng.random((nz, ny, nx)) for _ in range(7)]
    rows[3] = 6.5 + rng.random((nz, ny, nx))
    return rows, rng.standard_normal((nz, ny, nx))


def serial_oracle(nx, ny, nz, rows, rhs, nsweeps, omega):
    import cases
    from oracle_api import Oracle
    o = Oracle(cases.rt3d(4, Nx=nx, Ny=ny, Nz=nz))
    x, it, df = o.linear_solve("gauss_seidel", [r.reshape(-1) for r in rows], rhs.reshape(-1), 0.0, nsweeps - 1, omega)
    return x.reshape(nz, ny, nx), df


def _worker(rank, world, port, shape, nsweeps, omega, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny, nz = shape
    rows, rhs = make_system(nx, ny, nz, 11)
    slab = Slab(nx, ny, nz, world, rank)
    sl = slice(slab.k0, slab.k1)
    x, diff = slab_sor_reference(slab, [r[sl] for r in rows], rhs[sl], nsweeps, omega, dist)
    q.put((rank, x, diff))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_slab_pipelined_sor_equals_serial(world):
    shape, nsweeps, omega = (6, 5, 7), 5, 1.4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, nsweeps, omega, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = np.concatenate([r[1] for r in res], axis=0)
    diff = np.max(np.stack([r[2] for r in res]), axis=0)
    nx, ny, nz = shape
    rows, rhs = make_system(nx, ny, nz, 11)
    xs, dfs = serial_oracle(nx, ny, nz, rows, rhs, nsweeps, omega)
    assert np.array_equal(x, xs)
    assert diff[-1] == dfs
