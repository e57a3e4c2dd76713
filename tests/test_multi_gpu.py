"""z-slab decomposition on the device: the decomposed run must reproduce the single-GPU run bit for bit
-- every kernel is cell-local given the halo planes, and the ordered solvers keep the global hyperplane schedule.

The ranks live in this process (one thread each, hg_link_local).  With one visible GPU they share it (their
persistent solver kernels are limited to a part of the SMs so that they are co-resident); with more GPUs every
rank gets its own device and the halo planes / solver interface values travel over NVLink peer memory.
"""
import numpy as np
import pytest

import cases
from hydro_b200 import parallel

pytestmark = pytest.mark.gpu
FIELDS = ["VELOCITY_X", "VELOCITY_Y", "VELOCITY_Z", "PRESSURE", "VOLUME_FLUX", "PARTIAL_DENSITY_0", "PARTIAL_DENSITY_1",
          "DENSITY", "VISCOSITY", "FORCE_Y"]


def single(p, nsteps, fields=None):
    from hydro_b200.capi import Hydro
    h = Hydro(p)
    st = [h.step() for _ in range(nsteps)][-1]
    out = {n: h.get(n) for n in (fields or FIELDS)}
    h.close()
    return st, out


def devices_for(world):
    import torch
    n = torch.cuda.device_count()
    return (list(range(world)), 0) if n >= world else ([0] * world, 148 // world - 2)


@pytest.mark.parametrize("kernel", ["hyperplane", "tiled"])
@pytest.mark.parametrize("case,world", [("rt3d_16", 2), ("rt3d_20x12x17", 3), ("dam3d_32x10x12", 2), ("rt3d_tol", 2),
                                        ("rt3d_stf_20x12x16", 2), ("thermal3d_24x12x15", 3), ("rt3d_jacobi_16x12x14", 2)])
def test_slabs_equal_single_gpu(case, world, kernel, monkeypatch):
    # hyperplane: the neighbour-linked hyperplane kernels on both sides; tiled: the box-dataflow kernels (k_gs_tiled,
    # k_lu_tiled) with tagged interface values between the slabs -- all bitwise equal to the single-GPU run
    if kernel == "hyperplane":
        monkeypatch.setenv("HYDRO_GS_KERNEL", "hyperplane")
        monkeypatch.setenv("HYDRO_LU_KERNEL", "hyperplane")
    p = {"rt3d_16": cases.rt3d(16),
         "rt3d_20x12x17": cases.rt3d(8, Nx=20, Ny=12, Nz=17, lu_relaxed_num_iters_limit=30),
         "dam3d_32x10x12": cases.broken_dam_3d(32, 10, 12, lu_relaxed_num_iters_limit=40),
         # surface tension (CalcForce, hydro2d.hpp:1318-1370) and the temperature equation (heat.hpp:19-93) on slabs
         "rt3d_stf_20x12x16": cases.rt3d(8, Nx=20, Ny=12, Nz=16, lu_relaxed_num_iters_limit=20, sigma=0.07, tvd_split=1, sharp=0.03),
         "thermal3d_24x12x15": cases.rt3d(8, Nx=24, Ny=12, Nz=15, lu_relaxed_num_iters_limit=15, heat_enable=1,
                                          heat_box_lb=(-1., -1., -1.), heat_box_rt=(2., 0.01, 2.), heat_box_temperature=1.,
                                          conductivity_0=0.01, conductivity_1=0.05),
         "rt3d_jacobi_16x12x14": cases.rt3d(8, Nx=16, Ny=12, Nz=14, linear_solver_pressure="jacobi", lu_relaxed_relaxation_factor=0.9,
                                            lu_relaxed_num_iters_limit=60, lu_relaxed_tolerance=1e-6, convergence_tolerance=0.0,
                                            num_iterations_limit=3),
         "rt3d_tol": cases.rt3d(12, fixed_work=False, lu_relaxed_num_iters_limit=200, lu_relaxed_tolerance=1e-6,
                               num_iterations_limit=3, pressure_sweeps_per_check=32)}[case]
    fields = FIELDS + (["TEMPERATURE"] if p.get("heat_enable") else []) + (["STFORCE_X", "STFORCE_Y", "STFORCE_Z"] if p.get("sigma") else [])
    st1, f1 = single(p, 2, fields)
    devs, ctas = devices_for(world)
    stn, fn = parallel.run_local_ranks(p, world, 2, fields, devices=devs, solver_ctas=ctas)
    assert stn.simple_iterations == st1.simple_iterations
    assert stn.pressure_sweeps_total == st1.pressure_sweeps_total
    assert stn.pressure_last_diff == st1.pressure_last_diff
    assert stn.convergence_indicator == st1.convergence_indicator
    for n in fields:
        assert np.array_equal(fn[n], f1[n]), n
    for ph in range(2):
        assert abs(stn.volume[ph] - st1.volume[ph]) <= 1e-12 * abs(st1.volume[ph])


def test_single_gpu_matches_tiled_default():
    """The default single-GPU configuration (tile sweeps) against the 2-slab run."""
    p = cases.rt3d(24)
    st1, f1 = single(p, 1)
    devs, ctas = devices_for(2)
    stn, fn = parallel.run_local_ranks(p, 2, 1, FIELDS, devices=devs, solver_ctas=ctas)
    for n in FIELDS:
        assert np.array_equal(fn[n], f1[n]), n


@pytest.mark.parametrize("world", [2, 4])
def test_one_process_per_gpu_over_ipc_equals_single_gpu(world):
    """The path the scaling bench measures: one process per GPU (torchrun), CUDA IPC handles exchanged through
    torch.distributed, halo planes / interface values / reductions as NVLink peer stores.  bench.py --parity-only gathers
    the slab fields after two steps and compares them bit for bit with the single-GPU run of the same case (all ranks
    claim the box-dataflow kernels).  Needs `world` visible devices."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (the driver's GPU test box has one; gpurun --gpus %d runs it)" % (world, world))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29500 + world + os.getpid() % 200
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(root, "bench.py"),
                        "--gpus", str(world), "--parity-only", "--parity-size", "48"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-1500:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["parity_check"].startswith("bit-exact"), line
