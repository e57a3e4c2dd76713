"""Parity of the CUDA path (through the C ABI) against the oracle and the reference fixtures.

Tolerances: the device code keeps the reference's operation order and is compiled without FMA
contraction, so fields agree with the oracle to rounding; the stated bound is 1e-12 relative L2 per
field after the case's steps (north_star: "relative L2 of velocity/pressure/scalar after N steps"),
equal SIMPLE-iteration counts and equal linear-solver sweep counts.  CalcStat sums are parallel
reductions: 1e-11 relative.
"""
import os

import numpy as np
import pytest

import cases
from hydro_b200.config import F
from oracle_api import Oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-12

FIELDS = ["VELOCITY_X", "VELOCITY_Y", "VELOCITY_Z", "PRESSURE", "VOLUME_FLUX", "DENSITY", "VISCOSITY",
          "PARTIAL_DENSITY_0", "PARTIAL_DENSITY_1", "VOLUME_FRACTION_0", "FORCE_X", "FORCE_Y", "TEMPERATURE",
          "VELOCITY_PREV_X", "PRESSURE_PREV", "VOLUME_FLUX_PREV", "EXCLUDED"]
GOLD_KEYS = {"VELOCITY_X": "u0", "VELOCITY_Y": "u1", "VELOCITY_Z": "u2", "PRESSURE": "p", "VOLUME_FLUX": "flux",
             "PARTIAL_DENSITY_0": "pd0", "PARTIAL_DENSITY_1": "pd1", "TEMPERATURE": "temp", "DENSITY": "rho",
             "VISCOSITY": "mu"}


def rel_l2(a, b):
    n = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (n if n > 0 else 1.0)


def has_field(o, name):
    if name.endswith("_Z") and o.dim == 2:
        return False
    if name == "TEMPERATURE" and not o.cfg.heat_enable:
        return False
    if name[-1] in "12" and name[:-2] in ("PARTIAL_DENSITY", "VOLUME_FRACTION") and int(name[-1]) >= o.cfg.num_phases:
        return False
    return True


def vector_scale(cpu, name):
    """Components of a vector field are measured against the L2 norm of the whole vector field
    ("relative L2 of velocity"): a component that is ~0 by symmetry has no scale of its own."""
    for stem in ("VELOCITY_PREV_", "VELOCITY_", "FORCE_"):
        if name.startswith(stem) and name[len(stem):] in ("X", "Y", "Z"):
            comps = [cpu.get(stem + c) for c in "XYZ"[:cpu.dim]]
            return np.sqrt(sum(np.linalg.norm(c) ** 2 for c in comps))
    return None


def field_error(gpu, cpu, name, ref=None):
    a = gpu.get(name)
    b = cpu.get(name) if ref is None else ref
    assert np.all(np.isfinite(a)), name
    sc = vector_scale(cpu, name)
    if sc is None:
        sc = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (sc if sc > 0 else 1.0)


def compare_states(gpu, cpu, tol=TOL):
    for name in FIELDS:
        if not has_field(cpu, name):
            continue
        e = field_error(gpu, cpu, name)
        assert e <= tol, (name, e)


def run_both(p, nsteps, tol=TOL, stat_tol=1e-11):
    from hydro_b200.capi import Hydro
    gpu, cpu = Hydro(p), Oracle(p)
    compare_states(gpu, cpu, tol)   # initial state (module constructor)
    for _ in range(nsteps):
        sg, sc = gpu.step(), cpu.step()
        assert sg.simple_iterations == sc.simple_iterations
        assert sg.pressure_sweeps_total == sc.pressure_sweeps_total
        assert sg.advection_substeps == sc.advection_substeps
        rg, rc = gpu.residuals(), cpu.residuals()
        assert len(rg) == len(rc)
        np.testing.assert_allclose(rg, rc, rtol=1e-9, atol=1e-300)
        np.testing.assert_allclose(sg.dt, sc.dt, rtol=1e-13)
        np.testing.assert_allclose(sg.pressure_last_diff, sc.pressure_last_diff, rtol=1e-8, atol=1e-300)
        for i in range(cpu.cfg.num_phases):
            np.testing.assert_allclose([sg.volume[i], sg.mass[i], sg.pd_min[i], sg.pd_max[i], *sg.center[i], *sg.velocity[i]],
                                       [sc.volume[i], sc.mass[i], sc.pd_min[i], sc.pd_max[i], *sc.center[i], *sc.velocity[i]],
                                       rtol=stat_tol, atol=1e-14)
    compare_states(gpu, cpu, tol)
    return gpu, cpu


@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
def test_gpu_matches_oracle_and_reference_fixture(name):
    p, nsteps = cases.GOLDEN_CASES[name]
    # rt3d_8_jacobi_dtauto is an unstable flow (dt_auto lets dt jump): the 1-ulp difference between
    # device and host sin() in the initial velocity is amplified ~1e7 times in 3 steps; its exactness is
    # covered by test_gpu_bit_exact_from_identical_state instead
    loose = name == "rt3d_8_jacobi_dtauto"
    ill = cases.ILL_CONDITIONED.get(name)   # e.g. the mesh velocity from the differenced centre of a phase (see cases.py)
    gpu, cpu = run_both(p, nsteps, tol=1e-6 if loose else (ill or TOL), stat_tol=1e-6 if loose else (ill or 1e-11))
    g = np.load(os.path.join(GOLD, "ref_%s.npz" % name))
    for fname, key in GOLD_KEYS.items():
        if key in g.files and has_field(cpu, fname):
            e = field_error(gpu, cpu, fname, ref=g[key])
            assert e <= (1e-6 if loose else (ill or 1e-11)), (fname, e)


@pytest.mark.parametrize("name", ["rt3d_16", "dam3d_32x10x10", "thermal2d_32x16", "cavity_16", "rt3d_8_jacobi_dtauto",
                                  "dam2d_36x20", "rt3d_12x10x9", "cavity_12_lurelaxed", "cavity_12_velgs_plu", "rt3d_8_veljacobi",
                                  "thermal2d_24x12_vellur_heatgs", "rt3d_8_sharp_split", "dam2d_32x16_sharp", "channel2d_32x16_outlet",
                                  "rt3d_8_inlet_outlet", "rt3d_8_settling", "dam2d_32x16_settling", "rt3d_8_simpler", "cavity_16_simpler"])
def test_gpu_bit_exact_from_identical_state(name):
    """Started from bit-identical fields (the oracle's initial state uploaded through hg_set_field, which
    removes the 1-ulp difference between device and host sin() in the initial velocity), the CUDA path
    reproduces the oracle -- hence the reference -- to the last bit: same operation order, no FMA."""
    from hydro_b200.capi import Hydro
    p, nsteps = cases.GOLDEN_CASES[name]
    gpu, cpu = Hydro(p), Oracle(p)
    for f in ("VELOCITY_X", "VELOCITY_Y", "VELOCITY_Z", "VOLUME_FLUX", "PARTIAL_DENSITY_0", "PARTIAL_DENSITY_1"):
        if has_field(cpu, f):
            gpu.set(f, cpu.get(f))
    gpu.update_properties()
    for _ in range(nsteps):
        gpu.step(), cpu.step()
    for f in ("VELOCITY_X", "VELOCITY_Y", "VELOCITY_Z", "PRESSURE", "VOLUME_FLUX", "PARTIAL_DENSITY_0",
              "PARTIAL_DENSITY_1", "TEMPERATURE", "DENSITY", "VISCOSITY"):
        if has_field(cpu, f):
            assert np.array_equal(gpu.get(f), cpu.get(f)), f
    assert np.array_equal(gpu.residuals(), cpu.residuals())


def test_gpu_reproduces_reference_cavity_sample():
    """The reference's golden sample (examples/cavity/sample: 64^2, Re 3200, 1000 SIMPLE iterations x 101
    Gauss-Seidel sweeps) on the GPU: all 1000 residuals equal the shipped exp.iter_history.plt tokens and the
    ParaView file written from the device fields equals the shipped exp.field.0.vts byte for byte."""
    from hydro_b200 import output
    from hydro_b200.capi import Hydro
    g = np.load(os.path.join(GOLD, "cavity_sample.npz"))
    p = cases.cavity_kat()
    gpu = Hydro(p)
    st = gpu.step()
    assert st.simple_iterations == 1000
    tok = np.array(["%g" % v for v in gpu.residuals()])
    assert np.array_equal(tok, g["rs_tokens"])
    assert output.vts_text(gpu, p) == str(g["vts_text"])


@pytest.mark.parametrize("n", [32, 48])
def test_gpu_rt3d_medium(n):
    """W4 at sizes the oracle finishes in seconds; fixed work (3 SIMPLE x 101 sweeps)."""
    run_both(cases.rt3d(n), 2)


def test_gpu_rt3d_128_against_fast_oracle():
    """W4 at 128^3 (2.1 M cells, fixed work): the largest size the C restatement finishes in about ten seconds per
    step (liboracle_fast.so, -O3).  Same bound as the small cases: 1e-12 relative L2 per field, equal SIMPLE-iteration
    and sweep counts.  Covers the index arithmetic of every kernel at a size where the box dataflow runs many boxes per
    SM and several sweep groups (the 256^3 test below can only compare two GPU schedules with each other)."""
    from hydro_b200.capi import Hydro
    p = cases.rt3d(128)
    gpu, cpu = Hydro(p), Oracle(p, fast=True)
    sg, sc = gpu.step(), cpu.step()
    assert sg.simple_iterations == sc.simple_iterations == 3
    assert sg.pressure_sweeps_total == sc.pressure_sweeps_total
    np.testing.assert_allclose(gpu.residuals(), cpu.residuals(), rtol=1e-9)
    np.testing.assert_allclose(sg.pressure_last_diff, sc.pressure_last_diff, rtol=1e-8)
    compare_states(gpu, cpu, TOL)


def test_gpu_rt3d_256_against_fast_oracle():
    """BASELINE.json's full size against the oracle itself: RT-3D 256^3, one step of the bench workload (3 SIMPLE iterations x
    101 sweeps, advection, properties, statistics).  The serial C restatement needs about a minute and a half for it
    (liboracle_fast.so), which is why the other full-size test only compares GPU schedules with each other; this one closes the
    gap: 1e-12 relative L2 per field, equal iteration and sweep counts, at the size every number of bench.py is quoted on."""
    from hydro_b200.capi import Hydro
    p = cases.rt3d(256)
    gpu, cpu = Hydro(p), Oracle(p, fast=True)
    sg, sc = gpu.step(), cpu.step()
    assert sg.simple_iterations == sc.simple_iterations == 3
    assert sg.pressure_sweeps_total == sc.pressure_sweeps_total == 3 * 102
    np.testing.assert_allclose(gpu.residuals(), cpu.residuals(), rtol=1e-9)
    np.testing.assert_allclose(sg.pressure_last_diff, sc.pressure_last_diff, rtol=1e-8)
    compare_states(gpu, cpu, TOL)


def test_get_stats_does_not_advance_the_mesh_position():
    """CalcStat adds meshvel*dt to the mesh position (hydro2d.hpp:1526-1528) once per call: hg_calc_stat keeps that,
    hg_get_stats only returns the last statistics (the module constructor uses it)."""
    from hydro_b200.capi import Hydro
    p = cases.cavity(12, meshvel=(0.25, 0, 0), num_phases=1)
    gpu, cpu = Hydro(p), Oracle(p)
    a, b = gpu.get_stats(), gpu.get_stats()
    assert a.center[0][0] == b.center[0][0]
    sg, sc = gpu.step(), cpu.step()
    np.testing.assert_allclose(sg.center[0][0], sc.center[0][0], rtol=1e-12)
    assert gpu.get_stats().center[0][0] == sg.center[0][0]


def test_gpu_dam3d_as_shipped():
    """examples/broken_dam_3d as shipped (64x20x20, obstacle)."""
    run_both(cases.broken_dam_3d(64, 20, 20, lu_relaxed_num_iters_limit=60), 2)


def test_gpu_rt3d_early_stop_sweeps():
    """tolerance > 0: the pipelined sweeps must stop at exactly the reference's sweep count."""
    run_both(cases.rt3d(12, fixed_work=False, lu_relaxed_num_iters_limit=300, lu_relaxed_tolerance=1e-7,
                        num_iterations_limit=4, pressure_sweeps_per_check=32), 2)


@pytest.mark.parametrize("case", ["rt3d_40x36x20", "rt3d_70x19x33_tol", "dam3d_64x20x20", "rt3d_9x5x4"])
def test_gs_tiled_equals_hyperplane_kernel(case, monkeypatch):
    """The time-skewed tile sweeps (hg_gs_tiled.cuh) and the pipelined hyperplane sweeps (hg_solvers.cuh) are
    two schedules of the same lexicographic Gauss-Seidel: bitwise equal fields, sweep counts and norms."""
    from hydro_b200.capi import Hydro
    p = {"rt3d_40x36x20": cases.rt3d(8, Nx=40, Ny=36, Nz=20),
         "rt3d_70x19x33_tol": cases.rt3d(8, Nx=70, Ny=19, Nz=33, fixed_work=False, lu_relaxed_num_iters_limit=200,
                                         lu_relaxed_tolerance=1e-6, num_iterations_limit=3, pressure_sweeps_per_check=48),
         "dam3d_64x20x20": cases.broken_dam_3d(64, 20, 20, lu_relaxed_num_iters_limit=37),
         "rt3d_9x5x4": cases.rt3d(8, Nx=9, Ny=5, Nz=4, lu_relaxed_num_iters_limit=11)}[case]
    res = []
    for kern in ("tiled", "hyperplane"):
        monkeypatch.setenv("HYDRO_GS_KERNEL", kern)
        h = Hydro(p)
        st = [h.step() for _ in range(2)][-1]
        res.append((st, {n: h.get(n) for n in ("VELOCITY_X", "VELOCITY_Y", "VELOCITY_Z", "PRESSURE", "VOLUME_FLUX", "PARTIAL_DENSITY_1")}))
        h.close() if hasattr(h, "close") else None
    (sa, fa), (sb, fb) = res
    assert sa.pressure_sweeps_total == sb.pressure_sweeps_total and sa.simple_iterations == sb.simple_iterations
    assert sa.pressure_last_diff == sb.pressure_last_diff
    for n in fa:
        assert np.array_equal(fa[n], fb[n]), n


@pytest.mark.parametrize("case", ["rt3d_40x36x20", "dam3d_64x20x20", "rt3d_9x5x4", "thermal3d_33x18x12", "rt3d_70x40x9"])
def test_lu_tiled_equals_hyperplane_kernel(case, monkeypatch):
    """The column-box dataflow of the lu solver (hg_lu_tiled.cuh) and the hyperplane kernel with a grid barrier per plane
    (hg_solvers.cuh) are two schedules of the same forward + backward sweep (linear.hpp:533-566): bitwise equal."""
    from hydro_b200.capi import Hydro
    p = {"rt3d_40x36x20": cases.rt3d(8, Nx=40, Ny=36, Nz=20, lu_relaxed_num_iters_limit=9),
         "dam3d_64x20x20": cases.broken_dam_3d(64, 20, 20, lu_relaxed_num_iters_limit=12),
         "rt3d_9x5x4": cases.rt3d(8, Nx=9, Ny=5, Nz=4, lu_relaxed_num_iters_limit=11),
         "thermal3d_33x18x12": cases.rt3d(8, Nx=33, Ny=18, Nz=12, lu_relaxed_num_iters_limit=7, heat_enable=1,
                                          heat_box_lb=(-1., -1., -1.), heat_box_rt=(2., 0.01, 2.), heat_box_temperature=1.),
         "rt3d_70x40x9": cases.rt3d(8, Nx=70, Ny=40, Nz=9, lu_relaxed_num_iters_limit=5)}[case]
    names = ["VELOCITY_X", "VELOCITY_Y", "VELOCITY_Z", "PRESSURE", "VOLUME_FLUX", "PARTIAL_DENSITY_1"]
    if p.get("heat_enable"):
        names.append("TEMPERATURE")
    res = []
    for kern in ("tiled", "hyperplane"):
        monkeypatch.setenv("HYDRO_LU_KERNEL", kern)
        h = Hydro(p)
        st = [h.step() for _ in range(2)][-1]
        res.append((st, {n: h.get(n) for n in names}))
        h.close()
    (sa, fa), (sb, fb) = res
    assert sa.convergence_indicator == sb.convergence_indicator
    for n in fa:
        assert np.array_equal(fa[n], fb[n]), n


@pytest.mark.parametrize("case", ["rt3d_40x36x20", "dam3d_64x20x24", "rt3d_9x5x4", "rt3d_70x19x33_split_stf", "rt3d_33x40x41_pfix"])
def test_interior_kernels_equal_generic_kernels(case, monkeypatch):
    """hg_fast.cuh: the cells whose radius-2 stencil is all inner faces take fused interior kernels (gradients; source +
    assembly + transpose; fluxes + pressure rows + packing; correction; advection), the shell near walls / the obstacle /
    the fixed-pressure cell keeps the generic kernels over a cell list.  Same arithmetic, operation by operation: every
    field must equal the all-generic run (HYDRO_FAST=0) bit for bit."""
    from hydro_b200.capi import Hydro
    p = {"rt3d_40x36x20": cases.rt3d(8, Nx=40, Ny=36, Nz=20, lu_relaxed_num_iters_limit=9),
         "dam3d_64x20x24": cases.broken_dam_3d(64, 20, 24, lu_relaxed_num_iters_limit=12),
         "rt3d_9x5x4": cases.rt3d(8, Nx=9, Ny=5, Nz=4, lu_relaxed_num_iters_limit=11),
         "rt3d_70x19x33_split_stf": cases.rt3d(8, Nx=70, Ny=19, Nz=33, lu_relaxed_num_iters_limit=7, tvd_split=1, sigma=0.05,
                                               meshvel=(0.01, 0.02, -0.01), guess_extrapolation=0.5),
         "rt3d_33x40x41_pfix": cases.rt3d(8, Nx=33, Ny=40, Nz=41, lu_relaxed_num_iters_limit=8, pressure_fixed_enable=1,
                                          pressure_fixed_point=(0.5, 0.5, 0.5), pressure_fixed_value=0.25)}[case]
    names = ["VELOCITY_X", "VELOCITY_Y", "VELOCITY_Z", "PRESSURE", "VOLUME_FLUX", "PARTIAL_DENSITY_0", "PARTIAL_DENSITY_1"]
    res = []
    for fast in ("1", "0"):
        monkeypatch.setenv("HYDRO_FAST", fast)
        h = Hydro(p)
        st = [h.step() for _ in range(3)][-1]
        res.append((st, {n: h.get(n) for n in names}))
        h.close()
    (sa, fa), (sb, fb) = res
    assert sa.convergence_indicator == sb.convergence_indicator and sa.pressure_last_diff == sb.pressure_last_diff
    for n in names:
        assert np.array_equal(fa[n], fb[n]), n


def test_gpu_tvd_split_and_surface_tension():
    run_both(cases.broken_dam_2d(40, 24, tvd_split=1, sigma=0.07, lu_relaxed_num_iters_limit=40), 2, tol=1e-11)


def test_gpu_inlet_condition():
    run_both(cases.thermal_2d(24, 12, condition_left="inlet 0.3 0 0", condition_right="inlet 0.3 0 0"), 2)


# ---------------------------------------------------------------- kernel-level entries
def sin_field(n):
    return np.sin(np.arange(n, dtype=np.float64))   # test/benchmark/main.cpp:172 input pattern


@pytest.mark.parametrize("case", ["rt3d_12x10x9", "dam3d_32x10x10", "cavity_16"])
@pytest.mark.parametrize("cond", [0, 1, 2])
def test_interp_grad(case, cond):
    from hydro_b200.capi import Hydro
    p, _ = cases.GOLDEN_CASES[case]
    gpu, cpu = Hydro(p), Oracle(p)
    u = sin_field(cpu.nc)
    for a, b in zip(gpu.interp_grad(u, cond, 1), cpu.interp_grad(u, cond, 1)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("case", ["rt3d_12x10x9", "dam3d_32x10x10", "cavity_16"])
def test_smooth_field(case):
    from hydro_b200.capi import Hydro
    p, _ = cases.GOLDEN_CASES[case]
    gpu, cpu = Hydro(p), Oracle(p)
    u = sin_field(cpu.nc)
    for rep in (0, 1, 3):
        assert np.array_equal(gpu.smooth_field(u, rep), cpu.smooth_field(u, rep))


def random_rows(o, seed):
    """Diagonally dominant 7/5-point rows in the reference's term order z-,y-,x-,diag,x+,y+,z+."""
    rng = np.random.default_rng(seed)
    nc = o.nc
    rows = [-rng.random(nc) for _ in range(7)]
    rows[3] = 6.5 + rng.random(nc)
    if o.dim == 2:
        rows[0] = rows[6] = None
    return rows, rng.standard_normal(nc)


@pytest.mark.parametrize("case", ["rt3d_12x10x9", "cavity_16", "rt3d_16"])
@pytest.mark.parametrize("solver,tol,limit", [("lu", 0.0, 0), ("gauss_seidel", 0.0, 40), ("gauss_seidel", 1e-6, 500),
                                                ("jacobi", 1e-5, 300), ("jacobi", 0.0, 21), ("lu_relaxed", 1e-9, 40),
                                                ("lu_relaxed", 0.0, 5)])
def test_linear_solve(case, solver, tol, limit):
    from hydro_b200.capi import Hydro
    p, _ = cases.GOLDEN_CASES[case]
    gpu, cpu = Hydro(p), Oracle(p)
    rows, rhs = random_rows(cpu, 7)
    relax = 1.3 if solver == "gauss_seidel" else (0.4 if solver == "lu_relaxed" else 0.8)
    xg, ig, dg = gpu.linear_solve(solver, rows, rhs, tol, limit, relax)
    xc, ic, dc = cpu.linear_solve(solver, rows, rhs, tol, limit, relax)
    assert ig == ic
    assert np.array_equal(xg, xc)
    if solver != "lu":
        assert dg == dc


def test_error_convention():
    """Unsupported options fail in hg_create with a message instead of silently falling back."""
    from hydro_b200.capi import Hydro
    with pytest.raises(RuntimeError, match="simpler"):
        Hydro(cases.cavity(8, simpler=1, linear_solver_pressure="lu_relaxed"))
    with pytest.raises(RuntimeError, match="outlet"):
        Hydro(cases.rt3d(8, condition_right="outlet"), world_size=2, rank=0)
    with pytest.raises(RuntimeError, match="slabs support lu"):
        Hydro(cases.rt3d(8, linear_solver_velocity="jacobi"), world_size=2, rank=0)
    h = Hydro(cases.cavity(8))
    with pytest.raises(RuntimeError, match="bad field"):
        h.set("PRESSURE", np.zeros(3))
    h.set("PRESSURE", np.full(h.nc, np.nan))
    with pytest.raises(RuntimeError, match="NaN initial pressure"):
        h.step()


def test_set_get_roundtrip_and_idempotent_properties():
    from hydro_b200.capi import Hydro
    h = Hydro(cases.rt3d(8))
    for name in ("VELOCITY_X", "PRESSURE", "VOLUME_FLUX", "PARTIAL_DENSITY_1"):
        n = h.nf if F[name] == F["VOLUME_FLUX"] else h.nc
        v = 0.6 + 0.3 * sin_field(n)
        h.set(name, v)
        assert np.array_equal(h.get(name), v)
    h.update_properties()
    a = h.get("DENSITY").copy()
    h.update_properties()
    assert np.array_equal(h.get("DENSITY"), a)


def test_async_field_transfers_equal_blocking_ones():
    """hg_set_field_async / hg_get_field_async (pinned buffers, copy streams) against hg_set_field / hg_get_field."""
    import torch
    from hydro_b200.capi import Hydro
    p = cases.rt3d(12)
    a, b = Hydro(p), Hydro(p)
    a.step(), b.step()
    names_in = ["DENSITY", "VISCOSITY", "FORCE_Y"]
    names_out = ["VELOCITY_X", "VELOCITY_Y", "VELOCITY_Z", "PRESSURE"]
    rng = np.random.default_rng(3)
    pins = {n: torch.empty(a.nc, dtype=torch.float64).pin_memory() for n in names_in + names_out}
    for it in range(3):
        for n in names_in:
            v = a.get(n) * (1. + 1e-3 * rng.standard_normal(a.nc))
            a.set(n, v)
            pins[n].numpy()[:] = v
            b.set_from_async(n, pins[n].data_ptr())
        a.step(), b.step()
        for n in names_out:
            b.get_to_async(n, pins[n].data_ptr())
        b.synchronize()
        for n in names_out:
            assert np.array_equal(pins[n].numpy(), a.get(n)), (it, n)
    a.close(), b.close()


def test_step_begin_end_with_overlapped_uploads_equals_step():
    """hg_step = hg_step_begin + hg_step_end.  With the next step's property fields uploaded between the two halves
    (hg_set_field_async: the upload overlaps the running step and is applied in stream order after it) the run must equal, bit
    for bit, a run that sets the same fields with the blocking calls between whole steps."""
    import torch
    from hydro_b200.capi import Hydro
    p = cases.rt3d(24, lu_relaxed_num_iters_limit=20)
    a, b = Hydro(p), Hydro(p)
    ins = ["DENSITY", "VISCOSITY", "FORCE_Y"]
    names = ["VELOCITY_X", "VELOCITY_Y", "VELOCITY_Z", "PRESSURE", "VOLUME_FLUX"]
    rng = np.random.default_rng(5)
    base = {n: a.get(n) for n in ins}
    pert = [{n: base[n] * (1. + 0.01 * k + 1e-3 * rng.random(base[n].size)) for n in ins} for k in range(4)]
    pins = [{n: torch.from_numpy(pert[k][n].copy()).pin_memory() for n in ins} for k in range(4)]
    # a: blocking
    for k in range(4):
        for n in ins:
            a.set(n, pert[k][n])
        sa = a.step()
    # b: pipelined
    for n in ins:
        b.set_from_async(n, pins[0][n].data_ptr())
    b.step_begin()
    for k in range(4):
        if k + 1 < 4:
            for n in ins:
                b.set_from_async(n, pins[k + 1][n].data_ptr())
        sb = b.step_end()
        if k + 1 < 4:
            b.step_begin()
    b.synchronize()
    assert sa.convergence_indicator == sb.convergence_indicator and sa.pressure_sweeps_total == sb.pressure_sweeps_total
    for n in names:
        assert np.array_equal(a.get(n), b.get(n)), n
    with pytest.raises(RuntimeError, match="hg_step_begin"):
        b.step_end()
    a.close(), b.close()


def test_full_size_256_two_schedules_agree(monkeypatch):
    """BASELINE.json's full size (RT-3D 256^3, the bench workload: 3 SIMPLE iterations x 101 sweeps): the oracle cannot run
    it in seconds, so parity is carried by size-independent properties -- the box-dataflow kernels (k_gs_tiled,
    k_lu_tiled) and the hyperplane kernels (k_gs_persistent, k_lu_persistent) are independent schedules of the same
    lexicographic sweeps and must agree bit for bit; iteration / sweep counts must be the prescribed ones; the partial
    densities must conserve their volume integrals through the (conservative) advection step to round-off."""
    from hydro_b200.capi import Hydro
    p = cases.rt3d(256, fixed_work=True)
    names = ["VELOCITY_X", "VELOCITY_Y", "VELOCITY_Z", "PRESSURE", "PARTIAL_DENSITY_0", "PARTIAL_DENSITY_1"]
    res = []
    for kern in ("tiled", "hyperplane"):
        monkeypatch.setenv("HYDRO_GS_KERNEL", kern)
        monkeypatch.setenv("HYDRO_LU_KERNEL", kern)
        h = Hydro(p)
        m0 = [float(h.get("PARTIAL_DENSITY_%d" % q).sum()) for q in range(2)]
        st = h.step()
        f = {n: h.get(n) for n in names}
        h.close()
        # 101 sweeps per solve; the counter follows the reference's `iter` (linear.hpp:708-712): limit + 1, plus one
        assert st.simple_iterations == 3 and st.pressure_sweeps_total == 3 * 102 and st.advection_substeps == 1
        for q in range(2):
            assert abs(float(f["PARTIAL_DENSITY_%d" % q].sum()) - m0[q]) <= 1e-9 * abs(m0[q])
        res.append((st, f))
    (sa, fa), (sb, fb) = res
    assert sa.pressure_last_diff == sb.pressure_last_diff and sa.convergence_indicator == sb.convergence_indicator
    for n in names:
        assert np.array_equal(fa[n], fb[n]), n
        assert np.isfinite(fa[n]).all()
