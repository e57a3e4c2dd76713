"""Drop-in boundary, end to end: the reference's own console binary, rebuilt with ONE extra translation
unit (hydro_b200/host/hydro_gpu_module.cpp), runs the same .hydroconf script once with the CPU module
(`MODULE hydro3d`) and once with the GPU module (`MODULE hydro3d_gpu`).  The per-iteration residuals the
two modules log (`.....s=K, Rs=X`, hydro2d.hpp:1585-1586, 6 significant digits) must be identical text."""
import os
import re
import subprocess
import sys
import tempfile

import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
BIN = os.path.join(ROOT, "hydro_b200", "host", "_build", "hydro_gpu")
SHIM = os.path.join(ROOT, "hydro_b200", "host", "_build", "shim_check")


def run_console(params, nsteps):
    import refrun
    with tempfile.TemporaryDirectory() as tmp:
        p = type(params)(params)
        p["T"] = float(p["dt"]) * (nsteps - 0.5)
        p["max_frame_index"] = 0
        p["no_mesh_output"] = 1
        refrun.write_script(p, os.path.join(tmp, "start.hydroconf"), start=True)
        r = subprocess.run([BIN, "start.hydroconf"], cwd=tmp, capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, OMP_NUM_THREADS="4"))
        log = open(os.path.join(tmp, "exp.log")).read() if os.path.exists(os.path.join(tmp, "exp.log")) else ""
        assert "Experiment terminated" in log, (r.stdout[-800:], r.stderr[-800:], log[-800:])
        return re.findall(r"\.\.\.\.\.s=(\d+), Rs=(\S+)", log), log


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["rt3d", "cavity"])
def test_same_script_cpu_module_vs_gpu_module(case):
    assert os.access(BIN, os.X_OK), "hydro_b200/host/_build/hydro_gpu not built: __graft_entry__.build() makes it where /root/reference exists"
    if case == "rt3d":
        p, mods, nsteps = cases.rt3d(16), ("hydro3d", "hydro3d_gpu"), 3
    else:
        p, mods, nsteps = cases.cavity(32, num_iterations_limit=8, lu_relaxed_num_iters_limit=50), ("hydro2d", "hydro2d_gpu"), 3
    out = []
    for m in mods:
        p["MODULE"] = m
        rs, log = run_console(p, nsteps)
        out.append(rs)
    assert len(out[0]) >= nsteps and out[0] == out[1]


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["rt3d", "dam3d"])
def test_solver_shim_protocol_equals_hg_step(case):
    """Secondary boundary: hydro_b200/host/hydro_gpu.hpp mirrors solver::FluidSolver (fluid.hpp:202-253) and
    AdvectionSolverMulti (advection.hpp:62-84).  A C++ host program drives StartStep / IsConverged / MakeIteration /
    FinishStep / advection / properties / statistics the way hydro<Mesh>::step() does and dumps the getters' host mirrors;
    the fields must equal hg_step() on the same configuration bit for bit."""
    import numpy as np
    from hydro_b200.capi import Hydro
    assert os.access(SHIM, os.X_OK), "hydro_b200/host/_build/shim_check not built (make -C hydro_b200/host shim)"
    p = cases.rt3d(12) if case == "rt3d" else cases.broken_dam_3d(32, 10, 10, lu_relaxed_num_iters_limit=40, advection_dt_factor=1.0)
    nsteps = 2
    h = Hydro(p)
    for _ in range(nsteps):
        st = h.step()
    names = ["VELOCITY_X", "VELOCITY_Y", "VELOCITY_Z", "PRESSURE", "VOLUME_FLUX", "PARTIAL_DENSITY_0", "PARTIAL_DENSITY_1"]
    want = [h.get(n) for n in names]
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "cfg.bin"), "wb") as f:
            f.write(bytes(h.cfg))
        r = subprocess.run([SHIM, os.path.join(tmp, "cfg.bin"), str(nsteps), os.path.join(tmp, "out.bin")],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-800:]
        raw = open(os.path.join(tmp, "out.bin"), "rb").read()
    got, off = [], 0
    while off < len(raw):
        n = int(np.frombuffer(raw, dtype=np.uint64, count=1, offset=off)[0])
        got.append(np.frombuffer(raw, dtype=np.float64, count=n, offset=off + 8))
        off += 8 + 8 * n
    assert len(got) == len(want)
    for n, a, b in zip(names, got, want):
        assert np.array_equal(a, b), n
    m = re.search(r"indicator (\S+)", r.stdout)
    assert float(m.group(1)) == st.convergence_indicator


@pytest.mark.gpu
def test_module_writes_the_reference_output_files():
    """SURVEY 8f rank 2 inside the plugin: hydro_gpu<DIM>::write_results writes the ParaView frames, the .pvd collection and the
    scalar series on the reference's schedule (hydro2d.hpp:1623-1652, output_paraview.hpp:34-173, output.hpp:230-264).  The cavity
    case starts from identical fields on both paths (no transcendental initial condition), the GPU path is bit-exact on it, so
    every file must equal the CPU module's byte for byte."""
    import refrun
    assert os.access(BIN, os.X_OK), "hydro_b200/host/_build/hydro_gpu not built"
    files = {}
    for m in ("hydro2d", "hydro2d_gpu"):
        p = cases.cavity(24, num_iterations_limit=6, lu_relaxed_num_iters_limit=40)
        p["MODULE"] = m
        nsteps = 6
        p["T"] = float(p["dt"]) * (nsteps - 0.5)
        p["max_frame_index"] = 3
        p["max_frame_scalar_index"] = 6
        p["output_viscosity"] = 1
        with tempfile.TemporaryDirectory() as tmp:
            refrun.write_script(p, os.path.join(tmp, "start.hydroconf"), start=True)
            r = subprocess.run([BIN, "start.hydroconf"], cwd=tmp, capture_output=True, text=True, timeout=600,
                               env=dict(os.environ, OMP_NUM_THREADS="4"))
            log = open(os.path.join(tmp, "exp.log")).read() if os.path.exists(os.path.join(tmp, "exp.log")) else ""
            assert "Experiment terminated" in log, (r.stdout[-800:], r.stderr[-800:], log[-800:])
            files[m] = {f: open(os.path.join(tmp, f), "rb").read() for f in sorted(os.listdir(tmp))
                        if f.endswith((".vts", ".pvd", ".dat"))}
    a, b = files["hydro2d"], files["hydro2d_gpu"]
    assert sorted(a) == sorted(b) and len([f for f in a if f.endswith(".vts")]) >= 3, (sorted(a), sorted(b))
    for f in a:
        assert a[f] == b[f], f
