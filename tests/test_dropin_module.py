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
    if not os.access(BIN, os.X_OK):
        pytest.skip("hydro_b200/host/_build/hydro_gpu not built (needs /root/reference at build time)")
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
