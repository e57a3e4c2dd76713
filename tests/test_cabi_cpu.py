"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/hydro_gpu.h
declares; the ctypes structs match the C structs; parameter parsing mirrors the reference's names;
hg_create refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import cases
from hydro_b200 import capi
from hydro_b200.config import HgConfig, HgStepStats, Params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "hydro_gpu.h")).read()
    return sorted(set(re.findall(r"\b(hg_[a-z_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), s
    assert set(capi.SYMBOLS) <= set(syms)


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hydro_gpu.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(hg_config),sizeof(hg_step_stats),offsetof(hg_config,num_phases),offsetof(hg_config,world_size),'
                   'offsetof(hg_step_stats,center));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out == [C.sizeof(HgConfig), C.sizeof(HgStepStats), HgConfig.num_phases.offset, HgConfig.world_size.offset,
                   HgStepStats.center.offset]


def test_defaults_match_general_hydroconf():
    lib = capi.load_library()
    c = HgConfig()
    lib.hg_config_defaults(C.byref(c))
    p = Params().to_struct()
    for name, _ in HgConfig._fields_:
        a, b = getattr(c, name), getattr(p, name)
        if hasattr(a, "__len__"):
            a, b = [list(x) if hasattr(x, "__len__") else x for x in a], [list(x) if hasattr(x, "__len__") else x for x in b]
        assert a == b, name


def test_params_parse_hydroconf_text():
    p = Params()
    p.read_hydroconf("""
        set string MODULE hydro3d
        set int Nx 16   # comment
        set vect B (3.2, 1, 1)
        set double dt 0.0025
        set string condition_top "wall 1 0 0"
        set vect pressure_fixed_point (0, 1, 0)
        set int max_frame_index $(Nx)0
        del pressure_fixed_point
    """)
    c = p.to_struct()
    assert (c.dim, c.Nx, c.dt) == (3, 16, 0.0025)
    assert list(c.B) == [3.2, 1.0, 1.0]
    assert list(c.condition_velocity[3]) == [1.0, 0.0, 0.0]
    assert c.pressure_fixed_enable == 0 and p["max_frame_index"] == 160


def test_substitution_prints_values_like_the_console():
    """`$(name)` substitutes what the console prints for the parameter (console.cpp:464-509: stream default formatting, one
    token): rt/mfer.hydroconf:51,70 sets `double T 20` and `int max_frame_index $(T)0` = 200."""
    p = Params()
    p.read_hydroconf("""
        set double T 20
        set int max_frame_index $(T)0
        set double half 0.5
        set double dt_out $(half)
        set int n $T
    """)
    assert p["max_frame_index"] == 200 and p["dt_out"] == 0.5 and p["n"] == 20


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="needs the reference tree (build container)")
def test_every_example_script_of_the_reference_parses():
    import glob
    gen = open("/root/reference/examples/general.hydroconf").read()
    files = sorted(glob.glob("/root/reference/examples/*/*.hydroconf"))
    assert len(files) > 40
    for f in files:
        p = Params()
        p.read_hydroconf(gen)
        p.read_hydroconf(open(f).read())
    p = Params()
    p.read_hydroconf(gen)
    p.read_hydroconf(open("/root/reference/examples/broken_dam_3d/broken_dam_3d.hydroconf").read())
    c = p.to_struct()
    assert (c.dim, c.Nx, c.Ny, c.Nz) == (3, 167, 57, 62)
    # the cases as their start scripts layer them (run general / mfer / par / add files, then the remaining `set` lines up to
    # `init`): every case builds an hg_config except those that need images, radiation or chemistry, which fail loudly
    accepted, rejected = [], []
    ex = "/root/reference/examples"
    for d in sorted(os.listdir(ex)):
        st = os.path.join(ex, d, "start.hydroconf")
        if not os.path.exists(st):
            continue
        p = Params()
        try:
            for line in open(st).read().splitlines():
                m = re.match(r"\s*run\s+(\S+)", line)
                if m:
                    p.read_hydroconf(open(os.path.normpath(os.path.join(ex, d, m.group(1)))).read())
                else:
                    p.read_hydroconf(line)
                if line.strip() == "init":
                    break
            p.to_struct()
            accepted.append(d)
        except ValueError as e:
            assert "GPU path" in str(e), (d, e)
            rejected.append(d)
    assert accepted == ["broken_dam_2d", "broken_dam_3d", "cavity", "coal3d", "epfl", "mixing", "mixing3d", "mortazavi",
                        "mortazavi3d", "rt"]
    assert rejected == ["mixingimg", "reaction_2d", "reaction_3d", "rtimg", "vortimg"]


def test_meshvel_auto_maps_to_the_struct():
    """`set string meshvel_auto vx|vcx` (hydro2d.hpp:1510-1524; present = on, any other value is the reference's assert)."""
    c = cases.broken_dam_2d(16, 8, meshvel_auto="vcx", meshvel_weight=0.25).to_struct()
    assert (c.meshvel_auto, c.meshvel_weight) == (2, 0.25)
    assert cases.broken_dam_2d(16, 8, meshvel_auto="vx").to_struct().meshvel_auto == 1
    assert cases.broken_dam_2d(16, 8).to_struct().meshvel_auto == 0
    with pytest.raises(ValueError, match="Unknown meshvel_auto"):
        cases.broken_dam_2d(16, 8, meshvel_auto="phase0").to_struct()


def test_missing_parameter_raises_like_reference():
    p = Params()
    del p["Nx"]
    with pytest.raises(KeyError, match="'Nx' undefined"):
        p.to_struct()
    with pytest.raises(ValueError, match="Unknown linear solver"):
        Params(linear_solver_pressure="pardiso").to_struct()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        capi.Hydro(cases.cavity(8))


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hydro_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+\S*oracle|liboracle|#include\s+\"[^\"]*oracle|dlopen", txt, re.M), f


@pytest.mark.parametrize("key,val", [("compressible_enable", 1), ("deforming_velocity", 1), ("radiation_enable", 1),
                                     ("chemistry", "Nadirov"), ("chem_intensity", 0.5),
                                     ("imgu_init", "u.pgm"), ("velocity_is_carrier", 1), ("antidiffusion_factor", 0.3)])
def test_options_outside_the_gpu_path_are_rejected(key, val):
    """Options that change the reference's results (hydro2d.hpp:326-368, 1030-1217, 1294, 1387, 1511-1524) must not be
    dropped silently: Params.to_struct raises before a handle is created."""
    import cases
    p = cases.rt3d(8)
    p[key] = val
    with pytest.raises(ValueError, match="GPU path"):
        p.to_struct()
