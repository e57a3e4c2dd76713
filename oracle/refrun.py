"""TEST INFRASTRUCTURE ONLY: run the real reference (oracle/_ref/ref_dump) on a Params set.

Only usable where oracle/_ref has been built (`make -C oracle ref`, needs /root/reference at
build time).  ref_dump itself does not need /root/reference at run time: the script passed
to it carries every parameter explicitly (no `run general.hydroconf`).
"""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DUMP = os.path.join(HERE, "_ref", "ref_dump")
REF_HYDRO = os.path.join(HERE, "_ref", "hydro")


def available():
    return os.access(REF_DUMP, os.X_OK)


def write_script(params, path, extra=(), start=False):
    lines = ["set c_bool force_overwrite 1", "set c_bool force_yes 1", "set c_int max_threads_number 1",
             "ae exp1", "set string name exp"]
    lines += params.hydroconf_lines()
    lines += list(extra)
    lines += ["init"] + (["start"] if start else [])
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def run_ref_dump(params, nsteps, iters=True, threads=1):
    """Returns (dict name -> np.ndarray, stdout text)."""
    with tempfile.TemporaryDirectory() as tmp:
        script = os.path.join(tmp, "s.hydroconf")
        out = os.path.join(tmp, "out")
        os.mkdir(out)
        write_script(params, script, extra=["set bool no_output 1"])
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        cmd = [REF_DUMP, script, str(nsteps), out] + (["--iters"] if iters else [])
        r = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("ref_dump failed (%d): %s\n%s" % (r.returncode, r.stderr[-2000:], r.stdout[-2000:]))
        res = {fn[:-4]: np.load(os.path.join(out, fn)) for fn in os.listdir(out) if fn.endswith(".npy")}
        sweeps = [int(l.split("=")[1].split(",")[0]) for l in r.stdout.splitlines() if l.startswith("iter =")]
        res["lin_iters"] = np.array(sweeps, dtype=np.float64)
        return res, r.stdout


def run_reference_binary(params, nsteps, threads=None, timeout=3600):
    """Runs the UNMODIFIED reference binary (oracle/_ref/hydro) on `params` for `nsteps` time steps and
    returns (per-step wall seconds `t_all` from its log, dict of its MultiTimer totals, log text).
    Reference timing sources: control/module.cpp:121-132 (t_all), control/experiment.cpp:165-187."""
    import re
    if not os.access(REF_HYDRO, os.X_OK):
        raise RuntimeError("oracle/_ref/hydro not built")
    threads = threads or os.cpu_count()
    with tempfile.TemporaryDirectory() as tmp:
        script = os.path.join(tmp, "start.hydroconf")
        p = type(params)(params)
        p["T"] = float(p["dt"]) * (nsteps - 0.5)
        p["max_frame_index"] = 0
        p["no_mesh_output"] = 1
        write_script(p, script, start=True)
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        r = subprocess.run([REF_HYDRO, script], cwd=tmp, env=env, capture_output=True, text=True, timeout=timeout)
        logp = os.path.join(tmp, "exp.log")
        log = open(logp).read() if os.path.exists(logp) else ""
        if "Experiment terminated" not in log:
            raise RuntimeError("reference run failed: %s\n%s" % (r.stderr[-1000:], r.stdout[-1000:]))
        t_all = [float(x) for x in re.findall(r"t_all=([0-9.eE+-]+)", log)]
        timers = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"^([\w.\- ]+?) : ([0-9.eE+-]+)$", log, re.M)}
        return t_all, timers, log
