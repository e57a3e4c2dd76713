import sys, os, time
t0 = time.time()
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import cases
from oracle_api import Oracle
from hydro_b200.capi import Hydro
import test_gpu_parity as T
for name in ("dam2d_32x16_meshvel_vx", "dam3d_24x8x8_meshvel_vcx"):
    p, n = cases.GOLDEN_CASES[name]
    gpu, cpu = Hydro(p), Oracle(p)
    worst = 0.
    for _ in range(n):
        sg, sc = gpu.step(), cpu.step()
        assert sg.simple_iterations == sc.simple_iterations and sg.pressure_sweeps_total == sc.pressure_sweeps_total, (name, "counts")
    errs = {}
    for f in T.FIELDS:
        if T.has_field(cpu, f):
            errs[f] = T.field_error(gpu, cpu, f)
    g = np.load(os.path.join(T.GOLD, "ref_%s.npz" % name))
    gerr = {}
    for fname, key in T.GOLD_KEYS.items():
        if key in g.files and T.has_field(cpu, fname):
            gerr[fname] = T.field_error(gpu, cpu, fname, ref=g[key])
    st = max(abs(a - b) / max(abs(b), 1e-300) for a, b in zip([*sg.center[1], *sg.velocity[1]][:4], [*sc.center[1], *sc.velocity[1]][:4]) if b != 0)
    print(name, "max field err vs oracle %.2e (%s), vs reference fixture %.2e, stat rel %.2e" % (max(errs.values()), max(errs, key=errs.get), max(gerr.values()), st), flush=True)
print("elapsed %.1f s" % (time.time() - t0))
