"""Times the lu box kernel on small meshes (one box, chains of boxes): cycles per step and per link."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from hydro_b200.capi import Hydro
for (nx, ny, nz) in [(32, 8, 256), (32, 8, 1024), (64, 8, 256), (32, 16, 256), (32, 64, 256), (256, 8, 256), (256, 256, 64), (256, 256, 256)]:
    p = cases.rt3d(8, Nx=nx, Ny=ny, Nz=nz, lu_relaxed_num_iters_limit=3)
    h = Hydro(p)
    h.step()
    h.profile_enable(True); h.profile_read(1); h.profile_read(0)
    for _ in range(3):
        h.step()
    n, ms = h.profile_read(1)
    gn, gms = h.profile_read(0)
    steps = nx + ny + nz - 2
    print("%4dx%4dx%4d: lu %.4f ms per solve (2 launches) -> %.0f cycles per hyperplane step; gs(4 sweeps) %.3f ms" % (nx, ny, nz, ms / n, ms / n / 2 * 1e-3 * 1.965e9 / steps, gms / gn))
    h.close()
