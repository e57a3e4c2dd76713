"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle.so (the CPU restatement).

Mirrors hydro_b200.capi.Hydro (the product binding) call for call so that parity tests
issue identical calls against both.  Nothing under hydro_b200/ imports this file.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from hydro_b200.config import F, FACE_FIELDS, HgConfig, HgStepStats, Params

ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
_lib = None


def lib(fast=False):
    global _lib
    name = "liboracle_fast.so" if fast else "liboracle.so"
    path = os.path.join(ORACLE_DIR, name)
    src = os.path.join(ORACLE_DIR, "hydro_oracle.c")
    hdr = os.path.join(os.path.dirname(ORACLE_DIR), "include", "hydro_gpu.h")
    if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, name], stdout=subprocess.DEVNULL)
    if fast:
        return _bind(C.CDLL(path))
    if _lib is None:
        _lib = _bind(C.CDLL(path))
    return _lib


def _bind(l):
    dp = C.POINTER(C.c_double)
    l.ho_create.argtypes = [C.POINTER(HgConfig), C.POINTER(C.c_void_p)]
    l.ho_destroy.argtypes = [C.c_void_p]
    l.ho_last_error.argtypes = [C.c_void_p]
    l.ho_last_error.restype = C.c_char_p
    l.ho_num_cells.argtypes = [C.c_void_p]
    l.ho_num_cells.restype = C.c_size_t
    l.ho_num_faces.argtypes = [C.c_void_p]
    l.ho_num_faces.restype = C.c_size_t
    l.ho_set_field.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t]
    l.ho_get_field.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t]
    l.ho_step.argtypes = [C.c_void_p, C.POINTER(HgStepStats)]
    for n in ("ho_fluid_start_step", "ho_fluid_make_iteration", "ho_fluid_finish_step", "ho_advection_step",
              "ho_heat_step", "ho_update_properties"):
        getattr(l, n).argtypes = [C.c_void_p]
    l.ho_fluid_convergence_indicator.argtypes = [C.c_void_p, dp]
    l.ho_fluid_is_converged.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    l.ho_fluid_auto_time_step.argtypes = [C.c_void_p, dp]
    l.ho_set_time_step.argtypes = [C.c_void_p, C.c_double, C.c_double]
    l.ho_calc_stat.argtypes = [C.c_void_p, C.POINTER(HgStepStats)]
    l.ho_interp_grad.argtypes = [C.c_void_p, dp, C.c_int, C.c_int, dp, dp, dp]
    l.ho_linear_solve.argtypes = [C.c_void_p, C.c_int, C.POINTER(dp), dp, dp, C.c_double, C.c_int, C.c_double,
                                  C.POINTER(C.c_int), dp]
    l.ho_smooth_field.argtypes = [C.c_void_p, dp, C.c_int, dp]
    l.ho_last_residuals.argtypes = [C.c_void_p, dp, C.c_int, C.POINTER(C.c_int)]
    return l


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    """Same surface as hydro_b200.capi.Hydro."""
    prefix = "ho_"

    def __init__(self, params, fast=False):
        self.l = lib(fast)
        self.params = params if isinstance(params, Params) else Params(params)
        self.cfg = self.params.to_struct()
        h = C.c_void_p()
        rc = self.l.ho_create(C.byref(self.cfg), C.byref(h))
        if rc != 0:
            raise RuntimeError("ho_create failed: %s" % self.l.ho_last_error(None).decode())
        self.h = h
        self.dim = self.cfg.dim
        self.nc = self.l.ho_num_cells(h)
        self.nf = self.l.ho_num_faces(h)

    def close(self):
        if self.h:
            self.l.ho_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.l.ho_last_error(self.h).decode())

    def get(self, name):
        fid = F[name] if isinstance(name, str) else name
        n = self.nf if fid in FACE_FIELDS else self.nc
        a = np.empty(n, dtype=np.float64)
        self._chk(self.l.ho_get_field(self.h, fid, _dptr(a), n))
        return a

    def set(self, name, arr):
        fid = F[name] if isinstance(name, str) else name
        a = np.ascontiguousarray(arr, dtype=np.float64)
        self._chk(self.l.ho_set_field(self.h, fid, _dptr(a), a.size))

    def step(self):
        st = HgStepStats()
        self._chk(self.l.ho_step(self.h, C.byref(st)))
        return st

    def residuals(self):
        buf = np.empty(4096)
        n = C.c_int()
        self.l.ho_last_residuals(self.h, _dptr(buf), 4096, C.byref(n))
        return buf[:n.value].copy()

    def fluid_start_step(self):
        self._chk(self.l.ho_fluid_start_step(self.h))

    def fluid_make_iteration(self):
        self._chk(self.l.ho_fluid_make_iteration(self.h))

    def fluid_finish_step(self):
        self._chk(self.l.ho_fluid_finish_step(self.h))

    def fluid_convergence_indicator(self):
        v = C.c_double()
        self._chk(self.l.ho_fluid_convergence_indicator(self.h, C.byref(v)))
        return v.value

    def fluid_auto_time_step(self):
        v = C.c_double()
        self._chk(self.l.ho_fluid_auto_time_step(self.h, C.byref(v)))
        return v.value

    def advection_step(self):
        self._chk(self.l.ho_advection_step(self.h))

    def heat_step(self):
        self._chk(self.l.ho_heat_step(self.h))

    def update_properties(self):
        self._chk(self.l.ho_update_properties(self.h))

    def calc_stat(self):
        st = HgStepStats()
        self._chk(self.l.ho_calc_stat(self.h, C.byref(st)))
        return st

    def interp_grad(self, u, cond, comp=0):
        u = np.ascontiguousarray(u, dtype=np.float64)
        g = [np.zeros(self.nc) for _ in range(3)]
        self._chk(self.l.ho_interp_grad(self.h, _dptr(u), cond, comp, _dptr(g[0]), _dptr(g[1]), _dptr(g[2])))
        return g[:self.dim]

    def linear_solve(self, solver, coeffs, rhs, tol=0.0, limit=100, relax=1.9):
        from hydro_b200.config import LINEAR_SOLVERS
        sid = LINEAR_SOLVERS[solver] if isinstance(solver, str) else solver
        cs = [np.ascontiguousarray(c, dtype=np.float64) if c is not None else None for c in coeffs]
        arr = (C.POINTER(C.c_double) * 7)(*[_dptr(c) if c is not None else None for c in cs])
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros(self.nc)
        it = C.c_int()
        df = C.c_double()
        self._chk(self.l.ho_linear_solve(self.h, sid, arr, _dptr(rhs), _dptr(x), tol, limit, relax,
                                         C.byref(it), C.byref(df)))
        return x, it.value, df.value

    def smooth_field(self, u, repeat):
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.zeros(self.nc)
        self._chk(self.l.ho_smooth_field(self.h, _dptr(u), repeat, _dptr(out)))
        return out
