"""ctypes binding of libhydro_gpu.so (the C ABI in include/hydro_gpu.h).

`Hydro` is the Python-side mirror of the reference's module object: constructed from the same
parameters (`Params`, names as in .hydroconf files), `step()` = hydro<Mesh>::step()
(hydro2d.hpp:1531-1621), fine-grained calls = the solver::UnsteadyIterativeSolver protocol
(solver.hpp:710-751).  Errors surface as RuntimeError carrying hg_last_error(), the analogue of the
reference's `throw std::string(...)`.  There is no CPU fallback: a missing library or CUDA device
raises.
"""
import ctypes as C
import os

import numpy as np

from .config import F, FACE_FIELDS, LINEAR_SOLVERS, HgConfig, HgStepStats, Params

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HYDRO_GPU_LIB", os.path.join(_HERE, "libhydro_gpu.so"))
_lib = None

SYMBOLS = [
    "hg_config_defaults", "hg_create", "hg_destroy", "hg_last_error", "hg_num_cells", "hg_num_faces",
    "hg_set_field", "hg_get_field", "hg_set_field_async", "hg_get_field_async", "hg_step", "hg_step_begin", "hg_step_end", "hg_run", "hg_fluid_start_step", "hg_fluid_make_iteration",
    "hg_fluid_convergence_indicator", "hg_fluid_is_converged", "hg_last_residuals", "hg_fluid_finish_step",
    "hg_fluid_auto_time_step", "hg_set_time_step", "hg_advection_step", "hg_heat_step",
    "hg_update_properties", "hg_calc_stat", "hg_interp_grad", "hg_linear_solve", "hg_smooth_field",
    "hg_timers", "hg_timers_enable", "hg_launch_count", "hg_device_synchronize", "hg_event_record",
    "hg_event_elapsed_ms", "hg_profile_enable", "hg_profile_read",
    "hg_ipc_record_size", "hg_ipc_export", "hg_ipc_import", "hg_link_local", "hg_get_stats", "hg_solver_kernel_name",
    "hg_profile_read_clocks",
]


def load_library():
    """Loads libhydro_gpu.so (no compute).  Raises if it has not been built (see __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    l = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    l.hg_config_defaults.argtypes = [C.POINTER(HgConfig)]
    l.hg_create.argtypes = [C.POINTER(HgConfig), C.POINTER(C.c_void_p)]
    l.hg_destroy.argtypes = [C.c_void_p]
    l.hg_last_error.argtypes = [C.c_void_p]
    l.hg_last_error.restype = C.c_char_p
    l.hg_num_cells.argtypes = [C.c_void_p]
    l.hg_num_cells.restype = C.c_size_t
    l.hg_num_faces.argtypes = [C.c_void_p]
    l.hg_num_faces.restype = C.c_size_t
    l.hg_set_field.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t]
    l.hg_get_field.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t]
    l.hg_set_field_async.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t]
    l.hg_get_field_async.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t]
    l.hg_step.argtypes = [C.c_void_p, C.POINTER(HgStepStats)]
    l.hg_step_begin.argtypes = [C.c_void_p]
    l.hg_step_end.argtypes = [C.c_void_p, C.POINTER(HgStepStats)]
    l.hg_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(HgStepStats)]
    for n in ("hg_fluid_start_step", "hg_fluid_make_iteration", "hg_fluid_finish_step", "hg_advection_step",
              "hg_heat_step", "hg_update_properties", "hg_device_synchronize"):
        getattr(l, n).argtypes = [C.c_void_p]
    l.hg_fluid_convergence_indicator.argtypes = [C.c_void_p, dp]
    l.hg_fluid_is_converged.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    l.hg_last_residuals.argtypes = [C.c_void_p, dp, C.c_int, C.POINTER(C.c_int)]
    l.hg_fluid_auto_time_step.argtypes = [C.c_void_p, dp]
    l.hg_set_time_step.argtypes = [C.c_void_p, C.c_double, C.c_double]
    l.hg_calc_stat.argtypes = [C.c_void_p, C.POINTER(HgStepStats)]
    l.hg_get_stats.argtypes = [C.c_void_p, C.POINTER(HgStepStats)]
    l.hg_profile_read_clocks.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
    l.hg_solver_kernel_name.argtypes = [C.c_void_p, C.c_int]
    l.hg_solver_kernel_name.restype = C.c_char_p
    l.hg_interp_grad.argtypes = [C.c_void_p, dp, C.c_int, C.c_int, dp, dp, dp]
    l.hg_linear_solve.argtypes = [C.c_void_p, C.c_int, C.POINTER(dp), dp, dp, C.c_double, C.c_int, C.c_double,
                                  C.POINTER(C.c_int), dp]
    l.hg_smooth_field.argtypes = [C.c_void_p, dp, C.c_int, dp]
    l.hg_timers.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_int, C.POINTER(C.c_int)]
    l.hg_timers_enable.argtypes = [C.c_void_p, C.c_int]
    l.hg_event_record.argtypes = [C.c_void_p, C.c_int]
    l.hg_event_elapsed_ms.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
    l.hg_profile_enable.argtypes = [C.c_void_p, C.c_int]
    l.hg_profile_read.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), dp]
    l.hg_ipc_record_size.argtypes = []
    l.hg_ipc_record_size.restype = C.c_size_t
    l.hg_ipc_export.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    l.hg_ipc_import.argtypes = [C.c_void_p, C.c_void_p]
    l.hg_link_local.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    l.hg_launch_count.argtypes = [C.c_void_p]
    l.hg_launch_count.restype = C.c_longlong
    _lib = l
    return l


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Hydro:
    """Device-resident experiment; same surface as tests/oracle_api.Oracle."""

    def __init__(self, params, device=0, world_size=1, rank=0, solver_ctas=0):
        """world_size > 1: this handle owns z-slab `rank` (hydro_b200.parallel.slab_range); the ranks must be
        linked (link_ipc / link_local, collective) before anything else is called."""
        self.l = load_library()
        self.params = params if isinstance(params, Params) else Params(params)
        self.world_size, self.rank = world_size, rank
        self.cfg = self.params.to_struct(device=device, world_size=world_size, rank=rank, solver_ctas=solver_ctas)
        h = C.c_void_p()
        rc = self.l.hg_create(C.byref(self.cfg), C.byref(h))
        if rc != 0:
            raise RuntimeError("hg_create failed (%d): %s" % (rc, self.l.hg_last_error(None).decode()))
        self.h = h
        self.dim = self.cfg.dim
        self.nc = self.l.hg_num_cells(h)
        self.nf = self.l.hg_num_faces(h)
        self._res = []

    # -- slab decomposition ---------------------------------------------------
    def link_ipc(self, dist):
        """Separate processes (one per GPU): all-gather the CUDA IPC records through torch.distributed, import."""
        import torch
        n = self.l.hg_ipc_record_size()
        rec = (C.c_ubyte * n)()
        self._chk(self.l.hg_ipc_export(self.h, rec, n))
        mine = torch.frombuffer(bytearray(rec), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            mine = mine.cuda()
        parts = [torch.empty_like(mine) for _ in range(self.world_size)]
        dist.all_gather(parts, mine)
        blob = b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)
        buf = C.create_string_buffer(blob, len(blob))
        self._chk(self.l.hg_ipc_import(self.h, buf))

    def link_local(self, handles):
        """Ranks living in this process (one thread each): `handles` = all Hydro objects in rank order."""
        arr = (C.c_void_p * len(handles))(*[h.h.value for h in handles])
        self._chk(self.l.hg_link_local(self.h, arr))

    def close(self):
        if getattr(self, "h", None):
            self.l.hg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.l.hg_last_error(self.h).decode())

    # -- fields ---------------------------------------------------------------
    def get(self, name, out=None):
        fid = F[name] if isinstance(name, str) else name
        n = self.nf if fid in FACE_FIELDS else self.nc
        a = out if out is not None else np.empty(n, dtype=np.float64)
        self._chk(self.l.hg_get_field(self.h, fid, _dptr(a), n))
        return a

    def set(self, name, arr):
        fid = F[name] if isinstance(name, str) else name
        a = np.ascontiguousarray(arr, dtype=np.float64)
        self._chk(self.l.hg_set_field(self.h, fid, _dptr(a), a.size))

    # -- stepping ---------------------------------------------------------------
    def step(self):
        st = HgStepStats()
        self._chk(self.l.hg_step(self.h, C.byref(st)))
        return st

    def step_begin(self):
        """Queues one time step (hg_step_begin); step_end() waits for it.  Transfers queued in between overlap the step."""
        self._chk(self.l.hg_step_begin(self.h))

    def step_end(self):
        st = HgStepStats()
        self._chk(self.l.hg_step_end(self.h, C.byref(st)))
        return st

    def run(self, nsteps):
        st = HgStepStats()
        self._chk(self.l.hg_run(self.h, nsteps, C.byref(st)))
        return st

    def residuals(self):
        buf = np.empty(4096)
        n = C.c_int()
        self._chk(self.l.hg_last_residuals(self.h, _dptr(buf), 4096, C.byref(n)))
        return buf[:n.value].copy()

    def fluid_start_step(self):
        self._chk(self.l.hg_fluid_start_step(self.h))

    def fluid_make_iteration(self):
        self._chk(self.l.hg_fluid_make_iteration(self.h))

    def fluid_finish_step(self):
        self._chk(self.l.hg_fluid_finish_step(self.h))

    def fluid_is_converged(self):
        v = C.c_int()
        self._chk(self.l.hg_fluid_is_converged(self.h, C.byref(v)))
        return bool(v.value)

    def fluid_convergence_indicator(self):
        v = C.c_double()
        self._chk(self.l.hg_fluid_convergence_indicator(self.h, C.byref(v)))
        return v.value

    def fluid_auto_time_step(self):
        v = C.c_double()
        self._chk(self.l.hg_fluid_auto_time_step(self.h, C.byref(v)))
        return v.value

    def advection_step(self):
        self._chk(self.l.hg_advection_step(self.h))

    def heat_step(self):
        self._chk(self.l.hg_heat_step(self.h))

    def update_properties(self):
        self._chk(self.l.hg_update_properties(self.h))

    def calc_stat(self):
        st = HgStepStats()
        self._chk(self.l.hg_calc_stat(self.h, C.byref(st)))
        return st

    def get_stats(self):
        st = HgStepStats()
        self._chk(self.l.hg_get_stats(self.h, C.byref(st)))
        return st

    def solver_kernel_name(self, which=0):
        return self.l.hg_solver_kernel_name(self.h, which).decode()

    # -- kernel-level entries ---------------------------------------------------------
    def interp_grad(self, u, cond, comp=0):
        u = np.ascontiguousarray(u, dtype=np.float64)
        g = [np.zeros(self.nc) for _ in range(3)]
        self._chk(self.l.hg_interp_grad(self.h, _dptr(u), cond, comp, _dptr(g[0]), _dptr(g[1]), _dptr(g[2])))
        return g[:self.dim]

    def linear_solve(self, solver, coeffs, rhs, tol=0.0, limit=100, relax=1.9):
        sid = LINEAR_SOLVERS[solver] if isinstance(solver, str) else solver
        cs = [np.ascontiguousarray(c, dtype=np.float64) if c is not None else None for c in coeffs]
        arr = (C.POINTER(C.c_double) * 7)(*[_dptr(c) if c is not None else None for c in cs])
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros(self.nc)
        it = C.c_int()
        df = C.c_double()
        self._chk(self.l.hg_linear_solve(self.h, sid, arr, _dptr(rhs), _dptr(x), tol, limit, relax,
                                         C.byref(it), C.byref(df)))
        return x, it.value, df.value

    def smooth_field(self, u, repeat):
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.zeros(self.nc)
        self._chk(self.l.hg_smooth_field(self.h, _dptr(u), repeat, _dptr(out)))
        return out

    # -- instrumentation ---------------------------------------------------------------
    def timers_enable(self, on=True):
        self._chk(self.l.hg_timers_enable(self.h, int(on)))

    def timers(self):
        names = ((C.c_char * 64) * 64)()
        secs = (C.c_double * 64)()
        n = C.c_int()
        self._chk(self.l.hg_timers(self.h, C.cast(names, C.c_void_p), secs, 64, C.byref(n)))
        return {names[i].value.decode(): secs[i] for i in range(n.value)}

    def event_record(self, slot):
        self._chk(self.l.hg_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        v = C.c_double()
        self._chk(self.l.hg_event_elapsed_ms(self.h, a, b, C.byref(v)))
        return v.value

    def profile_enable(self, on=True):
        self._chk(self.l.hg_profile_enable(self.h, int(on)))

    def profile_read(self, which):
        n = C.c_int()
        ms = C.c_double()
        self._chk(self.l.hg_profile_read(self.h, which, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def profile_read_clocks(self):
        out = (C.c_ulonglong * 16)()
        self._chk(self.l.hg_profile_read_clocks(self.h, out))
        return list(out)

    def get_into(self, name, buf):
        """hg_get_field into a caller-provided (e.g. pinned) float64 buffer."""
        fid = F[name] if isinstance(name, str) else name
        self._chk(self.l.hg_get_field(self.h, fid, C.cast(buf.ctypes.data if hasattr(buf, "ctypes") else buf, C.POINTER(C.c_double)), self.nf if fid in FACE_FIELDS else self.nc))

    def set_from(self, name, ptr):
        """hg_set_field from a raw host pointer (int) of nc/nf doubles (e.g. pinned torch tensor)."""
        fid = F[name] if isinstance(name, str) else name
        self._chk(self.l.hg_set_field(self.h, fid, C.cast(ptr, C.POINTER(C.c_double)), self.nf if fid in FACE_FIELDS else self.nc))

    def get_to(self, name, ptr):
        fid = F[name] if isinstance(name, str) else name
        self._chk(self.l.hg_get_field(self.h, fid, C.cast(ptr, C.POINTER(C.c_double)), self.nf if fid in FACE_FIELDS else self.nc))

    def set_from_async(self, name, ptr):
        """hg_set_field_async from a raw PINNED host pointer; complete after synchronize()."""
        fid = F[name] if isinstance(name, str) else name
        self._chk(self.l.hg_set_field_async(self.h, fid, C.cast(ptr, C.POINTER(C.c_double)), self.nf if fid in FACE_FIELDS else self.nc))

    def get_to_async(self, name, ptr):
        """hg_get_field_async into a raw PINNED host pointer; valid after synchronize()."""
        fid = F[name] if isinstance(name, str) else name
        self._chk(self.l.hg_get_field_async(self.h, fid, C.cast(ptr, C.POINTER(C.c_double)), self.nf if fid in FACE_FIELDS else self.nc))

    def launch_count(self):
        return int(self.l.hg_launch_count(self.h))

    def synchronize(self):
        self._chk(self.l.hg_device_synchronize(self.h))
