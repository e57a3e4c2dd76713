#!/usr/bin/env python
"""bench.py -- cell-updates/s per time step (pressure solve included) of hydro's hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size n]

Workload (BASELINE.json configs[3]): Rayleigh-Taylor 3-D, `examples/rt` parameters with
MODULE hydro3d on a unit cube, n^3 cells (default 256^3), fixed work per step: 3 SIMPLE iterations x
(lu momentum solve + 101 Gauss-Seidel sweeps) + 1 advection sub-step + properties + statistics.
A "step" is one hydro<Mesh>::step() (hydro2d.hpp:1531-1621) through the C ABI.

impl b200      : the CUDA path (libhydro_gpu.so).  `value` = whole-job cell-updates/s with the state resident
                 in HBM; `e2e` = same through hg_set_field/hg_step/hg_get_field with pinned HOST buffers.
impl reference : the UNMODIFIED reference binary (oracle/_ref/hydro, OpenMP, all host threads) on a bounded
                 sample of the same workload (--cpu-size^3 cells), or the oracle port if it was not built.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "cell-updates/s per timestep incl. pressure solve"
UNIT = "cell-updates/s"


def workload_params(n, workload="rt", as_configured=False):
    """rt: W4 Rayleigh-Taylor 3-D n^3, fixed work (3 SIMPLE x 101 sweeps) or as configured (<= 20 SIMPLE iterations, tol 1e-4,
    Gauss-Seidel <= 1000 sweeps, tol 1e-5: SURVEY 8d).  dam: W5 examples/broken_dam_3d at n^3 (domain 3.2 x 1 x 1, obstacle
    box, 5 SIMPLE iterations, 10 advection sub-steps) with the sweeps fixed at 101 per solve."""
    import cases
    if workload == "dam":
        return cases.broken_dam_3d(n, n, n, lu_relaxed_num_iters_limit=100, lu_relaxed_tolerance=0.0, convergence_tolerance=0.0)
    if as_configured:
        return cases.rt3d(n, fixed_work=False, lu_relaxed_num_iters_limit=1000, lu_relaxed_tolerance=1e-5,
                          num_iterations_limit=20, convergence_tolerance=1e-4)
    return cases.rt3d(n, fixed_work=True)


def weak_mesh(n, world, shape="column"):
    """Mesh of the weak-scaling run, n^3 cells per GPU.  column (default): n x n x (n world) -- every z-slab is the single-GPU
    n^3 block, so the boxes of the sweep dataflow keep their height (measured on 4 GPUs: 6.7e8 cell-updates/s against 4.8e8
    for the cubic mesh, whose slabs are 64 planes high); the serial part, the wavefront length nx + ny + nz of the lu solve,
    grows with it.  cubic: y, x, z are doubled in turn (2: n x 2n x n, 4: 2n x 2n x n, 8: 2n x 2n x 2n)."""
    if shape == "column":
        return (n, n, n * world)
    m = [n, n, n]
    w, order, q = world, (1, 0, 2), 0
    while w > 1 and w % 2 == 0:
        m[order[q % 3]] *= 2
        w //= 2
        q += 1
    m[2] *= w
    return tuple(m)


def bytes_per_cell_step(n_simple, sweeps_per_solve, n_adv, heat):
    """Algorithmic HBM bytes per cell per step, SURVEY.md 8(d) / BASELINE.md section 3."""
    return n_simple * (664 + 32 * sweeps_per_solve) + 96 * n_adv + (240 if heat else 0) + 120


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_model():
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                return l.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_reference_run(n, steps, warmup, threads=None):
    """The reference's own CPU implementation on the box's host cores: returns (cell-updates/s, info)."""
    import refrun
    p = workload_params(n)
    cells = n ** 3
    cores = threads or os.cpu_count() or 1
    if os.access(refrun.REF_HYDRO, os.X_OK):
        t_all, timers, _ = refrun.run_reference_binary(p, steps + warmup, threads=cores)
        use = t_all[warmup:] if len(t_all) > warmup else t_all
        sec = sum(use) / len(use)
        info = {"kind": "reference", "cores": cores,
                "sample": "RT-3D %d^3 fixed-work, %d timed steps after %d warm-up, unmodified reference binary "
                          "(oracle/_ref/hydro, OpenMP %d threads; its linear solvers and advection are serial), "
                          "per-step t_all from its log" % (n, len(use), warmup, cores),
                "pressure_solve_share": (timers.get("fluid.6.pressure-solve", 0.) / timers["step"]) if timers.get("step") else None}
    else:
        from oracle_api import Oracle
        o = Oracle(p, fast=True)
        for _ in range(warmup):
            o.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            o.step()
        sec = (time.perf_counter() - t0) / steps
        info = {"kind": "port", "cores": 1,
                "sample": "RT-3D %d^3 fixed-work, %d timed steps after %d warm-up, oracle/hydro_oracle.c "
                          "(serial C restatement, -O3)" % (n, steps, warmup)}
    return cells / sec, sec, info


def run_reference(args, rank, world):
    if rank != 0:
        return
    n = args.cpu_size
    val, sec, info = cpu_reference_run(n, args.steps, max(args.warmup, 1))
    info["value"] = val
    info["unit"] = UNIT
    info["cpu_model"] = cpu_model()
    # OMP_NUM_THREADS 1 and all cores, 64^3 and 128^3 (bounded: 2 timed steps each)
    extra = []
    for (en, ethreads) in ((n, 1), (128, None)):
        try:
            v, sec2, inf2 = cpu_reference_run(en, 2, 1, threads=ethreads)
            extra.append({"size": en, "cores": inf2["cores"], "value": v, "ms_per_step": sec2 * 1e3, "kind": inf2["kind"]})
        except Exception as e:   # a report, never a gate
            extra.append({"size": en, "error": str(e)[:160]})
    info["extra"] = extra
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "RT-3D fixed-work (3 SIMPLE x 101 GS sweeps), bounded sample %d^3 on host cores" % n,
                       "cells": n ** 3, "parallelism": "OpenMP x%d" % info["cores"]},
            "cpu_baseline": info,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def slab_parity_check(dist, rank, world, local_rank, n=64, nsteps=2):
    """One process per GPU, CUDA IPC, NVLink peer stores -- the path the scaling numbers are measured on -- against the
    single-GPU run of the same case on rank 0: the gathered slab fields must be bit-identical."""
    import numpy as np
    import torch
    import cases
    from hydro_b200 import parallel
    from hydro_b200.capi import Hydro
    from hydro_b200.config import F, FACE_FIELDS
    p = cases.rt3d(n, Nx=n, Ny=n - 16, Nz=n)
    names = ["VELOCITY_X", "VELOCITY_Y", "VELOCITY_Z", "PRESSURE", "VOLUME_FLUX", "PARTIAL_DENSITY_1"]
    h = Hydro(p, device=local_rank, world_size=world, rank=rank)
    h.link_ipc(dist)
    st = None
    for _ in range(nsteps):
        st = h.step()
    parts = {}
    for nm in names:
        mine = torch.from_numpy(h.get(nm)).cuda()
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine.numel()], dtype=torch.int64, device="cuda"))
        mx = max(int(x.item()) for x in sizes)
        buf = torch.zeros(mx, dtype=torch.float64, device="cuda")
        buf[:mine.numel()] = mine
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf)
        parts[nm] = [o[:int(sz.item())].cpu().numpy() for o, sz in zip(outs, sizes)]
    h.close()
    if rank != 0:
        return None
    one = Hydro(p, device=local_rank)
    s1 = None
    for _ in range(nsteps):
        s1 = one.step()
    bad = []
    for nm in names:
        whole = (parallel.join_faces(parts[nm], n, n - 16, n, world) if F[nm] in FACE_FIELDS else parallel.join_cells(parts[nm]))
        if not np.array_equal(whole, one.get(nm)):
            bad.append(nm)
    if st.pressure_sweeps_total != s1.pressure_sweeps_total or st.convergence_indicator != s1.convergence_indicator:
        bad.append("counts")
    one.close()
    return "bit-exact (%dx%dx%d, %d steps, %d slabs over IPC vs 1 GPU)" % (n, n - 16, n, nsteps, world) if not bad else "MISMATCH: " + ",".join(bad)


def run_b200(args, rank, local_rank, world):
    import numpy as np
    import torch
    from hydro_b200.capi import Hydro
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    n = args.size
    p = workload_params(n, args.workload)
    strong = args.scaling == "strong"
    parity = None
    if world > 1:
        parity = slab_parity_check(dist, rank, world, local_rank, n=args.parity_size)
        if args.parity_only:
            if rank == 0:
                print(json.dumps({"parity_check": parity, "n_gpus": world}))
            dist.destroy_process_group()
            return
        if strong:
            # strong scaling: the n^3 mesh of the single-GPU run, cut into z-slabs
            mesh = (n, n, n)
        else:
            # weak scaling: n^3 cells per GPU, same cell size; z-slabs of a mesh that is kept as cubic as possible
            # (the lexicographic sweeps are a wavefront over i+j+k: its length, nx+ny+nz, is the serial part)
            mesh = weak_mesh(n, world, args.weak_mesh)
            fx, fy, fz = mesh[0] / n, mesh[1] / n, mesh[2] / n
            B, B1 = p["B"], p["B1"]
            p.update(Nx=mesh[0], Ny=mesh[1], Nz=mesh[2], B=(B[0] * fx, B[1] * fy, B[2] * fz), B1=(B1[0] * fx, B1[1] * fy, B1[2] * fz))
        h = Hydro(p, device=local_rank, world_size=world, rank=rank)
        h.link_ipc(dist)
    else:
        mesh = (n, n, n)
        h = Hydro(p, device=local_rank)
    cells = h.nc
    total_cells = mesh[0] * mesh[1] * mesh[2]
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        h.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(W):
        st = h.step()
    if os.environ.get("HYDRO_GT_CLOCK"):
        h.profile_read_clocks()
    n_simple, sweeps, n_adv = st.simple_iterations, st.pressure_sweeps_total, st.advection_substeps
    # ---- device-resident timing
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    h.profile_enable(True)
    h.profile_read(0), h.profile_read(1)
    barrier()
    launches0 = h.launch_count()
    h.event_record(0)
    for _ in range(K):
        st = h.step()
    h.event_record(1)
    barrier()
    ms = allmax(h.event_elapsed_ms(0, 1))
    launches = h.launch_count() - launches0
    gs_n, gs_ms = h.profile_read(0)
    lu_n, lu_ms = h.profile_read(1)
    h.profile_enable(False)
    if os.environ.get("HYDRO_BENCH_TIMERS") and rank == 0:
        # per-section device times of three more steps (events around every section: serialising, not part of any number)
        h.timers_enable(True)
        for _ in range(3):
            h.step()
        print("timers (s per 3 steps):", {k: round(v, 5) for k, v in sorted(h.timers().items())}, file=sys.stderr)
        h.timers_enable(False)
    elif os.environ.get("HYDRO_BENCH_TIMERS"):
        for _ in range(3):
            h.step()
    if world > 1 and os.environ.get("HYDRO_BENCH_RANKS"):
        print("rank %d: step %.2f ms, gs %.2f ms x %d, lu %.2f ms x %d" % (rank, h.event_elapsed_ms(0, 1) / K, gs_ms / max(gs_n, 1), gs_n, lu_ms / max(lu_n, 1), lu_n), file=sys.stderr)
    if os.environ.get("HYDRO_GT_CLOCK"):
        # instrumented build (-DGT_CLOCK): cycles per warp role, summed over warps: producer store+publish / poll+load /
        # barrier wait; sweep warps barrier wait / step
        # [0] sweep warps: steps, [1] their barrier waits, [2] interface waits (slabs), [3] producers: dependency polls, [4] loads (incl.
        # ghost planes), [5] barrier waits, [6] task time (one thread per task), [7] tasks
        print("gt_clocks rank %d" % rank, h.profile_read_clocks()[:8], "launches", gs_n, "gs_ms", gs_ms / max(gs_n, 1), file=sys.stderr)
    clocks = sampler.stop() if sampler else None
    sec_step = ms * 1e-3 / K
    value = total_cells / sec_step

    # ---- end to end through the C ABI with pinned host buffers (SURVEY 8b secondary boundary: the
    # fields FluidSimple reads from caller memory every iteration, hydro2d.hpp:449-463, are uploaded each step;
    # velocity + pressure are downloaded each step, as write_results / CalcStat consume them)
    ins = ["DENSITY", "VISCOSITY", "FORCE_X", "FORCE_Y", "FORCE_Z"]
    outs = ["VELOCITY_X", "VELOCITY_Y", "VELOCITY_Z", "PRESSURE"]
    pin_in = {k: torch.empty(cells, dtype=torch.float64).pin_memory() for k in ins}
    pin_out = {k: torch.empty(cells, dtype=torch.float64).pin_memory() for k in outs}
    for k in ins:
        h.get_to(k, pin_in[k].data_ptr())
    Ke = max(10, K)
    for k in ins:    # untimed: first use allocates the staging / snapshot buffers and the copy streams
        h.set_from_async(k, pin_in[k].data_ptr())
    for k in outs:
        h.get_to_async(k, pin_out[k].data_ptr())
    barrier()
    t0 = time.perf_counter()
    # Every step uploads its five input fields from pinned host memory and downloads its four result fields.  The uploads go
    # through a copy stream into staging buffers and are applied in stream order before the step that uses them; the step is
    # queued with hg_step_begin, so the upload of step n+1 is queued while step n computes (hg_step_end then waits for step n's
    # status block); the downloads copy a snapshot taken after the step and overlap the next step.  Ke steps, Ke uploads of
    # every input field, Ke downloads of every output field inside the timed region.
    for k in ins:
        h.set_from_async(k, pin_in[k].data_ptr())
    h.step_begin()
    for it in range(Ke):
        if it + 1 < Ke:
            for k in ins:
                h.set_from_async(k, pin_in[k].data_ptr())
        st2 = h.step_end()
        for k in outs:
            h.get_to_async(k, pin_out[k].data_ptr())
        if it + 1 < Ke:
            h.step_begin()
    barrier()   # hg_device_synchronize waits for the compute and both copy streams
    e2e_sec = allmax((time.perf_counter() - t0) / Ke)
    e2e = {"value": total_cells / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": len(ins) * cells * 8,
           "d2h_bytes_per_step": len(outs) * cells * 8 + 8 * 40, "ms_per_step": e2e_sec * 1e3, "steps": Ke,
           "timing": "host wall clock around pinned H2D (hg_set_field_async) + hg_step_begin/hg_step_end + D2H (hg_get_field_async) per step, the "
                     "upload of step n+1 queued while step n computes, all streams synchronised at the end, max over ranks"}
    kname = h.solver_kernel_name(0)
    lu_name = h.solver_kernel_name(1)

    # ---- the same mesh as configured (SURVEY 8d second accounting mode): <= 20 SIMPLE iterations (tol 1e-4), Gauss-Seidel
    # <= 1000 sweeps with tol 1e-5 -- the sweep chunks are checked on the host, an overshooting chunk is replayed
    asconf = None
    if world == 1 and args.workload == "rt" and not args.no_as_configured:
        h.close()
        del pin_in, pin_out
        ha = Hydro(workload_params(n, "rt", as_configured=True), device=local_rank)
        ha.step()
        ha.synchronize()
        ha.profile_enable(True)
        ha.profile_read(0), ha.profile_read(1)
        ha.event_record(0)
        sts = [ha.step() for _ in range(5)]
        ha.event_record(1)
        ha.synchronize()
        a_gs_n, a_gs_ms = ha.profile_read(0)
        a_lu_n, a_lu_ms = ha.profile_read(1)
        ha.profile_enable(False)
        ams = ha.event_elapsed_ms(0, 1) / len(sts)
        sim = sum(s_.simple_iterations for s_ in sts) / len(sts)
        swp = sum(s_.pressure_sweeps_total for s_ in sts) / len(sts)
        bp = 664. * sim + 32. * swp + 96. * sts[-1].advection_substeps + 120.
        asconf = {"workload": "RT-3D %d^3 as configured: <= 20 SIMPLE iterations (tol 1e-4), gauss_seidel <= 1000 sweeps (tol 1e-5)" % n,
                  "ms_per_step": ams, "value": cells / (ams * 1e-3), "unit": UNIT, "steps": len(sts),
                  "simple_iterations_per_step": sim, "pressure_sweeps_per_step": swp,
                  # sweep launches incl. the replayed chunks: (kernel time) / (sweeps the reference executes x time per sweep
                  # of the fixed-work launch) is the price of finding the stopping sweep after the fact
                  "sweep_launches_per_step": a_gs_n / len(sts), "sweep_kernel_ms_per_step": a_gs_ms / len(sts),
                  "lu_ms_per_step": a_lu_ms / len(sts),
                  "algorithmic_bytes_per_cell_step": bp, "frac_of_peak": bp * cells / (ams * 1e-3) / 1e9 / measured_peak()[0]}
        ha.close()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    # roofline of the dominant kernel: the Gauss-Seidel sweep kernel (time-skewed column boxes; on slabs the boxes of
    # neighbouring GPUs exchange tagged interface values), one launch per pressure solve = (limit+1) sweeps x 32
    # algorithmic bytes per cell-sweep
    sweeps_per_solve = p["lu_relaxed_num_iters_limit"] + 1
    gs_bytes = 32.0 * sweeps_per_solve * cells
    gs_avg_ms = gs_ms / gs_n if gs_n else float("nan")
    achieved = gs_bytes / (gs_avg_ms * 1e-3) / 1e9 if gs_n else None
    bpcs = bytes_per_cell_step(n_simple, sweeps_per_solve, n_adv, False)
    step_gbs = bpcs * total_cells / world / sec_step / 1e9
    roof = {"bound": "hbm", "kernel": "%s (lexicographic Gauss-Seidel/SOR as time-skewed column boxes, rows of a hyperplane staged once per "
                                      "box by TMA and shared by the sweeps in flight, %d sweeps per launch)" % (kname, sweeps_per_solve),
            "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
            "frac": achieved / peak if achieved else None, "traffic": None,
            "launches_timed": gs_n, "avg_launch_ms": gs_avg_ms, "share_of_step": gs_ms / ms if ms else None,
            "lu_kernel": lu_name, "lu_kernel_share_of_step": lu_ms / ms if ms else None,
            "lu_avg_solve_ms": lu_ms / lu_n if lu_n else None,
            "whole_step": {"algorithmic_bytes_per_cell_step": bpcs, "achieved_per_gpu": step_gbs, "frac": step_gbs / peak}}
    traffic_file = os.path.join(ROOT, "profiles", "gs_traffic.json")
    if os.path.exists(traffic_file):
        try:
            if world == 1 and args.workload == "rt":
                roof["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch_scaled_to", {}).get(str(n))
        except Exception:
            pass
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, sec, info = cpu_reference_run(args.cpu_size, 5, 1)
            info.update({"value": v, "unit": UNIT, "ms_per_step": sec * 1e3, "cpu_model": cpu_model()})
            cpu = info
        except Exception as e:  # the baseline is a report, never a gate
            cpu = {"error": str(e)[:200]}
    wname = ("RT-3D %d^3 fixed-work: 3 SIMPLE iterations x (lu momentum + 101 GS sweeps) + 1 advection sub-step + properties + "
             "stats per step (SURVEY 8d W4)" % n) if args.workload == "rt" else \
            ("broken_dam_3d %d^3 (SURVEY 8d W5: obstacle box, 5 SIMPLE iterations x (lu momentum + 101 GS sweeps), 10 advection "
             "sub-steps, smoothing 2/3/3)" % n)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": sec_step * 1e3, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "mesh": list(mesh),
                       "cells_per_gpu": cells, "simple_iterations": n_simple, "pressure_sweeps_per_step": sweeps,
                       "advection_substeps": n_adv,
                       "parallelism": "1 GPU" if world == 1 else
                       "z-slabs over %d GPUs, one process per GPU: %dx%dx%d cells, halo planes / solver interface values / "
                       "reductions over NVLink peer memory (%s scaling)" % ((world,) + tuple(mesh) + ("strong" if strong else "weak, %d^3 cells per GPU" % n,)),
                       "l2": "working set %.1f GB per GPU >> 126 MB L2 (inputs larger than L2, no flush needed)" % (cells * 8 * 90 / 1e9)},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
    if parity is not None:
        line["parity_check"] = parity
    if asconf is not None:
        line["as_configured"] = asconf
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--cpu-size", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-as-configured", action="store_true")
    ap.add_argument("--workload", default="rt", choices=["rt", "dam"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--weak-mesh", default="column", choices=["cubic", "column"])
    ap.add_argument("--parity-only", action="store_true", help="N > 1: only the slab-vs-single-GPU bit-exactness check over CUDA IPC")
    ap.add_argument("--parity-size", type=int, default=64)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
