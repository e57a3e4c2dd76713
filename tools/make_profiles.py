#!/usr/bin/env python
"""Builds the committed round-2 profile artefacts from gpurun_out/ (ncu reports are read here, no GPU needed):
   tools/make_profiles.py <bench tag> <gs ncu-rep> <lu ncu-rep> <fast-kernel ncu-reps...>"""
import csv, io, json, os, re, shutil, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
CELLS = 256 ** 3

def summ(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    return json.loads(out)

# ---- bench lines
for src, dst in ((tag + "_bench.json", "r02_bench_final.json"), (tag + "_reference.json", "r02_bench_reference_arm.json")):
    line = [l for l in open(os.path.join(G, src)).read().splitlines() if l.startswith("{")][-1]
    open(os.path.join(P, dst), "w").write(line + "\n")
# ---- launch list
shutil.copy(os.path.join(G, tag + "_launches.csv"), os.path.join(P, "r02_launches.csv"))
rows = [r for r in csv.reader(open(os.path.join(G, tag + "_launches.csv"))) if len(r) > 10]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
L = []
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    L.append((re.sub(r"\(.*", "", r[ki]).replace("void ", ""), v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(r[ui], 1e-6)))
idx = [n for n, (k, v) in enumerate(L) if k.startswith("k_status_pack")]
a, b = idx[-2] + 1, idx[-1] + 1
agg = collections.OrderedDict()
for k, v in L[a:b]:
    agg.setdefault(k, [0, 0.]); agg[k][0] += 1; agg[k][1] += v
tot = sum(v for k, v in L[a:b])
with open(os.path.join(P, "r02_launches_summary.md"), "w") as f:
    f.write("# Round 2 -- ncu launch list of one time step (RT-3D 256^3 fixed work)\n\n"
            "```\nncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/%s_launches.csv \\\n"
            "    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-as-configured\n```\n\n"
            "Raw list: `profiles/r02_launches.csv`.  The table is the last complete step in the list (%d launches, %.2f ms under ncu;\n"
            "the same step takes %.2f ms in `bench.py` without the profiler).  Per-launch times under ncu are cold-cache and serialised:\ncompare SHARES.\n\n"
            "| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n" % (tag, b - a, tot, json.load(open(os.path.join(P, "r02_bench_final.json")))["ms_per_step"]))
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.3f | %.3f | %.1f%% |\n" % (k, n, v, v / n, 100 * v / tot))
# ---- ncu summaries
gs = summ(sys.argv[2])[0]; lu = summ(sys.argv[3])[0]
json.dump(gs, open(os.path.join(P, "r02_gs_tiled_ncu.json"), "w"), indent=1)
json.dump(lu, open(os.path.join(P, "r02_lu_tiled_ncu.json"), "w"), indent=1)
fast = []
for rep in sys.argv[4:]:
    fast += summ(rep)
json.dump(fast, open(os.path.join(P, "r02_fast_kernels_ncu.json"), "w"), indent=1)
gs_alg = 32.0 * 101 * CELLS
json.dump({"kernel": "k_gs_tiled (rows of a hyperplane staged once per box by TMA, round 2)", "workload": "RT-3D 256^3, 101 sweeps per launch",
           "ncu": "--set full --clock-control none, one launch (-s 3 -c 1)", "duration_ms": gs["duration_ms"],
           "dram_bytes_read": gs["dram_read"], "dram_bytes_write": gs["dram_write"], "algorithmic_bytes_per_launch": gs_alg,
           "l2_hit_rate_pct": gs["l2_hit_pct"], "l2_to_sm_read_bytes": gs["l2_to_sm_read"], "warp_instructions": gs["inst_executed"],
           "registers_per_thread": gs["regs"], "dram_bytes_per_launch": gs["dram_read"] + gs["dram_write"],
           "traffic_over_algorithmic": (gs["dram_read"] + gs["dram_write"]) / gs_alg, "stall_cycles_per_issued_instruction": gs["stalls_per_issue"],
           "dram_bytes_per_launch_scaled_to": {"256": gs["dram_read"] + gs["dram_write"]}}, open(os.path.join(P, "gs_traffic.json"), "w"), indent=1)
# ---- per-kernel roofline table of the interior kernels
ALG = {"k_fa_grad": (22, "u(3) force(3) p -> G(9) fcr(3) gp(3)"), "k_fb_momentum": (40, "G(9) u(9: 3 layers) rho mu F(3) gp(3) fcr(3) -> A(7) R(3) dc"),
       "k_fc_flux_rows": (22, "u*(3) gp(3) fcr(3) force(3) p dc -> F*(3) rows(5)"), "k_fd_correct": (14, "p' dc F*(3) u(3) -> u(3) F(3)"),
       "k_fe_advect": (5, "pd F(3) -> pd"), "k_lu_tiled": (13.5, "A(7) R(3) -> X(3), + halo traffic; per direction"), "k_gs_tiled": (4 * 101, "32 B per cell-sweep x 101 sweeps")}
with open(os.path.join(P, "r02_stencil_kernels.md"), "w") as f:
    f.write("# Round 2 -- dominant kernels against their own HBM roofline (ncu --set full, RT-3D 256^3, one launch each)\n\n"
            "algorithmic bytes = (8-byte array passes over the 16 777 216 cells) x 8 B x cells; fraction = algorithmic bytes / duration / %.1f GB/s\n"
            "(MEASURED_PEAKS.json).  Durations are ncu's (cold cache, serialised).\n\n"
            "| kernel | passes | arrays | duration ms | DRAM read + write GB | algorithmic GB | algorithmic GB/s | fraction of peak | DRAM / algorithmic |\n|---|---:|---|---:|---:|---:|---:|---:|---:|\n" % PEAK)
    seen = set()
    for r in [gs, lu] + fast:
        name = re.sub(r"\(.*", "", r["kernel"]).replace("void ", "")
        base = re.sub(r"<.*", "", name)
        if base not in ALG or name in seen: continue
        seen.add(name)
        passes, what = ALG[base]
        alg = passes * 8.0 * CELLS
        dram = r["dram_read"] + r["dram_write"]
        f.write("| `%s` | %s | %s | %.3f | %.2f | %.2f | %.0f | %.2f | %.2f |\n" % (name, passes, what, r["duration_ms"], dram / 1e9, alg / 1e9,
                alg / 1e9 / (r["duration_ms"] * 1e-3), alg / 1e9 / (r["duration_ms"] * 1e-3) / PEAK, dram / alg))
# ---- SASS mnemonics
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "hydro_b200", "libhydro_gpu.so")], capture_output=True, text=True).stdout
cnt = collections.defaultdict(collections.Counter); fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: fn = m.group(1); continue
    m = re.search(r"\s(UTMALDG|UTMAPF|SYNCS|DFMA|DMUL|DADD|LDGSTS|CCTL|MUFU|BAR|LDS|STS|LDG|STG|ATOMG|REDG|SHFL)[.\s]", line)
    if m and fn: cnt[fn][m.group(1)] += 1
with open(os.path.join(P, "r02_sass_grep.txt"), "w") as f:
    f.write("cuobjdump -sass hydro_b200/libhydro_gpu.so : instruction counts per kernel (static).  UTMALDG = TMA tensor load, SYNCS = mbarrier,\n"
            "LDGSTS = cp.async, DFMA/DMUL/DADD = fp64 pipe, MUFU = reciprocal seed of the fp64 division, CCTL = cache control (prefetch / acquire)\n\n")
    for k in sorted(cnt):
        if any(s in k for s in ("k_gs_tiled", "k_lu_tiled", "k_f", "k_gt_", "k_stat", "k_slab", "k_mail")):
            f.write("%-70s %s\n" % (k[:70], " ".join("%s=%d" % kv for kv in sorted(cnt[k].items()))))
print("profiles written")
