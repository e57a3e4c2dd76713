#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU): key metrics of the first kernel in the report.
   tools/ncu_summary.py gpurun_out/x.ncu-rep [--json out.json]"""
import csv, io, json, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    def g(k, default=None):
        v = d.get(k)
        if v in (None, ""):
            return default
        try:
            return float(v.replace(",", ""))
        except ValueError:
            return v
    def scale(k):   # bytes with unit prefix
        v = g(k)
        if v is None: return None
        un = u.get(k, "")
        f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(un, 1)
        return v * f
    dur = g("gpu__time_duration.sum"); dun = u.get("gpu__time_duration.sum", "")
    dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(dun, 1) if dur else None
    r = {"kernel": d.get("Kernel Name"), "duration_ms": dur_ms,
         "dram_read": scale("dram__bytes_read.sum"), "dram_write": scale("dram__bytes_write.sum"),
         "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
         "l2_to_sm_read": scale("l1tex__m_xbar2l1tex_read_bytes.sum"),
         "regs": g("launch__registers_per_thread"), "smem_dyn": g("launch__shared_mem_per_block_dynamic"),
         "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
         "issue_active_pct": g("smsp__issue_active.avg.pct"), "inst_executed": g("smsp__inst_executed.sum"),
         "ipc": g("sm__inst_executed.avg.per_cycle_elapsed"),
         "fp64_pipe_pct": g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") or g("smsp__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
         "lsu_pipe_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
         "l1_data_pipe_lsu_pct": g("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
         "smem_wavefronts": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
         "dram_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
    st = {}
    for k in hdr:
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", k) or re.match(r"smsp__average_warp_latency_issue_stalled_(\w+)\.ratio", k)
        if m and d.get(k):
            try: st[m.group(1)] = round(float(d[k]), 3)
            except ValueError: pass
    r["stalls_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:10])
    out.append(r)
print(json.dumps(out, indent=1))
if "--json" in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
