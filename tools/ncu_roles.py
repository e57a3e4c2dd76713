#!/usr/bin/env python
"""Per warp role view of k_gs_tiled from an ncu source page (ncu -i x.ncu-rep --page source --csv --print-source sass > x.csv):
SASS lines are attributed to a role by how often they were executed (sweep warps: once per warp and step pair, i.e. ~8.6e6-9.3e6
times at 256^3 / 101 sweeps; lines both producer warps run: ~2.3e6; lines one producer warp runs: ~1.2e6).
   tools/ncu_roles.py plain.csv [link.csv]"""
import csv, sys, collections, re

def load(fn):
    rows = list(csv.reader(open(fn))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    return hdr, ix, rows[2:]

ROLES = (("sweep warps (8)", 5e6, 1e12), ("both producer warps + publisher", 1.8e6, 5e6), ("one producer warp", 0.8e6, 1.8e6))

def table(fn):
    hdr, ix, data = load(fn)
    out = []
    steps = None
    for name, lo, hi in ROLES:
        sel = [r for r in data if lo < float(r[ix["Instructions Executed"]] or 0) <= hi]
        inst = sum(float(r[ix["Instructions Executed"]] or 0) for r in sel)
        samp = sum(float(r[ix["# Samples"]] or 0) for r in sel)
        st = collections.Counter()
        for r in sel:
            for k in hdr:
                if k.startswith("stall_") and "Not Issued" not in k:
                    st[k[6:]] += float(r[ix[k]] or 0)
        ops = collections.Counter()
        for r in sel:
            s = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
            ops[s.split()[0].split(".")[0]] += 1
        out.append((name, len(sel), inst, samp, st, ops))
    return out

def show(title, t):
    print("### %s\n" % title)
    print("| role | SASS lines executed per step pair | warp instructions | stall samples | top stall reasons (samples) |")
    print("|---|---:|---:|---:|---|")
    for name, n, inst, samp, st, ops in t:
        top = ", ".join("%s %d" % (k, v) for k, v in st.most_common(5))
        print("| %s | %d | %.3e | %d | %s |" % (name, n, inst, samp, top))
    name, n, inst, samp, st, ops = t[0]
    fp = ops["DADD"] + ops["DMUL"] + ops["DFMA"] + ops["DSETP"] + ops["MUFU"]
    print("\nSweep-warp step pair (two steps, four updates each): %d lines = %d fp64 (DADD %d, DMUL %d, DFMA %d, DSETP %d, MUFU %d), "
          "%d LDS, %d STS, %d STG, %d FSEL, %d ISETP, %d other.\n" % (n, fp, ops["DADD"], ops["DMUL"], ops["DFMA"], ops["DSETP"], ops["MUFU"],
          ops["LDS"], ops["STS"], ops["STG"], ops["FSEL"], ops["ISETP"], n - fp - ops["LDS"] - ops["STS"] - ops["STG"] - ops["FSEL"] - ops["ISETP"]))

if __name__ == "__main__":
    show("k_gs_tiled<0> (one GPU)", table(sys.argv[1]))
    if len(sys.argv) > 2:
        show("k_gs_tiled<1> (slab instantiation, run on one GPU without neighbours: HYDRO_GT_FORCE_LINK=1)", table(sys.argv[2]))
