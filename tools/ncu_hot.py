#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu source page CSV: instructions executed and stall samples, top regions.
   tools/ncu_hot.py /tmp/src.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_inst = sum(float(r[ix["Instructions Executed"]] or 0) for r in data)
tot_samp = sum(float(r[ix["# Samples"]] or 0) for r in data)
print("total warp instructions %.3e, samples %d, SASS lines %d" % (tot_inst, tot_samp, len(data)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# contiguous regions with similar execution counts
print("top by samples:")
for r in sorted(data, key=lambda r: -float(r[ix["# Samples"]] or 0))[:n]:
    st = {k[6:]: int(float(r[ix[k]] or 0)) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and float(r[ix[k]] or 0) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%6s %-70s samp %6d (%.1f%%) inst %.2e %s" % (r[ix["Address"]][-5:], r[ix["Source"]][:70], float(r[ix["# Samples"]] or 0),
          100 * float(r[ix["# Samples"]] or 0) / tot_samp, float(r[ix["Instructions Executed"]] or 0), top))
