#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv` export.  usage: ncu_src_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(rows[0][1][:100])
hdr, data = rows[1], rows[2:]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[i_s] or 0) for r in data)
print("total samples", tot)
for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][i_s] or 0))[:n]:
    st = {h[6:]: int(r[i] or 0) for i, h in stall_cols if int(r[i] or 0) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print(f"{100*int(r[i_s])/tot:5.1f}% line{k:5d} ex={r[i_ex]:>9s} {r[i_src][:64]:64s} {st}")
