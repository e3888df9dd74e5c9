#!/usr/bin/env python
"""A few key metrics of each launch in an `ncu --page raw --csv` export.  usage: ncu_keys.py file.csv"""
import csv, sys
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__cycles_active.avg", "launch__registers_per_thread"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:60])
    for k in KEYS:
        if k in hdr: print("   ", k, r[hdr.index(k)])
