#!/usr/bin/env python
"""Per-source-line instruction / stall shares of one kernel from an .ncu-rep captured with --import-source on.
usage: tools/ncu_src_lines.py report.ncu-rep kernel_regex [N]"""
import csv
import io
import subprocess
import sys


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].strip()]      # source rows only (SASS rows have no line)
    i_ex, i_s = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
    tot, ts = sum(I(r[i_ex]) for r in data), sum(I(r[i_s]) for r in data)
    print(f"warp instructions {tot}, stall samples {ts}")
    for r in sorted(data, key=lambda r: -I(r[i_s]))[:n]:
        st = dict(sorted({h[6:]: I(r[i]) for i, h in stall if I(r[i]) > 0}.items(), key=lambda kv: -kv[1])[:3])
        print(f"{r[0]:>5s} stall {100 * I(r[i_s]) / ts:5.1f}%  instr {100 * I(r[i_ex]) / tot:5.1f}%  {r[1].strip()[:84]:84s} {st}")


if __name__ == "__main__":
    main()
