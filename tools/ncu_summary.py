#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export into a per-launch table (markdown).

usage: tools/ncu_summary.py gpurun_out/prof_<tag>_raw.csv [layer names json] > profiles/<tag>_ncu_summary.md
"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
    ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.sum.pct_of_peak_sustained_elapsed", "bf16 MMA ops % of peak"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("smsp__cycles_active.avg", "cycles"),
]

LAYERS = ["down1.net.0 (stem)", "down1.net.3", "down2.net.0", "down2.net.3", "down3.net.0", "down3.net.3",
          "down4.net.0", "down4.net.3", "bottleneck.net.0", "bottleneck.net.3", "up4", "conv4.net.0",
          "conv4.net.3", "up3", "conv3.net.0", "conv3.net.3", "up2", "conv2.net.0", "conv2.net.3", "up1",
          "conv1.net.0", "conv1.net.3 + out_conv"]
# the default plan folds every up-conv into the following conv (csrc/conv_phase*.cuh): 18 launches
LAYERS_FOLDED = [l if not l.endswith(".net.0") or not l.startswith("conv") else l + " + up" + l[4] + " (folded)"
                 for l in LAYERS if not l.startswith("up")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(c, n) for c, n in COLS if c in idx]
    print("| # | layer | kernel | " + " | ".join(f"{n} [{units[idx[c]]}]" for c, n in cols) + " |")
    print("|---|---|---|" + "---|" * len(cols))
    # the capture window may start mid-forward: anchor the labels on the stem kernel
    names = [r[idx["Kernel Name"]] for r in data]
    # stem = the A_STEM instantiation conv_tc_kernel<64, 1, 3, ...> (or the CUDA-core stem_conv_kernel)
    stem = next((i for i, n in enumerate(names)
                 if "stem" in n or "conv_tc_kernel<(int)64, (int)1, (int)3" in n or "conv_tc_kernel<64, 1, 3" in n), 0)
    layers = LAYERS_FOLDED if any("conv_phase" in n for n in names) else LAYERS
    for i, r in enumerate(data):
        name = names[i]
        short = name.split("(")[0].replace("ub::", "").replace("void ", "")[:60]
        layer = layers[(i - stem) % len(layers)]
        print(f"| {i} | {layer} | `{short}` | " + " | ".join(r[idx[c]] for c, _ in cols) + " |")


if __name__ == "__main__":
    main()
