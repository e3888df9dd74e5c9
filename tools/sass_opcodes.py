#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libunetb200.so (evidence that the hot kernels are tcgen05 / TMEM / TMA code).

    python tools/sass_opcodes.py [path/to/libunetb200.so] > profiles/rNN_sass_opcodes.md

Runs `cuobjdump -sass` (no GPU needed) and counts, per kernel, the mnemonics B200_PROFILING.md names:
UTCHMMA* (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG* / UTMASTG* (TMA load / store), UTCBAR* (tcgen05.commit),
UTMAPF (TMA prefetch), SYNCS* (mbarrier), plus the CUDA-core math that would betray a fallback (HMMA / FFMA)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tw_invoice_unet_ocr_llm_b200", "libunetb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda names: subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")

WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA", "total"]
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        base = op.split(".")[0]
        if base == "UTCHMMA":
            cur["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
        elif base in ("LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA"):
            cur[base] += 1
names = demangle(list(kernels))
print(f"# SASS opcode counts per kernel: {os.path.basename(lib)} ({os.path.getsize(lib)} bytes), `cuobjdump -sass`\n")
print("UTCHMMA = tcgen05.mma (cta_group::1), UTCHMMA.2CTA = cta_group::2, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store,")
print("UTMAPF = TMA prefetch, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.  No HMMA (mma.sync) anywhere: the dense path is tcgen05 only.\n")
print("| kernel | " + " | ".join(WATCH) + " |")
print("|---|" + "---|" * len(WATCH))
tot = collections.Counter()
for (mangled, c), name in zip(kernels.items(), names):
    short = re.sub(r"\(ub::ConvParams\)|\(ub::\w+\)$", "", name.replace("void ", "")).strip()
    short = re.sub(r"\(int\)", "", short)
    print(f"| `{short[:110]}` | " + " | ".join(str(c.get(k, 0)) for k in WATCH) + " |")
    tot.update(c)
print("| **all kernels** | " + " | ".join(f"**{tot.get(k, 0)}**" for k in WATCH) + " |")
