#!/bin/bash
# ncu evidence for the crop-enhancement kernels (csrc/enhance.cuh); usage: tools/profile_enhance.sh <tag>
tag=${1:-r01d}
out=gpurun_out
mkdir -p $out /tmp/ncu
CMD="python tools/enhance_bench.py --quick"
$CMD > $out/enh_plain_$tag.json 2> $out/enh_plain_$tag.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:enh_ -s 300 -c 25 --csv \
    --log-file $out/launches_enh_$tag.csv $CMD > $out/ncu_enh_launches_$tag.log 2>&1
echo "launch list rc=$?"
# one batch call (5 launches) of the 192-crop batch with the full set; the single-crop loop launches
# 60 x 5 kernels first, the e2e batch loop 13 x 5 more -> skip well into the device-only loop
ncu --set full --clock-control none --import-source on -k regex:enh_ -s 400 -c 5 \
    -o $out/enh_full_$tag -f $CMD > $out/ncu_enh_full_$tag.log 2>&1
echo "full set rc=$?"
ncu -i $out/enh_full_$tag.ncu-rep --page raw --csv > $out/enh_full_${tag}_raw.csv 2> /dev/null
ls -la $out | grep enh
