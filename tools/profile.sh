#!/bin/bash
# ncu evidence for the bench command (see /opt/skills/guides/B200_PROFILING.md).
# usage: tools/profile.sh <tag>
# Brings back only small artefacts (gpurun_out is capped at 64 MiB): CSV exports of the
# reports plus one source-level .ncu-rep of two kernels.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out /tmp/ncu
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
# 22 launches per forward; 3 warm-up forwards + 1 profile-priming -> skip 66, list two forwards
$CMD > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 66 -c 44 --csv \
    --log-file $out/launches_$tag.csv $CMD > $out/ncu_launches_$tag.log 2>&1
echo "launch list rc=$?"
# one full forward (22 launches) with the full metric set -> raw CSV only
ncu --set full --clock-control none -s 66 -c 22 -o /tmp/ncu/all -f $CMD > $out/ncu_all_$tag.log 2>&1
echo "full set rc=$?"
ncu -i /tmp/ncu/all.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2> /dev/null
ncu -i /tmp/ncu/all.ncu-rep --page details --csv > $out/prof_${tag}_details.csv 2> /dev/null
# source-level capture of two tensor-core kernels: conv launches #19 (conv1.net.0, Cout=64,
# weight-stationary) and #13 (conv3.net.0, Cout=256) of the 4th forward
ncu --set full --clock-control none --import-source on -k regex:conv_ -s 80 -c 1 \
    -o $out/prof_${tag}_conv3_0 -f $CMD > $out/ncu_src1_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_ -s 86 -c 1 \
    -o $out/prof_${tag}_conv1_0 -f $CMD > $out/ncu_src2_$tag.log 2>&1
echo "source rc=$?"
ls -la $out/ | tail -n 20
du -sh $out
tail -n 1 $out/plain_$tag.log | cut -c1-300
