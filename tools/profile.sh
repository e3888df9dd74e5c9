#!/bin/bash
# ncu evidence for the bench command (see /opt/skills/guides/B200_PROFILING.md).
# usage: tools/profile.sh <tag>   -> gpurun_out/launches_<tag>.csv, prof_<tag>_all.ncu-rep, prof_<tag>_src.ncu-rep
tag=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 66 -c 44 --csv \
    --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches_$tag.log 2>&1
echo "launch list rc=$?"
# one full forward (22 launches) with the full metric set
ncu --set full --clock-control none -s 66 -c 22 -o gpurun_out/prof_${tag}_all -f $CMD > gpurun_out/ncu_all_$tag.log 2>&1
echo "full set rc=$?"
# source-level capture of the two interesting tensor-core kernels: conv1.net.0 (Cout=64) and conv3.net.0 (Cout=256)
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 63 -c 21 \
    --launch-skip-before-match 0 -o gpurun_out/prof_${tag}_src -f $CMD > gpurun_out/ncu_src_$tag.log 2>&1
echo "source rc=$?"
ls -la gpurun_out/*.ncu-rep
cat gpurun_out/plain_$tag.log | tail -2
