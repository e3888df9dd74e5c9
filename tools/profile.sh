#!/bin/bash
# ncu evidence for the bench command (see /opt/skills/guides/B200_PROFILING.md).
# usage: tools/profile.sh <tag> [launches per forward, default 18]
# Brings back only small artefacts (gpurun_out is capped at 64 MiB): CSV exports of the reports.
tag=${1:-r02}
L=${2:-18}
out=gpurun_out
mkdir -p $out /tmp/ncu
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
# L launches per forward; 3 warm-up forwards + 1 profile-priming -> skip 4 L, list two forwards
$CMD > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s $((4 * L)) -c $((2 * L)) --csv \
    --log-file $out/launches_$tag.csv $CMD > $out/ncu_launches_$tag.log 2>&1
echo "launch list rc=$?"
# one full forward with the full metric set -> raw CSV only
ncu --set full --clock-control none -s $((4 * L)) -c $L -o /tmp/ncu/all -f $CMD > $out/ncu_all_$tag.log 2>&1
echo "full set rc=$?"
ncu -i /tmp/ncu/all.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2> /dev/null
# source-level capture of the folded kernels of levels 2 and 1 (conv2.net.0: conv_phase_multi, conv1.net.0: conv_phase_stack64), 5th forward
for spec in pm0:conv_phase_multi pm1:conv_phase_stack64; do
  IFS=: read -r j rx <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 1 \
      -o /tmp/ncu/$j -f $CMD > $out/ncu_${j}_$tag.log 2>&1
  ncu -i /tmp/ncu/$j.ncu-rep --page source --csv > $out/src_${j}_$tag.csv 2>/dev/null
done
echo "source rc=$?"
ls -la $out/ | tail -n 12
tail -n 1 $out/plain_$tag.log | cut -c1-300
