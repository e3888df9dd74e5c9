#!/bin/bash
# source-level ncu capture of selected launches of the 4th forward: args = launch indices (0..21)
mkdir -p gpurun_out /tmp/ncu
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_src.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/plain_src.log; exit 1; }
for j in "$@"; do
  s=$((66 + j))
  ncu --set full --clock-control none --import-source on -k regex:conv_tc -s $s -c 1 \
      -o /tmp/ncu/k$j -f $CMD > gpurun_out/ncu_k$j.log 2>&1
  echo "k$j rc=$?"
  ncu -i /tmp/ncu/k$j.ncu-rep --page source --csv > gpurun_out/src_k$j.csv 2>/dev/null
  ncu -i /tmp/ncu/k$j.ncu-rep --page raw --csv > gpurun_out/raw_k$j.csv 2>/dev/null
done
du -sh gpurun_out
