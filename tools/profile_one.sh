#!/bin/bash
# source-level ncu capture of ONE launch: args = kernel regex, launches to skip (of that regex), tag
out=gpurun_out
mkdir -p $out /tmp/ncu
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > $out/plain_one.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o /tmp/ncu/one -f $CMD > $out/ncu_one_$3.log 2>&1
echo "rc=$?"
ncu -i /tmp/ncu/one.ncu-rep --page source --csv > $out/src_$3.csv 2>/dev/null
ncu -i /tmp/ncu/one.ncu-rep --page raw --csv > $out/raw_$3.csv 2>/dev/null
ls -la $out/src_$3.csv $out/raw_$3.csv
