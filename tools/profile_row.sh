#!/bin/bash
# ncu capture of the three row-kernel launches (UNETB200_ROW64=3) of the 4th forward, with source
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out /tmp/ncu
export UNETB200_ROW64=${ROW64:-3}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > $out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_row -s 9 -c 3 -o /tmp/ncu/row -f $CMD > $out/ncu_row_$tag.log 2>&1
echo "rc=$?"
ncu -i /tmp/ncu/row.ncu-rep --page raw --csv > $out/prof_${tag}_row_raw.csv 2>/dev/null
cp /tmp/ncu/row.ncu-rep $out/prof_${tag}_row.ncu-rep
ls -la $out | tail -5; du -sh $out
tail -n 1 $out/plain_$tag.log | cut -c1-200
