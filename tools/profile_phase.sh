#!/bin/bash
# source-level ncu capture of ONE launch per configuration: args = "ENV=..,ENV=..:kernel-regex:skip:tag" ...
out=gpurun_out
mkdir -p $out /tmp/ncu
for spec in "$@"; do
  IFS=: read -r cfg rx skip tag <<< "$spec"
  CMD="env $(echo "$cfg" | tr ',' ' ') python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o /tmp/ncu/$tag -f $CMD > $out/ncu_$tag.log 2>&1
  echo "$tag rc=$?"
  ncu -i /tmp/ncu/$tag.ncu-rep --page source --csv > $out/src_$tag.csv 2>/dev/null
  ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > $out/raw_$tag.csv 2>/dev/null
done
ls -la $out/src_*.csv $out/raw_*.csv | tail
