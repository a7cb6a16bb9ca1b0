#!/bin/bash
# ad-hoc experiment runner (edited per experiment): block period / ring sweep of the headline and stream timelines
tag=${1:-x}
mkdir -p gpurun_out
{
for pr in "4 4" "5 4" "10 2" "10 4" "20 4" "4 5" "2 4" "4 8"; do
  set -- $pr
  echo "== period $1 ring $2 (K=20)"; timeout 300 python bench.py --steps 20 --warmup 5 --quick --period $1 --ring $2 2>&1 | tail -1
done
for pr in "4 4" "8 4" "12 4" "24 4" "8 2"; do
  set -- $pr
  echo "== period $1 ring $2 (K=240)"; timeout 300 python bench.py --steps 240 --warmup 24 --quick --period $1 --ring $2 2>&1 | tail -1
done
timeout 300 python tools/kernel_timeline.py --rows 150 > gpurun_out/timeline_${tag}_p4.log 2>&1
timeout 300 python tools/kernel_timeline.py --rows 150 --period 8 > gpurun_out/timeline_${tag}_p8.log 2>&1
head -8 gpurun_out/timeline_${tag}_p4.log gpurun_out/timeline_${tag}_p8.log
} 2>&1 | tee gpurun_out/exp_$tag.log
