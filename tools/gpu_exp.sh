#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
{
for legs in e2e e2e,exact e2e,sweep e2e,gae e2e,pipeline e2e,maps e2e,cpu; do
  echo "== legs=$legs"; python bench.py --steps 20 --warmup 5 --legs $legs 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('  e2e %.4g sync %.4g pipelined %.4g frac_of_link %.3f'%(e['value'], e['sync_value'], e['pipelined_value'], e['frac_of_link']))"
done
} 2>&1 | tee gpurun_out/exp_$tag.log
