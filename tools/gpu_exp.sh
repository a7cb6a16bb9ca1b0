#!/bin/bash
# ad-hoc measurement list (one line per experiment), see gpurun_out/exp_<tag>.log
tag=${1:-x}
mkdir -p gpurun_out
{
for v in "" radiation_ppo_b200/_C/var_*.so; do
  echo "== lib=$v"; RADSEARCH_B200_LIB=${v:+$PWD/$v} python bench.py --steps 20 --warmup 5 --legs sweep 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('  headline value %.4g  ms/step %.4f  kernel_ms %.4f  single %.4f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_ms_single_launch']), ' sweep:', [(s['n_envs'], round(s['kernel_ms'],4)) for s in d['sweep']])"
done
} 2>&1 | tee gpurun_out/exp_$tag.log
