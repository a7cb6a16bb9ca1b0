#!/bin/bash
# ad-hoc measurement list (one line per experiment), see gpurun_out/exp_<tag>.log
tag=${1:-x}
mkdir -p gpurun_out
{
echo "== shipped"; python bench.py --steps 20 --warmup 5 --quick 2>&1 | tail -1
echo "== shipped --no-prefetch"; python bench.py --steps 20 --warmup 5 --quick --no-prefetch 2>&1 | tail -1
echo "== shipped --no-graph"; python bench.py --steps 20 --warmup 5 --quick --no-graph 2>&1 | tail -1
echo "== shipped ring 2"; python bench.py --steps 20 --warmup 5 --quick --ring 2 2>&1 | tail -1
echo "== shipped ring 8"; python bench.py --steps 20 --warmup 5 --quick --ring 8 2>&1 | tail -1
for v in radiation_ppo_b200/_C/var_*.so; do
  echo "== $v"; RADSEARCH_B200_LIB=$PWD/$v python bench.py --steps 20 --warmup 5 --quick 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/exp_$tag.log
