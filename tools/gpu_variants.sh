#!/bin/bash
# Parity tests on the shipped library, then bench.py --quick for the shipped library and for every variant build under
# radiation_ppo_b200/_C/var_*.so (made with `python -m radiation_ppo_b200.build -DNAME=VALUE -ovar_x.so`), then one ncu capture.
#   gpurun --timeout 1500 -- 'bash tools/gpu_variants.sh r02c [noncu] [notests]'
tag=${1:-vX}
mkdir -p gpurun_out
if [ "$3" != "notests" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_$tag.log
fi
echo "== shipped"; timeout 300 python bench.py --steps 20 --warmup 5 --quick 2>&1 | tail -1
echo "== shipped, exact sampler"; timeout 300 python bench.py --steps 20 --warmup 5 --quick --exact-poisson 2>&1 | tail -1
echo "== shipped, ring 1 (state stays in L2)"; timeout 300 python bench.py --steps 20 --warmup 5 --quick --ring 1 2>&1 | tail -1
echo "== shipped, no resets"; timeout 300 python bench.py --steps 20 --warmup 5 --quick --episode-steps 30000 2>&1 | tail -1
for v in radiation_ppo_b200/_C/var_*.so; do
  echo "== $v"; RADSEARCH_B200_LIB=$PWD/$v timeout 300 python bench.py --steps 20 --warmup 5 --quick 2>&1 | tail -1
done
if [ "$2" != "noncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step1 -s 9 -c 2 -f -o gpurun_out/prof_step_$tag \
    python bench.py --steps 20 --warmup 5 --quick > gpurun_out/ncu_step_$tag.log 2>&1; echo "ncu step rc=$?"
fi
