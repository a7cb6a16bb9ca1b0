#!/bin/bash
# A/B timing of library builds on one GPU box: bench.py --quick for the shipped library and every radiation_ppo_b200/_C/var_*.so
# (python -m radiation_ppo_b200.build -DNAME=VALUE -ovar_x.so), optionally at other batch sizes, then the GPU tests and an ncu
# capture of the single-agent step kernel of the shipped library.
#   gpurun --timeout 1500 -- 'bash tools/gpu_ab.sh <tag> [tests] [ncu] [big]'
tag=${1:-ab}
mkdir -p gpurun_out
q="python bench.py --steps 20 --warmup 5 --quick"
echo "== shipped"; timeout 300 $q 2>&1 | tail -1
echo "== shipped, no resets"; timeout 300 $q --episode-steps 30000 2>&1 | tail -1
for v in radiation_ppo_b200/_C/var_*.so; do
  [ -f "$v" ] || continue
  echo "== $v"; RADSEARCH_B200_LIB=$PWD/$v timeout 300 $q 2>&1 | tail -1
done
if [[ " $* " == *" big "* ]]; then
  echo "== shipped, 262144 envs"; timeout 300 $q --envs-per-gpu 262144 --ring 2 2>&1 | tail -1
  for v in radiation_ppo_b200/_C/var_old*.so; do
    [ -f "$v" ] || continue
    echo "== $v, 262144 envs"; RADSEARCH_B200_LIB=$PWD/$v timeout 300 $q --envs-per-gpu 262144 --ring 2 2>&1 | tail -1
  done
fi
if [[ " $* " == *" tests "* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
  tail -3 gpurun_out/pytest_gpu_$tag.log
fi
if [[ " $* " == *" ncu "* ]]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:step1 -s 9 -c 2 -f -o gpurun_out/prof_step_$tag \
      $q > gpurun_out/ncu_step_$tag.log 2>&1; echo "ncu step rc=$?"
fi
