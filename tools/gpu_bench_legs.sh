#!/bin/bash
# Runs the parity tests, then every bench leg in its own process (a failing leg does not hide the others), then the ncu passes.
#   gpurun --timeout 1500 -- 'bash tools/gpu_bench_legs.sh r02b'
tag=${1:-vX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_$tag.log
for leg in none exact sweep gae pipeline maps e2e cpu; do
  timeout 600 python bench.py --steps 20 --warmup 5 --legs $leg > gpurun_out/bench_${tag}_$leg.log 2>&1; echo "bench[$leg] rc=$?"
  tail -c 2500 gpurun_out/bench_${tag}_$leg.log | grep -v "^$" | tail -12
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step1_kernel -s 8 -c 2 -f -o gpurun_out/prof_step_$tag \
    python bench.py --steps 20 --warmup 5 --quick > gpurun_out/ncu_step_$tag.log 2>&1; echo "ncu step rc=$?"
