#!/bin/bash
# ncu --set full capture of rs_prepare (reset_kernel<fast, 512, 1>) inside bench.py --quick, after a bench run of the same build
#   gpurun --timeout 900 -- 'bash tools/gpu_prof_prepare.sh <tag>'
tag=${1:-pp}
mkdir -p gpurun_out
true
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'reset_kernel.*512' -s 1 -c 1 -f \
    -o gpurun_out/prof_prepare_$tag python bench.py --steps 20 --warmup 5 --quick > gpurun_out/ncu_prepare_$tag.log 2>&1; echo "ncu prepare rc=$?"
