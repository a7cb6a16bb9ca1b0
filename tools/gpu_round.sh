#!/bin/bash
# One GPU-box round trip: parity tests, smoke, bench (ours + reference arm), then (unless "noncu") the ncu launch list and
# one ncu --set full capture of the step kernel.  Usage (from the repo root, through gpurun):
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r02a [noncu] [notests]'
# Everything lands in gpurun_out/; tools/ncu_summary.py turns the reports into profiles/*.txt here.
tag=${1:-vX}
mkdir -p gpurun_out
if [ "$3" != "notests" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$tag.log
tail -5 gpurun_out/pytest_gpu_$tag.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"
fi
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.log 2>&1; echo "bench rc=$?"
tail -c 4000 gpurun_out/bench_$tag.log
if [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 20 --warmup 5 --quick > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step1_kernel -s 8 -c 2 -f -o gpurun_out/prof_step_$tag \
    python bench.py --steps 20 --warmup 5 --quick > gpurun_out/ncu_step_$tag.log 2>&1; echo "ncu step rc=$?"
fi
