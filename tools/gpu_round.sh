#!/bin/bash
# One GPU-box round trip: parity tests, bench (ours + reference arm), ncu launch list, one ncu --set full capture of the
# step kernel and of the GAE kernel.  Usage (from the repo root, through gpurun):
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh v8'
# Everything lands in gpurun_out/; tools/ncu_summary.py turns it into profiles/*.txt here.
tag=${1:-vX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$tag.log
tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/bench_$tag.log 2>&1; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_$tag.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$tag.log 2>&1; echo "ref rc=$?"
tail -c 600 gpurun_out/bench_ref_$tag.log
if [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 8 -c 2 -f -o gpurun_out/prof_step_$tag \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step_$tag.log 2>&1; echo "ncu step rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gae_ -s 3 -c 1 -f -o gpurun_out/prof_gae_$tag \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gae_$tag.log 2>&1; echo "ncu gae rc=$?"
fi
