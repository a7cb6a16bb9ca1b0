set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench11.log 2>&1; tail -c 1800 gpurun_out/bench11.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 40 --launch-count 2 -o gpurun_out/prof_step_v6 -f python bench.py --steps 20 --warmup 12 --no-cpu-baseline > gpurun_out/ncu6.log 2>&1
