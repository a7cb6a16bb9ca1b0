timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 8 -c 1 -f -o gpurun_out/prof_step_v10 \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/ncu_step_v10.log 2>&1; echo "ncu step rc=$?"
