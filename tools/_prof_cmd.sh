timeout 900 ncu --set full --clock-control none --import-source on -k regex:reset_kernel -s 150 -c 10 -f -o gpurun_out/prof_reset_v31 \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/ncu_reset_v31.log 2>&1; echo "ncu reset rc=$?"
