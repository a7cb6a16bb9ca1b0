timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --no-cpu-baseline --steps 4000 > gpurun_out/bench_v11.log 2>&1; echo "rc=$?"; python - <<PY
import json
for l in open("gpurun_out/bench_v11.log"):
    if l.startswith("{"):
        d=json.loads(l); print("value %.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel_ms %.4f"%d["roofline"]["kernel_ms"], "e2e %.4g"%d["e2e"]["value"], d["clocks"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 8 -c 1 -f -o gpurun_out/prof_step_v11 \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/ncu_step_v11.log 2>&1; echo "ncu step rc=$?"
