timeout 900 python -m pytest tests/test_gpu_env.py tests/test_gpu_maps.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --steps 4000 > gpurun_out/bench_v30.log 2>&1; echo "rc=$?"; python - <<PY
import json
for l in open("gpurun_out/bench_v30.log"):
    if l.startswith("{"):
        d=json.loads(l); print("value %.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel_ms %.4f"%d["roofline"]["kernel_ms"], "frac %.4f"%d["roofline"]["frac"], d["maps"]["pipeline_env_steps_per_s"])
PY
done
