timeout 600 python -m pytest tests/test_gpu_maps.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --no-cpu-baseline --steps 2000 > gpurun_out/bench_v15.log 2>&1; echo "rc=$?"; python - <<PY
import json
for l in open("gpurun_out/bench_v15.log"):
    if l.startswith("{"):
        d=json.loads(l); print("value %.4g"%d["value"], "kernel_ms %.4f"%d["roofline"]["kernel_ms"], d["maps"])
PY
