timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for occ in 6 8; do RS_STEP_OCC=$occ timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench12_occ$occ.log 2>&1; python - <<PY
import json
d=json.loads(open("gpurun_out/bench12_occ$occ.log").read().strip().splitlines()[-1])
print("occ$occ", "value %.3e"%d["value"], "kernel_ms %.4f"%d["roofline"]["kernel_ms"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 40 --launch-count 2 -o gpurun_out/prof_step_v7 -f python bench.py --steps 20 --warmup 12 --no-cpu-baseline > gpurun_out/ncu7.log 2>&1
