timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
RS_PERIOD=4 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench13.log 2>&1; tail -c 1500 gpurun_out/bench13.log
