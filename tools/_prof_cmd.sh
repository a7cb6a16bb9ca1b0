timeout 600 python -m pytest tests/test_gpu_gae.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/gpu_experiments.py 2>&1 | tail -50
