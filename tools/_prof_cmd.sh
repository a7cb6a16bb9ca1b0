timeout 900 python -m pytest tests/test_gpu_env.py -m gpu -x -q -k soak 2>&1 | grep -v "^E    " | tail -30
