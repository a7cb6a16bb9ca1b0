timeout 900 python -m pytest tests/test_gpu_env.py -m gpu -x -q -k sharded 2>&1 | grep -v "^E    " | tail -20
