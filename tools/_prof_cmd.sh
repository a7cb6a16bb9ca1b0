timeout 900 python -m pytest tests/test_gpu_gae.py -m gpu -x -q -k pipeline 2>&1 | grep -v "^E    " | tail -30
