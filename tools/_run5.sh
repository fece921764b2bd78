set -x
timeout 600 python -m pytest tests/test_gpu_i8.py -x -q 2>&1 | tail -3
timeout 200 python tools/i8_time.py 2>&1 | tail -1
