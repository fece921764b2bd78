set -x
timeout 900 python -m pytest tests/test_gpu_chi2_i8.py -x -q 2>&1 | tail -15
timeout 300 python tools/time_chi2_i8.py 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_i8.py -x -q 2>&1 | tail -5
timeout 200 python tools/i8_time.py 2>&1 | tail -3
C1D_I8_MERGE=0 timeout 200 python tools/i8_time.py 2>&1 | tail -1
