set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_graph.py -x -q 2>&1 | tail -3
timeout 300 python tools/time_stages.py 65536 16384 4096 1024 2>&1 | tail -4
WORKLOAD=c1 timeout 300 python tools/time_stages.py 64 2>&1 | tail -1
