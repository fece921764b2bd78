set -x
timeout 900 python -m pytest tests/test_gpu_graph.py -x -q 2>&1 | tail -15
timeout 300 python tools/time_small.py 2>&1 | tail -7
