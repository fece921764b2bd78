set -x
timeout 900 python -m pytest tests/test_gpu_chi2_i8.py -x -q 2>&1 | tail -5
timeout 300 python tools/time_chi2_i8.py 2>&1 | tail -8
bash tools/prof_kernel.sh k_chi2_i8 chi2i8_prof4 2 python tools/chi2_one.py
