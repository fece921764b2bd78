set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_c.log 2>&1; tail -3 gpurun_out/r2_gputests_c.log
bash tools/prof_kernel.sh k_stage1 s1_prof 3 python tools/chi2_one.py
