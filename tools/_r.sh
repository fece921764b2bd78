python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_d.log 2>&1; tail -2 gpurun_out/r2_gputests_d.log
python bench.py > gpurun_out/r2_v8_bench_c4.json 2> gpurun_out/r2_v8_bench_c4.err; tail -c 300 gpurun_out/r2_v8_bench_c4.err
python tools/make_profiles.py r2f > gpurun_out/r2f_make_profiles.log 2>&1; tail -3 gpurun_out/r2f_make_profiles.log
