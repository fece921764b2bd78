set -x
python bench.py > gpurun_out/r2_v5_bench_c4.json 2> gpurun_out/r2_v5_bench_c4.err
tail -c 600 gpurun_out/r2_v5_bench_c4.err
python tools/make_profiles.py r2f > gpurun_out/r2f_make_profiles.log 2>&1
tail -5 gpurun_out/r2f_make_profiles.log
