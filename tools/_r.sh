python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2_v7_bench_c4.json 2> gpurun_out/r2_v7_bench_c4.err; tail -c 300 gpurun_out/r2_v7_bench_c4.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_v7_bench_reference_arm.json 2>/dev/null
python tools/make_profiles.py r2f > gpurun_out/r2f_make_profiles.log 2>&1; tail -3 gpurun_out/r2f_make_profiles.log
