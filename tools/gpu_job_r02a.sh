# First GPU job of round 2: everything written after round 1's GPU budget ran out gets its first run.
#   gpurun --timeout 1500 -- 'bash tools/gpu_job_r02a.sh'
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).
set -x
mkdir -p gpurun_out
# 1. the new GPU tests on their own first (a failure here must not hide the state of the rest), then the whole suite
python -m pytest tests/test_gpu_zz_bsp_bake.py tests/test_gpu_zz_kd_fast.py tests/test_gpu_zzz_cpp_bake.py tests/test_gpu_bump.py -q -s --durations=8 2>&1 | tail -40 > gpurun_out/r02a_pytest_new.log; tail -5 gpurun_out/r02a_pytest_new.log
python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 > gpurun_out/r02a_pytest_gpu.log; tail -4 gpurun_out/r02a_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# 2. bench (N=1) incl. the bsp_side child process
python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; tail -c 1500 gpurun_out/r02a_bench_n1.json; tail -3 gpurun_out/r02a_bench_n1.err
python tools/bsp_side_bench.py > gpurun_out/r02a_bsp_side.json 2> gpurun_out/r02a_bsp_side.err; cat gpurun_out/r02a_bsp_side.json; tail -3 gpurun_out/r02a_bsp_side.err
# 3. ncu: launch lists + metrics for K5 and the binned kd build
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 200 ncu --metrics $M --clock-control none -k regex:'k5_finalize' -c 6 --csv --page raw --log-file gpurun_out/r02a_k5.csv python tools/profile_target.py k5 > gpurun_out/r02a_k5.log 2>&1; tail -2 gpurun_out/r02a_k5.log; wc -c gpurun_out/r02a_k5.csv
timeout 300 ncu --metrics $M --clock-control none -k regex:'k_for_each|DeviceScan' -c 400 --csv --page raw --log-file gpurun_out/r02a_kdfast.csv python tools/profile_target.py kdfast > gpurun_out/r02a_kdfast.log 2>&1; tail -2 gpurun_out/r02a_kdfast.log; wc -c gpurun_out/r02a_kdfast.csv
# 4. compute-sanitizer on the new kernels (small sizes through the tests)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_zz_bsp_bake.py -q -k "k5 or add_bsp or radial" 2>&1 | tail -6 > gpurun_out/r02a_sanitizer_k5.txt; tail -3 gpurun_out/r02a_sanitizer_k5.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_zz_kd_fast.py -q -k "device_tree or same_tree" 2>&1 | tail -6 > gpurun_out/r02a_sanitizer_kdfast.txt; tail -3 gpurun_out/r02a_sanitizer_kdfast.txt
