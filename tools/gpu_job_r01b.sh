set -x
python -m pytest tests -m gpu -x -q --durations=10 2>&1 | tail -25 > gpurun_out/r01b_pytest_gpu.log; tail -4 gpurun_out/r01b_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r01b_bench_n1.json 2> gpurun_out/r01b_bench_n1.err; tail -c 600 gpurun_out/r01b_bench_n1.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
timeout 300 ncu --metrics $M --clock-control none -k regex:'k1_sky|k_point' --csv --page raw --log-file gpurun_out/r01b_sky.csv python tools/profile_target.py sky > gpurun_out/r01b_sky.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:'k2_visibility|k4_gather_short|k4_collect' -c 8 --csv --page raw --log-file gpurun_out/r01b_hier.csv python tools/profile_target.py hier > gpurun_out/r01b_hier.log 2>&1
ls -la gpurun_out | tail -8
