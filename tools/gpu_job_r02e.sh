# Round 2, GPU job 4: plain kernel back for one GPU, work-item kernel with steal pool, sort key layouts, geometric host chunks.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_radiosity.py tests/test_gpu_trace.py -x -q 2>&1 | tail -8 > gpurun_out/r02e_pytest_new.log; tail -8 gpurun_out/r02e_pytest_new.log
timeout 1200 python tools/r02_tune.py > gpurun_out/r02e_tune.log 2>&1; grep -v "^{" gpurun_out/r02e_tune.log | tail -90
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,l1tex__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio


timeout 400 ncu --metrics $M --clock-control none -k regex:'k4_gather_items' -c 14 --csv --page raw --log-file gpurun_out/r02e_k4.csv python tools/profile_target.py r2k4 > gpurun_out/r02e_k4.log 2>&1; tail -2 gpurun_out/r02e_k4.log; wc -c gpurun_out/r02e_k4.csv
