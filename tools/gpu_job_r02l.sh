# Round 2, GPU job: evidence for profiles/ -- launch list of the bench command, full ncu sets of the dominant kernels, metric passes
# for the late stages (K3, K5, radial filter, bump gather, binned kd build).
set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,l1tex__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
# 1. launch list of the bench command (N=1; numbers printed under ncu are not bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --no-large --no-cpu > gpurun_out/r02l_bench_under_ncu.log 2>&1; tail -2 gpurun_out/r02l_bench_under_ncu.log | cut -c1-300; wc -l gpurun_out/r02_final_launches.csv
# 2. full sets of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k4_gather' -c 9 -o gpurun_out/r02_k4_full python tools/profile_target.py r2k4 > gpurun_out/r02l_k4_full.log 2>&1; tail -2 gpurun_out/r02l_k4_full.log
ncu -i gpurun_out/r02_k4_full.ncu-rep --page raw --csv > gpurun_out/r02_k4_full_raw.csv 2>/dev/null; ls -la gpurun_out/r02_k4_full.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_test_lines' -c 8 -o gpurun_out/r02_k1_s1_full python tools/profile_target.py r2k1 > gpurun_out/r02l_k1_full.log 2>&1; tail -2 gpurun_out/r02l_k1_full.log
ncu -i gpurun_out/r02_k1_s1_full.ncu-rep --page raw --csv > gpurun_out/r02_k1_s1_full_raw.csv 2>/dev/null; ls -la gpurun_out/r02_k1_s1_full.ncu-rep
timeout 600 ncu --metrics $M --clock-control none -k regex:'k1_test_lines|k1_sort_keys|DeviceRadixSort' -c 40 --csv --page raw --log-file gpurun_out/r02_k1_s3.csv python tools/profile_target.py r2k1s3 > gpurun_out/r02l_k1_s3.log 2>&1; tail -1 gpurun_out/r02l_k1_s3.log
# 3. the late stages
timeout 600 ncu --metrics $M --clock-control none -k regex:'k3_' -c 24 --csv --page raw --log-file gpurun_out/r02_k3.csv python tools/profile_target.py k3 > gpurun_out/r02l_k3.log 2>&1; tail -1 gpurun_out/r02l_k3.log
timeout 300 ncu --metrics $M --clock-control none -k regex:'k5_finalize' -c 6 --csv --page raw --log-file gpurun_out/r02_k5.csv python tools/profile_target.py k5 > gpurun_out/r02l_k5.log 2>&1; tail -1 gpurun_out/r02l_k5.log
timeout 300 ncu --metrics $M --clock-control none -k regex:'k4_gather_bump' -c 4 --csv --page raw --log-file gpurun_out/r02_bump.csv python tools/profile_target.py bump > gpurun_out/r02l_bump.log 2>&1; tail -1 gpurun_out/r02l_bump.log
timeout 600 ncu --metrics $M --clock-control none -k regex:'k_radial|k5_finalize_patches|k_for_each|DeviceScan' -c 300 --csv --page raw --log-file gpurun_out/r02_bake_kdfast.csv python tools/bsp_side_bench.py > gpurun_out/r02l_bsp_side.log 2>&1; tail -c 600 gpurun_out/r02l_bsp_side.log
ls -la gpurun_out/ | tail -30
