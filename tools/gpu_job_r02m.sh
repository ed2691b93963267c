# compute-sanitizer over the round-2 kernels (small sizes through the tests)
set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_round2.py -q -x -k "gather_items or split_rows or simulated_peers or world1 or indexed_segments_equal or sorted_order_is_invisible or top_levels" 2>&1 | tail -8 > gpurun_out/r02_sanitizer_memcheck.txt; tail -6 gpurun_out/r02_sanitizer_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_round2.py -q -x -k "gather_items_any_plan_same_light and 2048 or simulated_peers" 2>&1 | tail -8 > gpurun_out/r02_sanitizer_racecheck.txt; tail -6 gpurun_out/r02_sanitizer_racecheck.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_group.py tests/test_gpu_bump.py -q -x 2>&1 | tail -8 > gpurun_out/r02_sanitizer_group.txt; tail -6 gpurun_out/r02_sanitizer_group.txt
timeout 600 python tools/k2_phase_probe.py 2>&1 | grep "phases\|wall" | tail -6
