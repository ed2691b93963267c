# 2 GPUs: hierarchical fused exchange, multi-GPU bump totals (process per GPU and in-process), whole suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -x -q --durations=5 2>&1 | tail -25 > gpurun_out/r02n_pytest_multi.log; tail -25 gpurun_out/r02n_pytest_multi.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02n_pytest_gpu.log; tail -8 gpurun_out/r02n_pytest_gpu.log
