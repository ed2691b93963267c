set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_trace.py tests/test_gpu_group.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -8 > gpurun_out/r02k_pytest.log; tail -8 gpurun_out/r02k_pytest.log
timeout 600 python tools/r02_tune.py --skip-k4 > gpurun_out/r02k_tune.log 2>&1; grep -v "^{" gpurun_out/r02k_tune.log | tail -30
