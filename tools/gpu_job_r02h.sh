# Round 2, GPU job 7 (2 GPUs): in-process multi-GPU handle, the whole GPU suite, smoke.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py -x -q --durations=6 2>&1 | tail -25 > gpurun_out/r02h_pytest_group.log; tail -25 gpurun_out/r02h_pytest_group.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 > gpurun_out/r02h_pytest_gpu.log; tail -8 gpurun_out/r02h_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
