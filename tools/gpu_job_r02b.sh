# Round 2, GPU job 1: new K1/K4 paths -- parity first, then the whole suite, then the tuning sweep.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q --durations=8 2>&1 | tail -40 > gpurun_out/r02b_pytest_new.log; tail -15 gpurun_out/r02b_pytest_new.log
timeout 900 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -30 > gpurun_out/r02b_pytest_gpu.log; tail -6 gpurun_out/r02b_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python tools/r02_tune.py > gpurun_out/r02b_tune.log 2>&1; tail -60 gpurun_out/r02b_tune.log
