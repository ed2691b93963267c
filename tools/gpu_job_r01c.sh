set -x
python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -14 > gpurun_out/r01c_pytest_gpu.log; tail -3 gpurun_out/r01c_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/r01c_bench_n1.json 2> gpurun_out/r01c_bench_n1.err; tail -c 300 gpurun_out/r01c_bench_n1.json; tail -3 gpurun_out/r01c_bench_n1.err
