set -x
mkdir -p gpurun_out
timeout 600 python tools/r02_c5_sim.py > gpurun_out/r02j_c5_sim.log 2>&1; tail -9 gpurun_out/r02j_c5_sim.log | cut -c1-600
timeout 600 python tools/r02_tune.py --skip-k4 > gpurun_out/r02j_tune.log 2>&1; grep -v "^{" gpurun_out/r02j_tune.log | tail -30
