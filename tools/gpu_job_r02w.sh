set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_multi.py tests/test_gpu_group.py tests/test_gpu_radiosity.py -x -q 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tools/multi_bounce_probe.py 2>&1 | grep "hierarchy\|^192\|^256"
