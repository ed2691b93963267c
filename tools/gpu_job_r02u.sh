set -x
timeout 900 python -m pytest tests/test_gpu_radiosity.py tests/test_gpu_pipeline.py tests/test_gpu_round2.py tests/test_gpu_bump.py -x -q 2>&1 | tail -5
timeout 900 python tools/k4_pack_probe.py 2>&1 | tail -12
