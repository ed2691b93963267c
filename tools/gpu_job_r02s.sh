set -x
timeout 900 python -m pytest tests/test_gpu_poly_form_factor.py tests/test_gpu_hier.py tests/test_gpu_radiosity.py tests/test_gpu_group.py -x -q 2>&1 | tail -15
