set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_radiosity.py tests/test_gpu_hier.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -6
VRAD_TIMING=1 timeout 600 python tools/k2_stream_probe.py 2>&1 | tail -30
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:k2_visibility -c 2 --csv --log-file gpurun_out/r02_k2_stream_ncu.csv python - <<'PY'
import sys
sys.path.insert(0, '.')
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
s = scenes.multi_room()
for flag in (1, 0):
    env = environment_from_scene(s); env.set_option("k2_stream", flag); env.build_transfers(s.pvs); env.close()
PY
python tools/ncu_table.py gpurun_out/r02_k2_stream_ncu.csv 2>&1 | tail -8
