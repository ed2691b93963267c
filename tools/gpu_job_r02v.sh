set -x
mkdir -p gpurun_out
cat > /tmp/pk_target.py <<'PY'
import sys
sys.path.insert(0, '.')
import torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
s = scenes.multi_room()
env = environment_from_scene(s)
env.set_option("k4_graph", 0)
env.build_transfers(s.pvs)
e0 = torch.full((s.n_patches, 3), 100.0, device="cuda"); o = torch.empty_like(e0)
for pack in (1, 0):
    env.set_option("k4_pack", pack)
    env.bounce(e0, 3, out=o, want_added=False)
env.close()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k4_gather' -c 6 -o gpurun_out/r02_k4_packed_full python /tmp/pk_target.py > gpurun_out/r02v.log 2>&1; tail -2 gpurun_out/r02v.log
ls -la gpurun_out/*.ncu-rep
