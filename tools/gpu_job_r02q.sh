set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py tests/test_gpu_hier.py tests/test_gpu_bump.py tests/test_gpu_trace.py tests/test_gpu_pipeline.py tests/test_gpu_cpp_driver.py -x -q 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tools/multi_bounce_probe.py 2>&1 | grep "hierarchy\|^192 1 12 1 1"
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
hs = scenes.multi_room_hier(nx=12, ny=11); tr = hs.meta["tree"]
h = environment_from_scene(hs); h.set_hierarchy(tr["parent"], tr["child1"], tr["child2"], tr["face"]); h.set_stream(torch.cuda.current_stream().cuda_stream)
h.build_transfers(hs.pvs)
he = torch.full((hs.n_patches, 3), 100.0, device="cuda"); ho = torch.empty_like(he); h.set_async(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for g in (1, 0):
    h.set_option("k4_graph", g); h.bounce(he, 100, out=ho, want_added=False)
    e0.record(); h.bounce(he, 100, out=ho, want_added=False); e1.record(); torch.cuda.synchronize()
    print("hier N=1 graph", g, e0.elapsed_time(e1) * 10, "us/bounce")
s1 = scenes.box_room(); env = environment_from_scene(s1, with_patches=False); env.set_stream(torch.cuda.current_stream().cuda_stream); env.set_async(True)
r = scenes.random_rays(s1, 1 << 24)
o, d, t = (torch.from_numpy(r[k]).cuda() for k in ("o", "d", "tmax"))
out = env.trace_rays(o, d, t)
e0.record()
for _ in range(3): env.trace_rays(o, d, t, out=out)
e1.record(); torch.cuda.synchronize()
print("closest-hit rays/s S1", 3 * (1 << 24) / (e0.elapsed_time(e1) * 1e-3))
PY
