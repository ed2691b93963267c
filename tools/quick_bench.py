"""Development micro-benchmark (not the driver contract; see bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene

def ev_time(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

which = sys.argv[1] if len(sys.argv) > 1 else "s1"
scene = scenes.box_room() if which == "s1" else (scenes.multi_room() if which == "s2" else scenes.outdoor())
t = time.time(); env = environment_from_scene(scene); print("build+upload s", time.time() - t, env.stats())
env.set_stream(torch.cuda.current_stream().cuda_stream); env.set_async(True)
n = 1 << 24
a, b = scenes.shadow_segments(scene, n)
ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
out = torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda")
best, med = ev_time(lambda: env.test_lines(ta, tb, out=out))
print(f"test_lines {which}: {n/best/1e6:.1f} Mrays/s best, {n/med/1e6:.1f} median ({best:.3f} ms)")
r = scenes.random_rays(scene, 1 << 22)
to, td, tm = torch.from_numpy(r["o"]).cuda(), torch.from_numpy(r["d"]).cuda(), torch.from_numpy(r["tmax"]).cuda()
outs = (torch.empty(1 << 22, dtype=torch.int32, device="cuda"), torch.empty(1 << 22, dtype=torch.int32, device="cuda"), torch.empty(1 << 22, dtype=torch.float32, device="cuda"))
best, med = ev_time(lambda: env.trace_rays(to, td, tm, out=outs))
print(f"trace_rays random {which}: {(1<<22)/best/1e6:.1f} Mrays/s best ({best:.3f} ms)")
if which == "s3":
    sys.exit(0)
# K2 + K4
env.set_async(False)
env.build_transfers(scene.pvs)
t = time.time(); nnz = env.build_transfers(scene.pvs); dt = time.time() - t
print(f"build_transfers: nnz={nnz} N={scene.n_patches} wall {dt:.3f}s kernel_ms={env.last_timing()}")
N = scene.n_patches
emit0 = torch.full((N, 3), 100.0, device="cuda"); tot = torch.empty_like(emit0)
env.set_async(True)
nb = 100
best, med = ev_time(lambda: env.bounce(emit0, nb, out=tot, want_added=False), iters=3, warm=1)
bytes_it = 8 * nnz + 40 * N
print(f"bounce: {nb/ (best/1e3):.1f} iters/s, {bytes_it*nb/(best/1e3)/1e9:.1f} GB/s algorithmic ({best/nb*1e3:.1f} us/iter)")
pos, nrm = (scene.luxel_pos, scene.luxel_normal) if scene.luxel_pos is not None else (scene.patch_origin, scene.patch_normal)
tp, tn = torch.from_numpy(pos).cuda(), torch.from_numpy(nrm).cuda()
rgb = torch.empty((pos.shape[0], 3), device="cuda")
env.set_async(False)
best, med = ev_time(lambda: env.direct_light(tp, tn, scene.lights, out=rgb), iters=3, warm=1)
print(f"direct_light: {pos.shape[0]} luxels x {len(scene.lights)} lights: {best:.3f} ms -> {pos.shape[0]*len(scene.lights)/best/1e3:.1f} Mrays/s (upper bound, culled pairs included)")
