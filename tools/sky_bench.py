"""Development micro-benchmark of vrad_test_lines_sky (not the driver contract; see bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import Environment

def ev_time(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

sc = scenes.sky_room(); m = sc.meta
g = Environment(); g.add_triangles(sc.tri_ids, sc.tri_verts, sc.tri_flags); g.set_triangle_colors(m["tri_colors"])
g.setup_acceleration_structure(); g.bsp_upload(m["bsp"]); g.process_sky_cameras(m["cams_origin"], m["cams_scale"])
g.set_stream(torch.cuda.current_stream().cuda_stream); g.set_async(True)
n = 1 << 23
a, b = scenes.sky_segments(sc, n)
ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
out = torch.empty(n, dtype=torch.float32, device="cuda")
bits = torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda")
for mode in (0, 1):
    best, _ = ev_time(lambda: g.test_lines(ta, tb, sky_mode=mode, out=bits))
    print(f"test_lines sky_mode={mode}: {n/best/1e6:.2f} Gsegments/s ({best:.3f} ms)")
for flags in (0, 1, 2, 3, 7):
    best, _ = ev_time(lambda: g.test_lines_sky(ta, tb, flags, 7, out=out))
    print(f"test_lines_sky flags={flags}: {n/best/1e6:.2f} Gsegments/s ({best:.3f} ms, launches {g.last_timing()[1]})")
pts = torch.rand((1 << 24, 3), device="cuda") * 1000 - 500
leaf = torch.empty(1 << 24, dtype=torch.int32, device="cuda")
import ctypes as C
from vrad_b200.lib import check, ptr
best, _ = ev_time(lambda: check(g._l.vrad_point_leafnum(g._h, C.c_int64(1 << 24), ptr(pts), ptr(leaf))))
print(f"point_leafnum: {(1<<24)/best/1e6:.1f} Gpoints/s ({best:.3f} ms; {(1<<24)*16/best/1e6:.1f} GB/s of 16 B/point)")
