"""Short workload for `ncu --set full` captures: a few launches at bench sizes (k1, k1s3, sky, hier, k5, kdfast, default = K2/K3/K4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene

what = sys.argv[1] if len(sys.argv) > 1 else "k1"
if what in ("k1", "k1s3"):
    s = scenes.box_room() if what == "k1" else scenes.outdoor(); env = environment_from_scene(s, with_patches=False)
    a, b = scenes.shadow_segments(s, 1 << 24)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.empty((1 << 24) // 32, dtype=torch.int32, device="cuda")
    for _ in range(4): env.test_lines(ta, tb, out=out)
    r = scenes.random_rays(s, 1 << 22)
    env.trace_rays(torch.from_numpy(r["o"]).cuda(), torch.from_numpy(r["d"]).cuda(), torch.from_numpy(r["tmax"]).cuda())
elif what in ("r2k1", "r2k1s3"):
    # round 2: sorted / unsorted / top-level-staged visibility kernels on the bench inputs, one launch each after a warm-up
    big = what == "r2k1s3"
    s = scenes.outdoor() if big else scenes.box_room(); env = environment_from_scene(s, with_patches=False)
    n = 1 << (22 if big else 24)
    a, b = scenes.shadow_segments(s, n, seed=0xC5 if big else 0xC0FFEE)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.empty(n // 32, dtype=torch.int32, device="cuda")
    for sort, top, st in ((0, 0, 1), (1, 0, 1), (0, 1023, 1), (0, 0, 0)):      # streaming unordered, ordered, top levels staged, round-1 per-chunk kernel
        env.set_option("k1_sort", sort); env.set_option("k1_top", top); env.set_option("k1_stream", st)
        for _ in range(2): env.test_lines(ta, tb, out=out)
elif what == "r2k4":
    # round 2: the multi-GPU gather kernel on the rank-3 slice of a simulated world-8 run, and the single-GPU kernel on the whole matrix
    s = scenes.multi_room()
    for world, rank in ((8, 3), (1, 0)):
        env = environment_from_scene(s, rank=rank, world=world)
        if world > 1: env.set_option("k4_sim_peers", 1)
        env.set_option("k4_graph", 0)
        env.build_transfers(s.pvs)
        e0 = torch.full((s.n_patches, 3), 100.0, device="cuda"); o = torch.empty_like(e0)
        env.bounce(e0, 6, out=o, want_added=False)
        env.close()
elif what == "k3":
    # K3 at both ends: 8 point/spot lights on the S1 luxels (C2), sun + 162-direction sky ambient on the 2 M luxels of S3 (C5)
    s = scenes.box_room(); env = environment_from_scene(s, with_patches=False)
    pos, nrm = torch.from_numpy(s.luxel_pos).cuda(), torch.from_numpy(s.luxel_normal).cuda()
    for _ in range(2): env.direct_light(pos, nrm, s.lights)
    env.close()
    s = scenes.outdoor(); env = environment_from_scene(s, with_patches=False)
    env.set_sky_dirs(np.loadtxt(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vrad_b200", "data", "anorms.txt"), dtype=np.float32))
    pos, nrm = torch.from_numpy(s.luxel_pos).cuda(), torch.from_numpy(s.luxel_normal).cuda()
    for _ in range(2): env.direct_light(pos, nrm, s.lights)
elif what == "bump":
    # bump-mapped (4-normal) gather on the C4 map: every patch bump-mapped, 4 bounces
    from vrad_b200.environment import bump_normals
    s = scenes.multi_room(); env = environment_from_scene(s)
    N = s.n_patches
    basis = np.zeros((N, 3, 3), np.float32)
    for f in np.unique(s.patch_normal, axis=0):
        sel = np.all(s.patch_normal == f, axis=1)
        sv = np.array([0, 1, 0], np.float32) if abs(f[0]) > 0.5 else np.array([1, 0, 0], np.float32)
        tv = np.cross(f, sv).astype(np.float32)
        basis[sel] = bump_normals(sv, tv, f, f)
    env.set_bump(np.ones(N, np.uint8), basis)
    env.build_transfers(s.pvs)
    e0 = torch.full((N, 3), 100.0, device="cuda")
    env.bounce(e0, 4)
elif what == "sky":
    from vrad_b200.environment import Environment
    s = scenes.sky_room(); m = s.meta
    env = Environment(); env.add_triangles(s.tri_ids, s.tri_verts, s.tri_flags); env.set_triangle_colors(m["tri_colors"])
    env.setup_acceleration_structure(); env.bsp_upload(m["bsp"]); env.process_sky_cameras(m["cams_origin"], m["cams_scale"])
    a, b = scenes.sky_segments(s, 1 << 23)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.empty(1 << 23, dtype=torch.float32, device="cuda")
    for _ in range(2): env.test_lines_sky(ta, tb, 3, 7, out=out)
elif what == "k5":
    import ctypes as C
    from vrad_b200 import lib as vlib
    from vrad_b200.environment import Environment
    env = Environment()
    n = 1 << 24
    d_dir = torch.rand((n, 3), device="cuda") * 400.0; d_ind = torch.rand((n, 3), device="cuda") * 100.0
    d_idx = torch.randint(0, 200000, (n,), device="cuda", dtype=torch.int32).sort().values      # neighbouring luxels share a patch
    d_tot = torch.rand((200000, 3), device="cuda") * 100.0
    d_out = torch.empty(n, dtype=torch.int32, device="cuda")
    l = vlib.load()
    for _ in range(3):
        vlib.check(l.vrad_lightmap_finalize(env._h, C.c_int64(n), vlib.ptr(d_dir), vlib.ptr(d_ind), vlib.ptr(d_out)))
        vlib.check(l.vrad_lightmap_finalize_patches(env._h, C.c_int64(n), vlib.ptr(d_dir), vlib.ptr(d_idx), C.c_int(200000), vlib.ptr(d_tot), vlib.ptr(d_out)))
elif what == "kdfast":
    from vrad_b200.environment import Environment
    s = scenes.outdoor()
    env = Environment(); env.add_triangles(s.tri_ids, s.tri_verts, s.tri_flags)
    print("binned device build seconds", env.build_fast(), env.stats())
    a, b = scenes.shadow_segments(s, 1 << 22)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.empty((1 << 22) // 32, dtype=torch.int32, device="cuda")
    for _ in range(2): env.test_lines(ta, tb, out=out)
elif what == "hier":
    s = scenes.multi_room_hier(nx=12, ny=11); t = s.meta["tree"]; env = environment_from_scene(s)
    env.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
    env.build_transfers(s.pvs)
    e0 = torch.full((s.n_patches, 3), 100.0, device="cuda")
    env.bounce(e0, 4)
else:
    s = scenes.multi_room(); env = environment_from_scene(s)
    nnz = env.build_transfers(s.pvs)
    e0 = torch.full((s.n_patches, 3), 100.0, device="cuda")
    env.bounce(e0, 6)
    pos, nrm = torch.from_numpy(s.patch_origin).cuda(), torch.from_numpy(s.patch_normal).cuda()
    env.direct_light(pos, nrm, s.lights)
torch.cuda.synchronize()
