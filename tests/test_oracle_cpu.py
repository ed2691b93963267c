"""CPU-only tests: the oracle against analytic known answers, its own brute force, and the
structural invariants of the kd build.  PARITY UNPINNED: the reference ships no golden vectors
for this path, so the arbiters are closed-form answers and brute force."""
import math

import numpy as np
import pytest

from oracle import pyoracle
from vrad_b200 import scenes


def _bits(a, n):
    return np.unpackbits(np.ascontiguousarray(a).view(np.uint8), bitorder="little")[:n]


def test_scene_s1_shape(s1_scene):
    assert s1_scene.n_tris == 996 and s1_scene.n_patches == 4096
    assert s1_scene.luxel_pos.shape == (16384, 3) and len(s1_scene.lights) == 8
    assert np.all(s1_scene.tri_ids & scenes.TRACE_ID_OPAQUE)
    again = scenes.box_room()
    assert np.array_equal(again.tri_verts, s1_scene.tri_verts) and np.array_equal(again.patch_refl, s1_scene.patch_refl)


def test_kd_tree_invariants(s1_scene, s1_oracle):
    t = s1_oracle.export()
    ch, sp, idx = t["children"], t["split"], t["tri_index"]
    leaf = (ch & 3) == 3
    assert leaf.sum() == s1_oracle.sizes()["n_leaves"]
    # optimisedkdnode.go: right child = left + 1, children appended adjacently, every node reachable once
    seen = np.zeros(len(ch), bool); stack = [0]
    covered = np.zeros(s1_scene.n_tris, bool)
    while stack:
        n = stack.pop()
        assert not seen[n]; seen[n] = True
        if leaf[n]:
            start, cnt = ch[n] >> 2, int(sp[n])
            assert sp[n] == float(cnt) and 0 <= start and start + cnt <= len(idx)
            covered[idx[start:start + cnt]] = True
        else:
            l = ch[n] >> 2
            assert 0 < l < len(ch) - 1
            stack += [l, l + 1]
    assert seen.all() and covered.all()
    assert s1_oracle.sizes()["max_depth"] <= 22                              # MAX_TREE_DEPTH 21 (+1 for the leaves)
    assert np.array_equal(t["aabb"], np.array([-512, -512, 0, 512, 512, 512], np.float32))


def test_intersection_format_known_answer():
    """Triangle (0,0,0),(1,0,0),(0,1,0): N=(0,0,1), D=0, drop z, keep (x,y); the two edge equations are
    barycentric coordinates: edge(p1,p2) = y (1 at p3), edge(p2,p3) = 1-x-y (1 at p1)."""
    o = pyoracle.OracleEnv()
    o.add_triangles([7, 8, 9], np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0], [5, 5, 5, 6, 5, 5, 5, 5, 9],
                                          [0, 0, 0, 0, 0, 1, 0, 1, 0]], np.float32))
    o.build()
    t = o.export()["tris"][0]
    assert tuple(t["n"]) == (0.0, 0.0, 1.0) and t["d"] == 0.0 and t["id"] == 7
    assert (t["sel0"], t["sel1"]) == (0, 1)
    assert np.allclose(t["e"], [0, 1, 0, -1, -1, 1])
    t2 = o.export()["tris"][2]               # normal along -x -> drop x, keep (y,z)
    assert abs(t2["n"][0]) == 1.0 and (t2["sel0"], t2["sel1"]) == (1, 2)


def test_axis_ray_known_answers(s1_scene, s1_oracle):
    """Rays along the room axes from a point above every occluder hit the walls at the analytic distance."""
    o = np.array([[10.0] * 6, [20.0] * 6, [300.0] * 6], np.float32)
    d = np.array([[1, -1, 0, 0, 0, 0], [0, 0, 1, -1, 0, 0], [0, 0, 0, 0, 1, -1]], np.float32)
    tmax = np.full(6, scenes.MAX_TRACE_LENGTH, np.float32)
    for fn in (s1_oracle.trace1, s1_oracle.trace4, s1_oracle.trace_brute):
        tri, sid, t = fn(o, d, tmax)
        assert np.allclose(t[:5], [502, 522, 492, 532, 212], rtol=1e-6)
        assert np.all(tri[:5] >= 0) and np.all(sid[:5] & scenes.TRACE_ID_OPAQUE)
    # down-ray: floor at t=300 unless a box (height <= 96) is below
    assert 204 <= t[5] <= 300


def test_kd_vs_brute_force_and_packet(s1_scene, s1_oracle, s2_small_scene, s2_small_oracle):
    for scene, orc in ((s1_scene, s1_oracle), (s2_small_scene, s2_small_oracle)):
        r = scenes.random_rays(scene, 30000, seed=123)
        b = orc.trace_brute(r["o"], r["d"], r["tmax"], threads=8)
        k = orc.trace1(r["o"], r["d"], r["tmax"])
        p = orc.trace4(r["o"], r["d"], r["tmax"], threads=8)
        assert np.array_equal(b[0], k[0]) and np.array_equal(b[2].view(np.uint32), k[2].view(np.uint32))
        assert np.array_equal(p[0], k[0]) and np.array_equal(p[2].view(np.uint32), k[2].view(np.uint32))
        c = orc.counters()
        assert c["nodes"] / 30000 < 60 and c["tris"] / 30000 < 40          # the tree actually prunes


def test_coherent_packets_match_single_rays(s1_scene, s1_oracle):
    """FourRays packets whose lanes share direction signs go through the coherent packet traversal."""
    n = 4096
    rng = scenes.SplitMix64(42)
    o = np.stack([rng.uniform(n, -400, 400), rng.uniform(n, -400, 400), rng.uniform(n, 120, 480)])
    d = np.stack([rng.uniform(n, 0.1, 1.0), rng.uniform(n, 0.1, 1.0), rng.uniform(n, -1.0, -0.1)])
    d /= np.linalg.norm(d, axis=0, keepdims=True)
    tmax = np.full(n, 5000, np.float32)
    k = s1_oracle.trace1(o, d.astype(np.float32), tmax)
    p = s1_oracle.trace4(o, d.astype(np.float32), tmax)
    assert np.array_equal(p[0], k[0]) and np.array_equal(p[2].view(np.uint32), k[2].view(np.uint32))


def test_test_lines_modes_agree(s1_scene, s1_oracle):
    n = 50001
    a, b = scenes.shadow_segments(s1_scene, n)
    v0 = s1_oracle.test_lines(a, b, threads=8)
    assert np.array_equal(v0, s1_oracle.test_lines(a, b, mode=1, threads=8))
    assert np.array_equal(v0, s1_oracle.test_lines(a, b, mode=2, threads=8))
    vis = _bits(v0, n).mean()
    assert 0.4 < vis < 0.95
    assert _bits(v0, ((n + 31) // 32) * 32)[n:].sum() == 0                   # tail bits clear


def test_parallel_plates_form_factor():
    """Two parallel unit-ish plates far apart: differential form factor cos*cos*A/(pi r^2)."""
    o = pyoracle.OracleEnv()
    o.add_triangles([scenes.TRACE_ID_OPAQUE], np.array([[1e4, 1e4, 1e4, 1e4 + 1, 1e4, 1e4, 1e4, 1e4 + 1, 1e4]], np.float32))
    o.build()
    origin = np.array([[0, 0, 0], [0, 0, 100], [60, 0, 100]], np.float32)
    normal = np.array([[0, 0, 1], [0, 0, -1], [0, 0, -1]], np.float32)
    pd = np.array([0, -100, -100], np.float32)
    area = np.array([4, 9, 16], np.float32)
    o.patches_upload(origin, normal, pd, area, np.full((3, 3), 0.5, np.float32))
    assert o.build_transfers() == 4                                          # 0<->1, 0<->2; 1,2 coplanar never see each other
    rp, col, w = o.transfers()
    assert list(rp) == [0, 2, 3, 4] and list(col) == [1, 2, 0, 0]
    r2 = 60.0 ** 2 + 100.0 ** 2
    cos = 100.0 / math.sqrt(r2)
    assert np.allclose(w, [9 / (math.pi * 1e4), 16 * cos * cos / (math.pi * r2), 4 / (math.pi * 1e4), 4 * cos * cos / (math.pi * r2)], rtol=1e-5)


def test_closed_room_energy(s1_scene, s1_oracle):
    """Closed box, uniform reflectivity rho, uniform emission E: after MakeScales every row sums to
    <= 1, so each bounce adds at most rho * previous; the total converges below E*rho/(1-rho)."""
    sel = slice(0, None, 4)
    n = s1_scene.patch_origin[sel].shape[0]
    rho = 0.5
    s1_oracle.patches_upload(s1_scene.patch_origin[sel], s1_scene.patch_normal[sel], s1_scene.patch_plane_dist[sel],
                             s1_scene.patch_area[sel], np.full((n, 3), rho, np.float32))
    nnz = s1_oracle.build_transfers(threads=8)
    rp, col, w = s1_oracle.transfers()
    assert nnz == rp[-1] and np.all(np.diff(rp) >= 0)
    rows = np.repeat(np.arange(n), np.diff(rp))
    assert np.all(rows != col)
    sums = np.bincount(rows, weights=w.astype(np.float64), minlength=n)
    assert sums.max() <= 1.0 + 1e-5
    E = 100.0
    prev = None
    for nb in (1, 2, 3, 30):
        total, added, done = s1_oracle.bounce(np.full((n, 3), E, np.float32), nb, threads=8)
        assert done == nb and np.all(total >= 0) and total.max() <= E * rho / (1 - rho) * (1 + 1e-4)
        if prev is not None:
            assert np.all(total >= prev - 1e-3)
        prev = total
    t_inf, added, done = s1_oracle.bounce(np.full((n, 3), E, np.float32), 200, early_out=True, threads=8)
    assert done < 200 and np.all(added < 1.0)
    # restore the fixture's patches for other tests
    s1_oracle.patches_upload(s1_scene.patch_origin, s1_scene.patch_normal, s1_scene.patch_plane_dist, s1_scene.patch_area,
                             s1_scene.patch_refl, s1_scene.patch_cluster, s1_scene.patch_flags)


def test_point_light_inverse_square():
    """Unoccluded floor under a point light: rgb = I * cos / (q d^2) (App. B.2)."""
    o = pyoracle.OracleEnv()
    o.add_triangles([scenes.TRACE_ID_OPAQUE] * 2, np.array([[-1e3, -1e3, -1, 1e3, -1e3, -1, 1e3, 1e3, -1],
                                                          [-1e3, -1e3, -1, 1e3, 1e3, -1, -1e3, 1e3, -1]], np.float32))
    o.build()
    L = np.zeros(1, scenes.LIGHT_DTYPE)
    L[0]["type"] = scenes.EMIT_POINT; L[0]["origin"] = (0, 0, 200); L[0]["intensity"] = (1e6, 2e6, 3e6)
    L[0]["quadratic_attn"] = 1.0; L[0]["end_fade"] = -1.0; L[0]["cap_dist"] = 1e22
    pos = np.array([[0, 0, 0], [150, 0, 0], [0, 0, -5]], np.float32)
    nrm = np.array([[0, 0, 1]] * 3, np.float32)
    rgb = o.direct_light(pos, nrm, L)
    assert np.allclose(rgb[0], np.array([1e6, 2e6, 3e6]) / 200 ** 2, rtol=1e-5)
    d2 = 150 ** 2 + 200 ** 2
    assert np.allclose(rgb[1], np.array([1e6, 2e6, 3e6]) * (200 / math.sqrt(d2)) / d2, rtol=1e-5)
    assert np.all(rgb[2] == 0)                                               # below the floor: shadowed


def test_gather_rows_matches_numpy():
    rng = np.random.default_rng(0)
    N = 200
    lens = rng.integers(0, 30, N)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    col = rng.integers(0, N, rp[-1]).astype(np.int32); w = rng.random(rp[-1]).astype(np.float32)
    emit = rng.random((N, 3)).astype(np.float32); refl = rng.random((N, 3)).astype(np.float32)
    out = pyoracle.gather_rows(0, N, rp, col, w, emit, refl)
    er = (emit * refl).astype(np.float64)
    ref = np.stack([np.bincount(np.repeat(np.arange(N), lens), weights=w * er[col, c], minlength=N) for c in range(3)], axis=1)
    assert np.allclose(out, ref, rtol=1e-5, atol=1e-6)
    part = pyoracle.gather_rows(50, 120, rp, col, w, emit, refl, threads=4)
    assert np.array_equal(part, out[50:120])
