"""CPU tests of the binned-SAH kd builder (vrad_b200/csrc/kd_fast.cu; RTE_FLAGS_FAST_TREE_GENERATION, raytracer/constants.go:5), run
through its host execution policy -- the same functors the device kernels run: the tree is a valid tree in the reference's packed
layout, holds every triangle wherever its box reaches, and the oracle's tracer finds the same hits walking it as walking the
exact builder's tree."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle
from vrad_b200 import scenes
from vrad_b200.environment import kd_build_binned_host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COST_TRAVERSAL, COST_INTERSECTION, MAX_DEPTH = 75.0, 167.0, 21      # raytracer/kdtree/constants.go:25-28


def _walk(tr, strict=True):
    """(leaves: list of (lo, hi, triangle ids, depth)), SAH cost, interior count -- by walking the packed arrays."""
    ch, sp, idx, aabb = tr["children"], tr["split"], tr["tri_index"], tr["aabb"].astype(np.float64)

    def area(lo, hi):
        d = hi - lo
        return 2 * (d[0] * d[1] + d[1] * d[2] + d[2] * d[0])
    root = area(aabb[:3], aabb[3:])
    leaves, cost, interior = [], 0.0, 0
    seen = np.zeros(ch.shape[0], bool)
    stack = [(0, aabb[:3].copy(), aabb[3:].copy(), 0)]
    while stack:
        n, lo, hi, depth = stack.pop()
        assert not seen[n]
        seen[n] = True
        axis = int(ch[n]) & 3
        if axis == 3:
            cnt, start = int(sp[n]), int(np.uint32(ch[n]) >> 2)
            assert sp[n] == cnt and start + cnt <= idx.shape[0]
            leaves.append((lo, hi, idx[start:start + cnt], depth))
            cost += COST_INTERSECTION * cnt * area(lo, hi) / root
        else:
            interior += 1
            left = int(np.uint32(ch[n]) >> 2)
            assert (lo[axis] < sp[n] < hi[axis]) if strict else (lo[axis] <= sp[n] <= hi[axis])     # the plane cuts the node's box
            cost += COST_TRAVERSAL * area(lo, hi) / root
            lhi, rlo = hi.copy(), lo.copy()
            lhi[axis] = sp[n]; rlo[axis] = sp[n]
            stack.append((left, lo, lhi, depth + 1)); stack.append((left + 1, rlo, hi, depth + 1))   # right child = left + 1
    assert seen.all()                                                  # no orphan nodes
    return leaves, cost, interior


@pytest.fixture(scope="module", params=["s1", "s2_small", "sky"])
def scene(request):
    return {"s1": lambda: scenes.box_room(), "s2_small": lambda: scenes.multi_room(nx=3, ny=2, boxes_per_room=30),
            "sky": lambda: scenes.sky_room()}[request.param]()


def test_binned_tree_is_valid_and_complete(scene):
    tr = kd_build_binned_host(scene.tri_verts)
    leaves, cost, interior = _walk(tr)
    assert len(leaves) == interior + 1 and tr["children"].shape[0] == 2 * interior + 1
    assert max(d for *_, d in leaves) == tr["max_depth"] <= MAX_DEPTH + 1
    v = scene.tri_verts.reshape(-1, 3, 3).astype(np.float64)
    tmin, tmax = v.min(axis=1), v.max(axis=1)
    assert np.allclose(tr["aabb"][:3], tmin.min(axis=0)) and np.allclose(tr["aabb"][3:], tmax.max(axis=0))
    # every triangle is listed in every leaf its bounding box reaches into (open overlap), and in no leaf it does not touch
    covered = np.zeros(v.shape[0], np.int64)
    for lo, hi, tris, _ in leaves:
        assert len(set(tris.tolist())) == len(tris)
        inside = np.all((tmin < hi) & (tmax > lo), axis=1)
        degenerate = np.any(tmin == tmax, axis=1)                      # flat boxes (axis-aligned triangles) sit on planes: closed test
        touching = np.all((tmin <= hi) & (tmax >= lo), axis=1)
        listed = np.zeros(v.shape[0], bool); listed[tris] = True
        assert np.all(listed[inside & ~degenerate]) and np.all(touching[tris])
        covered[tris] += 1
    assert np.all(covered >= 1)
    # the reference's leaf rule: a node that stayed a leaf with >= 3 triangles above the depth limit found no cheaper split; and the
    # whole tree is not worse, by the reference's own cost model, than the exact builder's
    o = pyoracle.env_from_scene(scene, with_patches=False)
    _, exact_cost, _ = _walk(o.export(), strict=False)          # the exact builder may put a plane on the box face
    assert cost <= 1.05 * exact_cost


def test_oracle_finds_the_same_hits_walking_it(scene):
    tr = kd_build_binned_host(scene.tri_verts)
    o = pyoracle.env_from_scene(scene, with_patches=False)
    r = scenes.random_rays(scene, 1 << 15)
    a = np.stack([r["o"][0], r["o"][1], r["o"][2]]); b = (a + np.stack(r["d"]) * np.minimum(r["tmax"], 4000.0)).astype(np.float32)
    exact = o.trace1(r["o"], r["d"], r["tmax"], threads=8)
    brute = o.trace_brute(r["o"], r["d"], r["tmax"], threads=8)
    vis_exact = o.test_lines(a, b, threads=8)
    o.replace_tree(tr["children"], tr["split"], tr["tri_index"], tr["aabb"])
    fast = o.trace1(r["o"], r["d"], r["tmax"], threads=8)
    vis_fast = o.test_lines(a, b, threads=8)
    # which triangle a ray hits does not depend on the tree -- except where a ray grazes a triangle edge lying in a split plane
    # (knife edge: a leaf touched in a single point may or may not be visited).  Random rays do not do that:
    assert np.array_equal(fast[0], exact[0]) and np.array_equal(fast[2].view(np.uint32), exact[2].view(np.uint32))
    assert np.array_equal(fast[0], brute[0]) and np.array_equal(fast[2].view(np.uint32), brute[2].view(np.uint32))
    assert np.array_equal(vis_fast, vis_exact)
    assert (fast[0] >= 0).mean() > 0.02


def test_knife_edge_rays_are_the_only_difference():
    """Shadow segments between points on a 32-unit grid do graze edges that lie in split planes; the two trees may then disagree on
    a handful of them, and the brute-force tracer shows those are edge hits (the hit sits on a triangle border)."""
    sc = scenes.multi_room(nx=3, ny=2, boxes_per_room=30)
    tr = kd_build_binned_host(sc.tri_verts)
    o = pyoracle.env_from_scene(sc, with_patches=False)
    a, b = scenes.shadow_segments(sc, 1 << 16)
    v_exact = np.unpackbits(o.test_lines(a, b, threads=8).view(np.uint8), bitorder="little")
    o.replace_tree(tr["children"], tr["split"], tr["tri_index"], tr["aabb"])
    v_fast = np.unpackbits(o.test_lines(a, b, threads=8).view(np.uint8), bitorder="little")
    diff = np.nonzero(v_exact != v_fast)[0]
    assert len(diff) <= 8                                               # of 65,536
    v = sc.tri_verts.reshape(-1, 3, 3).astype(np.float64)
    for i in diff:
        d = (b[:, i] - a[:, i]).astype(np.float64); L = np.linalg.norm(d)
        h = o.trace_brute(a[:, i:i + 1].copy(), (d / L).astype(np.float32).reshape(3, 1), np.float32([L]))
        if h[0][0] < 0:
            continue                                                    # the brute tracer's ray (direction rounded differently) slips past
        p = a[:, i] + d / L * float(h[2][0])
        t = v[h[0][0]]
        n = np.cross(t[1] - t[0], t[2] - t[0]); area2 = np.linalg.norm(n)
        bary = [np.linalg.norm(np.cross(t[(k + 1) % 3] - p, t[(k + 2) % 3] - p)) / area2 for k in range(3)]
        assert min(bary) < 1e-3, (i, bary)                             # on an edge of the triangle


def test_small_and_degenerate_inputs():
    tri = np.float32([[0, 0, 0, 1, 0, 0, 0, 1, 0]])
    for n in (0, 1, 2):
        tr = kd_build_binned_host(np.repeat(tri, n, axis=0))
        assert tr["children"].shape[0] == 1 and (tr["children"][0] & 3) == 3 and tr["split"][0] == n      # n < 3: one leaf (environment.go:241)
        assert list(tr["tri_index"]) == list(range(n))
    # 200 copies of one triangle cannot be separated: the cost rule keeps them in one leaf (or a few, never past the depth limit)
    tr = kd_build_binned_host(np.repeat(tri, 200, axis=0))
    leaves, _, _ = _walk(tr)
    assert tr["max_depth"] <= MAX_DEPTH + 1 and max(len(t) for _, _, t, _ in leaves) == 200
    # a flat scene (every triangle in z = 0): the box has no thickness on one axis, the other two still split
    rng = np.random.default_rng(4)
    xy = rng.uniform(0, 1000, (500, 1, 2)) + rng.uniform(-10, 10, (500, 3, 2))
    flat = np.concatenate([xy, np.zeros((500, 3, 1))], axis=2).astype(np.float32).reshape(500, 9)
    tr = kd_build_binned_host(flat)
    leaves, _, interior = _walk(tr)
    assert interior > 50 and all(int(c) & 3 in (0, 1, 3) for c in tr["children"])


def test_tree_does_not_depend_on_the_thread_count():
    """Integer bin counters and stable scans: the result is the same with one thread and with all of them."""
    code = ("import sys, hashlib, numpy as np; sys.path.insert(0, %r); from vrad_b200 import scenes; "
            "from vrad_b200.environment import kd_build_binned_host; "
            "t = kd_build_binned_host(scenes.multi_room(nx=2, ny=2, boxes_per_room=20).tri_verts); "
            "print(hashlib.sha256(t['children'].tobytes() + t['split'].tobytes() + t['tri_index'].tobytes()).hexdigest())") % ROOT
    digests = set()
    for threads in ("1", "3", "8"):
        env = dict(os.environ, OMP_NUM_THREADS=threads)
        digests.add(subprocess.run([sys.executable, "-c", code], check=True, capture_output=True, text=True, env=env).stdout.strip())
    assert len(digests) == 1 and len(digests.pop()) == 64
