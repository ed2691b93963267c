"""GPU tests of the binned-SAH kd build (vrad_env_build_fast, RTE_FLAGS_FAST_TREE_GENERATION): the device builds the same tree as the
host execution policy of the same source; K1 on that tree is bit-exact against the oracle's tracer walking the same tree; against the
exact builder's tree only knife-edge rays (a triangle edge lying in a split plane) may differ."""
import numpy as np
import pytest

from vrad_b200 import scenes

pytestmark = pytest.mark.gpu


def _env(scene, fast=None):
    from vrad_b200.environment import Environment
    g = Environment(0)
    g.add_triangles(scene.tri_ids, scene.tri_verts, scene.tri_flags)
    if fast is None:
        g.setup_acceleration_structure()
    else:
        g.build_fast(on_host=(fast == "host"))
    return g


@pytest.mark.parametrize("name", ["s1", "s2_small", "sky"])
def test_device_tree_equals_host_policy_tree(name):
    from vrad_b200.environment import kd_build_binned_host
    sc = {"s1": lambda: scenes.box_room(), "s2_small": lambda: scenes.multi_room(nx=3, ny=2, boxes_per_room=30), "sky": lambda: scenes.sky_room()}[name]()
    g = _env(sc, "device")
    t = g.download_tree(); ch, sp, idx = t["children"], t["split"], t["tri_index"]
    h = kd_build_binned_host(sc.tri_verts)
    assert np.array_equal(ch, h["children"]) and np.array_equal(sp.view(np.uint32), h["split"].view(np.uint32)) and np.array_equal(idx, h["tri_index"])
    st = g.stats()
    assert st["max_depth"] == h["max_depth"] and np.array_equal(st["aabb"], h["aabb"])
    gh = _env(sc, "host")                                              # the host policy through the environment: same arrays again
    t2 = gh.download_tree(); ch2, sp2, idx2 = t2["children"], t2["split"], t2["tri_index"]
    assert np.array_equal(ch, ch2) and np.array_equal(sp.view(np.uint32), sp2.view(np.uint32)) and np.array_equal(idx, idx2)
    g.close(); gh.close()


@pytest.mark.parametrize("name", ["s1", "s2_small"])
def test_k1_on_the_fast_tree_matches_the_oracle_on_the_same_tree(name):
    from oracle import pyoracle
    sc = {"s1": lambda: scenes.box_room(), "s2_small": lambda: scenes.multi_room(nx=3, ny=2, boxes_per_room=30)}[name]()
    g = _env(sc, "device")
    t = g.download_tree(); ch, sp, idx = t["children"], t["split"], t["tri_index"]
    o = pyoracle.env_from_scene(sc, with_patches=False)
    a, b = scenes.shadow_segments(sc, 1 << 17)
    exact_vis = o.test_lines(a, b, threads=8)
    o.replace_tree(ch, sp, idx, g.stats()["aabb"])
    vis = g.test_lines(a, b)
    assert np.array_equal(vis, o.test_lines(a, b, threads=8))          # bit-exact on the same tree
    r = scenes.random_rays(sc, 1 << 16)
    gt = g.trace_rays(r["o"], r["d"], r["tmax"]); ot = o.trace1(r["o"], r["d"], r["tmax"], threads=8)
    assert np.array_equal(gt[0], ot[0]) and np.array_equal(gt[2].view(np.uint32), ot[2].view(np.uint32))
    # against the exact builder's tree: only knife-edge segments may differ
    diff = np.unpackbits((vis ^ exact_vis).view(np.uint8)).sum()
    assert diff <= 16, diff                                            # of 131,072
    g.close()


def test_fast_build_at_c5_size():
    """1.03 M triangles: built on the device, validated, and K1 agrees with the exact tree's environment except on knife edges."""
    sc = scenes.outdoor()
    exact = _env(sc)
    fast = _env(sc, "device")
    st = fast.stats()
    print(f"\nC5 kd build: exact (host SAH) {exact.stats()['build_seconds']:.3f} s, binned on device {st['build_seconds']:.3f} s "
          f"({st['n_nodes']} nodes, {st['n_idx']} index entries, depth {st['max_depth']})")
    assert st["n_tris"] == sc.n_tris and st["max_depth"] <= 22 and st["n_idx"] >= sc.n_tris
    a, b = scenes.shadow_segments(sc, 1 << 20)
    ve, vf = exact.test_lines(a, b), fast.test_lines(a, b)
    diff = int(np.unpackbits((ve ^ vf).view(np.uint8)).sum())
    assert diff <= 64, diff                                            # of 1,048,576
    vis_frac = np.unpackbits(vf.view(np.uint8)).mean()
    assert 0.0 < vis_frac < 1.0
    exact.close(); fast.close()
