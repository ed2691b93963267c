"""GPU parity for the complete TestLineDoesHitSky surface and the BSP point queries:
every call goes through the C-ABI and must equal the CPU oracle (oracle/skytrace.cpp) bit for bit --
leaf numbers, clusters, camera tables and fractionVisible as raw float bits."""
import os

import numpy as np
import pytest

from vrad_b200 import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sky_scene():
    return scenes.sky_room()


@pytest.fixture(scope="module")
def sky_pair(sky_scene):
    from oracle import pyoracle
    from vrad_b200.environment import Environment
    m = sky_scene.meta
    g = Environment()
    g.add_triangles(sky_scene.tri_ids, sky_scene.tri_verts, sky_scene.tri_flags)
    g.set_triangle_colors(m["tri_colors"])                 # before the build: uploaded with the scene
    g.setup_acceleration_structure()
    g.bsp_upload(m["bsp"])
    o = pyoracle.OracleEnv()
    o.add_triangles(sky_scene.tri_ids, sky_scene.tri_verts, sky_scene.tri_flags)
    o.build(); o.set_triangle_colors(m["tri_colors"]); o.bsp_set(m["bsp"])
    assert g.process_sky_cameras(m["cams_origin"], m["cams_scale"]) == o.process_sky_cameras(m["cams_origin"], m["cams_scale"]) == 2
    yield g, o
    g.close()


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def test_point_queries_bit_exact(sky_scene, sky_pair):
    g, o = sky_pair
    rng = scenes.SplitMix64(123)
    n = 200000
    pts = np.stack([rng.uniform(n, -600, 9000), rng.uniform(n, -600, 4500), rng.uniform(n, -80, 600)], axis=1)
    # a share of points hugging the split planes (TEST_EPSILON logic of PointInLeaf, dist == 0 of PointLeafnum)
    k = n // 4
    pts[:k, 2] = rng.uniform(k, -0.06, 0.06)
    pts[k:2 * k, 0] = rng.uniform(k, -0.06, 0.06)
    pts[2 * k:2 * k + 1000] = np.array([0.0, 0.0, 0.0], np.float32)
    assert np.array_equal(g.point_leafnum(pts), o.point_leafnum(pts))
    assert np.array_equal(g.cluster_from_point(pts), o.cluster_from_point(pts))
    assert len(np.unique(g.point_leafnum(pts))) == 7


def test_sky_camera_tables(sky_pair):
    g, o = sky_pair
    for x, y in zip(g.sky_cameras(), o.sky_cameras()):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("flags,prop", [(0, -1), (0, 7), (1, 7), (2, 7), (3, 7), (5, 7), (7, -1)])
def test_fraction_visible_bit_exact(flags, prop, sky_scene, sky_pair):
    g, o = sky_pair
    n = (1 << 17) + 5
    a, b = scenes.sky_segments(sky_scene, n, seed=1000 + flags)
    gv = g.test_lines_sky(a, b, flags, prop)
    ov = o.test_lines_sky(a, b, flags, prop, threads=8)
    assert _same_bits(gv, ov)
    assert 0.03 < gv.mean() < 0.7
    if flags & 2:
        assert len(np.unique(gv)) > 6


@pytest.mark.parametrize("n", [1, 3, 31, 33, 129])
def test_ragged_sizes(n, sky_scene, sky_pair):
    g, o = sky_pair
    a, b = scenes.sky_segments(sky_scene, n, seed=n)
    assert _same_bits(g.test_lines_sky(a, b, 3, 7), o.test_lines_sky(a, b, 3, 7))
    assert _same_bits(g.test_lines_sky(a, b, 7, 7), o.test_lines_sky(a, b, 7, 7))


def test_mini_scene_known_answers_on_gpu():
    from vrad_b200.environment import Environment
    ids, verts, flags, cols = scenes.mini_sky_scene()
    g = Environment(); g.add_triangles(ids, verts, flags); g.setup_acceleration_structure()
    g.set_triangle_colors(cols)                            # after the build: uploaded on the spot
    for a, b, f, prop, want in scenes.MINI_SKY_CASES:
        got = g.test_lines_sky(np.array(a, np.float32).reshape(3, 1), np.array(b, np.float32).reshape(3, 1), f, prop)[0]
        assert got == np.float32(want), (a, b, f, prop, got, want)
    g.close()


def test_four_vectors_mirror(sky_scene, sky_pair):
    """trace.TestLineDoesHitSky host mirror: one FourVectors pair, leaf of lane 0 (testline.go:63)."""
    from vrad_b200.environment import test_line_does_hit_sky
    g, o = sky_pair
    a, b = scenes.sky_segments(sky_scene, 64, seed=8)
    for p in range(16):
        s, e = a[:, 4 * p:4 * p + 4], b[:, 4 * p:4 * p + 4]
        for rec in (True, False):
            got = test_line_does_hit_sky(g, s, e, can_recurse=rec, static_prop_to_skip=7, texture_shadows=True)
            want = o.test_lines_sky(np.ascontiguousarray(s), np.ascontiguousarray(e), 4 | 2 | (1 if rec else 0), 7)
            assert _same_bits(got, want)


def test_can_leaf_trace_to_sky(sky_scene, sky_pair):
    g, o = sky_pair
    dirs = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "vrad_b200", "data", "anorms.txt"), dtype=np.float32)
    g.set_sky_dirs(dirs)
    m = sky_scene.meta
    got = g.leafs_trace_to_sky(m["probe_mins"], m["probe_maxs"])
    want = o.leafs_trace_to_sky(m["probe_mins"], m["probe_maxs"], dirs, threads=8)
    assert np.array_equal(got, want)
    assert got[m["bsp"].leaf_mins.shape[0]] == 0 and got.sum() >= 3


def test_transparent_triangles_block_without_callback(sky_scene, sky_pair):
    """Trace4Rays / TestLine with callback == nil: a transparent triangle is an ordinary blocker."""
    g, o = sky_pair
    n = 1 << 16
    a, b = scenes.sky_segments(sky_scene, n, seed=2)
    for mode in (0, 1):
        assert np.array_equal(g.test_lines(a, b, sky_mode=mode), o.test_lines(a, b, sky_mode=mode, threads=8))
    r = scenes.random_rays(sky_scene, 1 << 15, seed=6)
    gt, gs, gd = g.trace_rays(r["o"], r["d"], r["tmax"])
    ot, os_, od = o.trace1(r["o"], r["d"], r["tmax"], threads=8)
    assert np.array_equal(gt, ot) and np.array_equal(gs, os_) and _same_bits(gd, od)


def test_device_resident_io(sky_scene, sky_pair):
    torch = pytest.importorskip("torch")
    g, o = sky_pair
    a, b = scenes.sky_segments(sky_scene, 50000, seed=3)
    out = g.test_lines_sky(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 3, 7)
    torch.cuda.synchronize()
    assert _same_bits(out.cpu().numpy(), o.test_lines_sky(a, b, 3, 7, threads=8))


def test_errors_are_loud(sky_scene):
    from vrad_b200.environment import Environment, VradError
    e = Environment()
    e.add_triangles(sky_scene.tri_ids, sky_scene.tri_verts, sky_scene.tri_flags)
    with pytest.raises(VradError):
        e.set_triangle_colors(np.zeros((3, 3), np.float32))            # wrong count
    e.setup_acceleration_structure()
    with pytest.raises(VradError):
        e.point_leafnum(np.zeros((4, 3), np.float32))                  # no BSP uploaded
    with pytest.raises(VradError):
        e.process_sky_cameras(np.zeros((1, 3), np.float32), np.ones(1, np.float32))
    bsp = sky_scene.meta["bsp"]
    bad = scenes.Bsp(bsp.node_plane, bsp.node_children.copy(), bsp.plane_normal, bsp.plane_dist, bsp.plane_type,
                     bsp.leaf_cluster, bsp.leaf_area, bsp.leaf_mins, bsp.leaf_maxs, bsp.n_areas)
    bad.node_children[3, 0] = 1                                        # child before its parent: a cycle
    with pytest.raises(VradError):
        e.bsp_upload(bad)
    with pytest.raises(VradError):
        e.leafs_trace_to_sky(np.zeros((1, 3), np.int16), np.ones((1, 3), np.int16))   # no sky directions set
    # without BSP/cameras the recursion flag is a no-op, not an error
    a, b = scenes.sky_segments(sky_scene, 100, seed=1)
    assert np.array_equal(e.test_lines_sky(a, b, 1, -1), e.test_lines_sky(a, b, 0, -1))
    e.close()


def _floor_luxels_and_lights():
    xs = np.arange(-496, 512, 32, dtype=np.float32)
    gx, gy = np.meshgrid(xs, xs, indexing="ij")
    pos = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, 1.0, np.float32)], axis=1)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (pos.shape[0], 1))
    L = np.zeros(4, dtype=scenes.LIGHT_DTYPE)
    L["start_fade"], L["end_fade"], L["cap_dist"] = 0.0, -1.0, 1.0e22
    sun = np.array([-0.4, -0.3, -0.87]); sun /= np.linalg.norm(sun)
    L[0]["type"] = scenes.EMIT_SKYLIGHT; L[0]["normal"] = sun; L[0]["intensity"] = (300, 280, 250)
    L[1]["type"] = scenes.EMIT_SKYAMBIENT; L[1]["intensity"] = (40, 50, 70)
    L[2]["type"] = scenes.EMIT_POINT; L[2]["origin"] = (100, 50, 500); L[2]["quadratic_attn"] = 1.0; L[2]["intensity"] = (4e6, 4e6, 4e6)
    L[3]["type"] = scenes.EMIT_SPOTLIGHT; L[3]["origin"] = (-200, -100, 505); L[3]["normal"] = (0, 0, -1)
    L[3]["stopdot"], L[3]["stopdot2"], L[3]["exponent"] = 0.8, 0.6, 1.0
    L[3]["quadratic_attn"] = 1.0; L[3]["intensity"] = (3e6, 2e6, 1e6)
    return pos, nrm, L


def test_direct_light_with_complete_sky_test(sky_scene, sky_pair):
    """K3 light rays through the complete TestLineDoesHitSky (vrad_set_light_trace_flags): sun + sky ambient with the
    3D-skybox recursion, every light with transparent-triangle coverage (dot *= fractionVisible).  Exponent 1 and no
    powf anywhere -> bit-exact against the oracle."""
    g, o = sky_pair
    dirs = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "vrad_b200", "data", "anorms.txt"), dtype=np.float32)
    g.set_sky_dirs(dirs); o.set_sky_dirs(dirs)
    pos, nrm, L = _floor_luxels_and_lights()
    res = {}
    try:
        for flags in (0, 1, 2, 3):
            g.set_light_trace_flags(flags); o.set_light_trace_flags(flags)
            gl = g.direct_light(pos, nrm, L)
            ol = o.direct_light(pos, nrm, L, threads=8)
            assert _same_bits(gl, ol), flags
            res[flags] = gl
    finally:
        g.set_light_trace_flags(0); o.set_light_trace_flags(0)
    assert (res[0] != res[1]).any(axis=1).sum() > 50          # the sky boxes take light away
    assert np.all(res[1].sum(axis=1) <= res[0].sum(axis=1) + 1e-3)
    assert (res[2].sum(axis=1) >= res[0].sum(axis=1) - 1e-3).all() and res[2].mean() > 2 * res[0].mean()   # panes stop being hard blockers
    from vrad_b200.environment import VradError
    with pytest.raises(VradError):
        g.set_light_trace_flags(4)                            # PACKET_LEAF has no meaning for light rays


def test_empty_inputs(sky_scene, sky_pair):
    g, _ = sky_pair
    z = np.zeros((3, 0), np.float32)
    assert g.test_lines_sky(z, z, 3, 7).shape == (0,)
    assert g.point_leafnum(np.zeros((0, 3), np.float32)).shape == (0,)
    assert g.cluster_from_point(np.zeros((0, 3), np.float32)).shape == (0,)
    assert g.leafs_trace_to_sky(np.zeros((0, 3), np.int16), np.zeros((0, 3), np.int16)).shape == (0,)
    assert g.process_sky_cameras(np.zeros((0, 3), np.float32), np.zeros(0, np.float32)) == 0      # no cameras: recursion becomes a no-op
    a = np.array([[0.0], [0.0], [100.0]], np.float32); b = np.array([[0.0], [0.0], [5000.0]], np.float32)
    assert g.test_lines_sky(a, b, 1, 7)[0] == g.test_lines_sky(a, b, 0, 7)[0]
    assert g.process_sky_cameras(sky_scene.meta["cams_origin"], sky_scene.meta["cams_scale"]) == 2   # restore the module fixture
