"""CPU-only tests of the complete TestLineDoesHitSky surface in the oracle (oracle/skytrace.cpp;
SURVEY.md section 8 f2) and of the PVS run-length code (f3).

PARITY UNPINNED (no golden vectors in the reference): the arbiters here are hand-derived known answers
and an independent numpy composition of the oracle's BRUTE-FORCE closest-hit tracer that restates
raytracer/trace/testline.go:42-93 array-wise."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle
from vrad_b200 import scenes

MAXLEN = np.float32(1.732050807569 * 32768.0)


@pytest.fixture(scope="module")
def sky_scene():
    return scenes.sky_room()


@pytest.fixture(scope="module")
def sky_oracle(sky_scene):
    o = pyoracle.OracleEnv()
    o.add_triangles(sky_scene.tri_ids, sky_scene.tri_verts, sky_scene.tri_flags)
    o.build()
    m = sky_scene.meta
    o.set_triangle_colors(m["tri_colors"])
    o.bsp_set(m["bsp"])
    assert o.process_sky_cameras(m["cams_origin"], m["cams_scale"]) == 2
    return o


def test_point_leafnum_known_answers(sky_oracle):
    """pointleaf.go:8-33 -- leaves in creation order: 0/1 sky boxes, 2 solid, 3..6 world."""
    pts = np.array([[0, 0, 100], [100, 50, 100], [-10, 10, 5], [-10, -10, 5], [0, 0, -5], [8192, 0, 5], [8192, 4096, 5],
                    [300, 300, 10], [0, 0, 0]], np.float32)
    # x >= 0 -> diagonal plane 0.6x+0.8y-100: (0,0) and (100,50)=100 -> exactly 0 -> front; (300,300) front
    assert list(sky_oracle.point_leafnum(pts)) == [4, 3, 5, 6, 2, 1, 0, 3, 4]
    # ClusterFromPoint: inside TEST_EPSILON of the floor plane the front side wins unless its cluster is -1
    near = np.array([[0, 0, 0.01], [0, 0, -0.01], [0, 0, -0.5], [-5, 5, 0.0]], np.float32)
    assert list(sky_oracle.cluster_from_point(near)) == [1, 1, -1, 2]
    assert list(sky_oracle.point_leafnum(near)) == [4, 2, 2, 5]


def test_process_sky_cameras(sky_oracle):
    """skycamera.go:10-49: the scale-0 entity is dropped, areas 2 and 3 get cameras 0 and 1."""
    cam_area, w2s, area_cam = sky_oracle.sky_cameras()
    assert list(cam_area) == [2, 3]
    assert list(w2s) == [np.float32(1 / 16), np.float32(1 / 32)]
    assert list(area_cam) == [-1, -1, 0, 1]


def test_mini_scene_known_answers():
    ids, verts, flags, cols = scenes.mini_sky_scene()
    o = pyoracle.OracleEnv()
    o.add_triangles(ids, verts, flags); o.build(); o.set_triangle_colors(cols)
    for a, b, f, prop, want in scenes.MINI_SKY_CASES:
        got = o.test_lines_sky(np.array(a, np.float32).reshape(3, 1), np.array(b, np.float32).reshape(3, 1), f, prop)[0]
        assert got == np.float32(want), (a, b, f, prop, got, want)


def _brute_fraction(o, scene, a, b, recurse, skip):
    """testline.go:22-93 restated array-wise on top of the brute-force tracer (no kd tree, no coverage)."""
    ids = scene.tri_ids
    bsp = scene.meta["bsp"]

    def occlusion(a, b):
        d = b - a
        len2 = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]
        ok = len2 != 0
        ln = np.sqrt(len2, dtype=np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = (np.float32(1.0) / ln).astype(np.float32)
            dn = (d * inv).astype(np.float32)
        dn[:, ~ok] = 1.0
        tri, _, t = o.trace_brute(a, np.ascontiguousarray(dn), ln, skip_id=skip, threads=8)
        occ = ((tri != -1) & (t < ln) & ((ids[np.maximum(tri, 0)] & scenes.TRACE_ID_SKY) == 0) & ok).astype(np.float32)
        return occ, ok

    occ, ok = occlusion(a, b)
    if recurse:
        cam_area, w2s, area_cam = o.sky_cameras()
        leaf = o.point_leafnum(a.T.copy())
        area = bsp.leaf_area[leaf]
        rec = ok & (occ < 1) & (area_cam[area] < 0)
        d = (b - a).astype(np.float32)
        magsq = d[0] * d[0]; magsq = d[1] * d[1] + magsq; magsq = d[2] * d[2] + magsq
        with np.errstate(divide="ignore", invalid="ignore"):
            rs = (1.0 / np.sqrt(magsq.astype(np.float64))).astype(np.float32)
            dn = (d * rs).astype(np.float32)
        cams = scene.meta["cams_origin"][scene.meta["cams_scale"] > 0]
        for c in range(len(w2s)):
            s0 = (cams[c][:, None] + a * w2s[c]).astype(np.float32)
            s1 = (dn * MAXLEN + s0).astype(np.float32)
            occ2, ok2 = occlusion(np.ascontiguousarray(s0[:, rec]), np.ascontiguousarray(s1[:, rec]))
            fv2 = np.where(ok2, 1 - np.clip(occ2, 0, 1), 1).astype(np.float32)
            occ[rec] = (occ[rec] + np.float32(1.0)) - fv2
    return (1 - np.clip(occ, 0, 1)).astype(np.float32)


@pytest.mark.parametrize("recurse", [False, True])
def test_sky_fraction_matches_brute_force_composition(recurse, sky_scene, sky_oracle):
    n = 20000
    a, b = scenes.sky_segments(sky_scene, n, seed=77)
    skip = scenes.TRACE_ID_STATICPROP | 7
    got = sky_oracle.test_lines_sky(a, b, flags=1 if recurse else 0, static_prop_to_skip=7, threads=8)
    want = _brute_fraction(sky_oracle, sky_scene, a, b, recurse, skip)
    assert np.array_equal(got, want)
    assert 0.05 < got.mean() < 0.6
    if recurse:
        flat = sky_oracle.test_lines_sky(a, b, flags=0, static_prop_to_skip=7, threads=8)
        assert (flat != got).sum() > 100 and np.all(got <= flat)          # the sky boxes only ever add occlusion
        # segments that start inside sky box 0 (its area has a camera) never recurse
        leaf = sky_oracle.point_leafnum(a.T.copy())
        assert np.array_equal(got[leaf == 1], flat[leaf == 1])


def test_no_recursion_equals_test_lines_sky_mode(sky_scene, sky_oracle):
    """flags=0 and no prop to skip is exactly the sky_mode=1 TestLine (testline.go:42-51)."""
    n = 30000
    a, b = scenes.sky_segments(sky_scene, n, seed=5)
    bits = sky_oracle.test_lines(a, b, sky_mode=1, threads=8)
    vis = np.unpackbits(bits.view(np.uint8), bitorder="little")[:n].astype(np.float32)
    # static_prop_to_skip = -1 makes the skip id all ones, which no triangle carries
    assert np.array_equal(sky_oracle.test_lines_sky(a, b, flags=0, static_prop_to_skip=-1, threads=8), vis)


def test_texture_shadow_coverage(sky_scene, sky_oracle):
    """coverageCount.go:16-48: occlusion = max(opaque hit, min(1, sum of colour.X of the panes crossed)).
    Independent sum: all transparent triangles tested directly with the TriIntersectData equations."""
    n = 20000
    a, b = scenes.sky_segments(sky_scene, n, seed=31)
    m = sky_scene.meta
    flat = sky_oracle.test_lines_sky(a, b, flags=0, static_prop_to_skip=7, threads=8)
    got = sky_oracle.test_lines_sky(a, b, flags=2, static_prop_to_skip=7, threads=8)
    tris = sky_oracle.export()["tris"]
    d = (b - a).astype(np.float64); ln = np.sqrt((d * d).sum(0)); ok = ln > 0
    dn = np.where(ok, d / np.where(ok, ln, 1), 0)
    cov = np.zeros(n)
    # without a callback panes block like opaque triangles: nearest opaque hit = brute force over the non-pane triangles
    for ti in range(m["n_opaque"], m["n_world"]):
        T = tris[ti]
        nrm = T["n"].astype(np.float64)
        ddn = dn.T @ nrm
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (float(T["d"]) - a.T.astype(np.float64) @ nrm) / ddn
            p = a.astype(np.float64) + dn * t
        c0, c1 = p[T["sel0"]], p[T["sel1"]]
        e = T["e"].astype(np.float64)
        b0 = e[0] * c0 + e[1] * c1 + e[2]; b1 = e[3] * c0 + e[4] * c1 + e[5]
        hit = ok & (np.abs(ddn) > 1e-7) & (t > 0) & (t < ln) & (b0 >= 0) & (b1 >= 0) & (b0 + b1 <= 1)
        cov += np.where(hit, float(m["tri_colors"][ti, 0]), 0.0)
    # where the segment is not blocked by an opaque, non-pane triangle the result is 1 - min(1, coverage)
    o2 = pyoracle.OracleEnv()
    keep = np.ones(sky_scene.n_tris, bool); keep[m["n_opaque"]:m["n_world"]] = False
    o2.add_triangles(sky_scene.tri_ids[keep], sky_scene.tri_verts[keep]); o2.build()
    free = o2.test_lines_sky(a, b, flags=0, static_prop_to_skip=7, threads=8) == 1.0
    # panes only sit inside the room, below the sky ceiling, so every pane crossed lies before the nearest sky hit
    want = 1.0 - np.minimum(1.0, cov)
    assert np.allclose(got[free], want[free], atol=2e-6)
    assert np.all(got[~free] == 0.0)
    assert np.all(got >= flat) and (got > flat).sum() > 200          # panes stop being hard blockers
    assert len(np.unique(got)) > 6                                   # fractional visibilities appear
    both = sky_oracle.test_lines_sky(a, b, flags=3, static_prop_to_skip=7, threads=8)
    assert np.all(both <= got + 1e-6)                               # (occ + 1) - fv2 rounds; the sky boxes only add occlusion
    assert (both < got - 0.1).sum() > 50


def test_packet_leaf_flag(sky_scene, sky_oracle):
    """testline.go:63 takes the leaf from lane 0: a packet whose lane 0 starts in a sky box (area with a
    camera) does not recurse on any lane; with lane 0 in the room all four lanes recurse."""
    a1 = np.array([[10.0, 8192.0], [20.0, 0.0], [50.0, 5.0]], np.float32)      # room point, sky-box point
    up = np.array([[0.3], [0.1], [1.0]], np.float32) * 20000
    for lane0 in (0, 1):
        order = [lane0, 0, 0, 0]
        a = np.ascontiguousarray(a1[:, order]); b = np.ascontiguousarray(a + up)
        per_lane = sky_oracle.test_lines_sky(a, b, flags=1, static_prop_to_skip=-1)
        packet = sky_oracle.test_lines_sky(a, b, flags=1 | 4, static_prop_to_skip=-1)
        flat = sky_oracle.test_lines_sky(a, b, flags=0, static_prop_to_skip=-1)
        if lane0 == 0:
            assert np.array_equal(packet, per_lane)
        else:
            assert np.array_equal(packet, flat)          # nobody recursed


def test_can_leaf_trace_to_sky(sky_scene, sky_oracle):
    """lightmap.go:425-451 on the BSP leaves + probe boxes: the probe inside the closed bunker cannot see sky."""
    m = sky_scene.meta
    dirs = _anorms()
    can = sky_oracle.leafs_trace_to_sky(m["probe_mins"], m["probe_maxs"], dirs, threads=8)
    nl = m["bsp"].leaf_mins.shape[0]
    assert can[nl + 0] == 0                         # inside the bunker
    assert can[0] == 1 and can[1] == 1              # the sky-box leaves see their own sky faces
    assert can[3:7].any()


def _anorms():
    import os
    return np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "vrad_b200", "data", "anorms.txt"), dtype=np.float32)


# ---- PVS run-length code (rad/lightmap/vis.go:54-94) --------------------------------------------

def _compress_vis(row: bytes) -> bytes:
    """The encoder every BSP compiler uses (Quake CompressVis): zero runs become (0, count<=255)."""
    out = bytearray(); j = 0
    while j < len(row):
        out.append(row[j])
        if row[j]:
            j += 1
            continue
        rep = 1
        j += 1
        while j < len(row) and row[j] == 0 and rep < 255:
            rep += 1; j += 1
        out.append(rep)
    return bytes(out)


def test_decompress_vis_known_answers_and_round_trip():
    row, used = pyoracle.decompress_vis(bytes([0xFF, 0x00, 0x03, 0x81]), 40)
    assert bytes(row) == bytes([0xFF, 0, 0, 0, 0x81]) and used == 4
    row, used = pyoracle.decompress_vis(bytes([0x00, 0xFF]), 20)              # overrun is clamped (vis.go:83-86)
    assert bytes(row) == bytes(3) and used == 2
    assert pyoracle.decompress_vis(bytes([0x00, 0x00]), 8)[1] < 0            # "0 repeat" is fatal (:78-80)
    rng = scenes.SplitMix64(99)
    for nc in (1, 7, 8, 9, 250, 2048, 5000):
        nb = (nc + 7) // 8
        dense = (rng.u64(nb) & np.uint64(0xFF)).astype(np.uint8)
        dense[rng.uniform(nb) < 0.7] = 0                                      # long zero runs, incl. > 255 bytes
        if nb > 600:
            dense[100:500] = 0
        comp = _compress_vis(dense.tobytes())
        row, used = pyoracle.decompress_vis(comp + b"\xAA\xBB", nc)
        assert row.tobytes() == dense.tobytes() and used == len(comp)


def test_product_host_vis_helpers_match_oracle():
    """vrad_decompress_vis / vrad_pvs_from_vis_lump are host-only: they run without a GPU."""
    from vrad_b200 import environment as E
    rng = scenes.SplitMix64(7)
    nc = 300
    nb = (nc + 7) // 8
    pvs = (rng.uniform(nc * nc) < 0.15).reshape(nc, nc)
    pvs[np.arange(nc), np.arange(nc)] = True
    lump = bytearray(); ofs = np.zeros((nc, 2), np.int32)
    for c in range(nc):
        bits = np.packbits(pvs[c], bitorder="little")
        assert bits.shape[0] == nb
        ofs[c, 0] = len(lump); ofs[c, 1] = -1
        lump += _compress_vis(bits.tobytes())
    ofs[17, 0] = -1                                                           # a cluster without vis data sees nothing
    for c in (0, 5, 299):
        got, used = E.decompress_vis(bytes(lump[ofs[c, 0]:]), nc)
        want, used_o = pyoracle.decompress_vis(bytes(lump[ofs[c, 0]:]), nc)
        assert got.tobytes() == want.tobytes() and used == used_o
    mat = E.pvs_from_vis_lump(nc, ofs, bytes(lump))
    want = pvs.astype(np.uint8); want[17] = 0
    assert np.array_equal(mat, want)
    with pytest.raises(E.VradError):
        E.decompress_vis(bytes([0, 0]), 8)
    with pytest.raises(E.VradError):
        E.decompress_vis(bytes([0xFF]), 64)                                   # input ends early
