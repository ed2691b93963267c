"""CPU tests of the BSP side of the path (include/vrad_bsp.h): the .bsp container, lumps -> triangles / face patches /
extents / tree tables / sky vis / smoothing normals / luxels, and the lighting lump write-back.  The product (host code in
libvradcuda.so, called through the C-ABI) is compared bit for bit with the oracle's function-by-function restatement
(oracle/bspside.py) on a synthetic BSP v20 map, and with hand-derived answers."""
import os
import struct

import numpy as np
import pytest

from oracle import bspside as O
from vrad_b200 import bspfile as B
from vrad_b200.lib import VradError


@pytest.fixture(scope="module")
def smap():
    return B.synthetic_map(3, 2, boxes_per_room=5, sky_rooms=(1,), bump_rooms=(0,))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _cube_map(side=64.0, contents=B.CONTENTS_SOLID, sky_side=None):
    """One leaf, one node, one axis-aligned cube brush [0, side]^3 (planes stored in +/- pairs)."""
    planes, sides = [], []
    for axis in range(3):
        n = [0.0, 0.0, 0.0]; n[axis] = 1.0
        planes.append((n, side, axis)); planes.append(([-c for c in n], -side, axis))        # +axis at `side`
        planes.append((n, 0.0, axis)); planes.append(([-c for c in n], -0.0, axis))          # pair for the face at 0
    P = np.zeros(len(planes), B.DPLANE)
    for k, (n, d, t) in enumerate(planes):
        P[k] = (n, d, t)
    tex = np.zeros(2, B.TEXINFO); tex[1]["flags"] = B.SURF_SKY
    for axis in range(3):
        sides.append((4 * axis, 0, 0, 0))          # +axis side, outward normal +axis at dist side
        sides.append((4 * axis + 3, 0, 0, 0))      # -axis side, outward normal -axis at dist 0
    S = np.zeros(6, B.DBRUSHSIDE)
    for k, t in enumerate(sides):
        S[k] = t
    if sky_side is not None:
        S[sky_side]["texinfo"] = 1
    br = np.zeros(1, B.DBRUSH); br[0] = (0, 6, contents)
    leafs = np.zeros(2, B.DLEAF); leafs[1]["numleafbrushes"] = 1; leafs[0]["cluster"] = -1
    nodes = np.zeros(1, B.DNODE); nodes[0]["planenum"] = 0; nodes[0]["children"] = (-1 - 1, -1 - 0)
    models = np.zeros(1, B.DMODEL)
    return B.Lumps(planes=P, texinfo=tex, texdata=np.zeros(1, B.DTEXDATA), brushsides=S, brushes=br, leafs=leafs, nodes=nodes, models=models,
                   leafbrushes=np.zeros(1, "<u2"))


# ---- container ---------------------------------------------------------------------------------------------------------
def test_bspfile_round_trip(smap, tmp_path):
    L, meta = smap
    path = str(tmp_path / "m.bsp")
    B.write_bsp(path, L, meta)
    raw = open(path, "rb").read()
    assert raw[:4] == b"VBSP" and struct.unpack_from("<i", raw, 4)[0] == 20
    for i in range(64):                                           # directory: aligned, in bounds, leaf lump version 1
        ofs, ln, ver, _ = struct.unpack_from("<iii4s", raw, 8 + 16 * i)
        assert ofs % 4 == 0 and ofs + ln <= len(raw) and (ln == 0 or ofs >= 1036)
        assert ver == (1 if i in (B.LUMP["LEAFS"], B.LUMP["FACES"]) else 0)
    f = B.BspFile(path)
    L2 = f.lumps()
    for k in L.a:
        assert np.array_equal(L.a[k], L2.a[k]), k
    assert L2.visdata.tobytes() == L.visdata.tobytes() and L2.n_areas == 2 and L2.n_clusters == 6
    assert f.get(B.LUMP["ENTITIES"])[0].rstrip(b"\0").decode() == meta["entities"]
    # replace a lump, save, reopen: every other lump is untouched
    f.set(B.LUMP["LIGHTING"], bytes(range(200)), version=1)
    path2 = str(tmp_path / "m2.bsp")
    f.save(path2); f.close()
    g = B.BspFile(path2)
    assert g.get(B.LUMP["LIGHTING"]) == (bytes(range(200)), 1)
    for k in L.a:
        assert np.array_equal(L.a[k], g.lumps().a[k]), k
    g.close()


def test_bspfile_rejects_bad_files(smap, tmp_path):
    L, meta = smap
    good = str(tmp_path / "g.bsp")
    B.write_bsp(good, L, meta)
    raw = bytearray(open(good, "rb").read())

    def opens(data):
        p = str(tmp_path / "bad.bsp")
        open(p, "wb").write(bytes(data))
        return B.BspFile(p)
    with pytest.raises(VradError):
        B.BspFile(str(tmp_path / "missing.bsp"))
    with pytest.raises(VradError):
        opens(raw[:500])                                          # shorter than a header
    bad = bytearray(raw); bad[:4] = b"IBSP"
    with pytest.raises(VradError):
        opens(bad)
    bad = bytearray(raw); struct.pack_into("<i", bad, 4, 17)
    with pytest.raises(VradError):
        opens(bad)                                                # unsupported version
    bad = bytearray(raw); struct.pack_into("<i", bad, 8 + 16 * B.LUMP["FACES"] + 4, len(raw))
    with pytest.raises(VradError):
        opens(bad)                                                # lump runs past the end
    f = opens(raw)
    faces = L.faces.copy(); faces[3]["texinfo"] = 999
    f.set(B.LUMP["FACES"], faces)
    with pytest.raises(VradError):
        f.lumps()                                                 # index outside another lump
    f.set(B.LUMP["FACES"], L.faces.tobytes()[:-3])
    with pytest.raises(VradError):
        f.lumps()                                                 # not a multiple of the record size
    f.set(B.LUMP["FACES"], L.faces); f.set(B.LUMP["LEAFS"], L.leafs, version=0)
    with pytest.raises(VradError):
        f.lumps()                                                 # 56-byte (version 0) leafs are not read
    f.close()


# ---- lumps -> triangles ------------------------------------------------------------------------------------------------
def test_raytrace_triangles_match_oracle(smap):
    L, meta = smap
    e = meta["brush_entity"]
    ids, verts = B.raytrace_triangles(L, [e["model"]], [e["origin"]], [e["angles"]])
    oids, overts = O.raytrace_triangles(L, [(e["model"], e["origin"], e["angles"])])
    assert np.array_equal(ids, oids) and np.array_equal(_bits(verts), _bits(overts))
    # 69 world brushes + the entity brush, 12 triangles each, minus the sky side of room 1's roof slab; 1 sky face
    assert (ids == B.TRACE_ID_OPAQUE).sum() == 12 * (L.brushes.shape[0]) - 2 and (ids == B.TRACE_ID_SKY).sum() == 2
    assert np.all(ids[-2:] == B.TRACE_ID_SKY)                     # sky faces come last (main.go:296-339)
    # the caster comes first, rotated 30 degrees about z and moved to its origin: its 8 corners
    c = verts[:12].reshape(-1, 3)
    ang = np.deg2rad(30.0)
    rot = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    corners = np.array([[x, y, z] for x in (-32, 32) for y in (-8, 8) for z in (0, 96)]) @ rot.T + e["origin"]
    for p in c:
        assert np.abs(corners - p).sum(axis=1).min() < 1e-3
    # without casters the world triangles are the same, 12 fewer
    ids0, verts0 = B.raytrace_triangles(L)
    assert np.array_equal(ids0, ids[12:]) and np.array_equal(verts0, verts[12:])
    # model indices outside (0, n_models) are ignored as brushmodelForEntity does
    ids1, _ = B.raytrace_triangles(L, [0, 7], [e["origin"]] * 2, [e["angles"]] * 2)
    assert np.array_equal(ids1, ids0)


def test_cube_brush_known_answer():
    side = 64.0
    ids, verts = B.raytrace_triangles(_cube_map(side))
    assert ids.shape[0] == 12 and np.all(ids == B.TRACE_ID_OPAQUE)
    assert set(np.unique(verts).tolist()) == {0.0, side}         # chopped at the axial planes without round-off (winding.go:153-157)
    a, b, c = verts[:, 0].astype(np.float64), verts[:, 1].astype(np.float64), verts[:, 2].astype(np.float64)
    n = np.cross(b - a, c - a)
    assert np.isclose(0.5 * np.linalg.norm(n, axis=1).sum(), 6 * side * side)
    centre = np.full(3, side / 2)
    assert np.all(np.einsum("ij,ij->i", n, (a + b + c) / 3 - centre) != 0)
    out = np.sign(np.einsum("ij,ij->i", n, (a + b + c) / 3 - centre))
    assert np.all(out == out[0])                                  # one consistent winding for the whole brush
    # a sky-textured side is not added; a non-opaque brush adds nothing (main.go:238-240,255)
    ids, verts = B.raytrace_triangles(_cube_map(side, sky_side=2))
    assert ids.shape[0] == 10 and not np.any(np.all(verts[:, :, 1] == side, axis=1))
    assert B.raytrace_triangles(_cube_map(side, contents=B.CONTENTS_WINDOW))[0].shape[0] == 0
    o_ids, o_verts = O.raytrace_triangles(_cube_map(side, sky_side=2))
    assert np.array_equal(ids, o_ids) and np.array_equal(_bits(verts), _bits(o_verts))


def test_oblique_brush_matches_oracle():
    """A wedge: the cube cut by an oblique plane -- general (non-axial) chops, where the interpolation arithmetic matters."""
    L = _cube_map(64.0)
    n = np.float32([0.6, 0.0, 0.8]); d = np.float32(70.0)
    P = np.concatenate([L.planes, np.zeros(2, B.DPLANE)])
    P[-2] = (n, d, 3); P[-1] = (-n, -d, 3)
    S = np.concatenate([L.brushsides, np.zeros(1, B.DBRUSHSIDE)]); S[-1] = (len(P) - 2, 0, 0, 0)
    br = L.brushes.copy(); br[0]["numsides"] = 7
    L2 = L.replace(planes=P, brushsides=S, brushes=br)
    ids, verts = B.raytrace_triangles(L2)
    o_ids, o_verts = O.raytrace_triangles(L2)
    assert ids.shape[0] > 12 and np.array_equal(ids, o_ids) and np.array_equal(_bits(verts), _bits(o_verts))
    assert np.all(verts.reshape(-1, 3).astype(np.float64) @ n.astype(np.float64) <= float(d) + 1e-2)     # everything behind the cut
    # bevel sides do not cut the other sides (main.go:264-266): the six cube faces come out whole again
    S2 = S.copy(); S2[-1]["bevel"] = 1
    _, vb = B.raytrace_triangles(L.replace(planes=P, brushsides=S2, brushes=br))
    whole = np.all(np.isin(vb.reshape(len(vb), -1), np.float32([0.0, 64.0])), axis=1)
    assert whole.sum() == 12 and np.all(whole[:12])


# ---- lumps -> face patches -----------------------------------------------------------------------------------------------
def test_face_patches_match_oracle(smap):
    L, meta = smap
    origins = np.zeros((2, 3), np.float32); origins[1] = meta["brush_entity"]["origin"]
    g = B.face_patches(L, origins)
    o = O.face_patches(L, origins)
    nf = len(o["windings"])
    assert g["faces"].shape[0] == nf == L.faces.shape[0]
    assert np.array_equal(g["face_number"], np.asarray(o["face_number"], np.int32))
    opts = np.asarray([p for w in o["windings"] for p in w], np.float32)
    assert np.array_equal(_bits(g["points"]), _bits(opts))
    assert np.array_equal(g["faces"]["n_points"], [len(w) for w in o["windings"]])
    assert np.array_equal(_bits(g["faces"]["normal"]), _bits(np.asarray(o["normal"], np.float32)))
    assert np.array_equal(_bits(g["faces"]["plane_dist"]), _bits(np.asarray(o["plane_dist"], np.float32)))
    assert np.array_equal(_bits(g["faces"]["lux_scale"]), _bits(np.asarray(o["lux_scale"], np.float32)))
    assert np.array_equal(g["faces"]["sky"], o["sky"]) and np.array_equal(g["faces"]["no_subdivide"], o["no_subdivide"])
    assert np.array_equal(_bits(g["reflectivity"]), _bits(np.asarray(o["reflectivity"], np.float32)))
    assert np.array_equal(g["needs_bump"], o["needs_bump"]) and np.array_equal(_bits(g["scale"]), _bits(np.asarray(o["scale"], np.float32)))
    assert np.all(g["base_area"] == 512 * 512)
    # hand-derived: luxel scale 1/16, texture scale 1/4; one sky face (NOLIGHT without LIGHT: no subdivision); room 0's floor is bumped
    assert np.all(g["faces"]["lux_scale"] == np.float32(1 / 16)) and np.all(g["scale"] == np.float32(0.25))
    assert g["faces"]["sky"].sum() == 1 and g["faces"]["no_subdivide"].sum() == 1 and g["needs_bump"].sum() == 1 and g["needs_bump"][0] == 1
    # the brush entity's faces are moved by the entity origin, and so are their planes (face.go:122-144)
    e = meta["brush_entity"]
    sel = np.isin(g["face_number"], np.arange(e["first_face"], e["first_face"] + e["n_faces"]))
    top = g["faces"][sel][0]
    assert np.allclose(top["normal"], (0, 0, 1)) and top["plane_dist"] == np.float32(96.0 + e["origin"][2])
    pts = g["points"][top["first_point"]:top["first_point"] + top["n_points"]]
    assert np.allclose(pts.mean(axis=0), (e["origin"][0], e["origin"][1], 96.0))


def test_face_patches_feed_subdivision(smap):
    from vrad_b200.environment import subdivide_patches
    L, _ = smap
    g = B.face_patches(L, None)
    t = subdivide_patches(g["faces"], g["points"], min_chop=4.0)
    leaves = t["child1"] == -1
    roots = t["parent"] == -1
    assert roots.sum() == L.faces.shape[0]
    # area is conserved from the face windings down to the leaf patches
    area_faces = sum(float(O.winding_area([O.vec(p) for p in g["points"][f["first_point"]:f["first_point"] + f["n_points"]]])) for f in g["faces"])
    assert np.isclose(t["area"][leaves].sum(dtype=np.float64), area_faces, rtol=1e-5)
    assert np.isclose(t["area"][roots].sum(dtype=np.float64), area_faces, rtol=1e-6)


def test_colinear_points_are_removed():
    """A square face with a redundant vertex in the middle of one edge (point.go:20-31)."""
    b = B._MapBuilder()
    ti = 0
    b.face([(0, 0, 0), (32, 0, 0), (64, 0, 0), (64, 64, 0), (0, 64, 0)], (0, 0, 1), ti)
    faces = np.zeros(1, B.DFACE); f = b.faces[0]
    faces[0]["planenum"] = f["planenum"]; faces[0]["firstedge"] = f["firstedge"]; faces[0]["numedges"] = 5; faces[0]["dispinfo"] = -1
    planes = np.zeros(len(b.planes), B.DPLANE)
    for k, (n, d, t) in enumerate(b.planes):
        planes[k] = (n, d, t)
    models = np.zeros(1, B.DMODEL); models[0]["numfaces"] = 1
    tex = np.zeros(1, B.TEXINFO); tex[0]["lightmap_vecs"][0][0] = 1 / 16; tex[0]["lightmap_vecs"][1][1] = 1 / 16
    L = B.Lumps(planes=planes, vertexes3=np.asarray(b.verts, np.float32), edges=np.asarray(b.edges, "<u2").view(B.DEDGE).reshape(-1),
                surfedges=np.asarray(b.surfedges, "<i4"), faces=faces, texinfo=tex, texdata=np.zeros(1, B.DTEXDATA), models=models)
    g = B.face_patches(L)
    assert g["faces"][0]["n_points"] == 4
    assert np.array_equal(g["points"], np.float32([(0, 0, 0), (64, 0, 0), (64, 64, 0), (0, 64, 0)]))
    o = O.face_patches(L)
    assert np.array_equal(g["points"], np.asarray(o["windings"][0], np.float32))


# ---- extents, tree tables ------------------------------------------------------------------------------------------------
def test_face_extents(smap):
    L, _ = smap
    mins, size, over = B.face_extents(L)
    omins, osize = O.face_extents(L)
    assert np.array_equal(mins, omins) and np.array_equal(size, osize) and over == 0
    # room 0's floor spans x 0..504, y 0..504 at 1/16 luxel per unit: mins 0, size ceil(31.5) = 32
    assert tuple(mins[0]) == (0, 0) and tuple(size[0]) == (32, 32)
    # the sky face keeps what the file stores (start.go:104-106)
    sky = int(np.nonzero(L.texinfo["flags"][L.faces["texinfo"]] & B.SURF_SKY)[0][0])
    assert tuple(size[sky]) == (0, 0)
    # denser luxels overflow the 126-luxel limit and are counted (face.go:66-87)
    t = L.texinfo.copy(); t["lightmap_vecs"][:, :, :3] *= 8
    assert B.face_extents(L.replace(texinfo=t))[2] > 0
    # the luxel-density rescale (start.go:21-64) brings them back
    t2 = B.rescale_lightmap_vecs(t, 1.0 / 16)
    assert np.array_equal(_bits(t2["lightmap_vecs"]), _bits(O.rescale_lightmap_vecs(t, 1.0 / 16)["lightmap_vecs"]))
    assert np.allclose(np.linalg.norm(t2["lightmap_vecs"][:, :, :3], axis=2), 1 / 16)
    assert np.array_equal(B.rescale_lightmap_vecs(t, 1.0)["lightmap_vecs"], t["lightmap_vecs"])        # no-op at density >= 1
    m2, s2, over2 = B.face_extents(L.replace(texinfo=t2))
    assert over2 == 0 and np.array_equal(s2, size)


def test_parents_and_cluster_table(smap):
    L, _ = smap
    npar, lpar = B.make_parents(L)
    onp, olp = O.make_parents(L)
    assert np.array_equal(npar, onp) and np.array_equal(lpar, olp)
    assert npar[0] == -1 and np.all(npar[1:] >= 0) and lpar[0] == -1 and np.all(lpar[1:7] >= 0)          # leaf 0 (solid) hangs nowhere
    for leaf in range(1, 7):                                      # walking up from every room leaf ends at the root
        n = lpar[leaf]
        assert (-1 - leaf) in L.nodes[n]["children"]
        while npar[n] != -1:
            assert n in L.nodes[npar[n]]["children"]; n = npar[n]
        assert n == 0
    first, leafs = B.cluster_table(L, L.n_clusters)
    table = O.build_cluster_table(L, L.n_clusters)
    assert [list(leafs[first[c]:first[c + 1]]) for c in range(L.n_clusters)] == table == [[1 + c] for c in range(6)]
    # a tree whose node is reached twice is refused
    bad = L.nodes.copy(); bad[1]["children"] = (2, 2)
    with pytest.raises(VradError):
        B.make_parents(L.replace(nodes=bad))


# ---- sky vis ---------------------------------------------------------------------------------------------------------------
def test_vis_for_light_environment(smap):
    L, meta = smap
    flags, pvs = B.vis_for_light_environment(L)
    oflags, opvs = O.build_vis_for_light_environment(L)
    assert np.array_equal(flags, oflags) and pvs.tobytes() == opvs
    # room 1 = (i 0, j 1) holds the sky face; its PVS row is the sky lights' PVS: rooms 0 (same column) and 3 (same row) + itself
    want = meta["pvs"][1]
    assert np.array_equal(np.unpackbits(pvs, bitorder="little")[:6], want)
    sees = np.nonzero(want)[0]
    for k in range(6):
        assert flags[1 + k] == (B.LEAF_FLAGS_SKY if k in sees else 0), k
    assert flags[0] == 0                                          # the solid leaf is never marked (lightmap.go:327-329)
    # no sky anywhere: no flags, no PVS
    L0, _ = B.synthetic_map(2, 1, boxes_per_room=1, with_brush_entity=False)
    f0, p0 = B.vis_for_light_environment(L0)
    assert not f0.any() and p0 is None and O.build_vis_for_light_environment(L0)[1] is None
    # a 2D-sky face marks SKY2D instead
    t = L.texinfo.copy(); t["flags"][t["flags"] & B.SURF_SKY != 0] |= B.SURF_SKY2D
    f2, _ = B.vis_for_light_environment(L.replace(texinfo=t))
    assert np.array_equal(f2, O.build_vis_for_light_environment(L.replace(texinfo=t))[0])
    assert f2[2] == B.LEAF_FLAGS_SKY2D and all(f2[1 + k] == B.LEAF_FLAGS_SKY2D for k in sees)
    # without vis data every leaf sees every sky leaf (GetVisCache, vis.go:11-20)
    fa, pa = B.vis_for_light_environment(L.replace(visdata=b""))
    assert np.all(fa[1:7] == B.LEAF_FLAGS_SKY) and pa is None


def test_radial_leafs_need_an_environment():
    """LEAF_FLAGS_RADIAL leafs that see no sky leaf go through CanLeafTraceToSky on the device (lightmap.go:373-381): without an
    environment that is an error, not a silent CPU path; radial leafs that already see sky need no trace."""
    L, _ = B.synthetic_map(3, 2, boxes_per_room=1, sky_rooms=(1,), radial_rooms=(4,), with_brush_entity=False)
    with pytest.raises(VradError) as ei:
        B.vis_for_light_environment(L)
    assert ei.value.status == -4
    L2, _ = B.synthetic_map(3, 2, boxes_per_room=1, sky_rooms=(1,), radial_rooms=(0,), with_brush_entity=False)
    f, _ = B.vis_for_light_environment(L2)
    assert f[1] == (B.LEAF_FLAGS_SKY | B.LEAF_FLAGS_RADIAL)
    # the oracle takes the trace as a callback
    fo, _ = O.build_vis_for_light_environment(L, can_leaf_trace_to_sky=lambda leaf: leaf == 5)
    assert fo[5] == (B.LEAF_FLAGS_SKY | B.LEAF_FLAGS_RADIAL)


# ---- smoothing normals ------------------------------------------------------------------------------------------------------
def test_pair_edges_match_oracle(smap):
    L, _ = smap
    for thr in (0.7071067, -1.0):
        vn, first, nb = B.pair_edges(L, thr)
        on, onb = O.pair_edges(L, thr)
        assert np.array_equal(_bits(vn), _bits(np.asarray([v for fn in on for v in fn], np.float32)))
        assert [list(nb[first[i]:first[i + 1]]) for i in range(L.faces.shape[0])] == onb
        normals, idx = B.save_vertex_normals(vn)
        onormals, oidx = O.save_vertex_normals(on)
        assert np.array_equal(_bits(normals), _bits(onormals)) and np.array_equal(idx, oidx)
        assert np.allclose(normals[idx], vn, atol=4e-3)           # equal within the 1e-5 squared-distance merge
        assert np.allclose(np.linalg.norm(vn, axis=1), 1, atol=1e-6)


def test_pair_edges_known_answers(smap):
    L, meta = smap
    # at the default 45-degree crease nothing in the axis-aligned map smooths across an edge: vertex normals = face normals,
    # but coplanar faces sharing vertices (wall pieces around a door) are neighbours
    vn, first, nb = B.pair_edges(L, 0.7071067)
    fn = L.planes["normal"][L.faces["planenum"]]
    plain = np.repeat(L.faces["smoothing_groups"] == 0, L.faces["numedges"])
    assert np.array_equal(vn[plain], np.repeat(fn, L.faces["numedges"], axis=0)[plain])
    for i in range(L.faces.shape[0]):
        for o in nb[first[i]:first[i + 1]]:
            assert np.dot(fn[i], fn[o]) > 0.7 or (L.faces[i]["smoothing_groups"] & L.faces[o]["smoothing_groups"])
    # with the threshold at -1 every face at a box corner joins in: the top face of an occluder box gets (+-1, +-1, 1)/sqrt(3)
    vn2, _, _ = B.pair_edges(L, -1.0)
    box_top = next(i for i in range(L.faces.shape[0]) if L.texinfo[L.faces[i]["texinfo"]]["texdata"] == 2 and fn[i][2] == 1)
    ofs = int(L.faces["numedges"][:box_top].sum())
    assert np.allclose(np.abs(vn2[ofs:ofs + 4]), 1 / np.sqrt(3), atol=1e-6) and np.all(vn2[ofs:ofs + 4, 2] > 0)
    # smoothing groups: solid walls of one room carry group 1 and meet at right angles -> smoothed whatever the threshold;
    # a hard-edge bit in common stops it (lightmap.go:160-172)
    wall = next(i for i in range(L.faces.shape[0]) if L.faces[i]["smoothing_groups"] == 1)
    o0 = int(L.faces["numedges"][:wall].sum())
    assert not np.array_equal(vn[o0:o0 + 4], np.repeat(fn[wall][None], 4, axis=0))
    hard = L.faces.copy(); hard["smoothing_groups"][hard["smoothing_groups"] == 1] = 0xff000001
    vh, _, _ = B.pair_edges(L.replace(faces=hard), 0.7071067)
    assert np.array_equal(vh[o0:o0 + 4], np.repeat(fn[wall][None], 4, axis=0))


def test_phong_normals(smap):
    L, _ = smap
    vn, _, _ = B.pair_edges(L, -1.0)
    on, _ = O.pair_edges(L, -1.0)
    g = B.face_patches(L)
    centroids = np.zeros((L.faces.shape[0], 3), np.float32)
    for k, f in enumerate(g["faces"]):
        centroids[g["face_number"][k]] = g["points"][f["first_point"]:f["first_point"] + f["n_points"]].astype(np.float64).mean(axis=0)
    rng = np.random.default_rng(5)
    faces = rng.integers(0, L.faces.shape[0], 300).astype(np.int32)
    pts = np.zeros((300, 3), np.float32)
    for k, fi in enumerate(faces):
        f = g["faces"][fi]
        w = g["points"][f["first_point"]:f["first_point"] + f["n_points"]].astype(np.float64)
        bary = rng.dirichlet(np.ones(len(w)))
        pts[k] = bary @ w
    out = B.phong_normals(L, vn, centroids, faces, pts, -1.0)
    want = np.asarray([O.get_phong_normal(L, on, centroids, int(fi), p, -1.0) for fi, p in zip(faces, pts)], np.float32)
    assert np.array_equal(_bits(out), _bits(want))
    assert np.allclose(np.linalg.norm(out, axis=1), 1, atol=1e-5)
    # known answers: at a face vertex the phong normal is that vertex's smoothed normal; threshold 1 turns smoothing off
    fi = 0
    ofs = 0
    v0 = L.vertexes3[O.edge_vertex(L, L.faces[fi], 0)]
    got = B.phong_normals(L, vn, centroids, [fi], [v0], -1.0)[0]
    assert np.allclose(got, vn[ofs], atol=1e-6)
    flat = B.phong_normals(L, vn, centroids, faces, pts, 1.0)
    assert np.array_equal(flat, L.planes["normal"][L.faces["planenum"][faces]])
    with pytest.raises(VradError):
        B.phong_normals(L, vn, centroids, [9999], [v0])


# ---- luxels and the lighting lump ------------------------------------------------------------------------------------------
def test_face_luxels_and_lighting_layout(smap, tmp_path):
    L, meta = smap
    mins, size, _ = B.face_extents(L)
    faces, first, nbytes = B.layout_lighting(L, mins, size)
    flags = L.texinfo["flags"][L.faces["texinfo"]]
    lit = (flags & (B.SURF_SKY | B.SURF_NOLIGHT)) == 0
    per_face = (size[:, 0] + 1) * (size[:, 1] + 1) * np.where(flags & B.SURF_BUMPLIGHT, 4, 1) * lit
    assert np.array_equal(np.diff(first), per_face) and nbytes == 4 * per_face.sum() + 4 * lit.sum()
    assert np.all(faces["lightofs"][~lit] == -1) and np.all(faces["styles"][~lit] == 255)
    assert np.all(faces["styles"][lit] == (0, 255, 255, 255)) and np.array_equal(faces["lm_size"], size)
    ofs = faces["lightofs"][lit]
    assert ofs[0] == 4 and np.array_equal(np.diff(ofs), 4 * per_face[lit][:-1] + 4)                   # average colour before each block
    L3 = L.replace(faces=faces)
    pos, nrm, lface = B.face_luxels(L3, mins, size, first)
    assert pos.shape[0] == first[-1] and np.array_equal(np.repeat(np.arange(L.faces.shape[0]), per_face), lface)
    # flat blocks equal the oracle's restatement bit for bit
    opos, onrm, oface = O.face_luxels(L3, mins, size)
    flat_sel = np.ones(pos.shape[0], bool)
    bump_face = int(np.nonzero(flags & B.SURF_BUMPLIGHT)[0][0])
    n_flat = (size[bump_face, 0] + 1) * (size[bump_face, 1] + 1)
    flat_sel[first[bump_face] + n_flat:first[bump_face + 1]] = False
    assert np.array_equal(_bits(pos[flat_sel]), _bits(opos)) and np.array_equal(_bits(nrm[flat_sel]), _bits(onrm)) and np.array_equal(lface[flat_sel], oface)
    # every sample is one unit in front of its face and maps back to integer lightmap coordinates mins + (s, t)
    pl = L.planes[L.faces["planenum"][lface]]
    assert np.allclose(np.einsum("ij,ij->i", pos.astype(np.float64), pl["normal"].astype(np.float64)) - pl["dist"], 1.0, atol=1e-3)
    lv = L.texinfo["lightmap_vecs"][L.faces["texinfo"][lface]].astype(np.float64)
    st = np.einsum("ijk,ik->ij", lv[:, :, :3], pos.astype(np.float64) - pl["normal"]) + lv[:, :, 3]
    assert np.allclose(st, np.round(st), atol=1e-4)
    k = first[5]
    w5 = size[5, 0] + 1
    assert np.allclose(st[k], mins[5]) and np.allclose(st[k + 1], mins[5] + (1, 0)) and np.allclose(st[k + w5], mins[5] + (0, 1))       # t major, s minor
    # the bump-mapped floor: 4 blocks, same positions, block 0 = the face normal, blocks 1..3 = the bump basis (sum = sqrt(3) n)
    blocks_p = pos[first[bump_face]:first[bump_face + 1]].reshape(4, n_flat, 3)
    blocks_n = nrm[first[bump_face]:first[bump_face + 1]].reshape(4, n_flat, 3)
    assert np.array_equal(blocks_p[0], blocks_p[1]) and np.array_equal(blocks_p[0], blocks_p[3])
    assert np.allclose(blocks_n[1:].sum(axis=0), np.sqrt(3) * blocks_n[0], atol=1e-5)
    # pack: colours land at lightofs, the average colour right before, and the file carries the lump
    rng = np.random.default_rng(11)
    rgb = rng.uniform(0, 400, (pos.shape[0], 3)).astype(np.float32)
    colors = B.color_to_rgbexp32(rgb)
    lump = B.pack_lighting(L3, first, colors, nbytes)
    assert len(lump) == nbytes
    for fi in (0, 5, bump_face, int(np.nonzero(lit)[0][-1])):
        o = int(faces["lightofs"][fi]); n = int(per_face[fi])
        assert lump[o:o + 4 * n] == colors[first[fi]:first[fi] + n].tobytes()
        flat_n = n // 4 if fi == bump_face else n
        avg = B.color_from_rgbexp32(colors[first[fi]:first[fi] + flat_n]).mean(axis=0, dtype=np.float64)
        got = B.color_from_rgbexp32(np.frombuffer(lump[o - 4:o], B.RGBEXP32))[0]
        assert np.allclose(got, avg, rtol=2e-2)
    path = str(tmp_path / "lit.bsp")
    B.write_bsp(path, L3, meta, lighting=lump)
    f = B.BspFile(path)
    assert f.get(B.LUMP["LIGHTING"])[0] == lump and np.array_equal(f.lumps().faces, faces)
    f.close()
    with pytest.raises(VradError):
        B.pack_lighting(L3, first, colors, nbytes - 8)           # the last face would not fit


def test_rgbexp32():
    rng = np.random.default_rng(3)
    mags = np.float32(2.0) ** rng.integers(-40, 40, (4000, 1)).astype(np.float32)
    rgb = (rng.uniform(0, 1, (4000, 3)).astype(np.float32) * mags).astype(np.float32)
    rgb[::17, 1] = 0; rgb[::29] = 0
    got = B.color_to_rgbexp32(rgb)
    want = np.asarray([O.pack_rgbexp32(c) for c in rgb], dtype=np.int64)
    assert np.array_equal(np.stack([got["r"], got["g"], got["b"], got["exponent"]], axis=1).astype(np.int64), want)
    back = B.color_from_rgbexp32(got)
    mx = rgb.max(axis=1)
    nz = mx > 0
    assert np.all(got[nz].view(np.uint8).reshape(-1, 4)[:, :3].max(axis=1) >= 128)                     # the largest mantissa is normalised
    assert np.all(back <= rgb * (1 + 1e-6)) and np.all(rgb[nz] - back[nz] <= (mx[nz] / 128)[:, None])   # truncation, one part in 128 of the largest
    # hand-derived: (1,2,3) -> largest 3 doubles 6 times to 192: exponent -6, mantissas 64,128,192
    c = B.color_to_rgbexp32([[1, 2, 3], [300, 200, 100], [255.5, 1, 1], [0, 0, 0], [-5, 4, np.nan], [np.inf, 1, 1], [1e-37, 0, 0], [256, 0, 0]])
    rows = [tuple(int(v) for v in (x["r"], x["g"], x["b"], x["exponent"])) for x in c]
    assert rows[0] == (64, 128, 192, -6) and rows[1] == (150, 100, 50, 1) and rows[2] == (255, 1, 1, 0) and rows[3] == (0, 0, 0, 0)
    assert rows[4] == (0, 128, 0, -5)                             # negative and NaN components count as zero
    assert rows[5][0] == 255 and rows[5][3] > 100 and rows[6] == (0, 0, 0, 0) and rows[7] == (128, 0, 0, 1)
    assert np.array_equal(B.color_from_rgbexp32(c[:2]), np.float32([[1, 2, 3], [300, 200, 100]]))


# ---- the file-driven bake, host side ------------------------------------------------------------------------------------
def test_luxel_nearest_patch():
    rng = np.random.default_rng(9)
    npatch, n = 60, 500
    patch_face = rng.integers(0, 5, npatch).astype(np.int32)
    child1 = np.where(rng.integers(0, 3, npatch) == 0, 1, -1).astype(np.int32)      # a third are interior patches
    origin = rng.uniform(0, 100, (npatch, 3)).astype(np.float32)
    lface = rng.integers(-1, 7, n).astype(np.int32)                                 # faces 5, 6 have no patch; -1 is no face
    pos = rng.uniform(0, 100, (n, 3)).astype(np.float32)
    got = B.luxel_nearest_patch(lface, pos, patch_face, origin, child1)
    for l in range(n):
        cand = [i for i in range(npatch) if patch_face[i] == lface[l] and child1[i] == -1]
        if not cand:
            assert got[l] == -1
            continue
        d = [float(np.sum((origin[i] - pos[l]).astype(np.float32) ** 2, dtype=np.float32)) for i in cand]
        assert got[l] in cand and np.isclose(float(np.sum((origin[got[l]] - pos[l]) ** 2)), min(d), rtol=1e-6)
    assert np.all(B.luxel_nearest_patch(lface, pos, patch_face, origin)[lface == 2] >= 0)      # child1 None: every patch is a leaf


def test_bake_prepare_host_pipeline(smap):
    from vrad_b200 import bake
    L, meta = smap
    ents = bake.parse_entities(meta["entities"])
    assert [e["classname"] for e in ents][:2] == ["worldspawn", "func_brush"] and sum(e["classname"] == "light" for e in ents) == 6
    cm, co, ca = bake.shadow_casters(ents)
    assert list(cm) == [1] and np.allclose(co[0], meta["brush_entity"]["origin"]) and np.allclose(ca[0], meta["brush_entity"]["angles"])
    prep = bake.prepare(L, meta["entities"])
    t = prep["tree"]
    N = t["origin"].shape[0]
    assert prep["tri_ids"].shape[0] == 12 * L.brushes.shape[0] - 2 + 2
    assert prep["lights"].shape[0] == 6 + 2 and list(prep["lights"]["type"][-2:]) == [3, 5]          # 6 point lights, sun + sky ambient
    assert prep["refl"].shape == (N, 3) and prep["cluster"].shape == (N,) and prep["cluster"].min() >= 0 and prep["cluster"].max() == 5
    # a patch's cluster is its room: the room its origin lies in
    room = (np.clip(t["origin"][:, 0] // 512, 0, 2) * 2 + np.clip(t["origin"][:, 1] // 512, 0, 1)).astype(np.int32)
    world = prep["face_of_patch"] < L.models[0]["numfaces"]
    assert np.array_equal(prep["cluster"][world], room[world])
    assert np.array_equal(prep["pvs"], meta["pvs"])
    # every luxel found a leaf patch on its own face, within that patch's reach (chop 4 luxels = 64 units -> < 64 away)
    lp = prep["lux_patch"]
    assert lp.min() >= 0 and np.all(t["child1"][lp] == -1) and np.array_equal(prep["face_of_patch"][lp], prep["lux_face"])
    assert np.linalg.norm(t["origin"][lp] - prep["lux_pos"], axis=1).max() < 64.0
    assert prep["flags"].sum() > 0 and np.all(prep["flags"][prep["face_of_patch"] != np.nonzero(L.texinfo["flags"][L.faces["texinfo"]] & B.SURF_SKY)[0][0]] == 0)
    assert prep["lump_bytes"] == 4 * prep["lux_pos"].shape[0] + 4 * (np.diff(prep["luxel_first"]) > 0).sum()


def test_cpp_driver_reads_the_same_file(smap, tmp_path):
    """The compiled C++ host mirror (integration/cpp/drive --bsp) goes through the same C-ABI entry points from a .bsp file and
    arrives at the same counts as the Python mirror -- no GPU involved."""
    import subprocess
    from vrad_b200 import bake
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    L, meta = smap
    path = str(tmp_path / "m.bsp")
    B.write_bsp(path, L, meta)
    subprocess.run(["make", "-C", os.path.join(root, "integration", "cpp")], check=True, capture_output=True)
    out = subprocess.run([os.path.join(root, "integration", "cpp", "drive"), "--bsp", path], check=True, capture_output=True, text=True).stdout.split()
    got = {out[i]: int(out[i + 1]) for i in range(1, len(out), 2)}
    prep = bake.prepare(L, meta["entities"])
    t = prep["tree"]
    assert got == dict(faces=L.faces.shape[0], brushes=L.brushes.shape[0], triangles=prep["tri_ids"].shape[0] - 12, patches=t["origin"].shape[0],
                       leaves=int((t["child1"] == -1).sum()), luxels=prep["lux_pos"].shape[0], lighting_bytes=prep["lump_bytes"], oversize=0,
                       neighbours=int(B.pair_edges(L)[2].shape[0]))
    bad = subprocess.run([os.path.join(root, "integration", "cpp", "drive"), "--bsp", str(tmp_path / "nope.bsp")], capture_output=True, text=True)
    assert bad.returncode == 1 and "cannot open" in bad.stderr


# ---- lights.rad -------------------------------------------------------------------------------------------------------------
LIGHTS_RAD = "\r\n".join([
    "lights/white001\t255 255 255 200",
    "",                                                        # the literal reader stops here; the intent is to skip
    "lights/fluorescentwarm002a 255 240 200 150",
    "lights/single 128",
    "lights/rgbonly\t 10 20 30",
    "hdr:lights/dual 255 255 255 100 255 128 0 800",
    "ldr:lights/dual 255 255 255 100",
    "lights/both 255 255 255 100 255 128 0 800",
    "noshadow tools/toolsnodraw.vmt",
    "noshadow glass/window01",
    "forcetextureshadow models/props/fence_d.mdl",
    "lights/bad -5 10 10 10",
    "lights/garbage",
    "lights/white001 255 255 255 400",                         # overrides the first line
    "metal/box01 200 180 160 50",                              # a material of the synthetic map
]) + "\r\n"


def test_lights_rad_parse():
    for hdr in (False, True):
        table, noshadow, forced = B.texlights_parse(LIGHTS_RAD, hdr)
        otable, onoshadow, oforced = O.read_lights_rad(LIGHTS_RAD, hdr)
        assert [t["name"].decode() for t in table] == [n for n, _ in otable]
        assert np.array_equal(_bits(table["value"]), _bits(np.asarray([v for _, v in otable], np.float32)))
        assert noshadow == onoshadow == ["tools/toolsnodraw", "glass/window01"] and forced == oforced == ["props/fence_d"]
    table, _, _ = B.texlights_parse(LIGHTS_RAD, False)
    got = {t["name"].decode(): t["value"] for t in table}
    assert list(got) == ["lights/white001", "lights/fluorescentwarm002a", "lights/single", "lights/rgbonly", "lights/dual", "lights/both", "metal/box01"]
    assert np.allclose(got["lights/white001"], 400.0)                                   # white, overridden to brightness 400
    lin = lambda c: (c / 255.0) ** 2.2 * 255
    assert np.allclose(got["lights/fluorescentwarm002a"], [lin(255) * 150 / 255, lin(240) * 150 / 255, lin(200) * 150 / 255], rtol=1e-6)
    assert np.allclose(got["lights/single"], lin(128), rtol=1e-6) and np.allclose(got["lights/rgbonly"], [lin(10), lin(20), lin(30)], rtol=1e-6)
    assert np.allclose(got["lights/dual"], 100.0)                                       # the ldr: line
    assert np.allclose(got["lights/both"], [lin(255) * 800 / 255, lin(128) * 800 / 255, 0.0], rtol=1e-6)      # 8 numbers: the second tuple (reader.go:126)
    hdr_table, _, _ = B.texlights_parse(LIGHTS_RAD, True)
    hdr_got = {t["name"].decode(): t["value"] for t in hdr_table}
    assert np.allclose(hdr_got["lights/dual"], [lin(255) * 800 / 255, lin(128) * 800 / 255, 0.0], rtol=1e-6)
    # an LF-only file parses the same; too many entries are refused (MAX_TEXLIGHTS)
    t2, _, _ = B.texlights_parse(LIGHTS_RAD.replace("\r\n", "\n"), False)
    assert np.array_equal(t2, table)
    with pytest.raises(VradError):
        B.texlights_parse("\n".join(f"lights/l{i} 255 255 255 {i + 1}" for i in range(129)))
    assert B.texlights_parse("")[0].shape[0] == 0


def test_texlights_give_faces_their_base_light(smap):
    from vrad_b200.environment import lights_from_patches, subdivide_patches
    L, meta = smap
    table, _, _ = B.texlights_parse(LIGHTS_RAD)
    fp = B.face_patches(L)
    base, faces = B.apply_texlights(L, meta["string_table"], meta["string_data"], "synth", table, fp["face_number"], fp["faces"])
    otable, _, _ = O.read_lights_rad(LIGHTS_RAD)
    names = meta["string_data"].split(b"\0")
    want = np.asarray([O.light_for_texture(names[int(L.texdata[int(L.texinfo[int(L.faces[fn]["texinfo"])]["texdata"])]["name_id"])].decode(), "synth", otable)
                       for fn in fp["face_number"]], np.float32)
    assert np.array_equal(_bits(base), _bits(want))
    # the occluder boxes use material 2 = "metal/box01": exactly their faces emit
    is_box = L.texinfo["texdata"][L.faces["texinfo"][fp["face_number"]]] == 2
    assert is_box.sum() > 0 and np.array_equal(np.any(base != 0, axis=1), is_box) and np.array_equal(faces["has_base_light"] == 1, is_box)
    assert np.array_equal(faces["no_subdivide"], fp["faces"]["no_subdivide"])            # no NOLIGHT face emits here
    # a cubemap-patched name resolves to the original material
    data = meta["string_data"].replace(b"metal/box01", b"maps/synth/metal/box01_12_-40_7")
    shift = len(b"maps/synth/metal/box01_12_-40_7") - len(b"metal/box01")
    st = meta["string_table"].copy(); st[3:] += shift
    base2, _ = B.apply_texlights(L, st, data, "synth", table, fp["face_number"])
    assert np.array_equal(base2, base)
    assert not B.apply_texlights(L, st, data, "othermap", table, fp["face_number"])[0].any()
    # and the patches made from them become surface lights (CreateDirectLights, lights.go:49-82)
    t = subdivide_patches(faces, fp["points"], min_chop=4.0)
    leaf = t["child1"] == -1
    lights = lights_from_patches(t["origin"], t["normal"], base[t["face"]], t["area"], np.repeat(fp["scale"][t["face"]][:, :1], 2, axis=1) * 0 + 1.0,
                                 fp["base_area"][t["face"]], t["child1"])
    assert lights.shape[0] == int((leaf & is_box[t["face"]]).sum()) and np.all(lights["type"] == 0)


def test_direct_light_honours_the_light_pvs():
    """DirectLight.PVS (AllocDLight / SetDLightVis, rad/lightmap/lights.go:118-161; PVSCheck, lightmap.go:413-422): a light reaches only
    the samples whose cluster its own cluster sees.  Checked on the oracle's environment through the bake's host logic."""
    from oracle import pyoracle
    from vrad_b200 import bake
    L, meta = B.synthetic_map(4, 1, boxes_per_room=3, pvs_radius=1, with_brush_entity=False)
    # one light per room at door height, in line with the door openings, so that it does shine through two doors in a row
    ents = "".join('{\n"classname" "light"\n"origin" "%g 256 128"\n"_light" "255 255 255 300"\n}\n' % (512 * r + 200) for r in range(4))
    prep = bake.prepare(L, ents)
    assert prep["lights"].shape[0] == 4 and prep["sky_pvs"] is None
    culled = bake.light(pyoracle.OracleEnv(), prep, bounces=1)
    full = bake.light(pyoracle.OracleEnv(), prep, bounces=1, use_light_pvs=False)
    assert np.all(culled["direct"] <= full["direct"]) and np.abs(culled["direct"] - full["direct"]).max() > 1.0     # light does pass two doors in a row
    # by hand: the samples of room r are lit by the lights of rooms r-1, r, r+1 only
    o = pyoracle.OracleEnv()
    o.add_triangles(prep["tri_ids"], prep["tri_verts"].reshape(-1, 9)); o.build()
    o.bsp_set(prep["bsp"])
    lux_room = o.cluster_from_point(prep["lux_pos"])                  # ClusterFromPoint of the sample, as upstream's BuildFacelights records it
    agree = lux_room == meta["face_room"][prep["lux_face"]]           # = the face's room, except for luxels hanging over the face's edge
    assert agree.mean() > 0.9 and lux_room.min() >= 0
    want = np.zeros_like(culled["direct"])
    for r in range(4):
        sel = np.nonzero(lux_room == r)[0]
        keep = np.abs(np.arange(4) - r) <= 1                          # light k sits in room k
        want[sel] = o.direct_light(prep["lux_pos"][sel], prep["lux_normal"][sel], prep["lights"][keep])
    assert np.array_equal(culled["direct"], want)
    # a PVS that sees everything changes nothing, bit for bit
    L2, meta2 = B.synthetic_map(4, 1, boxes_per_room=3, pvs_radius=9, with_brush_entity=False)
    prep2 = bake.prepare(L2, ents)
    a = bake.light(pyoracle.OracleEnv(), prep2, bounces=1); b = bake.light(pyoracle.OracleEnv(), prep2, bounces=1, use_light_pvs=False)
    assert np.array_equal(a["direct"], b["direct"]) and np.array_equal(a["total"], b["total"])


def test_target_faces_ldr_and_hdr(smap, tmp_path):
    """cache.SetTargetFaces (cmd/tasks/loadbsp/main.go:79-89): an HDR job lights LUMP_FACES_HDR (seeded from LUMP_FACES when empty) and
    writes LUMP_LIGHTING_HDR; the LDR lumps stay as they were."""
    L, meta = smap
    path = str(tmp_path / "m.bsp")
    B.write_bsp(path, L, meta)
    f = B.BspFile(path)
    assert f.set_target_faces(False) == (B.LUMP["FACES"], B.LUMP["LIGHTING"])
    assert f.get(B.LUMP["FACES_HDR"])[0] == b""
    assert f.set_target_faces(True) == (B.LUMP["FACES_HDR"], B.LUMP["LIGHTING_HDR"])
    assert f.get(B.LUMP["FACES_HDR"])[0] == L.faces.tobytes()
    hdr_faces = L.faces.copy(); hdr_faces["lightofs"] = 1234
    f.set(B.LUMP["FACES_HDR"], hdr_faces)
    assert np.array_equal(f.lumps().faces, hdr_faces)                 # the views follow the target
    f.set_target_faces(False)
    assert np.array_equal(f.lumps().faces, L.faces)
    f.close()


def test_lumps_from_elsewhere_are_validated(smap):
    """vrad_bsp_validate: the checks of vrad_bspfile_lumps for lumps that did not come through the container (the Go loader's cache)."""
    L, _ = smap
    L.validate()
    for name, field, value in (("faces", "planenum", 60000), ("faces", "texinfo", -1), ("texinfo", "texdata", 99), ("brushsides", "planenum", 65535),
                               ("leafs", "numleaffaces", 60000), ("brushes", "numsides", 100000), ("models", "numfaces", 100000)):
        arr = L.a[name].copy(); arr[field][0] = value
        with pytest.raises(VradError):
            L.replace(**{name: arr}).validate()
    se = L.surfedges.copy(); se[5] = -(L.edges.shape[0] + 3)
    with pytest.raises(VradError):
        L.replace(surfedges=se).validate()
    nodes = L.nodes.copy(); nodes[2]["children"] = (1, 1)              # a cycle / shared child: GetBrushRecursive would not terminate
    with pytest.raises(VradError):
        L.replace(nodes=nodes).validate()
    with pytest.raises(VradError):
        L.replace(visdata=np.int32(50).tobytes() + b"\0" * 16).validate()


# ---- bounced light per luxel: the radial filter -------------------------------------------------------------------------------
def test_radial_filter_matches_the_scatter_form(smap):
    """The gather kernel's functor (host policy) against the oracle's restatement of upstream's scatter (BuildPatchRadial /
    AddBouncedToRadial / SampleRadial): same sums in the same patch order, so bit for bit."""
    from vrad_b200 import bake
    L, meta = smap
    prep = bake.prepare(L, meta["entities"])
    t = prep["tree"]
    N = t["origin"].shape[0]
    rng = np.random.default_rng(12)
    totals = rng.uniform(0, 200, (N, 3)).astype(np.float32)
    mins, size, _ = B.face_extents(prep["lumps"])
    vn, nb_first, nb = B.pair_edges(L, 0.7071067)
    first, entries = B.radial_entries(prep["lumps"], mins, t, prep["face_of_patch"], prep["face_origin"], nb_first, nb)
    assert np.array_equal(first, prep["radial_first"]) and np.array_equal(entries, prep["radial_entries"])          # what the bake uses
    got = B.luxel_radial_light(None, prep["lux_face"], prep["luxel_first"], size, first, entries, totals)
    patch_lists = [[] for _ in range(L.faces.shape[0])]
    for p in range(N):
        if t["child1"][p] == -1:
            patch_lists[prep["face_of_patch"][p]].append(p)
    nbs = [list(nb[nb_first[f]:nb_first[f + 1]]) for f in range(L.faces.shape[0])]
    lit_faces = np.nonzero(np.diff(prep["luxel_first"]) > 0)[0]
    bump_face = int(np.nonzero(L.texinfo["flags"][L.faces["texinfo"]] & B.SURF_BUMPLIGHT)[0][0])
    ent_face = meta["brush_entity"]["first_face"]                         # a face of the brush model: its patches live at the entity's origin
    for f in [int(lit_faces[0]), int(lit_faces[7]), int(lit_faces[-1]), bump_face, ent_face] + [int(x) for x in lit_faces[20:24]]:
        want = O.build_patch_radial(L, f, mins, size, patch_lists, t, totals, nbs, prep["face_origin"])
        a = int(prep["luxel_first"][f])
        assert np.array_equal(_bits(got[a:a + want.shape[0]]), _bits(want)), f
        if f == bump_face:                                                 # without bump totals the three extra blocks repeat the flat one
            for b in range(1, 4):
                assert np.array_equal(got[a + b * want.shape[0]:a + (b + 1) * want.shape[0]], got[a:a + want.shape[0]])
    # with bump totals block b takes Light[b]
    bump = rng.uniform(0, 200, (N, 3, 3)).astype(np.float32)
    gotb = B.luxel_radial_light(None, prep["lux_face"], prep["luxel_first"], size, first, entries, totals, bump)
    nflat = (size[bump_face, 0] + 1) * (size[bump_face, 1] + 1)
    a = int(prep["luxel_first"][bump_face])
    for b in range(1, 4):
        want = O.build_patch_radial(L, bump_face, mins, size, patch_lists, t, bump[:, b - 1], nbs, prep["face_origin"])
        assert np.array_equal(_bits(gotb[a + b * nflat:a + (b + 1) * nflat]), _bits(want))
    assert np.array_equal(gotb[a:a + nflat], got[a:a + nflat])


def test_radial_filter_properties(smap):
    from vrad_b200 import bake
    L, meta = smap
    prep = bake.prepare(L, meta["entities"])
    t = prep["tree"]
    N = t["origin"].shape[0]
    mins, size, _ = B.face_extents(prep["lumps"])
    first, entries = B.radial_entries(prep["lumps"], mins, t, prep["face_of_patch"], prep["face_origin"])
    # entries: only leaf patches, each on its own face (no neighbours given), centre inside the face's luxel rectangle
    assert np.all(t["child1"][entries["patch"]] == -1)
    per_face = np.repeat(np.arange(L.faces.shape[0]), np.diff(first))
    assert np.array_equal(prep["face_of_patch"][entries["patch"]], per_face)
    assert np.all(entries["s"] > -0.01) and np.all(entries["s"] < size[per_face, 0] + 1.01) and np.all(entries["inv_ds"] <= 1.0) and np.all(entries["inv_dt"] <= 1.0)
    # a constant patch light comes back as the same constant on every luxel a patch reaches (weights normalise), zero elsewhere
    const = np.tile(np.float32([7.0, 11.0, 13.0]), (N, 1))
    out = B.luxel_radial_light(None, prep["lux_face"], prep["luxel_first"], size, first, entries, const)
    reached = out.any(axis=1)
    assert reached.mean() > 0.99 and np.allclose(out[reached], [7.0, 11.0, 13.0], rtol=1e-5)
    # the filter interpolates: every luxel lies between the smallest and the largest patch light of its face (+ neighbours)
    rng = np.random.default_rng(3)
    totals = rng.uniform(10, 20, (N, 3)).astype(np.float32)
    out = B.luxel_radial_light(None, prep["lux_face"], prep["luxel_first"], size, first, entries, totals)
    assert out[reached].min() >= 10.0 - 1e-3 and out[reached].max() <= 20.0 + 1e-3
    # and it follows the light: luxels next to a bright patch are brighter than those next to a dark one
    f = int(np.nonzero(np.diff(prep["luxel_first"]) > 500)[0][0])
    on_face = np.nonzero((prep["face_of_patch"] == f) & (t["child1"] == -1))[0]
    ramp = np.zeros((N, 3), np.float32)
    ramp[on_face] = (t["origin"][on_face, :1] - t["origin"][on_face, :1].min()) + (t["origin"][on_face, 1:2] - t["origin"][on_face, 1:2].min())
    out = B.luxel_radial_light(None, prep["lux_face"], prep["luxel_first"], size, first, entries, ramp)
    a, b = int(prep["luxel_first"][f]), int(prep["luxel_first"][f + 1])
    near = B.luxel_nearest_patch(prep["lux_face"][a:b], prep["lux_pos"][a:b], prep["face_of_patch"], t["origin"], t["child1"])
    assert np.corrcoef(out[a:b, 0], ramp[near, 0])[0, 1] > 0.98
    with pytest.raises(VradError):
        bad = entries.copy(); bad["patch"][0] = N
        B.luxel_radial_light(None, prep["lux_face"], prep["luxel_first"], size, first, bad, totals)


def test_bake_gives_child_patches_and_luxels_their_phong_normals(smap):
    """CreateChildPatch -> lightmap.GetPhongNormal (rad/patches/subdivide.go:385): children carry the smoothed normal at their origin,
    root patches the plane normal (face.go:152); luxel samples get the same treatment."""
    from vrad_b200 import bake
    L, meta = smap
    prep = bake.prepare(L, meta["entities"])
    t = prep["tree"]
    fop = prep["face_of_patch"]
    plane_n = L.planes["normal"][L.faces["planenum"][fop]]
    roots = t["parent"] == -1
    assert np.array_equal(t["normal"][roots], plane_n[roots])
    assert np.allclose(np.linalg.norm(t["normal"], axis=1), 1.0, atol=1e-5)
    smooth_face = L.faces["smoothing_groups"][fop] != 0                 # the doorless walls carry smoothing group 1 and meet at right angles
    bent = np.any(t["normal"] != plane_n, axis=1)
    assert not bent[~smooth_face].any() and bent[smooth_face & ~roots].mean() > 0.4
    on, _ = O.pair_edges(L, 0.7071067)
    kids = np.nonzero(smooth_face & ~roots)[0][::37]
    want = np.asarray([O.get_phong_normal(L, on, prep["face_centroids"], int(fop[p]), t["origin"][p] - prep["face_origin"][fop[p]]) for p in kids], np.float32)
    assert np.array_equal(_bits(t["normal"][kids]), _bits(want))
    # near a corner the normal leans towards the adjoining wall: a positive component along that wall's normal
    wall = int(np.nonzero((L.faces["smoothing_groups"] != 0))[0][0])
    on_wall = np.nonzero((fop == wall) & (t["child1"] == -1))[0]
    lean = np.abs(t["normal"][on_wall] - plane_n[on_wall]).max(axis=1)
    dist_to_centre = np.linalg.norm(t["origin"][on_wall] - prep["face_centroids"][wall], axis=1)
    assert lean[np.argmin(dist_to_centre)] < 0.1 and lean.max() > 0.2    # flat at the face centre, bent towards a smoothed corner
    # luxels: same normals as GetPhongNormal at the on-surface point
    lf = prep["lux_face"]
    sm_lux = np.nonzero(L.faces["smoothing_groups"][lf] != 0)[0][::211]
    flat_n = L.planes["normal"][L.faces["planenum"][lf]]
    pts = prep["lux_pos"][sm_lux] - flat_n[sm_lux]
    want = np.asarray([O.get_phong_normal(L, on, prep["face_centroids"], int(lf[i]), p) for i, p in zip(sm_lux, pts)], np.float32)
    assert np.allclose(prep["lux_normal"][sm_lux], want, atol=2e-6)
    plain = L.faces["smoothing_groups"][lf] == 0
    bump_extra = (L.texinfo["flags"][L.faces["texinfo"][lf]] & B.SURF_BUMPLIGHT) != 0
    assert np.array_equal(prep["lux_normal"][plain & ~bump_extra], flat_n[plain & ~bump_extra])


def test_samples_are_placed_on_their_faces(smap):
    L, meta = smap
    mins, size, _ = B.face_extents(L)
    faces, first, _ = B.layout_lighting(L, mins, size)
    L3 = L.replace(faces=faces)
    pos, nrm, lf = B.face_luxels(L3, mins, size, first)
    placed, st = B.place_samples(L3, mins, size, first, pos)
    moved = np.linalg.norm(placed - pos, axis=1) > 0
    assert 0.1 < moved.mean() < 0.4                                       # the luxels along the faces' edges
    # still one unit in front of the face plane
    pl = L.planes[L.faces["planenum"][lf]]
    assert np.allclose(np.einsum("ij,ij->i", placed.astype(np.float64), pl["normal"].astype(np.float64)) - pl["dist"], 1.0, atol=1e-3)
    # the returned lightmap coordinates are those of the new position
    lv = L.texinfo["lightmap_vecs"][L.faces["texinfo"][lf]].astype(np.float64)
    st_world = np.einsum("ijk,ik->ij", lv[:, :, :3], placed.astype(np.float64) - pl["normal"]) + lv[:, :, 3] - mins[lf]
    assert np.allclose(st_world, st, atol=2e-3)
    # every sample lies on its face (inside the winding, up to rounding) and within half a luxel of its grid point unless the cell misses the face
    checked = 0
    for f in np.nonzero(np.diff(first) > 0)[0][::9]:
        fc = L.faces[f]
        verts = np.asarray([L.vertexes3[O.edge_vertex(L, fc, k)] for k in range(int(fc["numedges"]))], np.float64)
        tx = L.texinfo[int(fc["texinfo"])]["lightmap_vecs"].astype(np.float64)
        poly = verts @ tx[:, :3].T + tx[:, 3] - mins[f]
        w = size[f, 0] + 1
        for k in range(int(first[f]), int(first[f]) + (size[f, 0] + 1) * (size[f, 1] + 1), 7):
            s, t = (k - int(first[f])) % w, (k - int(first[f])) // w
            cs, ct, area = O.place_sample(poly, s, t)
            assert abs(st[k, 0] - cs) < 1e-3 and abs(st[k, 1] - ct) < 1e-3, (f, s, t)
            # inside the convex outline: all cross products of one sign (or zero)
            e = np.roll(poly, -1, axis=0) - poly
            cr = e[:, 0] * (st[k, 1] - poly[:, 1]) - e[:, 1] * (st[k, 0] - poly[:, 0])
            assert (cr >= -1e-3).all() or (cr <= 1e-3).all()
            if area > 0:
                assert abs(st[k, 0] - s) <= 0.5 + 1e-4 and abs(st[k, 1] - t) <= 0.5 + 1e-4
            if area > 1 - 1e-9:
                assert not moved[k]                                        # a cell wholly inside the face keeps its grid point, bit for bit
            checked += 1
    assert checked > 500
    # bump-mapped faces: the four blocks move alike
    bump_face = int(np.nonzero(L.texinfo["flags"][L.faces["texinfo"]] & B.SURF_BUMPLIGHT)[0][0])
    n_flat = (size[bump_face, 0] + 1) * (size[bump_face, 1] + 1)
    blocks = placed[first[bump_face]:first[bump_face + 1]].reshape(4, n_flat, 3)
    assert np.array_equal(blocks[0], blocks[1]) and np.array_equal(blocks[0], blocks[3])


def _word_checksum(a) -> int:
    """bake::Checksum of integration/cpp/vrad_bake.hpp: sum of word_i * (2654435761 * i + 1) over the 32-bit words, wrapping at 2^64."""
    b = np.ascontiguousarray(a).tobytes()
    b += b"\0" * (-len(b) % 4)
    w = np.frombuffer(b, "<u4").astype(np.uint64)
    with np.errstate(over="ignore"):
        return int(np.sum(w * (np.uint64(2654435761) * np.arange(w.shape[0], dtype=np.uint64) + np.uint64(1)), dtype=np.uint64))


def test_cpp_bake_prepares_what_the_python_mirror_prepares(smap, tmp_path):
    """integration/cpp/vrad_bake.hpp (the C++ stand-in for the reference's Go host side) against vrad_b200/bake.py: every array the
    device stages take -- triangles, patch tree with phong normals, clusters, PVS, lights, luxel samples, radial entries, the laid-out
    face lump -- byte for byte, from the same .bsp file.  No GPU."""
    import subprocess
    from vrad_b200 import bake
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    L, meta = smap
    path = str(tmp_path / "m.bsp")
    B.write_bsp(path, L, meta)
    subprocess.run(["make", "-C", os.path.join(root, "integration", "cpp")], check=True, capture_output=True)
    for with_rad, switches in ((False, {}), (True, {}), (False, dict(luxel_density=0.05, smooth_degrees=100.0, chop=2.0, max_chop=8.0))):
        _compare_cpp_prepare(root, path, L, meta, with_rad, tmp_path, switches)
    # and on non-axial geometry (wedges), everything smoothed
    L2, meta2 = B.synthetic_map(2, 2, boxes_per_room=3, ramps=True, sky_rooms=(2,), bump_rooms=(1,))
    path2 = str(tmp_path / "ramps.bsp")
    B.write_bsp(path2, L2, meta2)
    _compare_cpp_prepare(root, path2, L2, meta2, False, tmp_path, dict(luxel_density=1.0, smooth_degrees=170.0, chop=4.0, max_chop=4.0))


def _compare_cpp_prepare(root, path, L, meta, with_rad, tmp_path, switches):
    import math
    import subprocess
    from vrad_b200 import bake
    cmd = [os.path.join(root, "integration", "cpp", "drive"), "--prepare", path]
    kw = {}
    if with_rad:
        rad = str(tmp_path / "lights.rad")
        open(rad, "w", newline="").write(LIGHTS_RAD)
        cmd += ["-lights", rad]
        kw = dict(lights_rad=LIGHTS_RAD, texdata_strings=(meta["string_table"], meta["string_data"]), map_name="m")
    if switches:                                                         # the reference's command-line switches (cmd/args.go:53-93)
        cmd += ["-luxeldensity", str(switches["luxel_density"]), "-smooth", str(switches["smooth_degrees"]), "-chop", str(switches["chop"]),
                "-maxchop", str(switches["max_chop"])]
        kw.update(luxel_density=switches["luxel_density"], min_chop=switches["chop"], max_chop=switches["max_chop"],
                  smoothing_threshold=float(np.float32(math.cos(math.radians(switches["smooth_degrees"])))))
    out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
    got = {l.split()[0]: [int(x) for x in l.split()[1:]] for l in out.splitlines()}
    prep = bake.prepare(L, meta["entities"], **kw)
    assert (prep["lights"]["type"] == 0).any() == with_rad             # surface lights only with the texlight table
    t = prep["tree"]
    mine = dict(tri_ids=prep["tri_ids"], tri_verts=prep["tri_verts"], origin=t["origin"], normal=t["normal"], plane_dist=t["plane_dist"], area=t["area"],
                parent=t["parent"], child1=t["child1"], needs_bump=prep["needs_bump"], bump_basis=prep["bump_basis"], face_of_patch=prep["face_of_patch"], cluster=prep["cluster"], flags=prep["flags"], refl=prep["refl"],
                pvs=prep["pvs"], sky_pvs=prep["sky_pvs"], lights=prep["lights"], lm_mins=prep["lm_mins"], lm_size=prep["lm_size"], lit_faces=prep["lumps"].faces,
                luxel_first=prep["luxel_first"], lux_pos=prep["lux_pos"], lux_normal=prep["lux_normal"], lux_face=prep["lux_face"],
                radial_first=prep["radial_first"], radial_entries=prep["radial_entries"])
    assert sorted(got) == sorted(list(mine) + ["lump_bytes"])
    for k, v in mine.items():
        assert got[k][0] == _word_checksum(v), k
        assert got[k][1] == np.asarray(v).size if v.dtype.names is None else got[k][1] == v.shape[0], k
    assert got["lump_bytes"] == [prep["lump_bytes"]]


def test_cpp_bake_end_to_end_on_the_oracle(smap, tmp_path):
    """bake::BakeFile (integration/cpp/vrad_bake.hpp) run WITHOUT a GPU: the test binary links tests/helpers/oracle_device_shim.cpp, which puts
    the CPU oracle behind the device entry points the C++ bake calls.  Its lit .bsp must equal what the Python mirror produces with the
    oracle's environment and the host twins of the radial filter and K5 -- lump for lump.  (On a GPU the same comparison runs against the
    real kernels: tests/test_gpu_zzz_cpp_bake.py.)"""
    import subprocess
    from oracle import pyoracle
    from vrad_b200 import bake
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pyoracle.build()
    exe = str(tmp_path / "drive_on_oracle")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fopenmp", "-I" + os.path.join(root, "include"), "-I" + os.path.join(root, "integration", "cpp"), "-o", exe,
                    os.path.join(root, "integration", "cpp", "drive_main.cpp"), os.path.join(root, "tests", "helpers", "oracle_device_shim.cpp"),
                    "-L" + os.path.join(root, "vrad_b200", "_lib"), "-lvradcuda", "-L" + os.path.join(root, "oracle", "_build"), "-loracle",
                    "-Wl,-rpath," + os.path.join(root, "vrad_b200", "_lib"), "-Wl,-rpath," + os.path.join(root, "oracle", "_build")], check=True, capture_output=True)
    L, meta = B.synthetic_map(2, 2, boxes_per_room=4, sky_rooms=(1,), bump_rooms=(0,), ramps=True)
    src, dst = str(tmp_path / "in.bsp"), str(tmp_path / "cpp.bsp")
    B.write_bsp(src, L, meta)
    anorms = os.path.join(root, "vrad_b200", "data", "anorms.txt")
    r = subprocess.run([exe, "--bake", src, dst, anorms], check=True, capture_output=True, text=True)
    words = r.stdout.split()
    got = {words[i]: int(words[i + 1]) for i in range(1, len(words), 2)}
    # the Python mirror on the oracle's environment
    prep = bake.prepare(L, meta["entities"], texdata_strings=(meta["string_table"], meta["string_data"]), map_name="in")
    lit = bake.light(pyoracle.OracleEnv(), prep, bounces=8)
    assert got["transfers"] == lit["nnz"] and got["bounces"] == lit["bounces_done"]
    assert got["direct"] == _word_checksum(lit["direct"]) and got["emit"] == _word_checksum(lit["emit0"]) and got["total"] == _word_checksum(lit["total"])
    assert lit["bump_totals"].any() and got["bump"] == _word_checksum(lit["bump_totals"])
    ind = B.luxel_radial_light(None, prep["lux_face"], prep["luxel_first"], prep["lm_size"], prep["radial_first"], prep["radial_entries"], lit["total"], lit["bump_totals"])
    lump = B.pack_lighting(prep["lumps"], prep["luxel_first"], B.color_to_rgbexp32(lit["direct"] + ind), prep["lump_bytes"])
    f = B.BspFile(dst)
    assert f.get(B.LUMP["LIGHTING"]) == (lump, 1)
    assert np.array_equal(f.lumps().faces, prep["lumps"].faces) and f.get(B.LUMP["FACES"])[1] == 1
    normals, indices = B.save_vertex_normals(prep["vertex_normals"])                 # SaveVertexNormals' lumps travel with the file
    assert f.get(B.LUMP["VERTNORMALS"])[0] == normals.tobytes() and f.get(B.LUMP["VERTNORMALINDICES"])[0] == indices.tobytes()
    assert indices.shape[0] == int(L.faces["numedges"].sum()) and normals.shape[0] < indices.shape[0]
    for k in L.a:
        if k != "faces":
            assert np.array_equal(L.a[k], f.lumps().a[k]), k
    f.close()


def test_corrupted_files_never_crash_the_host_pipeline(tmp_path):
    """120 corrupted copies of a .bsp (byte flips in lump payloads, broken directory entries, extreme 32-bit values) through
    vrad_bspfile_open -> vrad_bspfile_lumps (vrad_bsp_validate) -> the whole host pipeline: each is either rejected with a message or
    processed; the process survives.  (The same loop ran clean under ASan/UBSan: profiles/r01d_host_sanitizers.txt.)"""
    from vrad_b200 import bake
    L, meta = B.synthetic_map(2, 1, boxes_per_room=2, sky_rooms=(1,), bump_rooms=(0,))
    good = str(tmp_path / "good.bsp")
    B.write_bsp(good, L, meta)
    raw = bytearray(open(good, "rb").read())
    rng = np.random.default_rng(7)
    survived = rejected = 0
    for it in range(120):
        b = bytearray(raw)
        if it % 3 == 0:
            for _ in range(rng.integers(1, 8)):
                b[int(rng.integers(1036, len(b)))] = int(rng.integers(0, 256))
        elif it % 3 == 1:
            struct.pack_into("<i", b, 8 + 16 * int(rng.integers(0, 64)) + 4 * int(rng.integers(0, 2)), int(rng.integers(-5, len(b) + 100)))
        else:
            struct.pack_into("<i", b, int(rng.integers(1036, len(b) - 4)), int(rng.choice([-1, 0x7fffffff, -0x80000000, 65535, 1 << 20])))
        bad = str(tmp_path / "bad.bsp")
        open(bad, "wb").write(bytes(b))
        try:
            f = B.BspFile(bad)
            try:
                bake.prepare(f.lumps(), f.get(B.LUMP["ENTITIES"])[0].rstrip(b"\0").decode("utf-8", "replace"))
                survived += 1
            finally:
                f.close()
        except (VradError, ValueError, IndexError, KeyError, OverflowError):
            rejected += 1
    assert survived + rejected == 120 and rejected > 10 and survived > 10


def test_non_axial_geometry_through_the_whole_host_pipeline():
    """The same product-vs-oracle comparisons on a map with wedges: a sloping face whose lightmap axes are WORLD axes (as a BSP compiler
    projects them), so the texture normal differs from the face normal (InitLightinfo's distscale path), non-axial brush planes in the
    winding chopper, and smoothing across a 26.6-degree crease."""
    from vrad_b200 import bake
    L, meta = B.synthetic_map(2, 1, boxes_per_room=2, ramps=True, with_brush_entity=False)
    L.validate()
    slopes = np.nonzero(L.planes["type"][L.faces["planenum"]] == 3)[0]
    assert len(slopes) == 2                                              # one sloping face per room
    # triangles: the wedge brush has 5 sides -> 2 + 2 + 2 + 1 + 1 triangles
    ids, verts = B.raytrace_triangles(L)
    oids, overts = O.raytrace_triangles(L)
    assert np.array_equal(ids, oids) and np.array_equal(_bits(verts), _bits(overts))
    assert ids.shape[0] == 12 * (L.brushes.shape[0] - 2) + 8 * 2
    # face patches, extents
    g, o = B.face_patches(L), O.face_patches(L)
    assert np.array_equal(_bits(g["points"]), _bits(np.asarray([p for w in o["windings"] for p in w], np.float32)))
    assert np.array_equal(_bits(g["faces"]["plane_dist"]), _bits(np.asarray(o["plane_dist"], np.float32)))
    mins, size, over = B.face_extents(L)
    omins, osize = O.face_extents(L)
    assert np.array_equal(mins, omins) and np.array_equal(size, osize) and over == 0
    assert tuple(mins[slopes[0]]) == (2, 2) and tuple(size[slopes[0]]) == (9, 5)     # x 40..168, y 40..104 projected on xy at 16 units per luxel
    # smoothing: at 45 degrees the slope (26.6 degrees off the floor normal ... it meets the wedge's own vertical faces at > 45) stays
    # flat; at 20 degrees of cosine threshold 0.85 nothing changes; with everything smoothed product == oracle
    for thr in (0.7071067, -1.0):
        vn, first, nb = B.pair_edges(L, thr)
        on, onb = O.pair_edges(L, thr)
        assert np.array_equal(_bits(vn), _bits(np.asarray([v for fn in on for v in fn], np.float32)))
        assert [list(nb[first[i]:first[i + 1]]) for i in range(L.faces.shape[0])] == onb
    # luxels on the slope: on the plane + 1, integer lightmap coordinates, bit-equal to the oracle's frame
    faces, first, _ = B.layout_lighting(L, mins, size)
    L3 = L.replace(faces=faces)
    pos, nrm, lf = B.face_luxels(L3, mins, size, first)
    opos, onrm, oface = O.face_luxels(L3, mins, size)
    assert np.array_equal(_bits(pos), _bits(opos)) and np.array_equal(_bits(nrm), _bits(onrm)) and np.array_equal(lf, oface)
    sel = lf == slopes[0]
    pl = L.planes[L.faces["planenum"][slopes[0]]]
    on_plane = pos[sel].astype(np.float64) - pl["normal"].astype(np.float64)
    assert np.allclose(on_plane @ pl["normal"].astype(np.float64), pl["dist"], atol=1e-3)
    lv = L.texinfo["lightmap_vecs"][L.faces["texinfo"][slopes[0]]].astype(np.float64)
    st = on_plane @ lv[:, :3].T + lv[:, 3]
    assert np.allclose(st, np.round(st), atol=1e-4)
    assert abs(float(np.dot(np.cross(lv[1, :3], lv[0, :3]) / np.linalg.norm(np.cross(lv[1, :3], lv[0, :3])), pl["normal"]))) < 0.95     # texture normal != face normal
    # sample placement keeps them on the slope
    placed, st2 = B.place_samples(L3, mins, size, first, pos)
    assert np.allclose((placed[sel].astype(np.float64) - pl["normal"]) @ pl["normal"].astype(np.float64), pl["dist"], atol=1e-3)
    # the whole prepare + the oracle's light run through; the radial gather equals the scatter form on the sloping face
    prep = bake.prepare(L, meta["entities"])
    t = prep["tree"]
    totals = np.random.default_rng(1).uniform(0, 100, (t["origin"].shape[0], 3)).astype(np.float32)
    got = B.luxel_radial_light(None, prep["lux_face"], prep["luxel_first"], prep["lm_size"], prep["radial_first"], prep["radial_entries"], totals)
    patch_lists = [[] for _ in range(L.faces.shape[0])]
    for p in range(t["origin"].shape[0]):
        if t["child1"][p] == -1:
            patch_lists[prep["face_of_patch"][p]].append(p)
    vn, nb_first, nb = B.pair_edges(L)
    nbs = [list(nb[nb_first[f]:nb_first[f + 1]]) for f in range(L.faces.shape[0])]
    f = int(slopes[1])
    want = O.build_patch_radial(L, f, prep["lm_mins"], prep["lm_size"], patch_lists, t, totals, nbs, prep["face_origin"])
    a = int(prep["luxel_first"][f])
    assert np.array_equal(_bits(got[a:a + want.shape[0]]), _bits(want))


def test_radial_leafs_are_traced_for_sky(tmp_path):
    """BuildVisForLightEnvironment's last branch (rad/lightmap/lightmap.go:373-381): a LEAF_FLAGS_RADIAL leaf that sees no sky leaf through
    its PVS is traced with CanLeafTraceToSky.  The library's host code calls the tracer through vrad_leafs_trace_to_sky; here the oracle
    stands behind that entry point (tests/helpers/oracle_device_shim.cpp) and the result is compared with the oracle's restatement given
    the oracle's own CanLeafTraceToSky as the callback."""
    import subprocess
    from oracle import pyoracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pyoracle.build()
    exe = str(tmp_path / "vis_radial")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fopenmp", "-I" + os.path.join(root, "include"), "-o", exe,
                    os.path.join(root, "tests", "helpers", "vis_radial_main.cpp"), os.path.join(root, "tests", "helpers", "oracle_device_shim.cpp"),
                    "-L" + os.path.join(root, "vrad_b200", "_lib"), "-lvradcuda", "-L" + os.path.join(root, "oracle", "_build"), "-loracle",
                    "-Wl,-rpath," + os.path.join(root, "vrad_b200", "_lib"), "-Wl,-rpath," + os.path.join(root, "oracle", "_build")], check=True, capture_output=True)
    # room 1 has the sky ceiling; with a PVS radius of 0 no other room sees it through vis; rooms 2 and 3 are radial.  Room 3 = (i 1, j 1) is
    # next to room 1 = (0, 1) with a tall (480-unit) door in between, so rays from its centre do reach the sky; room 2 = (1, 0) is around the corner.
    L, meta = B.synthetic_map(2, 2, boxes_per_room=0, sky_rooms=(1,), radial_rooms=(2, 3), pvs_radius=0, with_brush_entity=False, door_h=480.0)
    path = str(tmp_path / "radial.bsp")
    B.write_bsp(path, L, meta)
    anorms = os.path.join(root, "vrad_b200", "data", "anorms.txt")
    out = subprocess.run([exe, path, anorms], check=True, capture_output=True, text=True).stdout
    got = np.asarray([int(x) for x in out.split()], np.uint8)
    # the oracle: its restatement of the function with its own tracer as the callback
    ids, verts = O.raytrace_triangles(L)
    o = pyoracle.OracleEnv(); o.add_triangles(ids, verts.reshape(-1, 9)); o.build()
    dirs = np.loadtxt(anorms, dtype=np.float32)
    can = lambda leaf: bool(o.leafs_trace_to_sky(L.leafs["mins"][leaf:leaf + 1], L.leafs["maxs"][leaf:leaf + 1], dirs, threads=4)[0])
    want, _ = O.build_vis_for_light_environment(L, can_leaf_trace_to_sky=can)
    assert np.array_equal(got, want)
    assert got[1 + 1] == B.LEAF_FLAGS_SKY and got[1 + 0] == 0                                  # the sky room itself; room 0 is neither radial nor in its PVS
    assert got[1 + 3] == (B.LEAF_FLAGS_RADIAL | B.LEAF_FLAGS_SKY)                               # traced: sees the sky through the door
    assert got[1 + 2] == B.LEAF_FLAGS_RADIAL                                                     # traced: no direction finds the sky
