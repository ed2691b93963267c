"""CPU-only: patch creation + subdivision (rad/patches/face.go:29-197, subdivide.go:25-437).  The product's host
function (vrad_patches_subdivide, flat arrays + work stack) against the oracle's literal recursive restatement and
against hand-derived answers.  No device is needed: this is host code on both sides."""
import numpy as np
import pytest

from oracle import pyoracle
from vrad_b200 import scenes
from vrad_b200.environment import PATCH_TREE_FIELDS, VradError, subdivide_patches


def _one_face(w, h, lux=1.0 / 16.0, chop=4.0, **kw):
    fl = scenes._FaceList(lux, chop)
    fl.quad((0, 0, 0), (w, 0, 0), (0, h, 0), (0, 0, 1), **kw)
    return fl.arrays()


def test_square_face_known_answer():
    """1024 x 1024 face, 16 units per luxel, chop 4 luxels: halves down to 32-unit squares (64 units = 4 luxels still
    splits because the rule is total >= chop, subdivide.go:199)."""
    faces, pts = _one_face(1024, 1024)
    t = subdivide_patches(faces, pts)
    leaf = t["child1"] == -1
    assert leaf.sum() == 1024 and len(leaf) == 2047
    assert np.all(t["area"][leaf] == 1024.0)
    ext = t["maxs"][leaf] - t["mins"][leaf]
    assert np.all(ext[:, :2] == 32.0) and np.all(ext[:, 2] == 0.0)
    # origins = cell centres of the 32-unit grid, each exactly once
    cells = set((int(o[0] // 32), int(o[1] // 32)) for o in t["origin"][leaf])
    assert len(cells) == 1024 and np.allclose(t["origin"][leaf][:, :2] % 32, 16.0, atol=1e-3)   # balance point: fp32 area weights
    # root: WindingCenter / WindingArea
    assert tuple(t["origin"][0]) == (512.0, 512.0, 0.0) and t["area"][0] == 1024.0 * 1024.0 and t["parent"][0] == -1
    # the widest axis splits first; ties go to the lower axis (x): child1 = front = larger x (ClipWindingEpsilon front side)
    c1, c2 = t["child1"][0], t["child2"][0]
    assert (c1, c2) == (1, 2) and t["mins"][c1][0] == 512.0 and t["maxs"][c2][0] == 512.0
    # depth-first numbering: child1's subtree is numbered before child2's children
    assert t["child1"][c1] == 3 and t["child1"][c2] == 3 + (2047 - 3) // 2
    # parents come before children; areas add up
    kids = np.nonzero(~leaf)[0]
    assert np.all(t["child1"][kids] > kids) and np.all(t["child2"][kids] == t["child1"][kids] + 1)
    assert np.allclose(t["area"][kids], t["area"][t["child1"][kids]] + t["area"][t["child2"][kids]], rtol=1e-6)
    assert np.all(t["parent"][t["child1"][kids]] == kids) and np.all(t["parent"][t["child2"][kids]] == kids)


def test_rules():
    # sky faces and PreventSubdivision faces stay whole (subdivide.go:183-186, 151-165)
    for kw in ({"sky": 1}, {"no_subdivide": 1}):
        t = subdivide_patches(*_one_face(1024, 512, **kw))
        assert len(t["area"]) == 1
    # a face smaller than the chop is not split; a 3-luxel x 1-luxel strip is not "made more square" at chop == minChop
    assert len(subdivide_patches(*_one_face(48, 48))["area"]) == 1
    assert len(subdivide_patches(*_one_face(48, 16))["area"]) == 1
    # chop 8 > minChop 4: a 112 x 16 strip (7 x 1 luxels) is split once by the "make more square" rule (:204-212)
    # which also halves the chop to 4; its 3.5-luxel children then stay whole
    t = subdivide_patches(*_one_face(112, 16, chop=8.0), min_chop=4.0)
    assert len(t["area"]) == 3 and list(t["chop"]) == [4.0, 4.0, 4.0]
    # degenerate face (zero area) gives no patch at all (face.go:47-51)
    fl = scenes._FaceList(); fl.quad((0, 0, 0), (10, 0, 0), (20, 0, 0), (0, 0, 1)); fl.quad((0, 0, 0), (64, 0, 0), (0, 64, 0), (0, 0, 1))
    t = subdivide_patches(*fl.arrays())
    assert t["face"][0] == 1 and np.all(t["face"] == 1)
    with pytest.raises(VradError):
        f, p = _one_face(64, 64); f["n_points"] = 2
        subdivide_patches(f, p)


def test_triangle_and_slanted_faces_match_oracle_and_conserve_area():
    fl = scenes._FaceList()
    fl.points += [(0.0, 0.0, 0.0), (700.0, 0.0, 100.0), (100.0, 500.0, 300.0)]                     # a slanted triangle
    n = np.cross(np.array([700, 0, 100.0]), np.array([100, 500, 300.0])); n /= np.linalg.norm(n)
    fl.faces.append((0, 3, tuple(np.float32(n)), np.float32(0.0), np.float32(1 / 16), np.float32(4.0), 0, 0, 0, 0))
    fl.points += [(0.0, 0.0, 0.0), (300.0, -40.0, 0.0), (420.0, 200.0, 0.0), (250.0, 420.0, 0.0), (-60.0, 260.0, 0.0)]   # a pentagon
    fl.faces.append((3, 5, (0.0, 0.0, 1.0), np.float32(0.0), np.float32(1 / 8), np.float32(4.0), 0, 0, 1, 0))
    faces, pts = fl.arrays()
    t = subdivide_patches(faces, pts)
    o = pyoracle.subdivide_patches(faces, pts)
    for k in PATCH_TREE_FIELDS:
        assert t[k].tobytes() == o[k].tobytes(), k
    leaf = t["child1"] == -1
    for f in (0, 1):
        assert np.isclose(t["area"][leaf & (t["face"] == f)].sum(), t["area"][f], rtol=1e-5)
    assert leaf.sum() > 100
    # every child winding lies inside its parent's bounds
    kids = np.nonzero(t["parent"] >= 0)[0]
    assert np.all(t["mins"][kids] >= t["mins"][t["parent"][kids]] - 1e-3) and np.all(t["maxs"][kids] <= t["maxs"][t["parent"][kids]] + 1e-3)


def test_multi_room_faces_match_oracle():
    faces, pts, face_room = scenes.room_faces(3, 2)
    t = subdivide_patches(faces, pts)
    o = pyoracle.subdivide_patches(faces, pts)
    for k in PATCH_TREE_FIELDS:
        assert t[k].tobytes() == o[k].tobytes(), k
    leaf = t["child1"] == -1
    # wall area with door openings: total leaf area == total face area
    assert np.isclose(t["area"][leaf].sum(), t["area"][t["parent"] == -1].sum(), rtol=1e-6)
    sc = scenes.multi_room_hier(nx=3, ny=2)
    assert sc.n_patches == len(leaf) and np.array_equal(sc.meta["tree"]["parent"], t["parent"])
    assert sc.patch_cluster.max() == 5 and np.all(sc.patch_refl[t["child1"][0]] == sc.patch_refl[0])   # children inherit the face's reflectivity


def test_hierarchical_solution_tracks_the_flat_one():
    """The hierarchy is an approximation of the leaf-to-leaf matrix (far emitters are merged into their parents): several
    times fewer transfers, the same bounced light within a few per cent on average (oracle, 2x1-room map)."""
    sc = scenes.multi_room_hier(nx=2, ny=1, boxes_per_room=6)
    t = sc.meta["tree"]
    leaf = t["child1"] == -1
    oh = pyoracle.env_from_scene(sc)
    oh.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
    nnz_h = oh.build_transfers(sc.pvs, threads=4)
    idx = np.nonzero(leaf)[0]
    of = pyoracle.OracleEnv(); of.add_triangles(sc.tri_ids, sc.tri_verts, sc.tri_flags); of.build()
    of.patches_upload(sc.patch_origin[idx], sc.patch_normal[idx], sc.patch_plane_dist[idx], sc.patch_area[idx], sc.patch_refl[idx],
                      sc.patch_cluster[idx], sc.patch_flags[idx])
    nnz_f = of.build_transfers(sc.pvs, threads=4)
    assert nnz_f > 2.5 * nnz_h
    emit = np.full((sc.n_patches, 3), 100.0, np.float32)
    th, _, _ = oh.bounce(emit, 8, threads=4)
    tf, _, _ = of.bounce(emit[idx], 8, threads=4)
    assert abs(th[leaf].mean() - tf.mean()) < 0.02 * tf.mean()
    assert np.median(np.abs(th[leaf] - tf) / np.maximum(tf, 1e-3)) < 0.05
    # energy bound of a closed scene with reflectivity <= 0.7: bounced light below E * rho / (1 - rho)
    assert th[leaf].max() < 100.0 * 0.7 / 0.3


def test_hierarchical_candidates_against_a_python_walk():
    """The emitters of a receiver row, re-derived with an independent Python restatement of upstream's TestPatchToPatch walk
    (descend while |origin_i - origin_j|^2 / 16 < area_j; roots of visible clusters; not the receiver's own face), the
    MakeTransfer form factor in float64, and the oracle's own TestLine for the shadow segment."""
    sc = scenes.multi_room_hier(nx=2, ny=1, boxes_per_room=6)
    t = sc.meta["tree"]
    o = pyoracle.env_from_scene(sc)
    o.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
    o.build_transfers(sc.pvs, threads=4)
    rowptr, col, w = o.transfers()
    O, Nn, A, D = sc.patch_origin.astype(np.float64), sc.patch_normal.astype(np.float64), sc.patch_area.astype(np.float64), sc.patch_plane_dist.astype(np.float64)
    roots = np.nonzero(t["parent"] == -1)[0]

    def walk(i, j, out):
        if t["child1"][j] != -1:
            d = sc.patch_origin[i] - sc.patch_origin[j]                       # fp32 like the oracle: the test sits on a threshold
            d2 = np.float32(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
            if np.float32(d2 * np.float32(0.0625)) < sc.patch_area[j]:
                walk(i, t["child1"][j], out); walk(i, t["child2"][j], out)
                return
        out.append(j)

    leaves = np.nonzero(t["child1"] == -1)[0]
    checked = 0
    for i in leaves[:: max(1, len(leaves) // 25)]:
        cand = []
        for r in roots:
            if sc.pvs[sc.patch_cluster[i], sc.patch_cluster[r]] and t["face"][r] != t["face"][i]:
                walk(i, r, cand)
        cand = sorted(c for c in cand if c != i)
        # form factor (float64) and the plane test; pairs within 1e-6 of a threshold are left to the oracle
        keep, fuzzy = [], set()
        for j in cand:
            side = O[j] @ Nn[i] - D[i] - 0.01
            dl = O[i] - O[j]; ln = np.linalg.norm(dl)
            if ln == 0:
                continue
            dl /= ln
            scale = -(dl @ Nn[i]) * (dl @ Nn[j]) / (np.pi * ln * ln)
            trans = A[j] * scale
            if abs(side) < 1e-4 or abs(trans - 1e-7) < 1e-9 or abs(scale) < 1e-12:
                fuzzy.add(j)
            if side > 0 and scale > 0 and trans > 1e-7:
                keep.append(j)
        row = col[rowptr[i]:rowptr[i + 1]]
        assert set(row) - fuzzy <= set(keep), i                                # nothing outside the walk's emitter set
        missing = [j for j in keep if j not in set(row) and j not in fuzzy]
        if missing:                                                             # every missing emitter must be shadowed
            a = np.stack([(sc.patch_origin[min(i, j)] + sc.patch_normal[min(i, j)]) for j in missing]).T.astype(np.float32)
            b = np.stack([(sc.patch_origin[max(i, j)] + sc.patch_normal[max(i, j)]) for j in missing]).T.astype(np.float32)
            bits = o.test_lines(np.ascontiguousarray(a), np.ascontiguousarray(b), mode=2)      # brute-force tracer
            vis = np.unpackbits(bits.view(np.uint8), bitorder="little")[: len(missing)]
            assert not vis.any(), (i, missing)
        checked += 1
    assert checked >= 20


def test_collect_light_hand_example():
    """One bounce on a hand-made hierarchy: receiver plate R (a root with two leaf children of areas 3 and 1, facing +z)
    under an emitter E (single leaf, facing -z).  Leaves gather area_E * cos*cos / (pi d^2); the root holds the
    area-weighted mean of its children (vrad.cpp CollectLight, App. B.4)."""
    o = pyoracle.OracleEnv()
    o.add_triangles([scenes.TRACE_ID_OPAQUE], np.array([[9000, 9000, 9000, 9001, 9000, 9000, 9000, 9001, 9000]], np.float32)); o.build()
    origin = np.array([[0, 0, 0], [-1, 0, 0], [3, 0, 0], [0, 0, 6]], np.float32)         # root, child1, child2, emitter (36/16 < area 4: the walk descends)
    normal = np.array([[0, 0, 1], [0, 0, 1], [0, 0, 1], [0, 0, -1]], np.float32)
    pdist = np.array([0, 0, 0, -6], np.float32)
    area = np.array([4, 3, 1, 2], np.float32)
    refl = np.full((4, 3), 0.5, np.float32)
    o.patches_upload(origin, normal, pdist, area, refl, np.zeros(4, np.int32))
    o.set_hierarchy([-1, 0, 0, -1], [1, -1, -1, -1], [2, -1, -1, -1], [0, 0, 0, 1])
    assert o.build_transfers(None) == 4                                                  # E <- both children, each child <- E
    rowptr, col, w = o.transfers()
    assert list(np.diff(rowptr)) == [0, 1, 1, 2] and list(col) == [3, 3, 1, 2]          # the root gathers nothing; E sees the two LEAVES (near root)

    def ff(i, j):
        d = origin[i].astype(np.float64) - origin[j]; ln = np.linalg.norm(d); d /= ln
        return area[j] * (-(d @ normal[i]) * (d @ normal[j])) / (np.pi * ln * ln)
    assert np.allclose(w, [ff(1, 3), ff(2, 3), ff(3, 1), ff(3, 2)], rtol=1e-6)
    emit = np.zeros((4, 3), np.float32); emit[3] = 100.0
    tot, added, done = o.bounce(emit, 1)
    t1, t2 = ff(1, 3) * 100 * 0.5, ff(2, 3) * 100 * 0.5
    assert np.allclose(tot[1], t1, rtol=1e-6) and np.allclose(tot[2], t2, rtol=1e-6) and np.allclose(tot[3], 0)
    assert np.allclose(tot[0], 0.75 * t1 + 0.25 * t2, rtol=1e-6)                         # CollectLight: area-weighted mean of the children
    assert np.allclose(added, t1 + t2, rtol=1e-6)                                        # only leaves count towards `added`
