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
