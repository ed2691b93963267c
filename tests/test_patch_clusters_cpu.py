"""Patch.ClusterNumber is assigned AFTER subdivision, per patch (rad/patches/subdivide.go:92-116): ClusterFromPoint(patch.Origin), and
for an origin in solid space the first winding point that is not; a patch that stays at -1 is in no cluster's child list, so the
transfer build leaves it out as receiver and as emitter.  Round-1 ADVICE: both bakes gave every patch the cluster of the first leaf
that lists its face, and vrad_build_transfers rejected the whole job on a -1.  CPU tests against the oracle (the device kernels'
form of the same rule is exercised by the GPU bake tests)."""
import numpy as np

from oracle import pyoracle
from vrad_b200 import bake, scenes


def _two_cluster_bsp():
    """x < 0: cluster 0; 0 <= x < 512: cluster 1; x >= 512: solid (cluster -1)."""
    b = scenes._BspBuilder()
    root = b.node(b.plane((1, 0, 0), 0.0, 0))
    right = b.node(b.plane((1, 0, 0), 512.0, 0))
    b.set_children(root, right, b.leaf(0, 0))              # front of x = 0 -> the right half, back -> cluster 0
    b.set_children(right, b.leaf(-1, 0), b.leaf(1, 0))     # front of x = 512 -> solid, back -> cluster 1
    return b.finish(1)


def _prep(origins, wind):
    first, count, pts = [], [], []
    for w in wind:
        first.append(len(pts)); count.append(len(w)); pts += list(w)
    tree = dict(origin=np.asarray(origins, np.float32), wind_first=np.asarray(first, np.int32), wind_count=np.asarray(count, np.int32),
                wind_points=np.asarray(pts, np.float32).reshape(-1, 3))
    return dict(tree=tree, pvs=np.ones((2, 2), np.uint8), bsp=_two_cluster_bsp())


def test_children_of_one_face_get_their_own_clusters_and_solid_origins_fall_back_to_the_winding():
    env = pyoracle.OracleEnv()
    # one floor face spanning x in [-256, 256] subdivided in two: the children sit in different leaves
    left, right = (-128.0, 0.0, 1.0), (128.0, 0.0, 1.0)
    quad = lambda x0, x1: [(x0, -64, 0), (x1, -64, 0), (x1, 64, 0), (x0, 64, 0)]
    # a patch whose origin is in solid but whose winding reaches back into cluster 1, and one entirely in solid
    in_solid_touching = (520.0, 0.0, 1.0)
    all_solid = (700.0, 0.0, 1.0)
    prep = _prep([left, right, in_solid_touching, all_solid],
                 [quad(-256, 0), quad(0, 256), [(600, -64, 0), (540, -64, 0), (500, 64, 0), (600, 64, 0)], quad(650, 750)])
    cl = bake.patch_clusters(env, prep)
    assert cl.tolist() == [0, 1, 1, -1]
    assert cl.dtype == np.int32
    # no vis data: every patch in cluster 0, nothing asked of the environment
    prep["pvs"] = None
    assert bake.patch_clusters(object(), prep).tolist() == [0, 0, 0, 0]


def test_cluster_minus_one_patches_neither_gather_nor_emit(s2_small_scene):
    scene = s2_small_scene
    o = pyoracle.env_from_scene(scene)
    full = o.build_transfers(scene.pvs, threads=8)
    rp, col, w = o.transfers()
    cluster = scene.patch_cluster.copy()
    out = np.array([5, 1000, 4321, scene.n_patches - 1])
    cluster[out] = -1
    o2 = pyoracle.OracleEnv()
    o2.add_triangles(scene.tri_ids, scene.tri_verts, scene.tri_flags); o2.build()
    o2.patches_upload(scene.patch_origin, scene.patch_normal, scene.patch_plane_dist, scene.patch_area, scene.patch_refl, cluster, scene.patch_flags)
    nnz = o2.build_transfers(scene.pvs, threads=8)
    rp2, col2, w2 = o2.transfers()
    assert nnz < full
    assert all(rp2[i + 1] == rp2[i] for i in out)                  # their rows are empty
    assert not np.isin(col2, out).any()                            # and nobody gathers from them
    # every other row: the old row minus the removed emitters, same order (weights renormalise with the row sum: MakeScales)
    for i in (0, 6, 999, 2500):
        old = col[rp[i]:rp[i + 1]]
        assert np.array_equal(col2[rp2[i]:rp2[i + 1]], old[~np.isin(old, out)])
