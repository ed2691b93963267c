"""End-to-end pass over the widened path through the C-ABI, the way the Go driver would call it:
faces -> MakePatchForFace + SubdividePatches (host) -> ClusterFromPoint on the device -> visibility lump -> PVS matrix (host)
-> patch upload + hierarchy -> direct light on the leaf patches -> hierarchical transfers -> bounces with CollectLight.
Every stage is compared with the oracle fed the same inputs; the final patch light within 1e-4 relative (north_star)."""
import numpy as np
import pytest

from vrad_b200 import scenes

pytestmark = pytest.mark.gpu


def test_faces_to_bounced_light():
    from oracle import pyoracle
    from vrad_b200.environment import Environment, pvs_from_vis_lump, subdivide_patches
    nx, ny = 3, 2
    base = scenes.multi_room(nx=nx, ny=ny)
    faces, pts, _ = scenes.room_faces(nx, ny)
    bsp = scenes.room_grid_bsp(nx, ny)
    lump, ofs = scenes.compress_vis_rows(base.pvs)

    g = Environment(); g.add_triangles(base.tri_ids, base.tri_verts, base.tri_flags); g.setup_acceleration_structure()
    o = pyoracle.OracleEnv(); o.add_triangles(base.tri_ids, base.tri_verts, base.tri_flags); o.build()
    # patches (host code in the library vs the oracle's literal restatement)
    t = subdivide_patches(faces, pts, min_chop=4.0)
    to = pyoracle.subdivide_patches(faces, pts, min_chop=4.0)
    assert all(t[k].tobytes() == to[k].tobytes() for k in t)
    N = t["origin"].shape[0]
    leaf = t["child1"] == -1
    # clusters: ClusterFromPoint(origin) (subdivide.go:108); every origin lies on a face, many on a room boundary plane
    g.bsp_upload(bsp); o.bsp_set(bsp)
    cluster = g.cluster_from_point(t["origin"])
    assert np.array_equal(cluster, o.cluster_from_point(t["origin"])) and cluster.min() >= 0
    # PVS: visibility lump -> matrix (vis.go:9-94)
    pvs = pvs_from_vis_lump(base.n_clusters, ofs, lump)
    assert np.array_equal(pvs, base.pvs)
    # reflectivity per face, inherited by the children (CreateChildPatch copies the parent)
    refl = np.minimum(scenes.SplitMix64(3).uniform(3 * len(faces), 0.2, 0.7).reshape(-1, 3), np.float32(0.99))[t["face"]].astype(np.float32)
    for env in (g, o):
        env.patches_upload(t["origin"], t["normal"], t["plane_dist"], t["area"], refl, cluster)
        env.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
    # direct light on the leaf patches (sample point lifted off the face like the transfer rays)
    pos = (t["origin"][leaf] + t["normal"][leaf]).astype(np.float32)
    dl_g = g.direct_light(pos, t["normal"][leaf], base.lights)
    dl_o = o.direct_light(pos, t["normal"][leaf], base.lights, threads=8)
    assert np.array_equal(dl_g.view(np.uint32), dl_o.view(np.uint32)) and (dl_g.sum(axis=1) > 0).mean() > 0.5
    # transfers + bounces; interior patches start from the area-weighted mean of their leaves (what CollectLight keeps)
    from vrad_b200 import sharding
    emit0 = np.zeros((N, 3), np.float32); emit0[leaf] = dl_g
    sharding.apply_collect(emit0, *sharding.collect_rows(t["parent"], t["child1"], t["child2"], t["area"]))
    assert g.build_transfers(pvs) == o.build_transfers(pvs, threads=8)
    rg, cg, wg = g.transfers_download(); ro, co, wo = o.transfers()
    assert np.array_equal(rg, ro) and np.array_equal(cg, co) and np.array_equal(wg.view(np.uint32), wo.view(np.uint32))
    tg, ag, dg = g.bounce(emit0, 100, early_out=True)
    tor, ao, do = o.bounce(emit0, 100, early_out=True, threads=8)
    assert dg == do and 2 <= dg < 100
    assert np.abs(tg - tor).max() <= 1e-4 * np.abs(tor).max()
    final = dl_g + tg[leaf]                                   # direct + bounced light per leaf patch
    assert np.isfinite(final).all() and (tg[leaf].sum(axis=1) > 0).mean() > 0.9
    # bounced energy stays below the closed-room bound E * rho / (1 - rho) with rho <= 0.7
    assert (tg[leaf] * t["area"][leaf, None]).sum() < (dl_g * t["area"][leaf, None]).sum() * 0.7 / 0.3
    g.close()
