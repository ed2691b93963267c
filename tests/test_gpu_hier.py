"""GPU parity for the patch hierarchy: hierarchical transfer build (upstream TestPatchToPatch
walk) and CollectLight with parent/child patches, through the C-ABI, against the CPU oracle.  Transfer rows must be
bit-exact (columns and weights); bounced light within 1e-4 relative (fp32, north_star)."""
import numpy as np
import pytest

from vrad_b200 import scenes

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.fixture(scope="module")
def hier_scene():
    return scenes.multi_room_hier(nx=3, ny=2)


@pytest.fixture(scope="module")
def hier_pair(hier_scene):
    from oracle import pyoracle
    from vrad_b200.environment import environment_from_scene
    t = hier_scene.meta["tree"]
    g = environment_from_scene(hier_scene)
    g.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
    o = pyoracle.env_from_scene(hier_scene)
    o.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
    nnz_g = g.build_transfers(hier_scene.pvs)
    nnz_o = o.build_transfers(hier_scene.pvs, threads=8)
    yield g, o, nnz_g, nnz_o
    g.close()


def test_hierarchical_transfers_bit_exact(hier_scene, hier_pair):
    g, o, nnz_g, nnz_o = hier_pair
    assert nnz_g == nnz_o and nnz_g > 500_000
    rg, cg, wg = g.transfers_download()
    ro, co, wo = o.transfers()
    assert np.array_equal(rg, ro) and np.array_equal(cg, co)
    assert np.array_equal(wg.view(np.uint32), wo.view(np.uint32))
    t = hier_scene.meta["tree"]
    leaf = t["child1"] == -1
    rl = np.diff(rg)
    assert np.all(rl[~leaf] == 0)                                   # only leaf patches gather
    assert (~leaf[cg]).mean() > 0.3                                 # far emitters are interior (bigger) patches
    rows = np.repeat(np.arange(len(rl)), rl)
    assert np.all(t["face"][rows] != t["face"][cg])                 # never from the receiver's own face
    assert np.all(np.diff(cg)[np.diff(rows) == 0] > 0)              # ascending patch order inside a row
    # an emitter and its ancestor never appear in the same row: the walk stops at exactly one level
    par = t["parent"][cg]
    key = rows.astype(np.int64) * len(rl) + cg
    assert not np.isin(rows.astype(np.int64) * len(rl) + par, key)[par >= 0].any()


def test_hierarchy_shrinks_the_matrix(hier_scene, hier_pair):
    """Same map, leaf patches only, flat build: several times more transfers than the hierarchical rows."""
    from vrad_b200.environment import Environment
    g, o, nnz_g, _ = hier_pair
    leaf = np.nonzero(hier_scene.meta["tree"]["child1"] == -1)[0]
    f = Environment(); f.add_triangles(hier_scene.tri_ids, hier_scene.tri_verts, hier_scene.tri_flags); f.setup_acceleration_structure()
    f.patches_upload(hier_scene.patch_origin[leaf], hier_scene.patch_normal[leaf], hier_scene.patch_plane_dist[leaf],
                     hier_scene.patch_area[leaf], hier_scene.patch_refl[leaf], hier_scene.patch_cluster[leaf])
    nnz_flat = f.build_transfers(hier_scene.pvs)
    f.close()
    assert nnz_flat > 3 * nnz_g


def test_bounce_with_collect_light(hier_scene, hier_pair):
    g, o, _, _ = hier_pair
    N = hier_scene.n_patches
    emit0 = scenes.SplitMix64(11).uniform(3 * N, 0.0, 200.0).reshape(N, 3)
    tg, ag, dg = g.bounce(emit0, 6)
    to, ao, do = o.bounce(emit0, 6, threads=8)
    assert dg == do == 6
    assert np.abs(tg - to).max() <= RTOL * np.abs(to).max()
    assert np.allclose(ag, ao, rtol=RTOL)
    # interior patches hold the area-weighted average of their children (CollectLight)
    t = hier_scene.meta["tree"]
    k = np.nonzero(t["child1"] != -1)[0]
    c1, c2 = t["child1"][k], t["child2"][k]
    a1, a2 = hier_scene.patch_area[c1], hier_scene.patch_area[c2]
    want = (tg[c1] * (a1 / (a1 + a2))[:, None]) + (tg[c2] * (a2 / (a1 + a2))[:, None])
    assert np.abs(tg[k] - want).max() <= 1e-4 * np.abs(tg).max()
    assert (tg[k] > 0).mean() > 0.9
    te, ae, de = g.bounce(emit0, 100, early_out=True)
    teo, aeo, deo = o.bounce(emit0, 100, early_out=True, threads=8)
    assert de == deo and np.abs(te - teo).max() <= RTOL * np.abs(teo).max()


def test_hierarchy_validation(hier_scene):
    from vrad_b200.environment import VradError, environment_from_scene
    t = hier_scene.meta["tree"]
    e = environment_from_scene(hier_scene)
    bad = t["child2"].copy(); bad[0] = -1
    with pytest.raises(VradError):
        e.set_hierarchy(t["parent"], t["child1"], bad)                          # one child only
    bad = t["parent"].copy(); bad[t["child1"][0]] = 5
    with pytest.raises(VradError):
        e.set_hierarchy(bad, t["child1"], t["child2"])                          # parent link does not match
    with pytest.raises(VradError):
        e.set_hierarchy(t["parent"][:10], t["child1"][:10], t["child2"][:10])   # wrong count
    refl = hier_scene.patch_refl.copy(); refl[t["child1"][0]] *= 0.5
    e.patches_upload(hier_scene.patch_origin, hier_scene.patch_normal, hier_scene.patch_plane_dist, hier_scene.patch_area, refl,
                     hier_scene.patch_cluster)
    with pytest.raises(VradError):
        e.set_hierarchy(t["parent"], t["child1"], t["child2"])                  # a child with its own reflectivity
    e.close()


def test_per_candidate_form_gives_the_same_rows(hier_scene, hier_pair, tmp_path):
    """The top-down kernel (default) and the per-candidate kernel (fallback when a shared-memory list overflows;
    forced with VRAD_K2_TOPDOWN=0, read once per process) must produce identical transfer lists."""
    import hashlib, os, subprocess, sys
    g = hier_pair[0]
    rg, cg, wg = g.transfers_download()
    want = hashlib.sha256(rg.tobytes() + cg.tobytes() + wg.tobytes()).hexdigest()
    code = (
        "import sys, hashlib; sys.path.insert(0, %r)\n"
        "from vrad_b200 import scenes\n"
        "from vrad_b200.environment import environment_from_scene\n"
        "s = scenes.multi_room_hier(nx=3, ny=2); t = s.meta['tree']\n"
        "e = environment_from_scene(s); e.set_hierarchy(t['parent'], t['child1'], t['child2'], t['face'])\n"
        "e.build_transfers(s.pvs); r, c, w = e.transfers_download()\n"
        "print('HASH', hashlib.sha256(r.tobytes() + c.tobytes() + w.tobytes()).hexdigest())\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, VRAD_K2_TOPDOWN="0"), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    got = [l.split()[1] for l in out.stdout.splitlines() if l.startswith("HASH")][0]
    assert got == want
