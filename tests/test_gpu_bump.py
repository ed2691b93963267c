"""GPU parity for the bump-mapped (4-normal) gather: TotalLight.Light[1..3] of the bump-mapped leaf patches through the C-ABI
against the oracle (upstream GatherLight's bump branch), within 1e-4 relative (fp32; the warp reduction reorders the sums)."""
import numpy as np
import pytest

from vrad_b200 import scenes

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _face_bases(normals):
    from vrad_b200.environment import bump_normals
    bn = np.zeros((normals.shape[0], 3, 3), np.float32)
    for n_ in np.unique(normals, axis=0):
        s = np.cross(n_, [0, 0, 1]) if abs(n_[2]) < 0.9 else np.float32([1, 0, 0])
        bn[np.all(normals == n_, axis=1)] = bump_normals(s, np.cross(n_, s), n_, n_)
    return bn


def _rel(a, b):
    return float(np.abs(a - b).max()) / max(float(np.abs(b).max()), 1e-30)


def test_bump_totals_flat_patches(s2_small_scene):
    from oracle import pyoracle
    from vrad_b200.environment import VradError, environment_from_scene
    sc = s2_small_scene
    N = sc.n_patches
    g = environment_from_scene(sc); o = pyoracle.env_from_scene(sc)
    with pytest.raises(VradError):
        g.bump_totals()                                                      # nothing set yet
    bn = _face_bases(sc.patch_normal)
    flags = (np.arange(N) % 3 != 0).astype(np.uint8)
    with pytest.raises(VradError):
        g.set_bump(flags[:10], bn[:10])                                      # wrong count
    g.set_bump(flags, bn); o.set_bump(flags, bn)
    assert g.build_transfers(sc.pvs) == o.build_transfers(sc.pvs, threads=8)
    emit = scenes.SplitMix64(21).uniform(3 * N, 0.0, 200.0).reshape(N, 3)
    tg, _, dg = g.bounce(emit, 4)
    to, _, do = o.bounce(emit, 4, threads=8)
    assert dg == do == 4 and _rel(tg, to) <= RTOL                            # the flat light is untouched by the bump pass
    bg, bo = g.bump_totals(), o.bump_totals()
    assert _rel(bg, bo) <= RTOL
    assert np.all(bg[flags == 0] == 0) and (bg[flags == 1].sum(axis=(1, 2)) > 0).mean() > 0.9
    lit = tg.sum(axis=1) > 1.0
    ratio = bg[lit & (flags == 1)].sum(axis=(1, 2)) / tg[lit & (flags == 1)].sum(axis=1)
    assert np.all(ratio >= np.sqrt(3) * (1 - 1e-3)) and np.all(ratio < 4.0)  # sum_b n_b = sqrt(3) n, negative cosines clipped
    # a second call starts the bump totals from zero again
    g.bounce(emit, 1); o.bounce(emit, 1, threads=8)
    assert _rel(g.bump_totals(), o.bump_totals()) <= RTOL
    g.close()


def test_bump_totals_with_patch_hierarchy():
    from oracle import pyoracle
    from vrad_b200.environment import environment_from_scene
    hs = scenes.multi_room_hier(nx=2, ny=1, boxes_per_room=6)
    t = hs.meta["tree"]
    N = hs.n_patches
    g = environment_from_scene(hs); o = pyoracle.env_from_scene(hs)
    for env in (g, o):
        env.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
    bn = _face_bases(hs.patch_normal)
    flags = np.ones(N, np.uint8)
    g.set_bump(flags, bn); o.set_bump(flags, bn)
    assert g.build_transfers(hs.pvs) == o.build_transfers(hs.pvs, threads=8)
    emit = scenes.SplitMix64(22).uniform(3 * N, 0.0, 200.0).reshape(N, 3)
    tg, _, _ = g.bounce(emit, 3)
    to, _, _ = o.bounce(emit, 3, threads=8)
    assert _rel(tg, to) <= RTOL
    bg, bo = g.bump_totals(), o.bump_totals()
    assert _rel(bg, bo) <= RTOL
    interior = t["child1"] != -1
    assert np.all(bg[interior] == 0) and (bg[~interior].sum(axis=(1, 2)) > 0).mean() > 0.8      # only leaf patches gather
    g.close()
