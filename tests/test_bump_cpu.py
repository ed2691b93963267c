"""CPU-only: the bump-mapped (4-normal) gather of the oracle -- upstream GatherLight's bump branch and GetBumpNormals, for which
the reference only carries the types (Patch.NeedsBumpMap common/types/patch.go:23, BumpLights common/types/bumpLights.go:8-10,
NUM_BUMP_VECTS common/constants/constants.go:33).  Hand-derived values and the basis identity sum_b n_b = sqrt(3) n."""
import numpy as np

from oracle import pyoracle
from vrad_b200 import scenes
from vrad_b200.environment import bump_normals


def test_bump_basis():
    b = bump_normals([1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 1])
    r23, r6, r2, r3 = np.float32(0.81649661), np.float32(0.40824822), np.float32(0.70710677), np.float32(0.57735026)
    assert np.allclose(b, [[r23, 0, r3], [-r6, r2, r3], [-r6, -r2, r3]], atol=1e-7)
    assert np.allclose(b.sum(axis=0), [0, 0, np.sqrt(3)], atol=1e-6) and np.allclose(np.linalg.norm(b, axis=1), 1, atol=1e-6)
    # left-handed texture axes mirror the second basis vector
    bl = bump_normals([1, 0, 0], [0, -1, 0], [0, 0, 1], [0, 0, 1])
    assert np.allclose(bl[0], b[0]) and np.allclose(bl[1], b[2]) and np.allclose(bl[2], b[1])
    # a tilted phong normal tilts the basis with it; the three normals keep their mutual angles
    ph = np.float32([0.1, -0.2, 0.97]); ph /= np.linalg.norm(ph)
    bt = bump_normals([1, 0, 0], [0, 1, 0], [0, 0, 1], ph)
    assert np.allclose(bt.sum(axis=0), np.sqrt(3) * ph, atol=1e-5) and np.allclose(bt @ bt.T, b @ b.T, atol=1e-5)
    for args in (([1, 0, 0], [0, 1, 0], [0, 0, 1], ph), ([0.3, 2, 0.1], [1, 0, -4], [0, 0, 1], [0, 0, 1])):
        assert bump_normals(*args).tobytes() == pyoracle.bump_normals(*args).tobytes()


def _plates():
    """A receiver at the origin facing +z and two emitters facing -z: one straight above, one 45 degrees off along +x."""
    o = pyoracle.OracleEnv()
    o.add_triangles([scenes.TRACE_ID_OPAQUE], np.array([[9000, 9000, 9000, 9001, 9000, 9000, 9000, 9001, 9000]], np.float32)); o.build()
    origin = np.array([[0, 0, 0], [0, 0, 100], [100, 0, 100]], np.float32)
    normal = np.array([[0, 0, 1], [0, 0, -1], [0, 0, -1]], np.float32)
    pdist = np.array([0, -100, -100], np.float32)
    area = np.array([16, 16, 16], np.float32)
    refl = np.full((3, 3), 0.5, np.float32)
    o.patches_upload(origin, normal, pdist, area, refl, np.zeros(3, np.int32))
    return o, origin, normal, area


def test_bump_gather_hand_example():
    o, origin, normal, area = _plates()
    basis = bump_normals([1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 1])
    bn = np.zeros((3, 3, 3), np.float32); bn[0] = basis
    o.set_bump([1, 0, 0], bn)
    assert o.build_transfers(None) == 4
    rowptr, col, w = o.transfers()
    assert list(col[rowptr[0]:rowptr[1]]) == [1, 2]
    emit = np.zeros((3, 3), np.float32); emit[1] = 100.0; emit[2] = [40.0, 0.0, 0.0]
    tot, _, _ = o.bounce(emit, 1)
    bt = o.bump_totals()
    w1, w2 = w[rowptr[0]], w[rowptr[0] + 1]
    d1 = np.float64([0, 0, 1]); d2 = np.float64([1, 0, 1]) / np.sqrt(2)
    want = np.zeros((3, 3))
    for d, wt, e in ((d1, w1, [50.0, 50.0, 50.0]), (d2, w2, [20.0, 0.0, 0.0])):      # emit * refl
        v = np.float64(e) * wt / d[2]                                                 # transfer with the flat cosine taken out again
        for b in range(3):
            dot = float(d @ basis[b].astype(np.float64))
            if dot > 0:
                want[b] += v * dot
    assert np.allclose(bt[0], want, rtol=1e-5) and np.all(bt[1:] == 0)
    assert np.allclose(tot[0], 50.0 * w1 + np.float64([20.0, 0, 0]) * w2, rtol=1e-6)  # Light[0] is the flat gather
    # both emitters lie inside the cone where all three bump cosines are positive: sum_b Light[b] = sqrt(3) * Light[0]
    assert np.allclose(bt[0].sum(axis=0), np.sqrt(3) * tot[0], rtol=1e-5)
    # bump totals accumulate over bounces like Light[0] does, only leaf patches with the flag get them
    tot2, _, _ = o.bounce(emit, 3)
    assert np.all(o.bump_totals()[0] >= bt[0] - 1e-6) and np.all(o.bump_totals()[1:] == 0)


def test_bump_totals_on_the_room_scene_track_the_flat_light(s2_small_scene, s2_small_oracle):
    """On the multi-room cut: every patch bump-mapped with its face's axis basis; where light arrives from the whole hemisphere
    the three bump sums together are sqrt(3) x the flat light plus what the clipped negative cosines leave out (grazing
    emitters), i.e. a ratio somewhat above sqrt(3)."""
    sc = s2_small_scene
    N = sc.n_patches
    bn = np.zeros((N, 3, 3), np.float32)
    for n_ in np.unique(sc.patch_normal, axis=0):
        s = np.cross(n_, [0, 0, 1]) if abs(n_[2]) < 0.9 else np.float32([1, 0, 0])
        t = np.cross(n_, s)
        bn[np.all(sc.patch_normal == n_, axis=1)] = bump_normals(s, t, n_, n_)
    s2_small_oracle.set_bump(np.ones(N, np.uint8), bn)
    if not hasattr(s2_small_oracle, "nnz"):
        s2_small_oracle.build_transfers(sc.pvs, threads=8)
    emit = np.full((N, 3), 100.0, np.float32)
    tot, _, _ = s2_small_oracle.bounce(emit, 3, threads=8)
    bt = s2_small_oracle.bump_totals()
    s2_small_oracle.set_bump(np.zeros(N, np.uint8), bn)                                # leave the shared fixture flat again
    lit = tot.sum(axis=1) > 1.0
    ratio = bt[lit].sum(axis=1).sum(axis=1) / tot[lit].sum(axis=1)
    assert np.all(ratio >= np.sqrt(3) * (1 - 1e-4)) and np.all(ratio < 4.0) and 1.8 < np.median(ratio) < 2.5
