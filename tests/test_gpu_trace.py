"""K1 parity on the GPU: every call goes through the C-ABI (libvradcuda.so) and is compared with
the CPU oracle on the same seeded inputs.  Hit indices, surface ids, hit distances and visibility
bits must match bit for bit."""
import numpy as np
import pytest

from vrad_b200 import scenes

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.unpackbits(np.ascontiguousarray(a).view(np.uint8), bitorder="little")


def test_tree_matches_oracle(s1_gpu, s1_oracle, s2_small_gpu, s2_small_oracle):
    for g, o in ((s1_gpu, s1_oracle), (s2_small_gpu, s2_small_oracle)):
        tg, to = g.download_tree(), o.export()
        assert np.array_equal(tg["children"], to["children"])
        assert np.array_equal(tg["split"].view(np.uint32), to["split"].view(np.uint32))
        assert np.array_equal(tg["tri_index"], to["tri_index"])
        assert tg["tris"].tobytes() == to["tris"].tobytes()
        assert np.array_equal(tg["aabb"], to["aabb"])


@pytest.mark.parametrize("which", ["s1", "s2"])
def test_closest_hit_bit_exact(which, s1_scene, s1_gpu, s1_oracle, s2_small_scene, s2_small_gpu, s2_small_oracle):
    scene, g, o = (s1_scene, s1_gpu, s1_oracle) if which == "s1" else (s2_small_scene, s2_small_gpu, s2_small_oracle)
    r = scenes.random_rays(scene, 1 << 18, seed=0xBEEF + len(which))
    gt, gs, gd = g.trace_rays(r["o"], r["d"], r["tmax"])
    ot, os_, od = o.trace1(r["o"], r["d"], r["tmax"], threads=8)
    assert np.array_equal(gt, ot)
    assert np.array_equal(gs, os_)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    assert (gt >= 0).mean() > 0.99          # closed rooms: (almost) every ray hits something


def test_closest_hit_vs_brute_force(s1_scene, s1_gpu, s1_oracle):
    r = scenes.random_rays(s1_scene, 1 << 15, seed=77)
    gt, _, gd = s1_gpu.trace_rays(r["o"], r["d"], r["tmax"])
    bt, _, bd = s1_oracle.trace_brute(r["o"], r["d"], r["tmax"], threads=8)
    assert np.array_equal(gt, bt)
    assert np.array_equal(gd.view(np.uint32), bd.view(np.uint32))


def test_skip_id_and_tmin(s1_scene, s1_gpu, s1_oracle):
    r = scenes.random_rays(s1_scene, 1 << 14, seed=5)
    tmin = np.full(r["tmax"].shape, 3.0, np.float32)
    # skipping the id every world triangle+1 carries (AddQuad's second triangle) changes results
    skip = scenes.TRACE_ID_OPAQUE + 1
    g = s1_gpu.trace_rays(r["o"], r["d"], r["tmax"], tmin=tmin, skip_id=skip)
    o = s1_oracle.trace1(r["o"], r["d"], r["tmax"], tmin=tmin, skip_id=skip, threads=8)
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1])
    assert np.array_equal(g[2].view(np.uint32), o[2].view(np.uint32))
    assert not np.any(g[1] == skip)


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, (1 << 18) + 7])
def test_test_lines_bit_exact_ragged(n, s1_scene, s1_gpu, s1_oracle):
    a, b = scenes.shadow_segments(s1_scene, n, seed=n)
    gb = s1_gpu.test_lines(a, b)
    ob = s1_oracle.test_lines(a, b, threads=8)
    assert np.array_equal(gb, ob)


def test_test_lines_s2_and_brute(s2_small_scene, s2_small_gpu, s2_small_oracle):
    a, b = scenes.shadow_segments(s2_small_scene, 1 << 17, seed=9)
    gb = s2_small_gpu.test_lines(a, b)
    assert np.array_equal(gb, s2_small_oracle.test_lines(a, b, threads=8))
    nb = 1 << 14
    bb = s2_small_oracle.test_lines(a[:, :nb].copy(), b[:, :nb].copy(), mode=2, threads=8)
    assert np.array_equal(gb[: nb // 32], bb)


def test_zero_length_and_degenerate_segments(s1_gpu, s1_oracle):
    a = np.array([[0, 10, 0, 100], [0, 10, 0, 100], [100, 50, 100, 100]], np.float32)
    b = a.copy()
    b[:, 2] = [0, 0, 600]        # leaves the room through the ceiling -> occluded
    b[:, 3] = [100, 100, 100.000001]
    gb = s1_gpu.test_lines(a, b)
    ob = s1_oracle.test_lines(a, b)
    assert np.array_equal(gb, ob)
    assert gb[0] & 1 and gb[0] & 2 and not (gb[0] & 4)


def test_trace4_packet_matches_result_layout(s1_scene, s1_gpu, s1_oracle):
    r = scenes.random_rays(s1_scene, 64, seed=3)
    for p in range(16):
        o = r["o"][:, 4 * p:4 * p + 4]; d = r["d"][:, 4 * p:4 * p + 4]
        tmin = np.zeros(4, np.float32); tmax = r["tmax"][4 * p:4 * p + 4]
        gi, gd, gn = s1_gpu.trace4_rays(o, d, tmin, tmax)
        oi, od, on = s1_oracle.trace4_packet(o, d, tmin, tmax)
        assert np.array_equal(gi, oi)
        assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
        assert np.array_equal(gn.view(np.uint32), on.view(np.uint32))


def test_sky_mode(s1_scene):
    """A room whose ceiling is TRACE_ID_SKY: up-rays see sky in sky_mode=1, are blocked in mode 0."""
    from oracle import pyoracle
    from vrad_b200.environment import Environment
    ids = s1_scene.tri_ids.copy()
    ids[2:4] = scenes.TRACE_ID_SKY           # the two ceiling triangles
    g = Environment(); g.add_triangles(ids, s1_scene.tri_verts); g.setup_acceleration_structure()
    o = pyoracle.OracleEnv(); o.add_triangles(ids, s1_scene.tri_verts); o.build()
    n = 1 << 14
    rng = scenes.SplitMix64(11)
    a = np.stack([rng.uniform(n, -500, 500), rng.uniform(n, -500, 500), rng.uniform(n, 100, 500)])
    b = a + np.stack([rng.uniform(n, -300, 300), rng.uniform(n, -300, 300), np.full(n, 20000.0, np.float32)])
    for mode in (0, 1):
        gb = g.test_lines(a, b, sky_mode=mode); ob = o.test_lines(a, b, sky_mode=mode, threads=8)
        assert np.array_equal(gb, ob)
    assert _bits(g.test_lines(a, b, sky_mode=0)).sum() == 0
    assert _bits(g.test_lines(a, b, sky_mode=1))[:n].mean() > 0.5
    g.close()


def test_device_resident_io(s1_scene, s1_gpu, s1_oracle):
    torch = pytest.importorskip("torch")
    a, b = scenes.shadow_segments(s1_scene, 100000, seed=21)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = s1_gpu.test_lines(ta, tb)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), s1_oracle.test_lines(a, b, threads=8))


def test_errors_are_loud(s1_scene):
    from vrad_b200.environment import Environment, VradError
    e = Environment()
    with pytest.raises(VradError):
        e.trace_rays(np.zeros((3, 4), np.float32), np.ones((3, 4), np.float32), np.ones(4, np.float32))   # not built
    with pytest.raises(VradError):
        e.test_lines(np.zeros((3, 4), np.float32), np.ones((3, 4), np.float32))                             # not built
    e.add_triangles(s1_scene.tri_ids, s1_scene.tri_verts)
    e.setup_acceleration_structure()
    with pytest.raises(VradError):
        e.setup_acceleration_structure()
    e.close()


def test_pipelined_host_path_matches_device_path(s1_scene, s1_gpu, s1_oracle):
    """Host batches >= 2^22 segments take the chunked, copy/compute-overlapped path: same bits."""
    torch = pytest.importorskip("torch")
    n = (1 << 22) + 12345
    a, b = scenes.shadow_segments(s1_scene, n, seed=4242)
    host_bits = s1_gpu.test_lines(a, b)
    dev_bits = s1_gpu.test_lines(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(host_bits, dev_bits.cpu().numpy().view(np.uint32))
    ns = 1 << 16
    assert np.array_equal(host_bits[: ns // 32], s1_oracle.test_lines(a[:, :ns].copy(), b[:, :ns].copy(), threads=8))
    tail0 = (n // 32) * 32 - 64
    assert np.array_equal(host_bits[tail0 // 32:], s1_oracle.test_lines(a[:, tail0:].copy(), b[:, tail0:].copy(), threads=8))
