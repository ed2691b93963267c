"""K2 / K3 / K4 parity on the GPU against the CPU oracle, through the C-ABI."""
import numpy as np
import pytest

from vrad_b200 import scenes

pytestmark = pytest.mark.gpu

RTOL = 1e-4      # north_star: lightmap and patch radiance within 1e-4 relative (fp32)


def _rel_err(a, b):
    scale = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / scale


@pytest.fixture(scope="module")
def s1_transfers(s1_gpu, s1_oracle):
    nnz_g = s1_gpu.build_transfers()
    nnz_o = s1_oracle.build_transfers(threads=8)
    return nnz_g, nnz_o


def test_transfers_s1_exact(s1_gpu, s1_oracle, s1_transfers):
    nnz_g, nnz_o = s1_transfers
    assert nnz_g == nnz_o and nnz_g > 1_000_000
    rg, cg, wg = s1_gpu.transfers_download()
    ro, co, wo = s1_oracle.transfers()
    assert np.array_equal(rg, ro)
    assert np.array_equal(cg, co)                                   # columns bit-exact
    assert np.array_equal(wg.view(np.uint32), wo.view(np.uint32))   # same op order -> weights bit-exact too
    sums = np.add.reduceat(wg.astype(np.float64), rg[:-1][np.diff(rg) > 0])
    assert sums.max() <= 1.0 + 1e-5                                 # MakeScales cap


def test_transfers_s2_pvs_exact(s2_small_scene, s2_small_gpu, s2_small_oracle):
    nnz_g = s2_small_gpu.build_transfers(s2_small_scene.pvs)
    nnz_o = s2_small_oracle.build_transfers(s2_small_scene.pvs, threads=8)
    assert nnz_g == nnz_o
    rg, cg, wg = s2_small_gpu.transfers_download()
    ro, co, wo = s2_small_oracle.transfers()
    assert np.array_equal(rg, ro) and np.array_equal(cg, co)
    assert np.array_equal(wg.view(np.uint32), wo.view(np.uint32))
    # PVS honoured: no transfer crosses between clusters that do not see each other
    rows = np.repeat(np.arange(len(rg) - 1), np.diff(rg))
    ci, cj = s2_small_scene.patch_cluster[rows], s2_small_scene.patch_cluster[cg]
    assert np.all(s2_small_scene.pvs[ci, cj] == 1)


def test_direct_light_s1(s1_scene, s1_gpu, s1_oracle):
    g = s1_gpu.direct_light(s1_scene.luxel_pos, s1_scene.luxel_normal, s1_scene.lights)
    o = s1_oracle.direct_light(s1_scene.luxel_pos, s1_scene.luxel_normal, s1_scene.lights, threads=8)
    assert _rel_err(g, o) <= RTOL
    assert np.array_equal(g.view(np.uint32), o.view(np.uint32))     # exponent 1 -> no powf -> bit-exact
    assert (g.sum(axis=1) > 0).mean() > 0.5


def test_direct_light_sky_and_fades(s1_scene):
    """Sky ceiling + sun + sky ambient + hard-falloff point light + exponent-2 spot."""
    from oracle import pyoracle
    from vrad_b200.environment import Environment
    ids = s1_scene.tri_ids.copy(); ids[2:4] = scenes.TRACE_ID_SKY
    g = Environment(); g.add_triangles(ids, s1_scene.tri_verts); g.setup_acceleration_structure()
    o = pyoracle.OracleEnv(); o.add_triangles(ids, s1_scene.tri_verts); o.build()
    dirs = np.loadtxt(__import__("os").path.join(__import__("os").path.dirname(__file__), "..", "vrad_b200", "data", "anorms.txt"), dtype=np.float32)
    assert dirs.shape == (162, 3)
    g.set_sky_dirs(dirs); o.set_sky_dirs(dirs)
    L = np.zeros(5, scenes.LIGHT_DTYPE)
    L["start_fade"], L["end_fade"], L["cap_dist"] = 0.0, -1.0, 1e22
    sun = np.array([-0.4, -0.3, -0.87]); sun /= np.linalg.norm(sun)
    L[0]["type"] = scenes.EMIT_SKYLIGHT; L[0]["normal"] = sun; L[0]["intensity"] = (300, 280, 250)
    L[1]["type"] = scenes.EMIT_SKYAMBIENT; L[1]["intensity"] = (40, 50, 70)
    L[2]["type"] = scenes.EMIT_POINT; L[2]["origin"] = (0, 0, 400); L[2]["intensity"] = (2e6, 2e6, 2e6)
    L[2]["quadratic_attn"] = 1.0; L[2]["start_fade"], L[2]["end_fade"] = 300.0, 600.0
    L[3]["type"] = scenes.EMIT_SPOTLIGHT; L[3]["origin"] = (200, -100, 450); L[3]["normal"] = (0, 0, -1)
    L[3]["stopdot"], L[3]["stopdot2"], L[3]["exponent"] = 0.9, 0.6, 2.0
    L[3]["quadratic_attn"] = 1.0; L[3]["intensity"] = (3e6, 1e6, 1e6)
    L[4]["type"] = scenes.EMIT_SURFACE; L[4]["origin"] = (-300, 300, 500); L[4]["normal"] = (0, 0, -1); L[4]["intensity"] = (1e6, 1e6, 3e6)
    pos, nrm = s1_scene.luxel_pos[::3], s1_scene.luxel_normal[::3]
    gg = g.direct_light(pos, nrm, L); oo = o.direct_light(pos, nrm, L, threads=8)
    assert _rel_err(gg, oo) <= RTOL
    for k in range(5):      # every light type contributes somewhere
        gk = g.direct_light(pos, nrm, L[k:k + 1]); ok = o.direct_light(pos, nrm, L[k:k + 1], threads=8)
        assert gk.max() > 0 and _rel_err(gk, ok) <= RTOL
    g.close()


def test_bounce_s1(s1_scene, s1_gpu, s1_oracle, s1_transfers):
    N = s1_scene.n_patches
    rng = scenes.SplitMix64(99)
    emit0 = rng.uniform(3 * N, 0.0, 200.0).reshape(N, 3)
    for nb in (1, 8):
        tg, ag, dg = s1_gpu.bounce(emit0, nb)
        to, ao, do = s1_oracle.bounce(emit0, nb, threads=8)
        assert dg == do == nb
        assert _rel_err(tg, to) <= RTOL
        assert np.allclose(ag, ao, rtol=RTOL)
    # energy: bounced light shrinks geometrically and early-out terminates before the cap
    tg, ag, dg = s1_gpu.bounce(emit0, 100, early_out=True)
    to, ao, do = s1_oracle.bounce(emit0, 100, early_out=True, threads=8)
    assert dg == do and dg < 100
    assert _rel_err(tg, to) <= RTOL


def test_bounce_uploaded_csr_and_linearity(s1_scene, s1_gpu, s1_oracle, s1_transfers):
    """K4 on an uploaded CSR (the path the multi-GPU shards use) + linearity in emit0."""
    from vrad_b200.environment import environment_from_scene
    ro, co, wo = s1_oracle.transfers()
    N = s1_scene.n_patches
    env = environment_from_scene(s1_scene)
    env.transfers_upload(0, N, ro, co, wo)
    rng = scenes.SplitMix64(5)
    e1 = rng.uniform(3 * N, 0.0, 100.0).reshape(N, 3); e2 = rng.uniform(3 * N, 0.0, 100.0).reshape(N, 3)
    t1, _, _ = env.bounce(e1, 4); t2, _, _ = env.bounce(e2, 4); t12, _, _ = env.bounce(e1 + e2, 4)
    assert _rel_err(t12, t1 + t2) <= RTOL
    to, _, _ = s1_oracle.bounce(e1, 4, threads=8)
    assert _rel_err(t1, to) <= RTOL
    r2, c2, w2 = env.transfers_download()
    assert np.array_equal(r2, ro) and np.array_equal(c2, co) and np.array_equal(w2, wo)
    env.close()


def test_bounce_ragged_rows_and_sky():
    """Hand-made CSR with empty rows, 1..9-entry rows and a sky patch, against the oracle gather."""
    from oracle import pyoracle
    from vrad_b200.environment import Environment
    N = 64
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 10, N); lens[5] = 0; lens[6] = 0
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    col = rng.integers(0, N, rowptr[-1]).astype(np.int32)
    w = rng.random(rowptr[-1]).astype(np.float32) * 0.1
    refl = rng.random((N, 3)).astype(np.float32) * 0.7
    flags = np.zeros(N, np.uint8); flags[7] = 1
    origin = rng.random((N, 3)).astype(np.float32); normal = np.tile(np.array([0, 0, 1], np.float32), (N, 1))
    env = Environment()
    env.add_triangles(np.array([1], np.int32), np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)); env.setup_acceleration_structure()
    env.patches_upload(origin, normal, np.zeros(N, np.float32), np.ones(N, np.float32), refl, None, flags)
    env.transfers_upload(0, N, rowptr, col, w)
    emit0 = rng.random((N, 3)).astype(np.float32) * 100
    tg, ag, _ = env.bounce(emit0, 3)
    # oracle: run the same recurrence with orc_gather_rows
    emit = emit0.copy(); emit[7] = 0            # k4_init: sky patches emit nothing
    total = np.zeros_like(emit0)
    for _ in range(3):
        add = pyoracle.gather_rows(0, N, rowptr, col, w, emit, refl)
        add[7] = 0
        total += add; emit = add
    assert _rel_err(tg, total) <= RTOL
    assert np.allclose(ag, emit.sum(axis=0), rtol=1e-4)
    env.close()

