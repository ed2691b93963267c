"""Round-2 paths on the GPU, through the C-ABI, against the CPU oracle on the same seeded inputs:
K1 with the coherence pre-pass (sorted order), with index pairs into a resident point table, with the top of the
kd tree staged in shared memory; K4 with split rows, item order, and the bounce loop replayed as a CUDA graph.
Visibility bits are bit-exact in every mode; K4 stays within the 1e-4 relative tolerance north_star states."""
import numpy as np
import pytest

from vrad_b200 import scenes
from vrad_b200.lib import VradError

pytestmark = pytest.mark.gpu


@pytest.fixture()
def s1_env(s1_scene):
    from vrad_b200.environment import environment_from_scene
    env = environment_from_scene(s1_scene)
    yield env
    env.close()


@pytest.mark.parametrize("n", [1, 33, 1000, 65536 + 5, (1 << 19) + 7])
@pytest.mark.parametrize("sky", [0, 1])
def test_sorted_order_is_invisible(n, sky, s1_scene, s1_env, s1_oracle):
    a, b = scenes.shadow_segments(s1_scene, n, seed=100 + n)
    ref = s1_oracle.test_lines(a, b, sky_mode=sky, threads=8)
    for mode in (0, 1):                      # never / always order the batch
        s1_env.set_option("k1_sort", mode)
        assert np.array_equal(s1_env.test_lines(a, b, sky_mode=sky), ref), f"k1_sort={mode}"


def test_sorted_order_device_buffers_and_degenerate_segments(s1_scene, s1_env, s1_oracle):
    import torch
    n = (1 << 18) + 3
    a, b = scenes.shadow_segments(s1_scene, n, seed=5)
    b[:, ::7] = a[:, ::7]                    # zero-length segments: visible by definition (testline.go:22-27)
    a[:, 11] = [-1e6, 2e6, 5.0]; b[:, 11] = [3e6, -2e6, 7.0]   # far outside the scene box: clamped cells, still traced
    ref = s1_oracle.test_lines(a, b, threads=8)
    s1_env.set_option("k1_sort", 1)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device="cuda")     # stale words must be overwritten, not or-ed into
    s1_env.test_lines(da, db, out=out)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref)


@pytest.mark.parametrize("sort", [0, 1])
def test_indexed_segments_equal_coordinates(sort, s1_scene, s1_env, s1_oracle):
    n = (1 << 18) + 9
    pts, pairs = scenes.shadow_segment_indices(s1_scene, n, seed=77)
    a, b = scenes.shadow_segments(s1_scene, n, seed=77)
    ref = s1_oracle.test_lines(a, b, threads=8)
    s1_env.set_option("k1_sort", sort)
    s1_env.points_upload(pts)
    assert np.array_equal(s1_env.test_lines_indexed(pairs), ref)
    import torch
    dp = torch.from_numpy(pairs).cuda()
    assert np.array_equal(s1_env.test_lines_indexed(dp).cpu().numpy().view(np.uint32), ref)


def test_indexed_segments_host_pipeline_and_errors(s1_scene, s1_env, s1_oracle):
    n = (1 << 22) + 64                         # >= 2^22 host pairs: chunked H2D overlapped with the traversal
    pts, pairs = scenes.shadow_segment_indices(s1_scene, n, seed=3)
    s1_env.points_upload(pts)
    got = s1_env.test_lines_indexed(pairs)
    ns = 1 << 16
    a, b = scenes.shadow_segments(s1_scene, n, seed=3)
    assert np.array_equal(got[: ns // 32], s1_oracle.test_lines(a[:, :ns].copy(), b[:, :ns].copy(), threads=8))
    assert np.array_equal(got, s1_env.test_lines(a, b))             # coordinate form on the same segments, all 2^22
    bad = pairs[:100].copy(); bad[17, 1] = pts.shape[0]
    with pytest.raises(VradError) as ei:
        s1_env.test_lines_indexed(bad)
    assert ei.value.status == -1
    import torch
    with pytest.raises(VradError):
        s1_env.test_lines_indexed(torch.from_numpy(bad).cuda())
    from vrad_b200.environment import environment_from_scene
    e2 = environment_from_scene(s1_scene, with_patches=False)
    with pytest.raises(VradError) as ei:                             # no point table yet
        e2.test_lines_indexed(pairs[:10])
    assert ei.value.status == -4
    e2.close()


@pytest.mark.parametrize("budget", [3, 64, 1023, 4000])
def test_top_levels_in_shared_memory(budget, s1_scene, s1_env, s1_oracle, s2_small_scene, s2_small_oracle):
    from vrad_b200.environment import environment_from_scene
    for scene, orc, env in ((s1_scene, s1_oracle, s1_env), (s2_small_scene, s2_small_oracle, None)):
        own = env is None
        if own:
            env = environment_from_scene(scene, with_patches=False)
        n = (1 << 17) + 1
        a, b = scenes.shadow_segments(scene, n, seed=budget)
        ref = orc.test_lines(a, b, threads=8)
        env.set_option("k1_top", budget)
        for sort in (0, 1):
            env.set_option("k1_sort", sort)
            assert np.array_equal(env.test_lines(a, b), ref), f"budget {budget} sort {sort}"
        env.set_option("k1_top", 0)
        assert np.array_equal(env.test_lines(a, b), ref)
        if own:
            env.close()


def _gather_reference(orc, scene, emit0, n):
    return orc.bounce(emit0, n, threads=8)


@pytest.mark.parametrize("seg,long_first,graph,block,persist", [(256, 0, 1, 256, 1), (256, 1, 0, 192, 0), (2048, 0, 1, 192, 1), (32768, 0, 0, 256, 0),
                                                                (512, 1, 1, 256, 1)])
def test_gather_items_any_plan_same_light(seg, long_first, graph, block, persist, s2_small_scene, s2_small_oracle):
    from vrad_b200.environment import environment_from_scene
    scene = s2_small_scene
    env = environment_from_scene(scene)
    env.set_option("k4_seg", seg); env.set_option("k4_long_first", long_first); env.set_option("k4_graph", graph)
    env.set_option("k4_block", block); env.set_option("k4_persist", persist)
    env.set_option("k4_items", 1)              # the multi-GPU (work-item) kernel on this one device: a one-rank peer table, exact results
    env.set_option("k4_pool", 0 if seg == 512 else 12)
    nnz = env.build_transfers(scene.pvs)
    assert nnz == s2_small_oracle.build_transfers(scene.pvs, threads=8)
    N = scene.n_patches
    emit0 = scenes.SplitMix64(21).uniform(3 * N, 0.0, 200.0).reshape(N, 3)
    for nb in (1, 7, 8):                       # below and above the graph threshold; odd and even (buffer parity)
        tg, ag, dg = env.bounce(emit0, nb)
        to, ao, do = _gather_reference(s2_small_oracle, scene, emit0, nb)
        assert dg == do == nb
        assert np.abs(tg - to).max() <= 1e-4 * np.abs(to).max()
        assert np.allclose(ag, ao, rtol=1e-4)
    t1, a1, _ = env.bounce(emit0, 8)           # replayed graph: identical bits (the plan fixes the summation order)
    t2, a2, _ = env.bounce(emit0, 8)
    assert np.array_equal(t1, t2) and np.array_equal(a1, a2)
    te, ae, de = env.bounce(emit0, 100, early_out=True)
    toe, aoe, doe = s2_small_oracle.bounce(emit0, 100, early_out=True, threads=8)
    assert de == doe and np.abs(te - toe).max() <= 1e-4 * np.abs(toe).max()
    # re-planning resident rows keeps the light
    env.set_option("k4_seg", 1024)
    t3, _, _ = env.bounce(emit0, 8)
    assert np.abs(t3 - t1).max() <= 1e-4 * np.abs(t1).max()
    env.close()


def test_split_rows_actually_occur(s1_scene, s1_oracle):
    """S1's rows average ~2000 transfers: with 256-entry items nearly every row is split, and the result must not move."""
    from vrad_b200.environment import environment_from_scene
    env = environment_from_scene(s1_scene)
    sel = slice(0, None, 2)
    args = (s1_scene.patch_origin[sel], s1_scene.patch_normal[sel], s1_scene.patch_plane_dist[sel], s1_scene.patch_area[sel], s1_scene.patch_refl[sel])
    env.patches_upload(*args)
    env.set_option("k4_seg", 256); env.set_option("k4_items", 1)
    nnz = env.build_transfers()
    n = args[0].shape[0]
    assert nnz / n > 512                         # rows long enough that 256-entry items split them
    from oracle import pyoracle
    o = pyoracle.env_from_scene(s1_scene)
    o.patches_upload(*args)
    assert o.build_transfers(threads=8) == nnz
    emit0 = scenes.SplitMix64(2).uniform(3 * n, 0.0, 100.0).reshape(n, 3)
    tg, ag, _ = env.bounce(emit0, 12)
    to, ao, _ = o.bounce(emit0, 12, threads=8)
    assert np.abs(tg - to).max() <= 1e-4 * np.abs(to).max() and np.allclose(ag, ao, rtol=1e-4)
    env.close()


def test_world1_rejects_a_partial_row_block(s2_small_scene, s2_small_oracle):
    """with world == 1 a transfer block other than [0, N) must be an error, not light on the wrong patches."""
    from vrad_b200.environment import environment_from_scene
    scene = s2_small_scene
    s2_small_oracle.build_transfers(scene.pvs, threads=8)
    rp, col, w = s2_small_oracle.transfers()
    N = scene.n_patches
    env = environment_from_scene(scene)
    a, b = 100, N - 50
    env.transfers_upload(a, b, rp[a:b + 1] - rp[a], col[rp[a]:rp[b]], w[rp[a]:rp[b]])
    with pytest.raises(VradError) as ei:
        env.bounce(np.ones((N, 3), np.float32), 2)
    assert ei.value.status == -4
    env.transfers_upload(0, N, rp, col, w)
    tg, _, _ = env.bounce(np.full((N, 3), 50.0, np.float32), 3)
    to, _, _ = s2_small_oracle.bounce(np.full((N, 3), 50.0, np.float32), 3, threads=8)
    assert np.abs(tg - to).max() <= 1e-4 * np.abs(to).max()
    env.close()


def test_simulated_peers_runs_the_multi_gpu_kernel_on_one_device(s2_small_scene):
    """k4_sim_peers: the world-4 slice of rank 1 through k4_gather_items<MULTI> (peer stores, in-kernel barrier, PDL chain,
    graph replay) on ONE device.  The light of the other ranks' rows is never refreshed, so only this rank's FIRST bounce is
    comparable: total after 1 bounce == the single-GPU rows of the block."""
    from vrad_b200.environment import Environment, environment_from_scene
    scene = s2_small_scene
    N = scene.n_patches
    emit0 = scenes.SplitMix64(4).uniform(3 * N, 0.0, 200.0).reshape(N, 3)
    full = environment_from_scene(scene)
    full.build_transfers(scene.pvs)
    t_full, _, _ = full.bounce(emit0, 1)
    full.close()
    env = environment_from_scene(scene, rank=1, world=4)
    env.set_option("k4_sim_peers", 1)
    env.build_transfers(scene.pvs)
    row0, row1, _ = env.transfers_info()
    assert 0 < row0 < row1 < N
    import torch
    d_emit = torch.from_numpy(emit0).cuda(); d_out = torch.empty_like(d_emit)
    env.bounce(d_emit, 1, out=d_out, want_added=False)
    got = d_out.cpu().numpy()
    assert np.abs(got[row0:row1] - t_full[row0:row1]).max() <= 1e-4 * np.abs(t_full).max()
    for graph in (0, 1):                           # 40 chained bounces, stream launches and graph replay: must terminate
        env.set_option("k4_graph", graph)
        for _ in range(2):
            env.bounce(d_emit, 40, out=d_out, want_added=False)
    torch.cuda.synchronize()
    assert np.isfinite(d_out.cpu().numpy()).all()
    env.close()


def test_packed_streams_rows_over_several_column_windows():
    """The 6-byte packed transfer streams (k4_pack): rows whose columns span several 65,536-column windows become several segments;
    the result must equal the {col,w} pair kernel's to rounding and a float64 gather, and the single-GPU kernel and the work-item
    (multi-GPU) kernel must agree bit for bit -- segments of a row are parts there, added in the same order."""
    from vrad_b200.environment import Environment
    rng = np.random.default_rng(5)
    N = 200_000
    env = Environment()
    env.add_triangles(np.int32([0]), np.float32([[0, 0, 0, 1, 0, 0, 0, 1, 0]])); env.setup_acceleration_structure()
    origin = rng.uniform(-100, 100, (N, 3)).astype(np.float32)
    normal = np.tile(np.float32([0, 0, 1]), (N, 1))
    refl = rng.uniform(0.2, 0.8, (N, 3)).astype(np.float32)
    env.patches_upload(origin, normal, np.zeros(N, np.float32), np.ones(N, np.float32), refl)
    # rows of 0 .. 900 entries: some empty, some inside one window, some over all four, some with > 32 in one window then a jump
    want = rng.integers(1, 900, N); want[np.arange(N) % 8 != 1] = 0; want[9] = 40_000     # most rows empty; one longer than a segment may be
    cols = []
    for i in range(N):
        if want[i] == 0: cols.append(np.empty(0, np.int32)); continue
        lo, hi = (max(0, i - 30_000), min(N, i + 30_000)) if i % 16 == 1 else (0, N)        # one window (mostly) / every window
        cols.append(np.unique(rng.integers(lo, hi, want[i])).astype(np.int32))
    lens = np.array([len(c) for c in cols])
    rp = np.zeros(N + 1, np.int64); np.cumsum(lens, out=rp[1:])
    col = np.concatenate(cols)
    w = rng.uniform(0.0, 1.0 / 900, rp[-1]).astype(np.float32)
    emit0 = rng.uniform(0, 100, (N, 3)).astype(np.float32)
    env.transfers_upload(0, N, rp, col, w)
    env.set_option("k4_short", 0)                  # rows this sparse would otherwise go to the short-row kernel
    out = {}
    for name, opts in (("pairs", {"k4_pack": 0}), ("packed", {"k4_pack": 1}), ("packed_items", {"k4_pack": 1, "k4_items": 1}),
                       ("pairs_items", {"k4_pack": 0, "k4_items": 1}), ("packed_items_seg512", {"k4_pack": 1, "k4_items": 1, "k4_seg": 512})):
        for k in ("k4_items", "k4_pack"): env.set_option(k, opts.get(k, 0))
        env.set_option("k4_seg", opts.get("k4_seg", 16384))
        out[name], _, _ = env.bounce(emit0, 3)
        out[name + "_1"], _, _ = env.bounce(emit0, 1)
    env.close()
    # float64 reference of three bounces
    import scipy.sparse as sp
    A = sp.csr_matrix((w.astype(np.float64), col, rp), shape=(N, N))
    emit = emit0.astype(np.float64); total = np.zeros((N, 3))
    for _ in range(3):
        add = A @ (emit * refl.astype(np.float64)); total += add; emit = add
    for name, t in out.items():
        if not name.endswith("_1"): assert np.abs(t - total).max() <= 1e-5 * np.abs(total).max(), name
    diff = np.nonzero((out["packed_1"] != out["packed_items_1"]).any(axis=1))[0]
    nwin = np.array([len(np.unique((c - c[0]) >> 16)) if len(c) else 0 for c in cols])
    cmp = {f"{a}~{b}": (int((out[a] != out[b]).any(axis=1).sum()), float(np.abs(out[a] - out[b]).max() / np.abs(out[a]).max()))
           for a, b in (("pairs_1", "pairs_items_1"), ("packed_1", "pairs_1"), ("packed_items_1", "pairs_items_1"), ("packed_1", "packed_items_1"))}
    assert cmp["packed_1~pairs_1"][0] > 1000 and cmp["packed_1~pairs_1"][1] < 1e-6          # the packed kernel did run: several-segment rows round differently
    assert cmp["pairs_1~pairs_items_1"][0] <= 1                                             # only the 36k-entry row is split in the pair plan
    assert diff.size == 0, (cmp, diff.size, diff[:8], nwin[diff[:8]], lens[diff[:8]], np.bincount(nwin[diff]), np.bincount(nwin))
    assert np.abs(out["packed_items_seg512"] - out["packed"]).max() <= 1e-6 * np.abs(total).max()


def test_block_row_streams_same_light(s2_small_scene, s2_small_oracle):
    """Block-row transfer streams (k4_pack = 2, the default): 4 consecutive rows share the union of their columns, one er[] gather serves
    the four.  Same light as the packed streams and the pairs to rounding, within 1e-4 of the oracle, and the single-GPU kernel and the
    work-item (multi-GPU) kernel agree bit for bit."""
    from vrad_b200.environment import environment_from_scene
    scene = s2_small_scene
    env = environment_from_scene(scene)
    nnz = env.build_transfers(scene.pvs)
    assert nnz == s2_small_oracle.build_transfers(scene.pvs, threads=8)
    pairs, packed, segs, blocked, rows = env.transfers_layout()
    assert blocked > 0 and packed == 0 and rows in (2, 4)   # block rows in use; their union is far below that many rows' worth of entries
    assert blocked * (2 + 4 * rows) < nnz * 6
    N = scene.n_patches
    emit0 = scenes.SplitMix64(33).uniform(3 * N, 0.0, 200.0).reshape(N, 3)
    out = {}
    for name, opts in (("block", {"k4_pack": 2}), ("block_items", {"k4_pack": 3, "k4_items": 1}), ("packed", {"k4_pack": 1}), ("pairs", {"k4_pack": 0})):
        for k in ("k4_items", "k4_pack"): env.set_option(k, opts.get(k, 0))
        out[name] = [env.bounce(emit0, nb) for nb in (1, 8)]            # below and above the graph threshold
    lay = {k: None for k in out}
    to1, ao1, _ = _gather_reference(s2_small_oracle, scene, emit0, 1)
    to8, ao8, _ = _gather_reference(s2_small_oracle, scene, emit0, 8)
    for name, ((t1, a1, _), (t8, a8, _)) in out.items():
        assert np.abs(t1 - to1).max() <= 1e-4 * np.abs(to1).max() and np.allclose(a1, ao1, rtol=1e-4), name
        assert np.abs(t8 - to8).max() <= 1e-4 * np.abs(to8).max() and np.allclose(a8, ao8, rtol=1e-4), name
    assert np.array_equal(out["block"][0][0], out["block_items"][0][0]) and np.array_equal(out["block"][1][0], out["block_items"][1][0])
    assert np.abs(out["block"][1][0] - out["pairs"][1][0]).max() <= 2e-6 * np.abs(to8).max()
    te, ae, de = (env.set_option("k4_pack", 2), env.bounce(emit0, 100, early_out=True))[1]
    toe, aoe, doe = s2_small_oracle.bounce(emit0, 100, early_out=True, threads=8)
    assert de == doe and np.abs(te - toe).max() <= 1e-4 * np.abs(toe).max()
    env.close()


def test_uploaded_rows_in_any_column_order_keep_the_pairs():
    """vrad_transfers_upload takes rows as given: columns unsorted or listed twice are legal for the {col,w} pair kernels; the packed and
    block-row streams (which rely on ascending columns) must then stay out of the way."""
    from vrad_b200.environment import Environment
    rng = np.random.default_rng(9)
    N = 4096
    env = Environment()
    env.add_triangles(np.int32([0]), np.float32([[0, 0, 0, 1, 0, 0, 0, 1, 0]])); env.setup_acceleration_structure()
    refl = rng.uniform(0.2, 0.8, (N, 3)).astype(np.float32)
    env.patches_upload(rng.uniform(-10, 10, (N, 3)).astype(np.float32), np.tile(np.float32([0, 0, 1]), (N, 1)), np.zeros(N, np.float32), np.ones(N, np.float32), refl)
    lens = rng.integers(0, 700, N)
    rp = np.zeros(N + 1, np.int64); np.cumsum(lens, out=rp[1:])
    col = rng.integers(0, N, rp[-1]).astype(np.int32)                  # any order, duplicates included
    w = rng.uniform(0.0, 1.0 / 700, rp[-1]).astype(np.float32)
    env.set_option("k4_short", 0)
    env.transfers_upload(0, N, rp, col, w)
    pairs, packed, segs, blocked, rows = env.transfers_layout()
    assert packed == 0 and blocked == 0
    emit0 = rng.uniform(0, 100, (N, 3)).astype(np.float32)
    import scipy.sparse as sp
    A = sp.csr_matrix((w.astype(np.float64), col, rp), shape=(N, N))    # duplicates are summed, as the gather does
    emit = emit0.astype(np.float64); total = np.zeros((N, 3))
    for _ in range(3):
        add = A @ (emit * refl.astype(np.float64)); total += add; emit = add
    for items in (0, 1):
        env.set_option("k4_items", items)
        t, _, _ = env.bounce(emit0, 3)
        assert np.abs(t - total).max() <= 1e-5 * np.abs(total).max(), items
    # the same rows sorted and de-duplicated may be packed again
    rows_sorted = [np.unique(col[rp[i]:rp[i + 1]]) for i in range(N)]
    rp2 = np.zeros(N + 1, np.int64); np.cumsum([len(r) for r in rows_sorted], out=rp2[1:])
    env.set_option("k4_items", 0)
    env.transfers_upload(0, N, rp2, np.concatenate(rows_sorted).astype(np.int32), np.full(rp2[-1], 1e-3, np.float32))
    pairs, packed, segs, blocked, rows = env.transfers_layout()
    assert packed > 0 or blocked > 0
    env.close()
