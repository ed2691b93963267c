"""World-size-2 gloo test (CPU): the row-sharded bounce recurrence with one all-gather per iteration
reproduces the unsharded result.  Compute per rank is the oracle's row gather; the exchange and the
partition rule are the product's host logic (vrad_b200/sharding.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vrad_b200 import sharding


def _make_problem(N=301, seed=1):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 40, N)
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    col = rng.integers(0, N, rowptr[-1]).astype(np.int32)
    w = (rng.random(rowptr[-1]) * 0.05).astype(np.float32)
    refl = (rng.random((N, 3)) * 0.7).astype(np.float32)
    emit0 = (rng.random((N, 3)) * 100).astype(np.float32)
    return rowptr, col, w, refl, emit0


def _worker(rank, world, port, n_bounces, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    rowptr, col, w, refl, emit0 = _make_problem()
    N = emit0.shape[0]
    row0, row1 = sharding.row_partition(N, world)[rank]
    rpr = sharding.rows_per_rank(N, world)
    lrp, lcol, lw = sharding.slice_csr(rowptr, col, w, row0, row1)
    emit = sharding.padded_gather_buffer(N, world)
    emit[:N] = emit0
    refl_pad = sharding.padded_gather_buffer(N, world); refl_pad[:N] = refl
    total = np.zeros((row1 - row0, 3), np.float32)
    for _ in range(n_bounces):
        add = pyoracle.gather_rows(0, row1 - row0, lrp, lcol, lw, emit, refl_pad)     # local rows only
        total += add
        mine = torch.zeros((rpr, 3), dtype=torch.float32); mine[: row1 - row0] = torch.from_numpy(add)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)                                                # the one collective per bounce
        emit = torch.cat(parts).numpy()
    tot = torch.zeros((rpr, 3), dtype=torch.float32); tot[: row1 - row0] = torch.from_numpy(total)
    parts = [torch.empty_like(tot) for _ in range(world)]
    dist.all_gather(parts, tot)
    if rank == 0:
        q.put(torch.cat(parts).numpy()[:N])
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_bounce_equals_unsharded(world):
    from oracle import pyoracle
    n_bounces = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_bounces, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rowptr, col, w, refl, emit0 = _make_problem()
    N = emit0.shape[0]
    emit = emit0.copy(); total = np.zeros_like(emit0)
    for _ in range(n_bounces):
        add = pyoracle.gather_rows(0, N, rowptr, col, w, emit, refl)
        total += add; emit = add
    assert np.array_equal(got, total)          # same per-row summation order -> bit-identical


def test_partition_helpers():
    assert sharding.range_partition(10, 3) == [(0, 3), (3, 6), (6, 10)]
    rp = np.array([0, 2, 2, 5, 9]); col = np.arange(9); w = np.arange(9, dtype=np.float32)
    lrp, lcol, lw = sharding.slice_csr(rp, col, w, 1, 3)
    assert list(lrp) == [0, 0, 3] and list(lcol) == [2, 3, 4]
    assert sharding.padded_gather_buffer(10, 4).shape == (16, 3)          # 4 ranks x 4 rows: blocks start on multiples of 4


# ---- patch hierarchy: rows sharded, interior patches recomputed on every rank after the exchange ----

def _hier_problem():
    from oracle import pyoracle
    from vrad_b200 import scenes
    sc = scenes.multi_room_hier(nx=2, ny=1, boxes_per_room=6)
    t = sc.meta["tree"]
    o = pyoracle.env_from_scene(sc)
    o.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
    o.build_transfers(sc.pvs, threads=4)
    rowptr, col, w = o.transfers()
    N = sc.n_patches
    emit0 = scenes.SplitMix64(5).uniform(3 * N, 0.0, 200.0).reshape(N, 3)
    return sc, t, o, rowptr, col, w, emit0


def _hier_worker(rank, world, port, n_bounces, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    sc, t, _, rowptr, col, w, emit0 = _hier_problem()
    N = emit0.shape[0]
    row0, row1 = sharding.row_partition(N, world)[rank]
    rpr = sharding.rows_per_rank(N, world)
    lrp, lcol, lw = sharding.slice_csr(rowptr, col, w, row0, row1)
    ids, ptr, leaf, wt = sharding.collect_rows(t["parent"], t["child1"], t["child2"], sc.patch_area)
    emit = sharding.padded_gather_buffer(N, world); emit[:N] = emit0
    refl_pad = sharding.padded_gather_buffer(N, world); refl_pad[:N] = sc.patch_refl
    total = np.zeros((row1 - row0, 3), np.float32)
    for _ in range(n_bounces):
        add = pyoracle.gather_rows(0, row1 - row0, lrp, lcol, lw, emit, refl_pad)     # interior rows are empty -> 0
        total += add
        mine = torch.zeros((rpr, 3), dtype=torch.float32); mine[: row1 - row0] = torch.from_numpy(add)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        emit = torch.cat(parts).numpy()
        sharding.apply_collect(emit, ids, ptr, leaf, wt)                            # CollectLight, interior patches, on every rank
    tot = torch.zeros((rpr, 3), dtype=torch.float32); tot[: row1 - row0] = torch.from_numpy(total)
    parts = [torch.empty_like(tot) for _ in range(world)]
    dist.all_gather(parts, tot)
    if rank == 0:
        q.put(sharding.apply_collect(torch.cat(parts).numpy()[:N].copy(), ids, ptr, leaf, wt))
    dist.destroy_process_group()


def test_sharded_hierarchical_bounce_equals_oracle():
    n_bounces = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_hier_worker, args=(r, 2, port, n_bounces, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc, t, o, rowptr, col, w, emit0 = _hier_problem()
    want, _, _ = o.bounce(emit0, n_bounces, threads=4)          # the oracle's literal reverse-order CollectLight
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    interior = t["child1"] != -1
    assert (want[interior] > 0).mean() > 0.9 and np.all(np.diff(rowptr)[interior] == 0)


def test_collect_rows_weights():
    # a root with children (area 3, area 1); the first child again split 2:1
    parent = [-1, 0, 0, 1, 1]; child1 = [1, 3, -1, -1, -1]; child2 = [2, 4, -1, -1, -1]
    area = np.array([4, 3, 1, 2, 1], np.float32)
    ids, ptr, leaf, wt = sharding.collect_rows(parent, child1, child2, area)
    assert list(ids) == [0, 1] and list(ptr) == [0, 3, 5]
    assert list(leaf[:3]) == [3, 4, 2] and np.allclose(wt[:3], [0.75 * 2 / 3, 0.75 / 3, 0.25])
    v = np.zeros((5, 3), np.float32); v[2:] = [[8, 8, 8], [2, 2, 2], [5, 5, 5]]
    sharding.apply_collect(v, ids, ptr, leaf, wt)
    assert np.allclose(v[1], 3.0) and np.allclose(v[0], 0.75 * 3 + 0.25 * 8)


# ---- the file-driven bake with luxels and patch origins sharded over ranks (vrad_b200/bake.py) -----------------------------------
def _bake_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    from vrad_b200 import bake, bspfile
    L, meta = bspfile.synthetic_map(2, 1, boxes_per_room=3, sky_rooms=(1,))
    prep = bake.prepare(L, meta["entities"])
    lit = bake.light(pyoracle.OracleEnv(), prep, bounces=3, rank=rank, world=world)     # K3 ranges per rank, all-gathered over gloo
    # the K5 leg's exchange (packed colours as int32 blocks of uneven length), with the host pack function standing in for the kernel
    parts = sharding.range_partition(lit["direct"].shape[0], world)
    a, b = parts[rank]
    ind = np.where(prep["lux_patch"][a:b, None] >= 0, lit["total"][np.maximum(prep["lux_patch"][a:b], 0)], np.float32(0)).astype(np.float32)
    mine = bspfile.color_to_rgbexp32(lit["direct"][a:b] + ind)
    colors = bake.all_gather_blocks(mine.view(np.int32), parts, rank).view(bspfile.RGBEXP32)
    if rank == 0:
        q.put((lit["direct"], lit["emit0"], lit["total"], colors))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_bake_equals_single_process(world):
    from oracle import pyoracle
    from vrad_b200 import bake, bspfile
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_bake_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    direct, emit0, total, colors = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    L, meta = bspfile.synthetic_map(2, 1, boxes_per_room=3, sky_rooms=(1,))
    prep = bake.prepare(L, meta["entities"])
    ref = bake.light(pyoracle.OracleEnv(), prep, bounces=3)
    # luxels and patch origins are independent work items: sharding them changes nothing, bit for bit
    assert np.array_equal(direct, ref["direct"]) and np.array_equal(emit0, ref["emit0"]) and np.array_equal(total, ref["total"])
    ind = np.where(prep["lux_patch"][:, None] >= 0, ref["total"][np.maximum(prep["lux_patch"], 0)], np.float32(0)).astype(np.float32)
    assert np.array_equal(colors, bspfile.color_to_rgbexp32(ref["direct"] + ind))
    assert direct.shape[0] % world != 0 or world == 2                         # uneven blocks are part of the case (world 3)
