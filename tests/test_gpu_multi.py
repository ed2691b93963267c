"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): one process per GPU, NCCL all-gather per
bounce inside libvradcuda, rows sharded by rank -- result must match the single-GPU run and the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["VRAD_ROOT"])
import numpy as np, torch, torch.distributed as dist
from vrad_b200 import scenes
from vrad_b200.environment import Environment, environment_from_scene
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
scene = scenes.multi_room(nx=3, ny=2)
env = environment_from_scene(scene, device=lr, rank=rank, world=world)
uid = [Environment.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
env.comm_init(uid[0])
nnz = env.build_transfers(scene.pvs)
row0, row1, _ = env.transfers_info()
N = scene.n_patches
emit0 = scenes.SplitMix64(7).uniform(3 * N, 0.0, 200.0).reshape(N, 3)
total, added, done = env.bounce(emit0, 6)
total_eo, added_eo, done_eo = env.bounce(emit0, 100, early_out=True)
rp, col, w = env.transfers_download()
np.savez(os.path.join(os.environ["VRAD_OUT"], f"rank{rank}.npz"), total=total, added=added, done=done, total_eo=total_eo,
         done_eo=done_eo, row0=row0, row1=row1, nnz=nnz, rp=rp, col=col, w=w)
env.close()
# the same map with its patch hierarchy (SubdividePatches trees): rows sharded, interior patches recomputed on every rank
hs = scenes.multi_room_hier(nx=3, ny=2)
t = hs.meta["tree"]
henv = environment_from_scene(hs, device=lr, rank=rank, world=world)
henv.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
uid2 = [Environment.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid2, src=0)
henv.comm_init(uid2[0])
hnnz = henv.build_transfers(hs.pvs)
hrow0, hrow1, _ = henv.transfers_info()
hN = hs.n_patches
hemit0 = scenes.SplitMix64(11).uniform(3 * hN, 0.0, 200.0).reshape(hN, 3)
htotal, hadded, hdone = henv.bounce(hemit0, 6)
hrp, hcol, hw = henv.transfers_download()
np.savez(os.path.join(os.environ["VRAD_OUT"], f"hier{rank}.npz"), total=htotal, added=hadded, row0=hrow0, row1=hrow1, nnz=hnnz,
         rp=hrp, col=hcol, w=hw)
for flag in (0, 1):          # all-gather pass per bounce, then the fused peer-store exchange of the leaf rows: same light
    henv.set_option("k4_hier_p2p", flag)
    ht2, _, _ = henv.bounce(hemit0, 6)
    np.save(os.path.join(os.environ["VRAD_OUT"], f"hier{rank}_p2p{flag}.npy"), ht2)
henv.close()
# bump totals with sharded rows
from vrad_b200.environment import bump_normals
bn = np.zeros((N, 3, 3), np.float32)
for n_ in np.unique(scene.patch_normal, axis=0):
    s_ = np.cross(n_, [0, 0, 1]) if abs(n_[2]) < 0.9 else np.float32([1, 0, 0])
    bn[np.all(scene.patch_normal == n_, axis=1)] = bump_normals(s_, np.cross(n_, s_), n_, n_)
bflags = (np.arange(N) % 3 != 0).astype(np.uint8)
benv = environment_from_scene(scene, device=lr, rank=rank, world=world)
uid3 = [Environment.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid3, src=0)
benv.comm_init(uid3[0])
benv.set_bump(bflags, bn)
benv.build_transfers(scene.pvs)
btotal, _, _ = benv.bounce(emit0, 4)
np.savez(os.path.join(os.environ["VRAD_OUT"], f"bump{rank}.npz"), total=btotal, bump=benv.bump_totals(), bn=bn, flags=bflags)
benv.close()
dist.destroy_process_group()
'''


def test_two_gpu_bounce_matches_single_and_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import pyoracle
    from vrad_b200 import scenes
    from vrad_b200.environment import row_partition
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, VRAD_ROOT=ROOT, VRAD_OUT=str(tmp_path))
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                    "--master-port", "29611", str(script)], check=True, env=env, timeout=600)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    scene = scenes.multi_room(nx=3, ny=2)
    N = scene.n_patches
    # the row blocks are contiguous, tile [0,N) in rank order and are balanced by (estimated) transfers, not rows
    parts = [(int(r0["row0"]), int(r0["row1"])), (int(r1["row0"]), int(r1["row1"]))]
    assert parts[0][0] == 0 and parts[0][1] == parts[1][0] and parts[1][1] == N
    assert abs(int(r0["nnz"]) - int(r1["nnz"])) < 0.05 * (int(r0["nnz"]) + int(r1["nnz"]))
    assert len(row_partition(N, 2)) == 2
    o = pyoracle.env_from_scene(scene)
    nnz = o.build_transfers(scene.pvs, threads=8)
    assert int(r0["nnz"]) + int(r1["nnz"]) == nnz
    rp, col, w = o.transfers()
    for r, (a, b) in zip((r0, r1), parts):           # each rank holds exactly its rows, bit-exact
        assert np.array_equal(r["rp"], rp[a:b + 1] - rp[a])
        assert np.array_equal(r["col"], col[rp[a]:rp[b]]) and np.array_equal(r["w"], w[rp[a]:rp[b]])
    emit0 = scenes.SplitMix64(7).uniform(3 * N, 0.0, 200.0).reshape(N, 3)
    to, ao, do = o.bounce(emit0, 6, threads=8)
    for r in (r0, r1):                                # every rank returns the full gathered result
        assert np.abs(r["total"] - to).max() <= 1e-4 * np.abs(to).max()
        assert np.allclose(r["added"], ao, rtol=1e-4) and int(r["done"]) == 6
    assert np.array_equal(r0["total"], r1["total"])
    te, ae, de = o.bounce(emit0, 100, early_out=True, threads=8)
    assert int(r0["done_eo"]) == int(r1["done_eo"]) == de
    assert np.abs(r0["total_eo"] - te).max() <= 1e-4 * np.abs(te).max()

    # hierarchical run of the same processes: rows bit-exact per rank, bounced light incl. the interior patches
    h0, h1 = np.load(tmp_path / "hier0.npz"), np.load(tmp_path / "hier1.npz")
    hs = scenes.multi_room_hier(nx=3, ny=2)
    t = hs.meta["tree"]
    ho = pyoracle.env_from_scene(hs)
    ho.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
    hnnz = ho.build_transfers(hs.pvs, threads=8)
    assert int(h0["nnz"]) + int(h1["nnz"]) == hnnz
    hrp, hcol, hw = ho.transfers()
    hparts = [(int(h0["row0"]), int(h0["row1"])), (int(h1["row0"]), int(h1["row1"]))]
    assert hparts[0][0] == 0 and hparts[0][1] == hparts[1][0] and hparts[1][1] == hs.n_patches
    for r, (a, b) in zip((h0, h1), hparts):
        assert np.array_equal(r["rp"], hrp[a:b + 1] - hrp[a])
        assert np.array_equal(r["col"], hcol[hrp[a]:hrp[b]]) and np.array_equal(r["w"], hw[hrp[a]:hrp[b]])
    hemit0 = scenes.SplitMix64(11).uniform(3 * hs.n_patches, 0.0, 200.0).reshape(hs.n_patches, 3)
    hto, hao, _ = ho.bounce(hemit0, 6, threads=8)
    for r in (h0, h1):
        assert np.abs(r["total"] - hto).max() <= 1e-4 * np.abs(hto).max()
        assert np.allclose(r["added"], hao, rtol=1e-4)
    assert np.array_equal(h0["total"], h1["total"])
    for r in (0, 1):
        for flag in (0, 1):
            t2 = np.load(tmp_path / f"hier{r}_p2p{flag}.npy")
            assert np.abs(t2 - hto).max() <= 1e-4 * np.abs(hto).max()
    # bump totals: every rank returns all rows' Light[1..3]
    b0, b1 = np.load(tmp_path / "bump0.npz"), np.load(tmp_path / "bump1.npz")
    ob = pyoracle.env_from_scene(scene)
    ob.set_bump(b0["flags"], b0["bn"])
    ob.build_transfers(scene.pvs, threads=8)
    tob, _, _ = ob.bounce(emit0, 4, threads=8)
    wb = ob.bump_totals()
    for r in (b0, b1):
        assert np.abs(r["total"] - tob).max() <= 1e-4 * np.abs(tob).max()
        assert np.abs(r["bump"] - wb).max() <= 1e-4 * np.abs(wb).max()
