"""Development probe (run under torchrun): us per bounce of the C4 gather at N ranks for the multi-GPU kernel's switches,
device-timed with CUDA events, max over ranks.  Prints one JSON line on rank 0 and writes gpurun_out/r02_multi_probe_n<N>.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from vrad_b200 import scenes
from vrad_b200.environment import Environment, environment_from_scene
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1: dist.init_process_group("nccl", device_id=dev)
s = scenes.multi_room(); env = environment_from_scene(s, device=lr, rank=rank, world=world)
env.set_stream(torch.cuda.current_stream().cuda_stream)
if world > 1:
    uid = [Environment.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0); env.comm_init(uid[0])
nnz = env.build_transfers(s.pvs)
row0, row1, _ = env.transfers_info()
N = s.n_patches
e0 = torch.from_numpy(scenes.SplitMix64(0xE1).uniform(3 * N, 0.0, 200.0).reshape(N, 3)).to(dev); out = torch.empty_like(e0)
env.set_async(True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {"world": world, "nnz_local_rank0": nnz, "rows_rank0": [row0, row1]}
ref = None
# (block, persist, pool, pdl, graph, pack)
for blk, per, pool, pdl, graph, pack in ((192, 1, 25, 1, 1, 1), (192, 1, 25, 1, 1, 0), (192, 1, 12, 1, 1, 1), (192, 1, 12, 1, 1, 0), (192, 1, 40, 1, 1, 1), (192, 1, 25, 0, 1, 1),
                                         (192, 1, 25, 1, 0, 1), (192, 1, 0, 1, 1, 1), (256, 1, 25, 1, 1, 1), (192, 0, 0, 1, 1, 1)):
    for k, v in (("k4_block", blk), ("k4_persist", per), ("k4_pool", pool), ("k4_pdl", pdl), ("k4_graph", graph), ("k4_pack", pack)):
        env.set_option(k, v)
    env.bounce(e0, 100, out=out, want_added=False)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(3):
        env.bounce(e0, 100, out=out, want_added=False)
    ev1.record(); torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1) / 300 * 1e3], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    got = out.cpu().numpy()
    if ref is None: ref = got
    res[f"block{blk}_persist{per}_pool{pool}_pdl{pdl}_graph{graph}_pack{pack}"] = {"us_per_bounce": float(t.item()), "max_rel_vs_first": float(np.abs(got - ref).max() / np.abs(ref).max())}
    if rank == 0: print(blk, per, pool, pdl, graph, pack, float(t.item()), float(np.abs(got - ref).max() / np.abs(ref).max()), flush=True)
env.close()
# the same map with its patch hierarchy: leaf rows by peer stores (k4_hier_p2p=1) against the all-gather pass per bounce (0)
hs = scenes.multi_room_hier(nx=12, ny=11); tr = hs.meta["tree"]
henv = environment_from_scene(hs, device=lr, rank=rank, world=world)
henv.set_hierarchy(tr["parent"], tr["child1"], tr["child2"], tr["face"])
henv.set_stream(torch.cuda.current_stream().cuda_stream)
if world > 1:
    uid = [Environment.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0); henv.comm_init(uid[0])
hnnz = henv.build_transfers(hs.pvs)
hN = hs.n_patches
he = torch.full((hN, 3), 100.0, device=dev); ho = torch.empty_like(he)
henv.set_async(True)
for flag in (1, 0):
    henv.set_option("k4_hier_p2p", flag)
    henv.bounce(he, 50, out=ho, want_added=False)
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); ev0.record()
    for _ in range(3):
        henv.bounce(he, 100, out=ho, want_added=False)
    ev1.record(); torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1) / 300 * 1e3], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[f"hierarchy_p2p{flag}_us_per_bounce"] = float(t.item())
    if rank == 0: print("hierarchy", flag, float(t.item()), flush=True)
henv.close()
if rank == 0:
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/r02_multi_probe_n{world}.json", "w"), indent=1)
if world > 1: dist.destroy_process_group()
