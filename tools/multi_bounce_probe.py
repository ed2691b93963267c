"""Development probe: per-bounce gather / exchange times at N ranks (run under torchrun with VRAD_TIMING=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from vrad_b200 import scenes
from vrad_b200.environment import Environment, environment_from_scene
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
s = scenes.multi_room(); env = environment_from_scene(s, device=lr, rank=rank, world=world)
if world > 1:
    uid = [Environment.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0); env.comm_init(uid[0])
nnz = env.build_transfers(s.pvs)
e0 = torch.full((s.n_patches, 3), 100.0, device="cuda"); out = torch.empty_like(e0)
for _ in range(3):
    env.bounce(e0, 40, out=out, want_added=False)
torch.cuda.synchronize()
env.close()
if world > 1: dist.destroy_process_group()
