"""torchrun check: one process per GPU, IPC peer mappings; every rank compares its slab with the oracle."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import terminal_fluid_sim_b200 as tfs
from oracle.oracle import OracleSim, frac_to_threshold
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dt = float(np.float32(1/60))
W, H, K = 2048, 512, 40
g = tfs.FluidSim(W, H, tfs.SimConfig(gravity=1.5), iterations=K, device=lr, rank=rank, world=world)
blobs = [None]*world
dist.all_gather_object(blobs, g.peer_export())
g.peer_attach(blobs[rank-1] if rank > 0 else None, blobs[rank+1] if rank < world-1 else None)
o = OracleSim(W, H, gravity=1.5, iterations=K)
thr = frac_to_threshold(0.03)
g.scene_random(9, thr); o.L.orc_scene_random(o.h, 9, thr)
g.fill_random_velocity(3, 80.0); o.fill_random_velocity(3, 80.0)
c_lo, c_hi, _, _ = g.get_slab()
ok = True
for st in range(3):
    g.next_step(dt); o.step(dt); g.sync()
    for f, n in ((tfs.FIELD_U,'u'),(tfs.FIELD_V,'v'),(tfs.FIELD_P,'p'),(tfs.FIELD_S,'s')):
        a = g.read_field(f); b = getattr(o, n)[c_lo*H:c_hi*H]
        ok &= bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
t = torch.tensor([1 if ok else 0], device=f"cuda:{lr}")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("slab_check_mp world=%d:" % world, "ALL OK" if int(t.item()) else "FAILED")
dist.barrier(); g.close(); dist.destroy_process_group()
