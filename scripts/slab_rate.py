"""torchrun: steady-state and barrier-bracketed step times of the slab run (max over ranks)."""
import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import torch, torch.distributed as dist
import terminal_fluid_sim_b200 as tfs
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dt = float(np.float32(1/60))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
g = tfs.FluidSim(N, N, tfs.SimConfig(), iterations=K, device=lr, rank=rank, world=world)
blobs = [None]*world
dist.all_gather_object(blobs, g.peer_export())
g.peer_attach(blobs[rank-1] if rank > 0 else None, blobs[rank+1] if rank < world-1 else None)
g.scene_random(0x5EED, (2 << 64)//100)
for _ in range(2): g.next_step(dt)
g.sync(); dist.barrier()
g.timer_begin()
for _ in range(steps): g.next_step(dt)
ms = g.timer_end()
t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{lr}")
mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
mn = t.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world} {N}^2 K={K} {os.environ.get('TFS_SYS_GRID','')} {os.environ.get('TFS_SYS_G','')}: {steps} steps, rank 0 {ms/steps:.1f} ms/step, "
          f"max over ranks {mx.item()/steps:.1f} ms/step (min {mn.item()/steps:.1f})")
dist.barrier(); g.close(); dist.destroy_process_group()
