"""torchrun: per-rank phase times of the slab run at config 5 (who waits for whom)."""
import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import torch, torch.distributed as dist
import terminal_fluid_sim_b200 as tfs
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dt = float(np.float32(1/60))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
g = tfs.FluidSim(N, N, tfs.SimConfig(), iterations=200, device=lr, rank=rank, world=world)
blobs = [None]*world
dist.all_gather_object(blobs, g.peer_export())
g.peer_attach(blobs[rank-1] if rank > 0 else None, blobs[rank+1] if rank < world-1 else None)
g.scene_random(0x5EED, (2 << 64)//100)
for _ in range(2): g.next_step(dt)
g.sync(); dist.barrier()
g.set_profiling(True)
g.timer_begin()
for _ in range(3): g.next_step(dt)
ms = g.timer_end()
ph = g.phase_ms()
out = [None]*world
dist.all_gather_object(out, (rank, ms/3, ph['gravity_ms']/3, ph['project_ms']/3, ph['advect_ms']/3))
if rank == 0:
    for o in out: print("rank %d: step %.1f ms | wait-before-project %.1f | project %.1f | advect+halo %.1f" % o)
dist.barrier(); g.close(); dist.destroy_process_group()
