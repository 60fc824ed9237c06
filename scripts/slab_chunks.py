"""torchrun experiment: block-cyclic band assignment through VIRTUAL ranks.  Each process (GPU) holds C slab handles,
virtual rank vr = c * world + rank, so the pass-0 wave of the exact sweep visits every GPU after world * (bands/chunk)
bands instead of after a whole slab.  Nothing in the library changes: a virtual rank is an ordinary x-slab handle.
  mode "check":  2048 x 512, K = 40 against the CPU oracle (bit-exact)               -- verified on 2 GPUs for C = 1, 2, 4
  mode "rate" :  N^2, K: barrier-bracketed ms/step, max over ranks
STATUS: experiment.  With C handles per GPU there are C persistent projection kernels per GPU and step; each must
leave CTA slots for the others, or chunk c's kernel fills the GPU with CTAs that wait (through the other GPUs) for
chunk c + world, whose kernel cannot become resident: that is a deadlock, observed at 16384^2 with C = 4 before the
slot budget below existed.  The budget (TFS_SYS_GRID = slots / C) has not been measured yet."""
import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import torch, torch.distributed as dist
import terminal_fluid_sim_b200 as tfs
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dt = float(np.float32(1/60))
mode = sys.argv[1]
C = int(sys.argv[2])
VW = C * world
if C > 1: os.environ.setdefault("TFS_SYS_GRID", str(max(1, (3 * 148) // C)))   # every chunk's kernel stays co-resident
def make(W, H, K, cfg):
    sims = [tfs.FluidSim(W, H, cfg, iterations=K, device=lr, rank=c * world + rank, world=VW) for c in range(C)]
    mine = [s.peer_export() for s in sims]
    allb = [None] * world
    dist.all_gather_object(allb, mine)
    blob = lambda vr: allb[vr % world][vr // world] if 0 <= vr < VW else None
    for c, s in enumerate(sims):
        vr = c * world + rank
        s.peer_attach(blob(vr - 1), blob(vr + 1))
    return sims
if mode == "check":
    from oracle.oracle import OracleSim, frac_to_threshold
    W, H, K = 2048, 512, 40
    sims = make(W, H, K, tfs.SimConfig(gravity=1.5))
    o = OracleSim(W, H, gravity=1.5, iterations=K)
    thr = frac_to_threshold(0.03)
    for s in sims: s.scene_random(9, thr); s.fill_random_velocity(3, 80.0)
    o.L.orc_scene_random(o.h, 9, thr); o.fill_random_velocity(3, 80.0)
    ok = True
    for st in range(4):
        for s in sims: s.next_step(dt)
        o.step(dt)
    for s in sims: s.sync()
    for s in sims:
        c_lo, c_hi, _, _ = s.get_slab()
        for f, n in ((tfs.FIELD_U,'u'),(tfs.FIELD_V,'v'),(tfs.FIELD_P,'p'),(tfs.FIELD_S,'s')):
            ok &= bool(np.array_equal(s.read_field(f).view(np.uint32), getattr(o, n)[c_lo*H:c_hi*H].view(np.uint32)))
    t = torch.tensor([1 if ok else 0], device=f"cuda:{lr}")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0: print(f"slab_chunks check world={world} C={C}: 4 unsynchronised steps:", "ALL OK" if int(t.item()) else "FAILED")
else:
    N, K, steps = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    sims = make(N, N, K, tfs.SimConfig())
    for s in sims: s.scene_random(0x5EED, (2 << 64)//100)
    for _ in range(2):
        for s in sims: s.next_step(dt)
    for s in sims: s.sync()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    t0 = time.perf_counter()
    for _ in range(steps):
        for s in sims: s.next_step(dt)
    for s in sims: s.sync()
    ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{lr}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"slab_chunks rate world={world} C={C} {N}^2 K={K}: {steps} steps, {t.item()/steps:.1f} ms/step (host clock, max over ranks)")
dist.barrier()
for s in sims: s.close()
dist.destroy_process_group()
