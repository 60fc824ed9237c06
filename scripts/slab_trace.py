"""torchrun, with a library built with TFS_NVCC_EXTRA=-DTFS_TASK_TRACE: per-task start/end times of the
projection kernel of the LAST step on every rank (who runs when, how long tasks take, how busy the slots are)."""
import os, sys, struct, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import terminal_fluid_sim_b200 as tfs
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dt = float(np.float32(1/60))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
g = tfs.FluidSim(N, N, tfs.SimConfig(), iterations=K, device=lr, rank=rank, world=world)
blobs = [None]*world
dist.all_gather_object(blobs, g.peer_export())
g.peer_attach(blobs[rank-1] if rank > 0 else None, blobs[rank+1] if rank < world-1 else None)
g.scene_random(0x5EED, (2 << 64)//100)
for _ in range(2): g.next_step(dt)
g.sync(); dist.barrier()
g.timer_begin()
for _ in range(3): g.next_step(dt)
ms = g.timer_end() / 3
g.sync(); dist.barrier()
os.makedirs("gpurun_out", exist_ok=True)
os.environ["TFS_TASK_TRACE_PREFIX"] = "gpurun_out/ttrace"
g.next_step(dt); g.sync()
dist.barrier()
c_lo, c_hi, _, _ = g.get_slab()
import glob
lines = []
for path in sorted(glob.glob("gpurun_out/ttrace.b*.bin")) if rank == 0 else []:
    raw = open(path, "rb").read()
    hdr = struct.unpack("8q", raw[:64]); n, n_passes, nbl, b0, grid, T, KB = hdr[:7]
    tasks = np.frombuffer(raw[64:64 + 8 * n], np.int32).reshape(n, 2)
    tt = np.frombuffer(raw[64 + 8 * n:], np.uint64).reshape(n, 4).astype(np.float64)
    st, en = tt[:, 0] * 1e-6, tt[:, 1] * 1e-6                       # ms
    t0 = st.min(); st -= t0; en -= t0
    dur = en - st
    span = en.max()
    lines.append(f"b0={b0:5d} bands={nbl} passes={n_passes} tasks={n} grid={grid}: span {span:7.1f} ms, task ms min/med/max "
                 f"{dur.min():.1f}/{np.median(dur):.1f}/{dur.max():.1f}, slot-busy {dur.sum() / (grid * span):.2f}, "
                 f"ideal(work/slots) {n * dur.min() / grid:.1f} ms")
    for p in range(n_passes):
        m = tasks[:, 0] == p
        lines.append(f"      pass {p:2d}: start {st[m].min():7.1f}..{st[m].max():7.1f}  end {en[m].min():7.1f}..{en[m].max():7.1f}  dur med {np.median(dur[m]):.1f} max {dur[m].max():.1f}")
if rank == 0:
    print(f"world {world}, {N}^2 K={K}: {ms:.1f} ms/step")
    print("\n".join(lines))
dist.barrier(); g.close(); dist.destroy_process_group()
