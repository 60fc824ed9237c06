import sys, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
N = int(sys.argv[1]); K = int(sys.argv[2]); tb = int(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dt = float(np.float32(1/60))
g = tfs.FluidSim(N, N, tfs.SimConfig(gravity=9.8), iterations=K, kernel=tfs.KERNEL_SYSTOLIC, temporal_block=tb)
g.next_step(dt); g.sync(); g.set_profiling(True)
for _ in range(reps): g.next_step(dt)
ph = g.phase_ms()
print(f"{N}^2 K={K} tb={tb}: project {ph['project_ms']/reps:.3f} ms, {K*N*N/(ph['project_ms']/reps)/1e6:.1f} Gupd/s")
