import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
from oracle.oracle import OracleSim
dt = float(np.float32(1/60))
# correctness on a developed, obstacle-laden flow
g = tfs.FluidSim(300, 200, tfs.SimConfig(gravity=3.0), iterations=10); o = OracleSim(300, 200, gravity=3.0, iterations=10)
g.scene_disc(60, 100, 20); o.scene_disc(60, 100, 20)
g.fill_random_velocity(5, 120.0); o.fill_random_velocity(5, 120.0)
for _ in range(3): g.next_step(dt); o.step(dt)
def par(W, H, amp, steps=3):
    g = tfs.FluidSim(W, H, tfs.SimConfig(gravity=3.0), iterations=6); o = OracleSim(W, H, gravity=3.0, iterations=6)
    from oracle.oracle import frac_to_threshold as ft; g.scene_random(7, ft(0.02)); o.scene_random(7, 0.02)
    g.fill_random_velocity(5, amp); o.fill_random_velocity(5, amp)
    for _ in range(steps): g.next_step(dt); o.step(dt)
    return all(np.array_equal(getattr(g,f).view(np.uint32), getattr(o,f).view(np.uint32)) for f in "uvps")
print("pair-kernel parity (H%256==0):", par(96, 256, 120.0), par(40, 512, 2000.0), par(300, 768, 30.0), par(64, 256, 0.5))
print("advect parity:", all(np.array_equal(getattr(g,f).view(np.uint32), getattr(o,f).view(np.uint32)) for f in "uvps"))
for N in (4096, 8192):
    g = tfs.FluidSim(N, N, tfs.SimConfig(gravity=9.8), iterations=20)
    for _ in range(3): g.next_step(dt)
    g.sync(); g.timer_begin()
    for _ in range(10): g.phase_advect(dt)
    ms = g.timer_end()/10
    print(f"{N}^2 advect {os.environ.get('TFS_ADVECT','fast32')}: {ms:.4f} ms = {25*N*N/ms/1e6:.0f} GB/s alg = {25*N*N/ms/1e6/6560.3:.3f} of roofline")
    g.close()
