import sys, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
dt = float(np.float32(1/60))
g = tfs.FluidSim(4096, 4096, tfs.SimConfig(gravity=9.8), iterations=20)
for _ in range(3): g.next_step(dt)
g.sync(); g.timer_begin()
for _ in range(5): g.phase_advect(dt)
print("advect ms", g.timer_end()/5)
