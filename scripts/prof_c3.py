"""Profiling target: config 3 (4096^2, K=100), a few steps (used under ncu)."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
dt = float(np.float32(1/60))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = tfs.FluidSim(4096, 4096, tfs.SimConfig(gravity=9.8), iterations=100)
for _ in range(3): g.next_step(dt)
g.sync(); g.timer_begin()
for _ in range(steps): g.next_step(dt)
print("ms/step", g.timer_end()/steps)
