"""Profiling target: config 5 (32768^2, 2 % obstacles, K=200), one warm-up step + N steps (used under ncu)."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
dt = float(np.float32(1/60))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g = tfs.FluidSim(32768, 32768, tfs.SimConfig(), iterations=200)
g.scene_random(0x5EED, (2 << 64)//100)
g.next_step(dt)
g.sync(); g.timer_begin()
for _ in range(steps): g.next_step(dt)
print("ms/step", g.timer_end()/steps)
