"""Profiling target: ONE band task alone on the GPU (16 x 131072, K=16): the per-step latency chain of a CTA."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
dt = float(np.float32(1/60))
g = tfs.FluidSim(16, 131072, tfs.SimConfig(gravity=9.8), iterations=16)
g.next_step(dt); g.sync(); g.timer_begin()
g.next_step(dt)
print("ms/step", g.timer_end())
