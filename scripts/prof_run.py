"""One short run of a workload for ncu (python scripts/prof_run.py c3|c4|c5 exact|rb [steps])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import terminal_fluid_sim_b200 as tfs
from bench import WORKLOADS, setup_scene, DT, MODES
w = WORKLOADS[sys.argv[1]]
mode = sys.argv[2] if len(sys.argv) > 2 else "exact"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
s = tfs.FluidSim(w["W"], w["H"], tfs.SimConfig(gravity=w["gravity"]), iterations=w["K"], mode=MODES[mode])
setup_scene(s, w)
for _ in range(steps):
    s.next_step(DT)
s.sync()
print("done", s.last_projection())
