"""Profiling target for ncu: a few steps of one workload, nothing else.
    python scripts/prof_run.py c3|c4|c5|c2|c1 [exact|rb] [steps]
    python scripts/prof_run.py oneband          # ONE band task alone on the GPU (16 x 131072, K = 16): the systolic step chain
Typical use (see profiles/README.md):
    ncu --set full --clock-control none --import-source on -k regex:k_project_rb --launch-skip 30 --launch-count 1 \
        -o gpurun_out/x python scripts/prof_run.py c3 rb 2"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import terminal_fluid_sim_b200 as tfs  # noqa: E402
from bench import DT, MODES, WORKLOADS, setup_scene  # noqa: E402

name = sys.argv[1]
mode = sys.argv[2] if len(sys.argv) > 2 else "exact"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
w = dict(W=16, H=131072, K=16, gravity=9.8, scene="none") if name == "oneband" else WORKLOADS[name]
s = tfs.FluidSim(w["W"], w["H"], tfs.SimConfig(gravity=w["gravity"]), iterations=w["K"], mode=MODES[mode])
setup_scene(s, w)
for _ in range(steps):
    s.next_step(DT)
s.sync()
print("done", s.last_projection())
