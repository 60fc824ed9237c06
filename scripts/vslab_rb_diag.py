"""Debug aid: red-black on two x-slabs of ONE device vs the oracle; prints which columns differ per K."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import terminal_fluid_sim_b200 as tfs
from oracle.oracle import OracleSim, frac_to_threshold
DT = float(np.float32(1 / 60))
W, H = 256, 64
for K in (4, 8, 12):
    for phase_only in (True, False):
        grp = tfs.SlabGroup(W, H, tfs.SimConfig(gravity=1.5), world=2, devices=[0, 0], iterations=K, mode=1)
        o = OracleSim(W, H, gravity=1.5, iterations=K, mode=1)
        thr = frac_to_threshold(0.05)
        grp.each("scene_random", 9, thr); o.L.orc_scene_random(o.h, 9, thr)
        grp.each("fill_random_velocity", 4, 50.0); o.fill_random_velocity(4, 50.0)
        grp.next_step(DT); o.step(DT)
        grp.sync(timeout_s=20)
        for f, n in ((tfs.FIELD_U, "u"), (tfs.FIELD_V, "v"), (tfs.FIELD_P, "p"), (tfs.FIELD_S, "s")):
            a = grp.read_field(f).reshape(W, H); b = np.array(getattr(o, n)).reshape(W, H)
            bad = np.nonzero((a.view(np.uint32) != b.view(np.uint32)).any(axis=1))[0]
            print("K", K, n, "bad columns:", (int(bad.min()), int(bad.max()), len(bad)) if len(bad) else "none", flush=True)
        print("slabs", [s.get_slab() for s in grp.sims])
        grp.close()
        break
