"""Debug aid: two x-slab handles on ONE device.  python scripts/vslab_debug.py exact|rb [reverse]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import terminal_fluid_sim_b200 as tfs
DT = float(np.float32(1 / 60))
mode = 1 if sys.argv[1] == "rb" else 0
rev = len(sys.argv) > 2 and sys.argv[2] == "reverse"
grp = tfs.SlabGroup(256, 64, tfs.SimConfig(gravity=1.5), world=2, devices=[0, 0], iterations=12, mode=mode)
for st in range(3):
    sims = grp.sims[::-1] if rev else grp.sims
    for s in sims:
        s.next_step(DT)
    t0 = time.time()
    while not all(s.idle() for s in grp.sims):
        if time.time() - t0 > 8:
            print("step", st, "STUCK: idle =", [s.idle() for s in grp.sims], flush=True)
            os._exit(3)
        time.sleep(0.01)
    print("step", st, "ok", [hex(int(x)) for x in grp.checksum()[:2]], flush=True)
print("done")
