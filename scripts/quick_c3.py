import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
from oracle.oracle import OracleSim, frac_to_threshold
dt = float(np.float32(1/60))
def same(a,b): return np.array_equal(a.view(np.uint32), b.view(np.uint32))
ok = True
for (W,H,K) in [(64,64,4),(100,100,50),(130,52,5),(40,300,9),(257,132,33),(512,512,20)]:
    g = tfs.FluidSim(W,H,tfs.SimConfig(gravity=2.5),iterations=K,kernel=tfs.KERNEL_SYSTOLIC); o = OracleSim(W,H,gravity=2.5,iterations=K)
    thr = frac_to_threshold(0.08)
    g.scene_random(W*131+H, thr); o.L.orc_scene_random(o.h, W*131+H, thr)
    g.fill_random_velocity(99, 40.0); o.fill_random_velocity(99, 40.0)
    for st in range(2):
        g.next_step(dt); o.step(dt)
    r = all(same(getattr(g,f), getattr(o,f)) for f in "uvps"); ok &= r
    print(W,H,K,"ok" if r else "MISMATCH", flush=True)
    g.close()
print("parity", "ALL OK" if ok else "FAILED")
for K in (100, 200):
    g = tfs.FluidSim(4096, 4096, tfs.SimConfig(gravity=9.8), iterations=K)
    for _ in range(3): g.next_step(dt)
    g.sync(); g.set_profiling(True)
    for _ in range(5): g.next_step(dt)
    pj = g.phase_ms()['project_ms']/5
    print(f"4096^2 K={K}: project {pj:.3f} ms = {25*K*4096*4096/pj/1e6/6560.3:.3f} of roofline", flush=True)
    g.close()
