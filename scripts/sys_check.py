"""Quick on-GPU check of the systolic kernel against the oracle, over the instantiated shapes."""
import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
from oracle.oracle import OracleSim, frac_to_threshold
dt = float(np.float32(1/60))
def same(a,b): return np.array_equal(a.view(np.uint32), b.view(np.uint32))
ok_all = True
shapes = [(4,4),(2,8),(4,2),(2,4),(4,3),(2,2),(5,2),(2,5),(4,8),(1,8)]
cases = [(64,64,4),(100,100,50),(130,50,5),(40,300,9),(257,129,33),(300,40,50),(512,512,20)]
if len(sys.argv) > 1: cases = cases[:int(sys.argv[1])]
for (kg,nw) in shapes:
    os.environ['TFS_SYS_KG']=str(kg); os.environ['TFS_SYS_NW']=str(nw)
    for (W,H,K) in cases:
        g = tfs.FluidSim(W,H,tfs.SimConfig(gravity=2.5),iterations=K,kernel=tfs.KERNEL_SYSTOLIC)
        o = OracleSim(W,H,gravity=2.5,iterations=K)
        thr = frac_to_threshold(0.08)
        g.scene_random(W*131+H, thr); o.L.orc_scene_random(o.h, W*131+H, thr)
        g.fill_random_velocity(99, 40.0); o.fill_random_velocity(99, 40.0)
        res = []
        for st in range(2):
            t0=time.time(); g.next_step(dt); g.sync(); el=time.time()-t0
            o.step(dt)
            res.append(same(g.u,o.u) and same(g.v,o.v) and same(g.p,o.p) and same(g.s,o.s))
        ok = all(res); ok_all &= ok
        print(f"KG={kg} NW={nw} {W}x{H} K={K}: {'ok' if ok else 'MISMATCH'} {el*1e3:.2f} ms", flush=True)
        if not ok:
            for name in 'uvp':
                a,b = getattr(g,name), getattr(o,name)
                bad = np.nonzero(a.view(np.uint32)!=b.view(np.uint32))[0]
                print("  ", name, "mismatches:", bad.size, [(int(k)//H, int(k)%H, float(a[k]), float(b[k])) for k in bad[:5]])
            break
print("ALL OK" if ok_all else "FAILED")
