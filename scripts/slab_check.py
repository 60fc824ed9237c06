"""2+-GPU check: x-slab handles (one per GPU, driven from one host thread) vs the CPU oracle."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
from oracle.oracle import OracleSim, frac_to_threshold
dt = float(np.float32(1/60))
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
def same(a,b): return np.array_equal(a.view(np.uint32), b.view(np.uint32))
ok_all = True
for (W,H,K,tb,g) in [(256,64,4,4,0.0),(256,128,20,8,2.5),(520,96,50,0,9.8),(1024,256,40,16,0.0),(1024,256,64,32,1.0)]:
    grp = tfs.SlabGroup(W,H,tfs.SimConfig(gravity=g),world=world,iterations=K,temporal_block=tb)
    o = OracleSim(W,H,gravity=g,iterations=K)
    thr = frac_to_threshold(0.05)
    grp.each("scene_random", 77, thr); o.L.orc_scene_random(o.h, 77, thr)
    grp.each("fill_random_velocity", 5, 60.0); o.fill_random_velocity(5, 60.0)
    print("slabs:", [s.get_slab() for s in grp.sims])
    assert np.array_equal(grp.read_field(tfs.FIELD_M), o.m)
    for st in range(3):
        grp.next_step(dt); o.step(dt)
        grp.sync()
        res = [same(grp.read_field(f), getattr(o, n)) for f, n in ((tfs.FIELD_U,'u'),(tfs.FIELD_V,'v'),(tfs.FIELD_P,'p'),(tfs.FIELD_S,'s'))]
        ok = all(res); ok_all &= ok
        print(f"{W}x{H} K={K} tb={tb} world={world} step {st}: {res}", flush=True)
        if not ok:
            for f, n in ((tfs.FIELD_U,'u'),(tfs.FIELD_V,'v'),(tfs.FIELD_P,'p'),(tfs.FIELD_S,'s')):
                a, b = grp.read_field(f), getattr(o, n)
                bad = np.nonzero(a.view(np.uint32) != b.view(np.uint32))[0]
                if bad.size: print("   ", n, bad.size, [(int(k)//H, int(k)%H, float(a[k]), float(b[k])) for k in bad[:6]])
            break
    if ok_all:   # several steps back to back without any host synchronisation: the ranks drift apart by their natural skew
        for st in range(5):
            grp.next_step(dt); o.step(dt)
        grp.sync()
        res = [same(grp.read_field(f), getattr(o, n)) for f, n in ((tfs.FIELD_U,'u'),(tfs.FIELD_V,'v'),(tfs.FIELD_P,'p'),(tfs.FIELD_S,'s'))]
        ok_all &= all(res)
        print(f"{W}x{H} K={K} tb={tb} world={world} 5 unsynchronised steps: {res}", flush=True)
    grp.close()
print("ALL OK" if ok_all else "FAILED")
