"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck): exits non-zero on a parity failure."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import terminal_fluid_sim_b200 as tfs
from oracle.oracle import OracleSim, frac_to_threshold
dt = float(np.float32(1/60))
ok = True
def same(g, o, fields="uvps"):
    return all(np.array_equal(getattr(g, f).view(np.uint32), getattr(o, f).view(np.uint32)) for f in fields)
cases = [(96, 256, 20, tfs.KERNEL_AUTO, 0), (130, 52, 9, tfs.KERNEL_AUTO, 0), (37, 23, 17, tfs.KERNEL_AUTO, 0),
         (64, 64, 6, tfs.KERNEL_DIAGONAL, 0), (64, 48, 8, tfs.KERNEL_AUTO, 1), (48, 512, 12, tfs.KERNEL_SYSTOLIC, 0)]
for (W, H, K, kern, mode) in cases:
    g = tfs.FluidSim(W, H, tfs.SimConfig(gravity=2.0), iterations=K, kernel=kern, mode=mode)
    o = OracleSim(W, H, gravity=2.0, iterations=K, mode=mode)
    g.scene_random(3, frac_to_threshold(0.05)); o.scene_random(3, 0.05)
    g.fill_random_velocity(4, 300.0); o.fill_random_velocity(4, 300.0)
    for _ in range(2):
        g.next_step(dt); o.step(dt)
    r = same(g, o)
    a, mn, mx = g.render_view(min(W, 40), min(H // 2, 20)); b, mn2, mx2 = o.render_view(min(W, 40), min(H // 2, 20))
    r &= bool(np.array_equal(a, b)) and (mn, mx) == (mn2, mx2)
    win = g.read_view(tfs.FIELD_S, 1, 2, min(W - 1, 9), min(H - 2, 11))
    r &= bool(np.array_equal(win.view(np.uint32), o.s.reshape(W, H)[1:1 + min(W - 1, 9), 2:2 + min(H - 2, 11)].ravel().view(np.uint32)))
    g.pressure_minmax(); g.divergence_linf()
    print(W, H, K, kern, mode, "ok" if r else "MISMATCH", flush=True)
    ok &= r
    g.close()
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
