import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
dt = float(np.float32(1/60))
cases = [(1024, 40), (4096, 100)]
shapes = [(4,4),(2,8),(4,2),(2,4),(4,3),(2,5),(4,8),(1,8),(5,2)]
if len(sys.argv) > 1 and sys.argv[1] != '-':
    shapes = [tuple(int(x) for x in a.split('x')) for a in sys.argv[1].split(',')]
if len(sys.argv) > 2:
    cases = [(int(a.split('x')[0]), int(a.split('x')[1])) for a in sys.argv[2].split(',')]
for N, K in cases:
    for kg, nw in shapes:
        os.environ['TFS_SYS_KG']=str(kg); os.environ['TFS_SYS_NW']=str(nw)
        g = tfs.FluidSim(N, N, tfs.SimConfig(gravity=9.8), iterations=K, kernel=tfs.KERNEL_SYSTOLIC)
        for _ in range(2): g.next_step(dt)
        g.sync(); g.set_profiling(True)
        n = 3
        for _ in range(n): g.next_step(dt)
        ph = g.phase_ms()
        cells = N*N
        pj = ph['project_ms']/n; ad = ph['advect_ms']/n
        print(f"{N}^2 K={K} KG={kg} NW={nw} (KB={kg*nw}): project {pj:.3f} ms = {25*K*cells/pj/1e6/6560.3:.2f} of roofline; {K*cells/pj/1e6:.1f} Gupd/s; advect {ad:.3f} ms", flush=True)
        g.close()
