import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
dt = float(np.float32(1/60))
for N, K in [(1024, 40), (4096, 100)]:
    for mode, kern, name in [(0, tfs.KERNEL_DIAGONAL, 'diag'), (1, 0, 'redblack')]:
        g = tfs.FluidSim(N, N, tfs.SimConfig(gravity=9.8), iterations=K, mode=mode, kernel=kern)
        for _ in range(2): g.next_step(dt)
        g.sync(); g.set_profiling(True)
        n = 3
        for _ in range(n): g.next_step(dt)
        ph = g.phase_ms()
        cells = N*N
        pj = ph['project_ms']/n; ad = ph['advect_ms']/n; gr = ph['gravity_ms']/n
        print(f"{N}^2 K={K} {name}: project {pj:.3f} ms ({25*K*cells/pj/1e6:.0f} GB/s alg), advect {ad:.3f} ms ({25*cells/ad/1e6:.0f} GB/s alg), gravity {gr:.3f} ms, launches {g.launch_count()}", flush=True)
        g.close()
