import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
dt = float(np.float32(1/60))
cases = [(1024, 40), (4096, 100)]
variants = [(0, tfs.KERNEL_SYSTOLIC, 4, 'sys4'), (0, tfs.KERNEL_SYSTOLIC, 8, 'sys8'), (0, tfs.KERNEL_SYSTOLIC, 10, 'sys10'), (0, tfs.KERNEL_SYSTOLIC, 16, 'sys16'), (1, 0, 0, 'redblack')]
if len(sys.argv) > 1:
    variants = [v for v in variants if v[3] in sys.argv[1].split(',')]
if len(sys.argv) > 2:
    cases = [(int(a.split('x')[0]), int(a.split('x')[1])) for a in sys.argv[2].split(',')]
for N, K in cases:
    for mode, kern, tb, name in variants:
        g = tfs.FluidSim(N, N, tfs.SimConfig(gravity=9.8), iterations=K, mode=mode, kernel=kern, temporal_block=tb)
        for _ in range(2): g.next_step(dt)
        g.sync(); g.set_profiling(True)
        n = 3
        for _ in range(n): g.next_step(dt)
        ph = g.phase_ms()
        cells = N*N
        pj = ph['project_ms']/n; ad = ph['advect_ms']/n; gr = ph['gravity_ms']/n
        print(f"{N}^2 K={K} {name}: project {pj:.3f} ms ({25*K*cells/pj/1e6:.0f} GB/s alg = {25*K*cells/pj/1e6/6560.3:.2f} of roofline; {K*cells/pj/1e6:.1f} Gupd/s), advect {ad:.3f} ms ({25*cells/ad/1e6:.0f} GB/s alg), gravity {gr:.3f} ms", flush=True)
        g.close()
