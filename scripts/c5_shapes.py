import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
dt = float(np.float32(1/60))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
for shp in sys.argv[1].split(','):
    kg, nw = shp.split('x')
    os.environ['TFS_SYS_KG']=kg; os.environ['TFS_SYS_NW']=nw
    g = tfs.FluidSim(N, N, tfs.SimConfig(), iterations=200)
    g.scene_random(0x5EED, (2 << 64)//100)
    g.next_step(dt); g.sync(); g.set_profiling(True)
    for _ in range(2): g.next_step(dt)
    ph = g.phase_ms(); pj = ph['project_ms']/2
    print(f"{N}^2 K=200 2% obstacles KG={kg} NW={nw}: project {pj:.1f} ms = {25*200*N*N/pj/1e6/6560.3:.3f} of roofline, {200*N*N/pj/1e6:.0f} Gupd/s", flush=True)
    g.close()
