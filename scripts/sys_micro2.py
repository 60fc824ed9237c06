import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
dt = float(np.float32(1/60))
def run(W,H,K,kg,nw,S):
    os.environ['TFS_SYS_KG']=str(kg); os.environ['TFS_SYS_NW']=str(nw); os.environ['TFS_SYS_S']=str(S)
    g = tfs.FluidSim(W,H,tfs.SimConfig(gravity=9.8),iterations=K,kernel=tfs.KERNEL_SYSTOLIC)
    g.phase_project(dt); g.sync()
    g.timer_begin()
    for _ in range(3): g.phase_project(dt)
    ms = g.timer_end()/3
    KB=kg*nw; T=H+32+KB
    print(f"W={W} H={H} K={K} KG={kg} NW={nw} S={S}: {ms:.3f} ms -> {ms*1e3/T*1965:.0f} cycles/step")
    g.close()
for S in (16,32,64):
    run(10,16384,16,4,4,S)
for S in (16,64):
    run(10,16384,8,1,8,S)
