import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
dt = float(np.float32(1/60))
def run(W,H,K,kg,nw,label):
    os.environ['TFS_SYS_KG']=str(kg); os.environ['TFS_SYS_NW']=str(nw)
    g = tfs.FluidSim(W,H,tfs.SimConfig(gravity=9.8),iterations=K,kernel=tfs.KERNEL_SYSTOLIC)
    g.phase_project(dt); g.sync()
    g.timer_begin()
    for _ in range(3): g.phase_project(dt)
    ms = g.timer_end()/3
    KB=kg*nw; T=H+32+KB; nb=(W+KB+31)//32; npass=(K+KB-1)//KB
    print(f"{label}: W={W} H={H} K={K} KG={kg} NW={nw} bands={nb} passes={npass} T={T}: {ms:.3f} ms -> {ms*1e3/T*1965:.0f} cycles per step-of-one-task")
    g.close()
for kg,nw in [(4,4),(2,8),(4,2),(1,8)]:
    KB=kg*nw
    run(10,16384,KB,kg,nw,"1band-1pass")
    run(40,16384,KB,kg,nw,"2band-1pass")
    run(300,16384,KB,kg,nw,"10band-1pass")
    run(10,16384,2*KB,kg,nw,"1band-2pass")
    run(10,16384,4*KB,kg,nw,"1band-4pass")
