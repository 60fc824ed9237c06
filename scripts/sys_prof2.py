import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
N,K,kg,nw = [int(x) for x in sys.argv[1:5]]
os.environ['TFS_SYS_KG']=str(kg); os.environ['TFS_SYS_NW']=str(nw)
dt = float(np.float32(1/60))
g = tfs.FluidSim(N,N,tfs.SimConfig(gravity=9.8),iterations=K,kernel=tfs.KERNEL_SYSTOLIC)
g.next_step(dt); g.sync()
g.timer_begin(); g.phase_project(dt); print("ms", g.timer_end())
