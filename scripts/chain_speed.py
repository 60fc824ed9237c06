"""How fast does one band (CTA) step when it runs alone, in a chain with one CTA per SM, and in a full GPU?"""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
import terminal_fluid_sim_b200 as tfs
dt = float(np.float32(1/60))
def run(W, H, K, note):
    g = tfs.FluidSim(W, H, tfs.SimConfig(gravity=9.8), iterations=K)
    for _ in range(2): g.next_step(dt)
    g.sync(); g.set_profiling(True)
    for _ in range(3): g.next_step(dt)
    pj = g.phase_ms()['project_ms'] / 3
    KB = 16
    bands = (W + KB + 31) // 32; passes = (K + KB - 1) // KB
    T = H + 32 + KB
    lag = 48
    crit = T + bands * lag      # steps on the critical path of one pass (all bands resident)
    print(f"{note}: {W}x{H} K={K}: {bands} bands x {passes} passes, project {pj:.3f} ms; if one wave: {pj*1e6/crit:.0f} ns per step "
          f"({pj*1e6/crit*1.9:.0f} cycles at 1.9 GHz)", flush=True)
    g.close()
run(16, 262144, 16, "1 band alone")
run(48, 262144, 16, "2 bands")
run(16 + 32 * 15, 65536, 16, "16 bands (1 CTA on 16 SMs)")
run(16 + 32 * 127, 32768, 16, "128 bands (1 CTA/SM)")
run(16 + 32 * 127, 32768, 32, "128 bands x 2 passes (2 CTAs/SM)")
run(16 + 32 * 127, 32768, 48, "128 bands x 3 passes (3 CTAs/SM)")
run(16 + 32 * 127, 32768, 96, "128 bands x 6 passes (2 rounds of 3 CTAs/SM)")
