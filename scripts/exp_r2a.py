"""Round-2 experiment batch A (one gpurun call): timings of the exact kernel shapes, the fused red-black kernel and
advection variants.  Knobs are environment variables read once per process, so every configuration is a subprocess."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, os
sys.path.insert(0, %r)
import numpy as np
import terminal_fluid_sim_b200 as tfs
DT = float(np.float32(1/60))
W, H, K, g, mode, steps, scene = json.loads(sys.argv[1])
s = tfs.FluidSim(W, H, tfs.SimConfig(gravity=g), iterations=K, mode=mode)
if scene == "random": s.scene_random(0x5EED, (2 << 64) // 100)
if scene == "lattice": s.scene_lattice(1024, 96)
for _ in range(3): s.next_step(DT)
s.sync(); s.set_profiling(True); s.timer_begin()
for _ in range(steps): s.next_step(DT)
ms = s.timer_end(); ph = s.phase_ms()
print(json.dumps({"ms_per_step": ms/steps, "project_ms": ph["project_ms"]/steps, "advect_ms": ph["advect_ms"]/steps, "kernel": s.last_projection(), "checksum": [hex(int(x)) for x in s.checksum()[:2]]}))
''' % ROOT

def run(tag, cfg, env=None):
    e = dict(os.environ); e.update(env or {})
    r = subprocess.run([sys.executable, "-c", CHILD, json.dumps(cfg)], env=e, capture_output=True, text=True, timeout=600)
    print(tag, cfg, env or {}, "->", r.stdout.strip() or r.stderr.strip()[-400:], flush=True)

def main():
    c3 = [4096, 4096, 100, 9.8, 0, 10, "none"]
    c5 = [32768, 32768, 200, 0.0, 0, 3, "random"]
    run("c3 exact default", c3)
    run("c3 exact (5,4)", c3, {"TFS_SYS_KG": "5", "TFS_SYS_NW": "4"})
    run("c3 exact (6,4)", c3, {"TFS_SYS_KG": "6", "TFS_SYS_NW": "4"})
    run("c3 exact (4,3)", c3, {"TFS_SYS_KG": "4", "TFS_SYS_NW": "3"})
    run("c3 rb fused", [4096, 4096, 100, 9.8, 1, 10, "none"])
    run("c3 rb fused lx256", [4096, 4096, 100, 9.8, 1, 10, "none"], {"TFS_RB_CHUNK": "256"})
    run("c3 rb fused lx64", [4096, 4096, 100, 9.8, 1, 10, "none"], {"TFS_RB_CHUNK": "64"})
    run("c3 rb two-launch", [4096, 4096, 100, 9.8, 1, 3, "none"], {"TFS_RB_VARIANT": "1"})
    for v in range(5):
        run("c3 advect variant %d" % v, [4096, 4096, 4, 9.8, 0, 20, "none"], {"TFS_ADV_VARIANT": str(v)})
    run("c4 exact", [16384, 16384, 50, 0.0, 0, 3, "lattice"])
    run("c4 rb", [16384, 16384, 50, 0.0, 1, 3, "lattice"])
    run("c5 exact default", c5)
    run("c5 exact (5,4)", c5, {"TFS_SYS_KG": "5", "TFS_SYS_NW": "4"})
    run("c5 rb", [32768, 32768, 200, 0.0, 1, 3, "random"])


if __name__ == "__main__":
    main()
