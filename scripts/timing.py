"""Timing tool (one GPU): per-step and per-phase device time of one or more configurations, each in its own process
so that the library's read-once tuning knobs can differ between them.

    python scripts/timing.py c3                      # BASELINE config by name (c1 .. c5), exact mode
    python scripts/timing.py c3:rb c4:rb c5          # several; ':rb' = red-black mode
    python scripts/timing.py 8192x8192:K=24:g=9.8    # an ad-hoc grid (scene none)
    python scripts/timing.py c3 TFS_SYS_KG=5 TFS_SYS_NW=4      # KEY=VALUE arguments become environment variables
    python scripts/timing.py chain                   # one band alone / one CTA per SM / full GPU: the systolic step chain

Every line of output is one JSON object: ms per step, projection / advection ms, the kernel that ran, a checksum."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import numpy as np
import terminal_fluid_sim_b200 as tfs
DT = float(np.float32(1/60))
W, H, K, g, mode, steps, scene = json.loads(sys.argv[1])
s = tfs.FluidSim(W, H, tfs.SimConfig(gravity=g), iterations=K, mode=mode)
if scene == "random": s.scene_random(0x5EED, (2 << 64) // 100)
if scene == "lattice": s.scene_lattice(1024, 96)
if scene == "disc": s.scene_disc(256, 512, 128)
for _ in range(3): s.next_step(DT)
s.sync(); s.set_profiling(True); s.timer_begin()
for _ in range(steps): s.next_step(DT)
ms = s.timer_end(); ph = s.phase_ms()
print(json.dumps({"ms_per_step": round(ms/steps, 4), "project_ms": round(ph["project_ms"]/steps, 4), "advect_ms": round(ph["advect_ms"]/steps, 4),
                  "gcell_steps_per_s": round(W*H*steps/ms/1e6, 4), "kernel": s.last_projection(), "checksum": ["%%016x" %% int(x) for x in s.checksum()[:2]]}))
''' % ROOT
NAMED = {"c1": (100, 100, 50, 0.0, "none", 50), "c2": (1024, 1024, 40, 0.0, "disc", 20), "c3": (4096, 4096, 100, 9.8, "none", 10),
         "c4": (16384, 16384, 50, 0.0, "lattice", 3), "c5": (32768, 32768, 200, 0.0, "random", 2)}


def run(tag, cfg, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, "-c", CHILD, json.dumps(cfg)], env=e, capture_output=True, text=True, timeout=900)
    res = json.loads(r.stdout) if r.returncode == 0 else {"error": r.stderr.strip()[-300:]}
    print(json.dumps({"config": tag, "env": env or {}, **res}), flush=True)


def parse(spec):
    parts = spec.split(":")
    mode = 1 if "rb" in parts[1:] else 0
    if parts[0] in NAMED:
        W, H, K, g, scene, steps = NAMED[parts[0]]
    else:
        W, H = (int(x) for x in parts[0].split("x"))
        K, g, scene, steps = 50, 0.0, "none", 5
    for p in parts[1:]:
        if p.startswith("K="):
            K = int(p[2:])
        if p.startswith("g="):
            g = float(p[2:])
        if p.startswith("steps="):
            steps = int(p[6:])
        if p.startswith("scene="):
            scene = p[6:]
    return [W, H, K, g, mode, steps, scene]


def is_env(a):
    return "=" in a and ":" not in a and a.split("=")[0].isupper()


if __name__ == "__main__":
    env = {a.split("=", 1)[0]: a.split("=", 1)[1] for a in sys.argv[1:] if is_env(a)}
    for spec in [a for a in sys.argv[1:] if not is_env(a)]:
        if spec == "chain":   # ~830 cycles per step for one band alone, ~1100 with 3 CTAs/SM
            for tag, cfg in (("1 band alone", [16, 262144, 16, 9.8, 0, 3, "none"]),
                             ("128 bands, 1 CTA/SM", [16 + 32 * 127, 32768, 16, 9.8, 0, 3, "none"]),
                             ("128 bands x 3 passes, 3 CTAs/SM", [16 + 32 * 127, 32768, 48, 9.8, 0, 3, "none"])):
                run(tag, cfg, env)
        else:
            run(spec, parse(spec), env)
