"""Mint tests/golden/*.npz from the independent NumPy restatement (oracle/np_oracle.py).

The reference (Rust) cannot be built or run in this environment and ships no golden vectors
of its own, so these fixtures are NOT outputs of the reference:
they are outputs of np_oracle.NpSim, which was written separately from oracle/fluid_oracle.c.
tests/test_oracle.py requires the C oracle to reproduce every array here bit-for-bit.

Run from the repo root:  python tests/golden/make_golden.py
Each case stores, per recorded step, u/v/p/s as raw float32 plus a sha256 over
(u|v|p|s) bytes for EVERY step, the mask, and the case parameters.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.np_oracle import NpSim, render_view  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
DT = float(np.float32(1.0 / 60.0))


def digest(sim) -> str:
    h = hashlib.sha256()
    for a in (sim.u, sim.v, sim.p, sim.s):
        h.update(np.ascontiguousarray(a, np.float32).tobytes())
    return h.hexdigest()


def disc(sim, cx, cy, r):
    for i in range(sim.W):
        for j in range(sim.H):
            dx, dy = 2 * i + 1 - 2 * cx, 2 * j + 1 - 2 * cy
            if dx * dx + dy * dy < 4 * r * r:
                sim.set_block(i, j)


def random_blocks(sim, seed, count):
    rng = np.random.default_rng(seed)
    for _ in range(count):
        sim.set_block(int(rng.integers(0, sim.W)), int(rng.integers(0, sim.H)))


def run_case(name, sim, steps, keep, params, dt=DT):
    out = {"mask": sim.m.copy(), "u0": sim.u.copy(), "v0": sim.v.copy(), "s0": sim.s.copy()}
    hashes = []
    for st in range(1, steps + 1):
        sim.step(dt)
        hashes.append(digest(sim))
        if st in keep:
            for f in "uvps":
                out[f"{f}_step{st}"] = getattr(sim, f).copy()
    params = dict(params, W=sim.W, H=sim.H, steps=steps, keep=sorted(keep), dt=dt, hashes=hashes)
    out["params"] = np.frombuffer(json.dumps(params).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "done", hashes[-1][:16])


def main():
    # config 1: the reference's own cargo-bench case (benches/benchmarks/sim_bench.rs:4-12), dt fixed
    p = dict(gravity=0.0, wind_speed=50.0, smoke_size=0.25, density=1000.0, iterations=50, omega=1.9)
    run_case("bench_100x100", NpSim(100, 100, **p), 40, {1, 10, 40}, p)

    # terminal-derived sizes (app.rs:51-55, layout.rs:41-43) with obstacles
    s = NpSim(80, 48, **p)
    disc(s, 24, 24, 7)
    run_case("term_80x48_disc", s, 20, {1, 20}, dict(p, scene="disc(24,24,7)"))

    p2 = dict(p, gravity=-9.8)
    s = NpSim(48, 42, **p2)
    random_blocks(s, 7, 160)
    run_case("term_48x42_random_gravity", s, 20, {1, 20}, dict(p2, scene="random_blocks(seed=7,count=160)"))

    # adversarial velocities: long backtraces exercise every clamp of sample_vector
    p3 = dict(p, iterations=10, gravity=3.0)
    s = NpSim(64, 40, **p3)
    random_blocks(s, 11, 100)
    rng = np.random.default_rng(5)
    s.u[:] = rng.uniform(-300, 300, s.u.size).astype(np.float32)
    s.v[:] = rng.uniform(-300, 300, s.v.size).astype(np.float32)
    s.s[:] = rng.uniform(0, 1, s.s.size).astype(np.float32)
    run_case("adversarial_64x40", s, 3, {1, 3}, dict(p3, scene="random_blocks(seed=11,count=100)+random fields"))

    # odd sizes (H not a multiple of 4, tiny grids, degenerate 2x2 default)
    p4 = dict(p, iterations=17, gravity=1.5)
    s = NpSim(37, 23, **p4)
    random_blocks(s, 3, 60)
    run_case("odd_37x23", s, 6, {1, 6}, dict(p4, scene="random_blocks(seed=3,count=60)"))
    run_case("default_2x2", NpSim(2, 2, **p), 2, {2}, p)
    run_case("thin_3x64", NpSim(3, 64, **p), 3, {3}, p)

    # edit path: resize quirk (simulator.rs:121-135), set_config (:67-71), restart (:55-57)
    s = NpSim(20, 10, **p)
    random_blocks(s, 2, 30)
    s.step(DT)
    s.resize(16, 12)
    m_after_resize = s.m.copy()
    s.step(DT)
    s.set_config(0.5, 20.0, 0.5, 900.0)
    u_after_cfg, s_after_cfg = s.u.copy(), s.s.copy()
    s.step(DT)
    fields_after = {f: getattr(s, f).copy() for f in "uvps"}
    s.restart_sim()
    np.savez_compressed(
        os.path.join(HERE, "edit_path.npz"),
        m_after_resize=m_after_resize, u_after_cfg=u_after_cfg, s_after_cfg=s_after_cfg,
        m_after_restart=s.m.copy(), u_after_restart=s.u.copy(), s_after_restart=s.s.copy(),
        **{f + "_after": a for f, a in fields_after.items()})
    print("edit_path done")
    render_goldens()


def render_goldens():
    """Terminal cells (src/ui/sim_renderer.rs:23-151) of the last stored step of three cases, from
    np_oracle.render_view; plus a synthetic field that reaches every branch of the colour map."""
    out = {}
    for name, aw, ah in (("term_80x48_disc", 80, 24), ("term_48x42_random_gravity", 48, 21), ("odd_37x23", 30, 11)):
        z = np.load(os.path.join(HERE, name + ".npz"))
        prm = json.loads(bytes(z["params"]).decode())
        st = max(prm["keep"])
        cells, mn, mx = render_view(z[f"p_step{st}"], z[f"s_step{st}"], z["mask"], prm["W"], prm["H"], aw, ah)
        out[name] = cells
        out[name + "_minmax"] = np.array([mn, mx], np.float32)
        out[name + "_area"] = np.array([aw, ah])
    W, H = 40, 36
    rng = np.random.default_rng(9)
    p = rng.uniform(-2000, 3000, W * H).astype(np.float32)
    p[:2] = 0
    s = rng.uniform(-0.2, 1.2, W * H).astype(np.float32)      # outside [0, 1]: the `as u8` saturation
    s[rng.random(W * H) < 0.5] = 0
    m = (rng.random(W * H) < 0.1).astype(np.uint8)
    p[5], s[6] = np.float32("nan"), np.float32("nan")
    cells, mn, mx = render_view(p, s, m, W, H, W, H // 2)
    out.update(synthetic=cells, synthetic_p=p, synthetic_s=s, synthetic_m=m, synthetic_minmax=np.array([mn, mx], np.float32))
    np.savez_compressed(os.path.join(HERE, "render_view.npz"), **out)
    print("render_view done")


if __name__ == "__main__":
    render_goldens() if "--render-only" in sys.argv else main()
