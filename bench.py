#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native simulator step (driver contract).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle)

One "step" = FluidSim::next_step (simulator.rs:59-65): gravity + `iterations` projection sweeps +
advection over the whole grid.  Metric: Gcell-steps/s = W*H*steps / seconds / 1e9.

Workloads (BASELINE.json configs):
    c5  32768^2, random obstacles (2 %, splitmix64 seed 0x5EED), 200 pressure iterations   [default]
    c3  4096^2 smoke plume (gravity 9.8), 100 pressure iterations
    c2  1024^2 wind tunnel with a disc, 40 pressure iterations
The default is c5, the configuration BASELINE.json scales 1 -> 8 GPUs ("strong" scaling: the grid is
fixed, x-slabs are split over the ranks); c3 numbers are measured in the same run and reported under
"extra".  Inputs are larger than L2 (state 18 GB at c5, 285 MB at c3), so no L2 flush is needed.

The printed JSON line carries the contract keys plus `roofline` (projection kernel: algorithmic
25 B/cell/iteration over the CUDA-event duration, against the measured HBM copy bandwidth) and
`cpu_baseline` (the C restatement of the reference timed on this box's host cores, on a bounded
sample).  `e2e` is the same metric through the public API with host buffers: per step one brush
edit uploaded from pinned memory (H2D), next_step, the renderer's visible view of smoke and
pressure plus the pressure min/max read back into pinned memory (D2H).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = float(np.float32(1.0 / 60.0))
MODES = {"exact": 0, "rb": 1}
MODE_NAME = {"exact": "wavefront_exact", "rb": "red_black (temporally blocked, tolerance-checked)"}
WORKLOADS = {
    "c1": dict(name="config1: the reference's cargo bench, 100x100, default config, K=50, dt=1/60", W=100, H=100, K=50,
               gravity=0.0, scene="none"),
    "c4": dict(name="config4: 16384^2 wind tunnel, discs r=96 on a 1024-pitch lattice, K=50, dt=1/60", W=16384, H=16384, K=50,
               gravity=0.0, scene="lattice"),
    "c5": dict(name="config5: 32768^2 random obstacles 2% (seed 0x5EED), K=200, dt=1/60", W=32768, H=32768, K=200,
               gravity=0.0, scene="random"),
    "c3": dict(name="config3: 4096^2 smoke plume, gravity 9.8, K=100, dt=1/60", W=4096, H=4096, K=100, gravity=9.8,
               scene="none"),
    "c2": dict(name="config2: 1024^2 wind tunnel, disc r=128 at (256,512), K=40, dt=1/60", W=1024, H=1024, K=40,
               gravity=0.0, scene="disc"),
}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def setup_scene(sim, wl):
    if wl["scene"] == "random":
        sim.scene_random(0x5EED, (2 << 64) // 100)
    elif wl["scene"] == "disc":
        sim.scene_disc(256, 512, 128)
    elif wl["scene"] == "lattice":
        sim.scene_lattice(1024, 96)


def hexsum(cs):
    return ["%016x" % int(x) for x in cs]


def view_window(W, H):
    """the renderer's visible window (sim_renderer.rs:45-51): x in [0, w), y in [H - h, H)"""
    w, h = min(W, 512), min(H, 256)
    return 0, H - h, w, h


def time_steps(sim, steps):
    sim.sync()
    sim.timer_begin()
    for _ in range(steps):
        sim.next_step(DT)
    return sim.timer_end()


def run_ours(args):
    import terminal_fluid_sim_b200 as tfs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    wl = WORKLOADS[args.workload]
    W, H, K = wl["W"], wl["H"], wl["K"]
    peak, peak_src = measured_peaks()

    if world > 1:
        return run_ours_slabs(args, tfs, dist, rank, world, local_rank, wl, peak, peak_src)

    sim = tfs.FluidSim(W, H, tfs.SimConfig(gravity=wl["gravity"]), iterations=K, device=local_rank, mode=MODES[args.mode])
    setup_scene(sim, wl)
    for _ in range(args.warmup):
        sim.next_step(DT)
    sim.sync()

    # ---- timed region: K steps, device time by CUDA events on the library's stream ------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    sim.set_profiling(True)
    l_before = sim.launch_count()
    ms = time_steps(sim, args.steps)
    l_after = sim.launch_count()
    phases = sim.phase_ms()
    sim.set_profiling(False)
    clocks = sampler.stop()
    cells = W * H
    value = cells * args.steps / (ms / 1e3) / 1e9
    proj_ms = phases["project_ms"] / max(1, phases["steps"])
    adv_ms = phases["advect_ms"] / max(1, phases["steps"])
    alg_bytes = 25.0 * K * cells                                  # 25 B / cell / iteration
    achieved = alg_bytes / (proj_ms / 1e3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of the projection kernel, per launch, from the ncu captures
    # committed under profiles/ (a separate run of the same configuration; never measured inside this timed run)
    traffic = {("c5", "exact"): (4.699e11, "profiles/r02c_c5_exact_dram.csv"),
               ("c3", "exact"): (2.246e9, "profiles/r02c_systolic4_c3.txt"),
               ("c5", "rb"): (1.922e12, "profiles/r02b_c5_rb_dram.csv (28.7 GB per pass x 67 passes)"),
               ("c3", "rb"): (1.438e10, "profiles/r02b_rb_c3.txt (0.423 GB per pass x 34 passes)"),
               ("c4", "rb"): (1.231e11, "profiles/r02b_rb_c4.txt (7.24 GB per pass x 17 passes)")}.get((args.workload, args.mode), (None, None))
    kname = "k_project_systolic4 (gravity + %d-iteration exact projection)" if args.mode == "exact" else \
        "k_project_rb (gravity + %d red-black iterations, 3 per launch; all launches of a step)"
    roofline = {"kernel": kname % K, "bound": "hbm",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic[0], "traffic_source": traffic[1], "peak_source": peak_src, "ms_per_launch": round(proj_ms, 4),
                "algorithmic_bytes_per_launch": alg_bytes,
                "advect": {"ms_per_launch": round(adv_ms, 4), "achieved": round(25.0 * cells / (adv_ms / 1e3) / 1e9, 1),
                           "frac": round(25.0 * cells / (adv_ms / 1e3) / 1e9 / peak, 4)}}

    # ---- e2e: public API with host buffers -----------------------------------------------------------
    # the frame consumer is the renderer (src/ui/sim_renderer.rs): a terminal area of aw x ah cells shows the
    # aw x 2*ah sim cells at the top left; tfs_render_view returns their colours (min/max scan included)
    vx, vy, vw, vh = view_window(W, H)
    aw, ah = vw, vh // 2
    pin_c = tfs.PinnedBuffer(max(1, aw * ah * 8), np.uint8)
    pin_m = tfs.PinnedBuffer(H, np.uint8)
    pin_m.array[:] = 0
    e2e_steps = max(1, args.steps)

    def e2e_step(i):
        x = 8 + (i % 8)                                            # a brush column that moves a little
        pin_m.array[:] = 0
        pin_m.array[H // 2 - 4:H // 2 + 4] = 1
        sim.upload_blocks(pin_m.array, offset=x * H)               # H2D (dirty range) -> flags refresh
        sim.next_step(DT)
        return sim.render_view(aw, ah, out=pin_c.array[:aw * ah * 8])   # device colour loop, D2H 8 B per terminal cell

    e2e_step(0)
    sim.sync()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i + 1)
    sim.sync()
    e2e_s = time.perf_counter() - t0
    e2e = {"value": round(cells * e2e_steps / e2e_s / 1e9, 4), "unit": "Gcell-steps/s",
           "h2d_bytes_per_step": int(H), "d2h_bytes_per_step": int(aw * ah * 8 + 8),
           "what": "upload_blocks(1 column) + next_step + render_view(%dx%d terminal cells = %dx%d sim window: pressure "
                   "min/max, gradient, smoke, blocks on the device); host wall clock" % (aw, ah, vw, vh)}
    pin_c.close(); pin_m.close()
    launches_per_step = (l_after - l_before) / max(1, args.steps)
    # parity evidence: order-independent checksum of the bit patterns of u, v, p, s after the whole deterministic
    # sequence (warm-up, timed steps, e2e steps with their brush edits).  The same command at N = 2, 4, 8 must print
    # the same eight words: the x-slab run then equals the single-GPU run bit for bit.
    parity = {"checksum": hexsum(sim.checksum()), "steps": args.warmup + args.steps + 1 + e2e_steps,
              "what": "sum / xor of 64-bit hashes of (cell index, f32 bits) of u, v, p, s over the whole grid (tfs_checksum)"}
    proj_info = sim.last_projection()
    sim.close()

    # ---- extra: config 3 (the >= 60 %-of-roofline target of the north star) and config 4, both modes ----
    extra = {}
    if args.workload == "c5" and not args.no_extra:
        extra["config3_4096"] = side_workload(tfs, "c3", local_rank, peak, steps=10, warmup=3)
        extra["config4_16384"] = side_workload(tfs, "c4", local_rank, peak, steps=5, warmup=3)
        # the same frame loop through include/fluid_sim.hpp, the C++ mirror of `impl FluidSim` (set_block brush,
        # next_step, render_view): its next_step() copies nothing back, the frame's D2H is the visible window only
        extra["e2e_cpp_mirror"] = {k: cpp_mirror_frames(WORKLOADS[k], frames) for k, frames in (("c3", 10), ("c5", 3))}

    cpu = cpu_baseline(args.workload, steps=2, warmup=1) if not args.no_cpu else None

    out = {
        "metric": "Gcell-steps/s", "value": round(value, 4), "unit": "Gcell-steps/s", "n_gpus": n_gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "grid": [W, H], "pressure_iterations": K, "mode": MODE_NAME[args.mode],
                   "omega": 1.9, "parallelism": "x-slabs x%d" % n_gpus,
                   "l2": "state (%.1f GB) exceeds L2, no flush needed" % (17 * cells / 1e9)},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(round(launches_per_step * args.steps)),
        "gpu_launches_per_step": launches_per_step, "clocks": clocks, "parity": parity,
        "projection_kernel": proj_info, "extra": extra,
    }
    if rank == 0:
        print(json.dumps(out))


def cpp_mirror_frames(w, frames, mode=0):
    from terminal_fluid_sim_b200 import build as tfs_build
    try:
        r = subprocess.run([tfs_build.E2E, str(w["W"]), str(w["H"]), str(w["K"]), str(w["gravity"]), w["scene"], str(frames), str(mode)],
                           capture_output=True, text=True, timeout=600)
        return json.loads(r.stdout) if r.returncode == 0 else {"error": (r.stderr or r.stdout)[-300:]}
    except Exception as e:  # the headline numbers do not depend on this leg
        return {"error": repr(e)}


def side_workload(tfs, key, device, peak, steps, warmup, group=None):
    """one of the other BASELINE configs measured in the same run, in both projection modes; `group` = (dist, rank,
    world) runs it on x-slabs.  Returns ms/step, Gcell-steps/s, per-phase roofline fractions and the checksum."""
    import numpy as _np
    w = WORKLOADS[key]
    cells = w["W"] * w["H"]
    out = {"workload": w["name"]}
    for mode in ("exact", "rb"):
        if group is None:
            s = tfs.FluidSim(w["W"], w["H"], tfs.SimConfig(gravity=w["gravity"]), iterations=w["K"], mode=MODES[mode], device=device)
        else:
            dist, rank, world = group
            s = tfs.FluidSim(w["W"], w["H"], tfs.SimConfig(gravity=w["gravity"]), iterations=w["K"], mode=MODES[mode], device=device,
                             rank=rank, world=world)
            blobs = [None] * world
            dist.all_gather_object(blobs, s.peer_export())
            s.peer_attach(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < world - 1 else None)
        setup_scene(s, w)
        s.sync()
        if group is not None:
            group[0].barrier()
        for _ in range(warmup):
            s.next_step(DT)
        s.sync()
        if group is not None:
            group[0].barrier()
        s.set_profiling(True)
        ms = time_steps(s, steps)
        ph = s.phase_ms()
        s.set_profiling(False)
        cs = s.checksum()
        pj, ad = ph["project_ms"] / max(1, ph["steps"]), ph["advect_ms"] / max(1, ph["steps"])
        if group is not None:
            import torch
            dist, rank, world = group
            t = torch.tensor([ms, pj, ad], dtype=torch.float64, device=f"cuda:{device}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, pj, ad = (float(x) for x in t.tolist())
            parts = [None] * world
            dist.all_gather_object(parts, [int(x) for x in cs])
            cs = tfs.combine_checksums([_np.array(p, dtype=_np.uint64) for p in parts])
            dist.barrier()
        n = 1 if group is None else group[2]
        r = {"ms_per_step": round(ms / steps, 4), "value": round(cells * steps / (ms / 1e3) / 1e9, 4), "unit": "Gcell-steps/s",
             "project": {"ms": round(pj, 4), "achieved_gbs_per_gpu": round(25.0 * w["K"] * cells / (pj / 1e3) / 1e9 / n, 1),
                         "frac": round(25.0 * w["K"] * cells / (pj / 1e3) / 1e9 / n / peak, 4)},
             "advect": {"ms": round(ad, 4), "achieved_gbs_per_gpu": round(25.0 * cells / (ad / 1e3) / 1e9 / n, 1),
                        "frac": round(25.0 * cells / (ad / 1e3) / 1e9 / n / peak, 4)},
             "kernel": s.last_projection(), "checksum_after_%d_steps" % (warmup + steps): hexsum(cs)}
        out[mode] = r
        s.sync()
        if group is not None:
            group[0].barrier()
        s.close()
    return out


def run_ours_slabs(args, tfs, dist, rank, world, local_rank, wl, peak, peak_src):
    """N > 1: one x-slab per rank / GPU.  Neighbours exchange halo columns, band records and progress
    counters by direct peer stores over NVLink (CUDA IPC mappings); torch.distributed is used only to
    hand the IPC blobs around, for the barrier and for the max-over-ranks of the device time."""
    import torch
    W, H, K = wl["W"], wl["H"], wl["K"]
    sim = tfs.FluidSim(W, H, tfs.SimConfig(gravity=wl["gravity"]), iterations=K, device=local_rank, rank=rank, world=world,
                       mode=MODES[args.mode])
    blobs = [None] * world
    dist.all_gather_object(blobs, sim.peer_export())
    sim.peer_attach(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < world - 1 else None)
    setup_scene(sim, wl)
    sim.sync()
    dist.barrier()
    for _ in range(args.warmup):
        sim.next_step(DT)
    sim.sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    sim.set_profiling(True)
    l_before = sim.launch_count()
    ms = time_steps(sim, args.steps)
    l_after = sim.launch_count()
    phases = sim.phase_ms()
    sim.set_profiling(False)
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([ms, phases["project_ms"] / max(1, phases["steps"]), phases["advect_ms"] / max(1, phases["steps"])],
                     dtype=torch.float64, device=f"cuda:{local_rank}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, proj_ms, adv_ms = (float(x) for x in t.tolist())
    clocks = sampler.stop() if rank == 0 else None
    cells = W * H
    value = cells * args.steps / (ms / 1e3) / 1e9
    alg_bytes = 25.0 * K * cells
    achieved = alg_bytes / (proj_ms / 1e3) / 1e9          # whole job, all GPUs
    roofline = {"kernel": "k_project_systolic4 (gravity + %d-iteration exact projection), %d x-slabs" % (K, world),
                "bound": "hbm", "achieved": round(achieved / world, 1), "peak": peak, "unit": "GB/s (per GPU)",
                "frac": round(achieved / world / peak, 4), "traffic": None, "peak_source": peak_src,
                "ms_per_launch": round(proj_ms, 4), "algorithmic_bytes_per_launch": alg_bytes / world,
                "advect": {"ms_per_launch": round(adv_ms, 4)}}
    # e2e: every rank reads its share of the visible window (rank 0 owns the left columns the renderer shows)
    c_lo, c_hi, _, _ = sim.get_slab()
    vx, vy, vw, vh = view_window(W, H)
    vw_r = max(0, min(vx + vw, c_hi) - max(vx, c_lo))
    pin_s, pin_p = tfs.PinnedBuffer(max(1, vw_r * vh)), tfs.PinnedBuffer(max(1, vw_r * vh))
    pin_m = tfs.PinnedBuffer(H, np.uint8)

    def e2e_step(i):
        x = 8 + (i % 8)
        pin_m.array[:] = 0
        pin_m.array[H // 2 - 4:H // 2 + 4] = 1
        sim.upload_blocks(pin_m.array, offset=x * H)
        sim.next_step(DT)
        if vw_r > 0:
            sim.read_view(tfs.FIELD_S, max(vx, c_lo), vy, vw_r, vh, out=pin_s.array)
            sim.read_view(tfs.FIELD_P, max(vx, c_lo), vy, vw_r, vh, out=pin_p.array)
        return sim.pressure_minmax()

    e2e_step(0)
    sim.sync()
    dist.barrier()
    t0 = time.perf_counter()
    for i in range(max(1, args.steps)):
        e2e_step(i + 1)
    sim.sync()
    dist.barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e = {"value": round(cells * max(1, args.steps) / e2e_s / 1e9, 4), "unit": "Gcell-steps/s",
           "h2d_bytes_per_step": int(H), "d2h_bytes_per_step": int(2 * vw * vh * 4 + 8 * world),
           "what": "per rank: upload_blocks(1 column) + next_step + read_view(own share of the %dx%d smoke/pressure window) "
                   "+ pressure_minmax; host wall clock, max over ranks" % (vw, vh)}
    launches_per_step = (l_after - l_before) / max(1, args.steps)
    pin_s.close(); pin_p.close(); pin_m.close()
    sim.sync()
    parts = [None] * world
    dist.all_gather_object(parts, [int(x) for x in sim.checksum()])
    parity = {"checksum": hexsum(tfs.combine_checksums([np.array(p, dtype=np.uint64) for p in parts])),
              "steps": args.warmup + args.steps + 1 + max(1, args.steps),
              "what": "sum / xor of 64-bit hashes of (cell index, f32 bits) of u, v, p, s over the whole grid: the per-slab "
                      "checksums (tfs_checksum) added / xored; equal to the N=1 line's when the x-slab run is bit-identical"}
    proj_info = sim.last_projection()
    dist.barrier()
    sim.close()
    extra = {}
    if args.workload == "c5" and not args.no_extra:
        extra["config4_16384"] = side_workload(tfs, "c4", local_rank, peak, steps=5, warmup=3, group=(dist, rank, world))
    out = {
        "metric": "Gcell-steps/s", "value": round(value, 4), "unit": "Gcell-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "grid": [W, H], "pressure_iterations": K, "mode": MODE_NAME[args.mode],
                   "omega": 1.9, "parallelism": "x-slabs x%d (peer stores over NVLink, no collective)" % world,
                   "l2": "state (%.1f GB) exceeds L2, no flush needed" % (17 * cells / 1e9)},
        "roofline": roofline, "cpu_baseline": None, "e2e": e2e,
        "gpu_launches": int(round(launches_per_step * args.steps)) * world, "gpu_launches_per_step_per_rank": launches_per_step,
        "clocks": clocks, "parity": parity, "projection_kernel": proj_info, "extra": extra,
    }
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


def cpu_sample(workload):
    """bounded sample of the workload for the CPU arm: same scene statistics on a smaller grid"""
    wl = WORKLOADS[workload]
    if workload == "c5":
        return dict(W=1024, H=1024, K=wl["K"], gravity=wl["gravity"], scene="random",
                    desc="1024x1024 sub-grid of config5 (same obstacle density and seed), K=200")
    if workload == "c3":
        return dict(W=1024, H=1024, K=wl["K"], gravity=wl["gravity"], scene="none",
                    desc="1024x1024 sub-grid of config3, K=100")
    if workload == "c4":
        return dict(W=2048, H=2048, K=wl["K"], gravity=wl["gravity"], scene="lattice",
                    desc="2048x2048 sub-grid of config4 (same lattice), K=50")
    if workload == "c1":
        return dict(W=100, H=100, K=wl["K"], gravity=0.0, scene="none", desc="full config1 (100x100, K=50)")
    return dict(W=wl["W"], H=wl["H"], K=wl["K"], gravity=wl["gravity"], scene="disc", desc="full config2")


def run_cpu(workload, steps, warmup):
    from oracle.oracle import OracleSim, frac_to_threshold
    sm = cpu_sample(workload)
    o = OracleSim(sm["W"], sm["H"], gravity=sm["gravity"], iterations=sm["K"])
    if sm["scene"] == "random":
        o.L.orc_scene_random(o.h, 0x5EED, frac_to_threshold(0.02))
    elif sm["scene"] == "disc":
        o.scene_disc(256, 512, 128)
    elif sm["scene"] == "lattice":
        o.scene_lattice(1024, 96)
    for _ in range(warmup):
        o.step(DT)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(DT)
    el = time.perf_counter() - t0
    cells = sm["W"] * sm["H"]
    return cells * steps / el / 1e9, el / steps * 1e3, sm


def cpu_baseline(workload, steps, warmup):
    value, ms, sm = run_cpu(workload, steps, warmup)
    return {"value": round(value, 6), "unit": "Gcell-steps/s", "cores": os.cpu_count(), "kind": "port",
            "sample": sm["desc"] + "; %d steps after %d warm-up; projection is serial as in the reference "
                                   "(simulator.rs:157-189), gravity/advection use all cores (OpenMP for rayon)" % (steps, warmup),
            "ms_per_step": round(ms, 2)}


def run_reference(args):
    """The reference's own CPU algorithm (C restatement: the Rust crate cannot be built here) on the
    box's host cores, same metric/config, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    value, ms, sm = run_cpu(args.workload, max(1, args.steps), max(0, args.warmup))
    line = {
        "impl": "reference", "metric": "Gcell-steps/s", "value": round(value, 6), "unit": "Gcell-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "grid": [wl["W"], wl["H"]], "pressure_iterations": wl["K"],
                   "mode": "reference order (lexicographic Gauss-Seidel)", "sample": sm["desc"]},
        "cpu_baseline": {"value": round(value, 6), "unit": "Gcell-steps/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": sm["desc"]},
        "e2e": {"value": round(value, 6), "unit": "Gcell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_sweep(args):
    """--workload all: the deterministic bench harness.  Every BASELINE config on one GPU, both
    projection modes, next to the CPU restatement of the reference - on the SAME grid for configs 1 and 2, on a
    stated sub-grid for the larger ones.  One JSON line; BASELINE.md's results table is filled from it."""
    import terminal_fluid_sim_b200 as tfs
    peak, peak_src = measured_peaks()
    rows = []
    plan = {"c1": (200, 20, 200, 20), "c2": (20, 5, 3, 1), "c3": (10, 3, 2, 1), "c4": (5, 3, 2, 1), "c5": (3, 3, 2, 1)}
    for key in ("c1", "c2", "c3", "c4", "c5"):
        gs, gw, cs_, cw = plan[key]
        row = side_workload(tfs, key, 0, peak, steps=gs, warmup=gw)
        row["key"] = key
        if not args.no_cpu:
            value, ms, sm = run_cpu(key, cs_, cw)
            row["cpu"] = {"value": round(value, 6), "unit": "Gcell-steps/s", "ms_per_step": round(ms, 3), "cores": os.cpu_count(),
                          "kind": "port", "sample": sm["desc"], "same_grid": sm["W"] == WORKLOADS[key]["W"]}
        rows.append(row)
    print(json.dumps({"sweep": rows, "n_gpus": 1, "dtype": "f32", "data": "synthetic", "peak_hbm_gbs": peak, "peak_source": peak_src,
                      "unit": "Gcell-steps/s", "dt": DT}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS) + ["all"])
    ap.add_argument("--mode", default="exact", choices=sorted(MODES), help="projection order of the headline run")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the config-3 extra measurements")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.workload == "all":
        run_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
