"""GPU parity tests added in round 2 (run with -m gpu on a B200), all through the C ABI:

* the exact mode against the CPU oracle ON the benchmark configurations (config 3 steps 1-3; a 2048-column slice of
  config 5), and the shipped block-ring kernel against the independent one-launch-per-diagonal kernel at 8192^2;
* KERNEL_AUTO runs the block-ring kernel for every iteration count (no silent change of kernel family);
* stale-workspace regressions (resize to a smaller grid, change of the iteration count);
* the temporally blocked red-black kernel: bit-identical to the oracle's red-black statement and to the simple
  one-launch-per-colour kernel; against the reference's order within the MEASURED tolerance (scripts/rb_tolerance.py);
* the order-independent checksum bench.py prints, against the same formula on the host;
* two x-slabs on ONE device (virtual ranks), so that the multi-GPU protocol is exercised on a 1-GPU box too.
"""
import numpy as np
import pytest

import terminal_fluid_sim_b200 as tfs
from oracle.oracle import OracleSim, frac_to_threshold
from tests.helpers import DT, bits_equal, first_mismatch

pytestmark = pytest.mark.gpu


def assert_same(g, o, what, fields="uvps"):
    H = o.height if hasattr(o, "height") else o.get_size()[1]
    for f in fields:
        a, b = getattr(g, f), getattr(o, f)
        assert bits_equal(a, b), f"{what}: {f}: " + first_mismatch(a, b, H)


# ---- exact mode on the benchmark configurations ---------------------------------------------------------------
def test_config3_4096_K100_steps_1_to_3_bit_exact_vs_oracle():
    """BASELINE config 3: 4096^2 smoke plume, gravity 9.8, K = 100; oracle comparison on steps 1-3."""
    W = H = 4096
    g = tfs.FluidSim(W, H, tfs.SimConfig(gravity=9.8), iterations=100)
    o = OracleSim(W, H, gravity=9.8, iterations=100)
    for st in range(1, 4):
        g.next_step(DT); o.step(DT)
        assert g.last_projection()["kernel"] == tfs.PROJ_SYSTOLIC_BLOCK_RING
        assert_same(g, o, f"config3 step {st}")
    assert np.array_equal(g.checksum(), tfs.checksum_of_fields(o.u, o.v, o.p, o.s))


def _wait(sim, seconds, what):
    """bounded wait: a protocol error must fail the test, not hang the box"""
    import time
    t0 = time.time()
    while not sim.idle():
        assert time.time() - t0 < seconds, f"{what}: the device did not finish within {seconds} s"
        time.sleep(0.01)


def test_config5_slice_512x32768_K200_bit_exact_vs_oracle():
    """A slice of BASELINE config 5 at its FULL height: 32768 rows, random obstacles 2 % with the config's seed (the mask
    depends on idx = x*H + y only, so these ARE config 5's first columns), K = 200, one step, against the CPU oracle.
    512 columns keep the serial CPU sweep at about a minute; the 2048-column slice is checked on the device below."""
    W, H, K = 512, 32768, 200
    thr = frac_to_threshold(0.02)
    g = tfs.FluidSim(W, H, iterations=K)
    o = OracleSim(W, H, iterations=K)
    g.scene_random(0x5EED, thr); o.L.orc_scene_random(o.h, 0x5EED, thr)
    assert np.array_equal(g.m, o.m)
    g.next_step(DT)
    _wait(g, 60, "config5 slice")
    o.step(DT)
    info = g.last_projection()
    assert info["kernel"] == tfs.PROJ_SYSTOLIC_BLOCK_RING and info["kg"] * info["nw"] == 16 and info["passes"] == 13
    assert_same(g, o, "config5 slice")


def test_config5_slice_2048x32768_K200_block_ring_equals_diagonal_kernel():
    """the 2048-column slice of config 5 (67 M cells, K = 200): shipped kernel vs the one-launch-per-diagonal kernel"""
    W, H, K = 2048, 32768, 200
    thr = frac_to_threshold(0.02)
    sims = []
    for kernel in (tfs.KERNEL_AUTO, tfs.KERNEL_DIAGONAL):
        s = tfs.FluidSim(W, H, iterations=K, kernel=kernel)
        s.scene_random(0x5EED, thr)
        s.next_step(DT)
        _wait(s, 120, f"kernel {kernel}")
        sims.append(s)
    assert np.array_equal(sims[0].checksum(), sims[1].checksum())


def test_block_ring_kernel_equals_diagonal_kernel_on_device_at_8192():
    """k_project_systolic4 against the independent k_project_wavefront (one launch per anti-diagonal), both on the
    device, at a size no CPU oracle reaches in a test: 8192^2, lattice of discs, gravity, K = 24."""
    W = H = 8192
    sims = []
    for kernel in (tfs.KERNEL_AUTO, tfs.KERNEL_DIAGONAL):
        s = tfs.FluidSim(W, H, tfs.SimConfig(gravity=9.8), iterations=24, kernel=kernel)
        s.scene_lattice(1024, 96)
        s.next_step(DT)
        sims.append(s)
    assert sims[0].last_projection()["kernel"] == tfs.PROJ_SYSTOLIC_BLOCK_RING
    assert sims[1].last_projection()["kernel"] == tfs.PROJ_DIAGONAL
    a, b = sims[0].checksum(), sims[1].checksum()
    assert np.array_equal(a, b), (a, b)
    assert bits_equal(sims[0].read_view(tfs.FIELD_P, 4000, 0, 64, H), sims[1].read_view(tfs.FIELD_P, 4000, 0, 64, H))


def test_auto_kernel_is_the_block_ring_kernel_for_every_iteration_count():
    """No silent fallback to the generic kernel: with H % 4 == 0 KERNEL_AUTO runs k_project_systolic4 for K = 1..64
    (and every one of them is bit-identical to the oracle).  224 x 64 = 14336 cells: above the 12800 cells that
    KERNEL_AUTO gives to the one-CTA kernel."""
    W, H = 224, 64
    for K in range(1, 65):
        g = tfs.FluidSim(W, H, tfs.SimConfig(gravity=-3.0), iterations=K)
        o = OracleSim(W, H, gravity=-3.0, iterations=K)
        g.scene_random(K, frac_to_threshold(0.06)); o.scene_random(K, 0.06)
        g.next_step(DT); o.step(DT)
        info = g.last_projection()
        assert info["kernel"] == tfs.PROJ_SYSTOLIC_BLOCK_RING, (K, info)
        assert info["passes"] * info["kg"] * info["nw"] >= K
        assert_same(g, o, f"K={K} shape {info}")
        g.close()
    g = tfs.FluidSim(W, 63, iterations=10)                    # H % 4 != 0: the generic kernel, and it says so (14112 cells)
    g.next_step(DT)
    assert g.last_projection()["kernel"] == tfs.PROJ_SYSTOLIC_GENERIC


def test_workspace_survives_resize_to_smaller_grid_and_iteration_change():
    """progress counters of the block-ring kernel must not be read from stale records when the
    workspace is reused with another layout (smaller grid, other shape, generic kernel first)."""
    g = tfs.FluidSim(128, 64, tfs.SimConfig(gravity=2.0), iterations=50, kernel=tfs.KERNEL_SYSTOLIC)   # not the one-CTA kernel
    o = OracleSim(128, 64, gravity=2.0, iterations=50)
    for sim in (g, o):
        sim.fill_random_velocity(3, 80.0)
    g.scene_random(5, frac_to_threshold(0.07)); o.scene_random(5, 0.07)
    for _ in range(2):
        g.next_step(DT); o.step(DT)
    assert_same(g, o, "before resize")
    g.resize(96, 48); o.resize(96, 48)
    assert np.array_equal(g.m, o.m)
    for st in range(3):
        g.next_step(DT); o.step(DT)
        assert_same(g, o, f"after resize to 96x48, step {st}")
    g.set_options(iterations=8); o.iterations = 8
    for st in range(2):
        g.next_step(DT); o.step(DT)
        assert_same(g, o, f"iterations 50 -> 8, step {st}")
    g.set_options(iterations=3, temporal_block=2); o.iterations = 3
    g.next_step(DT); o.step(DT)
    assert_same(g, o, "iterations 8 -> 3")
    g.resize(64, 30); o.resize(64, 30)                        # generic kernel (H % 4 != 0) shares the workspace
    g.next_step(DT); o.step(DT)
    assert_same(g, o, "64x30")
    g.resize(64, 32); o.resize(64, 32)
    g.set_options(iterations=50, temporal_block=0); o.iterations = 50
    g.next_step(DT); o.step(DT)
    assert_same(g, o, "64x32 after the generic kernel")


def test_checksum_equals_host_formula():
    g = tfs.FluidSim(200, 120, tfs.SimConfig(gravity=1.0), iterations=12)
    g.fill_random_velocity(11, 30.0)
    g.next_step(DT)
    assert np.array_equal(g.checksum(), tfs.checksum_of_fields(g.u, g.v, g.p, g.s))
    u = g.u
    u[[5, 6]] = u[[6, 5]]                                     # position-sensitive: swapping two values changes it
    if u[5] != u[6]:
        g2 = tfs.FluidSim(200, 120)
        g2.write_field(tfs.FIELD_U, u); g2.write_field(tfs.FIELD_V, g.v); g2.write_field(tfs.FIELD_P, g.p)
        g2.write_field(tfs.FIELD_S, g.s)
        assert not np.array_equal(g2.checksum()[:2], g.checksum()[:2])


# ---- red-black mode: temporally blocked kernel -------------------------------------------------------------------
RB_CASES = [(64, 64, 1, 0.0, 0.0), (100, 100, 50, 0.0, 0.05), (130, 50, 5, 2.5, 0.08), (40, 300, 9, -9.8, 0.1),
            (257, 130, 33, 1.0, 0.03), (300, 48, 50, 0.0, 0.0), (512, 512, 20, 9.8, 0.02), (3, 64, 7, 0.0, 0.0),
            (700, 200, 4, 0.0, 0.3), (48, 42, 50, 0.0, 0.02)]


@pytest.mark.parametrize("W,H,K,grav,frac", RB_CASES)
def test_redblack_fused_bit_exact_vs_its_cpu_statement(W, H, K, grav, frac):
    g = tfs.FluidSim(W, H, tfs.SimConfig(gravity=grav), iterations=K, mode=tfs.MODE_RED_BLACK)
    o = OracleSim(W, H, gravity=grav, iterations=K, mode=1)
    if frac:
        g.scene_random(W + H, frac_to_threshold(frac)); o.scene_random(W + H, frac)
    g.fill_random_velocity(21, 40.0); o.fill_random_velocity(21, 40.0)
    for st in range(3):
        g.next_step(DT); o.step(DT)
        want = tfs.PROJ_SMALL_FUSED if W * H <= 12800 else tfs.PROJ_REDBLACK_FUSED
        assert g.last_projection()["kernel"] == want
        assert_same(g, o, f"red-black {W}x{H} K={K} step {st}")
    g.phase_gravity(DT); o.add_gravity(DT)                    # the phases on their own (gravity not fused)
    g.phase_project(DT); o.make_incompressible(DT)
    assert_same(g, o, "phase_project", "uvp")


@pytest.mark.parametrize("W,H,K", [(3, 2, 1), (13, 4, 3), (12, 52, 4), (25, 50, 7), (97, 54, 2), (127, 104, 5), (253, 106, 3),
                                   (126, 52, 6), (5, 208, 4)])
def test_redblack_tiled_kernel_at_strip_and_chunk_boundaries(W, H, K):
    """k_project_rb itself (options.kernel = SYSTOLIC keeps the one-CTA kernel out) on grids whose height sits at the
    edges of its 52-row strips and whose width at the edges of its 12 ... 126-column chunks; K around the 3 fused iterations."""
    g = tfs.FluidSim(W, H, tfs.SimConfig(gravity=1.5), iterations=K, mode=tfs.MODE_RED_BLACK, kernel=tfs.KERNEL_SYSTOLIC)
    o = OracleSim(W, H, gravity=1.5, iterations=K, mode=1)
    g.scene_random(W * 3 + H, frac_to_threshold(0.07)); o.scene_random(W * 3 + H, 0.07)
    g.fill_random_velocity(8, 30.0); o.fill_random_velocity(8, 30.0)
    for st in range(3):
        g.next_step(DT); o.step(DT)
        assert g.last_projection()["kernel"] == tfs.PROJ_REDBLACK_FUSED
        assert_same(g, o, f"red-black {W}x{H} K={K} step {st}")


def test_redblack_odd_height_takes_the_simple_kernel_and_is_exact_too():
    g = tfs.FluidSim(190, 97, iterations=11, mode=tfs.MODE_RED_BLACK)
    o = OracleSim(190, 97, iterations=11, mode=1)
    g.scene_disc(30, 48, 9); o.scene_disc(30, 48, 9)
    for st in range(2):
        g.next_step(DT); o.step(DT)
        assert g.last_projection()["kernel"] == tfs.PROJ_REDBLACK_TWO_LAUNCH
        assert_same(g, o, f"odd H step {st}")


def test_redblack_fused_equals_simple_kernel_on_device_at_4096():
    W = H = 4096
    sims = []
    for kernel in (tfs.KERNEL_AUTO, tfs.KERNEL_DIAGONAL):
        s = tfs.FluidSim(W, H, tfs.SimConfig(gravity=9.8), iterations=30, mode=tfs.MODE_RED_BLACK, kernel=kernel)
        s.scene_lattice(512, 60)
        for _ in range(2):
            s.next_step(DT)
        sims.append(s)
    assert sims[0].last_projection()["kernel"] == tfs.PROJ_REDBLACK_FUSED
    assert sims[1].last_projection()["kernel"] == tfs.PROJ_REDBLACK_TWO_LAUNCH
    assert np.array_equal(sims[0].checksum(), sims[1].checksum())


def _linf(a, b):
    return float(np.max(np.abs(a - b)))


def test_redblack_vs_reference_order_config2_within_measured_tolerance():
    """BASELINE config 2 (1024^2, disc, K = 40, 20 steps).  Measured with scripts/rb_tolerance.py (both orders on the CPU
    oracle; the GPU modes are bit-identical to those): max over the 20 steps L-inf u 9.443 (step 1), v 8.247 (step 1),
    smoke 0.1697 (step 20), residual(rb) / residual(lex) between 0.94 and 0.99.  Tolerance = measured + 20 %."""
    ex = tfs.FluidSim(1024, 1024, iterations=40)
    rb = tfs.FluidSim(1024, 1024, iterations=40, mode=tfs.MODE_RED_BLACK)
    o_rb = OracleSim(1024, 1024, iterations=40, mode=1)
    for s in (ex, rb, o_rb):
        s.scene_disc(256, 512, 128)
    worst = [0.0, 0.0, 0.0]
    for st in range(1, 21):
        ex.next_step(DT); rb.next_step(DT)
        worst = [max(w, _linf(getattr(ex, f), getattr(rb, f))) for w, f in zip(worst, "uvs")]
        assert rb.divergence_linf() <= 1.2 * ex.divergence_linf(), st
        if st <= 3:
            o_rb.step(DT)
            assert_same(rb, o_rb, f"config2 red-black step {st}")
    assert worst[0] <= 9.443 * 1.2 and worst[1] <= 8.247 * 1.2 and worst[2] <= 0.1697 * 1.2, worst
    assert worst[0] >= 9.443 * 0.8 and worst[2] >= 0.1697 * 0.8, worst   # the drift is a property of the orders, not noise


# ---- two x-slabs on one device ----------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [tfs.MODE_WAVEFRONT_EXACT, tfs.MODE_RED_BLACK])
def test_two_virtual_slabs_on_one_device_match_oracle(mode):
    """Both slab handles live on device 0 (their kernels run concurrently and talk through the same flags, records and
    halo pushes as on two GPUs).  The grid is small enough for every CTA of both ranks to be resident at once, which
    a spin-waiting protocol needs on a shared GPU; the wait is bounded so a protocol error fails instead of hanging."""
    W, H, K = 256, 64, 12
    grp = tfs.SlabGroup(W, H, tfs.SimConfig(gravity=1.5), world=2, devices=[0, 0], iterations=K, mode=mode)
    o = OracleSim(W, H, gravity=1.5, iterations=K, mode=1 if mode == tfs.MODE_RED_BLACK else 0)
    thr = frac_to_threshold(0.05)
    grp.each("scene_random", 9, thr); o.L.orc_scene_random(o.h, 9, thr)
    grp.each("fill_random_velocity", 4, 50.0); o.fill_random_velocity(4, 50.0)
    for st in range(3):
        grp.next_step(DT); o.step(DT)
        grp.sync(timeout_s=60)
        for f, n in ((tfs.FIELD_U, "u"), (tfs.FIELD_V, "v"), (tfs.FIELD_P, "p"), (tfs.FIELD_S, "s")):
            a, b = grp.read_field(f), getattr(o, n)
            assert bits_equal(a, b), f"mode {mode} step {st} {n}: " + first_mismatch(a, b, H)
    assert np.array_equal(grp.checksum(), tfs.checksum_of_fields(o.u, o.v, o.p, o.s))
    grp.close()


# ---- the C++ host mirror (include/fluid_sim.hpp) ------------------------------------------------------------------
def test_cpp_mirror_frame_loop_computes_the_same_bits_as_the_ctypes_mirror():
    """host/e2e_mirror drives brush edits, next_step and render_view through include/fluid_sim.hpp; the ctypes mirror
    repeats the same calls.  Same state checksum, and the C++ next_step() does no whole-grid read-back (by construction:
    the mirrors are fetched lazily by the get_*_grid accessors only)."""
    import json
    import subprocess
    from terminal_fluid_sim_b200 import build as tfs_build
    W, H, K, frames = 512, 256, 20, 4
    out = subprocess.run([tfs_build.E2E, str(W), str(H), str(K), "1.5", "disc", str(frames), "0"], capture_output=True, text=True,
                         timeout=120)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout)
    g = tfs.FluidSim(W, H, tfs.SimConfig(gravity=1.5), iterations=K)
    g.scene_disc(256, 512, 128)
    for i in range(3 + frames):
        x = 8 + i % 8
        for y in range(H // 2 - 4, H // 2 + 4):
            (g.unset_block if i & 1 else g.set_block)(x, y)
        g.next_step(DT)
    assert got["checksum"] == ["%016x" % int(v) for v in g.checksum()]
    assert got["d2h_bytes_per_frame"] == 512 * 128 * 8 + 8


# ---- advection reach beyond the halo on x-slabs ------------------------------------------------------
@pytest.mark.parametrize("mode", [tfs.MODE_WAVEFRONT_EXACT, tfs.MODE_RED_BLACK])
@pytest.mark.parametrize("H,amp", [(64, 1400.0), (256, 1400.0), (256, 300.0)])
def test_slab_advection_reaches_into_the_neighbour_slab(mode, H, amp):
    """|vel|*dt up to 23 columns: more than the 10 stored halo columns, so the samples are loaded from the neighbour's
    memory (two slabs on one device; H = 256 takes the packed pair kernel, H = 64 the one-cell kernel)."""
    W, K = 384, 8
    grp = tfs.SlabGroup(W, H, world=2, devices=[0, 0], iterations=K, mode=mode)
    o = OracleSim(W, H, iterations=K, mode=1 if mode == tfs.MODE_RED_BLACK else 0)
    grp.each("fill_random_velocity", 31, amp); o.fill_random_velocity(31, amp)
    for st in range(2):
        grp.next_step(DT); o.step(DT)
        grp.sync(timeout_s=60)
        for f, n in ((tfs.FIELD_U, "u"), (tfs.FIELD_V, "v"), (tfs.FIELD_P, "p"), (tfs.FIELD_S, "s")):
            a, b = grp.read_field(f), getattr(o, n)
            assert bits_equal(a, b), f"mode {mode} H {H} amp {amp} step {st} {n}: " + first_mismatch(a, b, H)
    grp.close()


def test_slab_advection_beyond_the_reachable_columns_is_an_error_not_wrong_data():
    """exact mode: the left neighbour may only be read 32 columns deep (its earlier bands already run the next step);
    a back-trace of ~80 columns must surface as TFS_ERR_OUT_OF_RANGE at the next synchronising call"""
    grp = tfs.SlabGroup(384, 64, world=2, devices=[0, 0], iterations=4)
    grp.each("fill_random_velocity", 31, 5000.0)
    grp.next_step(DT)
    with pytest.raises(tfs.TfsError) as e:
        grp.sync(timeout_s=60)
    assert e.value.code == 2 and "back-trace" in str(e.value)
    grp.each("restart_sim")                                   # clears the condition
    grp.next_step(DT)
    grp.sync(timeout_s=60)
    grp.close()


# ---- hand-derived known answers (tests/kat.py) on the device -------------------------------------------------------
@pytest.mark.parametrize("H", [8, 16, 256])
def test_advection_and_sample_vector_known_answers_on_device(H):
    """values worked by hand from simulator.rs:192-298 (tests/kat.py), independent of both CPU restatements; H = 8 runs the
    one-cell kernel, 16 its L2-prefetching instantiation, 256 the packed two-cell kernel"""
    from tests import kat
    for name, (u, v, sm), expect in kat.advection_cases(H):
        g = tfs.FluidSim(8, H)
        g.write_field(tfs.FIELD_U, u); g.write_field(tfs.FIELD_V, v); g.write_field(tfs.FIELD_S, sm)
        g.phase_advect(kat.DT16)
        for f, x, y, want in expect:
            got = float(getattr(g, f)[x * H + y])
            assert got == want, f"{name} (H={H}): {f}[{x},{y}] = {got!r}, by hand {want!r}"


def test_smoke_pipe_and_wind_column_known_answers_on_device():
    from tests import kat
    for H, size, lo, hi in kat.SMOKE_PIPE:
        g = tfs.FluidSim(5, H, tfs.SimConfig(smoke_size=size, wind_speed=12.5))
        want = np.ones((5, H), np.float32)
        want[0, lo:hi] = 0.0
        assert np.array_equal(g.get_smoke_grid().reshape(5, H), want), (H, size)
        u = g.u.reshape(5, H)
        assert np.all(u[1] == np.float32(12.5)) and np.count_nonzero(u) == H


# ---- terminal-sized grids: the whole step in one CTA (k_step_small) ------------------------------------------------
SMALL_CASES = ["bench_100x100", "term_80x48_disc", "term_48x42_random_gravity", "adversarial_64x40", "odd_37x23", "default_2x2",
               "thin_3x64"]


@pytest.mark.parametrize("name", SMALL_CASES)
def test_small_grid_kernel_reproduces_goldens(name):
    """KERNEL_AUTO on the reference's own sizes (cargo bench 100 x 100, the terminal-derived 80 x 48 and 48 x 42): one launch
    per step, bit-identical to the goldens and to the oracle after every step"""
    from tests.helpers import load_golden, sim_kwargs
    z, prm = load_golden(name)
    kw = sim_kwargs(prm)
    g = tfs.FluidSim(prm["W"], prm["H"], tfs.SimConfig(kw["gravity"], kw["wind_speed"], kw["smoke_size"], kw["density"]),
                     iterations=kw["iterations"], omega=kw["omega"])
    o = OracleSim(prm["W"], prm["H"], kw["gravity"], kw["wind_speed"], kw["smoke_size"], kw["density"], iterations=kw["iterations"],
                  omega=kw["omega"])
    g.write_field(tfs.FIELD_M, z["mask"]); o.m[:] = z["mask"]
    for f, fid in (("u", tfs.FIELD_U), ("v", tfs.FIELD_V), ("s", tfs.FIELD_S)):
        g.write_field(fid, z[f + "0"]); getattr(o, f)[:] = z[f + "0"]
    launches = g.launch_count()
    for st in range(1, prm["steps"] + 1):
        g.next_step(prm["dt"]); o.step(prm["dt"])
        small = prm["W"] * prm["H"] <= 8192                       # exact mode: one CTA up to 8192 cells (100 x 100 takes the systolic kernel)
        assert (g.last_projection()["kernel"] == tfs.PROJ_SMALL_FUSED) == small
        assert_same(g, o, f"{name} step {st}")
        if st in prm["keep"]:
            for f in "uvps":
                assert bits_equal(getattr(g, f), z[f"{f}_step{st}"]), f"{name} golden {f} step {st}"
    if small:
        assert g.launch_count() - launches <= prm["steps"] + 1  # one launch per step (+ the flag refresh after the mask upload)


@pytest.mark.parametrize("W,H,K,mode", [(90, 91, 7, 0), (113, 113, 7, 1), (100, 100, 50, 1), (64, 128, 0, 0), (9, 900, 3, 0),
                                        (1400, 9, 3, 1), (5, 5, 4, 0)])
def test_small_grid_kernel_edge_sizes_and_modes(W, H, K, mode):
    g = tfs.FluidSim(W, H, tfs.SimConfig(gravity=-4.0), iterations=K, mode=mode)
    o = OracleSim(W, H, gravity=-4.0, iterations=K, mode=mode)
    g.scene_random(3, frac_to_threshold(0.1)); o.scene_random(3, 0.1)
    g.fill_random_velocity(8, 90.0); o.fill_random_velocity(8, 90.0)
    for st in range(3):
        g.next_step(DT); o.step(DT)
        assert g.last_projection()["kernel"] == tfs.PROJ_SMALL_FUSED
        assert_same(g, o, f"{W}x{H} K={K} mode {mode} step {st}")
    for Wb, Hb, mb in ((114, 113, 1), (91, 91, 0)):             # 12882 / 8281 cells: just above the red-black / exact limits
        big = tfs.FluidSim(Wb, Hb, iterations=4, mode=mb)
        big.next_step(DT)
        assert big.last_projection()["kernel"] != tfs.PROJ_SMALL_FUSED
