"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI via the
FluidSim mirror, against the CPU oracle on the same inputs and against the committed goldens.
Bar: bit-exact for the exact (wavefront) mode and for all integer work; stated tolerances for
the red-black mode."""
import numpy as np
import pytest

import terminal_fluid_sim_b200 as tfs
from oracle.oracle import OracleSim, frac_to_threshold
from tests.helpers import DT, bits_equal, first_mismatch, load_golden, sim_kwargs

pytestmark = pytest.mark.gpu

EXACT_KERNELS = [tfs.KERNEL_DIAGONAL, tfs.KERNEL_SYSTOLIC]
CASES = ["bench_100x100", "term_80x48_disc", "term_48x42_random_gravity", "adversarial_64x40",
         "odd_37x23", "default_2x2", "thin_3x64"]


def make_pair(W, H, *, gravity=0.0, wind_speed=50.0, smoke_size=0.25, density=1000.0, iterations=50,
              omega=1.9, mode=0, kernel=tfs.KERNEL_AUTO, temporal_block=0):
    g = tfs.FluidSim(W, H, tfs.SimConfig(gravity, wind_speed, smoke_size, density), iterations=iterations,
                     omega=omega, mode=mode, kernel=kernel, temporal_block=temporal_block)
    o = OracleSim(W, H, gravity, wind_speed, smoke_size, density, iterations=iterations, omega=omega, mode=mode)
    return g, o


def assert_same(g, o, what, fields="uvps"):
    H = o.height
    for f in fields:
        a, b = getattr(g, f), getattr(o, f)
        assert bits_equal(a, b), f"{what}: {f}: " + first_mismatch(a, b, H)


def skip_if_unavailable(g, kernel):
    if kernel == tfs.KERNEL_SYSTOLIC:
        try:
            g.phase_project(DT)
        except tfs.TfsError as e:
            pytest.skip(f"systolic kernel unavailable: {e}")


@pytest.mark.parametrize("kernel", EXACT_KERNELS)
@pytest.mark.parametrize("name", CASES)
def test_golden_cases_bit_exact(name, kernel):
    z, prm = load_golden(name)
    g, o = make_pair(prm["W"], prm["H"], kernel=kernel, **sim_kwargs(prm))
    skip_if_unavailable(tfs.FluidSim(8, 8, kernel=kernel), kernel)
    g.write_field(tfs.FIELD_M, z["mask"])
    o.m[:] = z["mask"]
    for f, fid in (("u", tfs.FIELD_U), ("v", tfs.FIELD_V), ("s", tfs.FIELD_S)):
        g.write_field(fid, z[f + "0"])
        getattr(o, f)[:] = z[f + "0"]
    assert np.array_equal(g.flags, o.cell_flags())
    for st in range(1, prm["steps"] + 1):
        g.next_step(prm["dt"])
        o.step(prm["dt"])
        assert_same(g, o, f"{name} step {st}")
        if st in prm["keep"]:
            for f in "uvps":
                assert bits_equal(getattr(g, f), z[f"{f}_step{st}"]), f"{name} golden {f} step {st}"


def test_phases_individually_bit_exact():
    g, o = make_pair(96, 72, gravity=-9.8, iterations=23, kernel=tfs.KERNEL_DIAGONAL)
    o.scene_disc(30, 36, 9)
    g.scene_disc(30, 36, 9)
    assert np.array_equal(g.m, o.m)
    for _ in range(3):
        g.phase_gravity(DT); o.add_gravity(DT)
        assert_same(g, o, "gravity", "v")
        g.phase_project(DT); o.make_incompressible(DT)
        assert_same(g, o, "project", "uvp")
        g.phase_advect(DT); o.move_velocity(DT)
        assert_same(g, o, "advect", "uvs")


@pytest.mark.parametrize("kernel", EXACT_KERNELS)
@pytest.mark.parametrize("W,H,K,tb", [(64, 64, 1, 0), (65, 63, 2, 0), (130, 50, 5, 3), (40, 300, 9, 4),
                                      (257, 129, 33, 16), (300, 40, 50, 7), (512, 512, 20, 0)])
def test_random_obstacles_and_fields_bit_exact(W, H, K, tb, kernel):
    g, o = make_pair(W, H, gravity=2.5, iterations=K, kernel=kernel, temporal_block=tb)
    skip_if_unavailable(tfs.FluidSim(8, 8, kernel=kernel), kernel)
    thr = frac_to_threshold(0.08)
    g.scene_random(W * 131 + H, thr); o.L.orc_scene_random(o.h, W * 131 + H, thr)
    g.fill_random_velocity(99, 40.0); o.fill_random_velocity(99, 40.0)
    assert np.array_equal(g.m, o.m) and np.array_equal(g.flags, o.cell_flags())
    assert_same(g, o, "initial", "uv")
    for st in range(2):
        g.next_step(DT); o.step(DT)
        assert_same(g, o, f"{W}x{H} K={K} step {st}")


@pytest.mark.parametrize("W,H,K", [(33, 4, 5), (8, 8, 20), (100, 12, 16), (5, 16, 50), (64, 20, 17), (3, 4, 1), (200, 24, 33)])
def test_systolic_kernel_on_short_columns_bit_exact(W, H, K):
    """The block-ring systolic kernel (H % 4 == 0) forced onto grids whose columns are shorter than one ring block or than
    the 32-lane skew: the feeder warp's priming rounds, the first / last ring block and the record intervals all degenerate."""
    g, o = make_pair(W, H, gravity=-3.0, iterations=K, kernel=tfs.KERNEL_SYSTOLIC)
    skip_if_unavailable(tfs.FluidSim(8, 8, kernel=tfs.KERNEL_SYSTOLIC), tfs.KERNEL_SYSTOLIC)
    thr = frac_to_threshold(0.1)
    g.scene_random(W * 7 + H, thr); o.L.orc_scene_random(o.h, W * 7 + H, thr)
    g.fill_random_velocity(5, 25.0); o.fill_random_velocity(5, 25.0)
    for st in range(3):
        g.next_step(DT); o.step(DT)
        assert_same(g, o, f"{W}x{H} K={K} step {st}")
    assert g.last_projection()["kernel"] in (tfs.PROJ_SYSTOLIC_BLOCK_RING, tfs.PROJ_NONE)


def test_config2_wind_tunnel_1024_bit_exact():
    """BASELINE config 2: 1024^2 wind tunnel, disc obstacle, K=40."""
    g, o = make_pair(1024, 1024, iterations=40)
    g.scene_disc(256, 512, 128); o.scene_disc(256, 512, 128)
    assert np.array_equal(g.m, o.m)
    for st in range(3):
        g.next_step(DT); o.step(DT)
        assert_same(g, o, f"config2 step {st}")
    assert abs(g.divergence_linf() - o.divergence_linf()) == 0.0
    p = o.p
    assert g.pressure_minmax() == (float(p.min()), float(p.max()))


def test_adversarial_velocities_bit_exact():
    g, o = make_pair(200, 120, iterations=8, gravity=9.8)
    g.fill_random_velocity(0xBEEF, 300.0); o.fill_random_velocity(0xBEEF, 300.0)
    for st in range(3):
        g.next_step(DT); o.step(DT)
        assert_same(g, o, f"adversarial step {st}")


def test_redblack_matches_its_cpu_statement_and_is_close_to_reference_order():
    W = H = 128
    g, o_rb = make_pair(W, H, mode=tfs.MODE_RED_BLACK, iterations=50)
    o_lex = OracleSim(W, H, iterations=50)
    for sim in (g, o_rb, o_lex):
        sim.scene_disc(40, 64, 12)
    for st in range(10):
        g.next_step(DT); o_rb.step(DT); o_lex.step(DT)
        assert_same(g, o_rb, f"red-black step {st}")
    # stated tolerance vs the reference's (lexicographic) order after 10 steps of K=50:
    # velocity / smoke L-inf within 25% of wind speed / 0.25, divergence residual within 2x
    assert np.max(np.abs(g.u - o_lex.u)) < 0.25 * 50.0
    assert np.max(np.abs(g.v - o_lex.v)) < 0.25 * 50.0
    assert np.max(np.abs(g.s - o_lex.s)) < 0.25
    assert g.divergence_linf() < 2.0 * o_lex.divergence_linf() + 1e-3


def test_edit_path_matches_reference_semantics():
    g, o = make_pair(20, 10)
    rng = np.random.default_rng(2)
    for _ in range(30):
        x, y = int(rng.integers(0, 20)), int(rng.integers(0, 10))
        g.set_block(x, y); o.set_block(x, y)
    g.next_step(DT); o.step(DT)
    g.resize(16, 12); o.resize(16, 12)                      # resize_block_grid no-remap quirk
    assert g.get_size() == (16, 12) and np.array_equal(g.m, o.m)
    assert_same(g, o, "after resize")
    g.next_step(DT); o.step(DT)
    g.set_config(tfs.SimConfig(0.5, 20.0, 0.5, 900.0)); o.set_config(0.5, 20.0, 0.5, 900.0)
    assert_same(g, o, "after set_config")
    g.next_step(DT); o.step(DT)
    assert_same(g, o, "step after set_config")
    g.unset_block(3, 3); o.unset_block(3, 3)
    g.restart_sim(); o.restart_sim()
    assert np.array_equal(g.m, o.m)
    assert_same(g, o, "after restart")
    z = np.load(__import__("os").path.join(__import__("tests.helpers", fromlist=["GOLDEN"]).GOLDEN, "edit_path.npz"))
    assert np.array_equal(g.m, z["m_after_restart"])
    with pytest.raises(IndexError):
        g.set_block(16, 0)                                  # the reference panics (simulator.rs:333)
    g.set_block(0, 13)                                      # flat index 13 is in range: allowed, like the reference
    assert g.m[13] == 1


def test_read_view_and_upload_blocks():
    g, o = make_pair(48, 40)
    for _ in range(2):
        g.next_step(DT); o.step(DT)
    s = o.s.reshape(48, 40)
    win = g.read_view(tfs.FIELD_S, 5, 7, 11, 13).reshape(11, 13)
    assert bits_equal(win, s[5:16, 7:20])
    full = g.read_view(tfs.FIELD_S, 3, 0, 20, 40).reshape(20, 40)
    assert bits_equal(full, s[3:23])
    with pytest.raises(tfs.TfsError):
        g.read_view(tfs.FIELD_S, 40, 0, 9, 40)
    m = o.m.copy(); m[100:140] = 1
    o.m[:] = m
    g.upload_blocks(m[100:140], offset=100)
    assert np.array_equal(g.m, o.m) and np.array_equal(g.flags, o.cell_flags())
    g.next_step(DT); o.step(DT)
    assert_same(g, o, "after upload_blocks")


def test_zero_dt_and_zero_iterations():
    g, o = make_pair(32, 32, iterations=0)
    g.next_step(DT); o.step(DT)
    assert_same(g, o, "K=0")
    g2, o2 = make_pair(32, 32, iterations=4)
    g2.next_step(0.0); o2.step(0.0)
    assert_same(g2, o2, "dt=0")          # p is inf/NaN on both sides (NaN payloads are not compared)


def test_division_shortcut_is_ieee_exact_for_all_inputs():
    assert tfs.selftest_division(0) == 0


def test_launch_counter_moves():
    g = tfs.FluidSim(64, 64)
    a = g.launch_count()
    g.next_step(DT); g.sync()
    assert g.launch_count() > a


# ---- renderer colour path on the device (src/ui/sim_renderer.rs:23-151) ---------------
@pytest.mark.parametrize("W,H,aw,ah", [(80, 48, 80, 24), (100, 100, 100, 50), (37, 23, 30, 11), (200, 150, 131, 70),
                                       (64, 64, 10, 3), (7, 2, 7, 1), (3, 1, 3, 0)])
def test_render_view_cells_bit_exact(W, H, aw, ah):
    g, o = make_pair(W, H, gravity=-9.8, iterations=10)
    rng = np.random.default_rng(W * 1000 + H)
    for _ in range(W * H // 15):
        x, y = int(rng.integers(0, W)), int(rng.integers(0, H))
        g.set_block(x, y); o.set_block(x, y)
    for _ in range(3):
        g.next_step(DT); o.step(DT)
    want, mn, mx = o.render_view(aw, ah)
    got, gmn, gmx = g.render_view(aw, ah)
    assert (gmn, gmx) == (mn, mx) == g.pressure_minmax()
    assert np.array_equal(got, want), f"{np.argwhere(got != want)[:4].tolist()}"
    if aw >= 8 and ah >= 2:                                  # a window that starts at x0 > 0, caller-supplied min/max
        sub, _, _ = g.render_view(aw - 5, ah - 1, x0=5, minmax=(mn, mx))
        assert np.array_equal(sub, want[:ah - 1, 5:aw])


def test_render_view_goldens_and_saturating_casts():
    z = np.load(__import__("os").path.join(__import__("tests.helpers", fromlist=["GOLDEN"]).GOLDEN, "render_view.npz"))
    g = tfs.FluidSim(40, 36)
    g.write_field(tfs.FIELD_P, z["synthetic_p"]); g.write_field(tfs.FIELD_S, z["synthetic_s"])
    g.write_field(tfs.FIELD_M, z["synthetic_m"])
    cells, mn, mx = g.render_view(40, 18)
    assert np.array_equal(cells, z["synthetic"])             # NaN pressure/smoke, smoke outside [0, 1], every colour arm
    assert np.array_equal(np.array([mn, mx], np.float32), z["synthetic_minmax"])
    for name in ("term_80x48_disc", "term_48x42_random_gravity", "odd_37x23"):
        zz, prm = load_golden(name)
        st = max(prm["keep"])
        s = tfs.FluidSim(prm["W"], prm["H"])
        s.write_field(tfs.FIELD_P, zz[f"p_step{st}"]); s.write_field(tfs.FIELD_S, zz[f"s_step{st}"])
        s.write_field(tfs.FIELD_M, zz["mask"])
        aw, ah = (int(v) for v in z[name + "_area"])
        cells, mn, mx = s.render_view(aw, ah)
        assert np.array_equal(cells, z[name]), name
        assert np.array_equal(np.array([mn, mx], np.float32), z[name + "_minmax"])


def test_render_view_rejects_what_the_reference_panics_on():
    g = tfs.FluidSim(8, 8)
    with pytest.raises(tfs.TfsError):
        g.render_view(8, 5)
    with pytest.raises(tfs.TfsError):
        g.render_view(9, 4)
    with pytest.raises(tfs.TfsError):
        g.render_view(4, 4, x0=5)
    cells, _, _ = g.render_view(0, 0)
    assert cells.size == 0


# ---- packed two-cells-per-thread advection (k_advect32_pair: taken when H % 256 == 0) ------------
@pytest.mark.parametrize("W,H,amp,frac", [(96, 256, 120.0, 0.02), (40, 512, 2000.0, 0.02), (300, 768, 30.0, 0.1),
                                          (64, 256, 0.5, 0.0), (5, 256, 50.0, 0.3)])
def test_pair_advection_kernel_bit_exact(W, H, amp, frac):
    g, o = make_pair(W, H, gravity=3.0, iterations=6)
    if frac:
        g.scene_random(7, frac_to_threshold(frac)); o.scene_random(7, frac)
    g.fill_random_velocity(5, amp); o.fill_random_velocity(5, amp)       # amp 2000: every clamp of sample_vector binds
    for st in range(3):
        g.phase_advect(DT); o.move_velocity(DT)
        assert_same(g, o, f"advect {st}", fields="uvs")
        g.next_step(DT); o.step(DT)
        assert_same(g, o, f"step {st}")


@pytest.mark.parametrize("kernel", EXACT_KERNELS)
def test_3x3_projection_known_answer_from_source_text(kernel):
    """The hand-worked value of tests/test_oracle.py::test_3x3_projection_worked_by_hand_from_the_source_text,
    on the device: correction = 1.9 * (50 / 3) at the only updatable cell of a 3x3 grid."""
    f = np.float32
    c = f(1.9) * (f(50.0) / f(3.0))
    pc = f(1000.0) / f(DT)
    g = tfs.FluidSim(3, 3, iterations=1, kernel=kernel)
    g.phase_project(DT)
    u, v, p = g.u, g.v, g.p
    assert u[4] == f(50.0) and u[7] == c and v[4] == -c and v[5] == c and p[4] == pc * c
    assert np.count_nonzero(v) == 2 and np.count_nonzero(p) == 1 and np.count_nonzero(u) == 4
