"""CPU tests: the C oracle against the committed golden vectors (minted by the independent
NumPy restatement), the two restatements against each other, and the wavefront-order proof."""
import hashlib
import os

import numpy as np
import pytest

from oracle.np_oracle import NpSim
from oracle.oracle import OracleSim, frac_to_threshold
from tests.helpers import DT, GOLDEN, bits_equal, first_mismatch, load_golden, sim_kwargs

CASES = ["bench_100x100", "term_80x48_disc", "term_48x42_random_gravity", "adversarial_64x40",
         "odd_37x23", "default_2x2", "thin_3x64"]


def _digest(o):
    h = hashlib.sha256()
    for a in (o.u, o.v, o.p, o.s):
        h.update(np.ascontiguousarray(a, np.float32).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_reproduces_golden(name):
    z, prm = load_golden(name)
    o = OracleSim(prm["W"], prm["H"], **sim_kwargs(prm))
    o.m[:] = z["mask"]
    o.u[:], o.v[:], o.s[:] = z["u0"], z["v0"], z["s0"]
    for st in range(1, prm["steps"] + 1):
        o.step(prm["dt"])
        assert _digest(o) == prm["hashes"][st - 1], f"{name}: step {st} digest differs"
        if st in prm["keep"]:
            for f in "uvps":
                assert bits_equal(getattr(o, f), z[f"{f}_step{st}"]), \
                    f"{name} {f} step {st}: " + first_mismatch(getattr(o, f), z[f"{f}_step{st}"], prm["H"])


def test_edit_path_golden():
    z = np.load(os.path.join(GOLDEN, "edit_path.npz"))
    o = OracleSim(20, 10)
    rng = np.random.default_rng(2)
    for _ in range(30):
        o.set_block(int(rng.integers(0, 20)), int(rng.integers(0, 10)))
    o.step(DT)
    o.resize(16, 12)
    assert np.array_equal(o.m, z["m_after_resize"])
    o.step(DT)
    o.set_config(0.5, 20.0, 0.5, 900.0)
    assert bits_equal(o.u, z["u_after_cfg"]) and bits_equal(o.s, z["s_after_cfg"])
    o.step(DT)
    for f in "uvps":
        assert bits_equal(getattr(o, f), z[f + "_after"])
    o.restart_sim()
    assert np.array_equal(o.m, z["m_after_restart"])
    assert bits_equal(o.u, z["u_after_restart"]) and bits_equal(o.s, z["s_after_restart"])


@pytest.mark.parametrize("W,H,K,g", [(13, 9, 7, 0.0), (20, 31, 12, 9.8), (9, 40, 5, -3.0)])
def test_wavefront_order_equals_lexicographic(W, H, K, g):
    """sweeping by t = i + j + 2k is bit-identical to the reference's
    lexicographic sweep.  NumPy literal triple loop vs NumPy wavefront vs C oracle."""
    rng = np.random.default_rng(W * H)
    a = NpSim(W, H, gravity=g, iterations=K)
    b = NpSim(W, H, gravity=g, iterations=K)
    c = OracleSim(W, H, gravity=g, iterations=K)
    for _ in range(W * H // 7):
        x, y = int(rng.integers(0, W)), int(rng.integers(0, H))
        a.set_block(x, y), b.set_block(x, y), c.set_block(x, y)
    for _ in range(3):
        a.step(DT, lexicographic=True)
        b.step(DT)
        c.step(DT)
        for f in "uvps":
            assert bits_equal(getattr(a, f), getattr(b, f)), f
            assert bits_equal(getattr(a, f), getattr(c, f)), f


def test_default_2x2_is_noop():
    o = OracleSim(2, 2)
    u0, s0 = o.u.copy(), o.s.copy()
    o.step(DT)
    assert bits_equal(o.u, u0) and bits_equal(o.s, s0)


def test_3x3_projection_worked_by_hand_from_the_source_text():
    """Known-answer value derived from simulator.rs by hand, independent of both restatements.

    3x3 grid, default config (wind 50, density 1000), one iteration, dt = 1/60, no gravity: the only cell that is
    neither block nor border is (1, 1), idx 4 (simulator.rs:161, :365-375).  Column 0 is solid (:113-119) and
    set_horizontal_speed wrote u[3..6) = 50 (:79-85).  Around idx 4: top 5, right 7, bottom 3, left 1 (:343-350);
    only the LEFT neighbour is a block, so sT = sR = sB = 1, sL = 0 and n = 3 (:165-170).
    divergence = ((u[7] - u[4]) + v[5]) - v[4] = (0 - 50) + 0 - 0 = -50 (:176-178);
    correction = 1.9 * (50 / 3) (:180); u[4] -= c*0, u[7] += c, v[4] -= c, v[5] += c (:181-185);
    p[4] += (1000 / dt) * c (:155, :186)."""
    f = np.float32
    c = f(1.9) * (f(50.0) / f(3.0))
    pc = f(1000.0) / f(DT)
    o = OracleSim(3, 3, iterations=1)
    o.make_incompressible(DT)
    u, v, p = o.u, o.v, o.p
    assert u[4] == f(50.0) and u[7] == c and v[4] == -c and v[5] == c and p[4] == pc * c
    assert u[3] == 50.0 and u[5] == 50.0                    # the border cells of column 1 keep their wind
    assert np.count_nonzero(v) == 2 and np.count_nonzero(p) == 1 and np.count_nonzero(u) == 4
    n = NpSim(3, 3, iterations=1)
    n.make_incompressible(DT)
    assert bits_equal(n.u, u) and bits_equal(n.v, v) and bits_equal(n.p, p)


def test_initial_state_matches_reference_constants():
    o = OracleSim(100, 100)
    H = 100
    assert np.all(o.u[H:2 * H] == 50.0) and np.all(o.u[:H] == 0) and np.all(o.u[2 * H:] == 0)
    assert np.all(o.m[:H] == 1) and np.all(o.m[H:] == 0)
    s = o.s
    assert np.all(s[37:62] == 0.0) and np.all(s[:37] == 1.0) and np.all(s[62:] == 1.0)  # rows 37..62 (SURVEY a-7)


def test_set_block_out_of_range_panics_like_reference():
    o = OracleSim(8, 8)
    with pytest.raises(IndexError):
        o.set_block(8, 0)


def test_zero_dt_gives_nonfinite_pressure_only():
    o = OracleSim(16, 16)
    o.step(0.0)
    assert np.all(np.isfinite(o.u)) and np.all(np.isfinite(o.v))
    assert not np.all(np.isfinite(o.p))  # pc = density / 0 = inf (SURVEY a-11)


def test_cell_flags_and_scenes_are_deterministic():
    o = OracleSim(64, 48)
    o.scene_disc(20, 24, 6)
    o.scene_random(0x5EED, 0.02)
    f = o.cell_flags()
    m = o.m.reshape(64, 48)
    assert f.reshape(64, 48)[0].max() == 0 and f.reshape(64, 48)[:, 0].max() == 0
    i, j = 10, 10
    if not m[i, j]:
        exp = 16 | (1 * (not m[i - 1, j])) | (2 * (not m[i + 1, j])) | (4 * (not m[i, j - 1])) | (8 * (not m[i, j + 1]))
        assert f[i * 48 + j] == exp
    assert frac_to_threshold(0.02) == (2 << 64) // 100
    o2 = OracleSim(64, 48)
    o2.scene_disc(20, 24, 6)
    o2.scene_random(0x5EED, 0.02)
    assert np.array_equal(o.m, o2.m)


def test_redblack_variant_converges_to_same_flow_class():
    """The red-black ordering is a different iteration, not the reference's; it must stay
    within a loose tolerance of the lexicographic result on a smooth case."""
    a = OracleSim(64, 64, iterations=50, mode=0)
    b = OracleSim(64, 64, iterations=50, mode=1)
    for _ in range(5):
        a.step(DT)
        b.step(DT)
    assert np.max(np.abs(a.u - b.u)) < 0.25 * 50.0
    assert abs(a.divergence_linf() - b.divergence_linf()) < 50.0


# ---- hand-derived known answers (tests/kat.py): they pin BOTH restatements to the source text ----------------------
from tests import kat


@pytest.mark.parametrize("H", [8, 256])
def test_advection_and_sample_vector_known_answers(H):
    for name, (u, v, sm), expect in kat.advection_cases(H):
        for make in (lambda: OracleSim(8, H), lambda: NpSim(8, H)):
            o = make()
            o.u[:], o.v[:], o.s[:] = u, v, sm
            o.move_velocity(kat.DT16)
            for f, x, y, want in expect:
                got = float(getattr(o, f)[x * H + y])
                assert got == want, f"{type(o).__name__} {name}: {f}[{x},{y}] = {got!r}, by hand {want!r}"


@pytest.mark.parametrize("H,size,lo,hi", kat.SMOKE_PIPE)
def test_smoke_pipe_and_wind_column_known_answers(H, size, lo, hi):
    for o in (OracleSim(5, H, smoke_size=size, wind_speed=12.5), NpSim(5, H, smoke_size=size, wind_speed=12.5)):
        s = np.array(o.s).reshape(5, H)
        want = np.ones((5, H), np.float32)
        want[0, lo:hi] = 0.0
        assert np.array_equal(s, want), (type(o).__name__, H, size)
        u = np.array(o.u).reshape(5, H)
        assert np.all(u[1] == np.float32(12.5)) and np.count_nonzero(u) == H
