"""Renderer colour path (src/ui/sim_renderer.rs:23-151): the C oracle against the
independent NumPy restatement, the committed golden cells, and values worked out by hand from
the reference's source text."""
import os

import numpy as np
import pytest

from oracle.np_oracle import render_view as np_render
from oracle.oracle import OracleSim
from tests.helpers import DT, GOLDEN, load_golden, sim_kwargs


def _oracle_from(p, s, m, W, H):
    o = OracleSim(W, H)
    o.p[:] = p; o.s[:] = s; o.m[:] = m
    return o


def test_gradient_known_answers_from_source_text():
    # get_linear_gradient (:116-140) worked by hand with min = 0, max = 8:
    #   value 0 -> normalized 0    -> type 0, cv 0   -> (0, 0, 1)       -> Rgb(0, 0, 255)
    #   value 1 -> normalized .125 -> type 0, cv .5  -> (0, .5, 1)      -> Rgb(0, 127, 255)
    #   value 2 -> normalized .25  -> type 1, cv 0   -> (0, 1, 1)       -> Rgb(0, 255, 255)
    #   value 4 -> normalized .5   -> type 2, cv 0   -> (0, 1, 0)       -> Rgb(0, 255, 0)
    #   value 7 -> normalized .875 -> type 3, cv .5  -> (1, .5, 0)      -> Rgb(255, 127, 0)
    #   value 8 -> normalized 1    -> type 4 (`_`)   -> (1, 0, 0)       -> Rgb(255, 0, 0)
    W, H = 4, 4
    p = np.zeros(W * H, np.float32); s = np.zeros(W * H, np.float32); m = np.zeros(W * H, np.uint8)
    # terminal cell (x, row): bg = sim (x, H-1-2row), fg = sim (x, H-2-2row)
    p[0 * H + 3], p[0 * H + 2] = 0, 1
    p[1 * H + 3], p[1 * H + 2] = 2, 4
    p[2 * H + 3], p[2 * H + 2] = 7, 8
    m[3 * H + 3] = 1                      # block -> THEME.sim_blocks
    s[2 * H + 3] = 0.5                    # get_color: reducer (255*0.5) as u8 = 127, saturating_sub
    cells, mn, mx = _oracle_from(p, s, m, W, H).render_view(4, 1)
    assert (mn, mx) == (0.0, 8.0)
    assert cells[0, 0].tolist() == [[0, 0, 255, 0], [0, 127, 255, 0]]
    assert cells[0, 1].tolist() == [[0, 255, 255, 0], [0, 255, 0, 0]]
    assert cells[0, 2].tolist() == [[128, 0, 0, 0], [255, 0, 0, 0]]     # (255,127,0) - 127 saturating
    assert cells[0, 3, 0].tolist() == [255, 255, 255, 1]
    # difference == 0 -> normalized 0.5 -> type 2, cv 0 -> green (:118-122).  The `else if` of :31-37
    # withholds only the FIRST element from max, so a constant field still gives min == max.
    o = OracleSim(4, 4); o.p[:] = 3.0; o.s[:] = 0; o.m[:] = 0
    cells, mn, mx = o.render_view(4, 2)
    assert (mn, mx) == (3.0, 3.0)
    assert (cells == [0, 255, 0, 0]).all()


@pytest.mark.parametrize("name", ["term_80x48_disc", "term_48x42_random_gravity", "odd_37x23"])
def test_c_oracle_reproduces_render_golden(name):
    z, prm = load_golden(name)
    g = np.load(os.path.join(GOLDEN, "render_view.npz"))
    st = max(prm["keep"])
    aw, ah = (int(v) for v in g[name + "_area"])
    o = _oracle_from(z[f"p_step{st}"], z[f"s_step{st}"], z["mask"], prm["W"], prm["H"])
    cells, mn, mx = o.render_view(aw, ah)
    assert np.array_equal(cells, g[name])
    assert np.array_equal(np.array([mn, mx], np.float32), g[name + "_minmax"])
    assert cells[..., :3].any() and (cells[..., 3] == 1).any()       # not a trivially black frame


def test_synthetic_field_covers_every_branch():
    g = np.load(os.path.join(GOLDEN, "render_view.npz"))
    p, s, m = g["synthetic_p"], g["synthetic_s"], g["synthetic_m"]
    cells, mn, mx = _oracle_from(p, s, m, 40, 36).render_view(40, 18)
    assert np.array_equal(cells, g["synthetic"])
    assert np.array_equal(np.array([mn, mx], np.float32), g["synthetic_minmax"])
    rgb = cells[cells[..., 3] == 0][:, :3]
    # all five colour_type arms occur: pure-blue..cyan, cyan..green, green..yellow, yellow..red, red
    assert ((rgb[:, 2] == 255) & (rgb[:, 0] == 0)).any() and ((rgb[:, 0] == 255) & (rgb[:, 2] == 0)).any()
    assert (rgb == [255, 0, 0]).all(1).any()


@pytest.mark.parametrize("W,H,aw,ah,seed", [(33, 70, 33, 35, 1), (64, 64, 10, 3, 2), (5, 9, 5, 4, 3), (7, 2, 7, 1, 4), (3, 1, 3, 0, 5)])
def test_c_and_numpy_restatements_agree_after_real_steps(W, H, aw, ah, seed):
    o = OracleSim(W, H, gravity=-9.8, iterations=8)
    rng = np.random.default_rng(seed)
    for _ in range(W * H // 12):
        o.set_block(int(rng.integers(0, W)), int(rng.integers(0, H)))
    for _ in range(3):
        o.step(DT)
    a, mn, mx = o.render_view(aw, ah)
    b, mn2, mx2 = np_render(o.p, o.s, o.m, W, H, aw, ah)
    assert np.array_equal(a, b) and (mn, mx) == (mn2, mx2)


def test_oversized_area_panics_like_reference():
    o = OracleSim(8, 8)
    with pytest.raises(IndexError):
        o.render_view(8, 5)       # y_index -= 1 underflows (sim_renderer.rs:60)
    with pytest.raises(IndexError):
        o.render_view(9, 4)       # index past the grid
