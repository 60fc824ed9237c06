"""CPU proof that the CUDA systolic schedule (oracle/systolic_model.py is its executable
specification) reproduces the reference's lexicographic sweep bit-for-bit, including band
hand-off, multi-pass ping-pong, the partial last pass and the fused gravity."""
import numpy as np
import pytest

from oracle.oracle import OracleSim
from oracle.systolic_model import systolic_project
from tests.helpers import DT, bits_equal

CASES = [(20, 15, 7, 3, 8, 0.0), (20, 15, 7, 3, 8, 9.8), (33, 21, 5, 4, 8, 2.0), (12, 40, 10, 4, 16, -3.0),
         (40, 12, 9, 2, 4, 1.0), (17, 17, 0, 3, 8, 1.0), (25, 30, 6, 6, 32, 0.5), (70, 9, 4, 4, 32, 0.0)]


@pytest.mark.parametrize("W,H,K,KB,LANES,g", CASES)
def test_schedule_is_bit_exact(W, H, K, KB, LANES, g):
    o = OracleSim(W, H, gravity=g, iterations=K)
    rng = np.random.default_rng(W * H + K)
    for _ in range(W * H // 6):
        o.set_block(int(rng.integers(0, W)), int(rng.integers(0, H)))
    o.fill_random_velocity(7, 30.0)
    flags, u0, v0 = o.cell_flags(), o.u.copy(), o.v.copy()
    pc = np.float32(1000.0) / np.float32(DT)
    gdt = np.float32(g) * np.float32(DT)
    u, v, p = systolic_project(u0, v0, flags, W, H, K, KB, LANES, np.float32(1.9), pc, gdt)
    o.add_gravity(DT)
    o.make_incompressible(DT)
    assert bits_equal(u, o.u) and bits_equal(v, o.v) and bits_equal(p, o.p)
