"""Multi-GPU x-slab tests (need >= 2 CUDA devices; skipped otherwise): the slab handles, driven from
one host thread, must reproduce the CPU oracle bit-for-bit, i.e. also the single-GPU result."""
import numpy as np
import pytest

import terminal_fluid_sim_b200 as tfs
from oracle.oracle import OracleSim, frac_to_threshold
from tests.helpers import DT, bits_equal, first_mismatch

pytestmark = pytest.mark.gpu


def _ndev():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


FIELDS = ((tfs.FIELD_U, "u"), (tfs.FIELD_V, "v"), (tfs.FIELD_P, "p"), (tfs.FIELD_S, "s"))


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("W,H,K,tb,g", [(256, 64, 4, 4, 0.0), (520, 96, 50, 0, 9.8), (1024, 256, 40, 16, 0.0),
                                       (1024, 256, 64, 32, 1.0)])   # KB = 32: band 0 cannot deliver the halo, whole-kernel flags
def test_slabs_match_oracle_bit_exact(world, W, H, K, tb, g):
    if _ndev() < world:
        pytest.skip(f"needs {world} GPUs")
    if (W + 1 + 31) // 32 < world:
        pytest.skip("grid too narrow for this many slabs")
    grp = tfs.SlabGroup(W, H, tfs.SimConfig(gravity=g), world=world, iterations=K, temporal_block=tb)
    o = OracleSim(W, H, gravity=g, iterations=K)
    thr = frac_to_threshold(0.05)
    grp.each("scene_random", 77, thr); o.L.orc_scene_random(o.h, 77, thr)
    grp.each("fill_random_velocity", 5, 60.0); o.fill_random_velocity(5, 60.0)
    assert np.array_equal(grp.read_field(tfs.FIELD_M), o.m)
    # the slabs tile the grid exactly
    bounds = [s.get_slab()[:2] for s in grp.sims]
    assert bounds[0][0] == 0 and bounds[-1][1] == W and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
    for st in range(3):
        grp.next_step(DT); o.step(DT)
        grp.sync()
        for f, n in FIELDS:
            a, b = grp.read_field(f), getattr(o, n)
            assert bits_equal(a, b), f"world={world} step {st} {n}: " + first_mismatch(a, b, H)
    # edits reach every slab that stores the cell; results stay exact afterwards
    x = bounds[0][1]  # first column of slab 1: stored by slabs 0 (halo) and 1 (owned)
    grp.each("set_block", x, H // 2); o.set_block(x, H // 2)
    grp.next_step(DT); o.step(DT); grp.sync()
    for f, n in FIELDS:
        assert bits_equal(grp.read_field(f), getattr(o, n)), n
    # several steps back to back without host synchronisation: the ranks drift apart by their natural skew and
    # only the device-side flags (in-kernel halo hand-offs included) order them
    for _ in range(5):
        grp.next_step(DT); o.step(DT)
    grp.sync()
    for f, n in FIELDS:
        a, b = grp.read_field(f), getattr(o, n)
        assert bits_equal(a, b), f"world={world} after 5 unsynchronised steps {n}: " + first_mismatch(a, b, H)
    grp.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_config4_like_lattice_slabs_equal_single_gpu_bitwise(world):
    """SURVEY 8d config 4, scaled down: wind tunnel with a lattice of discs, K = 50; the x-slab run must equal the
    single-GPU run bit for bit (the single-GPU kernel is itself pinned to the oracle by the other tests)."""
    if _ndev() < world:
        pytest.skip(f"needs {world} GPUs")
    W, H, K = 4096, 1024, 50
    one = tfs.FluidSim(W, H, iterations=K)
    grp = tfs.SlabGroup(W, H, world=world, iterations=K)
    one.scene_lattice(512, 48); grp.each("scene_lattice", 512, 48)
    for _ in range(3):
        one.next_step(DT); grp.next_step(DT)
    one.sync(); grp.sync()
    for f, n in FIELDS:
        a, b = grp.read_field(f), one.read_field(f)
        assert bits_equal(a, b), f"world={world} {n}: " + first_mismatch(a, b, H)
    assert np.array_equal(grp.read_field(tfs.FIELD_M), one.read_field(tfs.FIELD_M))
    one.close(); grp.close()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("W,H,K,g", [(256, 64, 4, 0.0), (520, 96, 50, 9.8), (1024, 256, 41, 0.0)])
def test_redblack_slabs_match_oracle_bit_exact(world, W, H, K, g):
    """red-black mode on x-slabs: every pass exchanges the 8 boundary columns of u and v with the neighbours"""
    if _ndev() < world:
        pytest.skip(f"needs {world} GPUs")
    if (W + 1 + 31) // 32 < world:
        pytest.skip("grid too narrow for this many slabs")
    grp = tfs.SlabGroup(W, H, tfs.SimConfig(gravity=g), world=world, iterations=K, mode=tfs.MODE_RED_BLACK)
    o = OracleSim(W, H, gravity=g, iterations=K, mode=1)
    thr = frac_to_threshold(0.05)
    grp.each("scene_random", 77, thr); o.L.orc_scene_random(o.h, 77, thr)
    grp.each("fill_random_velocity", 5, 60.0); o.fill_random_velocity(5, 60.0)
    for st in range(3):
        grp.next_step(DT); o.step(DT)
        grp.sync()
        for f, n in FIELDS:
            a, b = grp.read_field(f), getattr(o, n)
            assert bits_equal(a, b), f"world={world} step {st} {n}: " + first_mismatch(a, b, H)
    for _ in range(4):                                        # back to back, ordered by the device-side flags only
        grp.next_step(DT); o.step(DT)
    grp.sync()
    for f, n in FIELDS:
        assert bits_equal(grp.read_field(f), getattr(o, n)), n
    assert np.array_equal(grp.checksum(), tfs.checksum_of_fields(o.u, o.v, o.p, o.s))
    grp.close()


def test_slab_handle_rejects_what_it_cannot_do():
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    with pytest.raises(tfs.TfsError):
        tfs.FluidSim(256, 66, rank=0, world=2)              # H % 4 != 0
    with pytest.raises(tfs.TfsError):
        tfs.FluidSim(40, 64, rank=0, world=4)               # fewer bands than slabs
    s = tfs.FluidSim(256, 64, rank=0, world=2)
    with pytest.raises(tfs.TfsError):
        s.next_step(DT)                                     # not attached yet
    with pytest.raises(tfs.TfsError):
        s.resize(128, 64)
    with pytest.raises(tfs.TfsError):
        s.set_options(kernel=tfs.KERNEL_DIAGONAL)           # would index slab-local buffers with global columns
    with pytest.raises(tfs.TfsError):
        s.set_options(iterations=0)
    with pytest.raises(tfs.TfsError):
        tfs.FluidSim(256, 64, rank=0, world=2, kernel=tfs.KERNEL_DIAGONAL)
    s.close()
