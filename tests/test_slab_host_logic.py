"""CPU tests of the host side of the multi-GPU path: the slab partition (pure arithmetic in the C
ABI) and the blob hand-off pattern bench.py uses, run with world_size 2 and 4 over gloo."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

import terminal_fluid_sim_b200 as tfs
from terminal_fluid_sim_b200 import build as tfs_build


@pytest.fixture(scope="module", autouse=True)
def _built():
    tfs_build.build()


@pytest.mark.parametrize("W", [64, 100, 1024, 4096, 32768, 33000])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_slab_bounds_tile_the_grid_on_band_boundaries(W, world):
    if (W + 1 + 31) // 32 < world:
        pytest.skip("fewer bands than slabs")
    b = [tfs.slab_bounds(W, r, world) for r in range(world)]
    assert b[0][0] == 0 and b[-1][1] == W
    for (lo, hi), (lo2, _) in zip(b, b[1:]):
        assert hi == lo2 and hi > lo
        assert (hi + 1) % 32 == 0          # a slab starts where a band starts (band b covers columns 32b-1 ..)
    widths = [hi - lo for lo, hi in b]
    assert max(widths) - min(widths) <= 64  # even split, up to band granularity


def _worker(rank, world, port, W, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = tfs.slab_bounds(W, rank, world)
    blob = f"blob-of-rank-{rank}:{lo}:{hi}".encode().ljust(512, b"\0")     # what peer_export returns: 512 opaque bytes
    blobs = [None] * world
    dist.all_gather_object(blobs, blob)
    left = blobs[rank - 1] if rank > 0 else None
    right = blobs[rank + 1] if rank < world - 1 else None
    ok = True
    if left is not None:
        ok &= left.rstrip(b"\0").decode().endswith(f":{lo}")               # my left neighbour ends where I start
    if right is not None:
        ok &= right.rstrip(b"\0").decode().split(":")[1] == str(hi)         # my right neighbour starts where I end
    ok &= (left is None) == (rank == 0) and (right is None) == (rank == world - 1)
    # rank 0 gathers the per-rank column counts the way tests concatenate slabs
    counts = [None] * world
    dist.all_gather_object(counts, hi - lo)
    ok &= sum(counts) == W
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_blob_exchange_pattern_over_gloo(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 4096, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
