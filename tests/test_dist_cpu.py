"""CPU suite: host logic of the N>1 path with world_size-2 gloo: feature-tile ownership and the
communicator-id hand-off that GBRL.init_distributed performs through torch.distributed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def shard(n_tiles, rank, world):
    """Same arithmetic as prepare_workspace() in gbrl_b200/csrc/capi.cu: G_t tile groups x G_r row groups."""
    gt = min(world, n_tiles)
    while gt > 1 and world % gt != 0:
        gt -= 1
    gr = world // gt
    tg, rg = rank % gt, rank // gt
    return (n_tiles * tg) // gt, (n_tiles * (tg + 1)) // gt, gr, rg


def tile_range(n_tiles, rank, world):
    lo, hi, _, _ = shard(n_tiles, rank, world)
    return lo, hi


@pytest.mark.parametrize("n_tiles", [1, 2, 3, 4, 8, 13])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_work_ownership_partitions_all_tiles_and_chunks(n_tiles, world):
    """Every (tile, row-chunk) pair of a node is histogrammed by exactly one rank."""
    n_chunks = 11
    owner = {}
    for r in range(world):
        lo, hi, gr, rg = shard(n_tiles, r, world)
        assert 0 <= lo < hi <= n_tiles and 0 <= rg < gr
        for t in range(lo, hi):
            for c in range(rg, n_chunks, gr):
                assert (t, c) not in owner
                owner[(t, c)] = r
    assert len(owner) == n_tiles * n_chunks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # id hand-off: rank 0 owns the 128-byte id, everybody must end up with the same bytes
    buf = np.zeros(128, np.uint8)
    if rank == 0:
        buf[:] = np.arange(128, dtype=np.uint8) ^ 0x5A
    t = torch.from_numpy(buf)
    dist.broadcast(t, 0)
    # disjoint-slice all-reduce == all-gather: every rank fills only its tiles, integer sum is exact
    n_tiles, per_tile = 5, 7
    h = torch.zeros(n_tiles * per_tile, dtype=torch.int64)
    lo, hi = tile_range(n_tiles, rank, world)
    for tile in range(lo, hi):
        h[tile * per_tile:(tile + 1) * per_tile] = torch.arange(per_tile) + 1000 * tile - 3
    dist.all_reduce(h)
    exp = torch.cat([torch.arange(per_tile) + 1000 * tile - 3 for tile in range(n_tiles)])
    ok = bool(torch.equal(h, exp)) and bool((t.numpy() == (np.arange(128, dtype=np.uint8) ^ 0x5A)).all())
    out.put((rank, ok))
    dist.destroy_process_group()


def test_gloo_world2_id_broadcast_and_disjoint_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
