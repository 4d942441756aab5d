"""world_size-2 `gloo` test (CPU) of the multi-GPU plumbing: round-robin sharding of independent systems,
one all_gather of the solutions, max-over-ranks timing.  The solver itself is replaced by a stub here (the real
one needs a GPU and is covered by `-m gpu` tests); what is tested is the host logic bench.py runs at N > 1."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from russell_b200 import batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nsys, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = batch.shard(nsys, world, rank)
        # stub "solver": system i is diag(i+2) x = ones  ->  x = 1/(i+2)
        xs = torch.stack([torch.full((n,), 1.0 / (i + 2), dtype=torch.float64) for i in mine]) if mine else torch.zeros((0, n), dtype=torch.float64)
        allx = batch.gather_solutions(xs, nsys, world, rank)
        g = batch.SolutionGatherer(nsys, world, rank, n)  # the preallocated path bench.py uses
        for rep in range(2):  # buffers are reused from step to step
            g.block.zero_()
            if mine:
                g.block[: len(mine)] = xs
            assert torch.equal(g.gather(), allx)
        tmax = batch.max_over_ranks(10.0 + rank, world)
        np.save(os.path.join(out_dir, "x_%d.npy" % rank), allx.numpy())
        np.save(os.path.join(out_dir, "t_%d.npy" % rank), np.array([tmax]))
    finally:
        dist.destroy_process_group()


def test_shard_is_round_robin_and_complete():
    for nsys in (1, 2, 5, 8, 13):
        for world in (1, 2, 4, 8):
            seen = sorted(i for r in range(world) for i in batch.shard(nsys, world, r))
            assert seen == list(range(nsys))
            assert all(i % world == r for r in range(world) for i in batch.shard(nsys, world, r))
            assert batch.slots_per_rank(nsys, world) * world >= nsys


def test_gather_world_1():
    x = torch.arange(6, dtype=torch.float64).reshape(2, 3)
    assert torch.equal(batch.gather_solutions(x, 2, 1, 0), x)
    g = batch.SolutionGatherer(2, 1, 0, 3)
    g.block[:2] = x
    assert torch.equal(g.gather(), x)
    assert batch.max_over_ranks(3.5, 1) == 3.5


def test_two_ranks_gloo(tmp_path):
    world, nsys, n = 2, 5, 7  # uneven: rank 0 owns 3 systems, rank 1 owns 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nsys, n, str(tmp_path)), nprocs=world, join=True)
    want = np.stack([np.full(n, 1.0 / (i + 2)) for i in range(nsys)])
    for r in range(world):
        got = np.load(tmp_path / ("x_%d.npy" % r))
        assert np.array_equal(got, want)
        assert np.load(tmp_path / ("t_%d.npy" % r))[0] == 11.0  # max over ranks of 10 + rank
