"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py -- contiguous batch shards that tile the global
batch exactly once, rank-distinct inputs with rank-identical weights seeds, and max-over-ranks timing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench_workloads as B
    lo, hi = B.shard_range(16, rank, world)
    owned = torch.zeros(16, dtype=torch.int64)
    owned[lo:hi] = 1
    dist.all_reduce(owned)                                   # every unit owned exactly once
    ms = B.max_over_ranks(10.0 + 5.0 * rank, dist)           # slowest rank defines the step
    # inputs differ per rank, weights (seed 0) do not
    inp = B.msda_encoder_inputs(torch, 1, torch.device("cpu"), 1234 + rank, shapes_l=[(4, 4), (2, 2)])[0]
    w = torch.randn(8, generator=torch.Generator().manual_seed(0))
    gathered_in = [torch.zeros_like(inp) for _ in range(world)]
    gathered_w = [torch.zeros_like(w) for _ in range(world)]
    dist.all_gather(gathered_in, inp)
    dist.all_gather(gathered_w, w)
    q.put((rank, owned.tolist(), ms, bool(torch.equal(gathered_in[0], gathered_in[1])),
           bool(torch.equal(gathered_w[0], gathered_w[1]))))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_reduction():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, owned, ms, same_inputs, same_weights in res:
        assert owned == [1] * 16
        assert ms == 15.0
        assert not same_inputs and same_weights


def test_shard_range_rejects_ragged_batches():
    import bench_workloads as B
    import pytest
    assert B.shard_range(16, 3, 8) == (6, 8)
    with pytest.raises(ValueError):
        B.shard_range(10, 0, 4)
