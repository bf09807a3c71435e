"""CPU test of the N>1 path: 2 gloo ranks shard a batch i::world, compute locally, and the single per-step
all_gather restores global order (world_size 2, 127.0.0.1 rendezvous)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mm_interleaved_b200.distributed import gather_results, shard_range
    mine = shard_range(n_items, rank, world)
    data = torch.arange(n_items * 3, dtype=torch.float32).view(n_items, 3)
    local = data[mine] * 2.0 + 1.0                       # the "forward" of this rank's sequences
    full = gather_results(local, n_items, rank, world)
    q.put((rank, mine, full))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_items, world, port = 7, 2, _free_port()            # 7 % 2 != 0: ragged shards
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.arange(n_items * 3, dtype=torch.float32).view(n_items, 3) * 2.0 + 1.0
    owned = []
    for rank, mine, full in got:
        assert torch.equal(full, want)                    # every rank sees all results in global order
        owned += mine
    assert sorted(owned) == list(range(n_items))          # a partition: every sequence owned exactly once
