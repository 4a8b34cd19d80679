"""N > 1 host logic of bench.py on CPU: two gloo ranks shard the synthetic page set with no
overlap, and the job time is the max over ranks (world_size 2, 127.0.0.1 rendezvous)."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    import bench

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bench.rank_page_indices(rank, 8)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ms = bench.max_over_ranks(100.0 + 50.0 * rank, world, "cpu")
    dist.barrier()
    q.put((rank, gathered, ms, bench.job_throughput(8, world, 3, ms)))
    dist.destroy_process_group()


def test_two_ranks_shard_pages_and_take_max_time():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gathered, ms, val in out:
        flat = [i for shard in gathered for i in shard]
        assert sorted(flat) == list(range(16)) and len(set(flat)) == 16  # disjoint cover, no exchange
        assert ms == 150.0                                               # max over ranks
        assert abs(val - 8 * 2 * 3 / 0.150) < 1e-9
