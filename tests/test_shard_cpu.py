"""Host-side multi-GPU logic on CPU: round-robin assignment and the result gather, world_size 2 over gloo."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lemo_b200 import shard


def test_assign_partitions_all_sequences():
    for n, w in ((64, 8), (64, 1), (5, 2), (3, 4)):
        got = sorted(s for r in range(w) for s in shard.assign(n, w, r))
        assert got == list(range(n))
        assert all(shard.owner(s, w) == r for r in range(w) for s in shard.assign(n, w, r))
        sizes = [len(shard.assign(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n, T = 5, 4
    ids = shard.assign(n, world, rank)
    local = torch.stack([torch.full((T, 72), float(s)) for s in ids])
    out = shard.gather_results(ids, local, n, world, rank)
    if rank == 0:
        q.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_results_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.shape == (5, 4, 72)
    for s in range(5):
        assert np.all(out[s] == float(s))
