"""world_size-2 gloo tests (CPU) of the sharding / exchange host logic used by MultiRobotSampler in lock-step mode."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmd_b200.sampler import gather_peers, shard_robots


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_robots(n_total, world, rank)
    local = torch.stack([torch.full((64, 2), float(r)) + torch.arange(64)[:, None] * 1e-3 for r in range(lo, hi)]) \
        if hi > lo else torch.zeros(0, 64, 2)
    out = gather_peers(local, n_total)
    q.put((rank, lo, hi, out))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(n_total, port=None):
    port = port or _free_port()   # a fixed port collides with leftovers of earlier runs on a busy box
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    return sorted(res, key=lambda t: t[0])


def test_shard_robots_partitions_contiguously():
    for n, w in [(32, 8), (32, 2), (5, 2), (3, 4)]:
        blocks = [shard_robots(n, w, r) for r in range(w)]
        flat = [i for lo, hi in blocks for i in range(lo, hi)]
        assert flat == list(range(n))


def test_all_gather_of_representative_paths_even():
    res = _run(8)
    want = torch.stack([torch.full((64, 2), float(r)) + torch.arange(64)[:, None] * 1e-3 for r in range(8)])
    for rank, lo, hi, out in res:
        assert (lo, hi) == (rank * 4, rank * 4 + 4)
        assert torch.equal(out, want)


def test_all_gather_of_representative_paths_ragged():
    res = _run(5)
    want = torch.stack([torch.full((64, 2), float(r)) + torch.arange(64)[:, None] * 1e-3 for r in range(5)])
    for rank, lo, hi, out in res:
        assert torch.equal(out, want)
