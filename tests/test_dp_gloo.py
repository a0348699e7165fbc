"""CPU, world_size = 2 over gloo: the data-parallel host logic of the path (myriad_b200/dp.py) — sharding, the gradient
all-reduce of the flat buffer (runner_base.py:96-98 semantics) and the max-over-ranks timing rule."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from myriad_b200 import dp
    r, w = dp.init_process_group("gloo")
    assert (r, w) == (rank, world) and dp.dp_env() == (rank, world, rank)
    # per-rank gradients of a "model" whose second half is untouched on rank 1 (find_unused_parameters: zeros)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(1000, generator=g)
    if rank == 1:
        flat[500:] = 0
    mine = flat.clone()
    scale = dp.allreduce_flat_grads(flat)
    t = dp.max_over_ranks(1.0 + rank)
    out.put((rank, mine, flat * scale, t, dp.shard_indices(11, rank, world, seed=3)))
    dist.destroy_process_group()


def _run_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    return res


def test_two_rank_gradient_allreduce_and_sharding():
    # the rendezvous port is picked by binding to port 0 and releasing it: another process can take it in between, so a
    # failed rendezvous is retried on a fresh port (the assertions on the results below are never retried)
    res = None
    for attempt in range(3):
        try:
            res = _run_two_ranks()
            break
        except Exception:
            if attempt == 2:
                raise
    (_, g0, avg0, t0, s0), (_, g1, avg1, t1, s1) = res
    assert torch.allclose(avg0, (g0 + g1) / 2) and torch.equal(avg0, avg1), "every rank must hold the mean gradient"
    assert torch.allclose(avg0[500:], g0[500:] / 2), "a parameter unused on one rank averages with zeros, as DDP does"
    assert t0 == t1 == 2.0, "step time = slowest rank"
    assert len(s0) == len(s1) == 6 and set(s0) | set(s1) == set(range(11)), "shards cover the dataset, equal sizes (wrap-around pad)"


def test_shard_indices_match_distributed_sampler():
    from torch.utils.data import DistributedSampler
    from myriad_b200 import dp
    data = list(range(23))
    for world in (2, 4, 8):
        for rank in range(world):
            s = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=True, seed=5)
            s.set_epoch(0)
            assert list(s) == dp.shard_indices(len(data), rank, world, seed=5)
            s = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=False)
            assert list(s) == dp.shard_indices(len(data), rank, world)
