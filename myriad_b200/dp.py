"""Data-parallel plumbing of the hot path (one process per GPU): rank discovery from the torchrun
environment (reference: minigpt4/common/dist_utils.py:57-72), DistributedSampler-style sharding of a dataset
(runner_base.py:533-539) and the one collective of a training step — the all-reduce of the flat gradient buffer of the
trainable parameters (DDP with find_unused_parameters=True, runner_base.py:96-98: parameters untouched by a step
contribute zeros). Pure host logic: no CUDA needed (tests run it on 2 gloo ranks)."""
import os

import torch
import torch.distributed as dist


def dp_env():
    """(rank, world_size, local_rank) as torchrun exports them; (0, 1, 0) for a plain python launch."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend, device=None):
    """Rendezvous over MASTER_ADDR / MASTER_PORT from the environment (127.0.0.1 on a single node). No-op for world 1."""
    rank, world, _ = dp_env()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (device is not None and backend == "nccl") else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def shard_indices(n, rank, world, seed=None, drop_last=False):
    """Indices of `rank`'s shard of a dataset of n samples, torch.utils.data.DistributedSampler semantics: optional seeded
    shuffle shared by all ranks, padding by wrap-around so every rank draws the same number of samples."""
    if seed is None:
        idx = list(range(n))
    else:
        g = torch.Generator()
        g.manual_seed(seed)
        idx = torch.randperm(n, generator=g).tolist()
    if drop_last:
        per = n // world
        idx = idx[:per * world]
    else:
        per = (n + world - 1) // world
        pad = per * world - len(idx)
        if pad:
            idx += (idx * ((pad + len(idx) - 1) // len(idx)))[:pad]
    return idx[rank::world]


def allreduce_flat_grads(flat, world=None):
    """Sum the flat fp32 gradient buffer over all ranks in place (NCCL over NVLink on the GPU box, gloo in CPU tests) and
    return the factor the optimizer folds into its gradient unscale to turn the sum into DDP's mean."""
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return 1.0 / world


def max_over_ranks(value, device="cpu"):
    """Max of a python float over ranks (timing rule: a multi-GPU step takes as long as its slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class FlatGradDataParallel(torch.nn.Module):
    """Data-parallel wrapper the runner puts around the drop-in model when `run.distributed` (the reference wraps it in torch
    DDP with find_unused_parameters=True, runner_base.py:93-98). Exposes `.module` like DDP. It installs no per-parameter
    hooks: the model's fused backward (minigpt4/models/train_step.py) sees `dp_world > 1` and averages the ONE flat fp32
    gradient buffer of all trainable parameters across ranks (NCCL) before the nn.Parameter gradients are materialised, so
    parameters a rank's randomly drawn stage did not touch contribute zeros to the mean, exactly like DDP's unused-parameter
    handling."""

    def __init__(self, module):
        super().__init__()
        self.module = module
        module.dp_world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    @property
    def device(self):
        return self.module.device


def mean_flat_grads_(flat, world):
    """In-place mean over ranks of a flat gradient buffer."""
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.mul_(1.0 / world)
    return flat
