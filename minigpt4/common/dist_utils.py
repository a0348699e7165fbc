"""Process-group helpers with the reference's interface (minigpt4/common/dist_utils.py:17-137): one process per GPU
launched by torchrun, rank discovery from env://, NCCL backend (gloo when no GPU is present, for the CPU tests)."""
import datetime
import functools
import os

import torch
import torch.distributed as dist


def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process():
    return get_rank() == 0


def setup_for_distributed(is_master):
    import builtins
    builtin_print = builtins.print

    def print(*args, **kwargs):
        if is_master or kwargs.pop("force", False):
            builtin_print(*args, **kwargs)

    builtins.print = print


def init_distributed_mode(args):
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        args.rank, args.world_size = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        args.gpu = int(os.environ.get("LOCAL_RANK", 0))
    else:
        print("Not using distributed mode")
        args.distributed = False
        return
    args.distributed = True
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(args.gpu)
    args.dist_backend = "nccl" if use_cuda else "gloo"
    dist.init_process_group(backend=args.dist_backend, init_method=getattr(args, "dist_url", "env://"),
                            world_size=args.world_size, rank=args.rank, timeout=datetime.timedelta(days=365))
    dist.barrier()
    setup_for_distributed(args.rank == 0)


def get_dist_info():
    return get_rank(), get_world_size()


def main_process(func):
    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        if get_rank() == 0:
            return func(*args, **kwargs)
    return wrapper


def download_cached_file(url, check_hash=True, progress=False):
    raise RuntimeError("no network in this environment: place %s on disk and pass its path instead" % url)
