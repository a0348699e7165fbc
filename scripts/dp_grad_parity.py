"""Data-parallel gradient parity ON HARDWARE (SURVEY.md §4 "distributed"): launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_grad_parity.py
Each rank runs MyriadTrainer.forward_backward on ITS shard and the trainer's gradient all-reduce (NCCL, the overlapped tail +
the head of the flat buffer, exactly what optimizer_step does) and compares the averaged gradients with
  (a) the single-GPU gradients of the CONCATENATED batch (every rank computes them itself), same stage on all ranks;
  (b) the mean of the per-rank gradients when the ranks drew DIFFERENT stages (rank r: stage (1, 2, 0, ...)[r]) — DDP with
      find_unused_parameters=True (runner_base.py:96-98): parameters a rank's stage did not touch contribute zeros.
Prints one line per case; exit code 1 on mismatch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from myriad_b200 import synthetic as syn
from myriad_b200.dp import allreduce_flat_grads
from myriad_b200.training import MyriadTrainer


def reduced_grads(tr, world):
    early, tr._early = tr._early, None
    if early is not None:
        split, work = early
        allreduce_flat_grads(tr.flat_grads[:split])
        work.wait()
    else:
        allreduce_flat_grads(tr.flat_grads)
    torch.cuda.synchronize()
    return tr.flat_grads.clone() / world


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    d = syn.mid_dims(lora_r=8)
    sd = syn.make_state_dict(d, 0)
    Bp = 2
    ids_b, ids_a = syn.make_prompt_ids(d.llama.vocab)
    gen = torch.Generator().manual_seed(5)
    text_all = torch.randint(3, d.llama.vocab, (world * Bp, 8), generator=gen)
    tmask_all = torch.ones(world * Bp, 8, dtype=torch.long)  # equal target counts per rank: mean of rank means == global mean
    image_all, maps_all = syn.make_inputs(world * Bp, seed=21)
    sl = slice(rank * Bp, (rank + 1) * Bp)
    tr = MyriadTrainer(sd, d, device=dev, max_batch=world * Bp, max_seq=256)
    ok = True

    def rel(a, b):
        return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()

    # (a) same stage everywhere vs the concatenated batch on one GPU
    tr.forward_backward(image_all[sl].to(dev), maps_all[sl].to(dev), 1, ids_b, ids_a, text_all[sl], tmask_all[sl])
    got = reduced_grads(tr, world)
    tr.overlap_allreduce = False
    tr.forward_backward(image_all.to(dev), maps_all.to(dev), 1, ids_b, ids_a, text_all, tmask_all)
    want = tr.flat_grads.clone()
    e = rel(got, want)
    worst = max(((rel(got[o:o + n], want[o:o + n]), k) for k, (o, shape, _) in tr.segments.items()
                 for n in [int(torch.tensor(shape).prod())] if want[o:o + n].abs().max() > 0), default=(0.0, ""))
    if rank == 0:
        print("DP x%d same stage: mean of per-rank grads vs single-GPU grads of the concatenated batch: rel err %.2e overall, worst segment "
              "%.2e (%s)" % (world, e, worst[0], worst[1]), flush=True)
    ok &= worst[0] < 5e-2  # fp16 operands: the two batches split the token dimension of the wgrad GEMMs differently
    # (b) a different stage on every rank vs the mean of the separately computed per-rank gradients
    stages = [(1, 2, 0)[r % 3] for r in range(world)]
    tr.overlap_allreduce = False
    want = torch.zeros_like(got)
    for r in range(world):
        s2 = slice(r * Bp, (r + 1) * Bp)
        tr.forward_backward(image_all[s2].to(dev), maps_all[s2].to(dev), stages[r], ids_b, ids_a, text_all[s2], tmask_all[s2])
        want += tr.flat_grads / world
    tr.overlap_allreduce = True  # the tail of the buffer is reduced asynchronously while the backward continues
    tr.forward_backward(image_all[sl].to(dev), maps_all[sl].to(dev), stages[rank], ids_b, ids_a, text_all[sl], tmask_all[sl])
    got = reduced_grads(tr, world)
    e = rel(got, want)
    o, shape, _ = tr.segments["VETokenizer.meta_net.15.weight"]
    n = int(torch.tensor(shape).prod())
    zero_on = [r for r in range(world) if stages[r] == 2]
    if rank == 0:
        print("DP x%d stages %s: all-reduced mean vs mean of per-rank grads: rel err %.2e (ranks %s contribute zeros to VETokenizer)"
              % (world, stages, e, zero_on), flush=True)
    ok &= e < 1e-5
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
