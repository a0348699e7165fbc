#!/usr/bin/env python
"""Benchmark of the Myriad hot path on B200 (BASELINE.json metric: images/sec, forward + greedy decode).

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference ...                      the reference's CPU path (oracle port) on the host cores

Workload (config.workload = "myriad_generate_b4"; BASELINE.json configs[2]): full Myriad — EVA-ViT-g (39 blocks) +
LoraAdaptorV2/ln_vision + VEInstructor + Q-Former (81 queries) + llama_proj + VETokenizer + Vicuna-7B with LoRA r=8 —
Myriad.generate on a batch of 4 synthetic 224x224 images + anomaly maps per GPU with a 32-token prompt (prefill
S = 131) and 32 greedy new tokens. One step = one generate() call. Weights are seeded synthetic (no checkpoints exist
offline). Data-parallel replicas: no collective on the inference path => weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 4
NEW_TOKENS = 32
METRIC = "images/sec (fwd+greedy-decode)"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return {"hbm": float(p["hbm_gbs"]), "tensor": float(p["bf16_tflops"]), "tensor_sustained": float(p["bf16_tflops_sustained"]),
                "src": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm": 6650.0, "tensor": 1590.0, "tensor_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from myriad_b200 import kernels as K
    from myriad_b200 import synthetic as syn
    from myriad_b200.engine import MyriadEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dims = syn.full_dims(lora_r=8)
    t0 = time.time()
    eng = MyriadEngine(syn.LazyStateDict(dims, seed=0, device=dev), dims, device=dev, max_batch=BATCH, max_seq=256)
    torch.cuda.synchronize()
    t_load = time.time() - t0
    image, maps = syn.make_inputs(BATCH, seed=1234 + rank, device="cpu")
    image_h, maps_h = image.pin_memory(), maps.pin_memory()
    image_d, maps_d = image_h.to(dev), maps_h.to(dev)
    ids_b, ids_a = syn.make_prompt_ids(dims.llama.vocab)
    stops = ((835,), (2277, 29937))

    def step_resident():
        return eng.generate(image_d, maps_d, ids_b, ids_a, max_new_tokens=NEW_TOKENS, stop_seqs=stops)

    def step_e2e():
        # the call a user makes: host batch in (pinned), token ids out on the host
        im = image_h.to(dev, non_blocking=True)
        mp = maps_h.to(dev, non_blocking=True)
        return eng.generate(im, mp, ids_b, ids_a, max_new_tokens=NEW_TOKENS, stop_seqs=stops)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = K.launch_count()
        e0.record()
        ntok = 0
        for _ in range(steps):
            ntok += fn().numel()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ntok, K.launch_count() - n0

    if args.profile:  # under ncu: one warm-up (graph capture), one step, nothing else
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()  # ncu --profile-from-start off: only the one step below is captured
        ms, ntok, launches = timed(step_resident, 1)
        torch.cuda.profiler.stop()
        print(json.dumps({"profile_run": True, "ms_per_step_under_profiler": ms, "gpu_launches": launches}), flush=True)
        return
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, ntok, launches = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)

    value = world * BATCH * args.steps / (ms / 1e3)
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1e3)
    train = None
    if not args.no_train:
        train = run_train_leg(args, dims, dev, world, rank, barrier)
    peaks = _peaks()
    extra = {}
    roof = None
    if rank == 0:
        roof, extra = roofline_probe(eng, dims, dev, peaks)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_baseline import CpuSample
        torch.set_num_threads(os.cpu_count() or 1)
        cs = CpuSample()
        v, parts = cs.run(NEW_TOKENS)
        cpu = {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port", "sample": cs.describe(NEW_TOKENS),
               "parts_s": {k: round(x, 4) for k, x in parts.items()}}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 (fp32 accumulate / residual / softmax)", "data": "synthetic",
            "config": {"workload": "myriad_generate_b4", "model": "Myriad: EVA-ViT-g + Q-Former(81q) + Vicuna-7B LoRA r=8",
                       "batch_per_gpu": BATCH, "global_batch": BATCH * world, "prompt_tokens": 32, "prefill_len": 131,
                       "new_tokens": NEW_TOKENS, "parallelism": "dp%d replicas, no data-path collective" % world,
                       "l2": "inputs larger than L2: every step streams 13.5 GB of weights through a 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": image_h.numel() * 4 + maps_h.numel() * 4,
                    "d2h_bytes_per_step": BATCH * NEW_TOKENS * 4 + 8 * NEW_TOKENS, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "new_tokens_per_s": ntok / (ms / 1e3) * world, "weights_load_s": round(t_load, 1),
            "train": train,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train_leg(args, dims, dev, world, rank, barrier):
    """Second half of BASELINE.json's metric: train tokens/s. One step = Myriad.forward + backward + gradient all-reduce
    (NCCL, flat fp32 buffer of the trainable parameters) + fused AdamW on an MVTec-shaped batch per GPU (batch_size_train 4:
    2 normal + 2 augmented samples, runner_base.py:546-549; stage 1 = VEInstructor + VETokenizer both on the gradient path;
    LoRA r = 8 of loraadapter_simple_myriad_finetune.yaml). tokens = LLM sequence positions processed (B x L)."""
    import torch
    import torch.distributed as dist

    from myriad_b200 import synthetic as syn
    try:
        from myriad_b200.training import MyriadTrainer
        tr = MyriadTrainer(syn.LazyStateDict(dims, seed=0, device=dev), dims, device=dev, max_batch=BATCH, max_seq=256)
        image, maps = syn.make_inputs(BATCH, seed=4321 + rank, device="cpu")
        image_h, maps_h = image.pin_memory(), maps.pin_memory()
        ids_b, ids_a = syn.make_prompt_ids(dims.llama.vocab)
        g = torch.Generator().manual_seed(99 + rank)
        Lt = 32
        text = torch.randint(3, dims.llama.vocab, (BATCH, Lt), generator=g)
        tmask = torch.ones(BATCH, Lt, dtype=torch.long)
        text[:, 16:] = dims.llama.eos  # 16 answer tokens, right-padded with eos (= pad, myriad.py:182) to 32
        tmask[:, 16:] = 0

        def step():
            im, mp = image_h.to(dev, non_blocking=True), maps_h.to(dev, non_blocking=True)
            return tr.train_step(im, mp, 1, ids_b, ids_a, text, tmask)

        for _ in range(max(args.warmup, 3)):
            loss = step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = max(1, args.steps)
        e0.record()
        for _ in range(n):
            loss = step()
        loss_v = float(loss.item())  # device -> host read of the step's result inside the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        L = 1 + 6 + tr.num_image_tokens(1) + 26 + Lt
        out = {"metric": "train tokens/sec", "value": world * BATCH * L * n / (ms / 1e3), "unit": "tokens/s", "ms_per_step": ms / n,
               "images_per_s": world * BATCH * n / (ms / 1e3), "loss": loss_v,
               "config": {"workload": "myriad_stage2_lora_finetune_b4", "batch_per_gpu": BATCH, "seq_len": L, "stage": 1, "lora_r": 8,
                          "trainable_params": int(tr.flat_params.numel()), "allreduce_bytes_per_step": int(tr.flat_grads.numel()) * 4 if world > 1 else 0,
                          "optimizer": "fused AdamW (flat fp32 buffer)", "includes": "h2d of the batch, fwd, bwd, all-reduce, AdamW, loss d2h"}}
        del tr
        torch.cuda.empty_cache()
        return out
    except Exception as e:  # the inference line must still be printed
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}


def roofline_probe(eng, dims, dev, peaks):
    """Dominant kernel = the small-batch weight-streaming kernel (csrc/gemv.cu, `gemv_kernel`): ~75 % of a step is the 32
    decode steps, each streaming every LLaMA weight once through 4 launches per layer. Timed live with CUDA events over the
    real 32 layers' weights in layer order and in the decode step's own launch configuration (RMSNorm hand-over, in-place
    fp32 residual, SwiGLU epilogue; 13 GB >> L2, so nothing is re-served from cache). Also reports the tensor-bound
    ViT GEMM for context."""
    import torch

    from myriad_b200 import kernels as K
    l = dims.llama
    B = BATCH
    x = torch.randn(B, l.hidden, device=dev).half()
    a = torch.randn(B, l.inter, device=dev).half()
    wq = eng.llw.layers[0].wqkv.shape[0]  # 3 * hidden + 2 * lora_r (LoRA A rows ride along)
    qkv = torch.empty(B, wq, device=dev, dtype=torch.float16)
    o = torch.zeros(B, l.hidden, device=dev, dtype=torch.float32)
    act = torch.empty(B, l.inter, device=dev, dtype=torch.float16)
    fused = eng.fuse_small_batch_norm and B <= 4

    ya, yb = (torch.zeros(B, l.hidden, device=dev, dtype=torch.float16) for _ in range(2))
    ssa, ssb = (torch.zeros(K.NORM_SS_FLOATS, device=dev) for _ in range(2))

    def sweep():
        for L in eng.llw.layers:
            if fused:  # the decode step's own launch configuration: RMSNorm handed over between the projections
                K.gemm(yb, L.wqkv, out=qkv, w_static=True, norm_ss=(ssb, l.eps))
                K.gemm(x, L.wo, res=o, out=o, w_static=True, post_norm=(L.n2, ya, ssa))
                K.gemm(ya, L.wgu, act=K.ACT_SWIGLU, out=act, w_static=True, norm_ss=(ssa, l.eps))
                K.gemm(a, L.wd, res=o, out=o, w_static=True, post_norm=(L.n1, yb, ssb))
            else:
                K.gemm(x, L.wqkv, out=qkv, w_static=True)
                K.gemm(x, L.wo, res=o, out=o, w_static=True)
                K.gemm(x, L.wgu, act=K.ACT_SWIGLU, out=act, w_static=True)
                K.gemm(a, L.wd, res=o, out=o, w_static=True)

    for _ in range(2):
        sweep()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    o.zero_()
    e0.record()
    for _ in range(reps):
        sweep()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n_launch = 4 * l.layers
    wbytes = l.layers * 2 * (wq * l.hidden + l.hidden * l.hidden + 2 * l.inter * l.hidden + l.hidden * l.inter)
    # activations per layer: fp16 rows in (3 x hidden + inter), fp32 residual read + written by o / down, their fp16 hand-over
    # rows + gamma, qkv / act written in fp16
    abytes = l.layers * (B * 2 * (3 * l.hidden + l.inter) + 2 * 2 * B * 4 * l.hidden + 2 * (B * 2 * l.hidden + 4 * l.hidden) +
                         B * 2 * wq + B * 2 * l.inter)
    per_launch = (wbytes + abytes) / n_launch
    achieved = (wbytes + abytes) / (ms / 1e3) / 1e9
    kname = "gemv_kernel" if fused else "gemm_tc_kernel"
    roof = {"kernel": "%s (decode, T=%d: norm+qkv+loraA / o+res / norm+gate_up+swiglu / down+res of all 32 layers)" % (kname, B), "bound": "hbm",
            "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s", "frac": achieved / peaks["hbm"],
            # dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the four launches of a layer, from the committed
            # `ncu --set full` capture of these launches (profiles/r1_ncu_full_pass3.md rows 0-3: 90.38 + 3.66, 100.87 + 3.67,
            # 33.70 + 0.00, 180.43 + 3.41 MB); not re-measured by this run
            "traffic": 104.03e6 if fused else None,
            "peak_source": peaks["src"], "avg_launch_us": ms * 1e3 / n_launch, "algorithmic_bytes_per_launch": per_launch}
    # tensor-bound context: the ViT MLP GEMMs at the bench batch (T = B * 257)
    T, D, Hd = B * dims.vit.tokens, dims.vit.dim, dims.vit.mlp_hidden
    h = torch.randn(T, D, device=dev).half()
    m = torch.empty(T, Hd, device=dev, dtype=torch.float16)
    blk = eng.vitw.blocks
    for b in blk[:3]:
        K.gemm(h, b.fc1w, bias=b.fc1b, act=K.ACT_GELU, out=m)
    torch.cuda.synchronize()
    e0.record()
    for b in blk:
        K.gemm(h, b.fc1w, bias=b.fc1b, act=K.ACT_GELU, out=m)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / len(blk)
    tf = 2.0 * T * D * Hd / (ms2 / 1e3) / 1e12
    # attention context (north_star: tensor-pipe share of the attention kernels): the ViT flash-attention launch at the bench batch
    v = dims.vit
    qkv_a = torch.randn(T, 3 * D, device=dev).half()
    ctx_a = torch.empty(T, D, device=dev, dtype=torch.float16)
    sa = (3 * D, v.tokens * 3 * D, v.head_dim)

    def attn_once():
        K.attention(qkv_a, qkv_a[:, D:], qkv_a[:, 2 * D:], ctx_a, B, v.heads, v.tokens, v.tokens, v.head_dim, 1.0, sa, sa, sa,
                    (D, v.tokens * D, v.head_dim))

    for _ in range(3):
        attn_once()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        attn_once()
    e1.record()
    torch.cuda.synchronize()
    ms3 = e0.elapsed_time(e1) / 20
    tf_a = 4.0 * B * v.heads * v.tokens * v.tokens * v.head_dim / (ms3 / 1e3) / 1e12
    extra_attn = {"kernel": "attn_fwd_kernel (ViT, B=%d H=%d N=%d dh=%d)" % (B, v.heads, v.tokens, v.head_dim), "bound": "tensor",
                  "achieved": tf_a, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": tf_a / peaks["tensor"], "avg_launch_us": ms3 * 1e3,
                  "note": "latency-bound at this size (1.5 GFLOP per launch, 1 % of the step); tensor pipe 7 % active in profiles/"}
    extra = {"roofline_attention": extra_attn,
             "roofline_tensor": {"kernel": "gemm_tc_kernel (ViT fc1+GELU, T=%d F=%d K=%d)" % (T, Hd, D), "bound": "tensor",
                                 "achieved": tf, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": tf / peaks["tensor"],
                                 "avg_launch_us": ms2 * 1e3}}
    return roof, extra


def run_reference(args):
    """Reference arm for this tier: the reference's own CPU implementation of the path, i.e. the fp32 oracle port
    (the reference module files cannot travel to the GPU box), with all host threads, on a bounded sample per step."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.cpu_baseline import CpuSample
    torch.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every host core
    cs = CpuSample()
    vals = []
    steps = max(1, min(args.steps, 3))
    for _ in range(min(args.warmup, 1)):
        cs.run(NEW_TOKENS)
    for _ in range(steps):
        v, parts = cs.run(NEW_TOKENS)
        vals.append(v)
    v = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * BATCH / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32 (CPU)", "data": "synthetic",
            "config": {"workload": "myriad_generate_b4", "batch_per_gpu": BATCH, "prompt_tokens": 32, "prefill_len": 131,
                       "new_tokens": NEW_TOKENS},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": cs.describe(NEW_TOKENS)},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg (train tokens/s)")
    ap.add_argument("--profile", action="store_true", help="minimal run for ncu (numbers printed under a profiler are not bench values)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
