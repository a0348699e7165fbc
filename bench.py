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


QUESTION = ("<Img><ImageHere></Img> This image may be simulated by photo editing According to IAD expert opinions and corresponding visual "
            "descriptions find out if there are defects")


def _questions(tokenizer, n_after=26):
    """A question string whose wrapped prompt '###Human: <Img>' + image + '</Img> ... ###Assistant: ' tokenises to 6 + 26 = 32
    prompt tokens (north_star: 32-token prompts) under the tokenizer in use; words are dropped / repeated to hit the count."""
    words = QUESTION.split(" ")
    head, tail = words[0], words[1:]

    def n_tok(ws):
        text = "###Human: " + " ".join([head] + ws) + " ###Assistant: "
        after = text.split("<ImageHere>")[1]
        return tokenizer(after, return_tensors="pt", add_special_tokens=False).input_ids.shape[1]

    ws = list(tail)
    while n_tok(ws) > n_after and ws:
        ws.pop()
    while n_tok(ws) < n_after:
        ws.append("defects")
    assert n_tok(ws) == n_after, "cannot build a %d-token prompt tail" % n_after
    return " ".join([head] + ws)


def run_ours(args):
    import zlib

    import torch
    import torch.distributed as dist

    import minigpt4.models  # noqa: F401  (registers arch: myriad)
    from minigpt4.common.registry import registry
    from minigpt4.conversation.conversation import StoppingCriteriaSub
    from myriad_b200 import kernels as K
    from myriad_b200 import synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dims = syn.full_dims(lora_r=8)
    t0 = time.time()
    # the reference-facing plugin: the registry-registered class, built the way evaluation_aqa_dataset.py:255-256 builds it
    model = registry.get_model_class("myriad")(use_lora=True, llama_model="", max_txt_len=32, end_sym="###",
                                                weights=syn.LazyStateDict(dims, seed=0, device=dev)).to(dev).eval()
    eng = model.engine
    torch.cuda.synchronize()
    t_load = time.time() - t0
    image, maps = syn.make_inputs(BATCH, seed=1234 + rank, device="cpu")
    image_h, maps_h = image.pin_memory(), maps.pin_memory()
    image_d, maps_d = image_h.to(dev), maps_h.to(dev)
    question = _questions(model.llama_tokenizer)
    samples_h = {"image": image_h, "anomaly_maps": maps_h, "question": [question] * BATCH, "question2": [question] * BATCH,
                 "question3": [question] * BATCH, "scene": ["bottle"] * BATCH, "img_path": ["synthetic/%d.png" % i for i in range(BATCH)]}
    stops = ((835,), (2277, 29937))
    crit = [StoppingCriteriaSub(stops=[torch.tensor(s) for s in stops])]
    gen_kw = dict(max_new_tokens=NEW_TOKENS, stopping_criteria=crit, do_sample=True, top_p=0.01, temperature=1.0, min_length=1, use_cache=True)
    ids_b, ids_a = model._split_prompts(["###Human: " + question + " ###Assistant: "], dev)
    ids_b, ids_a = ids_b[0], ids_a[0]
    assert ids_b.numel() + ids_a.numel() == 32, (ids_b.numel(), ids_a.numel())

    def step_resident():
        return eng.generate(image_d, maps_d, ids_b, ids_a, max_new_tokens=NEW_TOKENS, stop_seqs=stops)

    def step_e2e():
        # the call a user of the reference makes (evaluation_aqa_dataset.py:333): host batch + question strings in, token ids out
        out = model.generate(samples_h, **gen_kw)
        return out["token_ids"].cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = K.launch_count()
        e0.record()
        ntok, last = 0, None
        for _ in range(steps):
            last = fn()
            ntok += last.numel()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ntok, K.launch_count() - n0, last

    if args.profile:  # under ncu: one warm-up (graph capture), one step, nothing else
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()  # ncu --profile-from-start off: only the one step below is captured
        ms, ntok, launches, _ = timed(step_resident, 1)
        torch.cuda.profiler.stop()
        print(json.dumps({"profile_run": True, "ms_per_step_under_profiler": ms, "gpu_launches": launches}), flush=True)
        return
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.decode_timing = []
    ms, ntok, launches, toks_timed = timed(step_resident, args.steps)
    decode_events, eng.decode_timing = eng.decode_timing, None
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e, _, _, toks_e2e = timed(step_e2e, args.steps)

    # correctness of the timed run: the graph-replayed decode must reproduce an eager (un-graphed) replay of the same launches
    # bit for bit, and the plugin call must produce the same ids as the engine call it wraps
    emb = eng.build_inputs_embeds(image_d, maps_d, 1, ids_b, ids_a)
    toks_eager = eng.greedy_decode(emb, NEW_TOKENS, stops, use_graph=False)
    assert toks_timed.tolist() == toks_eager.tolist(), "timed (CUDA-graph) tokens differ from the eager replay"
    assert toks_e2e.tolist() == toks_timed.tolist(), "plugin generate() tokens differ from the engine's"
    checksum = "%08x" % (zlib.crc32(toks_timed.numpy().astype("int64").tobytes()) & 0xFFFFFFFF)

    value = world * BATCH * args.steps / (ms / 1e3)
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1e3)
    peaks = _peaks()
    extra = {}
    roof = None
    if rank == 0:
        roof, extra = roofline_in_situ(eng, dims, dev, peaks, decode_events, toks_timed.shape[1])
    train = None
    if not args.no_train:
        train = run_train_leg(args, dims, dev, world, rank, barrier)
    cpu = eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_baseline import CpuSample
        torch.set_num_threads(os.cpu_count() or 1)
        cs = CpuSample()
        v, parts = cs.run(NEW_TOKENS)
        cpu = {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port", "sample": cs.describe(NEW_TOKENS),
               "parts_s": {k: round(x, 4) for k, x in parts.items()}}
        cpu.update(cs.extrapolation(NEW_TOKENS))
        del cs
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        try:
            from oracle.cpu_baseline import EagerGpuSample
            del model, eng
            torch.cuda.empty_cache()
            es = EagerGpuSample(dev)
            ev, ems, n_new = es.time(steps=3, warmup=1, new_tokens=NEW_TOKENS)
            eager = {"value": ev, "unit": "images/s", "ms_per_step": ems, "new_tokens": n_new,
                     "what": "the path being replaced: the same workload through plain PyTorch eager fp16 on this B200 (oracle port on "
                             ".half().cuda() tensors: cuBLAS / ATen kernels, torch.cat KV cache, per-step host sync of the HF greedy loop); "
                             "pinned host images in, token ids out — comparable with e2e"}
            del es
        except Exception as e:
            eager = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    expert = None
    if rank == 0 and world == 1 and not args.no_expert:
        try:
            expert = run_expert_leg(dev, peaks)
        except Exception as e:
            expert = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    if rank == 0:
        cfg = {"workload": "myriad_generate_b4", "model": "Myriad: EVA-ViT-g + Q-Former(81q) + Vicuna-7B LoRA r=8",
               "batch_per_gpu": BATCH, "global_batch": BATCH * world, "prompt_tokens": 32, "prefill_len": 131,
               "new_tokens": NEW_TOKENS, "parallelism": "dp%d replicas, no data-path collective" % world,
               "l2": "inputs larger than L2: every step streams 13.5 GB of weights through a 126 MB L2",
               "tokens_checksum": checksum, "tokens_verified": "timed CUDA-graph run == eager replay == plugin generate()"}
        if train and "value" in train:  # second half of BASELINE.json's metric, kept inside `config` so result parsers keep it
            cfg["train_tokens_per_s"] = train["value"]
            cfg["train_ms_per_step"] = train["ms_per_step"]
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 (fp32 accumulate / residual / softmax)", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": image_h.numel() * 4 + maps_h.numel() * 4,
                    "d2h_bytes_per_step": BATCH * NEW_TOKENS * 8 + 8 * NEW_TOKENS, "ms_per_step": ms_e2e / args.steps,
                    "api": "registry.get_model_class('myriad')(...).generate(samples) with pinned host tensors and question strings"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "gpu_eager_baseline": eager,
            "new_tokens_per_s": ntok / (ms / 1e3) * world, "weights_load_s": round(t_load, 1),
            "train": train, "vision_expert": expert,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_expert_leg(dev, peaks):
    """SURVEY.md §8 (f2), context next to the headline: the vision expert (adrefexpert_v2.py:245-301) at the benchmark batch - ImageBind-Huge
    vision trunk (32 blocks x 1280, four taps) + zero-shot and 1-shot map heads, the two calls Myriad makes per batch (myriad.py:342-343) with
    one trunk pass over the query images and the reference tokens cached. Device-resident inputs, CUDA events, 5 repetitions."""
    import torch

    from myriad_b200 import expert as X
    from myriad_b200 import synthetic as syn
    d = X.ExpertDims()
    sd = {k: syn.synth("vision_expert." + k, shape, std, 0, device=dev, mean=mean) for k, shape, std, mean in X.expert_state_dict_spec(d)}
    eng = X.VisionExpertEngine(sd, d, device=dev)
    del sd
    image, _ = syn.make_inputs(BATCH, seed=99, device="cpu")
    refs, _ = syn.make_inputs(BATCH, seed=98, device="cpu")
    image, refs = image.to(dev), refs.to(dev)
    text = X.make_text_features(BATCH, d).to(dev)
    refs_n = eng.encode_refs(refs, BATCH)
    for _ in range(2):
        eng.both(image, text, refs_n=refs_n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        (zm, _), (km, _) = eng.both(image, text, refs_n=refs_n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    T, D = BATCH * d.tokens, d.dim
    flop = (max(d.out_layers) + 1) * (2.0 * T * D * (3 * D + D + 2 * d.mlp_hidden) + 4.0 * BATCH * d.heads * d.tokens * d.tokens * d.head_dim)
    tf = flop / (ms / 1e3) / 1e12
    return {"what": "adrefexpert zero-shot + 1-shot maps for %d images: ImageBind-Huge trunk (one pass) + both heads, reference tokens cached" % BATCH,
            "ms_per_batch": ms, "images_per_s": BATCH / (ms / 1e3), "trunk_tflops": tf, "frac_of_tensor_peak": tf / peaks["tensor"],
            "maps_finite": bool(torch.isfinite(zm).all() and torch.isfinite(km).all())}


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


def _ncu_traffic(kernel="gemv_kernel"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, averaged over the launches in the committed
    `ncu --set full` capture profiles/r2_ncu_raw.csv (written by scripts/gpu_round.sh from the same decode-step launch
    configuration). None when the capture is absent."""
    import csv
    path = os.path.join(ROOT, "profiles", "r2_final_ncu_raw.csv")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r2_ncu_raw.csv")
    if not os.path.exists(path):
        return None, None
    try:
        rows = list(csv.reader(open(path)))
        hdr = rows[0]
        ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        unit_r, unit_w = rows[1][ir], rows[1][iw]
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = [float(r[ir]) * mult.get(unit_r, 1.0) + float(r[iw]) * mult.get(unit_w, 1.0) for r in rows[2:] if r[ik].startswith(kernel)]
        return (sum(vals) / len(vals), len(vals)) if vals else (None, 0)
    except Exception:
        return None, None


def roofline_in_situ(eng, dims, dev, peaks, decode_events, new_tokens):
    """Dominant kernel = the small-batch weight-streaming kernel (csrc/gemv.cu `gemv_kernel`): ~75 % of a step is the decode
    steps, each streaming every LLaMA weight once through 4 launches per layer + lm_head. Measured where the time is spent:
      * frac_step: (weight + KV-cache bytes of one decode step) / decode-step time, the step time taken with CUDA events
        around the graph replays INSIDE the timed region of the benchmark (engine.decode_timing);
      * frac (kernel): algorithmic bytes of the 129 gemv launches of one decode step / the sum of their durations inside
        the captured step, each launch timed by its own CTAs (%globaltimer: from the moment its predecessor released it,
        griddepcontrol.wait returning, to its last CTA done) — attention, arg-max and the launch boundaries are outside."""
    import ctypes

    import torch

    from myriad_b200 import kernels as K
    l = dims.llama
    B = BATCH
    wq = eng.llw.layers[0].wqkv.shape[0]  # 3 * hidden + 2 * lora_r (LoRA A rows ride along)
    w_layer = 2 * (wq * l.hidden + l.hidden * l.hidden + 2 * l.inter * l.hidden + l.hidden * l.inter)
    w_head = 2 * l.vocab * l.hidden
    # activations per layer: fp16 rows in (3 x hidden + inter), fp32 residual read + written by o / down, their fp16 hand-over
    # rows + gamma, qkv / act written in fp16
    a_layer = (B * 2 * (3 * l.hidden + l.inter) + 2 * 2 * B * 4 * l.hidden + 2 * (B * 2 * l.hidden + 4 * l.hidden) + B * 2 * wq + B * 2 * l.inter)
    gemv_bytes = l.layers * (w_layer + a_layer) + w_head + B * 2 * l.hidden + B * 4 * l.vocab
    n_gemv = 4 * l.layers + 1
    # ---- decode-step time inside the timed region
    torch.cuda.synchronize()
    tot_ms = sum(e0.elapsed_time(e1) for e0, e1, n in decode_events)
    tot_steps = sum(n for _, _, n in decode_events)
    step_ms = tot_ms / max(tot_steps, 1)
    S_mid = 131 + new_tokens // 2
    kv_bytes = 2 * l.layers * B * S_mid * l.hidden * 2  # K and V rows of the visible cache, fp16, mid-generation
    embed_bytes = B * l.hidden * 2
    step_bytes = l.layers * w_layer + w_head + kv_bytes + embed_bytes
    frac_step = step_bytes / (step_ms / 1e3) / 1e9 / peaks["hbm"] if step_ms > 0 else None
    # ---- per-launch durations from inside one captured decode step
    st = next(iter(eng._decode_graphs.values()))
    fused = st.graph is not None and eng.fuse_small_batch_norm and B <= 4 and st.mega is None
    kernel_us = kernel_us_after_wait = None
    if fused:
        n_slots = n_gemv + 8
        buf = torch.zeros(n_slots * 148 * 6, dtype=torch.int64, device=dev)
        snap = st.state.clone()
        K.lib().myr_gemm_set_trace(ctypes.c_void_p(buf.data_ptr()))
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                eng._decode_step(st)
        finally:
            K.lib().myr_gemm_set_trace(ctypes.c_void_p(0))
        st.state.copy_(snap)
        g.replay()
        g.replay()  # timestamps of the second replay (warm instruction cache) overwrite the first
        torch.cuda.synchronize()
        st.state.copy_(snap)
        t = buf.reshape(n_slots, 148, 6)[:n_gemv].cpu()
        live = t[:, :, 5] > 0
        # window of a launch: first CTA START (its weight ring is already being filled then, under the tail of the previous
        # kernel: programmatic dependent launch) to last CTA done. Windows of consecutive launches overlap by that prefetch
        # period, so their sum over-counts time: the fraction below is a lower bound of the kernel's own rate.
        durs, durs_wait = [], []
        for i in range(n_gemv):
            m = live[i]
            if m.any():
                durs.append((int(t[i, m, 5].max()) - int(t[i, m, 0].min())) / 1e3)
                durs_wait.append((int(t[i, m, 5].max()) - int(t[i, m, 1].min())) / 1e3)
        if len(durs) == n_gemv:
            kernel_us = sum(durs)
            kernel_us_after_wait = sum(durs_wait)
    traffic, n_cap = _ncu_traffic()
    if kernel_us:
        achieved = gemv_bytes / (kernel_us / 1e6) / 1e9
        avg_us = kernel_us / n_gemv
    else:  # tensor-core weight streaming path (MYR_GEMV=0 / MYR_MEGA=1): only the step-level figure is available
        achieved = step_bytes / (step_ms / 1e3) / 1e9
        avg_us = None
    roof = {"kernel": "gemv_kernel (decode, T=%d: norm+qkv+loraA / o+res / norm+gate_up+swiglu / down+res of all 32 layers + lm_head), timed "
                      "inside the captured decode step" % B, "bound": "hbm",
            "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s", "frac": achieved / peaks["hbm"],
            "frac_step": frac_step, "decode_step_ms": step_ms, "decode_steps_timed": tot_steps, "step_bytes": step_bytes,
            "traffic": traffic, "traffic_source": ("profiles/r2_final_ncu_raw.csv: mean dram read+write of %d gemv_kernel launches under ncu --set full" % n_cap)
            if traffic else None,
            "peak_source": peaks["src"], "avg_launch_us": avg_us, "launches_per_step": n_gemv,
            "avg_launch_us_after_dependency_wait": (kernel_us_after_wait / n_gemv) if kernel_us_after_wait else None,
            "algorithmic_bytes_per_launch": gemv_bytes / n_gemv}
    return roof, tensor_context(eng, dims, dev, peaks)


def tensor_context(eng, dims, dev, peaks):
    """Tensor-bound context kernels (north_star: tensor-pipe share of the GEMM / attention kernels): the ViT MLP GEMM and the ViT
    flash-attention launch at the bench batch, back to back over the 39 layers' own weights."""
    import torch

    from myriad_b200 import kernels as K
    B = BATCH
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    T, D, Hd = B * dims.vit.tokens, dims.vit.dim, dims.vit.mlp_hidden
    h = torch.randn(T, D, device=dev).half()
    m = torch.empty(T, Hd, device=dev, dtype=torch.float16)
    blk = eng.vitw.blocks
    for b in blk[:3]:
        K.gemm(h, b.fc1w, bias=b.fc1b, act=K.ACT_GELU, out=m)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()  # launches back to back on the device: a 10 us GEMM is shorter than the host's launch path
    with torch.cuda.graph(g):
        for b in blk:
            K.gemm(h, b.fc1w, bias=b.fc1b, act=K.ACT_GELU, out=m)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / len(blk)
    tf = 2.0 * T * D * Hd / (ms2 / 1e3) / 1e12
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
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        for _ in range(20):
            attn_once()
    g2.replay()
    torch.cuda.synchronize()
    e0.record()
    g2.replay()
    e1.record()
    torch.cuda.synchronize()
    ms3 = e0.elapsed_time(e1) / 20
    tf_a = 4.0 * B * v.heads * v.tokens * v.tokens * v.head_dim / (ms3 / 1e3) / 1e12
    return {"roofline_attention": {"kernel": "attn2_kernel (ViT, B=%d H=%d N=%d dh=%d)" % (B, v.heads, v.tokens, v.head_dim), "bound": "tensor",
                                   "achieved": tf_a, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": tf_a / peaks["tensor"],
                                   "avg_launch_us": ms3 * 1e3},
            "roofline_tensor": {"kernel": "ViT fc1+GELU GEMM (T=%d F=%d K=%d), 39 layers' weights back to back" % (T, Hd, D), "bound": "tensor",
                                "achieved": tf, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": tf / peaks["tensor"],
                                "avg_launch_us": ms2 * 1e3}}


def run_sweep(args):
    """BASELINE.json configs[4]: batch {1,4,16,32} x LLM sequence {256,1024,2048}: images/s (forward + greedy decode of NEW_TOKENS
    tokens) and prompt tokens/s, with the decode step's fraction of the HBM roofline, per GPU and whole-job (replicas). One JSON
    line per point; rank 0 also writes them to --sweep-out."""
    import torch
    import torch.distributed as dist

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
    peaks = _peaks()
    batches = [int(x) for x in args.sweep_batches.split(",")]
    seqs = [int(x) for x in args.sweep_seqs.split(",")]
    eng = MyriadEngine(syn.LazyStateDict(dims, seed=0, device=dev), dims, device=dev, max_batch=max(batches), max_seq=max(seqs) + NEW_TOKENS)
    l = dims.llama
    stops = ((835,), (2277, 29937))
    n_img = eng.num_image_tokens(1)
    wq = eng.llw.layers[0].wqkv.shape[0]
    w_step = l.layers * 2 * (wq * l.hidden + l.hidden * l.hidden + 3 * l.inter * l.hidden) + 2 * l.vocab * l.hidden
    out = []
    for B in batches:
        image, maps = syn.make_inputs(B, seed=1234 + rank, device="cpu")
        image_d, maps_d = image.to(dev), maps.to(dev)
        for S in seqs:
            ids_b, ids_a = syn.make_prompt_ids(l.vocab, n_before=6, n_after=S - 6 - n_img)

            def step():
                return eng.generate(image_d, maps_d, ids_b, ids_a, max_new_tokens=NEW_TOKENS, stop_seqs=stops)

            for _ in range(max(1, min(args.warmup, 2))):
                step()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            eng.decode_timing = []
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = max(1, min(args.steps, 3))
            e0.record()
            for _ in range(n):
                toks = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            ev, eng.decode_timing = eng.decode_timing, None
            dec_ms = sum(a.elapsed_time(b) for a, b, _ in ev) / max(1, sum(c for _, _, c in ev))
            kv = 2 * l.layers * B * (S + NEW_TOKENS // 2) * l.hidden * 2
            frac_step = (w_step + kv) / (dec_ms / 1e3) / 1e9 / peaks["hbm"] if dec_ms > 0 else None
            rec = {"sweep": True, "batch_per_gpu": B, "seq_len": S, "n_gpus": world, "new_tokens": int(toks.shape[1]), "ms_per_step": ms,
                   "images_per_s": world * B / (ms / 1e3), "prompt_tokens_per_s": world * B * S / (ms / 1e3),
                   "new_tokens_per_s": world * B * int(toks.shape[1]) / (ms / 1e3), "decode_step_ms": dec_ms,
                   "decode_frac_of_hbm_roofline": frac_step, "hbm_peak_gbs": peaks["hbm"]}
            out.append(rec)
            if rank == 0:
                print(json.dumps(rec), flush=True)
    if rank == 0 and args.sweep_out:
        with open(args.sweep_out, "w") as fh:
            json.dump(out, fh, indent=1)
    if world > 1:
        dist.destroy_process_group()


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
    steps = max(1, min(args.steps, 2))
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
            "cpu_baseline": dict({"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                  "sample": cs.describe(NEW_TOKENS)}, **cs.extrapolation(NEW_TOKENS)),
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
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the PyTorch-eager fp16 GPU baseline of the same workload")
    ap.add_argument("--no-expert", action="store_true", help="skip the vision-expert context measurement")
    ap.add_argument("--sweep", action="store_true", help="BASELINE configs[4]: batch x sequence throughput sweep instead of the headline line")
    ap.add_argument("--sweep-batches", default="1,4,16,32")
    ap.add_argument("--sweep-seqs", default="256,1024,2048")
    ap.add_argument("--sweep-out", default="")
    ap.add_argument("--profile", action="store_true", help="minimal run for ncu (numbers printed under a profiler are not bench values)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.sweep:
        run_sweep(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
