"""Training-side bridge of the drop-in `Myriad` class: `Myriad.forward(samples) -> {"loss"}` (reference
minigpt4/models/myriad.py:377-431) on the B200-native trainer (myriad_b200.training.MyriadTrainer).

The reference's runner owns the optimisation loop (base_task.py:233-271: `loss = model(samples)["loss"]`,
`scaler.scale(loss).backward()`, `optimizer.step()` over the parameters with requires_grad, DDP gradient averaging). To drop
in under that loop unchanged, the loss returned here is an autograd leaf-connected tensor: a custom autograd Function runs
the fused forward + backward of the device path once, and its `backward` hands each trainable nn.Parameter its gradient
(reference layout, multiplied by the incoming grad — i.e. GradScaler's scale and DDP's hooks see what they expect).
The master parameters stay the module's nn.Parameters; they are mirrored into the trainer's flat fp32 buffer before every
step (one device-to-device copy of ~460 MB at full size).

A loop that does not need torch's optimizer can call `MyriadTrainer.train_step` directly (bench.py does): that path adds
the flat-buffer NCCL all-reduce and the fused AdamW kernel.
"""
import torch


def _trainer_for(model):
    """One MyriadTrainer per model instance, built on first use from the same frozen weights as the inference engine."""
    tr = getattr(model, "_trainer", None)
    if tr is None:
        if not torch.cuda.is_available():
            raise RuntimeError("Myriad (B200-native) training needs a CUDA device; there is no CPU path")
        from myriad_b200.training import MyriadTrainer
        dev = torch.device("cuda", torch.cuda.current_device())
        tr = MyriadTrainer(model._Merged(model._frozen, model.trainable_state()), model.dims, device=dev,
                           max_batch=8, max_seq=512)
        tr.overlap_allreduce = False  # under the runner the whole flat buffer is averaged once, in _FusedStep.backward
        model._trainer = tr
    return tr


def _sync_params_to_trainer(tr, state):
    """nn.Parameter (reference layout) -> flat fp32 buffer (conv filters NHWC), then refresh the fp16 operand copies."""
    with torch.no_grad():
        for key, (off, shape, kind) in tr.segments.items():
            p = state[key].detach().to(tr.dev, torch.float32)
            if kind == "conv":
                p = p.permute(0, 2, 3, 1)
            n = 1
            for s in shape:
                n *= s
            tr.flat_params[off:off + n].copy_(p.reshape(-1))
    tr.refresh_trainables()


class _FusedStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, image, maps, stage, ids_before, ids_after, text_ids, text_mask, keys, *params):
        tr = _trainer_for(model)
        _sync_params_to_trainer(tr, dict(zip(keys, params)))
        # peft's lora_dropout is live only in train mode (reference: llama_model stays in train() under the runner, myriad.py:171-178)
        cfg = getattr(model, "lora_config", None) or {}
        tr.lora_dropout = float(cfg.get("lora_dropout", 0.0)) if model.training else 0.0
        loss = tr.forward_backward(image, maps, stage, ids_before, ids_after, text_ids, text_mask)
        ctx.trainer, ctx.keys, ctx.model = tr, keys, model
        ctx.devices = [p.device for p in params]
        return loss.detach().clone()

    @staticmethod
    def backward(ctx, grad_out):
        world = getattr(ctx.model, "dp_world", 1)
        if world > 1:  # data parallel (myriad_b200.dp.FlatGradDataParallel): one all-reduce of the flat buffer = DDP's mean
            from myriad_b200.dp import mean_flat_grads_
            mean_flat_grads_(ctx.trainer.flat_grads, world)
        grads = ctx.trainer.export_grads()  # reference layouts, already unscaled
        g = grad_out.to(ctx.trainer.dev)
        out = [(grads[k] * g).to(d) for k, d in zip(ctx.keys, ctx.devices)]
        return (None,) * 9 + tuple(out)


def myriad_training_forward(model, image, maps, stage, questions, text_inputs, double_prompts=False):
    """The body of Myriad.forward after prepare_sample / the stage and task draws (myriad.py:390-431)."""
    tok = model.llama_tokenizer
    prompts = ["###Human: " + q + " ###Assistant: " for q in questions]
    if double_prompts:
        prompts = prompts + prompts
    ids_before, ids_after = model._split_prompts(prompts, image.device)
    tok.padding_side = "right"
    text = [t + model.end_sym for t in text_inputs]
    enc = tok(text, return_tensors="pt", padding="longest", truncation=True, max_length=model.max_txt_len, add_special_tokens=False)
    state = model.trainable_state()
    keys = tuple(k for k in state if state[k].requires_grad)
    tr = _trainer_for(model)
    missing = [k for k in tr.segments if k not in keys]
    if missing:
        raise RuntimeError("trainable parameters expected by the device trainer are frozen or absent: %s" % missing[:4])
    keys = tuple(tr.segments.keys())
    dev = tr.dev
    return _FusedStep.apply(model, image.to(dev).float().contiguous(), maps.to(dev).float().contiguous(), stage, ids_before, ids_after,
                            enc.input_ids.cpu(), enc.attention_mask.cpu(), keys, *[state[k] for k in keys])
