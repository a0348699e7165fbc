"""Optimisation-loop pieces around MyriadTrainer: the reference's LR schedules
(minigpt4/common/optims.py:57-125), GradScaler's dynamic loss scale (runner_base.py:141-149 -> torch.cuda.amp.GradScaler:
init 2**16, x2 after 2000 clean steps, x0.5 on inf/nan with the step skipped) and checkpoint save / resume in the
reference's layout (runner_base.py:592-672: {"model": trainable state_dict, "optimizer", "config", "scaler", "epoch"}).
Host logic only; the arithmetic of a step (all-reduce, unscale, inf check, AdamW) runs on the device."""
import math
import os

import torch


class LinearWarmupCosineLR:
    """optims.py:57-97 (`linear_warmup_cosine_lr`): linear warm-up over `warmup_steps` steps of epoch 0 (the reference
    feeds the in-epoch step), then cosine decay over max_epoch * iters_per_epoch steps."""

    def __init__(self, max_epoch, iters_per_epoch, min_lr, init_lr, warmup_steps=0, warmup_start_lr=-1):
        self.max_epoch, self.iters_per_epoch, self.min_lr, self.init_lr = max_epoch, iters_per_epoch, min_lr, init_lr
        self.warmup_steps = warmup_steps
        self.warmup_start_lr = warmup_start_lr if warmup_start_lr >= 0 else init_lr

    def lr(self, cur_epoch, cur_step):
        total = cur_epoch * self.iters_per_epoch + cur_step
        if total < self.warmup_steps:
            return warmup_lr(cur_step, self.warmup_steps, self.warmup_start_lr, self.init_lr)
        return cosine_lr(total, self.max_epoch * self.iters_per_epoch, self.init_lr, self.min_lr)


class LinearWarmupStepLR:
    """optims.py:14-54 (`linear_warmup_step_lr`): warm-up during epoch 0, then init_lr * decay_rate**epoch floored at min_lr."""

    def __init__(self, max_epoch, min_lr, init_lr, decay_rate=1, warmup_start_lr=-1, warmup_steps=0):
        self.max_epoch, self.min_lr, self.init_lr, self.decay_rate = max_epoch, min_lr, init_lr, decay_rate
        self.warmup_steps = warmup_steps
        self.warmup_start_lr = warmup_start_lr if warmup_start_lr >= 0 else init_lr

    def lr(self, cur_epoch, cur_step):
        if cur_epoch == 0:
            return warmup_lr(cur_step, self.warmup_steps, self.warmup_start_lr, self.init_lr)
        return step_lr(cur_epoch, self.init_lr, self.min_lr, self.decay_rate)


def cosine_lr(step, max_step, init_lr, min_lr):      # optims.py:99-112
    return (init_lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * step / max_step)) + min_lr


def warmup_lr(step, max_step, init_lr, max_lr):      # optims.py:115-125
    return min(max_lr, init_lr + (max_lr - init_lr) * step / max(max_step, 1))


def step_lr(epoch, init_lr, min_lr, decay_rate):     # optims.py:128-132
    return max(min_lr, init_lr * (decay_rate ** epoch))


class DynamicLossScale:
    """torch.cuda.amp.GradScaler's scale schedule, driven by the device-side inf/nan flag of myr_adamw_step."""

    def __init__(self, init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.scale, self.growth_factor, self.backoff_factor, self.growth_interval = init_scale, growth_factor, backoff_factor, growth_interval
        self._good = 0

    def update(self, found_inf):
        if found_inf:
            self.scale *= self.backoff_factor
            self._good = 0
        else:
            self._good += 1
            if self._good >= self.growth_interval:
                self.scale *= self.growth_factor
                self._good = 0
        return self.scale

    def state_dict(self):
        return {"scale": self.scale, "growth_factor": self.growth_factor, "backoff_factor": self.backoff_factor,
                "growth_interval": self.growth_interval, "_growth_tracker": self._good}

    def load_state_dict(self, sd):
        self.scale, self._good = float(sd["scale"]), int(sd.get("_growth_tracker", 0))


def no_weight_decay(name, ndim):
    """runner_base.py:113-119: biases, 1-d tensors and anything with 'ln' / 'bn' in its name are not decayed."""
    return ndim < 2 or "bias" in name or "ln" in name or "bn" in name


def _is_main_process():
    import torch.distributed as dist
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


def save_checkpoint(trainer, path, epoch, config=None, scaler=None):
    """runner_base.py:592-628: only parameters with requires_grad are stored under "model" (reference key names and layouts,
    so `Myriad.load_state_dict(strict=False)` and the reference's own `ckpt:` loader read it). Rank 0 writes (the reference
    guards `_save_checkpoint` with @main_process); the other ranks return the path without touching the file system.

    "model", "config", "scaler", "epoch" follow the reference layout. "optimizer" does NOT: it is the fused optimizer's own
    record (kind = "myriad_b200.fused_adamw": step count, hyper-parameters, loss scale, flat AdamW moments by parameter
    name), which torch.optim.AdamW.load_state_dict cannot read — a checkpoint written here resumes through load_checkpoint
    below, and its "model" entry loads anywhere."""
    if not _is_main_process():
        return path
    if scaler is None:
        scaler = getattr(trainer, "scaler", None)
    if hasattr(trainer, "sync_loss_scale"):
        trainer.sync_loss_scale()
    names = list(trainer.segments.keys())
    obj = {
        "model": {k: v.cpu() for k, v in trainer.export_state_dict().items()},
        # flat AdamW moments by parameter name (same layouts as "model"); step counts the optimizer steps taken
        "optimizer": {"kind": "myriad_b200.fused_adamw", "step": trainer.opt_step, "hyper": dict(trainer.hp),
                      "loss_scale": float(trainer.loss_scale),
                      "exp_avg": {k: v.cpu() for k, v in trainer.export_flat(trainer.exp_avg).items()},
                      "exp_avg_sq": {k: v.cpu() for k, v in trainer.export_flat(trainer.exp_avg_sq).items()},
                      "param_names": names},
        "config": config if config is not None else {},
        "scaler": scaler.state_dict() if scaler is not None else None,
        "epoch": epoch,
    }
    tmp = path + ".tmp"
    torch.save(obj, tmp)
    os.replace(tmp, path)
    return path


def load_checkpoint(trainer, path, scaler=None):
    """runner_base.py:649-672: restores the trainable parameters, the optimizer state and the scaler; returns the epoch to
    resume from (stored epoch + 1)."""
    import warnings
    ck = torch.load(path, map_location="cpu", weights_only=False)
    missing = [k for k in trainer.segments if k not in ck["model"]]
    if missing:
        warnings.warn("checkpoint %s lacks %d trainable tensors (kept at their current values): %s ..." % (path, len(missing), missing[:3]))
    trainer.import_state_dict(ck["model"])
    opt = ck.get("optimizer")
    if isinstance(opt, dict) and opt.get("kind") == "myriad_b200.fused_adamw":
        trainer.opt_step = int(opt["step"])
        trainer.hp.update(opt.get("hyper", {}))
        trainer.import_flat(trainer.exp_avg, opt["exp_avg"])
        trainer.import_flat(trainer.exp_avg_sq, opt["exp_avg_sq"])
        if "loss_scale" in opt:
            trainer.loss_scale = float(opt["loss_scale"])
    elif isinstance(opt, dict) and "state" in opt and "param_groups" in opt:
        n = import_torch_adamw_state(trainer, opt, ck["model"])
        if n == 0:
            warnings.warn("checkpoint %s: torch AdamW state could not be matched to the trainable tensors; the optimizer restarts cold" % path)
    else:
        warnings.warn("checkpoint %s carries no usable optimizer state (kind %r): AdamW moments and step restart from zero"
                      % (path, opt.get("kind") if isinstance(opt, dict) else type(opt).__name__))
    if ck.get("scaler"):
        for sc in (scaler, getattr(trainer, "scaler", None)):
            if sc is not None:
                sc.load_state_dict(ck["scaler"])
        if hasattr(trainer, "scaler") and not (isinstance(opt, dict) and "loss_scale" in opt):
            trainer.loss_scale = float(trainer.scaler.scale)
    if hasattr(trainer, "scaler"):
        trainer.scaler.scale = trainer.loss_scale
    return int(ck.get("epoch", -1)) + 1


def import_torch_adamw_state(trainer, opt_state, model_state):
    """A checkpoint written by the reference runner (runner_base.py:606 `optimizer.state_dict()` of torch.optim.AdamW over the
    parameters with requires_grad, in named_parameters() order = the order of its "model" entry): copy exp_avg / exp_avg_sq /
    step into the flat buffers. Parameters are matched by position and checked by shape. Returns how many were imported."""
    keys = [k for k in model_state if k in trainer.segments]
    ids = [i for g in opt_state["param_groups"] for i in g["params"]]
    st = opt_state["state"]
    if len(ids) != len(keys):
        # the runner lists decayed parameters first, then the un-decayed ones (runner_base.py:111-133)
        return 0
    decayed = [k for k in keys if not no_weight_decay(k, model_state[k].ndim)]
    ordered = decayed + [k for k in keys if k not in decayed]
    m, v, n, step = {}, {}, 0, 0
    for pid, key in zip(ids, ordered):
        rec = st.get(pid)
        if rec is None or tuple(rec["exp_avg"].shape) != tuple(model_state[key].shape):
            continue
        m[key], v[key] = rec["exp_avg"], rec["exp_avg_sq"]
        step = max(step, int(rec["step"]))
        n += 1
    if n:
        trainer.import_flat(trainer.exp_avg, m)
        trainer.import_flat(trainer.exp_avg_sq, v)
        trainer.opt_step = step
    return n
