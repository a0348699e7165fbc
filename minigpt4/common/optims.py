"""`registry.get_lr_scheduler_class(run.lr_sched)` targets with the reference's constructor and `step(cur_epoch, cur_step)`
contract (minigpt4/common/optims.py:13-97). The schedule arithmetic lives once in myriad_b200/optim.py (also used by the
fused-optimizer loop); these classes only write the value into every optimizer param group, honouring a per-group
`init_lr` override like the reference's schedule functions (:99-125)."""
from minigpt4.common.registry import registry
from myriad_b200 import optim as _sched


def _set_lr(optimizer, fn, default_peak):
    for group in optimizer.param_groups:
        group["lr"] = fn(group.get("init_lr", default_peak))


@registry.register_lr_scheduler("linear_warmup_step_lr")
class LinearWarmupStepLRScheduler:
    def __init__(self, optimizer, max_epoch, min_lr, init_lr, decay_rate=1, warmup_start_lr=-1, warmup_steps=0, **kwargs):
        self.optimizer, self.max_epoch, self.min_lr, self.init_lr = optimizer, max_epoch, min_lr, init_lr
        self.decay_rate, self.warmup_steps = decay_rate, warmup_steps
        self.warmup_start_lr = warmup_start_lr if warmup_start_lr >= 0 else init_lr

    def step(self, cur_epoch, cur_step):
        if cur_epoch == 0:
            _set_lr(self.optimizer, lambda peak: _sched.warmup_lr(cur_step, self.warmup_steps, self.warmup_start_lr, peak), self.init_lr)
        else:  # the step schedule ignores per-group peaks (reference :128-132)
            lr = _sched.step_lr(cur_epoch, self.init_lr, self.min_lr, self.decay_rate)
            _set_lr(self.optimizer, lambda peak: lr, self.init_lr)


@registry.register_lr_scheduler("linear_warmup_cosine_lr")
class LinearWarmupCosineLRScheduler:
    def __init__(self, optimizer, max_epoch, iters_per_epoch, min_lr, init_lr, warmup_steps=0, warmup_start_lr=-1, **kwargs):
        self.optimizer, self.max_epoch, self.iters_per_epoch = optimizer, max_epoch, iters_per_epoch
        self.min_lr, self.init_lr, self.warmup_steps = min_lr, init_lr, warmup_steps
        self.warmup_start_lr = warmup_start_lr if warmup_start_lr >= 0 else init_lr

    def step(self, cur_epoch, cur_step):
        total = cur_epoch * self.iters_per_epoch + cur_step
        if total < self.warmup_steps:
            _set_lr(self.optimizer, lambda peak: _sched.warmup_lr(cur_step, self.warmup_steps, self.warmup_start_lr, peak), self.init_lr)
        else:
            horizon = self.max_epoch * self.iters_per_epoch
            _set_lr(self.optimizer, lambda peak: _sched.cosine_lr(total, horizon, peak, self.min_lr), self.init_lr)
