"""`BaseTask`: what the runner calls per epoch (reference minigpt4/tasks/base_task.py:19-343): `build_model`,
`build_datasets`, `train_step`, `train_epoch` / `train_iters` -> `_train_inner_loop`, `evaluation`, `save_result`.

The inner loop keeps the reference's contract step for step (:156-303): sample -> prepare_sample -> add epoch / iters keys ->
lr_scheduler.step -> autocast(enabled = scaler is not None) -> model(samples)["loss"] -> scaler.scale(loss).backward() ->
every accum_grad_iters: scaler.step(optimizer) / scaler.update() / zero_grad -> meters. The drop-in Myriad returns a loss
whose backward fills the trainable nn.Parameters' .grad (minigpt4/models/train_step.py), so torch's AdamW and GradScaler
work on it unchanged."""
import json
import logging
import os

import torch
import torch.distributed as dist

from minigpt4.common.dist_utils import get_rank, get_world_size, is_dist_avail_and_initialized, is_main_process
from minigpt4.common.logger import MetricLogger, SmoothedValue
from minigpt4.common.registry import registry
from minigpt4.datasets.data_utils import prepare_sample


class BaseTask:
    def __init__(self, **kwargs):
        self.inst_id_key = "instance_id"

    @classmethod
    def setup_task(cls, **kwargs):
        return cls()

    def build_model(self, cfg):
        model_cfg = cfg.model_cfg
        return registry.get_model_class(model_cfg.arch).from_config(model_cfg)

    def build_datasets(self, cfg):
        """{dataset name: {split: Dataset}} for every entry under `datasets:` (reference :36-66)."""
        datasets_cfg = cfg.datasets_cfg
        assert len(datasets_cfg) > 0, "At least one dataset has to be specified."
        out = {}
        for name in datasets_cfg:
            ds_cfg = datasets_cfg[name]
            built = registry.get_builder_class(name)(ds_cfg).build_datasets()
            built["train"].name = name
            if "sample_ratio" in ds_cfg:
                built["train"].sample_ratio = ds_cfg.sample_ratio
            out[name] = built
        return out

    def train_step(self, model, samples):
        return model(samples)

    def valid_step(self, model, samples):
        raise NotImplementedError

    def before_evaluation(self, model, dataset, **kwargs):
        model.before_evaluation(dataset=dataset, task_type=type(self))

    def after_evaluation(self, **kwargs):
        pass

    def inference_step(self):
        raise NotImplementedError

    def evaluation(self, model, data_loader, cuda_enabled=True):
        meters = MetricLogger(delimiter="  ")
        results = []
        for samples in meters.log_every(data_loader, 10, "Evaluation"):
            results.extend(self.valid_step(model=model, samples=prepare_sample(samples, cuda_enabled=cuda_enabled)))
        if is_dist_avail_and_initialized():
            dist.barrier()
        return results

    def train_epoch(self, epoch, model, data_loader, optimizer, lr_scheduler, scaler=None, cuda_enabled=False, log_freq=50,
                    accum_grad_iters=1):
        return self._train_inner_loop(epoch=epoch, iters_per_epoch=lr_scheduler.iters_per_epoch, model=model, data_loader=data_loader,
                                      optimizer=optimizer, scaler=scaler, lr_scheduler=lr_scheduler, log_freq=log_freq,
                                      cuda_enabled=cuda_enabled, accum_grad_iters=accum_grad_iters)

    def train_iters(self, epoch, start_iters, iters_per_inner_epoch, model, data_loader, optimizer, lr_scheduler, scaler=None,
                    cuda_enabled=False, log_freq=50, accum_grad_iters=1):
        return self._train_inner_loop(epoch=epoch, start_iters=start_iters, iters_per_epoch=iters_per_inner_epoch, model=model,
                                      data_loader=data_loader, optimizer=optimizer, scaler=scaler, lr_scheduler=lr_scheduler,
                                      log_freq=log_freq, cuda_enabled=cuda_enabled, accum_grad_iters=accum_grad_iters)

    def _train_inner_loop(self, epoch, iters_per_epoch, model, data_loader, optimizer, lr_scheduler, scaler=None, start_iters=None,
                          log_freq=50, cuda_enabled=False, accum_grad_iters=1):
        use_amp = scaler is not None
        batches = data_loader if hasattr(data_loader, "__next__") else iter(data_loader)
        meters = MetricLogger(delimiter="  ")
        meters.add_meter("lr", SmoothedValue(window_size=1, fmt="{value:.6f}"))
        meters.add_meter("loss", SmoothedValue(window_size=1, fmt="{value:.4f}"))
        logging.info("Start training epoch {}, {} iters per inner epoch.".format(epoch, iters_per_epoch))
        header = "Train: data epoch: [{}]".format(epoch)
        inner_epoch = epoch
        if start_iters is not None:  # iteration-based runner: the schedule follows the inner epoch
            inner_epoch = start_iters // iters_per_epoch
            header += "; inner epoch [{}]".format(inner_epoch)
        for i in meters.log_every(range(iters_per_epoch), log_freq, header):
            samples = prepare_sample(next(batches), cuda_enabled=cuda_enabled)
            samples.update({"epoch": inner_epoch, "num_iters_per_epoch": iters_per_epoch, "iters": i})
            lr_scheduler.step(cur_epoch=inner_epoch, cur_step=i)
            with torch.autocast("cuda", enabled=use_amp and torch.cuda.is_available()):
                outputs = self.train_step(model=model, samples=samples)
            extra = {}
            if isinstance(outputs, torch.Tensor):
                loss = outputs
            else:
                loss = outputs["loss"]
                extra = {k: outputs[k] for k in ("seg_loss", "llm_loss") if outputs.get(k) is not None}
            (scaler.scale(loss) if use_amp else loss).backward()
            if (i + 1) % accum_grad_iters == 0:
                if use_amp:
                    scaler.step(optimizer)
                    scaler.update()
                else:
                    optimizer.step()
                optimizer.zero_grad()
            meters.update(loss=loss.item(), lr=optimizer.param_groups[0]["lr"])
            for k, v in extra.items():
                if k not in meters.meters:
                    meters.add_meter(k, SmoothedValue(window_size=1, fmt="{value:.4f}"))
                meters.update(**{k: v})
        meters.synchronize_between_processes()
        logging.info("Averaged stats: " + meters.global_avg())
        return {k: "{:.3f}".format(m.global_avg) for k, m in meters.meters.items()}

    @staticmethod
    def save_result(result, result_dir, filename, remove_duplicate=""):
        """every rank writes its shard, rank 0 merges them (optionally de-duplicated on one key) (reference :305-343)"""
        with open(os.path.join(result_dir, "%s_rank%d.json" % (filename, get_rank())), "w") as fh:
            json.dump(result, fh)
        if is_dist_avail_and_initialized():
            dist.barrier()
        final = os.path.join(result_dir, "%s.json" % filename)
        if is_main_process():
            merged, seen = [], set()
            for rank in range(get_world_size()):
                with open(os.path.join(result_dir, "%s_rank%d.json" % (filename, rank))) as fh:
                    for rec in json.load(fh):
                        if remove_duplicate:
                            if rec[remove_duplicate] in seen:
                                continue
                            seen.add(rec[remove_duplicate])
                        merged.append(rec)
            with open(final, "w") as fh:
                json.dump(merged, fh)
            print("result file saved to %s" % final)
        return final
