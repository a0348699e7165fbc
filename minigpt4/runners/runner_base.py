"""`RunnerBase` (`run.runner: runner_base`, the default train.py:73-79 resolves): epochs, checkpoints, loaders and the
optimizer around the task loop — the reference's minigpt4/runners/runner_base.py:42-686 with the same config keys, the same
properties and the same checkpoint layout, re-expressed for the B200-native model:

  * data parallelism: the reference wraps the model in torch DDP with find_unused_parameters=True (:93-98). Here the
    model is wrapped in `myriad_b200.dp.FlatGradDataParallel`: the fused backward leaves every trainable gradient in ONE
    flat fp32 buffer, which is all-reduced (mean) over NCCL before the nn.Parameter .grad views are handed to the
    optimizer — the same semantics (parameters a rank's stage did not touch contribute zeros), one collective per step.
  * optimizer: torch.optim.AdamW over the parameters with requires_grad, split by the reference's no-weight-decay rule
    (`p.ndim < 2 or "bias" / "ln" / "bn" in name`, :111-119); GradScaler when `run.amp` (:141-149).
  * checkpoints: {"model": trainable state_dict, "optimizer", "config", "scaler", "epoch"} (:592-628), at most
    `run.max_checkpoints` kept; resume via `run.resume_ckpt_path` (:649-672).
"""
import datetime
import json
import logging
import os
import time
from pathlib import Path

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, DistributedSampler

from minigpt4.common.dist_utils import get_rank, get_world_size, is_main_process, main_process
from minigpt4.common.registry import registry
from minigpt4.common.utils import is_url
from minigpt4.datasets.data_utils import reorg_datasets_by_split
from minigpt4.datasets.datasets.dataloader_utils import IterLoader, MultiIterLoader
from myriad_b200.optim import no_weight_decay


@registry.register_runner("runner_base")
class RunnerBase:
    def __init__(self, cfg, task, model, datasets, job_id):
        self.config, self.job_id, self.task, self.datasets = cfg, job_id, task, datasets
        self._model, self._wrapped_model = model, None
        self._device = self._optimizer = self._scaler = self._dataloaders = self._lr_sched = None
        self.start_epoch = 0
        self.saved_history = []
        self.setup_output_dir()

    # ---------------------------------------------------------------------------------- config accessors
    def _run(self, key, default=None):
        return self.config.run_cfg.get(key, default)

    @property
    def device(self):
        if self._device is None:
            self._device = torch.device(self.config.run_cfg.device)
        return self._device

    @property
    def use_distributed(self):
        return self.config.run_cfg.distributed

    cuda_enabled = property(lambda self: self.device.type == "cuda")
    max_epoch = property(lambda self: int(self.config.run_cfg.max_epoch))
    log_freq = property(lambda self: int(self._run("log_freq", 50)))
    init_lr = property(lambda self: float(self.config.run_cfg.init_lr))
    min_lr = property(lambda self: float(self.config.run_cfg.min_lr))
    accum_grad_iters = property(lambda self: int(self._run("accum_grad_iters", 1)))
    test_splits = property(lambda self: self._run("test_splits", []))
    evaluate_only = property(lambda self: self.config.run_cfg.evaluate)
    use_dist_eval_sampler = property(lambda self: self._run("use_dist_eval_sampler", True))
    resume_ckpt_path = property(lambda self: self._run("resume_ckpt_path", None))

    @property
    def valid_splits(self):
        splits = self._run("valid_splits", [])
        if len(splits) == 0:
            logging.info("No validation splits found.")
        return splits

    @property
    def train_splits(self):
        splits = self._run("train_splits", [])
        if len(splits) == 0:
            logging.info("Empty train splits.")
        return splits

    # ------------------------------------------------------------------------------------------ objects
    @property
    def model(self):
        """the model on the run device, data-parallel wrapped when `run.distributed`"""
        if self._wrapped_model is None or self._model.device.type != self.device.type:
            self._model = self._model.to(self.device)
            if self.use_distributed:
                from myriad_b200.dp import FlatGradDataParallel
                self._wrapped_model = FlatGradDataParallel(self._model)
            else:
                self._wrapped_model = self._model
        return self._wrapped_model

    def unwrap_dist_model(self, model):
        return model.module if self.use_distributed else model

    @property
    def optimizer(self):
        if self._optimizer is None:
            decay, no_decay, n = [], [], 0
            for name, p in self.model.named_parameters():
                if not p.requires_grad:
                    continue
                (no_decay if no_weight_decay(name, p.ndim) else decay).append(p)
                n += p.numel()
            logging.info("number of trainable parameters: %d" % n)
            wd = float(self.config.run_cfg.weight_decay)
            groups = ([{"params": decay, "weight_decay": wd}] if decay else []) + ([{"params": no_decay, "weight_decay": 0}] if no_decay else [])
            self._optimizer = torch.optim.AdamW(groups, lr=self.init_lr, weight_decay=wd, betas=(0.9, self._run("beta2", 0.999)))
        return self._optimizer

    @property
    def scaler(self):
        if self._run("amp", False) and self._scaler is None:
            self._scaler = torch.amp.GradScaler("cuda", enabled=self.cuda_enabled)
        return self._scaler

    @property
    def lr_scheduler(self):
        if self._lr_sched is None:
            cls = registry.get_lr_scheduler_class(self.config.run_cfg.lr_sched)
            iters = self._run("iters_per_epoch", None)
            if iters is None:
                try:
                    iters = len(self.dataloaders["train"])
                except (AttributeError, TypeError):
                    iters = 10000
            self._lr_sched = cls(optimizer=self.optimizer, max_epoch=self.max_epoch, iters_per_epoch=iters, min_lr=self.min_lr,
                                 init_lr=self.init_lr, decay_rate=self._run("lr_decay_rate", None),
                                 warmup_start_lr=self._run("warmup_lr", -1), warmup_steps=self._run("warmup_steps", 0))
        return self._lr_sched

    @property
    def dataloaders(self):
        """{split: loader}; several training sets become one MultiIterLoader drawing by `sample_ratio` (reference :191-285)"""
        if self._dataloaders is None:
            self.datasets = reorg_datasets_by_split(self.datasets)
            for split, lst in self.datasets.items():
                if len(lst) == 1:
                    self.datasets[split] = lst[0]
                n = sum(len(d) for d in lst if hasattr(d, "__len__"))
                logging.info("Loaded {} records for {} split from the dataset.".format(n, split))
            run = self.config.run_cfg
            loaders = {}
            for split in sorted(self.datasets):
                ds = self.datasets[split]
                is_train = split in self.train_splits
                bsz = run.batch_size_train if split == "train" else run.batch_size_eval
                if isinstance(ds, (list, tuple)):
                    ratios = [d.sample_ratio for d in ds] if hasattr(ds[0], "sample_ratio") else None
                    loaders[split] = MultiIterLoader([self._create_loader(d, run.num_workers, bsz, is_train) for d in ds], ratios)
                else:
                    loaders[split] = self._create_loader(ds, run.num_workers, bsz, is_train)
            self._dataloaders = loaders
        return self._dataloaders

    def _create_loader(self, dataset, num_workers, bsz, is_train):
        sampler = None
        if self.use_distributed:
            sampler = DistributedSampler(dataset, shuffle=is_train, num_replicas=get_world_size(), rank=get_rank())
            if not self.use_dist_eval_sampler and not is_train:
                sampler = None
        # every AnomalyDetection item is a (normal, simulated-anomaly) PAIR, so b items make a batch of 2b images (:546-549)
        if getattr(dataset, "DatasetName", "") == "AnomalyDetection":
            bsz = max(1, bsz // 2)
        loader = DataLoader(dataset, batch_size=bsz, num_workers=num_workers, pin_memory=True, sampler=sampler,
                            shuffle=sampler is None and is_train, collate_fn=getattr(dataset, "collater", None), drop_last=is_train)
        return IterLoader(loader, use_distributed=self.use_distributed) if is_train else loader

    def create_loaders(self, datasets, num_workers, batch_sizes, is_trains, collate_fns=None, dataset_ratios=None):
        return [self._create_loader(d, num_workers, b, t) for d, b, t in zip(datasets, batch_sizes, is_trains)]

    @property
    def train_loader(self):
        return self.dataloaders["train"]

    def setup_output_dir(self):
        out = Path(registry.get_path("library_root")) / self.config.run_cfg.output_dir / self.job_id
        self.output_dir, self.result_dir = out, out / "result"
        self.result_dir.mkdir(parents=True, exist_ok=True)
        registry.register_path("result_dir", str(self.result_dir))
        registry.register_path("output_dir", str(self.output_dir))

    # --------------------------------------------------------------------------------------------- loop
    def train(self):
        t0 = time.time()
        best_metric, best_epoch = 0, 0
        self.log_config()
        if not self.evaluate_only and self.resume_ckpt_path is not None:
            self._load_checkpoint(self.resume_ckpt_path)
        cur_epoch = self.start_epoch
        for cur_epoch in range(self.start_epoch, self.max_epoch):
            if not self.evaluate_only:
                logging.info("Start training")
                self.log_stats(split_name="train", stats=self.train_epoch(cur_epoch))
            if len(self.valid_splits) > 0:
                for split in self.valid_splits:
                    logging.info("Evaluating on {}.".format(split))
                    val_log = self.eval_epoch(split_name=split, cur_epoch=cur_epoch)
                    if val_log is not None and is_main_process():
                        assert "agg_metrics" in val_log, "No agg_metrics found in validation log."
                        if val_log["agg_metrics"] > best_metric and split == "val":
                            best_epoch, best_metric = cur_epoch, val_log["agg_metrics"]
                            self._save_checkpoint(cur_epoch, is_best=True)
                        val_log.update({"best_epoch": best_epoch})
                        self.log_stats(val_log, split)
            elif not self.evaluate_only:
                self._save_checkpoint(cur_epoch, is_best=False)
            if self.evaluate_only:
                break
            if self.use_distributed:
                dist.barrier()
        self.evaluate(cur_epoch="best" if len(self.valid_splits) > 0 else cur_epoch, skip_reload=self.evaluate_only)
        logging.info("Training time {}".format(datetime.timedelta(seconds=int(time.time() - t0))))

    def evaluate(self, cur_epoch="best", skip_reload=False):
        if len(self.test_splits) > 0:
            return {s: self.eval_epoch(split_name=s, cur_epoch=cur_epoch, skip_reload=skip_reload) for s in self.test_splits}

    def train_epoch(self, epoch):
        self.model.train()
        return self.task.train_epoch(epoch=epoch, model=self.model, data_loader=self.train_loader, optimizer=self.optimizer,
                                     scaler=self.scaler, lr_scheduler=self.lr_scheduler, cuda_enabled=self.cuda_enabled,
                                     log_freq=self.log_freq, accum_grad_iters=self.accum_grad_iters)

    @torch.no_grad()
    def eval_epoch(self, split_name, cur_epoch, skip_reload=False):
        loader = self.dataloaders.get(split_name, None)
        assert loader, "data_loader for split {} is None.".format(split_name)
        model = self.unwrap_dist_model(self.model)
        if not skip_reload and cur_epoch == "best":
            model = self._reload_best_model(model)
        model.eval()
        self.task.before_evaluation(model=model, dataset=self.datasets[split_name])
        results = self.task.evaluation(model, loader)
        if results is not None:
            return self.task.after_evaluation(val_result=results, split_name=split_name, epoch=cur_epoch)

    # -------------------------------------------------------------------------------------- checkpoints
    @main_process
    def _save_checkpoint(self, cur_epoch, is_best=False):
        model = self.unwrap_dist_model(self.model)
        frozen = {k for k, p in model.named_parameters() if not p.requires_grad}  # only what is trained gets stored
        state = {k: v for k, v in model.state_dict().items() if k not in frozen}
        obj = {"model": state, "optimizer": self.optimizer.state_dict(), "config": self.config.to_dict(),
               "scaler": self.scaler.state_dict() if self.scaler else None, "epoch": cur_epoch}
        path = os.path.join(self.output_dir, "checkpoint_{}.pth".format("best" if is_best else cur_epoch))
        if len(self.saved_history) >= self._run("max_checkpoints", 1):  # rolling window of the newest checkpoints
            oldest = self.saved_history.pop(0)
            if oldest != path and os.path.exists(oldest):
                os.remove(oldest)
        self.saved_history.append(path)
        logging.info("Saving checkpoint at epoch {} to {}.".format(cur_epoch, path))
        torch.save(obj, path)

    def _reload_best_model(self, model):
        path = os.path.join(self.output_dir, "checkpoint_best.pth")
        logging.info("Loading checkpoint from {}.".format(path))
        model.load_state_dict(torch.load(path, map_location="cpu", weights_only=False)["model"], strict=False)
        return model

    def _load_checkpoint(self, url_or_filename):
        if is_url(url_or_filename) or not os.path.isfile(url_or_filename):
            raise RuntimeError("checkpoint url or path is invalid")
        ck = torch.load(url_or_filename, map_location=self.device, weights_only=False)
        self.unwrap_dist_model(self.model).load_state_dict(ck["model"], strict=False)
        self.optimizer.load_state_dict(ck["optimizer"])
        if self.scaler and ck.get("scaler"):
            self.scaler.load_state_dict(ck["scaler"])
        self.start_epoch = ck["epoch"] + 1
        logging.info("Resume checkpoint from {}".format(url_or_filename))

    @main_process
    def log_stats(self, stats, split_name):
        if isinstance(stats, dict):
            with open(os.path.join(self.output_dir, "log.txt"), "a") as fh:
                fh.write(json.dumps({"%s_%s" % (split_name, k): v for k, v in stats.items()}) + "\n")

    @main_process
    def log_config(self):
        with open(os.path.join(self.output_dir, "log.txt"), "a") as fh:
            fh.write(json.dumps(self.config.to_dict(), indent=4) + "\n")
