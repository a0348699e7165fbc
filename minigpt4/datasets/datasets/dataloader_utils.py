"""Loader wrappers the runner uses (reference minigpt4/datasets/datasets/dataloader_utils.py): `IterLoader` :145-181 turns
a DataLoader into an endless iterator that bumps the DistributedSampler epoch on wrap-around, `MultiIterLoader` :15-43
draws from several loaders by ratio. `PrefetchLoader` (:46-130, commented out at its only call site runner_base.py:556)
is a pass-through here: the step's host-to-device copy is already asynchronous from pinned memory."""
import random


class IterLoader:
    def __init__(self, dataloader, use_distributed=False):
        self._dataloader, self._use_distributed, self._epoch = dataloader, use_distributed, 0
        self.iter_loader = iter(dataloader)

    @property
    def epoch(self):
        return self._epoch

    def __next__(self):
        try:
            return next(self.iter_loader)
        except StopIteration:
            self._epoch += 1
            sampler = getattr(self._dataloader, "sampler", None)
            if self._use_distributed and hasattr(sampler, "set_epoch"):
                sampler.set_epoch(self._epoch)
            self.iter_loader = iter(self._dataloader)
            return next(self.iter_loader)

    def __iter__(self):
        return self

    def __len__(self):
        return len(self._dataloader)


class MultiIterLoader:
    def __init__(self, loaders, ratios=None):
        for ld in loaders:
            assert hasattr(ld, "__next__"), "Loader {} has no __next__ method.".format(ld)
        ratios = [1.0] * len(loaders) if ratios is None else [float(r) for r in ratios]
        assert len(ratios) == len(loaders)
        total = sum(ratios)
        self.loaders, self.ratios = loaders, [r / total for r in ratios]

    def __next__(self):
        return next(random.choices(self.loaders, self.ratios, k=1)[0])

    def __iter__(self):
        return self


class PrefetchLoader:
    def __init__(self, loader):
        self.loader = loader

    def __iter__(self):
        return iter(self.loader)

    def __len__(self):
        return len(self.loader)

    def __getattr__(self, name):
        return getattr(self.__dict__["loader"], name)
