"""Batch plumbing between the DataLoader and the model with the reference's names (minigpt4/datasets/data_utils.py:
`apply_to_sample` :66, `move_to_cuda` :83, `prepare_sample` :90, `reorg_datasets_by_split` :99, `concat_datasets` :125).
webdataset pipelines (ChainDataset :33-63) are not on the Myriad path and are not provided."""
import torch
from torch.utils.data import IterableDataset


def apply_to_sample(f, sample):
    """Map f over every tensor of a nested dict / list batch; everything else (strings, ints) passes through."""
    def walk(x):
        if torch.is_tensor(x):
            return f(x)
        if isinstance(x, dict):
            return {k: walk(v) for k, v in x.items()}
        if isinstance(x, list):
            return [walk(v) for v in x]
        return x
    return walk(sample) if len(sample) else {}


def move_to_cuda(sample):
    return apply_to_sample(lambda t: t.cuda(non_blocking=True), sample)


def prepare_sample(samples, cuda_enabled=True):
    return move_to_cuda(samples) if cuda_enabled else samples


def reorg_datasets_by_split(datasets):
    """{dataset name: {split: dataset}} -> {split: [datasets]}"""
    by_split = {}
    for per_split in datasets.values():
        for split, ds in per_split.items():
            by_split.setdefault(split, []).append(ds)
    return by_split


def concat_datasets(datasets):
    """{split: [datasets]} -> {split: dataset}: train splits are concatenated (map-style only), val / test must be single."""
    from minigpt4.datasets.datasets.base_dataset import ConcatDataset
    out = {}
    for split, lst in datasets.items():
        if split != "train":
            assert len(lst) == 1, "Do not support multiple {} datasets.".format(split)
            out[split] = lst[0]
            continue
        if any(isinstance(d, IterableDataset) for d in lst):
            raise NotImplementedError("iterable (webdataset) training sets are not supported on this path")
        out[split] = lst[0] if len(lst) == 1 else ConcatDataset(lst)
    return out
