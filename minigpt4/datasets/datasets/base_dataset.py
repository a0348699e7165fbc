"""Map-style dataset base with the reference's interface (minigpt4/datasets/datasets/base_dataset.py:30-120):
`annotation` list, `vis_processor` / `text_processor`, optional in-memory preload, `collater`, `set_processors`."""
import json
import os
from concurrent.futures import ThreadPoolExecutor

from PIL import Image
from torch.utils.data import ConcatDataset as _TorchConcat
from torch.utils.data import Dataset
from torch.utils.data.dataloader import default_collate


def imreader(args):
    """preload worker: decode one image into the shared annotation record (reference :18-22)"""
    i, records = args
    rec = records[i]
    rec["image"] = Image.open(rec["path"]).convert("RGB")
    rec["image"].load()


class BaseDataset(Dataset):
    def __init__(self, vis_processor=None, text_processor=None, vis_root=None, ann_paths=(), is_preload=False, preload_fn=imreader):
        self.vis_root, self.ann_paths = vis_root, list(ann_paths)
        self.load_annotations()
        self.vis_processor, self.text_processor = vis_processor, text_processor
        self.is_preload, self.preload_fn, self._cache = is_preload, preload_fn, None
        if is_preload:
            self.preload()

    def __len__(self):
        return len(self.annotation)

    def collater(self, samples):
        return default_collate(samples)

    def load_annotations(self):
        self.annotation = []
        for path in self.ann_paths:
            with open(path) as fh:
                self.annotation.extend(json.load(fh)["annotations"])

    def prepare_img(self, index):
        raise NotImplementedError("%s does not implement prepare_img" % type(self).__name__)

    def get_image_path(self, rel_path):
        return os.path.join(self.vis_root, rel_path)

    def construct_preload_maps(self):
        return [{"path": self.get_image_path(a["img_path"]), "rel_path": a["img_path"]} for a in self.annotation]

    def post_preload(self, results):
        self._cache = {rec["rel_path"]: rec["image"] for rec in results}

    def preload(self):
        records = self.construct_preload_maps()
        with ThreadPoolExecutor() as pool:
            list(pool.map(self.preload_fn, ((i, records) for i in range(len(records)))))
        self.post_preload(records)

    def set_processors(self, vis_processor, text_processor):
        self.vis_processor, self.text_processor = vis_processor, text_processor

    def _add_instance_ids(self, key="instance_id"):
        for i, ann in enumerate(self.annotation):
            ann[key] = str(i)


class ConcatDataset(_TorchConcat):
    def collater(self, samples):
        """collate on the keys every sample has (reference :105-120)"""
        shared = set.intersection(*(set(s) for s in samples))
        return self.datasets[0].collater([{k: s[k] for k in s if k in shared} for s in samples])
