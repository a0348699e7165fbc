"""Small helpers with the names train.py / evaluation_aqa_dataset.py / the runner import from the reference's
minigpt4/common/utils.py (`now` :35, `disable_torch_init` :41, `is_url` :50, `get_cache_path` :55, `get_abs_path` :59,
`load_json` :63, `makedir` :73). The reference's download helpers (:87-268) are not provided: there is no network on the
path this package serves; every file must already be on disk."""
import datetime
import json
import os
import re

from minigpt4.common.registry import registry


def now():
    """job id of a run: yyyymmddHHMM (train.py:84 creates it before init_distributed_mode so all ranks agree to the minute)."""
    return datetime.datetime.now().strftime("%Y%m%d%H%M")[:-1]


def disable_torch_init():
    """Skip the default nn.Linear / nn.LayerNorm initialisers: every weight is overwritten by a checkpoint right after."""
    import torch
    for cls in (torch.nn.Linear, torch.nn.LayerNorm):
        setattr(cls, "reset_parameters", lambda self: None)


def is_url(url_or_filename):
    return re.match(r"^(http|https|ftp)://", str(url_or_filename), re.IGNORECASE) is not None


def get_cache_path(rel_path):
    return os.path.expanduser(os.path.join(registry.get_path("cache_root"), rel_path))


def get_abs_path(rel_path):
    return os.path.join(registry.get_path("library_root"), rel_path)


def load_json(filename):
    with open(filename) as fh:
        return json.load(fh)


def makedir(dir_path):
    try:
        os.makedirs(dir_path, exist_ok=True)
        return True
    except OSError:
        return False


def download_url(*args, **kwargs):
    raise RuntimeError("no network on this path: place the file on disk and pass its path")


download_and_extract_archive = cache_url = download_url
