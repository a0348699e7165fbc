"""yaml configuration with the reference's merge order (minigpt4/common/config.py:16-151): model-class default yaml
<- user yaml `model:` <- `--options k=v` dotlist; datasets likewise. OmegaConf is not installed in this image, so
the used subset (`load`, `merge`, dotlist overrides, `.get`, attribute access, `to_container`) is provided by
`Node`, a dict with attribute access."""
import json
import re

import yaml

from minigpt4.common.registry import registry


class Node(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(o):
        if isinstance(o, dict):
            return Node({k: Node.wrap(v) for k, v in o.items()})
        if isinstance(o, list):
            return [Node.wrap(v) for v in o]
        return o


class _Loader(yaml.SafeLoader):
    """PyYAML follows YAML 1.1, where `1e-4` (no dot) is a string; OmegaConf — what the reference's yamls are written for —
    reads it as a float (`init_lr: 1e-4`, `warmup_lr: 1e-6` in every train config)."""


_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"^[-+]?(?:[0-9][0-9_]*\.[0-9_]*(?:[eE][-+]?[0-9]+)?|\.[0-9_]+(?:[eE][-+]?[0-9]+)?|[0-9][0-9_]*[eE][-+]?[0-9]+"
               r"|\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$"),
    list("-+0123456789."))


def load_yaml(path):
    with open(path) as fh:
        return Node.wrap(yaml.load(fh, Loader=_Loader) or {})


def merge(*nodes):
    out = Node()
    for n in nodes:
        for k, v in (n or {}).items():
            if isinstance(v, dict) and isinstance(out.get(k), dict):
                out[k] = merge(out[k], v)
            else:
                out[k] = Node.wrap(v)
    return out


def from_dotlist(opts):
    out = Node()
    for item in opts or []:
        key, _, val = item.partition("=")
        cur = out
        parts = key.split(".")
        for p in parts[:-1]:
            cur = cur.setdefault(p, Node())
        cur[parts[-1]] = yaml.load(val, Loader=_Loader)
    return out


def to_container(node):
    return json.loads(json.dumps(node))


class Config:
    def __init__(self, args):
        self.args = args
        registry.register("configuration", self)
        user = from_dotlist(self._opt_list(getattr(args, "options", None)))
        config = load_yaml(args.cfg_path)
        runner = Node({"run": config.get("run", Node())})
        model = self.build_model_config(config, **user)
        datasets = self.build_dataset_config(config)
        self.config = merge(runner, model, datasets, user)

    @staticmethod
    def _opt_list(opts):
        if not opts:
            return []
        if "=" in opts[0]:
            return list(opts)
        return ["%s=%s" % (k, v) for k, v in zip(opts[0::2], opts[1::2])]

    @staticmethod
    def build_model_config(config, **kwargs):
        model = config.get("model", None)
        assert model is not None, "Missing model configuration file."
        model_cls = registry.get_model_class(model.arch)
        assert model_cls is not None, "Model '%s' has not been registered." % model.arch
        model_type = (kwargs.get("model", None) or {}).get("model_type", None) or model.get("model_type", None)
        assert model_type is not None, "Missing model_type."
        default = load_yaml(model_cls.default_config_path(model_type=model_type))
        return merge(Node(), default, Node({"model": config["model"]}))

    @staticmethod
    def build_dataset_config(config):
        datasets = config.get("datasets", None)
        if datasets is None:
            return Node({"datasets": Node()})
        out = Node()
        for name in datasets:
            builder = registry.get_builder_class(name)
            default = Node()
            if builder is not None and hasattr(builder, "default_config_path"):
                default = load_yaml(builder.default_config_path(type=datasets[name].get("type", "default")))
            out = merge(out, default, Node({"datasets": Node({name: config["datasets"][name]})}))
        return out

    def get_config(self):
        return self.config

    @property
    def run_cfg(self):
        return self.config.run

    @property
    def datasets_cfg(self):
        return self.config.datasets

    @property
    def model_cfg(self):
        return self.config.model

    def to_dict(self):
        return to_container(self.config)

    def pretty_print(self):
        print(json.dumps(self.to_dict(), indent=2, sort_keys=True))
