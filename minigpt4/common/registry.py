"""Global class / path / state registry with the reference's interface (minigpt4/common/registry.py:9-329):
`@registry.register_model("myriad")`, `registry.get_model_class(name)`, `register_path/get_path`,
`register/get/unregister` with dotted names. Re-implemented from the interface, table-driven."""


class Registry:
    mapping = {
        "builder_name_mapping": {}, "task_name_mapping": {}, "processor_name_mapping": {}, "model_name_mapping": {},
        "lr_scheduler_name_mapping": {}, "runner_name_mapping": {}, "state": {}, "paths": {},
    }

    @classmethod
    def _register_class(cls, kind, name, base_check=None):
        table = cls.mapping[kind + "_name_mapping"]

        def wrap(klass):
            if base_check is not None:
                base = base_check()
                if base is not None and not issubclass(klass, base):
                    raise AssertionError("All %ss must inherit %s" % (kind, base.__name__))
            if name in table:
                raise KeyError("Name '%s' already registered for %s." % (name, table[name]))
            table[name] = klass
            return klass

        return wrap

    @classmethod
    def register_model(cls, name):
        def base():
            from minigpt4.models.base_model import BaseModel
            return BaseModel
        return cls._register_class("model", name, base)

    @classmethod
    def register_builder(cls, name):
        return cls._register_class("builder", name)

    @classmethod
    def register_task(cls, name):
        return cls._register_class("task", name)

    @classmethod
    def register_processor(cls, name):
        return cls._register_class("processor", name)

    @classmethod
    def register_lr_scheduler(cls, name):
        return cls._register_class("lr_scheduler", name)

    @classmethod
    def register_runner(cls, name):
        return cls._register_class("runner", name)

    @classmethod
    def register_path(cls, name, path):
        assert isinstance(path, str), "All path must be str."
        if name in cls.mapping["paths"]:
            raise KeyError("Name '%s' already registered." % name)
        cls.mapping["paths"][name] = path

    @classmethod
    def register(cls, name, obj):
        cur = cls.mapping["state"]
        parts = name.split(".")
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = obj

    @classmethod
    def get_model_class(cls, name):
        return cls.mapping["model_name_mapping"].get(name, None)

    @classmethod
    def get_builder_class(cls, name):
        return cls.mapping["builder_name_mapping"].get(name, None)

    @classmethod
    def get_task_class(cls, name):
        return cls.mapping["task_name_mapping"].get(name, None)

    @classmethod
    def get_processor_class(cls, name):
        return cls.mapping["processor_name_mapping"].get(name, None)

    @classmethod
    def get_lr_scheduler_class(cls, name):
        return cls.mapping["lr_scheduler_name_mapping"].get(name, None)

    @classmethod
    def get_runner_class(cls, name):
        return cls.mapping["runner_name_mapping"].get(name, None)

    @classmethod
    def list_models(cls):
        return sorted(cls.mapping["model_name_mapping"].keys())

    @classmethod
    def list_tasks(cls):
        return sorted(cls.mapping["task_name_mapping"].keys())

    @classmethod
    def list_processors(cls):
        return sorted(cls.mapping["processor_name_mapping"].keys())

    @classmethod
    def list_lr_schedulers(cls):
        return sorted(cls.mapping["lr_scheduler_name_mapping"].keys())

    @classmethod
    def list_runners(cls):
        return sorted(cls.mapping["runner_name_mapping"].keys())

    @classmethod
    def list_datasets(cls):
        return sorted(cls.mapping["builder_name_mapping"].keys())

    @classmethod
    def get_path(cls, name):
        return cls.mapping["paths"].get(name, None)

    @classmethod
    def get(cls, name, default=None, no_warning=False):
        value = cls.mapping["state"]
        for p in name.split("."):
            if not isinstance(value, dict) or p not in value:
                value = default
                break
            value = value[p]
        if "writer" in cls.mapping["state"] and value == default and not no_warning:
            cls.mapping["state"]["writer"].warning("Key %s is not present in registry, returning default value of %s" % (name, default))
        return value

    @classmethod
    def unregister(cls, name):
        return cls.mapping["state"].pop(name, None)


registry = Registry()
