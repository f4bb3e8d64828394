"""Minimal Registry with the call surface cmae/registry.py:8-152 and cmae/models/builder.py use
(test-only shim for mmengine.registry)."""


class Registry:
    def __init__(self, name, build_func=None, parent=None, scope=None, locations=None):
        self.name = name
        self.parent = parent
        self.scope = scope
        self.locations = locations or []
        self._module_dict = {}

    def _register(self, module, name=None, force=False):
        key = name or module.__name__
        if key in self._module_dict and not force:
            raise KeyError(f'{key} is already registered in {self.name}')
        self._module_dict[key] = module

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def deco(cls):
            self._register(cls, name, force)
            return cls
        return deco

    def get(self, key):
        if key in self._module_dict:
            return self._module_dict[key]
        if self.parent is not None:
            return self.parent.get(key)
        return None

    def build(self, cfg, *args, **kwargs):
        return build_from_cfg(cfg, self, *args, **kwargs)

    def __contains__(self, key):
        return self.get(key) is not None


def build_from_cfg(cfg, registry, default_args=None):
    cfg = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            cfg.setdefault(k, v)
    typ = cfg.pop('type')
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError(f'{typ} is not in the {registry.name} registry')
    return cls(**cfg)


_ROOT_NAMES = [
    'RUNNERS', 'RUNNER_CONSTRUCTORS', 'LOOPS', 'HOOKS', 'DATASETS', 'DATA_SAMPLERS', 'TRANSFORMS',
    'MODELS', 'MODEL_WRAPPERS', 'WEIGHT_INITIALIZERS', 'OPTIMIZERS', 'OPTIM_WRAPPERS',
    'OPTIM_WRAPPER_CONSTRUCTORS', 'PARAM_SCHEDULERS', 'METRICS', 'EVALUATOR', 'TASK_UTILS',
    'VISUALIZERS', 'VISBACKENDS', 'LOG_PROCESSORS', 'INFERENCERS', 'FUNCTIONS',
]
for _n in _ROOT_NAMES:
    globals()[_n] = Registry(_n.lower())
