"""BaseModule / BaseModel with mmengine's init_weights() recursion semantics (test-only shim).
Real behaviour restated: apply `init_cfg` (only the `Constant` initialiser on `_BatchNorm`/`GroupNorm`
layers is used by the reference, nonlinear_neck.py:46-51), then call `init_weights()` on direct
children that define it."""
import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm


def is_model_wrapper(model):
    return isinstance(model, (nn.DataParallel, nn.parallel.DistributedDataParallel))


def _apply_init_cfg(module, init_cfg):
    cfgs = init_cfg if isinstance(init_cfg, (list, tuple)) else [init_cfg]
    for cfg in cfgs:
        if cfg is None:
            continue
        if cfg.get('type') != 'Constant':
            raise NotImplementedError(f'shim only knows the Constant initialiser, got {cfg}')
        layers = cfg.get('layer', [])
        layers = [layers] if isinstance(layers, str) else layers
        for m in module.modules():
            hit = ('_BatchNorm' in layers and isinstance(m, _BatchNorm)) or \
                  ('GroupNorm' in layers and isinstance(m, nn.GroupNorm)) or \
                  (type(m).__name__ in layers)
            if hit:
                if getattr(m, 'weight', None) is not None:
                    nn.init.constant_(m.weight, cfg.get('val', 0))
                if getattr(m, 'bias', None) is not None:
                    nn.init.constant_(m.bias, cfg.get('bias', 0))


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg
        self._is_init = False

    def init_weights(self):
        if not self._is_init:
            if self.init_cfg:
                _apply_init_cfg(self, self.init_cfg)
            for m in self.children():
                if hasattr(m, 'init_weights') and not getattr(m, '_is_init', False):
                    m.init_weights()
            self._is_init = True


class BaseModel(BaseModule):
    def __init__(self, data_preprocessor=None, init_cfg=None):
        super().__init__(init_cfg=init_cfg)
