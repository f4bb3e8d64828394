"""Minimal stand-in for pytorch-lightning 1.6 (TEST INFRASTRUCTURE): just enough surface for the unmodified reference
Pretraining/MoCo/pl_bolts/models/self_supervised/moco/moco2_module.py to be imported and its Moco_v2 math (forward,
_momentum_update_key_encoder, _dequeue_and_enqueue, _compute_l_s) to be executed on CPU for minting golden vectors.
No training loop, no logging, no distributed plugins."""
import inspect

import torch.nn as nn


class AttributeDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class _NoPluginTrainer:
    """trainer stand-in: single process, no DDP plugin (the reference hard-codes gpus=1, moco2_module.py:452)."""
    training_type_plugin = None
    strategy = None
    datamodule = None
    logger = None


class LightningModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.hparams = AttributeDict()
        self.trainer = _NoPluginTrainer()

    def save_hyperparameters(self, *args, logger=True, ignore=()):
        frame = inspect.currentframe().f_back
        names = inspect.getargvalues(frame)
        for k in names.args:
            if k not in ('self',) and k not in ignore:
                self.hparams[k] = names.locals[k]

    def log_dict(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass


class LightningDataModule:
    def __init__(self, *a, **k):
        pass


class Trainer:
    def __init__(self, *a, **k):
        raise NotImplementedError('shim: no training loop')
