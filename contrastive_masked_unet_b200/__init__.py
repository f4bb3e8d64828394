"""cmunet-b200: B200-native (sm_100a) kernels behind the CM-UNet pretraining / fine-tuning hot path, exposed through
the reference's own `torch.nn.Module` and loss interfaces (drop-in for Pretraining/CM-UNet/cmae/models and
Finetuning/model.py + metrics.py).  The compute lives in libcmu_b200.so (C ABI: include/cmu_b200.h)."""
from ._lib import CmuError, lib  # noqa: F401
from .modules import (CM_UNet, CMUNetPretrainHead, DoubleConv, DownBlock, MaskStream, MODELS, MUNetPretrainDecoder,  # noqa: F401
                      NonLinearNeck, UNet_encoder, UpBlock, build, cmunet_config, concat_all_gather,
                      try_register_mmengine)
from .finetune import CrossEntropyLoss, DiceLoss, IoU, Loss, Metric, MultipliedLoss, SumOfLosses, UNet, soft_cldice  # noqa: F401

from .moco import Moco_v2, MocoUNetEncoder  # noqa: F401
from .data import CMUNetGpuPipeline, SampleParams  # noqa: F401

__version__ = '0.1.0'


def _apply_env_knobs():
    """CMU_DEBUG_KNOBS="10=1,9=1": A/B switches of include/cmu_b200.h:cmu_debug_set for whole-step experiments."""
    import os
    spec = os.environ.get('CMU_DEBUG_KNOBS', '')
    for kv in filter(None, spec.split(',')):
        k, v = kv.split('=')
        lib.cmu_debug_set(int(k), int(v))


_apply_env_knobs()
