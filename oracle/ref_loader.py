"""TEST INFRASTRUCTURE ONLY — loads the UNMODIFIED reference (read-only, /root/reference) in the build
container so that golden vectors can be minted from it (oracle/make_goldens.py) and the restatement in
oracle/cmunet_oracle.py can be pinned against it.  Nothing here is imported by the product package, and
nothing here works on the GPU box (the reference tree does not travel).

Recipe follows SURVEY.md Appendix D:
  * put oracle/ref_shim (mmengine/mmcv stubs) and Pretraining/CM-UNet on sys.path;
  * CPU execution patches for the device literals of the reference (quirk Q4:
    UNet_encoder.py:84,156, cmunet.py:128, cmunet_head.py:85): Tensor.cuda / Module.cuda -> identity,
    Tensor.to("cuda:0") -> identity, a 1-rank gloo process group;
  * S-generalisation: the literal `224, 224` at cmunet.py:130 is replaced by the image size through a
    subclass whose forward_train is the reference's own source with that single substitution.
"""
import contextlib
import inspect
import os
import runpy
import sys
import types

REF_ROOT = os.environ.get('CMU_REFERENCE_ROOT', '/root/reference')
CMU_DIR = os.path.join(REF_ROOT, 'Pretraining', 'CM-UNet')
FT_DIR = os.path.join(REF_ROOT, 'Finetuning')
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_shim')


def reference_available():
    return os.path.isdir(CMU_DIR) and os.path.isdir(FT_DIR)


def _ensure_paths():
    for p in (CMU_DIR, SHIM_DIR):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, CMU_DIR)
    sys.path.insert(0, SHIM_DIR)


_PATCHED = False


def apply_cpu_patches():
    """Q4 device literals -> CPU no-ops (only when no GPU is visible)."""
    global _PATCHED
    import torch
    if _PATCHED or torch.cuda.is_available():
        return
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _orig_to = torch.Tensor.to

    def _to(self, *args, **kwargs):
        args = tuple(a for a in args if not (isinstance(a, str) and a.startswith('cuda')))
        if isinstance(kwargs.get('device'), str) and kwargs['device'].startswith('cuda'):
            kwargs.pop('device')
        if not args and not kwargs:
            return self
        return _orig_to(self, *args, **kwargs)

    torch.Tensor.to = _to
    _PATCHED = True


def ensure_process_group(backend='gloo'):
    import torch.distributed as td
    if not td.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29577')
        td.init_process_group(backend, rank=int(os.environ.get('RANK', 0)),
                              world_size=int(os.environ.get('WORLD_SIZE', 1)))


def import_cmae():
    """Returns (cmae.models module, MODELS registry, model cfg dict from cmunet_config.py:5-42)."""
    _ensure_paths()
    import cmae  # noqa: F401
    import cmae.models as models
    from cmae.registry import MODELS
    cfg = runpy.run_path(os.path.join(CMU_DIR, 'configs', 'cmunet_config.py'))['model']
    return models, MODELS, cfg


def build_reference_cm_unet(img_size=224):
    """Reference CM_UNet built through its own registry from its own config; for img_size != 224 the
    S-generalised subclass (one literal substituted, projector.in_channels = S*S)."""
    import copy
    models, MODELS, cfg = import_cmae()
    from cmae.models import CM_UNet
    cfg = copy.deepcopy(cfg)
    cfg['neck']['projector']['in_channels'] = img_size * img_size
    cfg.pop('type')
    if img_size == 224:
        return CM_UNet(**cfg)
    src = inspect.getsource(CM_UNet.forward_train)
    assert '224, 224' in src
    src = inspect.cleandoc('\n' + src).replace('224, 224', 'img.shape[-2], img.shape[-1]')
    mod = sys.modules[CM_UNet.__module__]
    ns = dict(vars(mod))
    exec(src, ns)

    class CM_UNet_S(CM_UNet):
        forward_train = ns['forward_train']

    return CM_UNet_S(**cfg)


def import_finetune():
    """Reference Finetuning/model.py and metrics.py (metrics needs skimage stubs)."""
    for name in ('skimage', 'skimage.morphology', 'skimage.measure'):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.skeletonize = m.skeletonize_3d = m.find_contours = None
            sys.modules[name] = m
    if FT_DIR not in sys.path:
        sys.path.insert(0, FT_DIR)
    import importlib
    model = importlib.import_module('model')
    metrics = importlib.import_module('metrics')
    return model, metrics


def import_moco():
    """Reference Pretraining/MoCo/pl_bolts/models/self_supervised/moco/moco2_module.py, unmodified, loaded by path with
    stand-ins for what this container lacks: pytorch_lightning (oracle/ref_shim/pytorch_lightning), wandb, and the parts
    of the vendored pl_bolts tree the module imports but that are absent or unimportable."""
    _ensure_paths()
    moco_dir = os.path.join(REF_ROOT, 'Pretraining', 'MoCo', 'pl_bolts', 'models', 'self_supervised', 'moco')

    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(sys.modules[name], k, v)
        return sys.modules[name]

    stub('wandb')
    stub('pl_bolts')
    stub('pl_bolts.metrics', mean=lambda *a, **k: None, precision_at_k=lambda *a, **k: None)
    stub('pl_bolts.utils', _TORCHVISION_AVAILABLE=True, _PIL_AVAILABLE=True)
    stub('pl_bolts.utils.warnings', warn_missing_pkg=lambda *a, **k: None)
    stub('transforms', Moco2EvalImagenetTransforms=object, Moco2TrainImagenetTransforms=object)   # data transforms: unused
    if moco_dir not in sys.path:
        sys.path.insert(0, moco_dir)
    import importlib
    return importlib.import_module('moco2_module')


@contextlib.contextmanager
def seeded(seed):
    import numpy as np
    import torch
    torch.manual_seed(seed)
    np.random.seed(seed)
    yield
