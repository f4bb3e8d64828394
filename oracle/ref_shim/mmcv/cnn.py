"""mmcv.cnn.build_norm_layer as used at nonlinear_neck.py:58 (test-only shim).
SyncBN -> nn.SyncBatchNorm on CUDA; nn.BatchNorm1d on CPU (SyncBatchNorm rejects CPU tensors;
at world size 1 the two are numerically the same op)."""
import torch
import torch.nn as nn

FORCE_CPU_BN = True


def build_norm_layer(cfg, num_features, postfix=''):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    requires_grad = cfg.pop('requires_grad', True)
    cfg.setdefault('eps', 1e-5)
    if typ == 'SyncBN':
        layer = nn.BatchNorm1d(num_features, **cfg) if FORCE_CPU_BN else nn.SyncBatchNorm(num_features, **cfg)
        name = 'bn'
    elif typ in ('BN', 'BN1d'):
        layer = nn.BatchNorm1d(num_features, **cfg)
        name = 'bn'
    elif typ == 'BN2d':
        layer = nn.BatchNorm2d(num_features, **cfg)
        name = 'bn'
    else:
        raise NotImplementedError(typ)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return name + str(postfix), layer
