"""Drop-in fine-tuning model and losses: Finetuning/model.py:84-131 (`UNet`) and the loss part of
Finetuning/metrics.py (:19-82 Loss algebra and `__name__`s, :135-220 Dice / IoU, :503-504 CrossEntropyLoss).
Dice / IoU / CE of one (pred, gt) pair come from ONE fused reduction kernel (cached per input pair), CE's gradient
from the same launch.  Quirk Q7 is kept: the thresholded Dice / IoU terms carry no gradient; results are float64."""
import os
import re

import torch
import torch.nn as nn

from . import functional as Fn
from . import ops
from ._lib import lib
from .modules import DoubleConv, DownBlock, UpBlock, _head_fusable, register  # noqa: F401


@register
class UNet(nn.Module):
    """FT/model.py:84-131.  forward(x:(B,H,W)) -> (B, out_classes, H, W) fp32 logits."""

    def __init__(self, out_classes=2, up_sample_mode='conv_transpose'):
        super().__init__()
        self.up_sample_mode = up_sample_mode
        self.down_conv1 = DownBlock(1, 64)
        self.down_conv2 = DownBlock(64, 128)
        self.down_conv3 = DownBlock(128, 256)
        self.down_conv4 = DownBlock(256, 512)
        self.double_conv = DoubleConv(512, 1024)
        self.up_conv4 = UpBlock(1024, 512, self.up_sample_mode)
        self.up_conv3 = UpBlock(512, 256, self.up_sample_mode)
        self.up_conv2 = UpBlock(256, 128, self.up_sample_mode)
        self.up_conv1 = UpBlock(128, 64, self.up_sample_mode)
        self.conv_last = nn.Conv2d(64, out_classes, kernel_size=1)
        self.out_classes = out_classes

    def forward(self, x):
        ops._need_cuda(x)
        if self.out_classes != 2:
            raise NotImplementedError('the sm_100a output head is written for out_classes=2 (FT/train.py default)')
        x = x.unsqueeze(1)
        x, skip1_out = self.down_conv1(x)
        x, skip2_out = self.down_conv2(x)
        x, skip3_out = self.down_conv3(x)
        x, skip4_out = self.down_conv4(x)
        x = self.double_conv(x)
        x = self.up_conv4(x, skip4_out)
        x = self.up_conv3(x, skip3_out)
        x = self.up_conv2(x, skip2_out)
        if _head_fusable(self.conv_last) and os.environ.get('CMU_NO_HEAD_FUSION') != '1':
            return self.up_conv1(x, skip1_out, head=self.conv_last)      # BN + ReLU folded into conv_last (head_fused.cu)
        x = self.up_conv1(x, skip1_out)
        return Fn.Head1x1Fn.apply(Fn.to_act(x), self.conv_last.weight, self.conv_last.bias)


# ------------------------------------------------------------------------------------------------------- losses
class _SegLossFn(torch.autograd.Function):
    """(dice_loss, iou_loss, ce_loss) float64 0-d tensors from one launch; only ce_loss is differentiable."""

    @staticmethod
    def forward(ctx, logits, gt, dice_eps, beta, iou_eps):
        logits = logits.contiguous().float()
        gt = gt.contiguous()
        if gt.dtype != torch.float64:
            gt = gt.double()
        gs = torch.ones(1, dtype=torch.float32, device=logits.device)
        out, dl = ops.seg_losses(logits, gt, gs if ctx.needs_input_grad[0] else None, dice_eps, beta, iou_eps)
        if dl is not None:
            ctx.save_for_backward(dl)
        dice, iou, ce = out[0], out[1], out[2]
        ctx.mark_non_differentiable(dice, iou)
        return dice, iou, ce

    @staticmethod
    def backward(ctx, g_dice, g_iou, g_ce):
        (dl,) = ctx.saved_tensors
        return dl * g_ce.float(), None, None, None, None


class BaseObject(nn.Module):
    def __init__(self, name=None):
        super().__init__()
        self._name = name

    @property
    def __name__(self):
        if self._name is None:
            name = self.__class__.__name__
            s1 = re.sub('(.)([A-Z][a-z]+)', r'\1_\2', name)
            return re.sub('([a-z0-9])([A-Z])', r'\1_\2', s1).lower()
        return self._name


class Metric(BaseObject):
    pass


class Loss(BaseObject):
    def __add__(self, other):
        if isinstance(other, Loss):
            return SumOfLosses(self, other)
        raise ValueError('Loss should be inherited from `Loss` class')

    def __radd__(self, other):
        return self.__add__(other)

    def __mul__(self, value):
        if isinstance(value, (int, float)):
            return MultipliedLoss(self, value)
        raise ValueError('Loss should be inherited from `BaseLoss` class')

    def __rmul__(self, other):
        return self.__mul__(other)


class SumOfLosses(Loss):
    def __init__(self, l1, l2):
        super().__init__(name='{} + {}'.format(l1.__name__, l2.__name__))
        self.l1, self.l2 = l1, l2

    def __call__(self, *inputs):
        return self.l1.forward(*inputs) + self.l2.forward(*inputs)


class MultipliedLoss(Loss):
    def __init__(self, loss, multiplier):
        if len(loss.__name__.split('+')) > 1:
            name = '{} * ({})'.format(multiplier, loss.__name__)
        else:
            name = '{} * {}'.format(multiplier, loss.__name__)
        super().__init__(name=name)
        self.loss, self.multiplier = loss, multiplier

    def __call__(self, *inputs):
        return self.multiplier * self.loss.forward(*inputs)

    def forward(self, *inputs):
        return self.multiplier * self.loss.forward(*inputs)


_CACHE = {}


def _fused(y_pr, y_gt, dice_eps=1e-5, beta=1.0, iou_eps=1e-7):
    """All three reductions of a (pred, gt) pair in one launch; reused when Dice, CE and IoU are evaluated on the same
    tensors (FT/train.py:128-139 calls the loss and every metric per batch)."""
    key = (y_pr.data_ptr(), y_pr._version, y_gt.data_ptr(), y_gt._version, tuple(y_pr.shape), dice_eps, beta, iou_eps,
           y_pr.requires_grad and torch.is_grad_enabled())
    hit = _CACHE.get('k')
    if hit is not None and hit[0] == key and hit[1]() is y_pr:
        return hit[2]
    import weakref
    out = _SegLossFn.apply(y_pr, y_gt, dice_eps, beta, iou_eps)
    _CACHE['k'] = (key, weakref.ref(y_pr), out)
    return out


def _check_cfg(activation, threshold, ignore_channels, y_pr):
    if activation not in ('softmax', 'softmax2d') or threshold != 0.5 or list(ignore_channels or []) != [0] \
            or y_pr.shape[1] != 2:
        raise NotImplementedError('the fused sm_100a reduction implements the configuration of FT/train.py:455-465: '
                                  "2 classes, activation='softmax', threshold=0.5, ignore_channels=[0]")


class DiceLoss(Loss):
    def __init__(self, eps=1e-5, beta=1.0, activation=None, ignore_channels=None, threshold=None, **kwargs):
        super().__init__(**kwargs)
        self.eps, self.beta, self.activation = eps, beta, activation
        self.ignore_channels, self.threshold = ignore_channels, threshold

    def forward(self, y_pr, y_gt):
        _check_cfg(self.activation, self.threshold, self.ignore_channels, y_pr)
        return _fused(y_pr, y_gt, dice_eps=self.eps, beta=self.beta)[0]


class IoU(Metric):
    __name__ = 'iou_loss'

    def __init__(self, eps=1e-7, threshold=0.5, activation=None, ignore_channels=None, **kwargs):
        super().__init__(**kwargs)
        self.eps, self.threshold, self.activation, self.ignore_channels = eps, threshold, activation, ignore_channels

    def forward(self, y_pr, y_gt):
        _check_cfg(self.activation, self.threshold, self.ignore_channels, y_pr)
        return _fused(y_pr, y_gt, iou_eps=self.eps)[1]


class soft_cldice(Loss):
    """FT/metrics.py:401-430: soft-clDice of the thresholded prediction against the float64 target (eval metric of
    FT/train.py:464); 10 soft-skeleton iterations as the reference hard-codes (`iter_` is accepted and ignored there too)."""
    __name__ = 'soft_clDice'

    def __init__(self, iter_=3, smooth=1., exclude_background=False, threshold=0.5, activation=None, ignore_channels=None):
        super().__init__()
        if exclude_background:
            raise NotImplementedError('exclude_background=True empties the single foreground channel (FT/metrics.py:421-423)')
        self.iter, self.smooth, self.threshold = iter_, smooth, threshold
        self.activation, self.ignore_channels = activation, ignore_channels

    @torch.no_grad()
    def forward(self, y_pred, y_true):
        _check_cfg(self.activation, self.threshold, self.ignore_channels, y_pred)
        ops._need_cuda(y_pred, y_true)
        logits = y_pred.detach().contiguous().float()
        gt = y_true.detach().contiguous().double()
        n, _, h, w = logits.shape
        nbytes = lib.cmu_soft_cldice_workspace_bytes(n, h, w)
        ws = torch.empty(nbytes // 8, dtype=torch.float64, device=logits.device)
        out = torch.empty(1, dtype=torch.float64, device=logits.device)
        lib.cmu_soft_cldice(logits.data_ptr(), gt.data_ptr(), n, h, w, 10, float(self.smooth), ws.data_ptr(), nbytes,
                            out.data_ptr(), ops._stream())
        return out.reshape(())


class CrossEntropyLoss(Loss):
    """nn.CrossEntropyLoss with probability targets (FT/metrics.py:503-504), mean over N*H*W, float64 result."""

    def forward(self, y_pr, y_gt):
        if y_pr.shape[1] != 2:
            raise NotImplementedError('2-class soft-target cross-entropy only')
        return _fused(y_pr, y_gt)[2]
