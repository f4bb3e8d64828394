"""Drop-in `torch.nn.Module`s for the CM-UNet pretraining path: same class names, constructor kwargs, forward
signatures, return structures and state_dict keys as the reference
(Pretraining/CM-UNet/cmae/models/{backbones/UNet_encoder.py, necks/munet_neck.py, necks/nonlinear_neck.py,
heads/cmunet_head.py, algorithms/cmunet.py, algorithms/base.py}); the math runs in libcmu_b200.so.

The `nn.Conv2d` / `nn.BatchNorm2d` / `nn.Linear` children exist only as PARAMETER CONTAINERS (they give the
reference's key names, default initialisation and RNG consumption); their own forward is never called.
CUDA (sm_100a) only: CPU tensors raise, there is no fallback."""
import math
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import functional as Fn
from . import ops
from ._lib import CmuError, lib

MODELS = {}


def register(cls):
    MODELS[cls.__name__] = cls
    return cls


def build(cfg):
    """Minimal stand-in for mmengine's `MODELS.build(cfg)` (cmae/registry.py:83-84)."""
    if isinstance(cfg, nn.Module):
        return cfg
    cfg = dict(cfg)
    typ = cfg.pop('type')
    cls = MODELS[typ] if isinstance(typ, str) else typ
    return cls(**cfg)


# --------------------------------------------------------------------------------------------------------- mask stream
class MaskStream:
    """Device-resident MT19937 stream that continues numpy's legacy global RNG bit-for-bit
    (UNet_encoder.py:124 `np.random.shuffle`).  By default it is seeded lazily from `np.random.get_state()` at first
    use, so `np.random.seed(s)` before training gives the reference's masks (quirk Q2: online and target encoders
    share ONE stream, online first).

    pair_mode (set by CM_UNet): the stream is inherently sequential, so after the (online, target) pair of step t has
    been served the pair of step t+1 is generated on a side CUDA stream while step t computes -- same draws, same
    order, off the critical path.  The prefetch is speculative and always reversible: the state is snapshotted before
    the pair (`backup`) and between its halves (`mid`), so a call sequence that is not online -> target (a changed
    batch / size, `extract_feat` / mode='tensor' calls without a target call, `get_numpy_state`) rolls the stream back
    to its logical position and continues exactly where numpy would be."""

    def __init__(self):
        self.state = None
        self.pair_mode = False
        self._pref = None            # (key, mask, event, backup, mid): a prefetched pair nobody has consumed yet
        self._target_owed = None     # (tkey, mid): online half consumed, its target half already drawn speculatively
        self._side = None
        self._last_key = None
        self._deferred = None        # key of a pair to prefetch once the current step's backward has been enqueued
        self.defer_prefetch = True   # fire the prefetch from the end of backward instead of right after the target call
        self._grad_step = False      # the online call of this step ran with autograd on (a backward will follow)

    def _ensure(self, device):
        if self.state is None:
            self.set_numpy_state(np.random.get_state(), device)
        elif self.state.device != device:
            self._settle()
            self.state = self.state.to(device)

    def seed(self, seed, device='cuda'):
        self._pref = self._target_owed = None     # a new seed discards any speculation (after the side stream is done)
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        self.state = torch.empty(lib.cmu_mask_state_words(), dtype=torch.int32, device=device)
        lib.cmu_mask_seed(self.state.data_ptr(), int(seed) & 0xFFFFFFFF, ops._stream())

    def set_numpy_state(self, np_state, device='cuda'):
        assert np_state[0] == 'MT19937'
        self._pref = self._target_owed = None
        if self._side is not None and torch.cuda.is_available():
            torch.cuda.current_stream().wait_stream(self._side)
        words = np.concatenate([np.asarray(np_state[1], dtype=np.uint32), np.array([np_state[2]], dtype=np.uint32)])
        self.state = torch.from_numpy(words.view(np.int32).copy()).to(device)

    def get_numpy_state(self):
        """Logical position of the stream (speculatively drawn shuffles do not count)."""
        st = self.state
        if self._pref is not None:
            self._pref[2].synchronize()
            st = self._pref[3]
        elif self._target_owed is not None:
            st = self._target_owed[1]
        w = st.cpu().numpy().view(np.uint32)
        return ('MT19937', w[:624].copy(), int(w[624]), 0, 0.0)

    def fire_deferred_prefetch(self):
        """Called when the step's backward has been enqueued (from the first layer's backward node): the single-warp
        draw kernel then runs under the optimizer / EMA kernels instead of holding an SM while the persistent
        tensor-core kernels of the forward pass want all 148 (same-box A/B: the early prefetch cost as much as it hid)."""
        if self._deferred is not None and self._pref is None and self._target_owed is None:
            key, dev = self._deferred
            self._deferred = None
            self._prefetch(key, dev)

    def _settle(self):
        """Roll the device state back to the logical stream position (undo whatever was drawn speculatively)."""
        self._deferred = None
        if self._pref is not None:
            _, _, ev, backup, _ = self._pref
            torch.cuda.current_stream().wait_event(ev)
            self.state.copy_(backup)
            self._pref = None
        elif self._target_owed is not None:
            self.state.copy_(self._target_owed[1])        # keep the online half, undo the target half
        self._target_owed = None

    _drop_prefetch = _settle

    @staticmethod
    def _k(img_size, patch_size, mask_ratio):
        g = img_size // patch_size
        return min(int(mask_ratio * img_size * img_size) // (patch_size * patch_size), g * g)

    def _launch(self, batch, img_size, patch_size, k, n_shuffles, device, want_mask=True):
        mask = torch.empty(batch, img_size, img_size, dtype=torch.uint8, device=device) if want_mask else None
        g = img_size // patch_size
        nbytes = lib.cmu_mask_workspace_bytes(batch, g * g, k)
        ws = torch.empty(max(nbytes // 4, 1), dtype=torch.int32, device=device)
        lib.cmu_mask_generate(self.state.data_ptr(), ops._ptr(mask), ws.data_ptr(), batch, img_size, patch_size, k,
                              n_shuffles, ops._stream())
        return mask

    def _prefetch(self, key, device):
        batch, img_size, patch_size, k = key
        if self._side is None:
            self._side = torch.cuda.Stream(device=device)
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            backup = self.state.clone()
            mask = self._launch(batch, img_size, patch_size, k, batch, device)                    # online half
            mid = self.state.clone()
            self._launch(batch, img_size, patch_size, 0, batch, device, want_mask=False)          # target draws (Q2)
            ev = torch.cuda.Event()
            ev.record(self._side)
        self._pref = (key, mask, ev, backup, mid)

    def generate(self, batch, img_size, patch_size, mask_ratio, device):
        """One `create_random_patch_mask` call: consumes `batch` shuffles, returns ((B,S,S) uint8 mask, K)."""
        self._ensure(device)
        k = self._k(img_size, patch_size, mask_ratio)
        if not self.pair_mode:
            self._settle()
            return self._launch(batch, img_size, patch_size, k, batch, device), k
        if k > 0:                                             # online call
            key = (batch, img_size, patch_size, k)
            self._last_key = key
            self._grad_step = torch.is_grad_enabled()
            self._deferred = None
            if self._target_owed is not None:                 # previous online call was not followed by its target call
                self._settle()
            if self._pref is not None and self._pref[0] == key:
                _, mask, ev, _, mid = self._pref
                cur = torch.cuda.current_stream()
                cur.wait_event(ev)
                mask.record_stream(cur)
                self._pref = None
                self._target_owed = (key[:3], mid)
                return mask, k
            self._settle()
            return self._launch(batch, img_size, patch_size, k, batch, device), k
        # target call (mask_ratio 0): B shuffles are drawn and discarded (Q2)
        tkey = (batch, img_size, patch_size)
        if self._target_owed is not None and self._target_owed[0] == tkey:
            self._target_owed = None                          # already drawn by the prefetched pair
            mask = torch.zeros(batch, img_size, img_size, dtype=torch.uint8, device=device)
        else:
            self._settle()
            mask = self._launch(batch, img_size, patch_size, 0, batch, device)
        nxt = self._last_key
        if nxt is not None and nxt[:3] == tkey:
            if self.defer_prefetch and self._grad_step:
                self._deferred = (nxt, device)                # fired by fire_deferred_prefetch() at the end of backward
            else:
                self._prefetch(nxt, device)                   # pair of the next step, on the side stream
        return mask, 0

    def __getstate__(self):
        d = dict(self.__dict__)
        if d.get('_pref') is not None:
            d['_pref'][2].synchronize()
            d['state'] = d['_pref'][3]
        elif d.get('_target_owed') is not None:
            d['state'] = d['_target_owed'][1]
        d['_pref'] = d['_side'] = d['_target_owed'] = d['_deferred'] = None
        return d


# --------------------------------------------------------------------------------------------------------- conv blocks
def _bn_cfg(bn, pool, want_act=True):
    mom = bn.momentum if bn.momentum is not None else 0.1
    return Fn.BNConfig(bn.training, mom, bn.eps, pool, want_act)


def _conv_bn_relu(conv, bn, x0, x1=None, pool=False, want_act=True):
    if bn.training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return Fn.ConvBNReLUFn.apply(Fn.to_act(x0), None if x1 is None else Fn.to_act(x1), conv.weight, conv.bias, bn.weight,
                                 bn.bias, bn.running_mean, bn.running_var, _bn_cfg(bn, pool, want_act))


def _head_fusable(head):
    """conv_last shapes the fused decoder tail is written for (64 -> 2, 1x1): the only ones the reference configs use."""
    return (isinstance(head, nn.Conv2d) and head.in_channels == 64 and head.out_channels == 2 and
            head.kernel_size == (1, 1) and head.bias is not None)


def _conv_bn_relu_head(conv, bn, head, x0):
    if bn.training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return Fn.ConvBNReLUHeadFn.apply(Fn.to_act(x0), None, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean,
                                     bn.running_var, head.weight, head.bias, _bn_cfg(bn, False))


@register
class DoubleConv(nn.Module):
    """[Conv3x3 -> BatchNorm2d -> ReLU] x 2 (UNet_encoder.py:8-30)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True))
        self.in_channels, self.out_channels = in_channels, out_channels

    def run(self, x, skip=None, mask=None, pool=False, want_skip=True, on_backward=None, head=None):
        """x: NCHW-shaped tensor (fp32 (N,1,H,W) for the first layer); skip: second concat source; mask: (B,H,W) uint8
        whose image 0 masks the whole batch (first layer only)."""
        seq = self.double_conv
        ops._need_cuda(x)
        if self.in_channels == 1:
            bn = seq[1]
            if bn.training and bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
            cfg = _bn_cfg(bn, False)
            cfg.on_backward = on_backward
            a = Fn.FirstConvBNReLUFn.apply(x.reshape(x.shape[0], x.shape[-2], x.shape[-1]), mask, seq[0].weight,
                                           seq[0].bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, cfg)
        else:
            if mask is not None:
                raise CmuError('input masking is fused into the 1-channel first layer only')
            a = _conv_bn_relu(seq[0], seq[1], x, skip, False)
        if head is not None:      # decoder tail: second conv + BN + ReLU + conv_last in one autograd node (no `a` in HBM)
            return _conv_bn_relu_head(seq[3], seq[4], head, a)
        return _conv_bn_relu(seq[3], seq[4], a, None, pool, want_skip)

    def forward(self, x):
        return self.run(x)


@register
class DownBlock(nn.Module):
    """DoubleConv then MaxPool2d(2); returns (down, skip) (UNet_encoder.py:32-49)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.double_conv = DoubleConv(in_channels, out_channels)
        self.down_sample = nn.MaxPool2d(2)

    def forward(self, x, mask=None, want_skip=True, on_backward=None):
        """want_skip=False: the full-resolution skip tensor is not written and comes back as None (the frozen target
        encoder of CM_UNet and the MoCo encoders only use the pooled path; backward does not need the tensor)."""
        skip_out, down_out = self.double_conv.run(x, mask=mask, pool=True, want_skip=want_skip, on_backward=on_backward)
        return (down_out, skip_out)


@register
class UpBlock(nn.Module):
    """ConvTranspose2d(k2,s2) -> cat([up, skip]) -> DoubleConv (munet_neck.py:11-49).  The concat is never
    materialised: the following conv reads its K range from the two tensors."""

    def __init__(self, in_channels, out_channels, up_sample_mode):
        super().__init__()
        if up_sample_mode == 'conv_transpose':
            self.up_sample = nn.ConvTranspose2d(in_channels, out_channels, kernel_size=2, stride=2)
        elif up_sample_mode == 'bilinear':
            self.up_sample = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
        else:
            raise ValueError("Unsupported `up_sample_mode` (can take one of `conv_transpose` or `bilinear`)")
        self.up_sample_mode = up_sample_mode
        self.double_conv = DoubleConv(in_channels, out_channels)

    def forward(self, down_input, skip_input, head=None):
        """head: the decoder's `conv_last` (64 -> 2) to fuse behind the block -- then the block returns the head's output
        (N,2,H,W) fp32 instead of its own activation (which is never materialised)."""
        if self.up_sample_mode != 'conv_transpose':
            raise NotImplementedError('bilinear up-sampling has no sm_100a kernel (and is shape-inconsistent in the '
                                      'reference: munet_neck.py:29-33)')
        up = Fn.ConvT2x2Fn.apply(Fn.to_act(down_input), self.up_sample.weight, self.up_sample.bias)
        return self.double_conv.run(up, skip=skip_input, head=head)


def _kaiming_init(module):
    """UNet_encoder.py:86-104 / munet_neck.py:90-110."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Linear):
            nn.init.xavier_normal_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)


@register
class UNet_encoder(nn.Module):
    """UNet_encoder.py:51-158.  forward(x:(B,H,W)) -> (latent, mask uint8 (B,H,W) on x's device, [skip1..4])."""

    def __init__(self, out_classes=2, up_sample_mode='conv_transpose', patch_size=16, mask_ratio=0.65):
        super().__init__()
        self.up_sample_mode = up_sample_mode
        self.down_conv1 = DownBlock(1, 64)
        self.down_conv2 = DownBlock(64, 128)
        self.down_conv3 = DownBlock(128, 256)
        self.down_conv4 = DownBlock(256, 512)
        self.double_conv = DoubleConv(512, 1024)
        self.patch_size = patch_size
        self.mask_ratio = mask_ratio
        self.mask_stream = MaskStream()

    def forward(self, x, want_skips=True):
        """want_skips=False: a caller that only uses the latent (CM_UNet's frozen target path, under no_grad) lets the
        encoder skip the four full-resolution skip stores; the list then holds None entries."""
        ops._need_cuda(x)
        b, s = x.shape[0], x.shape[1]
        mask, k = self.mask_stream.generate(b, s, self.patch_size, self.mask_ratio, x.device)
        x = x.unsqueeze(1)
        x, skip1_out = self.down_conv1(x, mask=mask if k > 0 else None, want_skip=want_skips,     # fused x * (1 - mask[0])  (:156, Q1)
                                       on_backward=self.mask_stream.fire_deferred_prefetch if k > 0 else None)
        x, skip2_out = self.down_conv2(x, want_skip=want_skips)
        x, skip3_out = self.down_conv3(x, want_skip=want_skips)
        x, skip4_out = self.down_conv4(x, want_skip=want_skips)
        x = self.double_conv(x)
        return x, mask, [skip1_out, skip2_out, skip3_out, skip4_out]

    def init_weights(self):
        _kaiming_init(self)

    def create_random_patch_mask(self, batch_size, img_size=256):
        """UNet_encoder.py:106-139 — returns the mask as a numpy array like the reference (device stream, D2H copy)."""
        dev = self.mask_stream.state.device if self.mask_stream.state is not None else torch.device('cuda')
        mask, _ = self.mask_stream.generate(batch_size, img_size, self.patch_size, self.mask_ratio, dev)
        return mask.cpu().numpy()

    def random_masking(self, x):
        """UNet_encoder.py:141-158 (API compatibility; forward() uses the fused first-layer path instead)."""
        mask, _ = self.mask_stream.generate(x.shape[0], x.shape[2], self.patch_size, self.mask_ratio, x.device)
        return x * (1 - mask[0]), mask.cpu().numpy()


@register
class MUNetPretrainDecoder(nn.Module):
    """munet_neck.py:51-110.  forward(latent, [skip1..4]) -> (B, out_classes, H, W) fp32."""

    def __init__(self, out_classes=2, up_sample_mode='conv_transpose', init_cfg=None):
        super().__init__()
        self.up_sample_mode = up_sample_mode
        self.up_conv4 = UpBlock(1024, 512, self.up_sample_mode)
        self.up_conv3 = UpBlock(512, 256, self.up_sample_mode)
        self.up_conv2 = UpBlock(256, 128, self.up_sample_mode)
        self.up_conv1 = UpBlock(128, 64, self.up_sample_mode)
        self.conv_last = nn.Conv2d(64, out_classes, kernel_size=1)
        self.init_cfg = init_cfg

    def forward(self, x, skip):
        x = self.up_conv4(x, skip[3])
        x = self.up_conv3(x, skip[2])
        x = self.up_conv2(x, skip[1])
        if _head_fusable(self.conv_last) and os.environ.get('CMU_NO_HEAD_FUSION') != '1':     # A/B switch
            return self.up_conv1(x, skip[0], head=self.conv_last)
        x = self.up_conv1(x, skip[0])
        return Fn.Head1x1Fn.apply(Fn.to_act(x), self.conv_last.weight, self.conv_last.bias)

    def init_weights(self):
        """munet_neck.py:84-88: only recurses (quirk Q6: the decoder keeps PyTorch's default initialisation)."""
        return None


@register
class NonLinearNeck(nn.Module):
    """nonlinear_neck.py:7-103: fc0 - bn0 - [relu - fc_i - bn_i].  `norm_cfg` type SyncBN -> statistics are all-reduced
    over the default process group when world size > 1."""

    def __init__(self, in_channels, hid_channels, out_channels, num_layers=2, with_bias=False, with_last_bn=True,
                 with_last_bn_affine=True, with_last_bias=False, with_avg_pool=True,
                 norm_cfg=dict(type='SyncBN', eps=1e-6),
                 init_cfg=[dict(type='Constant', val=1, layer=['_BatchNorm', 'GroupNorm'])]):
        super().__init__()
        self.init_cfg = init_cfg
        self.with_avg_pool = with_avg_pool
        self.fc0_factor_exchange = False   # set by CM_UNet for its online projector (see Fn.LinearFn)
        if with_avg_pool:
            self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.relu = nn.ReLU(inplace=True)
        ncfg = dict(norm_cfg)
        self.sync = ncfg.pop('type', 'SyncBN') == 'SyncBN'
        ncfg.pop('requires_grad', None)
        self.fc0 = nn.Linear(in_channels, hid_channels, bias=with_bias)
        self.bn0 = nn.BatchNorm1d(hid_channels, **ncfg)
        self.fc_names, self.bn_names = [], []
        for i in range(1, num_layers):
            last = i == num_layers - 1
            this_channels = out_channels if last else hid_channels
            self.add_module(f'fc{i}', nn.Linear(hid_channels, this_channels, bias=with_last_bias if last else with_bias))
            if not last:
                self.add_module(f'bn{i}', nn.BatchNorm1d(this_channels, **ncfg))
                self.bn_names.append(f'bn{i}')
            elif with_last_bn:
                self.add_module(f'bn{i}', nn.BatchNorm1d(this_channels, **dict(ncfg, affine=with_last_bn_affine)))
                self.bn_names.append(f'bn{i}')
            else:
                self.bn_names.append(None)
            self.fc_names.append(f'fc{i}')

    def _bn(self, bn, x, relu):
        if bn.training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        mom = bn.momentum if bn.momentum is not None else 0.1
        return Fn.BN1dFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.training, mom, bn.eps, relu,
                               self.sync)

    def forward(self, x):
        ops._need_cuda(x)
        if self.with_avg_pool:
            raise NotImplementedError('with_avg_pool=True has no sm_100a kernel (cmunet_config.py uses False)')
        x = x[:, 0, :]
        x = x.reshape(x.size(0), -1)
        x = Fn.LinearFn.apply(x, self.fc0.weight, self.fc0.bias, self.fc0_factor_exchange)
        n_stage = len(self.fc_names)
        x = self._bn(self.bn0, x, relu=n_stage > 0)                     # the ReLU of the first loop turn is fused here
        for i, (fc_name, bn_name) in enumerate(zip(self.fc_names, self.bn_names)):
            fc = getattr(self, fc_name)
            x = Fn.LinearFn.apply(x, fc.weight, fc.bias)
            if bn_name is not None:
                x = self._bn(getattr(self, bn_name), x, relu=i + 1 < n_stage)
            elif i + 1 < n_stage:
                raise NotImplementedError('a ReLU that does not follow a BN layer')
        return x.unsqueeze(dim=1)

    def init_weights(self):
        for m in self.modules():   # init_cfg: Constant(val=1) on _BatchNorm layers
            if isinstance(m, nn.modules.batchnorm._BatchNorm) and m.weight is not None:
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


def _rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


@torch.no_grad()
def concat_all_gather(tensor):
    """cmunet_head.py:9-22."""
    _, world = _rank_world()
    if world == 1:
        return tensor
    parts = [torch.empty_like(tensor) for _ in range(world)]
    dist.all_gather(parts, tensor.contiguous())
    return torch.cat(parts, dim=0)


@register
class CMUNetPretrainHead(nn.Module):
    """cmunet_head.py:25-91.  forward(x, pred_pixel, mask_s, proj_s, proj_t) -> {'loss_ct', 'loss_rc'}."""

    def __init__(self, predictor, temperature=0.07, ct_weight=1.0, rc_weight=1.0, init_cfg=None):
        super().__init__()
        self.predictor = build(predictor)
        self.t = temperature
        self.ct_weight = ct_weight
        self.rc_weight = rc_weight
        self.criterion = nn.CrossEntropyLoss()

    def forward(self, x, pred_pixel, mask_s, proj_s, proj_t):
        ops._need_cuda(x, pred_pixel)
        loss_rc = Fn.MaskedMSEFn.apply(x, pred_pixel, mask_s, self.rc_weight)
        pred_s = self.predictor(proj_s).squeeze(dim=1)
        with torch.no_grad():
            z = ops.l2_normalize_rows(proj_t.squeeze(dim=1).contiguous().float())
            z_all = concat_all_gather(z)
        rank, _ = _rank_world()
        bs = pred_s.size(0)
        loss_ct = Fn.InfoNCEFn.apply(pred_s, z_all, bs * rank, self.t, self.ct_weight)
        return {'loss_ct': loss_ct, 'loss_rc': loss_rc}

    def init_weights(self):
        self.predictor.init_weights()


@register
class CM_UNet(nn.Module):
    """algorithms/cmunet.py:7-135 (+ algorithms/base.py:75-113 mode dispatch)."""

    def __init__(self, backbone, neck, head, base_momentum=0.996, init_cfg=None, target_cls=True,
                 persistent_reduce=False, **kwargs):
        super().__init__()
        assert neck is not None and head is not None
        self.init_cfg = init_cfg
        self.backbone = build(backbone['online'])
        self.target_backbone = build(backbone['target'])
        self.pixel_decoder = build(neck['pixel'])
        self.feature_decoder = build(neck['feature'])
        self.projector = build(neck['projector'])
        self.target_projector = build(neck['projector'])
        self.target_cls = target_cls
        self.head = build(head)
        self.base_momentum = base_momentum
        self.momentum = base_momentum
        for p in self.target_backbone.parameters():
            p.requires_grad = False
        for p in self.target_projector.parameters():
            p.requires_grad = False
        # one numpy-compatible mask stream shared by both encoders, online first (quirk Q2)
        self.target_backbone.mask_stream = self.backbone.mask_stream
        import os
        self.backbone.mask_stream.pair_mode = os.environ.get('CMU_NO_MASK_PREFETCH') != '1'   # A/B switch
        # Q3: the reference draws a fresh, untrained Conv2d(1024,256,1) in every forward_train (cmunet.py:128).
        # persistent_reduce=True (opt-in) keeps the first draw.
        self.persistent_reduce = persistent_reduce
        self._reduce = None
        self._ema_table = None
        import os
        self.multi_stream = os.environ.get('CMU_SINGLE_STREAM') != '1'
        self._aux_streams = None
        # Data parallel, opt-in (CMU_FC0_EXCHANGE=1): projector.fc0.weight holds 90 % of the trainable parameters
        # (1536 x S^2) and its gradient is a rank-(global batch) product.  DistributedDataParallel reads this attribute
        # and leaves the parameter out of its buckets; LinearFn all-gathers the two thin factors instead and builds the
        # averaged gradient on every rank.  Measured (B = 64/GPU, S = 512): -2 ms/step on 2 GPUs, +1 ms on 8 GPUs (the
        # 1.6 GB all-reduce is already hidden behind the backward pass, the all-gather is not), hence off by default.
        if os.environ.get('CMU_FC0_EXCHANGE') == '1':
            self.projector.fc0_factor_exchange = True
            self._ddp_params_and_buffers_to_ignore = ['projector.fc0.weight']

    # ----------------------------------------------------------------------------------- reference API
    def init_weights(self):
        """cmunet.py:61-76 after mmengine's BaseModule recursion over children (online encoder first)."""
        for m in (self.backbone, self.target_backbone, self.pixel_decoder, self.feature_decoder, self.projector,
                  self.target_projector, self.head):
            m.init_weights()
        with torch.no_grad():
            for pb, pm in zip(self.backbone.parameters(), self.target_backbone.parameters()):
                pm.copy_(pb)
                pm.requires_grad = False
            for pb, pm in zip(self.projector.parameters(), self.target_projector.parameters()):
                pm.copy_(pb)
                pm.requires_grad = False

    def _ema_pairs(self):
        pairs = list(zip(self.backbone.parameters(), self.target_backbone.parameters()))
        pairs += list(zip(self.projector.parameters(), self.target_projector.parameters()))
        return pairs

    @torch.no_grad()
    def momentum_update(self):
        """cmunet.py:78-92 as ONE multi-tensor launch, in place (theta_t = theta_t*m + theta_o*(1-m))."""
        pairs = self._ema_pairs()
        key = tuple((pm.data_ptr(), pb.data_ptr(), pm.numel()) for pb, pm in pairs)
        if self._ema_table is None or self._ema_table[0] != key:
            chunk = 1 << 16
            rows = []
            for dst, src, n in key:
                for off in range(0, n, chunk):
                    rows.append((dst + 4 * off, src + 4 * off, min(chunk, n - off)))
            dev = pairs[0][1].device
            self._ema_table = (key, torch.tensor(rows, dtype=torch.int64, device=dev), len(rows))
        ops._need_cuda(pairs[0][1])
        lib.cmu_ema_chunks(self._ema_table[1].data_ptr(), self._ema_table[2], float(self.momentum), ops._stream())

    def extract_feat(self, img):
        return self.backbone(img)

    def _reduce_params(self, device):
        if self._reduce is None or not self.persistent_reduce:
            conv = nn.Conv2d(1024, 256, kernel_size=1)             # same CPU torch-RNG consumption as cmunet.py:128
            w = ops.cast_bf16(conv.weight.detach().to(device).reshape(256, 1024))
            self._reduce = (w, conv.bias.detach().to(device).float().contiguous())
        return self._reduce

    def _target_branch(self, img, img_t):
        """Frozen target path (no grad): target encoder -> fresh 1x1 reduce (Q3) -> NCHW flatten -> target projector."""
        latent_t, _, _ = self.target_backbone(img_t, want_skips=False)
        rw, rb = self._reduce_params(img.device)
        lt = ops.conv1x1_fprop(Fn._nhwc(Fn.to_act(latent_t)), rw, rb)          # (B,h,w,256) act
        b, h, w, c = lt.shape
        flat = torch.empty(b, 1, img.shape[-2], img.shape[-1], dtype=torch.float32, device=img.device)
        assert c * h * w == img.shape[-2] * img.shape[-1]
        lib.cmu_nhwc_to_nchw_f32(lt.data_ptr(), flat.data_ptr(), b, h * w, c, ops._stream())   # :130 NCHW flatten
        return self.target_projector(flat)                     # mean over the single channel is the identity (:131)

    def forward_train(self, img, img_t=None, **kwargs):
        """cmunet.py:108-135.  The three branches that only meet in the head -- target path, pixel decoder, feature
        decoder + projector -- are enqueued on three CUDA streams, so the HBM-bound BatchNorm / element-wise kernels of
        one branch run under the tensor-core kernels of another (autograd replays each branch's backward on its own
        stream).  Host-side call order (mask stream: online first, then target; CPU RNG draw of Q3) is unchanged."""
        ops._need_cuda(img)
        latent_s, mask_s, skip_s = self.backbone(img)
        if not self.multi_stream:
            with torch.no_grad():
                proj_t = self._target_branch(img, img_t)
            pred_pixel = self.pixel_decoder(latent_s, skip_s)
            pred_feature = self.feature_decoder(latent_s, skip_s)
            proj_s = self.projector(Fn.ChannelMean2Fn.apply(pred_feature))
            return self.head(img, pred_pixel[:, 1], mask_s, proj_s, proj_t)
        main = torch.cuda.current_stream()
        if self._aux_streams is None or self._aux_streams[0].device != img.device:
            self._aux_streams = (torch.cuda.Stream(device=img.device), torch.cuda.Stream(device=img.device))
        s_tgt, s_feat = self._aux_streams
        s_tgt.wait_stream(main)
        s_feat.wait_stream(main)
        with torch.cuda.stream(s_tgt), torch.no_grad():
            proj_t = self._target_branch(img, img_t)
        with torch.cuda.stream(s_feat):
            pred_feature = self.feature_decoder(latent_s, skip_s)
            proj_s = self.projector(Fn.ChannelMean2Fn.apply(pred_feature))
        pred_pixel = self.pixel_decoder(latent_s, skip_s)
        main.wait_stream(s_tgt)
        main.wait_stream(s_feat)
        proj_t.record_stream(main)
        proj_s.record_stream(main)
        return self.head(img, pred_pixel[:, 1], mask_s, proj_s, proj_t)

    def forward(self, img, mode='loss', **kwargs):
        if mode == 'tensor':
            return self.extract_feat(img, **kwargs)
        elif mode == 'loss':
            return self.forward_train(img, **kwargs)
        else:
            raise RuntimeError(f'Invalid mode "{mode}".')

    def __getstate__(self):
        d = dict(self.__dict__)
        d['_ema_table'] = None
        d['_aux_streams'] = None
        d['_reduce'] = None if not self.persistent_reduce else d['_reduce']
        return d


def cmunet_config(img_size=224, patch_size=16, mask_ratio=0.65):
    """The model dict of Pretraining/CM-UNet/configs/cmunet_config.py:5-42 with projector.in_channels = S*S."""
    neck_cfg = dict(type='NonLinearNeck', hid_channels=1536, out_channels=256, num_layers=2, with_bias=True,
                    with_last_bn=False, with_avg_pool=False)
    return dict(
        type='CM_UNet',
        backbone=dict(online=dict(type='UNet_encoder', patch_size=patch_size, mask_ratio=mask_ratio),
                      target=dict(type='UNet_encoder', patch_size=patch_size, mask_ratio=0.0)),
        neck=dict(pixel=dict(type='MUNetPretrainDecoder'), feature=dict(type='MUNetPretrainDecoder'),
                  projector=dict(neck_cfg, in_channels=img_size * img_size)),
        head=dict(type='CMUNetPretrainHead', predictor=dict(neck_cfg, in_channels=256), temperature=0.07, ct_weight=1.0,
                  rc_weight=1.0))


def try_register_mmengine():
    """Registers the drop-in classes under the reference's registry names, replacing the reference's own entries.

    The reference builds its model with `cmae.registry.MODELS.build(cfg.model)` (cmae/models/builder.py:12-14); that
    registry is a CHILD of mmengine's root `MODELS` (cmae/registry.py:83-84) and a child's own entries shadow the
    parent's, so the swap has to happen in the child: `import cmae.models` first (the reference classes register
    themselves on import, without `force`), then `register_module(name=..., module=cls, force=True)` there.  mmengine's
    root registry gets the same entries for code that builds through `mmengine.registry.MODELS` directly.
    Returns the list of registries written (empty = neither `cmae` nor `mmengine` is importable)."""
    done = []
    try:
        import cmae.models  # noqa: F401  (must precede the forced registration)
        from cmae.registry import MODELS as CM
        for name, cls in MODELS.items():
            CM.register_module(name=name, module=cls, force=True)
        done.append('cmae.registry.MODELS')
    except ImportError:
        pass
    try:
        from mmengine.registry import MODELS as MM
        for name, cls in MODELS.items():
            MM.register_module(name=name, module=cls, force=True)
        done.append('mmengine.registry.MODELS')
    except ImportError:
        pass
    return done
