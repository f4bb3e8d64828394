"""ORACLE (test infrastructure; never imported by the product package): plain fp32 torch restatement of
the reference's CM-UNet pretraining step and fine-tuning UNet + losses.  It is the checker for the CUDA
path in tests/, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs.

Parity status: PINNED — tests/test_oracle_pinned.py compares this file against golden vectors minted by
executing the unmodified reference in the build container (oracle/make_goldens.py -> tests/golden/*.json).

Each function cites the reference lines it restates (paths relative to /root/reference):
  CMU = Pretraining/CM-UNet/cmae/models, FT = Finetuning.
State-dict keys, parameter creation order (hence RNG consumption under a fixed torch seed) and the
reference quirks Q1-Q8 are reproduced on purpose.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .mask_oracle import MT19937, patch_mask

ENC_CH = (1, 64, 128, 256, 512, 1024)


# ----------------------------------------------------------------------------------------------
# parameter containers (same key names as the reference: `double_conv.double_conv.{0,1,3,4}.*`)
# ----------------------------------------------------------------------------------------------
def _cbr_pair(cin, cout):
    """CMU/backbones/UNet_encoder.py:18-27 == FT/model.py:16-23."""
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(),
                         nn.Conv2d(cout, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU())


class _DC(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.double_conv = _cbr_pair(cin, cout)


class _Down(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.double_conv = _DC(cin, cout)


class _Up(nn.Module):
    """CMU/necks/munet_neck.py:25-33: ConvTranspose2d(k2,s2) then DoubleConv(in,out)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.up_sample = nn.ConvTranspose2d(cin, cout, 2, stride=2)
        self.double_conv = _DC(cin, cout)


def double_conv_fwd(dc, x):
    """Conv3x3(p1,bias) -> BatchNorm2d (train: batch stats, updates running stats) -> ReLU, twice.
    CMU/backbones/UNet_encoder.py:29-30."""
    seq = dc.double_conv
    for ci, bi in ((0, 1), (3, 4)):
        conv, bn = seq[ci], seq[bi]
        x = F.conv2d(x, conv.weight, conv.bias, padding=1)
        x = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias,
                         training=bn.training, momentum=bn.momentum, eps=bn.eps)
        if bn.training:
            bn.num_batches_tracked += 1
        x = torch.relu(x)
    return x


class OracleEncoder(nn.Module):
    """CMU/backbones/UNet_encoder.py:51-158."""

    def __init__(self, patch_size=16, mask_ratio=0.65, rng=None):
        super().__init__()
        for i in range(4):
            setattr(self, f'down_conv{i + 1}', _Down(ENC_CH[i], ENC_CH[i + 1]))
        self.double_conv = _DC(ENC_CH[4], ENC_CH[5])
        self.patch_size, self.mask_ratio = patch_size, mask_ratio
        self.rng = rng  # oracle.mask_oracle.MT19937 shared by online+target (global numpy stream, Q2)

    def init_weights(self):
        """UNet_encoder.py:86-104 (module traversal order of nn.Module.apply)."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        B, S = x.shape[0], x.shape[1]
        mask, _ = patch_mask(self.rng, B, S, self.patch_size, self.mask_ratio)  # :106-139
        mask_t = torch.from_numpy(mask).to(x.device)
        x = x.unsqueeze(1) * (1 - mask_t[0])          # :77, :156  (Q1: image 0's mask for the batch)
        skips = []
        for i in range(4):
            blk = getattr(self, f'down_conv{i + 1}')
            s = double_conv_fwd(blk.double_conv, x)    # :47
            skips.append(s)
            x = F.max_pool2d(s, 2)                     # :48
        x = double_conv_fwd(self.double_conv, x)       # :83
        return x, mask_t, skips


class OracleDecoder(nn.Module):
    """CMU/necks/munet_neck.py:51-82 (PyTorch default init is kept, quirk Q6)."""

    def __init__(self, out_classes=2):
        super().__init__()
        for i in (4, 3, 2, 1):
            setattr(self, f'up_conv{i}', _Up(ENC_CH[i + 1], ENC_CH[i]))
        self.conv_last = nn.Conv2d(64, out_classes, 1)

    def forward(self, x, skips):
        for i in (4, 3, 2, 1):
            blk = getattr(self, f'up_conv{i}')
            up = F.conv_transpose2d(x, blk.up_sample.weight, blk.up_sample.bias, stride=2)  # :46
            x = double_conv_fwd(blk.double_conv, torch.cat([up, skips[i - 1]], dim=1))       # :48-49
        return F.conv2d(x, self.conv_last.weight, self.conv_last.bias)                       # :81


class OracleNeck(nn.Module):
    """CMU/necks/nonlinear_neck.py:35-103 with the config of configs/cmunet_config.py:18-38:
    x[:,0,:] -> flatten -> fc0(+bias) -> (Sync)BN(eps 1e-6) -> ReLU -> fc1(no bias).
    sync_bn: nn.SyncBatchNorm semantics across ranks (what the reference runs on GPUs); False reproduces the CPU shim
    of oracle/ref_loader.py (unsynced BatchNorm1d), which is what the 2-rank CPU golden was minted with."""
    sync_bn = True

    def __init__(self, in_channels, hid_channels=1536, out_channels=256):
        super().__init__()
        self.fc0 = nn.Linear(in_channels, hid_channels, bias=True)
        self.bn0 = nn.BatchNorm1d(hid_channels, eps=1e-6)
        self.fc1 = nn.Linear(hid_channels, out_channels, bias=False)

    def forward(self, x):
        x = x[:, 0, :].reshape(x.size(0), -1)
        x = F.linear(x, self.fc0.weight, self.fc0.bias)
        bn = self.bn0
        if self.sync_bn and bn.training and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            x = _sync_bn_train(bn, x)
        else:
            x = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias,
                             training=bn.training, momentum=bn.momentum, eps=bn.eps)
            if bn.training:
                bn.num_batches_tracked += 1
        x = F.linear(torch.relu(x), self.fc1.weight)
        return x.unsqueeze(1)


def _sync_bn_train(bn, x):
    """nn.SyncBatchNorm train-mode semantics restated with autograd-aware all_reduce (world > 1)."""
    import torch.distributed as td
    import torch.distributed.nn.functional as tdf
    n_local = x.shape[0]
    W = td.get_world_size()
    s1 = tdf.all_reduce(x.sum(0))
    s2 = tdf.all_reduce((x * x).sum(0))
    n = n_local * W
    mean = s1 / n
    var = s2 / n - mean * mean
    with torch.no_grad():
        bn.running_mean.mul_(1 - bn.momentum).add_(bn.momentum * mean)
        bn.running_var.mul_(1 - bn.momentum).add_(bn.momentum * var * n / (n - 1))
        bn.num_batches_tracked += 1
    return (x - mean) * torch.rsqrt(var + bn.eps) * bn.weight + bn.bias


class OracleHead(nn.Module):
    """CMU/heads/cmunet_head.py:26-91."""

    def __init__(self, temperature=0.07, ct_weight=1.0, rc_weight=1.0):
        super().__init__()
        self.predictor = OracleNeck(256, 1536, 256)
        self.t, self.ct_weight, self.rc_weight = temperature, ct_weight, rc_weight

    def forward(self, x, pred_pixel, mask_s, proj_s, proj_t):
        with torch.no_grad():                                            # :64-67 (Q8)
            mean = x.mean(dim=-1, keepdim=True)
            var = x.var(dim=-1, keepdim=True)
            target = (x - mean) / (var + 1.e-6) ** .5
        loss_rc = (((pred_pixel - target) ** 2) * mask_s).sum() / mask_s.sum()   # :69-70
        p = F.normalize(self.predictor(proj_s).squeeze(1), dim=1, p=2)           # :72-74
        z = F.normalize(proj_t.squeeze(1), dim=1, p=2)                           # :75
        rank, Z = 0, z
        import torch.distributed as td
        if td.is_available() and td.is_initialized() and td.get_world_size() > 1:  # :9-22, :77
            with torch.no_grad():
                parts = [torch.empty_like(z) for _ in range(td.get_world_size())]
                td.all_gather(parts, z.contiguous())
                Z = torch.cat(parts, 0)
            rank = td.get_rank()
        score = p @ Z.detach().t() / self.t                                      # :79-81
        bs = score.size(0)
        label = torch.arange(bs, dtype=torch.long, device=score.device) + bs * rank   # :83-85
        return {'loss_ct': self.ct_weight * 2 * self.t * F.cross_entropy(score, label),   # :88
                'loss_rc': self.rc_weight * loss_rc}                                       # :89


class OracleCMUNet(nn.Module):
    """CMU/algorithms/cmunet.py:20-135 built with configs/cmunet_config.py:5-42 (projector.in_channels
    generalised to S*S).  Module creation order == reference (backbone, target_backbone, pixel_decoder,
    feature_decoder, projector, target_projector, head) so that a fixed torch seed yields identical
    initial weights."""

    def __init__(self, img_size=224, base_momentum=0.996, np_seed=None, mask_ratio=0.65, patch_size=16):
        super().__init__()
        self.rng = MT19937(np_seed)
        self.backbone = OracleEncoder(patch_size, mask_ratio, self.rng)
        self.target_backbone = OracleEncoder(patch_size, 0.0, self.rng)
        self.pixel_decoder = OracleDecoder()
        self.feature_decoder = OracleDecoder()
        self.projector = OracleNeck(img_size * img_size)
        self.target_projector = OracleNeck(img_size * img_size)
        self.head = OracleHead()
        self.base_momentum = self.momentum = base_momentum
        for p in list(self.target_backbone.parameters()) + list(self.target_projector.parameters()):
            p.requires_grad = False

    def init_weights(self):
        """cmunet.py:61-76 + mmengine BaseModule recursion: Kaiming on both encoders (online first),
        decoders/necks keep default init (Q6), then target <- online."""
        self.backbone.init_weights()
        self.target_backbone.init_weights()
        with torch.no_grad():
            for po, pt in zip(self.backbone.parameters(), self.target_backbone.parameters()):
                pt.copy_(po)
            for po, pt in zip(self.projector.parameters(), self.target_projector.parameters()):
                pt.copy_(po)

    @torch.no_grad()
    def momentum_update(self):
        """cmunet.py:78-92 — parameters only (BN buffers of the target are not averaged)."""
        m = self.momentum
        for src, dst in ((self.backbone, self.target_backbone), (self.projector, self.target_projector)):
            for po, pt in zip(src.parameters(), dst.parameters()):
                pt.data = pt.data * m + po.data * (1. - m)

    def forward_train(self, img, img_t, reduce_weight=None, reduce_bias=None):
        """cmunet.py:108-135.  Q3: the reference draws a fresh nn.Conv2d(1024,256,1) from the torch RNG on
        every call; pass (reduce_weight, reduce_bias) to inject a known draw, else one is drawn here the
        same way (same RNG consumption)."""
        latent_s, mask_s, skip_s = self.backbone(img)
        latent_t, _, _ = self.target_backbone(img_t)
        pred_pixel = self.pixel_decoder(latent_s, skip_s)
        pred_feature = self.feature_decoder(latent_s, skip_s)
        proj_s = self.projector(pred_feature.mean(dim=1, keepdim=True))       # :126
        if reduce_weight is None:
            rc = nn.Conv2d(1024, 256, kernel_size=1).to(latent_t.dtype).to(latent_t.device)   # :128
            reduce_weight, reduce_bias = rc.weight, rc.bias
        latent_t = F.conv2d(latent_t, reduce_weight, reduce_bias)              # :129
        B, S = img.shape[0], img.shape[-1]
        latent_t = latent_t.reshape(B, -1).reshape(B, 1, img.shape[-2], S)      # :130 (224 -> S)
        proj_t = self.target_projector(latent_t.mean(dim=1, keepdim=True))     # :131
        return self.head(img, pred_pixel[:, 1], mask_s, proj_s, proj_t)        # :133

    def forward(self, img, mode='loss', **kw):
        """algorithms/base.py:75-113."""
        if mode == 'loss':
            return self.forward_train(img, **kw)
        if mode == 'tensor':
            return self.backbone(img)
        raise RuntimeError(f'Invalid mode "{mode}".')


# ----------------------------------------------------------------------------------------------
# fine-tuning path (FT/model.py:84-131, FT/metrics.py)
# ----------------------------------------------------------------------------------------------
class OracleUNet(nn.Module):
    def __init__(self, out_classes=2):
        super().__init__()
        for i in range(4):
            setattr(self, f'down_conv{i + 1}', _Down(ENC_CH[i], ENC_CH[i + 1]))
        self.double_conv = _DC(ENC_CH[4], ENC_CH[5])
        for i in (4, 3, 2, 1):
            setattr(self, f'up_conv{i}', _Up(ENC_CH[i + 1], ENC_CH[i]))
        self.conv_last = nn.Conv2d(64, out_classes, 1)

    def forward(self, x):
        x = x.unsqueeze(1)                                                       # model.py:120
        skips = []
        for i in range(4):
            s = double_conv_fwd(getattr(self, f'down_conv{i + 1}').double_conv, x)
            skips.append(s)
            x = F.max_pool2d(s, 2)
        x = double_conv_fwd(self.double_conv, x)
        return OracleDecoder.forward(self, x, skips)                            # model.py:126-131


def dice_loss(logits, gt, eps=1e-5, beta=1.0, threshold=0.5, ignore_channels=(0,)):
    """FT/metrics.py:135-180 with activation='softmax' (implicit dim=1, Q7): non-differentiable."""
    pr = (torch.softmax(logits, dim=1) > threshold).type(logits.dtype)
    keep = [c for c in range(pr.shape[1]) if c not in ignore_channels]
    pr, gt = pr[:, keep], gt[:, keep]
    tp = torch.sum(gt * pr)
    fp = torch.sum(pr) - tp
    fn = torch.sum(gt) - tp
    b2 = beta ** 2
    return 1 - ((1 + b2) * tp + eps) / ((1 + b2) * tp + b2 * fn + fp + eps)


def iou_loss(logits, gt, eps=1e-7, threshold=0.5, ignore_channels=(0,), activation='softmax'):
    """FT/metrics.py:182-220."""
    pr = torch.softmax(logits, dim=1) if activation == 'softmax' else logits
    pr = (pr > threshold).type(logits.dtype)
    keep = [c for c in range(pr.shape[1]) if c not in ignore_channels]
    pr, gt = pr[:, keep], gt[:, keep]
    inter = torch.sum(gt * pr)
    union = torch.sum(gt) + torch.sum(pr) - inter + eps
    return 1 - (inter + eps) / union


def ce_prob_loss(logits, gt):
    """FT/metrics.py:503-504: nn.CrossEntropyLoss with float64 probability targets."""
    return F.cross_entropy(logits, gt)


def soft_skel(img, num_iter=10):
    """FT/metrics.py:447-492 (SoftSkeletonize, 2-D case): min/max-pool morphology, padding ignored by the pools."""
    def erode(t):
        return torch.min(-F.max_pool2d(-t, (3, 1), (1, 1), (1, 0)), -F.max_pool2d(-t, (1, 3), (1, 1), (0, 1)))

    def opened(t):
        return F.max_pool2d(erode(t), (3, 3), (1, 1), (1, 1))

    skel = F.relu(img - opened(img))
    for _ in range(num_iter):
        img = erode(img)
        delta = F.relu(img - opened(img))
        skel = skel + F.relu(delta - skel * delta)
    return skel


def cldice_loss(logits, gt, smooth=1.0, threshold=0.5, ignore_channels=(0,), num_iter=10):
    """FT/metrics.py:401-430 (soft_cldice with activation='softmax'): thresholded prediction vs float64 target."""
    pr = (torch.softmax(logits, dim=1) > threshold).type(logits.dtype)
    keep = [c for c in range(pr.shape[1]) if c not in ignore_channels]
    pr, gt = pr[:, keep], gt[:, keep]
    skel_pred, skel_true = soft_skel(pr, num_iter), soft_skel(gt, num_iter)
    tprec = (torch.sum(skel_pred * gt) + smooth) / (torch.sum(skel_pred) + smooth)
    tsens = (torch.sum(skel_true * pr) + smooth) / (torch.sum(skel_true) + smooth)
    return 1. - 2.0 * (tprec * tsens) / (tprec + tsens)


def cldice_inputs(n, h, w, seed):
    """Deterministic vessel-like test inputs: logits (n,2,h,w) fp32 and a one-hot float64 target from smoothed noise."""
    g = torch.Generator().manual_seed(seed)
    a = F.avg_pool2d(torch.rand(n, 1, h + 8, w + 8, generator=g), 9, 1)           # (n,1,h,w) smooth field
    b = F.avg_pool2d(torch.rand(n, 1, h + 8, w + 8, generator=g), 9, 1)
    y1 = ((a - 0.5).abs() < 0.012)                                                # thin iso-contours
    logits = torch.cat([torch.zeros(n, 1, h, w), 40.0 * (0.02 - (a - 0.5 + 0.05 * (b - 0.5)).abs())], 1)
    logits = logits + 0.05 * torch.randn(n, 2, h, w, generator=g)
    return logits.float(), torch.cat([~y1, y1], 1).double()


# ----------------------------------------------------------------------------------------------
# summaries used by the golden fixtures (small, deterministic fingerprints of big tensors)
# ----------------------------------------------------------------------------------------------
def fingerprint(t, n=8):
    """(l2 norm, sum, n samples at fixed pseudo-random flat indices) of a tensor, as python floats."""
    t = t.detach().double().flatten().cpu()
    g = np.random.RandomState(t.numel() % 100003)
    idx = g.randint(0, t.numel(), size=min(n, t.numel()))
    return {'norm': float(t.norm()), 'sum': float(t.sum()), 'idx': [int(i) for i in idx],
            'val': [float(t[int(i)]) for i in idx]}


def synthetic_batch(B, S, seed=1):
    """SURVEY §8d config 2 inputs: img = randn(B,S,S; seed), img_t = img + 0.1*randn."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, S, S, generator=g)
    img_t = img + 0.1 * torch.randn(B, S, S, generator=g)
    return img, img_t
