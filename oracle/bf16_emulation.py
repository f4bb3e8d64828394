"""ORACLE (test infrastructure; never imported by the product package): the pinned fp32 restatement of the CM-UNet
pretraining step (oracle/cmunet_oracle.py) with bf16 ROUNDING injected at chosen tensors while every arithmetic
operation stays fp32.

Why it exists: BASELINE.json's north star asks for bf16 operands with fp32 accumulation AND gradient cosine > 0.999
against the fp32 reference.  On this model the two conflict for the contrastive branch (all embeddings of a random-init
network are nearly collinear, the InfoNCE gradient is a difference of nearly equal vectors sharpened by 1/tau = 14):
rounding ONLY the convolution weights to bf16 -- the minimum any bf16 tensor-core GEMM does -- already moves the
loss_ct gradients of the encoder to cosine 0.93-0.97 (tools/noise_probe.py, profiles/r2_noise_probe_*.md).  This module
separates "the kernels compute the right function" from "bf16 operands perturb an ill-conditioned gradient":

  * with every switch off it IS the pinned oracle (tests/test_oracle_pinned.py::test_emulation_off_equals_oracle);
  * `faithful()` rounds exactly where the CUDA path stores or feeds bf16 (conv / ConvTranspose
    weights except the 1-channel first conv, raw conv outputs y, activations a, ConvTranspose outputs, the operands of
    the projector fc0 GEMMs when they run on the tensor-core engine, the target path's 1x1 reduce, gradients of
    activations); the CUDA path must agree with THAT to cosine > 0.999 per parameter
    (tests/test_model_gpu.py::test_pretrain_B64_S512_*), which pins every kernel of the step at model level;
  * single switches show which rounding costs what (tools/noise_probe.py).

Reference lines restated: the same as oracle/cmunet_oracle.py (CMU/backbones/UNet_encoder.py:8-158,
CMU/necks/munet_neck.py:11-82, CMU/necks/nonlinear_neck.py:88-103, CMU/algorithms/cmunet.py:108-135)."""
import torch
import torch.nn.functional as F


class _RoundFwd(torch.autograd.Function):
    """round the forward value to bf16, pass the gradient through unchanged"""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    """pass the value, round the gradient to bf16"""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _ReLURound(torch.autograd.Function):
    """relu then (optionally) bf16 rounding as ONE node that saves only its output (keeps the fp32 oracle's memory
    footprint at B = 64 @ 512^2)."""

    @staticmethod
    def forward(ctx, x, rnd):
        r = torch.relu(x)
        if rnd:
            r = r.bfloat16().float()
        ctx.save_for_backward(r)
        return r

    @staticmethod
    def backward(ctx, g):
        (r,) = ctx.saved_tensors
        return g * (r > 0), None


def rnd(x, fwd, bwd):
    if fwd:
        x = _RoundFwd.apply(x)
    if bwd and x.requires_grad:
        x = _RoundBwd.apply(x)
    return x


OFF = {'w': False, 'y': False, 'a': False, 'g': False}
ALL = {'w': True, 'y': True, 'a': True, 'g': True}


def switches(**kw):
    return dict(OFF, **kw)


def double_conv(dc, x, f, update_stats=False):
    """UNet_encoder.py:18-30.  w: conv weights (not the 1-channel first conv: that kernel multiplies in fp32),
    y: raw conv output, a: activation, g: gradients of y and a."""
    seq = dc.double_conv
    for ci, bi in ((0, 1), (3, 4)):
        conv, bn = seq[ci], seq[bi]
        w = rnd(conv.weight, f['w'] and conv.in_channels > 1, False)
        x = F.conv2d(x, w, conv.bias, padding=1)
        x = rnd(x, f['y'], f['g'])
        if update_stats:
            x = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, training=True, momentum=bn.momentum,
                             eps=bn.eps)
        else:
            x = F.batch_norm(x, None, None, bn.weight, bn.bias, training=True, eps=bn.eps)
        x = _ReLURound.apply(x, f['a'])
        x = rnd(x, False, f['g'])
    return x


def encoder(enc, x, mask0, f):
    """UNet_encoder.py:76-84,141-158 (Q1: image 0's mask for the whole batch)."""
    x = x.unsqueeze(1) * (1 - mask0)
    skips = []
    for i in range(4):
        s = double_conv(getattr(enc, f'down_conv{i + 1}').double_conv, x, f)
        skips.append(s)
        x = F.max_pool2d(s, 2)
    return double_conv(enc.double_conv, x, f), skips


def decoder(dec, x, skips, f):
    """munet_neck.py:74-82; conv_last multiplies fp32 weights (csrc/losses.cu head1x1)."""
    for i in (4, 3, 2, 1):
        blk = getattr(dec, f'up_conv{i}')
        up = F.conv_transpose2d(x, rnd(blk.up_sample.weight, f['w'], False), blk.up_sample.bias, stride=2)
        up = rnd(up, f['a'], f['g'])
        x = double_conv(blk.double_conv, torch.cat([up, skips[i - 1]], 1), f)
    return F.conv2d(x, dec.conv_last.weight, dec.conv_last.bias)


def neck(nk, x, fc0_bf16):
    """nonlinear_neck.py:88-103 (single rank).  fc0_bf16: both fc0 operands (and, in backward, dy) are bf16."""
    x = x[:, 0, :].reshape(x.size(0), -1)
    x = F.linear(rnd(x, fc0_bf16, False), rnd(nk.fc0.weight, fc0_bf16, False), nk.fc0.bias)
    x = rnd(x, False, fc0_bf16)
    x = F.batch_norm(x, None, None, nk.bn0.weight, nk.bn0.bias, training=True, eps=nk.bn0.eps)
    return F.linear(torch.relu(x), nk.fc1.weight).unsqueeze(1)


def forward_train(o, img, img_t, mask, reduce_w, reduce_b, f_enc=OFF, f_dec=OFF, f_tgt=OFF, fc0_s=False, fc0_t=False):
    """cmunet.py:108-135 on the parameters of an `OracleCMUNet` `o`.  mask: (B,S,S) uint8 numpy array (the online
    call's patch mask).  Does not touch BN running statistics or the oracle's RNG stream."""
    mask_t = torch.from_numpy(mask).to(img.device)
    latent_s, skips = encoder(o.backbone, img, mask_t[0], f_enc)
    with torch.no_grad():
        latent_t, _ = encoder(o.target_backbone, img_t, torch.zeros_like(mask_t[0]), f_tgt)
    pred_pixel = decoder(o.pixel_decoder, latent_s, skips, f_dec)
    pred_feature = decoder(o.feature_decoder, latent_s, skips, f_dec)
    proj_s = neck(o.projector, pred_feature.mean(dim=1, keepdim=True), fc0_s)
    with torch.no_grad():
        lt = F.conv2d(latent_t, rnd(reduce_w, f_tgt['w'], False), reduce_b)            # :128-129 (Q3)
        lt = rnd(lt, f_tgt['a'], False)
        B, S = img.shape[0], img.shape[-1]
        proj_t = neck(o.target_projector, lt.reshape(B, 1, img.shape[-2], S), fc0_t)   # :130-131
    return o.head(img, pred_pixel[:, 1], mask_t, proj_s, proj_t)                       # :133


def faithful(batch):
    """keyword arguments of `forward_train` that place the roundings where the CUDA path has them.  The projector fc0
    GEMMs use the bf16 tensor-core engine only for batches that are multiples of 64 (ops.tc_linear_ok), else fp32."""
    tc = batch % 64 == 0 and batch <= 128
    return dict(f_enc=ALL, f_dec=ALL, f_tgt=ALL, fc0_s=tc, fc0_t=tc)
