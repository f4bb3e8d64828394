"""torch.autograd.Function wrappers: each fused op of the hot path, forward and backward, as calls into the C ABI.

Public tensors keep the reference's NCHW *shape*; activations travel as bf16 tensors whose memory is NHWC
(channels-last strides), so `t.permute(0, 2, 3, 1)` is the contiguous (N,H,W,C) buffer the kernels use and nothing
is ever transposed between layers.  Backward of train-mode BatchNorm follows SURVEY.md Appendix C; the conv bias in
front of a train-mode BN has an analytically zero gradient (returned as zeros)."""
import torch
import torch.distributed as dist

from . import ops
from ._lib import CmuError

BF16 = torch.bfloat16


def to_act(x):
    """NCHW-shaped tensor -> NCHW-shaped bf16 tensor with NHWC memory (zero-copy when it already is one)."""
    if x.dtype == BF16 and x.dim() == 4 and x.permute(0, 2, 3, 1).is_contiguous():
        return x
    ops._need_cuda(x)
    return x.to(BF16).contiguous(memory_format=torch.channels_last) if x.shape[1] > 1 else \
        x.to(BF16).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


def _nhwc(x):
    """NCHW-shaped channels-last bf16 tensor -> contiguous (N,H,W,C) view."""
    t = x.permute(0, 2, 3, 1)
    if not t.is_contiguous():
        t = t.contiguous()
    return t


def _nchw_view(a):
    return a.permute(0, 3, 1, 2)


class BNConfig:
    __slots__ = ('training', 'momentum', 'eps', 'pool', 'grad', 'want_act', 'on_backward')

    def __init__(self, training, momentum, eps, pool, want_act=True):
        self.training, self.momentum, self.eps, self.pool = training, momentum, eps, pool
        self.grad = torch.is_grad_enabled()     # captured at call time (grad mode is always off inside Function.forward)
        # pooled layers whose full-resolution activation (the skip tensor) nobody reads: skip its store (backward
        # recomputes ReLU mask and pool routing from y, so the tensor is not needed for autograd either)
        self.want_act = want_act or not pool
        self.on_backward = None                 # callable run once this node's backward kernels have been enqueued


def _conv_bias_grad(dy, bn_training):
    """Gradient of the conv bias in front of a BatchNorm: analytically zero under batch statistics (the mean removes any
    per-channel constant); with frozen / eval-mode statistics it is the per-channel sum of dy."""
    n, h, w, c = dy.shape
    if bn_training:
        return torch.zeros(c, dtype=torch.float32, device=dy.device)
    return ops.colsum_bf16(n * h * w, c, dy)


class ConvBNReLUFn(torch.autograd.Function):
    """Conv3x3(pad 1) -> BatchNorm2d -> ReLU [-> MaxPool2d(2)] on (x0 | x1) (channel concat, x1 optional).
    UNet_encoder.py:18-30,44-49 / munet_neck.py:48-49."""

    @staticmethod
    def forward(ctx, x0, x1, weight, bias, gamma, beta, running_mean, running_var, cfg):
        a0 = _nhwc(x0)
        a1 = _nhwc(x1) if x1 is not None else None
        need_grad = cfg.grad and any(ctx.needs_input_grad)
        wf, wd = ops.pack_conv3x3(weight, need_dgrad=need_grad)
        y, stats = ops.conv3x3_fprop(a0, a1, wf, want_stats=cfg.training)
        scale, shift, mean, rstd = ops.bn_finalize(stats, gamma, beta, bias, running_mean, running_var, cfg.momentum,
                                                   cfg.eps, cfg.training)
        act, pooled = ops.bn_relu_apply(y, scale, shift, cfg.pool, cfg.want_act)
        if need_grad:
            ctx.save_for_backward(a0, a1, y, scale, shift, mean, rstd, wd)
            ctx.c0 = a0.shape[3]
            ctx.c1 = 0 if a1 is None else a1.shape[3]
        ctx.pool, ctx.bn_training = cfg.pool, cfg.training
        if cfg.pool:
            return (_nchw_view(act) if act is not None else None), _nchw_view(pooled)
        return _nchw_view(act)

    @staticmethod
    def backward(ctx, d_act, d_pool=None):
        a0, a1, y, scale, shift, mean, rstd, wd = ctx.saved_tensors
        da = _nhwc(to_act(d_act)) if d_act is not None else None
        dp = _nhwc(to_act(d_pool)) if (ctx.pool and d_pool is not None) else None
        if da is None and dp is None:
            return (None,) * 9
        if da is None and not ctx.pool:
            return (None,) * 9
        dy, dgamma, dbeta = ops.bn_relu_bwd(da, dp, y, scale, shift, mean, rstd, ctx.bn_training)
        dx0 = dx1 = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            g0, g1 = ops.conv3x3_dgrad(dy, wd, ctx.c0, ctx.c1)
            dx0 = _nchw_view(g0) if ctx.needs_input_grad[0] else None
            dx1 = _nchw_view(g1) if (g1 is not None and ctx.needs_input_grad[1]) else None
        dw = ops.conv3x3_wgrad(a0, a1, dy) if ctx.needs_input_grad[2] else None
        dbias = _conv_bias_grad(dy, ctx.bn_training) if ctx.needs_input_grad[3] else None
        return dx0, dx1, dw, dbias, dgamma, dbeta, None, None, None


class ConvBNReLUHeadFn(torch.autograd.Function):
    """Conv3x3(pad 1) -> BatchNorm2d -> ReLU -> Conv2d(64, 2, 1): the last DoubleConv layer of a decoder fused with
    `conv_last` (munet_neck.py:48-49,72,81 == FT/model.py:79-81,131).  The activated 64-channel tensor and its gradient
    never reach HBM: BN apply + ReLU run in the prologue of the 1x1 head (forward) and of the BatchNorm backward passes,
    which rebuild da = W_head^T dout on the fly (csrc/head_fused.cu).  -> (N,2,H,W) fp32."""

    @staticmethod
    def forward(ctx, x0, x1, weight, bias, gamma, beta, running_mean, running_var, head_w, head_b, cfg):
        a0 = _nhwc(x0)
        a1 = _nhwc(x1) if x1 is not None else None
        need_grad = cfg.grad and any(ctx.needs_input_grad)
        wf, wd = ops.pack_conv3x3(weight, need_dgrad=need_grad)
        y, stats = ops.conv3x3_fprop(a0, a1, wf, want_stats=cfg.training)
        scale, shift, mean, rstd = ops.bn_finalize(stats, gamma, beta, bias, running_mean, running_var, cfg.momentum,
                                                   cfg.eps, cfg.training)
        out = ops.bn_relu_head_fwd(y, scale, shift, head_w, head_b)
        if need_grad:
            ctx.save_for_backward(a0, a1, y, scale, shift, mean, rstd, wd, head_w)
            ctx.c0 = a0.shape[3]
            ctx.c1 = 0 if a1 is None else a1.shape[3]
        ctx.bn_training = cfg.training
        return out

    @staticmethod
    def backward(ctx, d_out):
        a0, a1, y, scale, shift, mean, rstd, wd, head_w = ctx.saved_tensors
        dy, dgamma, dbeta, dhw, dhb = ops.bn_relu_head_bwd(y, scale, shift, mean, rstd, head_w, d_out, ctx.bn_training)
        dx0 = dx1 = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            g0, g1 = ops.conv3x3_dgrad(dy, wd, ctx.c0, ctx.c1)
            dx0 = _nchw_view(g0) if ctx.needs_input_grad[0] else None
            dx1 = _nchw_view(g1) if (g1 is not None and ctx.needs_input_grad[1]) else None
        dw = ops.conv3x3_wgrad(a0, a1, dy) if ctx.needs_input_grad[2] else None
        dbias = _conv_bias_grad(dy, ctx.bn_training) if ctx.needs_input_grad[3] else None
        return (dx0, dx1, dw, dbias, dgamma, dbeta, None, None, dhw.reshape(head_w.shape).clone(), dhb.clone(), None)


class FirstConvBNReLUFn(torch.autograd.Function):
    """Cin = 1 first layer: (x * (1 - mask[0])) -> Conv3x3 -> BN -> ReLU.  x: (N,H,W) fp32; mask: (B,H,W) uint8 or
    None (UNet_encoder.py:77,156; quirk Q1)."""

    @staticmethod
    def forward(ctx, x, mask, weight, bias, gamma, beta, running_mean, running_var, cfg):
        x = x.contiguous().float()
        y, stats = ops.conv3x3_c1_fprop(x, mask, weight, want_stats=cfg.training)
        scale, shift, mean, rstd = ops.bn_finalize(stats, gamma, beta, bias, running_mean, running_var, cfg.momentum,
                                                   cfg.eps, cfg.training)
        act, _ = ops.bn_relu_apply(y, scale, shift, False)
        if cfg.grad and any(ctx.needs_input_grad):
            ctx.save_for_backward(x, mask, y, scale, shift, mean, rstd)
        ctx.bn_training = cfg.training
        ctx.on_backward = cfg.on_backward
        return _nchw_view(act)

    @staticmethod
    def backward(ctx, d_act):
        x, mask, y, scale, shift, mean, rstd = ctx.saved_tensors
        da = _nhwc(to_act(d_act))
        dy, dgamma, dbeta = ops.bn_relu_bwd(da, None, y, scale, shift, mean, rstd, ctx.bn_training)
        dw = ops.conv3x3_c1_wgrad(x, mask, dy) if ctx.needs_input_grad[2] else None
        dbias = _conv_bias_grad(dy, ctx.bn_training) if ctx.needs_input_grad[3] else None
        if ctx.on_backward is not None:
            ctx.on_backward()        # last node of the step's backward (e.g. the mask stream's deferred prefetch)
        # the input image never needs a gradient on this path (f1 "not needed")
        return None, None, dw, dbias, dgamma, dbeta, None, None, None


class ConvT2x2Fn(torch.autograd.Function):
    """ConvTranspose2d(k=2, s=2) + bias (munet_neck.py:28,46)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        a = _nhwc(x)
        need_grad = any(ctx.needs_input_grad)
        wf, wd = ops.pack_convT2x2(weight, need_dgrad=need_grad)
        y = ops.convT2x2_fprop(a, wf, bias.detach().float().contiguous() if bias is not None else None)
        if need_grad:
            ctx.save_for_backward(a, wd)
        return _nchw_view(y)

    @staticmethod
    def backward(ctx, d_y):
        a, wd = ctx.saved_tensors
        dy = _nhwc(to_act(d_y))
        dx = _nchw_view(ops.convT2x2_dgrad(dy, wd)) if ctx.needs_input_grad[0] else None
        dw = ops.convT2x2_wgrad(a, dy) if ctx.needs_input_grad[1] else None
        db = None
        if ctx.needs_input_grad[2]:
            n, h, w, c = dy.shape
            # (taking these sums in the epilogue of the dgrad that produces dy -- ops.conv3x3_dgrad(want_colsum0=True) --
            # was measured: the short-K dual-output dgrads are epilogue-bound, +2.9 ms/step for the 1.5 ms saved here)
            db = ops.colsum_bf16(n * h * w, c, dy)
        return dx, dw, db


class Head1x1Fn(torch.autograd.Function):
    """conv_last: Conv2d(64, 2, 1) on the activated act tensor -> (N,2,H,W) fp32 (munet_neck.py:72,81)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        a = _nhwc(x)
        out = ops.head1x1_fprop(a, weight, bias)
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(a, weight)
        return out

    @staticmethod
    def backward(ctx, d_out):
        a, weight = ctx.saved_tensors
        da, dw, db = ops.head1x1_bwd(a, weight, d_out)
        return _nchw_view(da), dw.reshape(weight.shape).clone(), db.clone()


class ChannelMean2Fn(torch.autograd.Function):
    """torch.mean(x, dim=1, keepdim=True) for the 2-channel decoder output (cmunet.py:126)."""

    @staticmethod
    def forward(ctx, x):
        n, c, h, w = x.shape
        assert c == 2
        x = x.contiguous().float()
        y = torch.empty(n, 1, h, w, dtype=torch.float32, device=x.device)
        ops.lib.cmu_channel_mean2(x.data_ptr(), y.data_ptr(), n, h * w, ops._stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        n, _, h, w = dy.shape
        dy = dy.contiguous().float()
        dx = torch.empty(n, 2, h, w, dtype=torch.float32, device=dy.device)
        ops.lib.cmu_channel_mean2_bwd(dy.data_ptr(), dx.data_ptr(), n, h * w, ops._stream())
        return dx


class LinearFn(torch.autograd.Function):
    """nn.Linear on a (M,K) fp32 matrix (nonlinear_neck.py:94,99).  Large layers (projector.fc0: K = S*S) run on the
    tcgen05 engine in bf16 with fp32 accumulation: y = (x^T)^T W^T, dx = (dy^T)^T W, dW = dy^T x are all "reduce over
    matrix rows" GEMMs once the small operand is transposed; small layers use the fp32 SIMT SGEMM."""

    @staticmethod
    def forward(ctx, x, weight, bias, exchange=False):
        """exchange=True (data parallel, weight excluded from DDP's all-reduce): the weight gradient of a linear layer is
        the rank-(global batch) product dy^T x, so the ranks all-gather the two thin factors (M x N and M x K) and each
        computes the already-averaged dW locally -- for projector.fc0 (1536 x S^2 = 1.6 GB of fp32 gradient at S = 512)
        that replaces a 1.6 GB all-reduce by a 34 MB/rank all-gather; the result is identical on every rank."""
        ctx.exchange = bool(exchange) and _world() > 1
        x = x.contiguous().float()
        w = weight.detach().contiguous().float()
        m, k = x.shape
        n = w.shape[0]
        b = bias.detach().contiguous().float() if bias is not None else None
        ctx.tc = ops.tc_linear_ok(m, k, n)
        if ctx.tc:
            need_dx = ctx.needs_input_grad[0]
            wt16, w16 = ops.transpose_cast(w, also_plain=need_dx)          # (K,N) [, (N,K)] bf16
            xt16, x16 = ops.transpose_cast(x, also_plain=ctx.needs_input_grad[1])   # (K,M) [, (M,K)]
            y = ops.gemm_tn(xt16, wt16, bias=b)                            # (M,N) = x W^T + b
            ctx.save_for_backward(x16, w16)
        else:
            y = ops.linear_fwd(x, w, b)
            ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = dw = None
        if ctx.tc:
            dyt16, dy16 = ops.transpose_cast(dy, also_plain=True)          # (N,M), (M,N) bf16
            if ctx.needs_input_grad[0]:
                dx = ops.gemm_tn(dyt16, w)                                 # (M,K) = dy W
            if ctx.needs_input_grad[1]:
                if ctx.exchange:
                    dy16, x = _gather_rows(dy16), _gather_rows(x)
                dw = ops.gemm_tn(dy16, x)                                  # (N,K) = dy^T x
        else:
            dx = ops.linear_dgrad(dy, w) if ctx.needs_input_grad[0] else None
            if ctx.needs_input_grad[1]:
                dw = ops.linear_wgrad(_gather_rows(dy), _gather_rows(x)) if ctx.exchange else ops.linear_wgrad(dy, x)
        if dw is not None and ctx.exchange:
            dw = dw / _world()                                             # DDP averages gradients over the ranks
        db = ops.colsum(dy) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db, None


def _gather_rows(t):
    """(M, C) on every rank -> (world*M, C), rank-major (no gradient)."""
    t = t.contiguous()
    out = torch.empty(_world() * t.shape[0], t.shape[1], dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t)
    return out


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


class BN1dFn(torch.autograd.Function):
    """(Sync)BatchNorm over the rows of a (M,C) matrix, optionally fused with the following ReLU
    (nonlinear_neck.py:95-97).  With an initialised process group of world size > 1 the (sum, sumsq) and the backward
    (sum dz, sum dz*xhat) pairs are all-reduced between the two kernels -> nn.SyncBatchNorm semantics."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, momentum, eps, relu, sync):
        x = x.contiguous().float()
        m, c = x.shape
        dev = x.device
        lib, st = ops.lib, ops._stream()
        if gamma is None:
            gamma = torch.ones(c, device=dev)
        if beta is None:
            beta = torch.zeros(c, device=dev)
        y = torch.empty_like(x)
        mean, rstd = torch.empty(c, device=dev), torch.empty(c, device=dev)
        stats = torch.empty(2, c, device=dev)
        count = float(m)
        world = _world() if (sync and training) else 1
        if training:
            lib.cmu_bn1d_stats(x.data_ptr(), m, c, stats.data_ptr(), st)
            if world > 1:
                dist.all_reduce(stats)
                count = float(m * world)
        lib.cmu_bn1d_apply(x.data_ptr(), stats.data_ptr(), count, m, c, gamma.data_ptr(), beta.data_ptr(),
                           ops._ptr(running_mean), ops._ptr(running_var), float(momentum), float(eps), int(training),
                           int(relu), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), st)
        ctx.save_for_backward(x, y, gamma, mean, rstd)
        ctx.relu, ctx.count, ctx.world, ctx.training = relu, count, world, training
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, mean, rstd = ctx.saved_tensors
        if not ctx.training:
            raise CmuError('backward through eval-mode BatchNorm1d is not implemented')
        dy = dy.contiguous().float()
        m, c = x.shape
        lib, st = ops.lib, ops._stream()
        sums = torch.empty(2, c, device=x.device)
        lib.cmu_bn1d_bwd_stats(dy.data_ptr(), y.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), m, c,
                               int(ctx.relu), sums.data_ptr(), st)
        local = sums.clone() if ctx.world > 1 else sums
        if ctx.world > 1:
            dist.all_reduce(sums)
        dx = torch.empty_like(x)
        lib.cmu_bn1d_bwd_apply(dy.data_ptr(), y.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                               gamma.data_ptr(), sums.data_ptr(), ctx.count, m, c, int(ctx.relu), dx.data_ptr(), st)
        # parameter gradients are LOCAL sums (the data-parallel gradient all-reduce averages them afterwards)
        dgamma = local[1] if ctx.needs_input_grad[1] else None
        dbeta = local[0] if ctx.needs_input_grad[2] else None
        return dx, dgamma, dbeta, None, None, None, None, None, None, None


class MaskedMSEFn(torch.autograd.Function):
    """loss_rc = rc_weight * sum(((pred - t)^2) * mask) / sum(mask) with t the row-normalised image
    (cmunet_head.py:62-70,89).  pred is any (B,H,W) fp32 view whose rows are contiguous (e.g. pred_pixel[:, 1])."""

    @staticmethod
    def forward(ctx, x, pred, mask, rc_weight):
        x = x.contiguous().float()
        mask = mask.contiguous()
        if pred.dtype != torch.float32 or pred.stride(2) != 1 or pred.stride(1) != pred.shape[2]:
            pred = pred.float().contiguous()
        b, h, w = x.shape
        acc = torch.empty(2, dtype=torch.float64, device=x.device)
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        ops.lib.cmu_masked_mse_fwd(x.data_ptr(), pred.data_ptr(), pred.stride(0), mask.data_ptr(), acc.data_ptr(),
                                   float(rc_weight), loss.data_ptr(), b, h, w, ops._stream())
        ctx.save_for_backward(x, pred, mask, acc)
        ctx.rc_weight = float(rc_weight)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        x, pred, mask, acc = ctx.saved_tensors
        b, h, w = x.shape
        gs = (g.float() * ctx.rc_weight).reshape(1).contiguous()
        dpred = torch.empty(b, h, w, dtype=torch.float32, device=x.device)
        ops.lib.cmu_masked_mse_bwd(x.data_ptr(), pred.data_ptr(), pred.stride(0), mask.data_ptr(), acc.data_ptr(),
                                   gs.data_ptr(), dpred.data_ptr(), h * w, b, h, w, ops._stream())
        return None, dpred, None, None


class InfoNCEFn(torch.autograd.Function):
    """loss_ct = ct_weight * 2 * tau * CE(normalize(q) @ Z^T / tau, arange(B) + B*rank) (cmunet_head.py:74-88).
    z_all: L2-normalised, all-gathered keys (no gradient, :77-79).  Logits stay on chip; dq is produced in the same
    kernel and scaled by the upstream gradient in backward."""

    @staticmethod
    def forward(ctx, q, z_all, label_offset, tau, ct_weight):
        q = q.contiguous().float()
        z_all = z_all.contiguous().float()
        loss, dq = ops.infonce(q, z_all, label_offset, tau, ct_weight, need_grad=ctx.needs_input_grad[0])
        if dq is not None:
            ctx.save_for_backward(dq)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dq,) = ctx.saved_tensors
        return dq * g, None, None, None, None
