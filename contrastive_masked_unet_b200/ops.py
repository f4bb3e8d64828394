"""Tensor-level wrappers over the C ABI (include/cmu_b200.h).  torch is used for device memory and streams only;
every computation below is a call into libcmu_b200.so.  Activations ("act") are NHWC bf16 tensors of shape
(N, H, W, C).  No fallbacks: a CPU tensor or a missing library raises."""
import ctypes

import torch

from ._lib import CmuError, lib

BF16 = torch.bfloat16


def _ptr(t):
    return 0 if t is None else t.data_ptr()


_CUDA_OK = None


def _stream():
    global _CUDA_OK
    if _CUDA_OK is None:
        _CUDA_OK = torch.cuda.is_available()
    if not _CUDA_OK:
        raise CmuError('contrastive_masked_unet_b200: no CUDA device -- this package runs on sm_100a only '
                       '(there is no CPU fallback)')
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise CmuError('contrastive_masked_unet_b200: this path runs on CUDA (sm_100a) only; got a CPU tensor '
                           '(there is no CPU fallback)')


def _act(t):
    assert t.dtype == BF16 and t.dim() == 4 and t.is_contiguous(), (t.dtype, t.shape, t.stride())
    return t


def device_check():
    lib.cmu_device_check()


class KernelTimer:
    """Optional CUDA-event timing of the tensor-core launches (bench.py roofline): set `ops.PROFILER = KernelTimer()`."""

    def __init__(self):
        self.records = []

    class _Region:
        def __init__(self, timer, name, flops):
            self.t, self.name, self.flops = timer, name, flops

        def __enter__(self):
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

        def __exit__(self, *exc):
            self.e1.record()
            self.t.records.append((self.name, self.flops, self.e0, self.e1))

    def region(self, name, flops):
        return KernelTimer._Region(self, name, flops)

    def summarize(self):
        torch.cuda.synchronize()
        out = {}
        for name, flops, e0, e1 in self.records:
            c, ms, fl = out.get(name, (0, 0.0, 0.0))
            out[name] = (c + 1, ms + e0.elapsed_time(e1), fl + flops)
        return out


class _NoRegion:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


PROFILER = None
_NOREGION = _NoRegion()


def _timed(name, flops):
    return PROFILER.region(name, flops) if PROFILER is not None else _NOREGION


# ------------------------------------------------------------------------------------------- weights
def pack_conv3x3(w, need_dgrad=True):
    """(Cout,Cin,3,3) fp32 -> (wf [9,Cout,Cin], wd [9,Cin,Cout]) bf16 GEMM operands."""
    _need_cuda(w)
    cout, cin = w.shape[0], w.shape[1]
    w = w.detach().contiguous().float()
    wf = torch.empty(9, cout, cin, dtype=BF16, device=w.device)
    wd = torch.empty(9, cin, cout, dtype=BF16, device=w.device) if need_dgrad else None
    lib.cmu_pack_conv3x3_weights(_ptr(w), cout, cin, _ptr(wf), _ptr(wd), _stream())
    return wf, wd


def pack_convT2x2(w, need_dgrad=True):
    """(Cin,Cout,2,2) fp32 -> (wf [4*Cout,Cin], wd [Cin,4*Cout]) bf16."""
    _need_cuda(w)
    cin, cout = w.shape[0], w.shape[1]
    w = w.detach().contiguous().float()
    wf = torch.empty(4 * cout, cin, dtype=BF16, device=w.device)
    wd = torch.empty(cin, 4 * cout, dtype=BF16, device=w.device) if need_dgrad else None
    lib.cmu_pack_convT2x2_weights(_ptr(w), cin, cout, _ptr(wf), _ptr(wd), _stream())
    return wf, wd


def cast_bf16(x):
    _need_cuda(x)
    x = x.detach().contiguous().float()
    y = torch.empty(x.shape, dtype=BF16, device=x.device)
    lib.cmu_cast_f32_to_bf16(_ptr(x), _ptr(y), x.numel(), _stream())
    return y


# ------------------------------------------------------------------------------------------- conv 3x3
class ConvStats:
    """Per-CTA (sum, sumsq) partials written by a conv epilogue: partial[grid][2][bn_tile]."""
    __slots__ = ('partial', 'grid', 'bn_tile', 'count')

    def __init__(self, partial, grid, bn_tile, count):
        self.partial, self.grid, self.bn_tile, self.count = partial, grid, bn_tile, count


def conv3x3_fprop(x0, x1, wf, want_stats=True):
    _need_cuda(x0, wf)
    _act(x0)
    n, h, w, c0 = x0.shape
    c1 = 0 if x1 is None else _act(x1).shape[3]
    cout = wf.shape[1]
    assert wf.shape[2] == c0 + c1
    y = torch.empty(n, h, w, cout, dtype=BF16, device=x0.device)
    stats = None
    g, b = ctypes.c_int(0), ctypes.c_int(0)
    partial = None
    if want_stats:
        partial = torch.empty(max(lib.cmu_conv_max_grid(), 64) * 2 * max(128, cout), dtype=torch.float32,
                              device=x0.device)
    with _timed('conv3x3_fprop', 2.0 * n * h * w * cout * 9 * (c0 + c1)):
        lib.cmu_conv3x3_fprop(_ptr(x0), c0, _ptr(x1), c1, n, h, w, _ptr(wf), cout, _ptr(y), _ptr(partial),
                              ctypes.byref(g), ctypes.byref(b), _stream())
    if want_stats:
        stats = ConvStats(partial, g.value, b.value, float(n * h * w))
    return y, stats


def conv3x3_dgrad(dy, wd, c0, c1=0, want_colsum0=False):
    """-> (dx0, dx1) [, colsum0]: with want_colsum0 the per-channel sum of dx0 over all pixels (fp32 (c0,), taken in the
    epilogue from the stored bf16 values) is returned as well -- the bias gradient of a ConvTranspose2d that produced x0."""
    _need_cuda(dy, wd)
    _act(dy)
    n, h, w, cout = dy.shape
    dx0 = torch.empty(n, h, w, c0, dtype=BF16, device=dy.device)
    dx1 = torch.empty(n, h, w, c1, dtype=BF16, device=dy.device) if c1 else None
    if not want_colsum0:
        with _timed('conv3x3_dgrad', 2.0 * n * h * w * cout * 9 * (c0 + c1)):
            lib.cmu_conv3x3_dgrad(_ptr(dy), cout, n, h, w, _ptr(wd), _ptr(dx0), c0, _ptr(dx1), c1, _stream())
        return dx0, dx1
    partial = torch.empty(max(lib.cmu_conv_max_grid(), 64) * 2 * max(128, c0 + c1), dtype=torch.float32, device=dy.device)
    g, b = ctypes.c_int(0), ctypes.c_int(0)
    with _timed('conv3x3_dgrad', 2.0 * n * h * w * cout * 9 * (c0 + c1)):
        lib.cmu_conv3x3_dgrad_sums(_ptr(dy), cout, n, h, w, _ptr(wd), _ptr(dx0), c0, _ptr(dx1), c1, _ptr(partial),
                                   ctypes.byref(g), ctypes.byref(b), _stream())
    colsum0 = torch.empty(c0, dtype=torch.float32, device=dy.device)
    lib.cmu_stats_colsum(_ptr(partial), g.value, b.value, c0 + c1, c0, _ptr(colsum0), _stream())
    return dx0, dx1, colsum0


def conv3x3_wgrad(x0, x1, dy, dw=None, accumulate=False):
    _need_cuda(x0, dy)
    n, h, w, c0 = _act(x0).shape
    c1 = 0 if x1 is None else _act(x1).shape[3]
    cout = _act(dy).shape[3]
    if dw is None:
        dw = torch.empty(cout, c0 + c1, 3, 3, dtype=torch.float32, device=dy.device)
        accumulate = False
    nbytes = lib.cmu_conv3x3_wgrad_workspace_bytes(c0 + c1, cout, n, h, w)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=dy.device)
    with _timed('conv3x3_wgrad', 2.0 * n * h * w * cout * 9 * (c0 + c1)):
        lib.cmu_conv3x3_wgrad(_ptr(x0), c0, _ptr(x1), c1, _ptr(dy), cout, n, h, w, _ptr(ws), nbytes, _ptr(dw),
                              int(accumulate), _stream())
    return dw


def conv3x3_c1_fprop(x, mask0, w, want_stats=True):
    """x (N,H,W) fp32, mask0 (H,W) uint8 or None, w (64,1,3,3) fp32 -> act (N,H,W,64)."""
    _need_cuda(x, w)
    x = x.contiguous().float()
    n, h, wd = x.shape
    cout = w.shape[0]
    y = torch.empty(n, h, wd, cout, dtype=BF16, device=x.device)
    grid = lib.cmu_conv3x3_c1_grid()
    partial = torch.empty(grid * 2 * cout, dtype=torch.float32, device=x.device) if want_stats else None
    wc = w.detach().contiguous().float()
    lib.cmu_conv3x3_c1_fprop(_ptr(x), _ptr(mask0), _ptr(wc), cout, _ptr(y), _ptr(partial), n, h, wd, _stream())
    return y, (ConvStats(partial, grid, cout, float(n * h * wd)) if want_stats else None)


def conv3x3_c1_wgrad(x, mask0, dy):
    _need_cuda(x, dy)
    x = x.contiguous().float()
    n, h, wd = x.shape
    cout = _act(dy).shape[3]
    grid = lib.cmu_conv3x3_c1_grid()
    partial = torch.empty(grid * cout * 9, dtype=torch.float32, device=x.device)
    dw = torch.empty(cout, 1, 3, 3, dtype=torch.float32, device=x.device)
    lib.cmu_conv3x3_c1_wgrad(_ptr(x), _ptr(mask0), _ptr(dy), cout, _ptr(partial), _ptr(dw), 0, n, h, wd, _stream())
    return dw


# ------------------------------------------------------------------------------------------- BatchNorm2d
def bn_finalize(stats, gamma, beta, conv_bias, running_mean, running_var, momentum, eps, training):
    """-> (scale, shift, mean, rstd), each (C,) fp32; updates the running statistics in place when training."""
    c = gamma.numel()
    dev = gamma.device
    _need_cuda(gamma)
    out = torch.empty(4, c, dtype=torch.float32, device=dev)
    if training:
        lib.cmu_bn_finalize(_ptr(stats.partial), stats.grid, stats.bn_tile, c, float(stats.count), _ptr(gamma),
                            _ptr(beta), _ptr(conv_bias), _ptr(running_mean), _ptr(running_var), float(momentum),
                            float(eps), 1, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), _stream())
    else:
        lib.cmu_bn_finalize(0, 0, 0, c, 1.0, _ptr(gamma), _ptr(beta), _ptr(conv_bias), _ptr(running_mean),
                            _ptr(running_var), float(momentum), float(eps), 0, _ptr(out[0]), _ptr(out[1]),
                            _ptr(out[2]), _ptr(out[3]), _stream())
    return out[0], out[1], out[2], out[3]


def bn_relu_apply(y, scale, shift, pool=False, want_act=True):
    """want_act=False (with pool=True): only the pooled map is written (no-grad paths whose skip tensor nobody reads)."""
    n, h, w, c = _act(y).shape
    a = torch.empty_like(y) if (want_act or not pool) else None
    pooled = torch.empty(n, h // 2, w // 2, c, dtype=BF16, device=y.device) if pool else None
    lib.cmu_bn_relu_apply(_ptr(y), _ptr(scale), _ptr(shift), _ptr(a), _ptr(pooled), n, h, w, c, _stream())
    return a, pooled


def bn_relu_bwd(da, dpool, y, scale, shift, mean, rstd, training=True):
    """-> (dy act, dgamma (C,), dbeta (C,))."""
    n, h, w, c = _act(y).shape
    grid = lib.cmu_bn_bwd_grid()
    partial = torch.empty(grid * 2 * c, dtype=torch.float32, device=y.device)
    sums = torch.empty(2, c, dtype=torch.float32, device=y.device)
    dy = torch.empty_like(y)
    lib.cmu_bn_relu_bwd(_ptr(da), _ptr(dpool), _ptr(y), _ptr(scale), _ptr(shift), _ptr(mean), _ptr(rstd),
                        _ptr(partial), _ptr(sums), _ptr(dy), n, h, w, c, int(training), _stream())
    return dy, sums[1], sums[0]


# ------------------------------------------------------------------------------------------- ConvTranspose 2x2 s2
def convT2x2_fprop(x, wf, bias):
    n, h, w, cin = _act(x).shape
    cout = wf.shape[0] // 4
    y = torch.empty(n, 2 * h, 2 * w, cout, dtype=BF16, device=x.device)
    with _timed('convT_fprop', 2.0 * n * h * w * cin * 4 * cout):
        lib.cmu_convT2x2_fprop(_ptr(x), cin, n, h, w, _ptr(wf), cout, _ptr(bias), _ptr(y), _stream())
    return y


def convT2x2_dgrad(dy, wd):
    n, h2, w2, cout = _act(dy).shape
    cin = wd.shape[0]
    h, w = h2 // 2, w2 // 2
    dx = torch.empty(n, h, w, cin, dtype=BF16, device=dy.device)
    with _timed('convT_dgrad', 2.0 * n * h * w * cin * 4 * cout):
        lib.cmu_convT2x2_dgrad(_ptr(dy), cout, n, h, w, _ptr(wd), cin, _ptr(dx), _stream())
    return dx


def convT2x2_wgrad(x, dy):
    n, h, w, cin = _act(x).shape
    cout = _act(dy).shape[3]
    dw = torch.empty(cin, cout, 2, 2, dtype=torch.float32, device=x.device)
    nbytes = lib.cmu_convT2x2_wgrad_workspace_bytes(cin, cout, n, h, w)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=x.device)
    with _timed('convT_wgrad', 2.0 * n * h * w * cin * 4 * cout):
        lib.cmu_convT2x2_wgrad(_ptr(x), cin, _ptr(dy), cout, n, h, w, _ptr(ws), nbytes, _ptr(dw), 0, _stream())
    return dw


def colsum_bf16(x2d_rows, c, x):
    grid = lib.cmu_bn_bwd_grid()
    partial = torch.empty(grid * c, dtype=torch.float32, device=x.device)
    out = torch.empty(c, dtype=torch.float32, device=x.device)
    lib.cmu_colsum_bf16(_ptr(x), x2d_rows, c, _ptr(partial), _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------- 1x1
def conv1x1_fprop(x, w_bf16, bias, name='conv1x1_fprop'):
    n, h, w, cin = _act(x).shape
    cout = w_bf16.shape[0]
    y = torch.empty(n, h, w, cout, dtype=BF16, device=x.device)
    with _timed(name, 2.0 * n * h * w * cin * cout):
        lib.cmu_conv1x1_fprop(_ptr(x), cin, n, h, w, _ptr(w_bf16), cout, _ptr(bias), _ptr(y), _stream())
    return y


def conv1x1_fprop_exp(x, w_bf16, inv_temperature, name='conv1x1_fprop_exp'):
    """y[p][c] = exp((x[p] . w[c] - 1) * inv_temperature) (bf16) and per-CTA partial column sums of y.
    -> (y, partial, grid, bn_tile)"""
    n, h, w, cin = _act(x).shape
    cout = w_bf16.shape[0]
    y = torch.empty(n, h, w, cout, dtype=BF16, device=x.device)
    partial = torch.empty(max(lib.cmu_conv_max_grid(), 64) * 2 * max(128, cout), dtype=torch.float32, device=x.device)
    g, b = ctypes.c_int(0), ctypes.c_int(0)
    with _timed(name, 2.0 * n * h * w * cin * cout):
        lib.cmu_conv1x1_fprop_exp(_ptr(x), cin, n, h, w, _ptr(w_bf16), cout, float(inv_temperature), _ptr(y), _ptr(partial),
                                  ctypes.byref(g), ctypes.byref(b), _stream())
    return y, partial, g.value, b.value


def head1x1_fprop(a, w, b):
    """a act (N,H,W,64); w (2,64[,1,1]) fp32; -> (N,2,H,W) fp32 NCHW."""
    n, h, wd, cin = _act(a).shape
    w2 = w.detach().reshape(w.shape[0], -1).contiguous().float()
    out = torch.empty(n, w2.shape[0], h, wd, dtype=torch.float32, device=a.device)
    lib.cmu_head1x1_fprop(_ptr(a), _ptr(w2), _ptr(b.detach().contiguous().float()), _ptr(out), n, h, wd, cin,
                          w2.shape[0], _stream())
    return out


def head1x1_bwd(a, w, dout):
    n, h, wd, cin = _act(a).shape
    w2 = w.detach().reshape(w.shape[0], -1).contiguous().float()
    dout = dout.contiguous().float()
    da = torch.empty_like(a)
    acc = torch.empty(130, dtype=torch.float32, device=a.device)
    lib.cmu_head1x1_bwd(_ptr(a), _ptr(w2), _ptr(dout), _ptr(da), _ptr(acc), n, h, wd, cin, w2.shape[0], _stream())
    return da, acc[:128].view(2, 64), acc[128:130]


def bn_relu_head_fwd(y, scale, shift, w, b):
    """conv_last(relu(y * scale + shift)) for a 64-channel raw conv output y (act) -> (N,2,H,W) fp32; the activated
    tensor is never written."""
    n, h, wd, cin = _act(y).shape
    w2 = w.detach().reshape(w.shape[0], -1).contiguous().float()
    out = torch.empty(n, w2.shape[0], h, wd, dtype=torch.float32, device=y.device)
    lib.cmu_bn_relu_head_fwd(_ptr(y), _ptr(scale), _ptr(shift), _ptr(w2), _ptr(b.detach().contiguous().float()), _ptr(out),
                             n, h, wd, cin, w2.shape[0], _stream())
    return out


def bn_relu_head_bwd(y, scale, shift, mean, rstd, w, dout, training=True):
    """-> (dy act, dgamma (64,), dbeta (64,), d conv_last.weight (2,64), d conv_last.bias (2,))."""
    n, h, wd, cin = _act(y).shape
    w2 = w.detach().reshape(w.shape[0], -1).contiguous().float()
    dout = dout.contiguous().float()
    partial = torch.empty(lib.cmu_bn_relu_head_grid() * 258, dtype=torch.float32, device=y.device)
    sums = torch.empty(2, cin, dtype=torch.float32, device=y.device)
    acc = torch.empty(130, dtype=torch.float32, device=y.device)
    dy = torch.empty_like(y)
    lib.cmu_bn_relu_head_bwd(_ptr(y), _ptr(scale), _ptr(shift), _ptr(mean), _ptr(rstd), _ptr(w2), _ptr(dout), _ptr(partial),
                             _ptr(sums), _ptr(acc), _ptr(dy), n, h, wd, cin, w2.shape[0], int(training), _stream())
    return dy, sums[1], sums[0], acc[:128].view(2, 64), acc[128:130]


# ------------------------------------------------------------------------------------------- linear / BN1d
def sgemm(a, sam, sak, b, sbn, sbk, m, n, k, bias=None, out=None, accumulate=False):
    dev = a.device
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=dev)
        accumulate = False
    nbytes = lib.cmu_sgemm_workspace_bytes(m, n, k)
    ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=dev)
    lib.cmu_sgemm(_ptr(a), sam, sak, _ptr(b), sbn, sbk, _ptr(out), n, _ptr(bias), m, n, k, int(accumulate), _ptr(ws),
                  nbytes, _stream())
    return out


def linear_fwd(x, w, bias):
    """x (M,K) fp32, w (N,K) fp32 -> (M,N)."""
    m, k = x.shape
    n = w.shape[0]
    return sgemm(x, k, 1, w, k, 1, m, n, k, bias=bias)


def linear_dgrad(dy, w):
    """dy (M,N), w (N,K) -> dx (M,K)."""
    m, n = dy.shape
    k = w.shape[1]
    return sgemm(dy, n, 1, w, 1, k, m, k, n)


def linear_wgrad(dy, x):
    """dy (M,N), x (M,K) -> dw (N,K)."""
    m, n = dy.shape
    k = x.shape[1]
    return sgemm(dy, 1, n, x, 1, k, n, k, m)


def transpose_cast(x, also_plain=False):
    """fp32 (R,C) -> bf16 (C,R) [and optionally the untransposed bf16 (R,C)] in one pass over x."""
    x = x.detach().contiguous().float()
    r, c = x.shape
    yt = torch.empty(c, r, dtype=BF16, device=x.device)
    y = torch.empty(r, c, dtype=BF16, device=x.device) if also_plain else None
    lib.cmu_transpose_cast_bf16(_ptr(x), _ptr(yt), _ptr(y), r, c, _stream())
    return yt, y


def gemm_tn(q, p, bias=None, out=None, accumulate=False):
    """out (QC,PC) fp32 = q^T p (+ bias[PC]); q (R,QC) bf16, p (R,PC) bf16 -- tcgen05 (K2 engine, rows = reduction)."""
    assert q.dtype == BF16 and p.dtype == BF16 and q.is_contiguous() and p.is_contiguous() and q.shape[0] == p.shape[0]
    rows, qc = q.shape
    pc = p.shape[1]
    if out is None:
        out = torch.empty(qc, pc, dtype=torch.float32, device=q.device)
        accumulate = False
    nbytes = lib.cmu_gemm_tn_workspace_bytes(qc, pc, rows)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=q.device) if nbytes else None
    with _timed('gemm_tn', 2.0 * rows * qc * pc):
        lib.cmu_gemm_tn_bf16(_ptr(q), qc, _ptr(p), pc, rows, _ptr(out), _ptr(bias), int(accumulate), _ptr(ws), nbytes,
                             _stream())
    return out


def tc_linear_ok(m, k, n):
    """Shapes for which the projector linears run on the tensor-core engine (else the SIMT SGEMM)."""
    return m % 64 == 0 and k % 64 == 0 and n % 64 == 0 and m <= 128 and k * n >= (1 << 22)


def colsum(x):
    m, n = x.shape
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    lib.cmu_colsum(_ptr(x), m, n, _ptr(out), 0, _stream())
    return out


# ------------------------------------------------------------------------------------------- losses
def masked_mse_fwd(x, pred_full, mask, rc_weight):
    """x (B,H,W) fp32; pred_full (B,2,H,W) fp32 (channel 1 is used, cmunet.py:133); mask (B,H,W) uint8."""
    b, h, w = x.shape
    acc = torch.empty(2, dtype=torch.float64, device=x.device)
    loss = torch.empty(1, dtype=torch.float32, device=x.device)
    lib.cmu_masked_mse_fwd(_ptr(x), pred_full.data_ptr() + h * w * 4, 2 * h * w, _ptr(mask), _ptr(acc),
                           float(rc_weight), _ptr(loss), b, h, w, _stream())
    return loss, acc


def masked_mse_bwd(x, pred_full, mask, acc, gscale):
    b, h, w = x.shape
    dpred = torch.zeros_like(pred_full)
    lib.cmu_masked_mse_bwd(_ptr(x), pred_full.data_ptr() + h * w * 4, 2 * h * w, _ptr(mask), _ptr(acc), _ptr(gscale),
                           dpred.data_ptr() + h * w * 4, 2 * h * w, b, h, w, _stream())
    return dpred


def l2_normalize_rows(x):
    y = torch.empty_like(x)
    lib.cmu_l2_normalize_rows(_ptr(x), _ptr(y), x.shape[0], x.shape[1], _stream())
    return y


def infonce(q, z, label_offset, tau, ct_weight, need_grad=True):
    bsz, dim = q.shape
    rows = torch.empty(bsz, dtype=torch.float32, device=q.device)
    loss = torch.empty(1, dtype=torch.float32, device=q.device)
    dq = torch.empty_like(q) if need_grad else None
    lib.cmu_infonce_fwd_bwd(_ptr(q), _ptr(z), bsz, z.shape[0], dim, int(label_offset), float(tau), float(ct_weight),
                            _ptr(rows), _ptr(loss), _ptr(dq), _stream())
    return loss, dq


def seg_losses(logits, gt, gscale=None, dice_eps=1e-5, beta=1.0, iou_eps=1e-7):
    n, c, h, w = logits.shape
    assert c == 2 and gt.dtype == torch.float64
    acc = torch.empty(4, dtype=torch.float64, device=logits.device)
    out = torch.empty(3, dtype=torch.float64, device=logits.device)
    dl = torch.empty_like(logits) if gscale is not None else None
    lib.cmu_seg_losses(_ptr(logits), _ptr(gt), _ptr(acc), _ptr(out), _ptr(dl), _ptr(gscale), n, h, w, float(dice_eps),
                       float(beta), float(iou_eps), _stream())
    return out, dl
