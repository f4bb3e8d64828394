"""Kernel-level parity checks (GPU): every C-ABI kernel against a plain fp32 torch statement of the same op on the
same bf16-rounded inputs, or against the oracle (oracle/) where the op is reference-specific.  Each check returns a
dict of error measures and raises AssertionError on failure.  Used by tests/test_kernels_gpu.py (pytest -m gpu) and
by tools/run_checks.py (one subprocess per check, so a faulting kernel cannot take the other checks down)."""
import numpy as np
import torch
import torch.nn.functional as F

from contrastive_masked_unet_b200 import ops
from contrastive_masked_unet_b200._lib import lib

DEV = 'cuda'
BF16 = torch.bfloat16
# fp32 references must really be fp32 (cuDNN / cuBLAS would otherwise use TF32)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def nhwc(x):
    """(N,C,H,W) fp32 -> act (N,H,W,C) bf16 contiguous."""
    return x.permute(0, 2, 3, 1).contiguous().to(BF16)


def nchw(a):
    return a.float().permute(0, 3, 1, 2).contiguous()


def rel_err(got, ref):
    got, ref = got.double(), ref.double()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def cos(got, ref):
    got, ref = got.double().flatten(), ref.double().flatten()
    return float((got @ ref) / (got.norm() * ref.norm()).clamp_min(1e-30))


def _gen(seed):
    return torch.Generator(device='cpu').manual_seed(seed)


def _randn(shape, g, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).to(DEV)


def reduce_stats(st, c):
    """ConvStats partial [grid][2][bn_tile] -> (sum (C,), sumsq (C,)) in float64."""
    p = st.partial[: st.grid * 2 * st.bn_tile].view(st.grid, 2, st.bn_tile).double()
    n_tiles = c // st.bn_tile
    s1 = torch.zeros(c, dtype=torch.float64, device=p.device)
    s2 = torch.zeros(c, dtype=torch.float64, device=p.device)
    for nt in range(n_tiles):
        sel = p[nt::n_tiles]
        s1[nt * st.bn_tile:(nt + 1) * st.bn_tile] = sel[:, 0].sum(0)
        s2[nt * st.bn_tile:(nt + 1) * st.bn_tile] = sel[:, 1].sum(0)
    return s1, s2


# ------------------------------------------------------------------------------------------------ conv3x3
def check_conv3x3(n=2, h=24, w=40, c0=64, c1=0, cout=64, seed=0, tol=1.5e-2):
    g = _gen(seed)
    x0 = _randn((n, c0, h, w), g)
    x1 = _randn((n, c1, h, w), g) if c1 else None
    wt = _randn((cout, c0 + c1, 3, 3), g, (2.0 / (9 * (c0 + c1))) ** 0.5)
    dyr = _randn((n, cout, h, w), g)
    a0, a1 = nhwc(x0), (nhwc(x1) if c1 else None)
    dy = nhwc(dyr)
    wf, wd = ops.pack_conv3x3(wt)
    # reference on the bf16-rounded operands
    xr = torch.cat([nchw(a0)] + ([nchw(a1)] if c1 else []), 1).requires_grad_(True)
    wr = wt.to(BF16).float().requires_grad_(True)
    yr = F.conv2d(xr, wr, padding=1)
    yr.backward(nchw(dy))
    y, st = ops.conv3x3_fprop(a0, a1, wf, want_stats=True)
    dx0, dx1 = ops.conv3x3_dgrad(dy, wd, c0, c1)
    # same dgrad with the column sums of dx0 taken in the epilogue (ConvTranspose bias gradient hand-off)
    dx0b, dx1b, cs0 = ops.conv3x3_dgrad(dy, wd, c0, c1, want_colsum0=True)
    dw = ops.conv3x3_wgrad(a0, a1, dy)
    torch.cuda.synchronize()
    res = {'fprop': rel_err(nchw(y), yr.detach())}
    res['dgrad_colsum0'] = rel_err(cs0, nchw(dx0).double().sum((0, 2, 3)))
    assert torch.equal(dx0, dx0b) and (dx1 is None or torch.equal(dx1, dx1b)) and res['dgrad_colsum0'] < 1e-4, res
    s1, s2 = reduce_stats(st, cout)
    # the epilogue takes the statistics over the bf16 outputs it stores (the values BatchNorm then normalises)
    yk = nchw(y).double()
    res['stats_sum'] = rel_err(s1, yk.sum((0, 2, 3)))
    res['stats_sq'] = rel_err(s2, (yk ** 2).sum((0, 2, 3)))
    res['stats_sq_vs_fp32'] = rel_err(s2, (yr.detach().double() ** 2).sum((0, 2, 3)))
    dxg = torch.cat([nchw(dx0)] + ([nchw(dx1)] if c1 else []), 1)
    res['dgrad'] = rel_err(dxg, xr.grad)
    res['wgrad'] = rel_err(dw, wr.grad)
    res['wgrad_cos'] = cos(dw, wr.grad)
    assert res['fprop'] < tol and res['dgrad'] < tol, res
    assert res['wgrad'] < tol and res['wgrad_cos'] > 0.9999, res
    assert res['stats_sum'] < 1e-4 and res['stats_sq'] < 1e-4 and res['stats_sq_vs_fp32'] < 5e-3, res
    return res


def check_conv3x3_fprop_only(n=2, h=24, w=40, c0=64, c1=0, cout=64, seed=0, tol=1.5e-2):
    g = _gen(seed)
    x0 = _randn((n, c0, h, w), g)
    x1 = _randn((n, c1, h, w), g) if c1 else None
    wt = _randn((cout, c0 + c1, 3, 3), g, (2.0 / (9 * (c0 + c1))) ** 0.5)
    a0, a1 = nhwc(x0), (nhwc(x1) if c1 else None)
    wf, _ = ops.pack_conv3x3(wt)
    xr = torch.cat([nchw(a0)] + ([nchw(a1)] if c1 else []), 1)
    yr = F.conv2d(xr, wt.to(BF16).float(), padding=1)
    y, st = ops.conv3x3_fprop(a0, a1, wf, want_stats=True)
    torch.cuda.synchronize()
    res = {'fprop': rel_err(nchw(y), yr)}
    s1, s2 = reduce_stats(st, cout)
    yk = nchw(y).double()
    res['stats_sum'] = rel_err(s1, yk.sum((0, 2, 3)))
    res['stats_sq'] = rel_err(s2, (yk ** 2).sum((0, 2, 3)))
    res['stats_sq_vs_fp32'] = rel_err(s2, (yr.double() ** 2).sum((0, 2, 3)))
    assert res['fprop'] < tol, res
    assert res['stats_sum'] < 1e-4 and res['stats_sq'] < 1e-4 and res['stats_sq_vs_fp32'] < 5e-3, res
    return res


def check_conv3x3_wgrad_only(n=2, h=24, w=40, c0=64, c1=0, cout=64, seed=0, tol=1.5e-2):
    g = _gen(seed)
    x0 = _randn((n, c0, h, w), g)
    x1 = _randn((n, c1, h, w), g) if c1 else None
    dyr = _randn((n, cout, h, w), g)
    a0, a1 = nhwc(x0), (nhwc(x1) if c1 else None)
    dy = nhwc(dyr)
    xr = torch.cat([nchw(a0)] + ([nchw(a1)] if c1 else []), 1)
    wr = torch.zeros(cout, c0 + c1, 3, 3, device=DEV, requires_grad=True)
    F.conv2d(xr, wr, padding=1).backward(nchw(dy))
    dw = ops.conv3x3_wgrad(a0, a1, dy)
    torch.cuda.synchronize()
    res = {'wgrad': rel_err(dw, wr.grad), 'wgrad_cos': cos(dw, wr.grad)}
    assert res['wgrad'] < tol and res['wgrad_cos'] > 0.9999, res
    return res


def check_conv3x3_c1(n=3, h=32, w=48, seed=1):
    g = _gen(seed)
    x = _randn((n, h, w), g)
    mask0 = (torch.rand(h, w, generator=g) > 0.5).to(torch.uint8).to(DEV)
    wt = _randn((64, 1, 3, 3), g, 0.3)
    dy = nhwc(_randn((n, 64, h, w), g))
    xm = (x * (1 - mask0.float())).unsqueeze(1)
    wr = wt.clone().requires_grad_(True)
    yr = F.conv2d(xm, wr, padding=1)
    yr.backward(nchw(dy))
    y, st = ops.conv3x3_c1_fprop(x, mask0, wt)
    dw = ops.conv3x3_c1_wgrad(x, mask0, dy)
    y2, _ = ops.conv3x3_c1_fprop(x, None, wt, want_stats=False)
    torch.cuda.synchronize()
    s1, s2 = reduce_stats(st, 64)
    res = {'fprop': rel_err(nchw(y), yr.detach()), 'wgrad': rel_err(dw, wr.grad),
           'stats_sum': rel_err(s1, yr.detach().double().sum((0, 2, 3))),
           'stats_sq': rel_err(s2, (yr.detach().double() ** 2).sum((0, 2, 3))),
           'nomask': rel_err(nchw(y2), F.conv2d(x.unsqueeze(1), wt, padding=1))}
    assert res['fprop'] < 1e-2 and res['wgrad'] < 2e-3 and res['stats_sum'] < 1e-3 and res['stats_sq'] < 1e-3, res
    assert res['nomask'] < 1e-2, res
    return res


# ------------------------------------------------------------------------------------------------ convT / 1x1
def check_convT(n=2, h=12, w=20, cin=128, cout=64, seed=2, tol=1.5e-2):
    g = _gen(seed)
    x = nhwc(_randn((n, cin, h, w), g))
    wt = _randn((cin, cout, 2, 2), g, (1.0 / cin) ** 0.5)
    bias = _randn((cout,), g, 0.1)
    dy = nhwc(_randn((n, cout, 2 * h, 2 * w), g))
    wf, wd = ops.pack_convT2x2(wt)
    xr = nchw(x).requires_grad_(True)
    wr = wt.to(BF16).float().requires_grad_(True)
    yr = F.conv_transpose2d(xr, wr, bias, stride=2)
    yr.backward(nchw(dy))
    y = ops.convT2x2_fprop(x, wf, bias)
    dx = ops.convT2x2_dgrad(dy, wd)
    dw = ops.convT2x2_wgrad(x, dy)
    db = ops.colsum_bf16(n * 4 * h * w, cout, dy)
    torch.cuda.synchronize()
    res = {'fprop': rel_err(nchw(y), yr.detach()), 'dgrad': rel_err(nchw(dx), xr.grad), 'wgrad': rel_err(dw, wr.grad),
           'wgrad_cos': cos(dw, wr.grad), 'dbias': rel_err(db, nchw(dy).sum((0, 2, 3)))}
    assert res['fprop'] < tol and res['dgrad'] < tol and res['wgrad'] < tol and res['dbias'] < 1e-3, res
    return res


def check_conv1x1(n=2, h=8, w=8, cin=1024, cout=256, seed=3):
    g = _gen(seed)
    x = nhwc(_randn((n, cin, h, w), g))
    wt = _randn((cout, cin, 1, 1), g, (1.0 / cin) ** 0.5)
    bias = _randn((cout,), g, 0.1)
    wb = ops.cast_bf16(wt.view(cout, cin))
    y = ops.conv1x1_fprop(x, wb, bias)
    yr = F.conv2d(nchw(x), wt.to(BF16).float(), bias)
    t = torch.empty(n, cout, h * w, dtype=BF16, device=DEV)
    lib.cmu_nhwc_to_nchw_bf16(y.data_ptr(), t.data_ptr(), n, h * w, cout, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    res = {'fprop': rel_err(nchw(y), yr), 'transpose': float((t.float().view(n, cout, h, w) - nchw(y)).abs().max())}
    assert res['fprop'] < 1.5e-2 and res['transpose'] == 0.0, res
    return res


def check_head1x1(n=2, h=20, w=28, seed=4):
    g = _gen(seed)
    a = nhwc(_randn((n, 64, h, w), g).abs())
    wt = _randn((2, 64, 1, 1), g, 0.2)
    b = _randn((2,), g, 0.1)
    dout = _randn((n, 2, h, w), g)
    ar = nchw(a).requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    yr = F.conv2d(ar, wr, br)
    yr.backward(dout)
    y = ops.head1x1_fprop(a, wt, b)
    da, dw, db = ops.head1x1_bwd(a, wt, dout)
    torch.cuda.synchronize()
    res = {'fwd': rel_err(y, yr.detach()), 'da': rel_err(nchw(da), ar.grad), 'dw': rel_err(dw, wr.grad.view(2, 64)),
           'db': rel_err(db, br.grad)}
    assert res['fwd'] < 1e-4 and res['da'] < 1e-2 and res['dw'] < 1e-4 and res['db'] < 1e-4, res
    return res


# ------------------------------------------------------------------------------------------------ BN / ReLU / pool
def check_bn(n=3, h=16, w=24, c=64, pool=True, seed=5):
    g = _gen(seed)
    yv = _randn((n, c, h, w), g, 2.0) + 0.5
    gamma = (torch.rand(c, generator=g) + 0.5).to(DEV)
    beta = _randn((c,), g, 0.3)
    cbias = _randn((c,), g, 0.2)
    y = nhwc(yv)
    yf = nchw(y)
    # statistics as a conv epilogue would deliver them (one partial row)
    st = ops.ConvStats(torch.cat([yf.sum((0, 2, 3)), (yf ** 2).sum((0, 2, 3))]).float().contiguous(), 1, c, float(n * h * w))
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    scale, shift, mean, rstd = ops.bn_finalize(st, gamma, beta, cbias, rm, rv, 0.1, 1e-5, True)
    a, pooled = ops.bn_relu_apply(y, scale, shift, pool)
    # reference (conv bias re-added: it must cancel in train mode and show up in running_mean)
    yin = (yf + cbias.view(1, -1, 1, 1)).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_r, rv_r = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    ar = F.relu(F.batch_norm(yin, rm_r, rv_r, gr, br, True, 0.1, 1e-5))
    da = nhwc(_randn((n, c, h, w), g))
    res = {}
    if pool:
        # pool the bf16-rounded activations so that ties / argmax match what the kernel sees
        ar_q = ar + (nchw(a) - ar).detach()
        pr = F.max_pool2d(ar_q, 2)
        dp = nhwc(_randn((n, c, h // 2, w // 2), g))
        (ar_q * nchw(da)).sum().backward(retain_graph=True)
        (pr * nchw(dp)).sum().backward()
        res['pool'] = rel_err(nchw(pooled), pr.detach())
    else:
        dp = None
        (ar * nchw(da)).sum().backward()
    dy, dgamma, dbeta = ops.bn_relu_bwd(da, dp, y, scale, shift, mean, rstd)
    torch.cuda.synchronize()
    res.update({'act': rel_err(nchw(a), ar.detach()), 'rm': rel_err(rm, rm_r), 'rv': rel_err(rv, rv_r),
                'dy': rel_err(nchw(dy), yin.grad), 'dgamma': rel_err(dgamma, gr.grad), 'dbeta': rel_err(dbeta, br.grad)})
    assert res['act'] < 1e-2 and res['rm'] < 1e-4 and res['rv'] < 1e-4, res
    assert res['dy'] < 2e-2 and res['dgamma'] < 2e-3 and res['dbeta'] < 2e-3, res
    if pool:
        assert res['pool'] < 1e-6, res
    # eval mode: running statistics, bias folded into the shift
    sc2, sh2, _, _ = ops.bn_finalize(None, gamma, beta, cbias, rm, rv, 0.1, 1e-5, False)
    a2, _ = ops.bn_relu_apply(y, sc2, sh2, False)
    ar2 = F.relu(F.batch_norm(yf + cbias.view(1, -1, 1, 1), rm_r, rv_r, gamma, beta, False, 0.1, 1e-5))
    res['eval'] = rel_err(nchw(a2), ar2)
    assert res['eval'] < 1e-2, res
    return res


def check_bn_relu_head(n=3, h=20, w=28, seed=21, training=True):
    """fused decoder tail (BN apply + ReLU folded into conv_last, forward and backward) vs the unfused kernels of this
    library (bit-compatible up to summation order) and vs an fp32 torch statement."""
    c = 64
    g = _gen(seed)
    yv = _randn((n, c, h, w), g, 1.5) + 0.3
    gamma = (torch.rand(c, generator=g) + 0.5).to(DEV)
    beta = _randn((c,), g, 0.3)
    hw_ = _randn((2, c, 1, 1), g, 0.2)
    hb = _randn((2,), g, 0.1)
    dout = _randn((n, 2, h, w), g)
    y = nhwc(yv)
    yf = nchw(y)
    st = ops.ConvStats(torch.cat([yf.sum((0, 2, 3)), (yf ** 2).sum((0, 2, 3))]).float().contiguous(), 1, c, float(n * h * w))
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    scale, shift, mean, rstd = ops.bn_finalize(st, gamma, beta, None, rm, rv, 0.1, 1e-5, True)
    if not training:
        scale, shift, mean, rstd = ops.bn_finalize(None, gamma, beta, None, rm, rv, 0.1, 1e-5, False)
    # fused
    out = ops.bn_relu_head_fwd(y, scale, shift, hw_, hb)
    dy, dgamma, dbeta, dhw, dhb = ops.bn_relu_head_bwd(y, scale, shift, mean, rstd, hw_, dout, training)
    # unfused kernels of the library
    a, _ = ops.bn_relu_apply(y, scale, shift, False)
    out_u = ops.head1x1_fprop(a, hw_, hb)
    da_u, dhw_u, dhb_u = ops.head1x1_bwd(a, hw_, dout)
    dy_u, dgamma_u, dbeta_u = ops.bn_relu_bwd(da_u, None, y, scale, shift, mean, rstd, training)
    # fp32 torch statement on the same bf16 y
    yin = yf.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    hwr, hbr = hw_.clone().requires_grad_(True), hb.clone().requires_grad_(True)
    if training:
        ar = F.relu(F.batch_norm(yin, None, None, gr, br, True, 0.1, 1e-5))
    else:
        ar = F.relu(F.batch_norm(yin, rm.clone(), rv.clone(), gr, br, False, 0.1, 1e-5))
    pr = F.conv2d(ar, hwr, hbr)
    (pr * dout).sum().backward()
    torch.cuda.synchronize()
    res = {'fwd_vs_unfused': rel_err(out, out_u), 'dy_vs_unfused': rel_err(nchw(dy), nchw(dy_u)),
           'dgamma_vs_unfused': rel_err(dgamma, dgamma_u), 'dbeta_vs_unfused': rel_err(dbeta, dbeta_u),
           'dhw_vs_unfused': rel_err(dhw, dhw_u), 'dhb_vs_unfused': rel_err(dhb, dhb_u),
           'fwd': rel_err(out, pr.detach()), 'dy': rel_err(nchw(dy), yin.grad), 'dgamma': rel_err(dgamma, gr.grad),
           'dbeta': rel_err(dbeta, br.grad), 'dhw': rel_err(dhw.reshape(2, c), hwr.grad.reshape(2, c)), 'dhb': rel_err(dhb, hbr.grad)}
    # dy is stored in bf16: the per-channel sums of the two paths differ in their last fp32 bit (atomics order), which can
    # flip the rounding of single elements by one bf16 ulp -> 1e-3 of the largest element; everything else is fp32
    assert res['dy_vs_unfused'] < 1e-3, res
    assert max(res[k] for k in res if k.endswith('_vs_unfused') and k != 'dy_vs_unfused') < 2e-5, res
    assert res['fwd'] < 1e-2 and res['dy'] < 2e-2 and res['dgamma'] < 5e-3 and res['dbeta'] < 5e-3, res
    assert res['dhw'] < 5e-3 and res['dhb'] < 1e-5, res
    return res


# ------------------------------------------------------------------------------------------------ head losses
def check_masked_mse(b=5, s=48, seed=6):
    from oracle import cmunet_oracle as O
    g = _gen(seed)
    x = _randn((b, s, s), g)
    pred = _randn((b, 2, s, s), g).requires_grad_(True)
    mask = (torch.rand(b, s, s, generator=g) > 0.35).to(torch.uint8).to(DEV)
    with torch.no_grad():
        mean = x.mean(-1, keepdim=True)
        var = x.var(-1, keepdim=True)
        tgt = (x - mean) / (var + 1e-6) ** .5
    lr = 0.7 * (((pred[:, 1] - tgt) ** 2) * mask).sum() / mask.sum()
    lr.backward()
    loss, acc = ops.masked_mse_fwd(x, pred.detach(), mask, 0.7)
    gs = torch.tensor([0.7], device=DEV)
    dp = ops.masked_mse_bwd(x, pred.detach(), mask, acc, gs)
    torch.cuda.synchronize()
    res = {'loss': abs(float(loss) - float(lr)) / abs(float(lr)), 'dpred': rel_err(dp, pred.grad),
           'ch0_zero': float(dp[:, 0].abs().max())}
    assert res['loss'] < 1e-5 and res['dpred'] < 1e-4 and res['ch0_zero'] == 0.0, res
    return res


def check_infonce(b=16, world=4, rank=2, seed=7):
    g = _gen(seed)
    q = _randn((b, 256), g).requires_grad_(True)
    z = F.normalize(_randn((b * world, 256), g), dim=1)
    tau, ctw = 0.07, 1.3
    p = F.normalize(q, dim=1)
    score = p @ z.t() / tau
    label = torch.arange(b, device=DEV) + b * rank
    lr = ctw * 2 * tau * F.cross_entropy(score, label)
    lr.backward()
    loss, dq = ops.infonce(q.detach(), z, b * rank, tau, ctw)
    zn = ops.l2_normalize_rows(_randn((7, 256), _gen(seed + 1)))
    torch.cuda.synchronize()
    res = {'loss': abs(float(loss) - float(lr)) / abs(float(lr)), 'dq': rel_err(dq, q.grad),
           'norm': float((zn.norm(dim=1) - 1).abs().max())}
    assert res['loss'] < 1e-5 and res['dq'] < 1e-4 and res['norm'] < 1e-5, res
    return res


def check_seg_losses(n=3, s=40, seed=8):
    from oracle import cmunet_oracle as O
    g = _gen(seed)
    logits = _randn((n, 2, s, s), g).requires_grad_(True)
    y1 = (torch.rand(n, 1, s, s, generator=g) > 0.8).to(DEV)
    gt = torch.cat([~y1, y1], 1).double()
    dice, iou, ce = O.dice_loss(logits, gt), O.iou_loss(logits, gt), O.ce_prob_loss(logits, gt)
    ce.backward()
    gs = torch.tensor([1.0], device=DEV)
    out, dl = ops.seg_losses(logits.detach(), gt, gs)
    torch.cuda.synchronize()
    res = {'dice': abs(float(out[0]) - float(dice)), 'iou': abs(float(out[1]) - float(iou)),
           'ce': abs(float(out[2]) - float(ce)) / float(ce), 'dlogits': rel_err(dl, logits.grad)}
    assert res['dice'] < 1e-12 and res['iou'] < 1e-12 and res['ce'] < 1e-6 and res['dlogits'] < 1e-4, res
    return res


# ------------------------------------------------------------------------------------------------ mask generator
def check_mask(seed=60, b=6, s=128, ps=16, ratio=0.65, steps=2):
    from oracle.mask_oracle import MT19937, num_masked_patches, patch_mask
    st = torch.zeros(lib.cmu_mask_state_words(), dtype=torch.int32, device=DEV)
    stream = torch.cuda.current_stream().cuda_stream
    lib.cmu_mask_seed(st.data_ptr(), seed, stream)
    rng = MT19937(seed)
    k = num_masked_patches(s, ps, ratio)
    p = (s // ps) ** 2
    bad = 0
    for _ in range(steps):
        mask = torch.empty(b, s, s, dtype=torch.uint8, device=DEV)
        ws = torch.empty(max(lib.cmu_mask_workspace_bytes(b, p, k) // 4, 1), dtype=torch.int32, device=DEV)
        lib.cmu_mask_generate(st.data_ptr(), mask.data_ptr(), ws.data_ptr(), b, s, ps, k, b, stream)
        lib.cmu_mask_generate(st.data_ptr(), 0, 0, b, s, ps, 0, b, stream)      # target encoder: draws only (Q2)
        ref, _ = patch_mask(rng, b, s, ps, ratio)
        patch_mask(rng, b, s, ps, 0.0)
        bad += int((mask.cpu().numpy() != ref).sum())
    torch.cuda.synchronize()
    key, pos = rng.get_state()
    dev_state = st.cpu().numpy().astype(np.uint32)
    # positions may differ by a pending regeneration (624 == "regenerate first"): compare the next draws instead
    nxt_ref = [rng.next_u32() for _ in range(3)]
    r2 = MT19937()
    r2.set_state(dev_state[:624], int(dev_state[624]))
    nxt_dev = [r2.next_u32() for _ in range(3)]
    res = {'mismatch_bytes': bad, 'stream_ok': nxt_ref == nxt_dev, 'K': k, 'P': p}
    assert bad == 0 and nxt_ref == nxt_dev, res
    return res


def check_mask_pair_mode(seed=60, b=5, s=96, steps=4):
    """CM_UNet's prefetching stream (next step's online+target pair generated on a side stream) serves exactly the
    masks of the sequential reference order, survives a batch-size change, and reports the logical stream position."""
    import contrastive_masked_unet_b200 as C
    from oracle.mask_oracle import MT19937, patch_mask
    ms = C.MaskStream()
    ms.pair_mode = True
    ms.seed(seed, DEV)
    rng = MT19937(seed)
    dev = torch.device(DEV)
    bad = 0
    batches = [b, b, b, b + 2, b + 2, b]
    for step_i, bb in enumerate(batches[:steps + 2]):
        m_on, k = ms.generate(bb, s, 16, 0.65, dev)
        m_tg, k0 = ms.generate(bb, s, 16, 0.0, dev)
        if step_i != 1:
            ms.fire_deferred_prefetch()        # what the first layer's backward node does at the end of a training step
        # (step 1: no backward ran -> the deferred prefetch is dropped and the next pair is generated synchronously)
        ref, _ = patch_mask(rng, bb, s, 16, 0.65)
        patch_mask(rng, bb, s, 16, 0.0)
        torch.cuda.synchronize()
        bad += int((m_on.cpu().numpy() != ref).sum()) + int(m_tg.sum())
    st = ms.get_numpy_state()
    r2 = MT19937()
    r2.set_state(st[1], st[2])
    ok = [r2.next_u32() for _ in range(3)] == [rng.next_u32() for _ in range(3)]
    res = {'mismatch_bytes': bad, 'logical_stream_ok': ok, 'prefetch_outstanding': ms._pref is not None}
    assert bad == 0 and ok and res['prefetch_outstanding'], res
    return res


def check_mask_interleaved_calls(seed=62, b=4, s=96):
    """Speculative pair prefetch stays bit-exact with numpy's order for call sequences that are NOT online -> target:
    mode='tensor' style online-only calls after a training step, get_numpy_state in between, a size change, pickling
    (the prefetched target half must be rolled back when no target call follows)."""
    import pickle
    import contrastive_masked_unet_b200 as C
    from oracle.mask_oracle import MT19937, patch_mask
    ms = C.MaskStream()
    ms.pair_mode = True
    ms.seed(seed, DEV)
    rng = MT19937(seed)
    dev = torch.device(DEV)
    bad = 0
    ok = True

    def online(bb, ss):
        nonlocal bad
        m, _ = ms.generate(bb, ss, 16, 0.65, dev)
        ref, _ = patch_mask(rng, bb, ss, 16, 0.65)
        torch.cuda.synchronize()
        bad += int((m.cpu().numpy() != ref).sum())

    def target(bb, ss):
        nonlocal bad
        m, _ = ms.generate(bb, ss, 16, 0.0, dev)
        ms.fire_deferred_prefetch()        # end of the training step's backward
        patch_mask(rng, bb, ss, 16, 0.0)
        bad += int(m.sum())

    def position_ok():
        st = ms.get_numpy_state()
        r2 = MT19937()
        r2.set_state(st[1], st[2])
        key, pos = rng.get_state()
        r3 = MT19937()
        r3.set_state(key, pos)
        return [r2.next_u32() for _ in range(3)] == [r3.next_u32() for _ in range(3)]

    online(b, s); target(b, s)            # training step -> pair of the next step is prefetched
    online(b, s)                          # extract_feat: consumes the prefetched online half, no target call follows
    ok &= position_ok()                   # logical position = +B, although +2B were drawn speculatively
    online(b, s)                          # second extract_feat: must start at +B (roll back the speculative target half)
    ok &= position_ok()
    online(b, s); target(b, s)            # training resumes
    online(b, s); target(b, s)
    ms2 = pickle.loads(pickle.dumps(ms))  # pickled at a point where a prefetched pair is outstanding
    ok &= position_ok()
    online(b + 1, s); target(b + 1, s)    # batch change drops the prefetch
    online(b + 1, s)
    ms.pair_mode = False
    online(b, s + 32)                     # plain mode after an owed target half
    ok &= position_ok()
    # the unpickled copy continues from the logical position it was saved at
    rng2 = MT19937(seed)
    for _ in range(4):
        patch_mask(rng2, b, s, 16, 0.65)
    for _ in range(3):
        patch_mask(rng2, b, s, 16, 0.0)
    # (sequence so far: on,tg,on,on,on,tg,on,tg = 5 online + 3 target calls of batch b)
    patch_mask(rng2, b, s, 16, 0.65)
    m2, _ = ms2.generate(b, s, 16, 0.65, dev)
    ref2, _ = patch_mask(rng2, b, s, 16, 0.65)
    torch.cuda.synchronize()
    bad2 = int((m2.cpu().numpy() != ref2).sum())
    res = {'mismatch_bytes': bad, 'positions_ok': bool(ok), 'unpickled_mismatch': bad2}
    assert bad == 0 and ok and bad2 == 0, res
    return res


def time_mask(b=64, s=512, reps=5):
    """ms per training-step pair (B online + B discarded target shuffles) of the mask generator."""
    st = torch.zeros(lib.cmu_mask_state_words(), dtype=torch.int32, device=DEV)
    stream = torch.cuda.current_stream().cuda_stream
    lib.cmu_mask_seed(st.data_ptr(), 60, stream)
    p = (s // 16) ** 2
    k = int(0.65 * s * s) // 256
    mask = torch.empty(b, s, s, dtype=torch.uint8, device=DEV)
    ws = torch.empty(lib.cmu_mask_workspace_bytes(b, p, k) // 4, dtype=torch.int32, device=DEV)
    ts = []
    for _ in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.cmu_mask_generate(st.data_ptr(), mask.data_ptr(), ws.data_ptr(), b, s, 16, k, b, stream)
        lib.cmu_mask_generate(st.data_ptr(), 0, 0, b, s, 16, 0, b, stream)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return {'ms_per_pair': sorted(ts[2:])[len(ts[2:]) // 2], 'B': b, 'S': s}


# ------------------------------------------------------------------------------------------------ linear / BN1d / optim
def check_linear(m=8, k=4096 * 3, n=192, seed=9):
    g = _gen(seed)
    x = _randn((m, k), g)
    w = _randn((n, k), g, k ** -0.5)
    b = _randn((n,), g)
    dy = _randn((m, n), g)
    y = ops.linear_fwd(x, w, b)
    dx = ops.linear_dgrad(dy, w)
    dw = ops.linear_wgrad(dy, x)
    db = ops.colsum(dy)
    torch.cuda.synchronize()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    res = {'fwd': rel_err(y, (x.double() @ w.double().t() + b.double())), 'dx': rel_err(dx, dy.double() @ w.double()),
           'dw': rel_err(dw, dy.double().t() @ x.double()), 'db': rel_err(db, dy.sum(0))}
    torch.backends.cuda.matmul.allow_tf32 = prev
    assert max(res.values()) < 1e-4, res
    return res


def check_gemm_tn(rows=1000, qc=64, pc=256, seed=19, bias=True):
    g = _gen(seed)
    q = _randn((rows, qc), g)
    p = _randn((rows, pc), g)
    b = _randn((pc,), g) if bias else None
    qt16, q16 = ops.transpose_cast(q.t().contiguous(), also_plain=True)    # qt16: (rows,qc) from a (qc,rows) source
    p16 = ops.cast_bf16(p)
    out = ops.gemm_tn(qt16, p16, bias=b)
    ref = qt16.double().t() @ p16.double() + (b.double() if bias else 0)
    torch.cuda.synchronize()
    res = {'gemm': rel_err(out, ref), 'tcast': float((qt16.float() - q.to(BF16).float()).abs().max()),
           'cast': float((q16.float() - q.t().to(BF16).float()).abs().max())}
    assert res['gemm'] < 1e-4 and res['tcast'] == 0.0 and res['cast'] == 0.0, res
    return res


def check_linear_tc(m=64, k=4096, n=1536, seed=20):
    """LinearFn on the tensor-core path vs fp32 autograd on the bf16-rounded operands."""
    from contrastive_masked_unet_b200 import functional as Fn
    g = _gen(seed)
    x = _randn((m, k), g).to(BF16).float().requires_grad_(True)
    w = _randn((n, k), g, k ** -0.5).to(BF16).float().requires_grad_(True)
    b = _randn((n,), g).requires_grad_(True)
    dy = _randn((m, n), g).to(BF16).float()
    assert ops.tc_linear_ok(m, k, n)
    y = Fn.LinearFn.apply(x, w, b)
    y.backward(dy)
    gx, gw, gb = x.grad.clone(), w.grad.clone(), b.grad.clone()
    x.grad = w.grad = b.grad = None
    yr = torch.nn.functional.linear(x.double(), w.double(), b.double())
    yr.backward(dy.double())
    torch.cuda.synchronize()
    res = {'y': rel_err(y, yr.detach()), 'dx': rel_err(gx, x.grad), 'dw': rel_err(gw, w.grad), 'db': rel_err(gb, b.grad)}
    assert max(res.values()) < 1e-4, res
    return res


def check_moco_head(n=64, d=1024, kneg=8192, seed=40):
    """MoCo queue head (configs[3]) vs the oracle restatement: loss, dq, enqueue."""
    from contrastive_masked_unet_b200.moco import MocoLossFn
    from oracle import moco_oracle as MO
    g = _gen(seed)
    q = _randn((n, d), g).requires_grad_(True)
    k = F.normalize(_randn((n, d), g), dim=1)
    queue = F.normalize(_randn((d, kneg), g), dim=0)
    rows = queue.t().contiguous().to(BF16)
    qref = rows.float().t().contiguous()                       # the oracle sees the bf16-rounded queue
    lr = MO.moco_loss(q, k, qref, 0.07)
    lr.backward()
    q2 = q.detach().clone().requires_grad_(True)
    loss = MocoLossFn.apply(q2, k, rows, 0.07)
    loss.backward()
    # enqueue
    keys = F.normalize(_randn((n, d), _gen(seed + 1)), dim=1)
    qref2 = queue.clone()
    ptr = MO.dequeue_and_enqueue(qref2, kneg - n, keys)
    qdev = queue.clone()
    lib.cmu_queue_enqueue(keys.data_ptr(), n, d, kneg, kneg - n, rows.data_ptr(), qdev.data_ptr(),
                          torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    res = {'loss': abs(float(loss) - float(lr)) / float(lr), 'dq': rel_err(q2.grad, q.grad), 'dq_cos': cos(q2.grad, q.grad),
           'enqueue_ref': float((qdev - qref2).abs().max()),
           'enqueue_rows': float((rows[kneg - n:].float() - keys.to(BF16).float()).abs().max()), 'ptr': ptr}
    assert res['loss'] < 2e-3 and res['dq_cos'] > 0.999 and res['dq'] < 5e-2, res
    assert res['enqueue_ref'] == 0.0 and res['enqueue_rows'] == 0.0 and ptr == 0, res
    return res


def check_spatial_mean(n=3, h=8, w=6, c=1024, seed=41):
    from contrastive_masked_unet_b200.moco import SpatialMeanFn
    g = _gen(seed)
    x = nhwc(_randn((n, c, h, w), g))
    xr = nchw(x).requires_grad_(True)
    out = SpatialMeanFn.apply(x.permute(0, 3, 1, 2).requires_grad_(True))
    ref = torch.mean(xr, dim=[2, 3])
    torch.cuda.synchronize()
    res = {'mean': rel_err(out, ref.detach())}
    assert res['mean'] < 1e-5, res
    return res


def check_cldice():
    """soft-clDice metric vs the oracle and the golden values minted from the reference (bit-exact on binary masks)."""
    import json
    import os
    import contrastive_masked_unet_b200 as C
    from oracle import cmunet_oracle as O
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'cldice.json')))['cases']
    m = C.soft_cldice(threshold=0.5, activation='softmax', ignore_channels=[0])
    res = {'name': m.__name__}
    for c in gold:
        logits, gt = O.cldice_inputs(c['n'], c['h'], c['w'], c['seed'])
        v = m(logits.cuda(), gt.cuda())
        ref = O.cldice_loss(logits, gt)
        res[f"{c['n']}x{c['h']}x{c['w']}"] = (float(v), float(ref), c['cldice'])
        assert v.dtype == torch.float64 and abs(float(v) - c['cldice']) < 1e-12 and abs(float(v) - float(ref)) < 1e-12, res
    # soft (non-binary) target: float64 morphology must still agree
    g = _gen(77)
    logits = torch.randn(2, 2, 70, 45, generator=g).cuda()
    y1 = torch.rand(2, 1, 70, 45, generator=g).double().cuda()
    gt = torch.cat([1 - y1, y1], 1)
    v, ref = m(logits, gt), O.cldice_loss(logits, gt)
    res['soft_target'] = (float(v), float(ref))
    assert abs(float(v) - float(ref)) < 1e-9, res
    return res


def check_bn1d(m=12, c=1536, seed=10):
    g = _gen(seed)
    x = _randn((m, c), g, 2.0)
    gamma = (torch.rand(c, generator=g) + 0.5).to(DEV)
    beta = _randn((c,), g, 0.3)
    dy = _randn((m, c), g)
    stream = torch.cuda.current_stream().cuda_stream
    stats = torch.empty(2, c, device=DEV)
    y = torch.empty_like(x)
    mean, rstd = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    lib.cmu_bn1d_stats(x.data_ptr(), m, c, stats.data_ptr(), stream)
    lib.cmu_bn1d_apply(x.data_ptr(), stats.data_ptr(), float(m), m, c, gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(),
                       rv.data_ptr(), 0.1, 1e-6, 1, 1, y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), stream)
    sums = torch.empty(2, c, device=DEV)
    dx = torch.empty_like(x)
    lib.cmu_bn1d_bwd_stats(dy.data_ptr(), y.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), m, c, 1,
                           sums.data_ptr(), stream)
    lib.cmu_bn1d_bwd_apply(dy.data_ptr(), y.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                           sums.data_ptr(), float(m), m, c, 1, dx.data_ptr(), stream)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_r, rv_r = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    yr = F.relu(F.batch_norm(xr, rm_r, rv_r, gr, br, True, 0.1, 1e-6))
    yr.backward(dy)
    torch.cuda.synchronize()
    res = {'y': rel_err(y, yr.detach()), 'dx': rel_err(dx, xr.grad), 'dgamma': rel_err(sums[1], gr.grad),
           'dbeta': rel_err(sums[0], br.grad), 'rm': rel_err(rm, rm_r), 'rv': rel_err(rv, rv_r)}
    assert max(res.values()) < 2e-4, res
    return res


def check_optim(seed=11):
    g = _gen(seed)
    sizes = [5, 1024, 70001, 33]
    stream = torch.cuda.current_stream().cuda_stream
    dst = [_randn((s,), g) for s in sizes]
    src = [_randn((s,), g) for s in sizes]
    ref = [d * 0.996 + s * (1. - 0.996) for d, s in zip(dst, src)]
    rows = []
    CH = 65536
    for d, s in zip(dst, src):
        for off in range(0, d.numel(), CH):
            rows.append([d.data_ptr() + 4 * off, s.data_ptr() + 4 * off, min(CH, d.numel() - off)])
    table = torch.tensor(rows, dtype=torch.int64, device=DEV)
    lib.cmu_ema_chunks(table.data_ptr(), len(rows), 0.996, stream)
    torch.cuda.synchronize()
    res = {'ema': max(float((d - r).abs().max()) for d, r in zip(dst, ref))}
    # AdamW vs torch.optim.AdamW, 3 steps, decay on tensor 0 and 2 only
    ps = [_randn((s,), g).requires_grad_(True) for s in sizes]
    mine = [p.detach().clone() for p in ps]
    ms = [torch.zeros_like(p) for p in mine]
    vs = [torch.zeros_like(p) for p in mine]
    opt = torch.optim.AdamW([{'params': [ps[0], ps[2]], 'weight_decay': 0.05}, {'params': [ps[1], ps[3]], 'weight_decay': 0.0}],
                            lr=1e-2, betas=(0.9, 0.95), eps=1e-8)
    for step in range(1, 4):
        grads = [_randn((s,), g) for s in sizes]
        for p, gr in zip(ps, grads):
            p.grad = gr.clone()
        opt.step()
        rows = []
        for i, (p, gr, m_, v_) in enumerate(zip(mine, grads, ms, vs)):
            for off in range(0, p.numel(), CH):
                rows.append([p.data_ptr() + 4 * off, gr.data_ptr() + 4 * off, m_.data_ptr() + 4 * off,
                             v_.data_ptr() + 4 * off, min(CH, p.numel() - off), 1 if i in (0, 2) else 0])
        table = torch.tensor(rows, dtype=torch.int64, device=DEV)
        lib.cmu_adamw_chunks(table.data_ptr(), len(rows), 1e-2, 0.9, 0.95, 1e-8, 0.05, step, 1.0, stream)
        torch.cuda.synchronize()
    res['adamw'] = max(float((a - b.detach()).abs().max()) for a, b in zip(mine, ps))
    assert res['ema'] < 1e-6 and res['adamw'] < 2e-5, res
    return res


def check_optim_amp(seed=12):
    """FusedAdamW + DynamicLossScaler against torch.optim.AdamW + torch.amp.GradScaler over 6 steps with an overflow
    injected at step 3: skipped update, halved scale, bias correction by the count of successful steps, growth after
    `growth_interval` clean steps (mmengine AmpOptimWrapper(loss_scale='dynamic'), cmunet_config.py:76-78)."""
    from contrastive_masked_unet_b200.optim import DynamicLossScaler, FusedAdamW
    g = _gen(seed)
    sizes = [7, 3000, 70001]
    ref = [_randn((s,), g).requires_grad_(True) for s in sizes]
    mine = [p.detach().clone().requires_grad_(True) for p in ref]
    opt_r = torch.optim.AdamW(ref, lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    sc_r = torch.amp.GradScaler('cuda', init_scale=1024.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2)
    opt_m = FusedAdamW([(f'w{i}', p) for i, p in enumerate(mine)], lr=1e-2, weight_decay=0.05, no_decay_keys=())
    sc_m = DynamicLossScaler(DEV, init_scale=1024.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2)
    scales = []
    for step in range(6):
        grads = [_randn((s,), g) for s in sizes]
        if step == 2:
            grads[1][17] = float('inf')
        sc_r.scale(torch.zeros(1, device=DEV))            # GradScaler initialises its device scale lazily in scale()
        s_now = sc_r.get_scale()
        for p, q, gr in zip(ref, mine, grads):
            p.grad = (gr * s_now).clone()
            q.grad = (gr * sc_m.get_scale()).clone()
        sc_r.step(opt_r)
        sc_r.update()
        opt_m.step(scaler=sc_m)
        torch.cuda.synchronize()
        scales.append((sc_m.get_scale(), sc_r.get_scale()))
    res = {'scales': scales, 'steps_taken': sc_m.steps_taken(),
           'param_err': max(float((a.detach() - b.detach()).abs().max()) for a, b in zip(mine, ref))}
    assert all(a == b for a, b in scales), res
    assert res['steps_taken'] == 5 and res['param_err'] < 2e-5, res
    assert all(bool(torch.isfinite(p).all()) for p in mine)
    return res


def check_sgd(seed=13):
    from contrastive_masked_unet_b200.optim import FusedSGD
    g = _gen(seed)
    sizes = [5, 4097, 70001]
    ref = [_randn((s,), g).requires_grad_(True) for s in sizes]
    mine = [p.detach().clone().requires_grad_(True) for p in ref]
    opt_r = torch.optim.SGD(ref, lr=0.03, momentum=0.9, weight_decay=1e-4)
    opt_m = FusedSGD([(f'w{i}', p) for i, p in enumerate(mine)], lr=0.03, momentum=0.9, weight_decay=1e-4)
    for _ in range(4):
        grads = [_randn((s,), g) for s in sizes]
        for p, q, gr in zip(ref, mine, grads):
            p.grad = gr.clone()
            q.grad = gr.clone()
        opt_r.step()
        opt_m.step()
    torch.cuda.synchronize()
    res = {'sgd': max(float((a.detach() - b.detach()).abs().max()) for a, b in zip(mine, ref))}
    assert res['sgd'] < 1e-5, res
    return res


CHECKS = {
    'conv3x3_64_64': lambda: check_conv3x3(2, 24, 40, 64, 0, 64),
    'conv3x3_128_128': lambda: check_conv3x3(2, 16, 16, 128, 0, 128, seed=1),
    'conv3x3_cat_64+64_64': lambda: check_conv3x3(2, 24, 24, 64, 64, 64, seed=2),
    'conv3x3_256_128_ragged': lambda: check_conv3x3(3, 14, 14, 256, 0, 128, seed=3),
    'conv3x3_512_1024_tiny': lambda: check_conv3x3(2, 4, 4, 512, 0, 1024, seed=4),
    'conv3x3_cat_512+512_512': lambda: check_conv3x3(1, 8, 8, 512, 512, 512, seed=5),
    'conv3x3_64_128_big': lambda: check_conv3x3(4, 128, 128, 64, 0, 128, seed=6),
    'conv3x3_64_64_resident_weights': lambda: check_conv3x3(6, 128, 128, 64, 0, 64, seed=7),
    'conv3x3_cat_64+64_64_resident': lambda: check_conv3x3(6, 128, 128, 64, 64, 64, seed=8),
    'conv3x3_64_128_resident': lambda: check_conv3x3(6, 128, 128, 64, 0, 128, seed=9),
    'convT_128_64_resident': lambda: check_convT(8, 64, 128, 128, 64, seed=14),
    'conv3x3_pair_128_128': lambda: check_conv3x3(6, 128, 128, 128, 0, 128, seed=30),
    'conv3x3_pair_cat_128+128_256': lambda: check_conv3x3(8, 96, 96, 128, 128, 256, seed=31),
    'conv3x3_pair_odd_tiles_256_256': lambda: check_conv3x3(7, 88, 72, 256, 0, 256, seed=32),
    'conv3x3_pair_64_128': lambda: check_conv3x3(6, 128, 128, 64, 0, 128, seed=33),
    'conv3x3_pair_64_64': lambda: check_conv3x3(6, 128, 128, 64, 0, 64, seed=34),
    'conv3x3_pair_cat_64+64_64': lambda: check_conv3x3(5, 136, 120, 64, 64, 64, seed=35),
    # single-patch staging (N = 64 tiles) on ragged shapes: W and H not multiples of the 8 x 16 tile, odd tile count
    'conv3x3_pair_64_64_ragged': lambda: check_conv3x3(5, 100, 92, 64, 0, 64, seed=36),
    'conv3x3_pair_128_64_ragged_odd': lambda: check_conv3x3(3, 140, 83, 128, 0, 64, seed=37),
    # BASELINE.json configs[1] full sizes (per-GPU batch 64 @512^2): first-level and bottleneck-level layers
    'conv3x3_full_size_enc1_conv2': lambda: check_conv3x3(64, 512, 512, 64, 0, 64, seed=50),
    'conv3x3_full_size_up4_conv1': lambda: check_conv3x3(64, 64, 64, 512, 512, 512, seed=51),
    'convT_full_size_up1': lambda: check_convT(64, 256, 256, 128, 64, seed=52),
    'conv3x3_fprop_only_64_64': lambda: check_conv3x3_fprop_only(2, 24, 40, 64, 0, 64),
    'conv3x3_wgrad_only_64_64': lambda: check_conv3x3_wgrad_only(2, 24, 40, 64, 0, 64),
    'conv3x3_wgrad_only_128_128': lambda: check_conv3x3_wgrad_only(2, 16, 16, 128, 0, 128, seed=1),
    'conv3x3_c1': check_conv3x3_c1,
    'conv3x3_c1_ragged_rows': lambda: check_conv3x3_c1(2, 30, 52, seed=2),   # H % 4 != 0, W % 32 != 0
    'conv3x3_c1_224': lambda: check_conv3x3_c1(2, 224, 224, seed=3),
    'convT_128_64': lambda: check_convT(2, 12, 20, 128, 64),
    'convT_256_128_pair_ragged': lambda: check_convT(3, 120, 136, 256, 128, seed=15),
    'convT_512_256_pair': lambda: check_convT(17, 32, 32, 512, 256, seed=16),
    'convT_1024_512': lambda: check_convT(2, 4, 4, 1024, 512, seed=12),
    'convT_256_128_ragged': lambda: check_convT(2, 14, 14, 256, 128, seed=13),
    'conv1x1_1024_256': check_conv1x1,
    'head1x1': check_head1x1,
    'bn_relu_head_fused': check_bn_relu_head,
    'bn_relu_head_fused_ragged': lambda: check_bn_relu_head(2, 13, 37, seed=22),
    'bn_relu_head_fused_eval': lambda: check_bn_relu_head(2, 16, 16, seed=23, training=False),
    'bn_pool': lambda: check_bn(3, 16, 24, 64, True),
    'bn_nopool_c256': lambda: check_bn(2, 10, 6, 256, False, seed=15),
    'bn_pool_c1024': lambda: check_bn(2, 4, 4, 1024, True, seed=16),
    'masked_mse': check_masked_mse,
    'infonce': check_infonce,
    'infonce_w1': lambda: check_infonce(64, 1, 0, seed=17),
    'seg_losses': check_seg_losses,
    'mask_128': check_mask,
    'mask_512_b64': lambda: check_mask(60, 64, 512, 16, 0.65, 2),
    'mask_224': lambda: check_mask(61, 4, 224, 16, 0.65, 2),
    'mask_ps8': lambda: check_mask(3, 2, 128, 8, 0.5, 1),
    'mask_pair_mode_prefetch': check_mask_pair_mode,
    'mask_interleaved_calls': check_mask_interleaved_calls,
    'mask_1024_p4096': lambda: check_mask(7, 3, 1024, 16, 0.65, 1),
    'mask_tiny_p4': lambda: check_mask(8, 5, 32, 16, 0.65, 2),
    'linear': check_linear,
    'linear_small': lambda: check_linear(64, 256, 1536, seed=18),
    'gemm_tn': check_gemm_tn,
    'gemm_tn_big_n': lambda: check_gemm_tn(64, 1536, 4096 + 64, seed=21, bias=False),
    'gemm_tn_splitk': lambda: check_gemm_tn(50000, 64, 128, seed=22),
    'linear_tc': check_linear_tc,
    'linear_tc_128rows': lambda: check_linear_tc(128, 8192, 512, seed=23),
    'moco_head_8k': check_moco_head,
    'moco_head_64k': lambda: check_moco_head(64, 1024, 65536, seed=42),
    'spatial_mean': check_spatial_mean,
    'cldice': check_cldice,
    'bn1d': check_bn1d,
    'optim': check_optim,
    'optim_amp_dynamic_loss_scale': check_optim_amp,
    'sgd_momentum': check_sgd,
}
