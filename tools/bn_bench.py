"""Micro-benchmark of the HBM-bound BatchNorm backward / column-sum kernels at the benchmark shapes (CUDA events, L2 flushed).
    python tools/bn_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from contrastive_masked_unet_b200 import ops  # noqa: E402

BF16 = torch.bfloat16


def timeit(fn, flush, iters=7):
    fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    dev = 'cuda'
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for B, S, C in ((64, 512, 64), (64, 256, 128), (64, 128, 256), (64, 64, 512)):
        y = torch.randn(B, S, S, C, device=dev).to(BF16)
        da = torch.randn(B, S, S, C, device=dev).to(BF16)
        dp = torch.randn(B, S // 2, S // 2, C, device=dev).to(BF16)
        scale, shift = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
        mean, rstd = torch.randn(C, device=dev) * 0.1, torch.rand(C, device=dev) + 0.5
        nbytes = y.numel() * 2
        t1 = timeit(lambda: ops.bn_relu_bwd(da, None, y, scale, shift, mean, rstd), flush)
        t2 = timeit(lambda: ops.bn_relu_bwd(da, dp, y, scale, shift, mean, rstd), flush)
        t3 = timeit(lambda: ops.colsum_bf16(B * S * S, C, y), flush)
        t4 = timeit(lambda: ops.bn_relu_apply(y, scale, shift, False), flush)
        print(f'B{B} S{S} C{C}: bn_bwd {t1:.3f} ms ({5 * nbytes / t1 / 1e9:.2f} TB/s)  bn_bwd+pool {t2:.3f} ms '
              f'({5.5 * nbytes / t2 / 1e9:.2f} TB/s)  colsum {t3:.3f} ms ({nbytes / t3 / 1e9:.2f} TB/s)  '
              f'bn_relu {t4:.3f} ms ({2 * nbytes / t4 / 1e9:.2f} TB/s)', flush=True)


def misc():
    dev = 'cuda'
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B, S = 64, 512
    a = torch.randn(B, S, S, 64, device=dev).to(BF16)
    w, b = torch.randn(2, 64, device=dev), torch.randn(2, device=dev)
    dout = torch.randn(B, 2, S, S, device=dev)
    nb = a.numel() * 2
    t1 = timeit(lambda: ops.head1x1_fprop(a, w, b), flush)
    t2 = timeit(lambda: ops.head1x1_bwd(a, w, dout), flush)
    x = torch.randn(1536, S * S, device=dev)
    t3 = timeit(lambda: ops.transpose_cast(x, False), flush)
    t4 = timeit(lambda: ops.transpose_cast(x, True), flush)
    xb = x.numel()
    img = torch.randn(B, S, S, device=dev)
    mask0 = (torch.rand(S, S, device=dev) > 0.35).to(torch.uint8)
    w1 = torch.randn(64, 1, 3, 3, device=dev) * 0.3
    t5 = timeit(lambda: ops.conv3x3_c1_fprop(img, mask0, w1), flush)
    t6 = timeit(lambda: ops.conv3x3_c1_wgrad(img, mask0, a), flush)
    dy = torch.randn(64, 1536, device=dev)
    wfc = torch.randn(1536, S * S, device=dev) * 0.01
    xin = torch.randn(64, S * S, device=dev)
    from contrastive_masked_unet_b200 import functional as Fn
    def lin():
        xx = xin.clone().requires_grad_(True)
        ww = wfc.requires_grad_(True)
        ww.grad = None
        y = Fn.LinearFn.apply(xx, ww, None)
        y.backward(dy)
    t7 = timeit(lin, flush)
    print(f'projector fc0 fwd+bwd (64 x {S * S} -> 1536): {t7:.3f} ms', flush=True)
    print(f'conv_c1 fprop {t5:.3f} ms ({nb / t5 / 1e9:.2f} TB/s)  wgrad {t6:.3f} ms ({nb / t6 / 1e9:.2f} TB/s)', flush=True)
    print(f'head1x1 fwd {t1:.3f} ms ({(nb + dout.numel() * 4) / t1 / 1e9:.2f} TB/s)  bwd {t2:.3f} ms '
          f'({(2 * nb + dout.numel() * 4) / t2 / 1e9:.2f} TB/s)  transpose_cast {t3:.3f} ms ({6 * xb / t3 / 1e9:.2f} TB/s)  '
          f'+plain {t4:.3f} ms ({8 * xb / t4 / 1e9:.2f} TB/s)', flush=True)


def head():
    """fused decoder tail (csrc/head_fused.cu) at the benchmark shape"""
    dev = 'cuda'
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B, S, C = 64, 512, 64
    y = torch.randn(B, S, S, C, device=dev).to(BF16)
    scale, shift = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    mean, rstd = torch.randn(C, device=dev) * 0.1, torch.rand(C, device=dev) + 0.5
    w, b = torch.randn(2, C, device=dev), torch.randn(2, device=dev)
    dout = torch.randn(B, 2, S, S, device=dev)
    nb, ob = y.numel() * 2, dout.numel() * 4
    t1 = timeit(lambda: ops.bn_relu_head_fwd(y, scale, shift, w, b), flush)
    t2 = timeit(lambda: ops.bn_relu_head_bwd(y, scale, shift, mean, rstd, w, dout), flush)
    print(f'fused tail fwd {t1:.3f} ms ({(nb + ob) / t1 / 1e9:.2f} TB/s: y read, out written)  '
          f'bwd (reduce + apply) {t2:.3f} ms ({(3 * nb + 2 * ob) / t2 / 1e9:.2f} TB/s: y, dout read twice, dy written)', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'head':
        head()
    else:
        misc()
        main()
        head()
