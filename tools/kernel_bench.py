"""Per-layer micro-benchmark of the tcgen05 kernels at the benchmark shapes (B=64, S=512), all variants in ONE process
(same GPU, same clocks) for A/B comparisons.  CUDA events, L2 flushed between iterations.
    python tools/kernel_bench.py [--knobs "3=1;1=64"] [--only enc1,up1] [--batch 64]
Prints ms and TFLOP/s per layer for the default configuration and for every knob setting given."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from contrastive_masked_unet_b200 import ops  # noqa: E402
from contrastive_masked_unet_b200._lib import lib  # noqa: E402

BF16 = torch.bfloat16


def layers(B):
    # (name, kind, cin0, cin1, cout, S)
    L = [('enc1.c2', 'conv', 64, 0, 64, 512), ('enc2.c1', 'conv', 64, 0, 128, 256), ('enc2.c2', 'conv', 128, 0, 128, 256),
         ('enc3.c2', 'conv', 256, 0, 256, 128), ('enc4.c2', 'conv', 512, 0, 512, 64), ('enc5.c2', 'conv', 1024, 0, 1024, 32),
         ('up4.c1', 'conv', 512, 512, 512, 64), ('up2.c1', 'conv', 128, 128, 128, 256), ('up1.c1', 'conv', 64, 64, 64, 512),
         ('up1.T', 'convT', 128, 0, 64, 256), ('up2.T', 'convT', 256, 0, 128, 128), ('up4.T', 'convT', 1024, 0, 512, 32)]
    return L


def time_fn(fn, flush, iters=5, warm=2):
    for _ in range(warm):
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
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--knobs', default='')
    ap.add_argument('--only', default='')
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--reps', type=int, default=7)
    args = ap.parse_args()
    B = args.batch
    dev = 'cuda'
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    variants = [('default', [])] + [(kv, [tuple(int(x) for x in one.split('=')) for one in kv.split(',')])
                                    for kv in filter(None, args.knobs.split(';'))]
    results = {}
    for name, kind, c0, c1, cout, S in layers(B):
        if args.only and not any(s in name for s in args.only.split(',')):
            continue
        g = torch.Generator(device='cpu').manual_seed(0)
        if kind == 'conv':
            x0 = torch.randn(B, S, S, c0, device=dev).to(BF16)
            x1 = torch.randn(B, S, S, c1, device=dev).to(BF16) if c1 else None
            dy = torch.randn(B, S, S, cout, device=dev).to(BF16)
            w = torch.randn(cout, c0 + c1, 3, 3, device=dev) * 0.05
            wf, wd = ops.pack_conv3x3(w)
            flops = 2.0 * B * S * S * cout * 9 * (c0 + c1)
            fns = {'fprop': lambda: ops.conv3x3_fprop(x0, x1, wf, True), 'dgrad': lambda: ops.conv3x3_dgrad(dy, wd, c0, c1),
                   'wgrad': lambda: ops.conv3x3_wgrad(x0, x1, dy)}
        else:
            x0 = torch.randn(B, S, S, c0, device=dev).to(BF16)
            dy = torch.randn(B, 2 * S, 2 * S, cout, device=dev).to(BF16)
            w = torch.randn(c0, cout, 2, 2, device=dev) * 0.05
            bias = torch.zeros(cout, device=dev)
            wf, wd = ops.pack_convT2x2(w)
            flops = 2.0 * B * S * S * c0 * 4 * cout
            fns = {'fprop': lambda: ops.convT2x2_fprop(x0, wf, bias), 'dgrad': lambda: ops.convT2x2_dgrad(dy, wd),
                   'wgrad': lambda: ops.convT2x2_wgrad(x0, dy)}
        # variants are interleaved (A B A B ...) so that clock / power drift hits all of them alike; median per variant
        for op, fn in fns.items():
            samples = {vname: [] for vname, _ in variants}
            for vname, knobs in variants:       # warm-up of every variant
                for k, v in knobs:
                    lib.cmu_debug_set(k, v)
                fn()
                for k, v in knobs:
                    lib.cmu_debug_set(k, 0)
            for _ in range(args.reps):
                for vname, knobs in variants:
                    for k, v in knobs:
                        lib.cmu_debug_set(k, v)
                    samples[vname].append(time_fn(fn, flush, iters=1, warm=0))
                    for k, v in knobs:
                        lib.cmu_debug_set(k, 0)
            for vname, _ in variants:
                ts = sorted(samples[vname])
                ms = ts[len(ts) // 2]
                results[f'{name}.{op}[{vname}]'] = (ms, flops / ms / 1e9)
        line = f'{name:8s}'
        for op in fns:
            line += f' | {op}:'
            for vname, _ in variants:
                ms, tf = results[f'{name}.{op}[{vname}]']
                line += f' {ms:6.3f}ms {tf:6.0f}TF'
        print(line, flush=True)
        del x0, dy
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, 'gpurun_out', 'kernel_bench.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
