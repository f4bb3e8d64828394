#!/usr/bin/env python
"""Diagnostic: time the GPU data pipeline (contrastive_masked_unet_b200/data.py) for one B = 64 batch
of raw 512 x 512 images and, beside it, the same per-sample work done with Pillow + numpy on ONE host core (what one
DataLoader worker of the reference does, cmae/datasets/cmunet_dataset.py:74-88).  Writes
gpurun_out/data_pipeline_timing.json.  Not part of the product path."""
import json
import os
import random as pyrandom
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import contrastive_masked_unet_b200 as C  # noqa: E402
from oracle import data_oracle as D  # noqa: E402


def pillow_sample(raw, prm):
    """One sample through Pillow itself (Image.resize(BICUBIC)), the host-side twin of `D.sample_pipeline`."""
    from PIL import Image
    mode = 'L' if raw.dtype == np.uint8 else 'F'
    base = np.asarray(Image.fromarray(raw, mode=mode).resize((256, 256), Image.BICUBIC))
    oh, ow, th, tw = prm['crop']
    crop = np.ascontiguousarray(base[oh:oh + th, ow:ow + tw])
    src = np.asarray(Image.fromarray(crop, mode=mode).resize((256, 256), Image.BICUBIC))
    if prm['flip']:
        src = np.flip(src, axis=1)
    img = D.shift_pixel(src, 0, 0)
    img_t = D.gauss_noise(D.shift_pixel(src.copy(), *prm['shift']), prm['noise'])
    return np.ascontiguousarray(img), img_t


def main():
    n, iters = 64, 20
    out = {'batch': n, 'raw': '512x512', 'iters': iters, 'cases': {}}
    pipe = C.CMUNetGpuPipeline()
    for name, dt in (('uint8', np.uint8), ('float32', np.float32)):
        raws = np.stack([D.synthetic_raw(dt, 100 + i % 4) for i in range(n)])
        np.random.seed(5)
        pyrandom.seed(5)
        t0 = time.perf_counter()
        prm = pipe.draw_params(n)
        t_draw = (time.perf_counter() - t0) * 1e3
        dev_raw = torch.from_numpy(raws).cuda()
        pinned = torch.from_numpy(raws).pin_memory()
        for _ in range(3):
            pipe(dev_raw, prm, noise_seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            pipe(dev_raw, prm, noise_seed=i)
        e1.record()
        torch.cuda.synchronize()
        gpu_ms = e0.elapsed_time(e1) / iters
        # the same with the host->device copy of the raw batch inside the timed region
        e0.record()
        for i in range(iters):
            pipe(pinned.cuda(non_blocking=True), prm, noise_seed=i)
        e1.record()
        torch.cuda.synchronize()
        gpu_h2d_ms = e0.elapsed_time(e1) / iters
        # algorithmic bytes: raw read + base write/read + src write/read + two float32 outputs
        el = raws.dtype.itemsize
        alg = n * (512 * 512 * el + 2 * 2 * 256 * 256 * el + 2 * 224 * 224 * 4)
        # host: one core, Pillow
        np.random.seed(5)
        pyrandom.seed(5)
        host_prm = [D.draw_sample_params() for _ in range(16)]
        t0 = time.perf_counter()
        for i in range(16):
            pillow_sample(raws[i], host_prm[i])
        cpu_ms_per_sample = (time.perf_counter() - t0) * 1e3 / 16
        out['cases'][name] = {
            'gpu_ms_per_batch': round(gpu_ms, 4), 'gpu_img_per_s': round(n / gpu_ms * 1e3, 1),
            'gpu_ms_per_batch_with_h2d': round(gpu_h2d_ms, 4), 'host_draw_params_ms': round(t_draw, 3),
            'algorithmic_MB': round(alg / 1e6, 2), 'algorithmic_GBps': round(alg / gpu_ms / 1e6, 1),
            'pillow_1core_ms_per_sample': round(cpu_ms_per_sample, 3),
            'pillow_1core_img_per_s': round(1e3 / cpu_ms_per_sample, 1),
        }
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'data_pipeline_timing.json'), 'w'), indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
