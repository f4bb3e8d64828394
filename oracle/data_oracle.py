"""ORACLE (test infrastructure; never imported by the product package): CPU restatement of the reference's per-sample
data pipeline for CM-UNet pretraining.

Restates (paths relative to /root/reference/Pretraining/CM-UNet):
  * cmae/datasets/cmunet_dataset.py:74-88  __getitem__: np.load -> PIL resize (256,256) BICUBIC -> RandomResizedCrop ->
    RandomFlip -> {ShiftPixel(0) | ShiftPixel(31) + GaussNoise} -> (img, img_t)
  * cmae/datasets/pipelines/processing.py:464-505 (rand_crop_params), :570-590 (crop + PIL bicubic resize to 256x256),
    :97-121 (ShiftPixel: python `random.randint` twice, crop [ph:ph+224, pw:pw+224])
  * cmae/datasets/pipelines/auto_augment.py:1137-1155 (GaussNoise.transform: sigma = max(img)/10,
    out = img + sigma * np.random.randn(*img.shape), cast back to img.dtype; `prob` is never consulted, the noise is
    ALWAYS applied)
and the third-party algorithm the resize calls into, which is not part of the reference tree:
  * Pillow (pinned: pillow via torchvision in CMU/environment.yml; container has 12.2.0) `Image.resize(size, BICUBIC)` =
    libImaging/Resample.c: `precompute_coeffs` (support = 2 * max(scale, 1), windows [xmin, xmax), weights normalised to
    sum 1), horizontal pass then vertical pass; mode "L" (uint8): 22-bit fixed-point coefficients, accumulate in int32
    from 1 << 21, `>> 22`, clip to [0, 255]; mode "F" (float32): double accumulation, stored as float32 after each pass.
  * mmcv 2.2.0 `RandomFlip(prob=0.5)` [mmcv, from memory: un-vendored]: one `np.random.choice(['horizontal', None],
    p=[0.5, 0.5])`; `mmcv.imflip(img, 'horizontal')` = np.flip(img, axis=1); `mmcv.imcrop` = array slicing.

Parity status: PINNED.  tests/test_oracle_data.py holds `pil_resize` bit-exactly to Pillow itself (uint8 and float32, up-
and down-scaling, crops) and holds the pipeline to tests/golden/data_pipeline.json, minted by oracle/make_goldens_data.py
from the reference's own transform classes (ShiftPixel, GaussNoise, RandomResizedCrop) loaded through the shim."""
import math
import random as pyrandom

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x, a=-0.5):
    """Resample.c bicubic_filter."""
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, in0, in1, out_size):
    """Resample.c precompute_coeffs for the bicubic filter (support 2.0).  -> (bounds [(xmin, n)], weights [list of float])."""
    scale = filterscale = (in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    bounds, kk = [], []
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        ww = 0.0
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = []
        for x in range(xmax):
            w = bicubic_filter((x + xmin - center + 0.5) * ss)
            k.append(w)
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        bounds.append((xmin, xmax))
        kk.append(k)
    return bounds, kk


def _pass_u8(img, bounds, kk, axis):
    """one separable pass on a uint8 plane (normalize_coeffs_8bpc + ImagingResampleHorizontal/Vertical_8bpc)."""
    h, w = img.shape
    src = img.astype(np.int64)
    n_out = len(bounds)
    out = np.empty((h, n_out) if axis == 1 else (n_out, w), dtype=np.uint8)
    for o, ((lo, n), k) in enumerate(zip(bounds, kk)):
        ki = [int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS)) for v in k]
        acc = np.full(h if axis == 1 else w, 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for t in range(n):
            acc = acc + (src[:, lo + t] if axis == 1 else src[lo + t, :]) * ki[t]
        # int32 wrap-around cannot occur (|sum ki| ~ 2^22, pixels < 2^8); clip8 = clamp(acc >> 22, 0, 255)
        v = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
        if axis == 1:
            out[:, o] = v
        else:
            out[o, :] = v
    return out


def _pass_f32(img, bounds, kk, axis):
    """one separable pass on a float32 plane (ImagingResample*_32bpc: double accumulation in tap order, float32 store)."""
    h, w = img.shape
    src = img.astype(np.float64)
    n_out = len(bounds)
    out = np.empty((h, n_out) if axis == 1 else (n_out, w), dtype=np.float32)
    for o, ((lo, n), k) in enumerate(zip(bounds, kk)):
        acc = np.zeros(h if axis == 1 else w, dtype=np.float64)
        for t in range(n):
            acc = acc + (src[:, lo + t] if axis == 1 else src[lo + t, :]) * k[t]
        if axis == 1:
            out[:, o] = acc.astype(np.float32)
        else:
            out[o, :] = acc.astype(np.float32)
    return out


def pil_resize(img, out_hw, box=None):
    """`Image.fromarray(img).resize((ow, oh), Image.BICUBIC, box)` for a 2-D uint8 ("L") or float32 ("F") array.
    box = (x0, y0, x1, y1) in source pixels (default: the whole image)."""
    assert img.ndim == 2 and img.dtype in (np.uint8, np.float32)
    oh, ow = out_hw
    ih, iw = img.shape
    x0, y0, x1, y1 = box if box is not None else (0, 0, iw, ih)
    need_h = ow != iw or x0 != 0 or x1 != iw
    need_v = oh != ih or y0 != 0 or y1 != ih
    bh, kh = precompute_coeffs(iw, x0, x1, ow)
    bv, kv = precompute_coeffs(ih, y0, y1, oh)
    pas = _pass_u8 if img.dtype == np.uint8 else _pass_f32
    cur = img
    if need_h:
        # Pillow only resamples the source rows the vertical pass will read; shift the vertical bounds accordingly
        first = bv[0][0]
        last = bv[-1][0] + bv[-1][1]
        cur = pas(cur[first:last] if need_v else cur, bh, kh, axis=1)
        if need_v:
            bv = [(lo - first, n) for lo, n in bv]
    if need_v:
        cur = pas(cur, bv, kv, axis=0)
    if not need_h and not need_v:
        cur = img.copy()
    return cur


def rand_crop_params(h, w, crop_ratio_range=(0.2, 1.0), aspect_ratio_range=(3. / 4., 4. / 3.), max_attempts=10):
    """processing.py:464-505 on numpy's GLOBAL legacy RNG (same draw order)."""
    area = h * w
    for _ in range(max_attempts):
        target_area = np.random.uniform(*crop_ratio_range) * area
        log_ratio = (math.log(aspect_ratio_range[0]), math.log(aspect_ratio_range[1]))
        aspect_ratio = math.exp(np.random.uniform(*log_ratio))
        target_w = int(round(math.sqrt(target_area * aspect_ratio)))
        target_h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < target_w <= w and 0 < target_h <= h:
            offset_h = np.random.randint(0, h - target_h + 1)
            offset_w = np.random.randint(0, w - target_w + 1)
            return offset_h, offset_w, target_h, target_w
    in_ratio = float(w) / float(h)
    if in_ratio < min(aspect_ratio_range):
        target_w = w
        target_h = int(round(target_w / min(aspect_ratio_range)))
    elif in_ratio > max(aspect_ratio_range):
        target_h = h
        target_w = int(round(target_h * max(aspect_ratio_range)))
    else:
        target_w, target_h = w, h
    return (h - target_h) // 2, (w - target_w) // 2, target_h, target_w


def draw_sample_params(base=256, pixel=31):
    """The random parameters of ONE `__getitem__` call, drawn in the reference's order from the same generators
    (numpy global legacy RNG: crop box, flip; python `random`: the two ShiftPixel calls).  The GaussNoise field itself
    (np.random.randn(224, 224), drawn last) is returned as `noise` so that parity is defined."""
    oh, ow, th, tw = rand_crop_params(base, base)
    flip = np.random.choice(['horizontal', None], p=[0.5, 0.5]) == 'horizontal'   # [mmcv RandomFlip, from memory]
    pyrandom.randint(0, 0)
    pyrandom.randint(0, 0)                                                        # self.shift (pixel=0), processing.py:111-112
    ph = pyrandom.randint(0, pixel)
    pw = pyrandom.randint(0, pixel)
    noise = np.random.randn(224, 224)
    return {'crop': (oh, ow, th, tw), 'flip': bool(flip), 'shift': (ph, pw), 'noise': noise}


def shift_pixel(img, ph, pw):
    """processing.py:109-121 (the 224 literal)."""
    assert ph + 224 < img.shape[0] and pw + 224 < img.shape[0]
    return img[ph:ph + 224, pw:pw + 224]


def gauss_noise(img, noise):
    """auto_augment.py:1148-1154 with the randn field made explicit."""
    sigma = np.max(img) / 10
    out = img + sigma * noise
    return np.array(out, dtype=img.dtype)


def sample_pipeline(raw, params):
    """cmunet_dataset.py:74-88 for one raw 2-D array with explicit random parameters -> (img, img_t), dtype of `raw`."""
    base = pil_resize(raw, (256, 256))
    oh, ow, th, tw = params['crop']
    crop = np.ascontiguousarray(base[oh:oh + th, ow:ow + tw])                     # mmcv.imcrop, inclusive bbox -> slicing
    src = pil_resize(crop, (256, 256))
    if params['flip']:
        src = np.flip(src, axis=1)
    img = shift_pixel(src, 0, 0)
    img_t = gauss_noise(shift_pixel(src.copy(), *params['shift']), params['noise'])
    return np.ascontiguousarray(img), img_t


def synthetic_raw(dtype, seed):
    """Deterministic synthetic "angiogram" (smooth vessel-like structure + noise), 512 x 512, uint8 or float32: the raw
    input of the golden cases (legacy RandomState: stable across numpy versions)."""
    r = np.random.RandomState(seed)
    y, x = np.mgrid[0:512, 0:512].astype(np.float64)
    img = 120 + 60 * np.sin(x / 37.0 + 2 * np.sin(y / 53.0)) * np.cos(y / 29.0) + 12 * r.randn(512, 512)
    img -= 80 * np.exp(-((x - 0.6 * y - 60) ** 2) / 90.0)
    if dtype == np.uint8:
        return np.clip(img, 0, 255).astype(np.uint8)
    return (img / 255.0).astype(np.float32)
