"""GPU data pipeline for CM-UNet pretraining: the per-sample work of the reference's
`CMUNetDataset.__getitem__` (Pretraining/CM-UNet/cmae/datasets/cmunet_dataset.py:60-88 with the train pipeline of
configs/cmunet_config.py:48-53) for a whole batch on the device.

At ~400 img/s per GPU the reference's 8 PIL/cv2 DataLoader workers cannot feed the step; here the host only draws the
random parameters -- from the same generators and in the same order as the reference, so that results are defined and
comparable -- and three kernels do the pixel work (`cmu_pil_resize_bicubic` x2, `cmu_aug_shift_flip_noise`):

    raw (N,H0,W0) uint8|float32 --Pillow bicubic--> 256x256 --RandomResizedCrop box + Pillow bicubic--> 256x256
        --RandomFlip--> {ShiftPixel(0) -> img (N,224,224)} , {ShiftPixel(31) + GaussNoise -> img_t (N,224,224)}

Outputs are float32 (what the model consumes); their VALUES are those of the reference's arrays in the raw dtype
(uint8: Pillow's fixed-point resampling and numpy's wrapping cast after the noise), bit for bit when the noise field is
passed explicitly.  CUDA only, no fallback."""
import math
import random as pyrandom

import numpy as np
import torch

from . import ops
from ._lib import CmuError, lib


class SampleParams:
    """Random parameters of a batch: crop boxes (oh, ow, th, tw), flip flags, ShiftPixel offsets (ph, pw) and, when
    drawn on the host, the GaussNoise fields."""
    __slots__ = ('crop', 'flip', 'shift', 'noise')

    def __init__(self, crop, flip, shift, noise=None):
        self.crop, self.flip, self.shift, self.noise = crop, flip, shift, noise

    def __len__(self):
        return len(self.flip)


def _rand_crop_params(h, w, crop_ratio_range, aspect_ratio_range, max_attempts):
    """RandomResizedCrop.rand_crop_params (cmae/datasets/pipelines/processing.py:464-505): numpy GLOBAL legacy RNG."""
    area = h * w
    for _ in range(max_attempts):
        target_area = np.random.uniform(*crop_ratio_range) * area
        log_ratio = (math.log(aspect_ratio_range[0]), math.log(aspect_ratio_range[1]))
        aspect_ratio = math.exp(np.random.uniform(*log_ratio))
        tw = int(round(math.sqrt(target_area * aspect_ratio)))
        th = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < tw <= w and 0 < th <= h:
            return np.random.randint(0, h - th + 1), np.random.randint(0, w - tw + 1), th, tw
    in_ratio = float(w) / float(h)
    if in_ratio < min(aspect_ratio_range):
        tw, th = w, int(round(w / min(aspect_ratio_range)))
    elif in_ratio > max(aspect_ratio_range):
        th, tw = h, int(round(h * max(aspect_ratio_range)))
    else:
        tw, th = w, h
    return (h - th) // 2, (w - tw) // 2, th, tw


class CMUNetGpuPipeline:
    """Batch version of `CMUNetDataset.__getitem__`.  `base`, `out` and `pixel` are the reference's literals
    (256: cmunet_dataset.py:78 / cmunet_config.py:49; 224: processing.py:117; 31: cmunet_config.py:66)."""

    def __init__(self, base=256, out=224, pixel=31, crop_ratio_range=(0.2, 1.0), aspect_ratio_range=(3. / 4., 4. / 3.),
                 max_attempts=10, flip_prob=0.5):
        if out + pixel >= base:
            raise ValueError('ShiftPixel asserts pixel + size < width (processing.py:113-114)')
        self.base, self.out, self.pixel = base, out, pixel
        self.crop_ratio_range, self.aspect_ratio_range = crop_ratio_range, aspect_ratio_range
        self.max_attempts, self.flip_prob = max_attempts, flip_prob

    def draw_params(self, n, host_noise=False):
        """Per sample, in the reference's order: crop box and flip from numpy's global RNG, the two ShiftPixel calls
        from python's `random` (the pixel=0 call consumes two draws as well), then -- host_noise=True -- the
        np.random.randn(out, out) field of GaussNoise.  host_noise=False leaves the field to the device (Philox)."""
        crop, flip, shift, noise = [], [], [], []
        for _ in range(n):
            crop.append(_rand_crop_params(self.base, self.base, self.crop_ratio_range, self.aspect_ratio_range,
                                          self.max_attempts))
            # mmcv RandomFlip(prob=0.5): one np.random.choice over ['horizontal', None]  [mmcv 2.2.0, un-vendored]
            flip.append(bool(np.random.choice(['horizontal', None], p=[self.flip_prob, 1 - self.flip_prob]) == 'horizontal'))
            pyrandom.randint(0, 0)
            pyrandom.randint(0, 0)
            shift.append((pyrandom.randint(0, self.pixel), pyrandom.randint(0, self.pixel)))
            if host_noise:
                noise.append(np.random.randn(self.out, self.out))
        return SampleParams(crop, flip, shift, np.stack(noise) if host_noise else None)

    def __call__(self, raw, params, noise_seed=0):
        """raw: (N,H0,W0) uint8 or float32 CUDA tensor -> (img, img_t), float32 (N,out,out)."""
        ops._need_cuda(raw)
        if raw.dim() != 3 or raw.dtype not in (torch.uint8, torch.float32):
            raise CmuError('CMUNetGpuPipeline: raw must be a (N,H,W) uint8 or float32 tensor')
        if len(params) != raw.shape[0]:
            raise CmuError('CMUNetGpuPipeline: one parameter set per image')
        raw = raw.contiguous()
        n, h0, w0 = raw.shape
        dev, st = raw.device, ops._stream()
        dt = 0 if raw.dtype == torch.uint8 else 1
        b = self.base
        base = torch.empty(n, b, b, dtype=raw.dtype, device=dev)
        tmp = torch.empty(n, max(h0, b), b, dtype=raw.dtype, device=dev)
        lib.cmu_pil_resize_bicubic(raw.data_ptr(), dt, n, h0, w0, 0, tmp.data_ptr(), base.data_ptr(), b, b, st)
        boxes = torch.tensor([[ow, oh, tw, th] for (oh, ow, th, tw) in params.crop], dtype=torch.int32).to(dev, non_blocking=True)
        src = torch.empty(n, b, b, dtype=raw.dtype, device=dev)
        lib.cmu_pil_resize_bicubic(base.data_ptr(), dt, n, b, b, boxes.data_ptr(), tmp.data_ptr(), src.data_ptr(), b, b, st)
        prm = torch.tensor([[int(f), ph, pw, 0] for f, (ph, pw) in zip(params.flip, params.shift)],
                           dtype=torch.int32).to(dev, non_blocking=True)
        noise = None
        if params.noise is not None:
            noise = torch.from_numpy(np.ascontiguousarray(params.noise, dtype=np.float64)).to(dev)
        img = torch.empty(n, self.out, self.out, dtype=torch.float32, device=dev)
        img_t = torch.empty_like(img)
        lib.cmu_aug_shift_flip_noise(src.data_ptr(), dt, n, b, b, prm.data_ptr(), ops._ptr(noise), int(noise_seed) & (2 ** 64 - 1),
                                     self.out, img.data_ptr(), img_t.data_ptr(), st)
        return img, img_t
