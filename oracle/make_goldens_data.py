"""TEST INFRASTRUCTURE: mints tests/golden/data_pipeline.json (sha256 + probes of every stage) by running the reference's OWN data-pipeline code in the
build container (read-only /root/reference, mmcv/mmengine/timm stand-ins under oracle/ref_shim):

  * `Image.fromarray(raw).resize((256, 256), resample=Image.BICUBIC)`      cmae/datasets/cmunet_dataset.py:77-78
  * `RandomResizedCrop(scale=256, crop_ratio_range=(0.2, 1.0), backend='pillow', interpolation='bicubic')`
    with its `rand_crop_params` draws from numpy's global RNG              cmae/datasets/pipelines/processing.py:399-590
  * horizontal flip (mmcv RandomFlip is un-vendored: np.flip on the drawn decision)
  * `ShiftPixel(pixel=0)` / `ShiftPixel(pixel=31)` (python `random`)        processing.py:97-121
  * `GaussNoise(magnitude_range=(0.1, 2.0), magnitude_std='inf', prob=0.5)` auto_augment.py:1137-1155
for two synthetic raw images (uint8 and float32, 512x512) under fixed seeds.  The drawn parameters are captured by
replaying the same seeds through oracle/data_oracle.draw_sample_params (same generators, same order) and stored next
to the outputs, so that the oracle and the CUDA kernels can be held to the stored arrays with explicit parameters.

    python oracle/make_goldens_data.py"""
import importlib
import os
import random as pyrandom
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import ref_loader as RL  # noqa: E402


def load_reference_transforms():
    RL._ensure_paths()
    import cmae  # noqa: F401
    proc = importlib.import_module('cmae.datasets.pipelines.processing')
    aug = importlib.import_module('cmae.datasets.pipelines.auto_augment')
    return proc, aug


def reference_sample(proc, aug, raw, seed):
    from PIL import Image
    np.random.seed(seed)
    pyrandom.seed(seed)
    image = Image.fromarray(raw).resize((256, 256), resample=Image.BICUBIC)
    results = {'img': np.asarray(image)}
    rrc = proc.RandomResizedCrop(scale=256, crop_ratio_range=(0.2, 1.0), backend='pillow', interpolation='bicubic')
    src = rrc(results)
    flip = np.random.choice(['horizontal', None], p=[0.5, 0.5]) == 'horizontal'    # mmcv RandomFlip [from memory]
    if flip:
        src['img'] = np.flip(src['img'], axis=1)
    patch_results = {'img': src['img']}
    img_t_results = {'img': src['img'].copy()}
    shift0 = proc.ShiftPixel(pixel=0)
    shift31 = proc.ShiftPixel(pixel=31)
    noise = aug.GaussNoise(magnitude_range=(0.1, 2.0), magnitude_std='inf', prob=0.5)
    img = shift0(patch_results)['img']
    img_t = noise(shift31(img_t_results))['img']
    return np.ascontiguousarray(img), np.ascontiguousarray(img_t), np.asarray(image)


def fingerprint(a):
    import hashlib
    a = np.ascontiguousarray(a)
    return {'sha256': hashlib.sha256(a.tobytes()).hexdigest(), 'dtype': str(a.dtype), 'shape': list(a.shape),
            'sum': float(a.astype(np.float64).sum()), 'corner': [float(v) for v in a[:2, :3].reshape(-1)],
            'probe': [float(a[i, j]) for i, j in ((17, 201), (100, 100), (223, 0), (150, 37))]}


def main():
    import json
    from oracle import data_oracle as D
    synthetic_raw = D.synthetic_raw
    proc, aug = load_reference_transforms()
    out = {}
    for name, dtype, seed in (('u8', np.uint8, 5), ('f32', np.float32, 6), ('u8b', np.uint8, 7)):
        raw = synthetic_raw(dtype, 100 + seed)
        img, img_t, base = reference_sample(proc, aug, raw, seed)
        # the parameters the reference drew: replay the same seeds through the restated draw order
        np.random.seed(seed)
        pyrandom.seed(seed)
        prm = D.draw_sample_params()
        out[name] = {'seed': seed, 'raw_seed': 100 + seed, 'dtype': np.dtype(dtype).name, 'crop': [int(v) for v in prm['crop']],
                     'flip': bool(prm['flip']), 'shift': [int(v) for v in prm['shift']], 'raw': fingerprint(raw),
                     'base256': fingerprint(base), 'img': fingerprint(img), 'img_t': fingerprint(img_t)}
        o_img, o_img_t = D.sample_pipeline(raw, prm)
        print(name, 'crop', prm['crop'], 'flip', prm['flip'], 'shift', prm['shift'], '| oracle == reference:',
              np.array_equal(o_img, img), np.array_equal(o_img_t, img_t))
    path = os.path.join(ROOT, 'tests', 'golden', 'data_pipeline.json')
    import PIL
    json.dump({'meta': {'numpy': np.__version__, 'pillow': PIL.__version__, 'generated_by': 'oracle/make_goldens_data.py'},
               'cases': out}, open(path, 'w'), indent=1)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
