"""CPU: the data-pipeline oracle (oracle/data_oracle.py) against Pillow itself and against the golden
fingerprints minted from the reference's own transform classes (oracle/make_goldens_data.py)."""
import hashlib
import json
import os
import random as pyrandom

import numpy as np
import pytest

from oracle import data_oracle as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'data_pipeline.json')))['cases']


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize('dtype', [np.uint8, np.float32])
def test_pil_resize_restatement_is_bit_exact_with_pillow(dtype):
    Image = pytest.importorskip('PIL.Image')
    rng = np.random.RandomState(0)
    for ih, iw, oh, ow in [(512, 512, 256, 256), (100, 137, 256, 256), (256, 256, 256, 256), (300, 256, 256, 256),
                           (64, 64, 256, 256), (333, 777, 256, 256)]:
        a = (rng.rand(ih, iw) * 255).astype(np.uint8) if dtype == np.uint8 else (rng.randn(ih, iw) * 50 + 100).astype(np.float32)
        ref = np.asarray(Image.fromarray(a).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(ref, D.pil_resize(a, (oh, ow))), (dtype, ih, iw)


@pytest.mark.parametrize('name', sorted(GOLD))
def test_pipeline_matches_reference_golden(name):
    g = GOLD[name]
    raw = D.synthetic_raw(np.dtype(g['dtype']).type, g['raw_seed'])
    assert _sha(raw) == g['raw']['sha256']
    # the parameters: same generators, same draw order as the reference's __getitem__
    np.random.seed(g['seed'])
    pyrandom.seed(g['seed'])
    prm = D.draw_sample_params()
    assert list(prm['crop']) == g['crop'] and prm['flip'] == g['flip'] and list(prm['shift']) == g['shift']
    assert _sha(D.pil_resize(raw, (256, 256))) == g['base256']['sha256']
    img, img_t = D.sample_pipeline(raw, prm)
    assert img.shape == (224, 224) and img.dtype == raw.dtype
    assert _sha(img) == g['img']['sha256'], (float(img.astype(np.float64).sum()), g['img']['sum'])
    assert _sha(img_t) == g['img_t']['sha256'], (float(img_t.astype(np.float64).sum()), g['img_t']['sum'])


def test_fallback_central_crop_and_ragged_sizes():
    # an aspect range that can never fit -> central-crop fallback (processing.py:492-505); tiny and non-square inputs
    np.random.seed(1)
    oh, ow, th, tw = D.rand_crop_params(16, 256, crop_ratio_range=(4.0, 5.0))
    assert (oh, ow, th, tw) == (0, 117, 16, 21)
    a = np.arange(12, dtype=np.uint8).reshape(3, 4)
    assert D.pil_resize(a, (3, 4)).tolist() == a.tolist()
    assert D.pil_resize(a, (6, 9)).shape == (6, 9)
