"""pytest -m gpu: the GPU data pipeline (contrastive_masked_unet_b200/data.py) against the pinned oracle
(oracle/data_oracle.py) and the golden fingerprints minted from the reference's own transforms -- BIT-EXACT."""
import hashlib
import json
import os
import random as pyrandom

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'data_pipeline.json')))['cases']


def _gpu():
    if not torch.cuda.is_available():
        pytest.skip('needs a B200')


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize('name', sorted(GOLD))
def test_pipeline_bit_exact_vs_reference_golden(name):
    _gpu()
    import contrastive_masked_unet_b200 as C
    from oracle import data_oracle as D
    g = GOLD[name]
    dt = np.dtype(g['dtype']).type
    raw = D.synthetic_raw(dt, g['raw_seed'])
    pipe = C.CMUNetGpuPipeline()
    np.random.seed(g['seed'])
    pyrandom.seed(g['seed'])
    prm = pipe.draw_params(1, host_noise=True)          # same generators, same order as the reference __getitem__
    assert list(prm.crop[0]) == g['crop'] and prm.flip[0] == g['flip'] and list(prm.shift[0]) == g['shift']
    img, img_t = pipe(torch.from_numpy(raw)[None].cuda(), prm)
    torch.cuda.synchronize()
    img_n = img[0].cpu().numpy().astype(dt)
    img_t_n = img_t[0].cpu().numpy().astype(dt)
    assert _sha(img_n) == g['img']['sha256'], (float(img_n.astype(np.float64).sum()), g['img']['sum'])
    assert _sha(img_t_n) == g['img_t']['sha256'], (float(img_t_n.astype(np.float64).sum()), g['img_t']['sum'])


@pytest.mark.parametrize('dtype', [np.uint8, np.float32])
def test_batch_vs_oracle_ragged_sizes_and_boxes(dtype):
    """a batch with different crop boxes / flips / shifts, a non-square raw size, up- and down-scaling in one call"""
    _gpu()
    import contrastive_masked_unet_b200 as C
    from oracle import data_oracle as D
    rng = np.random.RandomState(3)
    n, h0, w0 = 5, 300, 417
    raw = (rng.rand(n, h0, w0) * 255).astype(np.uint8) if dtype == np.uint8 else (rng.randn(n, h0, w0) * 40 + 90).astype(np.float32)
    pipe = C.CMUNetGpuPipeline()
    np.random.seed(11)
    pyrandom.seed(11)
    prm = pipe.draw_params(n, host_noise=True)
    prm.crop[0] = (0, 0, 256, 256)                      # identity crop (no resampling in either direction)
    prm.crop[1] = (10, 200, 240, 56)                    # narrow box: strong horizontal up-scaling
    img, img_t = pipe(torch.from_numpy(raw).cuda(), prm)
    torch.cuda.synchronize()
    for i in range(n):
        p = {'crop': prm.crop[i], 'flip': prm.flip[i], 'shift': prm.shift[i], 'noise': prm.noise[i]}
        o_img, o_img_t = D.sample_pipeline(raw[i], p)
        assert np.array_equal(img[i].cpu().numpy().astype(dtype), o_img), (i, p['crop'])
        assert np.array_equal(img_t[i].cpu().numpy().astype(dtype), o_img_t), (i, p['crop'], p['shift'])


def test_resize_kernel_full_size_batch_matches_oracle_on_probes():
    """B = 64 raw 512 x 512 images (one training batch): the first resize against the oracle on a few planes"""
    _gpu()
    from contrastive_masked_unet_b200 import ops
    from contrastive_masked_unet_b200._lib import lib
    from oracle import data_oracle as D
    rng = np.random.RandomState(5)
    raw = (rng.rand(64, 512, 512) * 255).astype(np.uint8)
    d = torch.from_numpy(raw).cuda()
    tmp = torch.empty(64, 512, 256, dtype=torch.uint8, device='cuda')
    out = torch.empty(64, 256, 256, dtype=torch.uint8, device='cuda')
    lib.cmu_pil_resize_bicubic(d.data_ptr(), 0, 64, 512, 512, 0, tmp.data_ptr(), out.data_ptr(), 256, 256, ops._stream())
    torch.cuda.synchronize()
    for i in (0, 31, 63):
        assert np.array_equal(out[i].cpu().numpy(), D.pil_resize(raw[i], (256, 256)))


def test_device_noise_is_standard_normal_and_feeds_the_model():
    """production mode (Philox noise on the device): statistics of the field, determinism per seed, and one pretraining
    step of the drop-in model consuming the pipeline's output"""
    _gpu()
    import contrastive_masked_unet_b200 as C
    from oracle import data_oracle as D
    raw = np.stack([D.synthetic_raw(np.float32, 200 + i) for i in range(8)])
    pipe = C.CMUNetGpuPipeline()
    np.random.seed(2)
    pyrandom.seed(2)
    prm = pipe.draw_params(8)
    d = torch.from_numpy(raw).cuda()
    img, img_t = pipe(d, prm, noise_seed=1234)
    img2, img_t2 = pipe(d, prm, noise_seed=1234)
    _, img_t3 = pipe(d, prm, noise_seed=99)
    assert torch.equal(img_t, img_t2) and torch.equal(img, img2) and not torch.equal(img_t, img_t3)
    # recover z = (img_t - shifted crop) / sigma for a sample whose shift is known
    zs = []
    for i in range(8):
        src = D.pil_resize(D.pil_resize(raw[i], (256, 256))[prm.crop[i][0]:prm.crop[i][0] + prm.crop[i][2],
                                                          prm.crop[i][1]:prm.crop[i][1] + prm.crop[i][3]].copy(), (256, 256))
        if prm.flip[i]:
            src = np.flip(src, axis=1)
        ph, pw = prm.shift[i]
        c = src[ph:ph + 224, pw:pw + 224].astype(np.float64)
        zs.append((img_t[i].cpu().numpy().astype(np.float64) - c) / (float(c.max()) / 10))
    z = np.stack(zs)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01, (z.mean(), z.std())
    torch.manual_seed(0)
    m = C.build(C.cmunet_config(224)).cuda().train()
    m.init_weights()
    out = m(img, mode='loss', img_t=img_t)
    (out['loss_ct'] + out['loss_rc']).backward()
    assert torch.isfinite(out['loss_ct']) and torch.isfinite(out['loss_rc'])
