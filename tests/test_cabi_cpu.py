"""CPU-only checks of the boundary: the C-ABI library builds/loads and exports every symbol include/cmu_b200.h declares
(no compute calls), the drop-in modules mirror the reference's state_dict / init / error behaviour, and the product path
fails loudly without CUDA (no CPU fallback)."""
import io
import json
import os
import subprocess
import sys

import pytest
import torch

import contrastive_masked_unet_b200 as C
from contrastive_masked_unet_b200._lib import LIB_PATH, parse_header
from oracle import cmunet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB_PATH):
        from contrastive_masked_unet_b200._buildlib import build
        build()
    import ctypes
    dll = ctypes.CDLL(LIB_PATH)
    protos = parse_header()
    assert len(protos) >= 45
    for name in protos:
        assert hasattr(dll, name), f'{name} declared in include/cmu_b200.h but not exported'
    out = subprocess.run(['nm', '-D', '--defined-only', LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if ' T ' in l and 'cmu_' in l}
    assert set(protos) <= exported
    assert C.lib.cmu_version() >= 100
    assert C.lib.cmu_mask_state_words() == 625


def test_kernels_are_blackwell_native():
    """SASS of the shipped library contains tcgen05 MMA, TMEM loads and TMA loads/stores (no legacy HMMA)."""
    out = subprocess.run(['cuobjdump', '-sass', LIB_PATH], capture_output=True, text=True).stdout
    assert 'UTCHMMA' in out and 'LDTM' in out and 'UTMALDG' in out and 'UTMASTG' in out
    assert 'HMMA.16816' not in out


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(C.CmuError):
        C.lib.cmu_device_check()
    m = C.UNet()
    with pytest.raises(C.CmuError):
        m(torch.rand(1, 32, 32))


def test_state_dict_keys_and_init_match_reference():
    g = json.load(open(os.path.join(GOLD, 'pretrain.json')))['cases'][0]
    torch.manual_seed(g['seed'])
    m = C.build(C.cmunet_config(g['S']))
    m.init_weights()
    assert [k for k, _ in m.named_parameters()] == g['param_keys']
    assert sum(p.numel() for p in m.parameters()) == g['n_params']
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == g['n_trainable']
    from tests.test_oracle_pinned import check_fp
    for k, p in m.named_parameters():
        check_fp(p, g['init'][k], rtol=1e-6)
    torch.manual_seed(g['seed'])
    o = O.OracleCMUNet(img_size=g['S'], np_seed=g['seed'])
    o.init_weights()
    assert list(m.state_dict().keys()) == list(o.state_dict().keys())
    # state dicts are interchangeable with the oracle/reference layout
    m.load_state_dict(o.state_dict(), strict=True)
    assert all(not p.requires_grad for p in m.target_backbone.parameters())
    assert all(not p.requires_grad for p in m.target_projector.parameters())


def test_finetune_unet_keys_pickle_and_loss_names():
    torch.manual_seed(0)
    u = C.UNet()
    torch.manual_seed(0)
    ou = O.OracleUNet()
    assert list(u.state_dict().keys()) == list(ou.state_dict().keys())
    assert all(torch.equal(a, b) for a, b in zip(u.state_dict().values(), ou.state_dict().values()))
    buf = io.BytesIO()
    torch.save(u, buf)                      # FT/train.py:212 pickles the whole module
    buf.seek(0)
    u2 = torch.load(buf, weights_only=False)
    assert list(u2.state_dict().keys()) == list(u.state_dict().keys())
    loss = C.DiceLoss(activation='softmax', threshold=0.5, ignore_channels=[0]) + C.CrossEntropyLoss()
    assert loss.__name__ == 'dice_loss + cross_entropy_loss'
    assert C.IoU().__name__ == 'iou_loss'
    assert (2 * C.CrossEntropyLoss()).__name__ == '2 * cross_entropy_loss'
    with pytest.raises(ValueError):
        C.DiceLoss() + 3


def test_mode_dispatch_and_upblock_errors():
    m = C.build(C.cmunet_config(64))
    with pytest.raises(RuntimeError, match='Invalid mode'):
        m(torch.zeros(1, 64, 64), mode='predict')
    with pytest.raises(ValueError):
        C.UpBlock(8, 4, 'nearest')
    assert m.momentum == m.base_momentum == 0.996


def test_mask_stream_numpy_handover():
    import numpy as np
    np.random.seed(60)
    ms = C.MaskStream()
    ms.set_numpy_state(np.random.get_state(), device='cpu')
    st = ms.get_numpy_state()
    assert st[2] == 624 and (st[1] == np.random.get_state()[1]).all()


def test_checkpoint_handoff_pretrain_to_finetune(tmp_path):
    """Finetuning/train.py:262-273: a pretraining checkpoint in mmengine layout loads into the fine-tune UNet."""
    from contrastive_masked_unet_b200.checkpoint import load_pretrained_into_unet, save_pretrain_checkpoint
    torch.manual_seed(3)
    m = C.build(C.cmunet_config(64))
    m.init_weights()
    path = str(tmp_path / 'epoch_1.pth')
    ck = save_pretrain_checkpoint(m, path)
    assert 'mmengine_version' in ck['meta'] and any(k.startswith('backbone.') for k in ck['state_dict'])
    torch.manual_seed(4)
    u = C.UNet()
    last_w = u.conv_last.weight.detach().clone()
    res = load_pretrained_into_unet(u, path)
    assert set(res.missing_keys) == {'conv_last.weight', 'conv_last.bias'}          # dropped on purpose (:271-272)
    assert torch.equal(u.conv_last.weight, last_w)
    assert torch.equal(u.down_conv3.double_conv.double_conv[0].weight, m.backbone.down_conv3.double_conv.double_conv[0].weight)
    assert torch.equal(u.up_conv2.up_sample.weight, m.pixel_decoder.up_conv2.up_sample.weight)
    assert torch.equal(u.double_conv.double_conv[4].running_var, m.backbone.double_conv.double_conv[4].running_var)
    # and the same file loads in the oracle (== reference layout) UNet
    ou = O.OracleUNet()
    load_pretrained_into_unet(ou, path)
    assert torch.equal(ou.up_conv4.double_conv.double_conv[3].weight, m.pixel_decoder.up_conv4.double_conv.double_conv[3].weight)


def test_host_planners_need_no_gpu():
    """Workspace / grid planners are pure host code: callable without a device, positive and monotonic."""
    L = C.lib
    w1 = L.cmu_conv3x3_wgrad_workspace_bytes(64, 64, 64, 512, 512)
    w2 = L.cmu_conv3x3_wgrad_workspace_bytes(512, 512, 64, 64, 64)
    assert w1 >= 9 * 64 * 64 * 4 and w2 >= 9 * 512 * 512 * 4
    assert L.cmu_convT2x2_wgrad_workspace_bytes(128, 64, 64, 256, 256) >= 4 * 128 * 64 * 4
    assert L.cmu_sgemm_workspace_bytes(64, 256, 1536) % (64 * 256 * 4) == 0
    assert L.cmu_sgemm_workspace_bytes(64, 256, 64) == 64 * 256 * 4                 # small K: no split
    assert L.cmu_gemm_tn_workspace_bytes(64, 1536, 262144) >= 64 * 1536 * 4
    assert L.cmu_soft_cldice_workspace_bytes(4, 256, 256) == (4 * 2 * 4 * 256 * 256 + 8) * 8
    assert L.cmu_bn_bwd_grid() > 0 and L.cmu_conv_max_grid() > 0 and L.cmu_conv3x3_c1_grid() > 0


def test_moco_constructor_matches_reference_rng_order_and_keys():
    """Drop-in Moco_v2: same default-init weights and queue as the reference constructor under the same seed
    (tests/golden/moco.json is minted from the unmodified reference), same buffers, encoder_k frozen; CPU input raises."""
    g = json.load(open(os.path.join(GOLD, 'moco.json')))['case']
    torch.manual_seed(g['seed'])
    m = C.Moco_v2(emb_dim=1024, num_negatives=g['K'])
    assert [k for k, _ in m.encoder_q.named_parameters()] == list(g['init_q'].keys())
    for k, p in m.encoder_q.named_parameters():
        assert float(p.detach().double().norm()) == pytest.approx(g['init_q'][k]['norm'], rel=1e-6)
    assert float(m.queue.double().sum()) == pytest.approx(g['init_queue']['sum'], rel=1e-6, abs=1e-6)
    assert set(dict(m.named_buffers())) >= {'queue', 'queue_ptr', 'val_queue', 'val_queue_ptr'}
    assert tuple(m.queue.shape) == (1024, g['K']) and int(m.queue_ptr) == 0
    assert all(not p.requires_grad for p in m.encoder_k.parameters())
    assert all(p.requires_grad for p in m.encoder_q.parameters())
    if not torch.cuda.is_available():
        with pytest.raises(C.CmuError):
            m.training_step(torch.rand(64, 32, 32), torch.rand(64, 32, 32))
    with pytest.raises(NotImplementedError):
        C.Moco_v2(use_mlp=True, num_negatives=64)


def test_data_pipeline_host_draws_match_the_oracle_order():
    """CMUNetGpuPipeline.draw_params consumes numpy's global RNG and python's `random` exactly like the restated
    reference order (oracle/data_oracle.draw_sample_params, pinned to the reference by tests/test_oracle_data.py)."""
    import random as pyrandom
    import numpy as np
    from oracle import data_oracle as D
    pipe = C.CMUNetGpuPipeline()
    np.random.seed(21)
    pyrandom.seed(21)
    prm = pipe.draw_params(6, host_noise=True)
    np.random.seed(21)
    pyrandom.seed(21)
    for i in range(6):
        o = D.draw_sample_params()
        assert tuple(prm.crop[i]) == tuple(o['crop']) and prm.flip[i] == o['flip'] and tuple(prm.shift[i]) == tuple(o['shift'])
        assert np.array_equal(prm.noise[i], o['noise'])
    # both generators are left at the same position as after six reference-order draws
    nxt, nxt_py = np.random.rand(), pyrandom.random()
    np.random.seed(21)
    pyrandom.seed(21)
    for _ in range(6):
        D.draw_sample_params()
    assert np.random.rand() == nxt and pyrandom.random() == nxt_py
    with pytest.raises(ValueError):
        C.CMUNetGpuPipeline(base=256, out=240, pixel=31)
    if not torch.cuda.is_available():
        with pytest.raises(C.CmuError):
            pipe(torch.zeros(1, 64, 64, dtype=torch.uint8), pipe.draw_params(1))


def test_checkpoint_handoff_moco_to_finetune(tmp_path):
    """Finetuning/train.py:286-296 (".ckpt" branch): a MoCo-v2 checkpoint hands its QUERY encoder to the fine-tune UNet
    (`encoder_q.` stripped, `encoder_k.*` ignored); a checkpoint with no usable encoder tensor raises instead of silently
    fine-tuning from random weights."""
    from contrastive_masked_unet_b200.checkpoint import (finetune_state_dict, load_pretrained_into_unet,
                                                         save_moco_checkpoint)
    torch.manual_seed(9)
    m = C.Moco_v2(emb_dim=1024, num_negatives=128)
    with torch.no_grad():                       # make the two encoders differ so that a mix-up would show
        for p in m.encoder_k.parameters():
            p.add_(1.0)
    path = str(tmp_path / 'moco_epoch_3.ckpt')
    ck = save_moco_checkpoint(m, path, epoch=3)
    assert any(k.startswith('encoder_q.') for k in ck['state_dict']) and 'queue' in ck['state_dict']
    mapped = finetune_state_dict(ck, path)
    assert all(not k.startswith('encoder_') for k in mapped) and 'down_conv1.double_conv.double_conv.0.weight' in mapped
    torch.manual_seed(10)
    u = C.UNet()
    up_w = u.up_conv3.up_sample.weight.detach().clone()
    res = load_pretrained_into_unet(u, path)
    assert torch.equal(u.down_conv2.double_conv.double_conv[3].weight, m.encoder_q.down_conv2.double_conv.double_conv[3].weight)
    assert not torch.equal(u.down_conv2.double_conv.double_conv[3].weight, m.encoder_k.down_conv2.double_conv.double_conv[3].weight)
    assert torch.equal(u.up_conv3.up_sample.weight, up_w)                       # decoder untouched: encoder-only hand-off
    assert all(k.startswith(('up_conv', 'conv_last')) for k in res.missing_keys)
    # detection by key prefix when the extension is not .ckpt (this package's own state_dict)
    res2 = load_pretrained_into_unet(C.UNet(), {'state_dict': m.state_dict()})
    assert all(k.startswith(('up_conv', 'conv_last')) for k in res2.missing_keys)
    # the same file loads in the oracle (== reference layout) UNet
    load_pretrained_into_unet(O.OracleUNet(), path)
    # nothing usable -> loud failure
    with pytest.raises(ValueError, match='no tensor for any encoder key'):
        load_pretrained_into_unet(C.UNet(), {'state_dict': {'encoder_x.down_conv1.weight': torch.zeros(1), 'foo.bar': torch.zeros(1)}})


def test_moco_queue_cache_signature_follows_buffer_writes():
    """The bf16 working copy of the MoCo queue is keyed on the buffers' version counters: load_state_dict / in-place
    writes invalidate it, the module's own bookkeeping does not."""
    torch.manual_seed(1)
    m = C.Moco_v2(emb_dim=1024, num_negatives=256)
    s0 = m._queue_sig()
    m.queue_ptr[0] = 64
    assert m._queue_sig() != s0
    s1 = m._queue_sig()
    m.load_state_dict({k: v.clone() for k, v in m.state_dict().items()})
    assert m._queue_sig() != s1
    assert m.shuffle_bn is True


def test_model_factory_survives_importing_the_library_builder():
    """`pkg.build(cfg)` (the stand-in for MODELS.build) keeps working after the driver's build() and after an explicit
    import of the `build` submodule, which Python would otherwise bind over the factory."""
    import importlib
    import __graft_entry__ as G
    G.build()
    assert type(C.build(C.cmunet_config(64))).__name__ == 'CM_UNet'
    importlib.import_module('contrastive_masked_unet_b200.build')
    import contrastive_masked_unet_b200 as P
    assert type(P.build(P.cmunet_config(64))).__name__ == 'CM_UNet'
    from contrastive_masked_unet_b200.build import build as build_library
    assert build_library.__module__.endswith('_buildlib')
