"""Mints tests/golden/*.json by EXECUTING THE UNMODIFIED REFERENCE (/root/reference) on CPU in the build
container (see oracle/ref_loader.py for the shim + patches).  Run:  python -m oracle.make_goldens
The fixtures are small fingerprints (losses, per-tensor norm/sum/sampled entries, mask hashes); inputs and
initial weights are re-derivable from the recorded seeds because the oracle/product modules create their
parameters in the reference's order.

Test infrastructure only.  The reference has no tests or golden vectors of its own.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.cmunet_oracle import fingerprint, synthetic_batch  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def gold_masks():
    """create_random_patch_mask KATs (UNet_encoder.py:106-139) incl. Q2 call ordering."""
    _, MODELS, cfg = ref_loader.import_cmae()
    from cmae.models.backbones.UNet_encoder import UNet_encoder
    out = []
    for seed, B, S in [(60, 4, 224), (61, 4, 224), (60, 64, 512), (61, 8, 512), (42, 2, 1024), (7, 3, 64),
                       (60, 5, 256)]:
        online = UNet_encoder(patch_size=16, mask_ratio=0.65)
        target = UNet_encoder(patch_size=16, mask_ratio=0.0)
        np.random.seed(seed)
        rec = {'seed': seed, 'B': B, 'S': S, 'patch_size': 16, 'mask_ratio': 0.65, 'steps': []}
        for step in range(2):
            m_on = online.create_random_patch_mask(B, S)
            m_tg = target.create_random_patch_mask(B, S)
            rec['steps'].append({'online_sha16': sha16(m_on), 'online_sum_per_image': int(m_on[0].sum()),
                                 'online_img0_rowsum_head': [int(v) for v in m_on[0].sum(1)[:4]],
                                 'target_sum': int(m_tg.sum())})
        # position of the numpy stream after 2 steps: next raw 32-bit draws
        rec['next_u32'] = [int(v) for v in np.random.randint(0, 2 ** 32, size=4, dtype=np.uint64)]
        out.append(rec)
    # other ratios / patch sizes
    for seed, B, S, ps, ratio in [(3, 2, 128, 8, 0.5), (3, 2, 128, 32, 0.9), (9, 2, 96, 16, 1.0), (9, 2, 96, 16, 0.3)]:
        enc = UNet_encoder(patch_size=ps, mask_ratio=ratio)
        np.random.seed(seed)
        m = enc.create_random_patch_mask(B, S)
        out.append({'seed': seed, 'B': B, 'S': S, 'patch_size': ps, 'mask_ratio': ratio,
                    'steps': [{'online_sha16': sha16(m), 'online_sum_per_image': int(m[0].sum())}]})
    return out


def _fp_named(named):
    return {k: fingerprint(v) for k, v in named}


def gold_pretrain(S, B, seed=60, data_seed=1):
    """One full reference step: init_weights -> forward_train -> backward -> momentum_update."""
    ref_loader.apply_cpu_patches()
    ref_loader.ensure_process_group()
    torch.manual_seed(seed)
    np.random.seed(seed)
    t0 = time.time()
    model = ref_loader.build_reference_cm_unet(S)
    model.init_weights()
    model.train()
    img, img_t = synthetic_batch(B, S, data_seed)
    rec = {'S': S, 'B': B, 'seed': seed, 'data_seed': data_seed,
           'n_params': int(sum(p.numel() for p in model.parameters())),
           'n_trainable': int(sum(p.numel() for p in model.parameters() if p.requires_grad)),
           'param_keys': [k for k, _ in model.named_parameters()],
           'init': _fp_named(model.named_parameters())}
    torch.manual_seed(seed + 1000)          # pins the Q3 reduce_channels draw (first torch RNG use in fwd)
    losses = model(img, mode='loss', img_t=img_t)
    rec['loss_ct'] = float(losses['loss_ct'])
    rec['loss_rc'] = float(losses['loss_rc'])
    (losses['loss_ct'] + losses['loss_rc']).backward()
    rec['grad'] = _fp_named((k, p.grad) for k, p in model.named_parameters() if p.grad is not None)
    rec['no_grad_keys'] = [k for k, p in model.named_parameters() if p.grad is None]
    rec['buffers_after'] = _fp_named((k, b.float()) for k, b in model.named_buffers())
    model.momentum_update()
    rec['target_after_ema'] = _fp_named((k, p) for k, p in model.named_parameters()
                                        if k.startswith('target_'))
    rec['seconds'] = time.time() - t0
    return rec


def gold_modules(seed=5):
    """Per-module forward/grad fingerprints on tiny shapes (reference DoubleConv/DownBlock/UpBlock/
    MUNetPretrainDecoder/NonLinearNeck/CMUNetPretrainHead)."""
    ref_loader.apply_cpu_patches()
    ref_loader.ensure_process_group()
    _, MODELS, cfg = ref_loader.import_cmae()
    out = {}
    torch.manual_seed(seed)
    dc = MODELS.build(dict(type='DoubleConv', in_channels=8, out_channels=16)).train()
    x = torch.randn(2, 8, 12, 20, requires_grad=True)
    y = dc(x)
    (y * torch.linspace(0, 1, y.numel()).view_as(y)).sum().backward()
    out['double_conv'] = {'seed': seed, 'out': fingerprint(y), 'dx': fingerprint(x.grad),
                          'grads': _fp_named((k, p.grad) for k, p in dc.named_parameters()),
                          'buffers': _fp_named((k, b.float()) for k, b in dc.named_buffers())}
    torch.manual_seed(seed)
    up = MODELS.build(dict(type='UpBlock', in_channels=16, out_channels=8, up_sample_mode='conv_transpose')).train()
    d = torch.randn(2, 16, 6, 10, requires_grad=True)
    s = torch.randn(2, 8, 12, 20, requires_grad=True)
    y = up(d, s)
    (y * torch.linspace(-1, 1, y.numel()).view_as(y)).sum().backward()
    out['up_block'] = {'seed': seed, 'out': fingerprint(y), 'd_down': fingerprint(d.grad), 'd_skip': fingerprint(s.grad),
                       'grads': _fp_named((k, p.grad) for k, p in up.named_parameters())}
    torch.manual_seed(seed)
    head = MODELS.build(cfg['head']).train()
    B, S = 6, 32
    img = torch.randn(B, S, S)
    pred = torch.randn(B, 2, S, S, requires_grad=True)
    mask = (torch.rand(B, S, S) > 0.4).to(torch.uint8)
    ps = torch.randn(B, 1, 256, requires_grad=True)
    pt = torch.randn(B, 1, 256)
    losses = head(img, pred[:, 1], mask, ps, pt)
    (losses['loss_ct'] + losses['loss_rc']).backward()
    out['head'] = {'seed': seed, 'B': B, 'S': S, 'loss_ct': float(losses['loss_ct']), 'loss_rc': float(losses['loss_rc']),
                   'd_pred': fingerprint(pred.grad), 'd_proj_s': fingerprint(ps.grad),
                   'grads': _fp_named((k, p.grad) for k, p in head.named_parameters())}
    return out


def gold_finetune(seed=0):
    """BASELINE.json configs[0]: FT/model.py UNet fwd+bwd, B=4, 256², Dice(thr .5, ignore ch0)+CE."""
    model_mod, metrics = ref_loader.import_finetune()
    torch.manual_seed(seed)
    net = model_mod.UNet().train()
    x = torch.rand(4, 256, 256)
    y1 = (torch.rand(4, 1, 256, 256) > 0.9)
    y = torch.cat([~y1, y1], 1).double()
    dice = metrics.DiceLoss(activation='softmax', threshold=0.5, ignore_channels=[0])
    ce = metrics.CrossEntropyLoss()
    iou = metrics.IoU(activation='softmax', threshold=0.5, ignore_channels=[0])
    loss = dice + ce
    t0 = time.time()
    pred = net.forward(x)
    total = loss(pred, y)
    total.backward()
    rec = {'seed': seed, 'B': 4, 'S': 256, 'loss_name': loss.__name__, 'dice_name': dice.__name__,
           'ce_name': ce.__name__, 'iou_name': iou.__name__,
           'dice_loss': float(dice(pred, y)), 'ce_loss': float(ce(pred, y)), 'iou_loss': float(iou(pred, y)),
           'total': float(total), 'total_dtype': str(total.dtype),
           'dice_requires_grad': bool(dice(pred, y).requires_grad),
           'pred': fingerprint(pred), 'init': _fp_named(net.named_parameters()),
           'grad': _fp_named((k, p.grad) for k, p in net.named_parameters()),
           'seconds': time.time() - t0}
    net.eval()
    with torch.no_grad():
        rec['pred_eval'] = fingerprint(net.forward(x))
    return rec


def gold_moco(seed=7, N=64, S=32, K=1024, steps=2):
    """BASELINE.json configs[3] scaled down: the reference Moco_v2 (moco2_module.py) for two training steps on CPU, with a
    plain SGD step on encoder_q in between so that the momentum update of the second step is non-trivial."""
    from oracle.moco_oracle import moco_inputs
    mm = ref_loader.import_moco()
    torch.manual_seed(seed)
    mod = mm.Moco_v2(emb_dim=1024, num_negatives=K, batch_size=N).train()
    rec = {'seed': seed, 'N': N, 'S': S, 'K': K, 'init_q': _fp_named(mod.encoder_q.named_parameters()),
           'init_queue': fingerprint(mod.queue), 'steps': []}
    for step in range(steps):
        img_q, img_k = moco_inputs(N, S, step)
        mod.zero_grad()
        mod._momentum_update_key_encoder()                                  # training_step, moco2_module.py:297
        logits, labels, keys, _ = mod(img_q=img_q, img_k=img_k, queue=mod.queue)
        loss = mod._compute_l_s(logits, labels, keys, queue=mod.queue)        # enqueue + cross-entropy
        loss.backward()
        rec['steps'].append({'loss': float(loss), 'logits': fingerprint(logits), 'keys': fingerprint(keys),
                             'labels_dtype': str(labels.dtype), 'queue_ptr': int(mod.queue_ptr),
                             'queue': fingerprint(mod.queue),
                             'grad_q': _fp_named((k, p.grad) for k, p in mod.encoder_q.named_parameters()),
                             'enc_k': _fp_named(mod.encoder_k.named_parameters())})
        with torch.no_grad():
            for p in mod.encoder_q.parameters():
                p -= 0.05 * p.grad
    return rec


def gold_cldice():
    """FT/metrics.py:401-430 soft_cldice (the eval metric of FT/train.py:464) on deterministic synthetic inputs."""
    from oracle.cmunet_oracle import cldice_inputs
    _, metrics = ref_loader.import_finetune()
    m = metrics.soft_cldice(threshold=0.5, activation='softmax', ignore_channels=[0])
    cases = []
    for n, h, w, seed in ((2, 64, 64, 3), (3, 48, 80, 4), (1, 33, 21, 5)):
        logits, gt = cldice_inputs(n, h, w, seed)
        val = m(logits, gt)
        pr = (torch.softmax(logits, 1) > 0.5).float()[:, 1:]
        cases.append({'n': n, 'h': h, 'w': w, 'seed': seed, 'name': m.__name__, 'cldice': float(val), 'dtype': str(val.dtype),
                      'pred_pixels': float(pr.sum()), 'gt_pixels': float(gt[:, 1].sum()),
                      'skel_pred_sum': float(m.soft_skeletonize(pr).sum()),
                      'skel_true_sum': float(m.soft_skeletonize(gt[:, 1:]).sum())})
    return cases


def main():
    if not ref_loader.reference_available():
        raise SystemExit('reference tree not found: goldens can only be minted in the build container')
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ['masks', 'modules', 'finetune', 'pretrain']
    meta = {'torch': torch.__version__, 'numpy': np.__version__, 'generated_by': 'oracle/make_goldens.py'}
    if 'masks' in which:
        json.dump({'meta': meta, 'cases': gold_masks()}, open(os.path.join(GOLD, 'masks.json'), 'w'), indent=1)
        print('masks done')
    if 'modules' in which:
        json.dump({'meta': meta, 'cases': gold_modules()}, open(os.path.join(GOLD, 'modules.json'), 'w'), indent=1)
        print('modules done')
    if 'moco' in which:
        json.dump({'meta': meta, 'case': gold_moco()}, open(os.path.join(GOLD, 'moco.json'), 'w'), indent=1)
        print('moco done')
    if 'cldice' in which:
        json.dump({'meta': meta, 'cases': gold_cldice()}, open(os.path.join(GOLD, 'cldice.json'), 'w'), indent=1)
        print('cldice done')
    if 'finetune' in which:
        json.dump({'meta': meta, 'case': gold_finetune()}, open(os.path.join(GOLD, 'finetune.json'), 'w'), indent=1)
        print('finetune done')
    if 'pretrain' in which:
        cases = [gold_pretrain(64, 8), gold_pretrain(224, 4), gold_pretrain(512, 2), gold_pretrain(224, 8)]
        json.dump({'meta': meta, 'cases': cases}, open(os.path.join(GOLD, 'pretrain.json'), 'w'), indent=1)
        print('pretrain done', [c['seconds'] for c in cases])


if __name__ == '__main__':
    main()
