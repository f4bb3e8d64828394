"""Pins oracle/cmunet_oracle.py (plain-torch restatement) against golden vectors minted by executing the
unmodified reference (oracle/make_goldens.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cmunet_oracle as O

G = os.path.join(os.path.dirname(__file__), 'golden')
PRE = json.load(open(os.path.join(G, 'pretrain.json')))['cases']
FT = json.load(open(os.path.join(G, 'finetune.json')))['case']
MOD = json.load(open(os.path.join(G, 'modules.json')))['cases']


def check_fp(t, fp, rtol=2e-4, atol=1e-6):
    t = t.detach().double().flatten().cpu()
    scale = max(fp['norm'], 1e-30)
    assert abs(float(t.norm()) - fp['norm']) <= rtol * scale + atol, (float(t.norm()), fp['norm'])
    vals = t[torch.tensor(fp['idx'])].numpy()
    np.testing.assert_allclose(vals, np.array(fp['val']), rtol=rtol * 50, atol=atol + 2e-4 * scale / max(1.0, t.numel() ** 0.5))


def run_pretrain_case(c):
    torch.manual_seed(c['seed'])
    m = O.OracleCMUNet(img_size=c['S'], np_seed=c['seed'])
    m.init_weights()
    m.train()
    assert [k for k, _ in m.named_parameters()] == c['param_keys']
    assert sum(p.numel() for p in m.parameters()) == c['n_params']
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == c['n_trainable']
    for k, p in m.named_parameters():
        check_fp(p, c['init'][k], rtol=1e-6)
    img, img_t = O.synthetic_batch(c['B'], c['S'], c['data_seed'])
    torch.manual_seed(c['seed'] + 1000)
    out = m(img, mode='loss', img_t=img_t)
    assert float(out['loss_ct']) == pytest.approx(c['loss_ct'], rel=1e-4)
    assert float(out['loss_rc']) == pytest.approx(c['loss_rc'], rel=1e-4)
    (out['loss_ct'] + out['loss_rc']).backward()
    assert sorted(k for k, p in m.named_parameters() if p.grad is None) == sorted(c['no_grad_keys'])
    for k, p in m.named_parameters():
        if p.grad is not None:
            check_fp(p.grad, c['grad'][k], rtol=2e-3)
    for k, b in m.named_buffers():
        check_fp(b.float(), c['buffers_after'][k], rtol=1e-4)
    m.momentum_update()
    for k, p in m.named_parameters():
        if k.startswith('target_'):
            check_fp(p, c['target_after_ema'][k], rtol=1e-6)


def test_pretrain_S64():
    run_pretrain_case(PRE[0])


def test_pretrain_S224_reference_native():
    run_pretrain_case(PRE[1])


@pytest.mark.slow
def test_pretrain_S512():
    run_pretrain_case(PRE[2])


@pytest.mark.slow
def test_pretrain_S224_B8():
    run_pretrain_case(PRE[3])


def test_finetune_config1():
    c = FT
    torch.manual_seed(c['seed'])
    net = O.OracleUNet().train()
    for k, p in net.named_parameters():
        check_fp(p, c['init'][k], rtol=1e-6)
    x = torch.rand(4, 256, 256)
    y1 = (torch.rand(4, 1, 256, 256) > 0.9)
    y = torch.cat([~y1, y1], 1).double()
    pred = net(x)
    check_fp(pred, c['pred'], rtol=1e-4)
    dice, ce, iou = O.dice_loss(pred, y), O.ce_prob_loss(pred, y), O.iou_loss(pred, y)
    assert float(dice) == pytest.approx(c['dice_loss'], rel=1e-6)
    assert float(ce) == pytest.approx(c['ce_loss'], rel=1e-5)
    assert float(iou) == pytest.approx(c['iou_loss'], rel=1e-6)
    total = dice + ce
    assert str(total.dtype) == c['total_dtype'] == 'torch.float64'
    assert not dice.requires_grad and c['dice_requires_grad'] is False      # Q7
    total.backward()
    for k, p in net.named_parameters():
        check_fp(p.grad, c['grad'][k], rtol=2e-3)
    net.eval()
    with torch.no_grad():
        check_fp(net(x), c['pred_eval'], rtol=1e-4)


def test_module_double_conv():
    c = MOD['double_conv']
    torch.manual_seed(c['seed'])
    dc = O._DC(8, 16).train()
    x = torch.randn(2, 8, 12, 20, requires_grad=True)
    y = O.double_conv_fwd(dc, x)
    (y * torch.linspace(0, 1, y.numel()).view_as(y)).sum().backward()
    check_fp(y, c['out'])
    check_fp(x.grad, c['dx'], rtol=1e-3)
    for k, p in dc.named_parameters():
        check_fp(p.grad, c['grads'][k], rtol=1e-3, atol=1e-4)
    for k, b in dc.named_buffers():
        check_fp(b.float(), c['buffers'][k])


def test_module_head():
    c = MOD['head']
    torch.manual_seed(c['seed'])
    head = O.OracleHead().train()
    B, S = c['B'], c['S']
    img = torch.randn(B, S, S)
    pred = torch.randn(B, 2, S, S, requires_grad=True)
    mask = (torch.rand(B, S, S) > 0.4).to(torch.uint8)
    ps = torch.randn(B, 1, 256, requires_grad=True)
    pt = torch.randn(B, 1, 256)
    out = head(img, pred[:, 1], mask, ps, pt)
    assert float(out['loss_ct']) == pytest.approx(c['loss_ct'], rel=1e-5)
    assert float(out['loss_rc']) == pytest.approx(c['loss_rc'], rel=1e-5)
    (out['loss_ct'] + out['loss_rc']).backward()
    check_fp(pred.grad, c['d_pred'])
    check_fp(ps.grad, c['d_proj_s'], rtol=1e-3)


def test_moco_oracle_matches_reference_golden():
    """OracleMoco (moco2_module.py restated) against tests/golden/moco.json, minted by running the unmodified reference
    Moco_v2 for two training steps (oracle/make_goldens.py gold_moco)."""
    from oracle import moco_oracle as MO
    c = json.load(open(os.path.join(G, 'moco.json')))['case']
    torch.manual_seed(c['seed'])
    m = MO.OracleMoco(emb_dim=1024, num_negatives=c['K']).train()
    for k, p in m.encoder_q.named_parameters():
        check_fp(p, c['init_q'][k], rtol=1e-6)
    check_fp(m.queue, c['init_queue'], rtol=1e-6)
    for step, gs in enumerate(c['steps']):
        img_q, img_k = MO.moco_inputs(c['N'], c['S'], step)
        m.zero_grad()
        loss, logits, keys = m.training_step(img_q, img_k)
        loss.backward()
        assert float(loss) == pytest.approx(gs['loss'], rel=1e-4)
        check_fp(logits, gs['logits'])
        check_fp(keys, gs['keys'])
        assert int(m.queue_ptr) == gs['queue_ptr']
        check_fp(m.queue, gs['queue'])
        for k, p in m.encoder_q.named_parameters():
            if not (k.endswith('double_conv.0.bias') or k.endswith('double_conv.3.bias')):   # analytically zero grads
                check_fp(p.grad, gs['grad_q'][k], rtol=2e-3)
        for k, p in m.encoder_k.named_parameters():
            check_fp(p, gs['enc_k'][k], rtol=1e-5)
        with torch.no_grad():
            for p in m.encoder_q.parameters():
                p -= 0.05 * p.grad


def test_moco_oracle_self_consistency():
    """moco_loss algebra against a manual log-sum-exp and the enqueue pointer arithmetic of moco2_module.py:160-175."""
    from oracle import moco_oracle as MO
    g = torch.Generator().manual_seed(3)
    q = torch.randn(8, 32, generator=g)
    k = torch.nn.functional.normalize(torch.randn(8, 32, generator=g), dim=1)
    queue = torch.nn.functional.normalize(torch.randn(32, 64, generator=g), dim=0)
    qh = torch.nn.functional.normalize(q, dim=1)
    logits = torch.cat([(qh * k).sum(1, keepdim=True), qh @ queue], 1) / 0.07
    manual = (torch.logsumexp(logits, 1) - logits[:, 0]).mean()
    assert abs(float(MO.moco_loss(q, k, queue, 0.07)) - float(manual)) < 1e-6
    ptr = 56
    ptr = MO.dequeue_and_enqueue(queue, ptr, k)
    assert ptr == 0 and torch.equal(queue[:, 56:], k.T)


def test_cldice_oracle_matches_reference_golden():
    """oracle.cldice_loss (FT/metrics.py:401-492 restated) against values minted from the reference's soft_cldice."""
    import json, os
    from oracle import cmunet_oracle as O
    gold = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'cldice.json')))['cases']
    for c in gold:
        logits, gt = O.cldice_inputs(c['n'], c['h'], c['w'], c['seed'])
        pr = (torch.softmax(logits, 1) > 0.5).float()[:, 1:]
        assert float(pr.sum()) == c['pred_pixels'] and float(gt[:, 1].sum()) == c['gt_pixels']
        assert float(O.soft_skel(pr).sum()) == c['skel_pred_sum']
        assert float(O.soft_skel(gt[:, 1:]).sum()) == c['skel_true_sum']
        v = O.cldice_loss(logits, gt)
        assert str(v.dtype) == c['dtype'] and abs(float(v) - c['cldice']) < 1e-12


def test_emulation_off_equals_oracle():
    """oracle/bf16_emulation.py with every rounding switch off IS the pinned oracle (same losses, same gradients); with
    the switches on it differs (the switches do something) but stays within bf16 distance."""
    from oracle import bf16_emulation as E
    from oracle import cmunet_oracle as O
    from oracle.mask_oracle import MT19937, patch_mask
    S, B, seed = 32, 4, 60
    torch.manual_seed(seed)
    o = O.OracleCMUNet(img_size=S, np_seed=seed)
    o.init_weights()
    o.train()
    img, img_t = O.synthetic_batch(B, S, 1)
    torch.manual_seed(seed + 1000)
    ref = o(img, mode='loss', img_t=img_t)
    (ref['loss_ct'] + ref['loss_rc']).backward()
    g_ref = {k: p.grad.clone() for k, p in o.named_parameters() if p.grad is not None}
    o.zero_grad()
    mask, _ = patch_mask(MT19937(seed), B, S, 16, 0.65)
    torch.manual_seed(seed + 1000)
    rc = torch.nn.Conv2d(1024, 256, kernel_size=1)
    out = E.forward_train(o, img, img_t, mask, rc.weight.detach(), rc.bias.detach())
    assert float(out['loss_ct']) == pytest.approx(float(ref['loss_ct']), rel=1e-6)
    assert float(out['loss_rc']) == pytest.approx(float(ref['loss_rc']), rel=1e-6)
    (out['loss_ct'] + out['loss_rc']).backward()
    for k, p in o.named_parameters():
        if k in g_ref:
            assert torch.allclose(p.grad, g_ref[k], rtol=1e-4, atol=1e-7), k
    o.zero_grad()
    out2 = E.forward_train(o, img, img_t, mask, rc.weight.detach(), rc.bias.detach(), **E.faithful(64))
    assert float(out2['loss_rc']) != float(ref['loss_rc'])
    assert float(out2['loss_rc']) == pytest.approx(float(ref['loss_rc']), rel=1e-2)
