"""pytest -m gpu: drop-in modules vs the pinned oracle and the golden values minted from the reference."""
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def _gpu():
    if not torch.cuda.is_available():
        pytest.skip('needs a B200')


def test_pretrain_step_S64_B8_vs_oracle_and_golden():
    _gpu()
    from tests import model_checks as M
    rep = M.pretrain_parity(64, 8, golden=M.golden_pretrain(64, 8))
    assert not rep['fails'], (rep['fails'][:8], {k: v for k, v in rep.items() if k not in ('fails', 'cos_table')})


def test_pretrain_step_S224_B8_reference_native_size():
    """224 x 224 is the size the unmodified reference runs at (cmunet.py:130 literal).  B = 8: with fewer rows the
    SyncBN of the projection head (statistics over B rows) makes loss_ct ill-conditioned in ANY reduced precision."""
    _gpu()
    from tests import model_checks as M
    rep = M.pretrain_parity(224, 8, golden=M.golden_pretrain(224, 8))
    assert not rep['fails'], (rep['fails'][:8], {k: v for k, v in rep.items() if k not in ('fails', 'cos_table')})


def test_pretrain_step_S128_B16():
    _gpu()
    from tests import model_checks as M
    rep = M.pretrain_parity(128, 16, seed=61, data_seed=3)
    assert not rep['fails'], (rep['fails'][:8], {k: v for k, v in rep.items() if k not in ('fails', 'cos_table')})


def test_finetune_config1_vs_oracle_and_golden():
    _gpu()
    from tests import model_checks as M
    rep = M.finetune_parity(4, 256, golden=M.golden_finetune())
    assert not rep['fails'], (rep['fails'][:8], {k: v for k, v in rep.items() if k not in ('fails', 'cos_table')})
    assert rep['name'] == 'dice_loss + cross_entropy_loss'


def test_standalone_blocks_accept_fp32_nchw():
    """DoubleConv / DownBlock / UpBlock work as standalone modules on plain NCHW fp32 tensors."""
    _gpu()
    import contrastive_masked_unet_b200 as C
    from oracle import cmunet_oracle as O
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(5)
    up = C.UpBlock(128, 64, 'conv_transpose').cuda().train()
    torch.manual_seed(5)
    ref = O._Up(128, 64).cuda().train()
    d = torch.randn(2, 128, 12, 10, device='cuda', requires_grad=True)
    s = torch.randn(2, 64, 24, 20, device='cuda', requires_grad=True)
    y = up(d, s)
    import torch.nn.functional as F
    yr = O.double_conv_fwd(ref.double_conv, torch.cat([F.conv_transpose2d(d, ref.up_sample.weight, ref.up_sample.bias, stride=2), s], 1))
    assert y.shape == yr.shape
    assert (y.float() - yr).abs().max() <= 5e-2 * yr.abs().max()
    w = torch.linspace(-1, 1, y.numel(), device='cuda').view_as(yr)
    (y.float() * w).sum().backward()
    gd, gs = d.grad.clone(), s.grad.clone()
    d.grad = s.grad = None
    (yr * w).sum().backward()
    cos = torch.nn.functional.cosine_similarity
    # the fp32 inputs are quantised to bf16 at the module boundary, hence 0.995 here (kernel-level checks: > 0.9999)
    assert cos(gd.flatten(), d.grad.flatten(), dim=0) > 0.995
    assert cos(gs.flatten(), s.grad.flatten(), dim=0) > 0.995
    with pytest.raises(ValueError):
        C.UpBlock(8, 4, 'nearest')


def test_cpu_tensor_raises_no_fallback():
    _gpu()
    import contrastive_masked_unet_b200 as C
    net = C.UNet().cuda()
    with pytest.raises(C.CmuError):
        net(torch.rand(1, 32, 32))


def test_pretrain_step_S512_B2_vs_golden_losses():
    """512 x 512 (the benchmark size) against the golden minted from the S-generalised reference at B=2.  With two rows
    the projection-head BatchNorm is degenerate (x_hat = +-1), so loss_ct is only sanity-checked; loss_rc is exact to
    bf16 tolerance."""
    _gpu()
    from tests import model_checks as M
    g = M.golden_pretrain(512, 2)
    rep = M.pretrain_parity(512, 2, golden=None, autocast_ref=False, grad_cos=-1.0)
    assert abs(rep['loss_rc'][0] - g['loss_rc']) <= 1e-2 * g['loss_rc'], rep['loss_rc']
    assert abs(rep['loss_rc'][0] - rep['loss_rc'][1]) <= 1e-2 * rep['loss_rc'][1]
    assert abs(rep['loss_ct'][0] - g['loss_ct']) <= 0.25 * g['loss_ct'] + 0.02, (rep['loss_ct'], g['loss_ct'])
    assert rep['mask_mismatch'] == 0 and rep['ema_max_abs_err'] <= 1e-6


def test_pretrain_B64_S512_benchmark_config_losses_and_split_gradients():
    """The configuration the headline metric is quoted on (BASELINE.json configs[1]: B = 64 @ 512^2), model level:
    loss_ct / loss_rc within 1e-2 of the fp32 oracle with NO widening, and per-parameter gradient cosines for the
    loss_rc-only, loss_ct-only and summed backward passes (rules: tests/model_checks.check_split_parity; numbers:
    profiles/r2_grad_parity_S512_B64.md).  The fp32 oracle, the bf16 emulation, the drop-in and torch's bf16 autocast
    run one after the other (the fp32 runs need ~135 GB)."""
    _gpu()
    from tests import model_checks as M
    rep = M.split_grad_parity(512, 64)
    fails = M.check_split_parity(rep, band=0.02, loss_rtol=1e-2)
    summ = M.summarize_split(rep)
    assert not fails, (fails[:8], summ)
    # reconstruction branch: every decoder parameter from up_conv2 outwards and the first two encoder levels reach 0.999
    for k, r in rep['table'].items():
        if r.get('zero_by_construction') or r.get('rc') is None:
            continue
        if k.startswith(('pixel_decoder.up_conv1.d', 'pixel_decoder.up_conv2.d', 'pixel_decoder.conv_last.w',
                         'backbone.down_conv1.', 'backbone.down_conv2.')):
            assert r['rc'] >= 0.999, (k, r)


def test_pretrain_S128_B16_split_gradients():
    _gpu()
    from tests import model_checks as M
    rep = M.split_grad_parity(128, 16, seed=61, data_seed=3)
    # small batch: SyncBN statistics over 16 rows, the run-to-run spread of every bf16 realisation is wider -> 0.04;
    # loss_ct: 1.5e-2 (the B = 64 test holds the 1e-2 bar)
    fails = M.check_split_parity(rep, band=0.04, loss_rtol=1.5e-2)
    assert not fails, (fails[:8], M.summarize_split(rep))


def test_finetune_S1024_config5_shape():
    """BASELINE.json configs[4] resolution (1024 x 1024) on one GPU, batch 1: forward/backward parity with the oracle."""
    _gpu()
    from tests import model_checks as M
    rep = M.finetune_parity(1, 1024, seed=3)
    assert not rep['fails'], (rep['fails'][:8], {k: v for k, v in rep.items() if k != 'fails'})


def test_moco_queue_head_two_steps_vs_oracle():
    """BASELINE.json configs[3] scaled down (N=64, S=64, K=4096): loss, encoder_q grads, EMA of encoder_k, queue."""
    _gpu()
    from tests import model_checks as M
    rep = M.moco_parity(64, 64, 4096)
    assert not rep['fails'], rep


def test_moco_two_steps_vs_reference_golden():
    """configs[3] scaled down, against values minted from the unmodified reference Moco_v2 (tests/golden/moco.json)."""
    _gpu()
    from tests import model_checks as M
    rep = M.moco_vs_golden()
    assert not rep['fails'], rep


def test_frozen_bn_conv_bias_gets_its_gradient():
    """eval-mode (frozen) BatchNorm: the conv bias gradient is the per-channel sum of dy, not the analytic zero of the
    batch-statistics case (functional.py ConvBNReLUFn.backward)."""
    _gpu()
    import contrastive_masked_unet_b200 as C
    from oracle import cmunet_oracle as O
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(11)
    dc = C.DoubleConv(64, 64).cuda()
    torch.manual_seed(11)
    ref = O._DC(64, 64).cuda()
    for m in (dc, ref):
        with torch.no_grad():
            for bn in (m.double_conv[1], m.double_conv[4]):
                bn.running_mean.normal_(0, 0.2)
                bn.running_var.uniform_(0.5, 1.5)
                bn.weight.uniform_(0.5, 1.5)
                bn.bias.normal_(0, 0.2)
    ref.load_state_dict(dc.state_dict())
    dc.eval()
    ref.eval()
    x = torch.randn(3, 64, 24, 20, device='cuda')
    y = dc(x)
    yr = O.double_conv_fwd(ref, x)
    w = torch.linspace(-1, 1, y.numel(), device='cuda').view_as(yr)
    (y.float() * w).sum().backward()
    (yr * w).sum().backward()
    cos = torch.nn.functional.cosine_similarity
    for i in (0, 3):
        g, gr = dc.double_conv[i].bias.grad, ref.double_conv[i].bias.grad
        assert float(gr.norm()) > 0
        assert cos(g.flatten(), gr.flatten(), dim=0) > 0.999, (i, g[:4], gr[:4])
        gw, gwr = dc.double_conv[i].weight.grad, ref.double_conv[i].weight.grad
        # fp32 inputs are quantised to bf16 at the module boundary (same bar as test_standalone_blocks_accept_fp32_nchw)
        assert cos(gw.flatten(), gwr.flatten(), dim=0) > 0.995


def test_moco_resume_from_state_dict_uses_loaded_queue_and_pointer():
    """step, save, step, load the saved state, step again: the second model continues exactly like a fresh model that
    loaded the same state (queue contents, pointer, bf16 working copy) -- ADVICE r1, moco.py `_queue_rows` cache."""
    _gpu()
    import contrastive_masked_unet_b200 as C
    torch.manual_seed(3)
    m = C.Moco_v2(emb_dim=1024, num_negatives=512).cuda().train()
    g = torch.Generator().manual_seed(5)
    batches = [(torch.rand(64, 32, 32, generator=g).cuda(), torch.rand(64, 32, 32, generator=g).cuda()) for _ in range(3)]
    m.training_step(*batches[0]).backward()
    saved = {k: v.clone() for k, v in m.state_dict().items()}
    m.training_step(*batches[1]).backward()                 # moves the pointer and the cached rows past the saved state
    assert int(m.queue_ptr) == 128
    m.load_state_dict(saved)                                # resume
    assert int(m.queue_ptr) == 64
    loss_a = m.training_step(*batches[2])
    torch.manual_seed(3)
    fresh = C.Moco_v2(emb_dim=1024, num_negatives=512).cuda().train()
    fresh.load_state_dict(saved)
    loss_b = fresh.training_step(*batches[2])
    torch.cuda.synchronize()
    # (the conv statistics use shared-memory float atomics: two runs agree to rounding, not to the bit)
    assert abs(float(loss_a) - float(loss_b)) <= 1e-4 * abs(float(loss_b)), (float(loss_a), float(loss_b))
    assert int(m.queue_ptr) == int(fresh.queue_ptr) == 128
    assert torch.equal(m.queue[:, 128:], fresh.queue[:, 128:]) and torch.equal(m.queue[:, 128:], saved['queue'][:, 128:])
    assert float((m.queue[:, :128] - fresh.queue[:, :128]).abs().max()) < 2e-2      # keys of a bf16 encoder
    assert float((m._rows.float() - m.queue.t()).abs().max()) < 1e-2
    # without the cache invalidation the stale rows / pointer would have been used: the step-2 keys sit in columns 64..127
    assert not torch.equal(saved['queue'][:, 64:128], m.queue[:, 64:128])


def test_mask_stream_position_after_training_steps_with_deferred_prefetch():
    """three fwd+bwd steps (the prefetch of the next pair is fired from the backward thread), then the stream position
    and the next online mask equal numpy's after 3 x (B online + B target) shuffles -- bit-exact (Q2)."""
    _gpu()
    import numpy as np
    import contrastive_masked_unet_b200 as C
    from oracle.mask_oracle import MT19937, patch_mask
    S, B, seed = 64, 6, 77
    torch.manual_seed(seed)
    np.random.seed(seed)
    m = C.build(C.cmunet_config(S))
    m.init_weights()
    m = m.cuda().train()
    rng = MT19937(seed)
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        img = torch.randn(B, S, S, generator=g).cuda()
        out = m(img, mode='loss', img_t=img + 0.1)
        (out['loss_ct'] + out['loss_rc']).backward()
        patch_mask(rng, B, S, 16, 0.65)
        patch_mask(rng, B, S, 16, 0.0)
        assert m.backbone.mask_stream._pref is not None, 'the deferred prefetch must have fired at the end of backward'
    torch.cuda.synchronize()
    st = m.backbone.mask_stream.get_numpy_state()
    r2 = MT19937()
    r2.set_state(st[1], st[2])
    key, pos = rng.get_state()
    r3 = MT19937()
    r3.set_state(key, pos)
    assert [r2.next_u32() for _ in range(4)] == [r3.next_u32() for _ in range(4)]
    _, mask, _ = m(torch.randn(B, S, S, generator=g).cuda(), mode='tensor')      # consumes the prefetched online half
    ref, _ = patch_mask(rng, B, S, 16, 0.65)
    assert np.array_equal(mask.cpu().numpy(), ref)
    # no target call followed: the speculatively drawn target half is rolled back for the next call
    _, mask2, _ = m(torch.randn(B, S, S, generator=g).cuda(), mode='tensor')
    ref2, _ = patch_mask(rng, B, S, 16, 0.65)
    assert np.array_equal(mask2.cpu().numpy(), ref2)
