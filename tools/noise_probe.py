"""Where does the bf16 gradient noise of the contrastive branch come from?  Diagnostic (not a test): the fp32 oracle is
re-run with bf16 ROUNDING injected at one class of tensors at a time (forward values and/or gradients; the arithmetic
stays fp32) and the loss_ct / loss_rc gradients are compared with the clean fp32 run.

    python tools/noise_probe.py S B  -> gpurun_out/noise_probe_S{S}_B{B}.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import cmunet_oracle as O  # noqa: E402
from tests import model_checks as M  # noqa: E402

DEVICE = "cuda" if torch.cuda.is_available() else "cpu"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


from oracle import bf16_emulation as E  # noqa: E402


def run(o, img, img_t, mask, rw, rb, **kw):
    out = E.forward_train(o, img, img_t, mask, rw, rb, **kw)
    named = [(k, p) for k, p in o.named_parameters() if p.requires_grad]
    ps = [p for _, p in named]
    g_rc = torch.autograd.grad(out['loss_rc'], ps, retain_graph=True, allow_unused=True)
    g_ct = torch.autograd.grad(out['loss_ct'], ps, allow_unused=True)
    return ({'loss_ct': float(out['loss_ct']), 'loss_rc': float(out['loss_rc'])},
            {k: (a, b) for (k, _), a, b in zip(named, g_rc, g_ct)})


GROUPS = {'enc12': ('backbone.down_conv1', 'backbone.down_conv2'), 'enc345': ('backbone.down_conv3', 'backbone.down_conv4', 'backbone.double_conv'),
          'fdec': ('feature_decoder',), 'pdec': ('pixel_decoder',), 'proj': ('projector',), 'pred': ('head.predictor',)}


def group_cos(g, g0, which):
    out = {}
    for name, prefixes in GROUPS.items():
        num = da = db = 0.0
        worst = 2.0
        for k in g0:
            if not k.startswith(prefixes) or M.is_zero_grad_key(k) or k.endswith('fc0.bias') or k.endswith('conv_last.bias') \
                    or k.endswith('up_sample.bias'):
                continue
            a, b = g[k][which], g0[k][which]
            if a is None or b is None:
                continue
            a, b = a.double().flatten(), b.double().flatten()
            num += float(a @ b); da += float(a @ a); db += float(b @ b)
            worst = min(worst, float(a @ b) / max((float(a @ a) * float(b @ b)) ** 0.5, 1e-300))
        if da > 0:
            out[name] = (round(num / (da * db) ** 0.5, 5), round(worst, 5))
    return out


def main():
    S, B = int(sys.argv[1]), int(sys.argv[2])
    seed = 60
    torch.manual_seed(seed)
    o = O.OracleCMUNet(img_size=S, np_seed=seed)
    o.init_weights()
    o = o.to(DEVICE).train()
    img, img_t = O.synthetic_batch(B, S, 1)
    img, img_t = img.to(DEVICE), img_t.to(DEVICE)
    from oracle.mask_oracle import MT19937, patch_mask
    mask, _ = patch_mask(MT19937(seed), B, S, 16, 0.65)
    torch.manual_seed(seed + 1000)
    rc = torch.nn.Conv2d(1024, 256, 1).to(DEVICE)
    rw, rb = rc.weight.detach(), rc.bias.detach()
    ON, ALL = E.switches, E.ALL
    cases = {
        'clean': {},
        'all': dict(f_enc=ALL, f_dec=ALL, f_tgt=ALL, fc0_s=True, fc0_t=True),
        'target_path_only': dict(f_tgt=ALL, fc0_t=True),
        'target_encoder_only': dict(f_tgt=ALL),
        'fc0_target_only': dict(fc0_t=True),
        'fc0_online_only': dict(fc0_s=True),
        'online_fwd_acts_only': dict(f_enc=ON(a=True, y=True), f_dec=ON(a=True, y=True)),
        'online_weights_only': dict(f_enc=ON(w=True), f_dec=ON(w=True)),
        'online_grads_only': dict(f_enc=ON(g=True), f_dec=ON(g=True)),
        'encoder_all_only': dict(f_enc=ALL),
        'decoder_all_only': dict(f_dec=ALL),
        'all_but_target_path': dict(f_enc=ALL, f_dec=ALL, fc0_s=True),
    }
    res = {}
    g0 = None
    for name, kw in cases.items():
        losses, g = run(o, img, img_t, mask, rw, rb, **kw)
        if g0 is None:
            g0 = g
        res[name] = {'losses': losses, 'ct': group_cos(g, g0, 1), 'rc': group_cos(g, g0, 0)}
        print(name, json.dumps(res[name]), flush=True)
        if name != 'clean':
            del g
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', f'noise_probe_S{S}_B{B}.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
