"""Per-parameter gradient cosine report: CUDA path vs fp32 oracle, and (for context) the oracle under torch bf16
autocast vs the fp32 oracle.  python tools/grad_report.py S B [S B ...]  -> gpurun_out/grad_report.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import model_checks as M  # noqa: E402
from oracle import cmunet_oracle as O  # noqa: E402


def autocast_vs_fp32(S, B, seed=60, data_seed=1):
    import numpy as np
    torch.manual_seed(seed)
    o = O.OracleCMUNet(img_size=S, np_seed=seed); o.init_weights(); o = o.cuda().train()
    torch.manual_seed(seed)
    a = O.OracleCMUNet(img_size=S, np_seed=seed); a.init_weights(); a = a.cuda().train()
    img, img_t = O.synthetic_batch(B, S, data_seed)
    img, img_t = img.cuda(), img_t.cuda()
    torch.manual_seed(seed + 1000)
    lo = o(img, mode='loss', img_t=img_t); (lo['loss_ct'] + lo['loss_rc']).backward()
    torch.manual_seed(seed + 1000)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        la = a(img, mode='loss', img_t=img_t)
    (la['loss_ct'] + la['loss_rc']).backward()
    po = dict(o.named_parameters())
    tab = {}
    for k, p in a.named_parameters():
        if p.grad is None or M.is_zero_grad_key(k) or float(po[k].grad.norm()) < 1e-7:
            continue
        tab[k] = M.cosine(p.grad, po[k].grad)
    return {'loss_ct': (float(la['loss_ct']), float(lo['loss_ct'])), 'loss_rc': (float(la['loss_rc']), float(lo['loss_rc'])), 'cos': tab}


def main():
    args = [int(a) for a in sys.argv[1:]] or [64, 8]
    out = []
    for S, B in zip(args[0::2], args[1::2]):
        rep = M.pretrain_parity(S, B, verbose=True)
        ac = autocast_vs_fp32(S, B)
        rep['autocast'] = ac
        out.append(rep)
        print(f'== S={S} B={B} loss_ct {rep["loss_ct"]} loss_rc {rep["loss_rc"]} autocast ct {ac["loss_ct"]} rc {ac["loss_rc"]}')
        for k, c in rep['cos_table'].items():
            print(f'{c:.5f}  {ac["cos"].get(k, float("nan")):.5f}  {k}')
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'grad_report.json'), 'w'), indent=1, default=str)


if __name__ == '__main__':
    main()
