"""Mints tests/golden/head_2rank.json: the UNMODIFIED reference CMUNetPretrainHead (cmunet_head.py:47-91) executed by
two gloo ranks on CPU (all-gathered negatives, labels offset by bs*rank).  Run: python -m oracle.make_goldens_dist
Test infrastructure only (build container only: needs /root/reference)."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def head_inputs(rank, B=6, S=32):
    g = torch.Generator().manual_seed(100 + rank)
    img = torch.randn(B, S, S, generator=g)
    pred = torch.randn(B, 2, S, S, generator=g)
    mask = (torch.rand(B, S, S, generator=g) > 0.4).to(torch.uint8)
    ps = torch.randn(B, 1, 256, generator=g)
    pt = torch.randn(B, 1, 256, generator=g)
    return img, pred, mask, ps, pt


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from oracle import ref_loader
    ref_loader.apply_cpu_patches()
    dist.init_process_group('gloo', rank=rank, world_size=world)
    _, MODELS, cfg = ref_loader.import_cmae()
    torch.manual_seed(5)                      # same predictor weights on both ranks
    head = MODELS.build(cfg['head']).train()
    img, pred, mask, ps, pt = head_inputs(rank)
    pred.requires_grad_(True)
    ps.requires_grad_(True)
    losses = head(img, pred[:, 1], mask, ps, pt)
    (losses['loss_ct'] + losses['loss_rc']).backward()
    from oracle.cmunet_oracle import fingerprint
    out[rank] = {'loss_ct': float(losses['loss_ct']), 'loss_rc': float(losses['loss_rc']),
                 'd_proj_s': fingerprint(ps.grad), 'd_pred': fingerprint(pred.grad)}
    dist.destroy_process_group()


def main():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29631, out), nprocs=2, join=True)
    res = {'meta': {'generated_by': 'oracle/make_goldens_dist.py', 'note': 'SyncBN -> BatchNorm1d on CPU (unsynced), as '
                    'in the shim; what is pinned here is the gather order and the bs*rank label offset'},
           'ranks': {str(k): v for k, v in out.items()}}
    json.dump(res, open(os.path.join(ROOT, 'tests', 'golden', 'head_2rank.json'), 'w'), indent=1)
    print(res['ranks'])


if __name__ == '__main__':
    main()
