"""2-rank data-parallel parity worker (run by tests/test_dist_gpu.py through torch.distributed.run): the drop-in CM_UNet
under DDP vs the oracle on the same ranks (NCCL): per-rank losses (all-gathered negatives + bs*rank labels + SyncBN
statistics over the global batch) and the all-reduced gradients."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402


def main():
    rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import contrastive_masked_unet_b200 as C
    from oracle import cmunet_oracle as O
    S, B, seed = 64, 16, 60
    torch.manual_seed(seed)
    np.random.seed(seed + rank)
    m = C.build(C.cmunet_config(S))
    m.init_weights()
    m = m.to(dev).train()
    torch.manual_seed(seed)
    o = O.OracleCMUNet(img_size=S, np_seed=seed + rank)
    o.init_weights()
    o = o.to(dev).train()
    dm = DDP(m, device_ids=[local], broadcast_buffers=False)
    do = DDP(o, device_ids=[local], broadcast_buffers=False)
    img, img_t = O.synthetic_batch(B, S, 1 + rank)
    img, img_t = img.to(dev), img_t.to(dev)
    torch.manual_seed(seed + 1000)
    lo = do(img, mode='loss', img_t=img_t)
    (lo['loss_ct'] + lo['loss_rc']).backward()
    torch.manual_seed(seed + 1000)
    lm = dm(img, mode='loss', img_t=img_t)
    (lm['loss_ct'] + lm['loss_rc']).backward()
    torch.cuda.synchronize()
    po = dict(o.named_parameters())
    num = den_a = den_b = 0.0
    for k, p in m.named_parameters():
        if p.grad is None or k.endswith('double_conv.0.bias') or k.endswith('double_conv.3.bias'):
            continue
        a, b = p.grad.double().flatten(), po[k].grad.double().flatten()
        num += float(a @ b)
        den_a += float(a @ a)
        den_b += float(b @ b)
    # gradients must be identical across ranks after the all-reduce
    flat = torch.cat([p.grad.flatten() for p in m.parameters() if p.grad is not None])[:100000].clone()
    ref = flat.clone()
    dist.broadcast(ref, 0)
    res = {'rank': rank, 'loss_ct': (float(lm['loss_ct']), float(lo['loss_ct'])),
           'loss_rc': (float(lm['loss_rc']), float(lo['loss_rc'])),
           'grad_cos_all': num / (den_a ** 0.5 * den_b ** 0.5), 'grads_equal_across_ranks': bool(torch.equal(flat, ref)),
           'bn_rm_err': float((m.projector.bn0.running_mean - o.projector.bn0.running_mean).abs().max() /
                              o.projector.bn0.running_mean.abs().max())}
    # MoCo queue head across ranks (moco2_module.py:160-175 + :404-413): keys of ALL ranks enter every rank's queue in rank
    # order, so the queues stay identical; the first 2N columns are the two ranks' normalised keys.
    torch.manual_seed(7)
    N, K = 64, 1024
    moco = C.Moco_v2(emb_dim=1024, num_negatives=K).to(dev).train()
    gq = torch.Generator().manual_seed(100 + rank)
    iq, ik = torch.rand(N, 64, 64, generator=gq).to(dev), torch.rand(N, 64, 64, generator=gq).to(dev)
    loss = moco.training_step(iq, ik)
    loss.backward()
    with torch.no_grad():
        k_local = torch.nn.functional.normalize(moco.encoder_k(ik).float(), dim=1)
    qsum = moco.queue[:, :world * N].double().sum().reshape(1)
    qs = [torch.zeros_like(qsum) for _ in range(world)]
    dist.all_gather(qs, qsum)
    torch.cuda.synchronize()
    res['moco'] = {'loss': float(loss), 'ptr': int(moco.queue_ptr),
                   'queues_equal': bool(all(torch.equal(q, qs[0]) for q in qs)),
                   'own_keys_err': float((moco.queue[:, rank * N:(rank + 1) * N].t() - k_local).abs().max()),
                   'rows_match_queue': float((moco._rows[:world * N].float() - moco.queue[:, :world * N].t()).abs().max())}
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        print('DIST_RESULT ' + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
