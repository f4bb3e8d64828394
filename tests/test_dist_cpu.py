"""world_size-2 gloo tests on CPU for the data-parallel host logic of the head: gather order of the
negatives, the bs*rank label offset, and SyncBN statistics over the global batch -- pinned against a golden minted from
the unmodified reference head run by two gloo ranks (oracle/make_goldens_dist.py)."""
import json
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'head_2rank.json')))['ranks']


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import torch.nn.functional as F
    from oracle import cmunet_oracle as O
    from oracle.make_goldens_dist import head_inputs
    from contrastive_masked_unet_b200.modules import _rank_world, concat_all_gather
    res = {}
    # (a) oracle head, 2 ranks, unsynced BN (the CPU shim the golden was minted with)
    O.OracleNeck.sync_bn = False
    torch.manual_seed(5)
    head = O.OracleHead().train()
    img, pred, mask, ps, pt = head_inputs(rank)
    pred.requires_grad_(True)
    ps.requires_grad_(True)
    losses = head(img, pred[:, 1], mask, ps, pt)
    (losses['loss_ct'] + losses['loss_rc']).backward()
    res['loss_ct'], res['loss_rc'] = float(losses['loss_ct']), float(losses['loss_rc'])
    res['d_proj_s_norm'] = float(ps.grad.double().norm())
    # (b) product gather helper: rank order, no gradient
    z = torch.full((3, 4), float(rank), requires_grad=True)
    allz = concat_all_gather(z)
    res['gather_ok'] = bool(allz.shape == (6, 4) and not allz.requires_grad and
                            torch.equal(allz[:3], torch.zeros(3, 4)) and torch.equal(allz[3:], torch.ones(3, 4)))
    res['rank_world'] = _rank_world()
    # (c) synced BN of the oracle == BN over the concatenated global batch
    O.OracleNeck.sync_bn = True
    torch.manual_seed(7)
    neck = O.OracleNeck(64, 32, 16).train()
    g = torch.Generator().manual_seed(200 + rank)
    x = torch.randn(5, 1, 8, 8, generator=g)
    y = neck(x)
    xs = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(xs, x)
    torch.manual_seed(7)
    ref = O.OracleNeck(64, 32, 16).train()
    O.OracleNeck.sync_bn = False
    yr = ref(torch.cat(xs, 0))[rank * 5:(rank + 1) * 5]
    res['syncbn_err'] = float((y - yr).abs().max())
    res['syncbn_rv_err'] = float((neck.bn0.running_var - ref.bn0.running_var).abs().max())
    # (d) product host logic of MoCo's shuffle-BN (moco.py `_batch_shuffle_ddp` / `_batch_unshuffle_ddp` ==
    # moco2_module.py:177-222): one permutation for all ranks, every sample visits exactly one rank, unshuffle restores
    # each rank's own rows in order
    import contrastive_masked_unet_b200 as C
    torch.manual_seed(100 + rank)                      # different local seeds: the permutation must come from rank 0
    m = C.Moco_v2(emb_dim=1024, num_negatives=64)
    xb = torch.arange(6, dtype=torch.float32).reshape(6, 1) + 100.0 * rank          # rows tagged (rank, index)
    shuffled, idx_unshuffle = m._batch_shuffle_ddp(xb)
    gathered = [torch.empty_like(shuffled) for _ in range(world)]
    dist.all_gather(gathered, shuffled)
    union = torch.cat(gathered).flatten().sort().values
    expect = torch.cat([torch.arange(6.) + 100.0 * r for r in range(world)])
    res['shuffle_is_permutation'] = bool(torch.equal(union, expect))
    res['shuffle_mixes_ranks'] = bool((shuffled.flatten() // 100 != rank).any())
    restored = m._batch_unshuffle_ddp(shuffled * 2.0, idx_unshuffle)               # "encoder" = x -> 2x
    res['unshuffle_ok'] = bool(torch.equal(restored, xb * 2.0))
    idx_all = [torch.empty_like(idx_unshuffle) for _ in range(world)]
    dist.all_gather(idx_all, idx_unshuffle)
    res['same_permutation_on_all_ranks'] = bool(all(torch.equal(i, idx_all[0]) for i in idx_all))
    out[rank] = res
    dist.destroy_process_group()


def test_two_rank_head_semantics_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29641, out), nprocs=2, join=True)
    for r in (0, 1):
        g, o = GOLD[str(r)], out[r]
        assert o['loss_ct'] == pytest.approx(g['loss_ct'], rel=1e-5)
        assert o['loss_rc'] == pytest.approx(g['loss_rc'], rel=1e-5)
        assert o['d_proj_s_norm'] == pytest.approx(g['d_proj_s']['norm'], rel=1e-4)
        assert o['gather_ok'] and o['rank_world'] == (r, 2)
        assert o['syncbn_err'] < 1e-5 and o['syncbn_rv_err'] < 1e-5
        assert o['shuffle_is_permutation'] and o['unshuffle_ok'] and o['same_permutation_on_all_ranks'], o
    assert out[0]['shuffle_mixes_ranks'] or out[1]['shuffle_mixes_ranks']
    assert out[0]['loss_ct'] != out[1]['loss_ct']      # the label offset makes the ranks' losses differ
