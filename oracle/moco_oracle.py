"""ORACLE (test infrastructure): plain fp32 torch restatement of the MoCo-v2 queue head
(Pretraining/MoCo/pl_bolts/models/self_supervised/moco/moco2_module.py:224-270 forward, :160-175 enqueue, :284 loss;
moco_data_module.py:47-66 encoder).

Parity status: PINNED.  The unmodified reference module is loaded in the build container through
oracle/ref_loader.import_moco() (stand-ins for pytorch-lightning / wandb / the missing pl_bolts pieces; the Moco_v2 class
itself runs as written) and oracle/make_goldens.py mints tests/golden/moco.json from it: two training steps (momentum
update, forward, InfoNCE loss, backward, dequeue/enqueue).  tests/test_oracle_pinned.py checks OracleMoco against it."""
import copy
import torch
import torch.nn.functional as F

from .cmunet_oracle import OracleEncoder, double_conv_fwd


class OracleMoco(torch.nn.Module):
    """moco2_module.py:51-148 (constructor order = RNG order: one base encoder with torch's default init, deep-copied into
    encoder_q / encoder_k, then `queue` and `val_queue` as unit-norm randn columns)."""

    def __init__(self, emb_dim=1024, num_negatives=65536, encoder_momentum=0.999, softmax_temperature=0.07):
        super().__init__()
        base = OracleEncoder()
        self.encoder_q, self.encoder_k = copy.deepcopy(base), copy.deepcopy(base)
        for pq, pk in zip(self.encoder_q.parameters(), self.encoder_k.parameters()):
            pk.data.copy_(pq.data)
            pk.requires_grad = False
        self.register_buffer('queue', F.normalize(torch.randn(emb_dim, num_negatives), dim=0))
        self.register_buffer('queue_ptr', torch.zeros(1, dtype=torch.long))
        self.register_buffer('val_queue', F.normalize(torch.randn(emb_dim, num_negatives), dim=0))
        self.register_buffer('val_queue_ptr', torch.zeros(1, dtype=torch.long))
        self.m, self.t = encoder_momentum, softmax_temperature

    @torch.no_grad()
    def momentum_update(self):
        """:153-158"""
        for pq, pk in zip(self.encoder_q.parameters(), self.encoder_k.parameters()):
            pk.data = pk.data * self.m + pq.data * (1.0 - self.m)

    def training_step(self, img_q, img_k):
        """:287-309 + :272-285: EMA, forward, enqueue, cross-entropy.  -> (loss, logits, k)"""
        self.momentum_update()
        q = F.normalize(moco_encoder_fwd(self.encoder_q, img_q), dim=1)
        with torch.no_grad():
            k = F.normalize(moco_encoder_fwd(self.encoder_k, img_k), dim=1)
        l_pos = torch.einsum('nc,nc->n', [q, k]).unsqueeze(-1)
        l_neg = torch.einsum('nc,ck->nk', [q, self.queue.clone().detach()])
        logits = torch.cat([l_pos, l_neg], dim=1) / self.t
        self.queue_ptr[0] = dequeue_and_enqueue(self.queue, int(self.queue_ptr), k)
        loss = F.cross_entropy(logits.float(), torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device))
        return loss, logits, k


def moco_inputs(n, s, step, seed=11):
    g = torch.Generator().manual_seed(seed + step)
    return torch.rand(n, 1, s, s, generator=g), torch.rand(n, 1, s, s, generator=g)


def moco_encoder_fwd(enc, x):
    """moco_data_module.py:58-66 (an OracleEncoder's layers without masking) -> (N, 1024)."""
    if x.dim() == 3:
        x = x.unsqueeze(1)
    for i in range(4):
        s = double_conv_fwd(getattr(enc, f'down_conv{i + 1}').double_conv, x)
        x = F.max_pool2d(s, 2)
    x = double_conv_fwd(enc.double_conv, x)
    return torch.mean(x, dim=[2, 3])


def moco_loss(q, k, queue, temperature=0.07):
    """q (N,D) raw, k (N,D) normalised keys, queue (D,K) as in the reference buffer -> scalar loss (:236-270, :284)."""
    q = F.normalize(q, dim=1)
    l_pos = torch.einsum('nc,nc->n', [q, k]).unsqueeze(-1)
    l_neg = torch.einsum('nc,ck->nk', [q, queue.clone().detach()])
    logits = torch.cat([l_pos, l_neg], dim=1) / temperature
    labels = torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device)
    return F.cross_entropy(logits.float(), labels)


def dequeue_and_enqueue(queue, ptr, keys):
    """:160-175 (single rank)."""
    n = keys.shape[0]
    assert queue.shape[1] % n == 0
    queue[:, ptr:ptr + n] = keys.T
    return (ptr + n) % queue.shape[1]
