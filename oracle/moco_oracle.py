"""ORACLE (test infrastructure): plain fp32 torch restatement of the MoCo-v2 queue head
(Pretraining/MoCo/pl_bolts/models/self_supervised/moco/moco2_module.py:224-270 forward, :160-175 enqueue, :284 loss;
moco_data_module.py:47-66 encoder).

Parity status: UNPINNED.  The reference module cannot be imported in the build container (it needs pytorch-lightning 1.6
and the vendored pl_bolts subset imports modules that are not in the tree), and the reference has no
tests or golden vectors; this file restates the published lines and is checked only for self-consistency."""
import torch
import torch.nn.functional as F

from .cmunet_oracle import OracleEncoder, double_conv_fwd


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
