"""Fused multi-tensor AdamW over the C ABI (`cmu_adamw_chunks`), configured like the reference's optimizer
(Pretraining/CM-UNet/configs/cmunet_config.py:76-91: AdamW betas (0.9, 0.95), weight decay 0.05, no decay on parameters
whose name contains 'bias' -- mmengine `paramwise_cfg.custom_keys`).  One launch updates every parameter."""
import torch

from . import ops
from ._lib import lib

_CHUNK = 1 << 16
NO_DECAY_KEYS = ('ln', 'bias', 'pos_embed', 'mask_token', 'cls_token')


class FusedAdamW:
    def __init__(self, named_params, lr=1.5e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05,
                 no_decay_keys=NO_DECAY_KEYS):
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.entries = []
        for name, p in named_params:
            if not p.requires_grad:
                continue
            decay = 0 if any(k in name for k in no_decay_keys) else 1
            self.entries.append((name, p, decay))
        self.state = {}
        self.step_count = 0
        self._table = None

    def zero_grad(self, set_to_none=True):
        for _, p, _ in self.entries:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def _build_table(self):
        rows, key = [], []
        for name, p, decay in self.entries:
            if p.grad is None:
                continue
            g = p.grad
            assert g.is_contiguous() and p.is_contiguous() and g.dtype == torch.float32 and p.dtype == torch.float32
            if name not in self.state:
                self.state[name] = (torch.zeros_like(p), torch.zeros_like(p))
            m, v = self.state[name]
            key.append((p.data_ptr(), g.data_ptr()))
            n = p.numel()
            for off in range(0, n, _CHUNK):
                rows.append((p.data_ptr() + 4 * off, g.data_ptr() + 4 * off, m.data_ptr() + 4 * off,
                             v.data_ptr() + 4 * off, min(_CHUNK, n - off), decay))
        return tuple(key), rows

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for _, p, _ in self.entries if p.grad is not None)
        if self._table is None or self._table[0] != key:
            k2, rows = self._build_table()
            dev = self.entries[0][1].device
            self._table = (k2, torch.tensor(rows, dtype=torch.int64, device=dev), len(rows))
        self.step_count += 1
        lib.cmu_adamw_chunks(self._table[1].data_ptr(), self._table[2], float(self.lr), float(self.betas[0]),
                             float(self.betas[1]), float(self.eps), float(self.weight_decay), self.step_count,
                             float(grad_scale), ops._stream())
