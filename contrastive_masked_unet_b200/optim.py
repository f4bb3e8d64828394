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
    def step(self, grad_scale=1.0, scaler=None):
        """scaler: a `DynamicLossScaler` -- the gradients are unscaled by its current scale, the step is skipped on the
        device when any gradient is non-finite and the scale is updated (GradScaler.step + update), all without a host
        synchronisation; `step_count` then counts attempted steps, the bias correction uses the device-side count of
        successful ones."""
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for _, p, _ in self.entries if p.grad is not None)
        if self._table is None or self._table[0] != key:
            k2, rows = self._build_table()
            dev = self.entries[0][1].device
            self._table = (k2, torch.tensor(rows, dtype=torch.int64, device=dev), len(rows))
        self.step_count += 1
        if scaler is not None:
            lib.cmu_adamw_chunks_amp(self._table[1].data_ptr(), self._table[2], float(self.lr), float(self.betas[0]),
                                     float(self.betas[1]), float(self.eps), float(self.weight_decay),
                                     scaler.state.data_ptr(), float(scaler.growth_factor), float(scaler.backoff_factor),
                                     int(scaler.growth_interval), ops._stream())
            return
        lib.cmu_adamw_chunks(self._table[1].data_ptr(), self._table[2], float(self.lr), float(self.betas[0]),
                             float(self.betas[1]), float(self.eps), float(self.weight_decay), self.step_count,
                             float(grad_scale), ops._stream())


class DynamicLossScaler:
    """Device-resident dynamic loss scale with the semantics of `torch.amp.GradScaler` / mmengine's
    `AmpOptimWrapper(loss_scale='dynamic')` (cmunet_config.py:76-78): `scale(loss)` multiplies by the current scale,
    `FusedAdamW.step(scaler=...)` unscales, skips the update when a gradient overflowed, halves the scale after an
    overflow and doubles it after `growth_interval` clean steps.  bf16 training does not need it (same exponent range as
    fp32); it exists so that a reference config using fp16 AMP maps one to one."""

    def __init__(self, device='cuda', init_scale=2.0 ** 16, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, growth_interval
        self.state = torch.zeros(5, dtype=torch.int32, device=device)
        lib.cmu_amp_init(self.state.data_ptr(), float(init_scale), ops._stream())

    def scale(self, loss):
        return loss * self.state[0:1].view(torch.float32).reshape(())

    def get_scale(self):
        return float(self.state[0:1].view(torch.float32))

    def steps_taken(self):
        return int(self.state[2])

    def found_inf(self):
        return bool(int(self.state[1]))


class FusedSGD:
    """torch.optim.SGD(momentum, weight_decay) as one multi-tensor launch (`cmu_sgd_chunks`): the MoCo-v2 optimizer
    (MOCO/moco2_module.py configure_optimizers: lr 0.03, momentum 0.9, weight decay 1e-4)."""

    def __init__(self, named_params, lr=0.03, momentum=0.9, weight_decay=1e-4):
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.entries = [(n, p) for n, p in named_params if p.requires_grad]
        self.state = {}
        self.step_count = 0
        self._table = None

    def zero_grad(self, set_to_none=True):
        for _, p in self.entries:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for _, p in self.entries if p.grad is not None)
        if self._table is None or self._table[0] != key:
            rows = []
            for name, p in self.entries:
                if p.grad is None:
                    continue
                g = p.grad
                assert g.is_contiguous() and p.is_contiguous() and g.dtype == torch.float32 and p.dtype == torch.float32
                if name not in self.state:
                    self.state[name] = torch.zeros_like(p)
                buf, n = self.state[name], p.numel()
                for off in range(0, n, _CHUNK):
                    rows.append((p.data_ptr() + 4 * off, g.data_ptr() + 4 * off, buf.data_ptr() + 4 * off, min(_CHUNK, n - off)))
            self._table = (key, torch.tensor(rows, dtype=torch.int64, device=self.entries[0][1].device), len(rows))
        lib.cmu_sgd_chunks(self._table[1].data_ptr(), self._table[2], float(self.lr), float(self.momentum),
                           float(self.weight_decay), int(self.step_count == 0), ops._stream())
        self.step_count += 1
