"""MoCo-v2 momentum-encoder UNet with a negative queue (BASELINE.json configs[3]):
drop-in for the math of Pretraining/MoCo/pl_bolts/models/self_supervised/moco/moco2_module.py:51-309 without the
Lightning shell (constructor keywords, buffers `queue` (D,K) / `queue_ptr` / `val_queue*`, `encoder_q` / `encoder_k`
state_dict keys are kept).

The N x (1+K) logits never reach HBM: the tcgen05 1x1 kernel (queue rows are the GEMM-M "pixels", the normalised
queries the weight matrix) applies exp((logit - 1) / T) in its epilogue and stores the unnormalised softmax terms
E[K][N] in bf16 together with per-CTA partial column sums (the denominators); a tiny kernel finishes loss, p(positive)
and the per-query scale, and dq = (E^T Queue) * scale runs on the tcgen05 row-reduction GEMM."""
import torch
import torch.nn as nn

from . import functional as Fn
from . import ops
from ._lib import CmuError, lib
from .functional import _world
from .modules import DoubleConv, DownBlock, concat_all_gather, register

BF16 = torch.bfloat16


class SpatialMeanFn(torch.autograd.Function):
    """torch.mean(x, dim=[2, 3]) on an act tensor (moco_data_module.py:65) -> (N, C) fp32."""

    @staticmethod
    def forward(ctx, x):
        a = Fn._nhwc(x)
        n, h, w, c = a.shape
        out = torch.empty(n, c, dtype=torch.float32, device=a.device)
        lib.cmu_spatial_mean(a.data_ptr(), out.data_ptr(), n, h * w, c, ops._stream())
        ctx.shape = (n, h, w, c)
        return out

    @staticmethod
    def backward(ctx, dout):
        n, h, w, c = ctx.shape
        dout = dout.contiguous().float()
        dx = torch.empty(n, h, w, c, dtype=BF16, device=dout.device)
        lib.cmu_spatial_mean_bwd(dout.data_ptr(), dx.data_ptr(), n, h * w, c, ops._stream())
        return Fn._nchw_view(dx)


class MocoLossFn(torch.autograd.Function):
    """loss = CE([q_hat.k, q_hat.Queue] / T, label 0), q_hat = normalize(q) (moco2_module.py:236-270,284).
    q: (N,D) raw query features; k: (N,D) normalised keys (no grad); queue_rows: (K,D) bf16, one negative per row."""

    @staticmethod
    def forward(ctx, q, k, queue_rows, temperature):
        q = q.contiguous().float()
        k = k.contiguous().float()
        n, d = q.shape
        kneg = queue_rows.shape[0]
        if n % 64 != 0 or d % 64 != 0 or queue_rows.dtype != BF16 or not queue_rows.is_contiguous():
            raise CmuError('MoCo head: batch and embedding dim must be multiples of 64, queue rows bf16 (K,D)')
        dev, st = q.device, ops._stream()
        qh16 = torch.empty(n, d, dtype=BF16, device=dev)
        qh = torch.empty_like(q)
        qnorm = torch.empty(n, device=dev)
        lpos = torch.empty(n, device=dev)
        lib.cmu_moco_prep(q.data_ptr(), k.data_ptr(), n, d, qh16.data_ptr(), qh.data_ptr(), qnorm.data_ptr(),
                          lpos.data_ptr(), st)
        wv = 64 if kneg % 64 == 0 else 8 if kneg % 8 == 0 else 1                     # any (H,W) factorisation works
        # E[k][n] = exp((Queue_k . q_hat_n - 1) / T) straight from the tcgen05 epilogue (bf16) + partial column sums:
        # the logits never reach HBM; E is the operand of the backward GEMM
        e, part, grid, bn = ops.conv1x1_fprop_exp(queue_rows.view(1, kneg // wv, wv, d), qh16, 1.0 / float(temperature),
                                                  name='moco_logits')
        need = ctx.needs_input_grad[0]
        rows = torch.empty(n, device=dev)
        loss = torch.empty(1, device=dev)
        ppos = torch.empty(n, device=dev)
        rscale = torch.empty(n, device=dev)
        lib.cmu_moco_finish(part.data_ptr(), grid, bn, n, lpos.data_ptr(), float(temperature), rows.data_ptr(),
                            loss.data_ptr(), ppos.data_ptr(), rscale.data_ptr(), st)
        if need:
            dq_neg = ops.gemm_tn(e.view(kneg, n), queue_rows)                      # (N,D) = E^T Queue (unnormalised)
            dq = torch.empty_like(q)
            lib.cmu_moco_dq_scaled(dq_neg.data_ptr(), rscale.data_ptr(), k.data_ptr(), qh.data_ptr(), qnorm.data_ptr(),
                                   ppos.data_ptr(), n, d, float(temperature), dq.data_ptr(), st)
            ctx.save_for_backward(dq)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dq,) = ctx.saved_tensors
        return dq * g, None, None, None


@register
class MocoUNetEncoder(nn.Module):
    """Pretraining/MoCo/.../moco_data_module.py:47-66: UNet encoder + global average pool -> (N, 1024)."""

    def __init__(self, out_classes=2, up_sample_mode='conv_transpose'):
        super().__init__()
        self.up_sample_mode = up_sample_mode
        self.down_conv1 = DownBlock(1, 64)
        self.down_conv2 = DownBlock(64, 128)
        self.down_conv3 = DownBlock(128, 256)
        self.down_conv4 = DownBlock(256, 512)
        self.double_conv = DoubleConv(512, 1024)

    def forward(self, x):
        if x.dim() == 3:
            x = x.unsqueeze(1)
        x, _ = self.down_conv1(x, want_skip=False)      # the skip tensors have no consumer here: never stored
        x, _ = self.down_conv2(x, want_skip=False)
        x, _ = self.down_conv3(x, want_skip=False)
        x, _ = self.down_conv4(x, want_skip=False)
        x = self.double_conv(x)
        return SpatialMeanFn.apply(Fn.to_act(x))


@register
class Moco_v2(nn.Module):
    def __init__(self, base_encoder=None, emb_dim=1024, num_negatives=65536, encoder_momentum=0.999,
                 softmax_temperature=0.07, learning_rate=0.03, momentum=0.9, weight_decay=1e-4, data_dir='./',
                 batch_size=256, use_mlp=False, num_workers=8, *args, **kwargs):
        super().__init__()
        if use_mlp:
            raise NotImplementedError('use_mlp needs an fc layer the UNet encoder does not have (moco2_module.py:110-113)')
        import copy
        self.hparams = dict(emb_dim=emb_dim, num_negatives=num_negatives, encoder_momentum=encoder_momentum,
                            softmax_temperature=softmax_temperature, learning_rate=learning_rate, momentum=momentum,
                            weight_decay=weight_decay, batch_size=batch_size)
        base = base_encoder if isinstance(base_encoder, nn.Module) else MocoUNetEncoder()
        self.encoder_q = copy.deepcopy(base)                         # :140-146
        self.encoder_k = copy.deepcopy(base)
        for pq, pk in zip(self.encoder_q.parameters(), self.encoder_k.parameters()):
            pk.data.copy_(pq.data)
            pk.requires_grad = False
        self.register_buffer('queue', nn.functional.normalize(torch.randn(emb_dim, num_negatives), dim=0))   # :120-121
        self.register_buffer('queue_ptr', torch.zeros(1, dtype=torch.long))
        self.register_buffer('val_queue', nn.functional.normalize(torch.randn(emb_dim, num_negatives), dim=0))
        self.register_buffer('val_queue_ptr', torch.zeros(1, dtype=torch.long))
        self._rows = None            # bf16 (K, D) working copy of `queue` (one negative per row)
        self._ptr = 0
        self._sig = None
        self.shuffle_bn = True       # batch shuffle across ranks whenever a process group with world > 1 is active
        self._ema_table = None

    # ------------------------------------------------------------------ queue
    def _queue_sig(self):
        # in-place writes (load_state_dict's copy_, user code) bump the version counters; no device sync involved
        return (self.queue.data_ptr(), self.queue._version, self.queue_ptr._version)

    def _queue_rows(self):
        """bf16 (K, D) working copy of `queue` + host copy of `queue_ptr`, rebuilt whenever the buffers were written by
        anything but this module's own enqueue (checkpoint resume, `.to()`, code written against the reference API)."""
        if self._rows is None or self._rows.device != self.queue.device or self._sig != self._queue_sig():
            self._rows = self.queue.t().contiguous().to(BF16)
            self._ptr = int(self.queue_ptr)
            self._sig = self._queue_sig()
        return self._rows

    @torch.no_grad()
    def _dequeue_and_enqueue(self, keys):
        """moco2_module.py:160-175 (keys all-gathered across ranks first)."""
        keys = concat_all_gather(keys.contiguous().float())
        n, d = keys.shape
        kneg = self.hparams['num_negatives']
        assert kneg % n == 0
        rows = self._queue_rows()
        lib.cmu_queue_enqueue(keys.data_ptr(), n, d, kneg, self._ptr, rows.data_ptr(), self.queue.data_ptr(), ops._stream())
        self._ptr = (self._ptr + n) % kneg
        self.queue_ptr[0] = self._ptr
        self._sig = self._queue_sig()        # our own writes (raw-pointer kernel + the pointer store) keep the cache valid

    # ------------------------------------------------------------------ shuffle-BN (moco2_module.py:177-222)
    @torch.no_grad()
    def _batch_shuffle_ddp(self, x):
        """All-gather the key images, draw ONE permutation on rank 0 (CPU torch RNG, like `torch.randperm(n).cuda()` in
        the reference), broadcast it, and keep this rank's slice: the per-GPU BatchNorm of `encoder_k` then sees a
        different sample set than `encoder_q` (the information leak the shuffle exists to prevent)."""
        import torch.distributed as dist
        n_this = x.shape[0]
        x_all = concat_all_gather(x.contiguous())
        n_all = x_all.shape[0]
        idx_shuffle = torch.randperm(n_all).to(x.device)
        dist.broadcast(idx_shuffle, src=0)
        idx_unshuffle = torch.argsort(idx_shuffle)
        idx_this = idx_shuffle.view(n_all // n_this, -1)[dist.get_rank()]
        return x_all[idx_this], idx_unshuffle

    @torch.no_grad()
    def _batch_unshuffle_ddp(self, x, idx_unshuffle):
        import torch.distributed as dist
        n_this = x.shape[0]
        x_all = concat_all_gather(x.contiguous())
        idx_this = idx_unshuffle.view(x_all.shape[0] // n_this, -1)[dist.get_rank()]
        return x_all[idx_this]

    @torch.no_grad()
    def _momentum_update_key_encoder(self):
        """moco2_module.py:153-158 as one multi-tensor launch."""
        pairs = list(zip(self.encoder_q.parameters(), self.encoder_k.parameters()))
        key = tuple((pk.data_ptr(), pq.data_ptr(), pk.numel()) for pq, pk in pairs)
        if self._ema_table is None or self._ema_table[0] != key:
            rows = []
            for dst, src, n in key:
                for off in range(0, n, 1 << 16):
                    rows.append((dst + 4 * off, src + 4 * off, min(1 << 16, n - off)))
            self._ema_table = (key, torch.tensor(rows, dtype=torch.int64, device=pairs[0][0].device), len(rows))
        lib.cmu_ema_chunks(self._ema_table[1].data_ptr(), self._ema_table[2], float(self.hparams['encoder_momentum']),
                           ops._stream())

    # ------------------------------------------------------------------ step
    def forward(self, img_q, img_k, queue=None, encoder_q=None):
        """-> (loss, k, q).  Unlike moco2_module.py:224-270 the (N, 1+K) logits tensor is not returned: it never exists.
        encoder_q: optional wrapper of `self.encoder_q` to call instead (DistributedDataParallel around the only
        sub-module that has gradients)."""
        ops._need_cuda(img_q, img_k)
        q = (encoder_q if encoder_q is not None else self.encoder_q)(img_q)
        with torch.no_grad():
            shuffle = self.shuffle_bn and _world() > 1            # moco2_module.py:246-255 (`_use_ddp_or_ddp2`)
            if shuffle:
                img_k, idx_unshuffle = self._batch_shuffle_ddp(img_k)
            k = ops.l2_normalize_rows(self.encoder_k(img_k).contiguous().float())
            if shuffle:
                k = self._batch_unshuffle_ddp(k, idx_unshuffle)
        loss = MocoLossFn.apply(q, k, self._queue_rows(), self.hparams['softmax_temperature'])
        return loss, k, q

    def training_step(self, img_q, img_k, encoder_q=None):
        """moco2_module.py:287-309 without the Lightning plumbing: EMA of the key encoder, loss, dequeue/enqueue."""
        ops._need_cuda(img_q, img_k)
        self._momentum_update_key_encoder()
        loss, k, _ = self.forward(img_q, img_k, encoder_q=encoder_q)
        self._dequeue_and_enqueue(k)
        return loss

    def __getstate__(self):
        d = dict(self.__dict__)
        d['_rows'] = None
        d['_sig'] = None
        d['_ema_table'] = None
        return d
