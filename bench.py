#!/usr/bin/env python
"""bench.py — CM-UNet pretraining throughput (BASELINE.json metric: images/sec @512² on 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W    (the reference's CPU path = pinned oracle port)

A "step" is one full pretraining iteration of BASELINE.json configs[1] on every GPU: device patch-mask generation ->
online + target encoder -> two decoders -> projector / InfoNCE + masked MSE -> backward (-> gradient all-reduce over
NCCL when N>1, overlapped by DDP) -> fused AdamW -> EMA of the target networks.  Per-GPU batch 64 of synthetic
1x512x512 images (weak scaling), random-init weights of the real architecture, bf16 activations / fp32 accumulate.
`value` times K steps with inputs resident in HBM (CUDA events, max over ranks); `e2e` repeats the K steps through the
public module API with pinned-host inputs copied H2D and the losses read back D2H inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_IMAGE_512 = 2040.9e9          # algorithmic FLOPs of one pretraining step per image @512²
METRIC = 'CM-UNet pretrain images/sec @512^2'


def algorithmic_flops_per_image(S):
    return FLOPS_PER_IMAGE_512 * (S / 512.0) ** 2


K1_DRAM_BYTES_PER_LAUNCH = 1.915e9   # measured, see roofline.traffic_note


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'bf16_tflops': d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1400.0)), 'hbm_gbs': d.get('hbm_gbs', 6650.0),
                'source': 'MEASURED_PEAKS.json (sustained bf16; kernels are timed inside a long step)'}
    return {'bf16_tflops': 1400.0, 'hbm_gbs': 6650.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._stop_evt = index, [], set(), threading.Event()
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap', nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: 'hw_power_brake'}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.1)
        except Exception as e:  # NVML unavailable: report that instead of inventing numbers
            self.reasons.add(f'nvml_unavailable:{type(e).__name__}')

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


# ---------------------------------------------------------------------------------------------------- reference arm
def cpu_oracle_step_time(B, S, steps, warmup, threads):
    """fwd+bwd of the pinned oracle port (oracle/cmunet_oracle.py == the reference's CM_UNet, fp32) on the host cores."""
    import torch
    from oracle import cmunet_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(60)
    m = O.OracleCMUNet(img_size=S, np_seed=60)
    m.init_weights()
    m.train()
    img, img_t = O.synthetic_batch(B, S, 1)
    times = []
    for it in range(warmup + steps):
        t0 = time.time()
        for p in m.parameters():
            p.grad = None
        out = m(img, mode='loss', img_t=img_t)
        (out['loss_ct'] + out['loss_rc']).backward()
        m.momentum_update()
        dt = time.time() - t0
        if it >= warmup:
            times.append(dt)
    return sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    B, S = 2, args.size
    # bounded sample: the oracle at B=2, S=512 costs ~10-30 s/step on 8-32 host cores.  K and W are honoured up to
    # 8 timed + 2 warm-up steps so that the whole run stays within a few minutes (the cap is reported in `config`).
    warm = max(1, min(args.warmup, 2))
    steps = max(1, min(args.steps, 8))
    t = cpu_oracle_step_time(B, S, steps, warm, cores)
    val = B / t
    line = {'metric': METRIC, 'value': val, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': f'configs[1] CM-UNet pretraining step (mask+fwd+bwd+EMA) at {S}x{S}, CPU, bounded sample',
                       'per_step_batch': B, 'img_size': S, 'timed_steps_actually_run': steps, 'warmup_actually_run': warm},
            'cpu_baseline': {'value': val, 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                             'sample': f'{steps} timed fwd+bwd+EMA steps of the oracle port at B={B}, S={S} (fp32, torch CPU)'},
            'e2e': {'value': val, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    import contrastive_masked_unet_b200 as C
    from contrastive_masked_unet_b200 import ops
    from contrastive_masked_unet_b200.optim import FusedAdamW
    C.lib.cmu_device_check()

    B, S = args.batch, args.size
    torch.manual_seed(60)
    np.random.seed(60 + rank)                      # cmunet_config.py:133 diff_rank_seed
    model = C.build(C.cmunet_config(S))
    model.init_weights()
    model = model.to(dev).train()
    core = model
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        model = DDP(core, device_ids=[local_rank], broadcast_buffers=False, gradient_as_bucket_view=True,
                    bucket_cap_mb=64)
    opt = FusedAdamW(core.named_parameters(), lr=1.5e-4)

    # synthetic data: a few distinct batches (each 2 x 67 MB at B=64, S=512; activations per step are tens of GB, far
    # beyond the 126 MB L2, so nothing is served from cache between steps)
    g = torch.Generator().manual_seed(1 + rank)
    n_pool = 2
    host_img = [torch.randn(B, S, S, generator=g).pin_memory() for _ in range(n_pool)]
    host_img_t = [(h + 0.1 * torch.randn(B, S, S, generator=g)).pin_memory() for h in host_img]
    dev_img = [h.to(dev) for h in host_img]
    dev_img_t = [h.to(dev) for h in host_img_t]

    def step(img, img_t):
        opt.zero_grad(set_to_none=True)
        out = model(img, mode='loss', img_t=img_t)
        (out['loss_ct'] + out['loss_rc']).backward()
        opt.step()
        core.momentum_update()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(dev_img[i % n_pool], dev_img_t[i % n_pool])
    barrier()

    # ---- timed region 1: device-resident inputs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = C.lib.cmu_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        out = step(dev_img[i % n_pool], dev_img_t[i % n_pool])
    e1.record()
    barrier()
    launches = C.lib.cmu_launch_count() - n0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms)
    loss_ct, loss_rc = float(out['loss_ct']), float(out['loss_rc'])

    # ---- timed region 2: end to end through the public API, pinned host inputs, losses read back
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(args.steps):
        img = host_img[i % n_pool].to(dev, non_blocking=True)
        img_t = host_img_t[i % n_pool].to(dev, non_blocking=True)
        o = step(img, img_t)
        host_losses = torch.stack([o['loss_ct'].detach(), o['loss_rc'].detach()]).cpu()   # D2H read of the step's result
    e3.record()
    barrier()
    t2 = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    ms_e2e = float(t2)

    # ---- roofline pass: the same K steps with CUDA events around every tensor-core launch.  The production schedule runs
    # the target encoder and the two decoders on three streams; there an event pair also spans the time a kernel waits for
    # SMs held by another stream's persistent kernel, so per-kernel durations are taken with the branches serialised on
    # one stream (same kernels, same order, same clocks / power state).
    timer = ops.KernelTimer()
    core.multi_stream = False
    step(dev_img[0], dev_img_t[0])
    barrier()
    ops.PROFILER = timer
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for i in range(args.steps):
        step(dev_img[i % n_pool], dev_img_t[i % n_pool])
    e5.record()
    barrier()
    ops.PROFILER = None
    core.multi_stream = os.environ.get('CMU_SINGLE_STREAM') != '1'
    ms_prof = e4.elapsed_time(e5)

    if rank == 0:
        imgs = B * world * args.steps
        value = imgs / (ms / 1e3)
        peaks = measured_peaks()
        summ = timer.summarize()
        tc_ms = sum(v[1] for v in summ.values())
        tc_fl = sum(v[2] for v in summ.values())
        # dominant kernel family: the tcgen05 implicit-GEMM convolution (K1: fprop + dgrad of every 3x3 conv)
        k1 = [summ.get('conv3x3_fprop', (0, 0.0, 0.0)), summ.get('conv3x3_dgrad', (0, 0.0, 0.0))]
        k1_ms, k1_fl, k1_n = sum(v[1] for v in k1), sum(v[2] for v in k1), sum(v[0] for v in k1)
        achieved = (k1_fl / (k1_ms / 1e3)) / 1e12 if k1_ms > 0 else 0.0
        roofline = {'bound': 'tensor', 'kernel': 'k1_kernel (tcgen05 implicit-GEMM conv3x3 fprop+dgrad)',
                    'achieved': achieved, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                    'frac': achieved / peaks['bf16_tflops'],
                    'traffic': K1_DRAM_BYTES_PER_LAUNCH if (B, S) == (64, 512) else None, 'peak_source': peaks['source'],
                    'traffic_note': 'bytes per K1 launch, dram__bytes_read.sum + dram__bytes_write.sum averaged over the 60 '
                                    'k1_pair launches of one step of this command (ncu, profiles/r1_k2_dram_traffic.md, '
                                    'profiles/r1_step_dram_traffic.json); algorithmic bytes (every activation tile loaded '
                                    'once, every output stored once) are the same 1.9 GB: no wasted re-reads',
                    'launches': k1_n, 'avg_launch_ms': k1_ms / max(1, k1_n), 'share_of_step': k1_ms / ms_prof,
                    'measured': f'CUDA events around every launch in a second pass of {args.steps} steps on ONE stream '
                                f'({ms_prof / args.steps:.1f} ms/step; the 3-stream production schedule of the timed region '
                                f'takes {ms / args.steps:.1f} ms/step)',
                    'all_tensor_core_kernels': {k: {'launches': v[0], 'ms': round(v[1], 3),
                                                    'tflops': round(v[2] / (v[1] / 1e3) / 1e12, 1) if v[1] > 0 else None}
                                                for k, v in summ.items()},
                    'tensor_core_share_of_step': tc_ms / ms_prof,
                    'whole_step_algorithmic_tflops': algorithmic_flops_per_image(S) * value / 1e12,
                    'whole_step_frac_of_peak': algorithmic_flops_per_image(S) * value / 1e12 / peaks['bf16_tflops']}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            t = cpu_oracle_step_time(2, S, 1, 1 if S <= 256 else 0, cores)
            cpu = {'value': 2 / t, 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                   'sample': f'1 fwd+bwd+EMA step of the oracle port (fp32 torch CPU) at B=2, S={S}'}
        line = {'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
                'config': {'workload': 'configs[1]: CM-UNet pretraining step (patch mask + fwd + bwd + InfoNCE + masked MSE '
                                       '+ AdamW + EMA), per-GPU batch 64 of synthetic 1x512x512' if (B, S) == (64, 512) else
                                       f'CM-UNet pretraining step at per-GPU batch {B}, {S}x{S}',
                           'per_gpu_batch': B, 'global_batch': B * world, 'img_size': S, 'parallelism': f'dp{world}',
                           'l2': 'inputs 2x67 MB/step and >50 GB of activations per step: far larger than the 126 MB L2',
                           'loss_ct': loss_ct, 'loss_rc': loss_rc},
                'clocks': clocks,
                'e2e': {'value': imgs / (ms_e2e / 1e3), 'unit': 'images/s', 'h2d_bytes_per_step': 2 * B * S * S * 4,
                        'd2h_bytes_per_step': 8, 'ms_per_step': ms_e2e / args.steps},
                'gpu_launches': int(launches),
                'roofline': roofline}
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3                            # timing rule: at least 3 warm-up steps
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
