#!/usr/bin/env python
"""bench.py -- IB-INN hot-path throughput on B200 (contract: see the task statement / DESIGN.md).

  python bench.py --gpus N --steps K --warmup W            # ibinn_b200 (CUDA, one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)

A step = one CIFAR-10 IB-INN (beta = 1) training step (forward, IB loss, backward, clip 8.0, SGD) on a
synthetic batch of 128 images per GPU (BASELINE.json configs[1]).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = 'CIFAR-10 IB-INN beta_ramp beta_1.0000 training step (fwd + IB loss + bwd + clip 8.0 + SGD), fp32, synthetic 32x32x3'
PER_GPU_BATCH = 128
TRAIN_FLOP_PER_IMG = 1.175e9      # BASELINE.md section 2 (fwd + dgrad + wgrad, 2 x MAC)


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return dict(hbm=float(p['hbm_gbs']), bf16=float(p['bf16_tflops']), bf16_sustained=float(p.get('bf16_tflops_sustained', p['bf16_tflops'])),
                    source='measured')
    except Exception:
        return dict(hbm=6650., bf16=1590., bf16_sustained=1400., source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace('.', '').isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons, 'samples': len(sm)}


def cpu_reference(steps, warmup, budget_s=None):
    """The reference's training step restated by the oracle, fp32, all host threads."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import cpu_train_step as C
    import ibinn_oracle as O
    tr = C.CpuTrainer(O.Cfg(), PER_GPU_BATCH, beta=1.0)
    if budget_s is not None:
        t0 = time.perf_counter()
        tr.step()
        first = time.perf_counter() - t0
        steps = max(2, min(steps, int(budget_s / max(first, 1e-3))))
        warmup = 0
    sec = C.time_steps(tr, steps, warmup)
    return dict(value=PER_GPU_BATCH / sec, sec_per_step=sec, cores=tr.threads, steps=steps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-out', default=None, help='write the per-kernel-class device-time table here (json)')
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    warmup = max(a.warmup, 3)

    if a.impl == 'reference':
        if rank != 0:
            return
        r = cpu_reference(a.steps, warmup)
        line = {'impl': 'reference', 'metric': 'train_images_per_sec', 'value': r['value'], 'unit': 'img/s', 'n_gpus': a.gpus,
                'steps': a.steps, 'warmup': warmup, 'ms_per_step': r['sec_per_step'] * 1e3, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
                'config': {'workload': WORKLOAD, 'per_gpu_batch': PER_GPU_BATCH, 'global_batch': PER_GPU_BATCH,
                           'note': 'reference algorithm (oracle port of the pure-PyTorch reference) on host CPU cores; one step = one full batch of 128'},
                'cpu_baseline': {'value': r['value'], 'unit': 'img/s', 'cores': r['cores'], 'kind': 'port',
                                 'sample': '%d full training steps of batch %d' % (a.steps, PER_GPU_BATCH)},
                'e2e': {'value': r['value'], 'unit': 'img/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line), flush=True)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ROOT, 'ibinn_b200', 'csrc', 'libibinn_b200.so')):
        ge.build()
    from ibinn_b200 import _lib, config
    from ibinn_b200.engine import TrainStep
    from ibinn_b200.model import GenerativeClassifier
    if not torch.cuda.is_available():
        raise SystemExit('bench.py (impl ours) needs a B200: there is no CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'          # NCCL prints its version banner on stdout: keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(0)
    np.random.seed(0)
    args = config.make_args()                      # the reference's default CIFAR-10 config, beta_IB = 1.0
    model = GenerativeClassifier(args).to(dev)
    model.train()
    B = PER_GPU_BATCH
    g = torch.Generator().manual_seed(1 + rank)
    xh = torch.randn(B, 3, 32, 32, generator=g).pin_memory()
    lab = torch.randint(0, 10, (B,), generator=g)
    yh = (torch.zeros(B, 10).scatter_(1, lab.view(-1, 1), 1.) * 0.98 + 0.002).pin_memory()
    step = TrainStep(model, batch_size=B, beta=1.0, use_graph=not a.no_graph)
    step.load(xh, yh)

    # ---- per-kernel-class device times: one eager step bracketed with CUDA events (after a warm step)
    side, step.side_streams = step.side_streams, None      # serialised on one stream: every launch gets its own event pair
    step._eager()
    torch.cuda.synchronize()
    _lib.profile(True)
    n0 = _lib.lib().ibinn_launch_count()
    step._eager()
    prof = _lib.profile(False)
    step.side_streams = side
    table = prof.summary()
    eager_launches = int(_lib.lib().ibinn_launch_count() - n0)
    total_ms = sum(t for t, _ in table.values())

    # ---- timed region: K steps, inputs resident in HBM
    for _ in range(warmup):
        step.run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
        step.run()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of the logged scalars, every step
    res_h = torch.zeros(4).pin_memory()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    f0.record()
    for _ in range(a.steps):
        out = step.step(xh, yh)
        res_h.copy_(torch.stack([out['L_x_tr'], out['L_y_tr'], out['acc_tr'], out['grad_norm']]))   # synchronous read, like train.py:120-121
    f1.record()
    torch.cuda.synchronize()
    ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    loss_ok = bool(np.isfinite(res_h.numpy()).all())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    ms_step = ms / a.steps
    value = world * B / (ms_step * 1e-3)
    # ---- roofline of the dominant kernel.  The profiler keys carry the launch shapes
    # ('ibinn_conv_forward[k3 s1 N 32->32 @16x16 pre2 epi1]'), so the algorithmic FLOPs (2 x MAC, counted once although the
    # 3xTF32 arithmetic issues 2-3 MMAs per product) are exact per launch; the duration is the CUDA-event time of the same launches.
    import re

    def tag_flops(key):
        m = re.search(r'\[k(\d) s(\d) (?:[NT] )?(\d+)->(\d+) @(\d+)x(\d+)', key)
        if not m:
            return 0.
        k, _, ci, co, h, w = (int(v) for v in m.groups())
        return 2. * B * h * w * ci * co * k * k
    fam = {'conv_tc_kernel (subnet conv forward + dgrad)': lambda k: k.startswith('ibinn_conv_forward[') and ' s1 ' in k,
           'wgrad_tc_kernel (subnet conv weight gradient)': lambda k: k.startswith('ibinn_conv_wgrad[') and ' s1 ' in k}
    stats = {}
    for name, pred in fam.items():
        keys = [k for k in table if pred(k)]
        stats[name] = dict(ms=sum(table[k][0] for k in keys), n=sum(table[k][1] for k in keys), flops=sum(tag_flops(k) * table[k][1] for k in keys))
    top = max(stats, key=lambda n: stats[n]['ms'])
    st = stats[top]
    ach = st['flops'] / (st['ms'] * 1e-3) / 1e12 if st['ms'] > 0 else 0.
    traffic = None
    try:
        ncu = json.load(open(os.path.join(ROOT, 'profiles', 'r01_ncu_full.json')))
        traffic = ncu[top.split(' ')[0]]['dram_bytes_per_launch']
    except Exception:
        pass
    # HBM-bound kernel families at the training batch: algorithmic bytes / CUDA-event time of the same
    # launches.  At batch 128 a launch moves 3-8 MB (L2-resident), so these are latency-bound; scripts/hbm_ops_bench.py measures the
    # same kernels at the scoring batch (profiles/r01_hbm_ops.json).
    n_param = sum(q.numel() for q in model.parameters() if q.requires_grad)
    hbm_bytes = {'ibinn_aio_coupling_forward': 8 * 32768 * B + 48 * 36864 * B,      # 8 blocks of 32 768 B + 48 of 36 864 B per image
                 'ibinn_aio_coupling_backward': 8 * 53248 * B + 48 * 61440 * B,
                 'ibinn_sgd_step': 20. * n_param, 'ibinn_grad_norm_clip': 4. * n_param}
    hbm_fam = {}
    for k, nbytes in hbm_bytes.items():
        if k in table and table[k][0] > 0:
            gbs = nbytes / (table[k][0] * 1e-3) / 1e9
            hbm_fam[k] = {'ms_per_step': table[k][0], 'launches': table[k][1], 'bytes_per_step': nbytes, 'gbs': gbs, 'frac_of_hbm_peak': gbs / pk['hbm']}
    roofline = {'bound': 'tensor', 'kernel': top, 'hbm_families': hbm_fam, 'launches_per_step': st['n'], 'avg_launch_us': st['ms'] * 1e3 / max(st['n'], 1),
                'flops_per_launch': st['flops'] / max(st['n'], 1), 'achieved': ach, 'peak': pk['bf16_sustained'], 'unit': 'TFLOP/s',
                'frac': ach / pk['bf16_sustained'], 'traffic': traffic,
                'peak_source': pk['source'] + ' bf16 dense sustained (MEASURED_PEAKS.json has no TF32 figure; nominal TF32 dense is half of it)',
                'share_of_step': st['ms'] / total_ms if total_ms else None,
                'families': {n: {'ms_per_step': v['ms'], 'launches': v['n'], 'tflops': (v['flops'] / (v['ms'] * 1e-3) / 1e12 if v['ms'] > 0 else 0.)}
                             for n, v in stats.items()},
                'note': 'fp32-accurate 3xTF32 on tcgen05; channel widths are 16/32/64, so one MMA uses N <= 128 of 256 columns and the kernels are '
                        'latency- rather than throughput-bound (DESIGN.md section 7); time = CUDA-event durations of these launches in one eager '
                        'step, traffic = ncu dram bytes per launch (profiles/)'}
    line = {'metric': 'train_images_per_sec', 'value': value, 'unit': 'img/s', 'n_gpus': world, 'steps': a.steps, 'warmup': warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'per_gpu_batch': B, 'global_batch': B * world, 'parallelism': 'dp%d' % world,
                       'cuda_graph': step.use_graph, 'dropout': 'as shipped (Dropout2d 0.25 stage 2, Dropout 0.5 fc)',
                       'l2': 'no explicit flush: one step streams %.0f MB of saved activations + 3 x 45 MB of parameter state, larger than the 126 MB L2'
                             % (torch.cuda.max_memory_allocated() / 2 ** 20)},
            'clocks': clocks,
            'e2e': {'value': world * B / (ms_e2e / a.steps * 1e-3), 'unit': 'img/s', 'h2d_bytes_per_step': int(xh.numel() * 4 + yh.numel() * 4),
                    'd2h_bytes_per_step': 16},
            'gpu_launches': (step.launches if step.launches is not None else eager_launches) * a.steps,
            'gpu_launches_per_step': step.launches if step.launches is not None else eager_launches,
            'roofline': roofline, 'finite': loss_ok}
    if not a.no_cpu_baseline and world == 1:
        r = cpu_reference(10, 1, budget_s=15.)
        line['cpu_baseline'] = {'value': r['value'], 'unit': 'img/s', 'cores': r['cores'], 'kind': 'port',
                                'sample': '%d full training steps of batch %d (oracle fp32 port of the reference step)' % (r['steps'], B)}
    if a.profile_out:
        rows = sorted(((k, t, n) for k, (t, n) in table.items()), key=lambda r: -r[1])
        json.dump({'total_ms': total_ms, 'launch_calls': eager_launches, 'rows': rows}, open(a.profile_out, 'w'), indent=1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
