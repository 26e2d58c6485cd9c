#!/usr/bin/env python
"""bench.py -- IB-INN hot-path throughput on B200 (contract: see the task statement / DESIGN.md section 4).

  python bench.py --gpus N --steps K --warmup W                  # ibinn_b200 (CUDA, one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...        # the reference's own files on the host CPU cores
  python bench.py --workload {cifar10_train|cifar100_train|mnist_fwd|cifar10_infer|cifar10_beta_sweep}

Default workload = BASELINE.json configs[1]: one CIFAR-10 IB-INN (beta = 1) training step (forward, IB loss, backward,
clip 8.0, SGD) on a synthetic batch of 128 images per GPU.  The other workloads are configs[2] (CIFAR-100, 100-class head),
configs[0] (MNIST forward + IB loss, batch 64), configs[3] (eval-mode scoring at batch 2048) and configs[4] (the beta ramp:
independent replicas, one process per GPU, no collective).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BETA_RAMP = [0.6002, 0.8435, 1.0000, 1.1855, 1.6662, 2.3419, 3.2915, 4.6261, 6.5019]   # reference gen_beta_ramp.sh:11-19

# name -> description, ini overrides, per-GPU batch, kind, algorithmic FLOPs per image (2 x MAC; BASELINE.md section 2 / SURVEY 8d)
WORKLOADS = {
    'cifar10_train': dict(
        text='CIFAR-10 IB-INN beta_ramp beta_1.0000 training step (fwd + IB loss + bwd + clip 8.0 + SGD), fp32, synthetic 32x32x3',
        overrides={}, batch=128, kind='train', flop_per_img=1.175e9, classes=10),
    'cifar100_train': dict(
        text='CIFAR-100 IB-INN (100-class GMM head) training step (fwd + IB loss + bwd + clip 8.0 + SGD), fp32, synthetic 32x32x3, batch-sharded',
        overrides={'data': {'dataset': 'CIFAR100'}}, batch=128, kind='train', flop_per_img=1.175e9 + 3 * 2 * 3072 * 100, classes=100),
    'mnist_fwd': dict(
        text='MNIST IB-INN (experiments_configs/mnist, beta=1) forward + IB loss, batch 64, fp32, synthetic 28x28',
        overrides={'data': {'dataset': 'MNIST', 'label_smoothing': '0.01'}, 'model': {'act_norm': '0.80'}}, batch=64, kind='fwd',
        flop_per_img=0.2064e9, classes=10),
    'cifar10_infer': dict(
        text='CIFAR-10 IB-INN inference log-likelihood / OoD scoring (eval mode, no backward), fp32, synthetic 32x32x3, batch 2048',
        overrides={}, batch=2048, kind='infer', flop_per_img=0.3916e9, classes=10),
    'cifar10_beta_sweep': dict(
        text='CIFAR-10 beta sweep (9 betas 0.6-6.5 of beta_ramp) independent training replicas, fp32, synthetic 32x32x3, batch 128 each',
        overrides={}, batch=128, kind='sweep', flop_per_img=1.175e9, classes=10),
}
METRIC = {'train': 'train_images_per_sec', 'sweep': 'train_images_per_sec', 'fwd': 'forward_images_per_sec', 'infer': 'inference_images_per_sec'}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return dict(hbm=float(p['hbm_gbs']), bf16=float(p['bf16_tflops']), bf16_sustained=float(p.get('bf16_tflops_sustained', p['bf16_tflops'])),
                    source='measured')
    except Exception:
        return dict(hbm=6650., bf16=1590., bf16_sustained=1400., source='fallback')


def measure_tf32_peak(torch, dev, seconds=1.0):
    """Dense TF32 tensor throughput of this GPU: cuBLAS `torch.matmul` 8192^3 with TF32 allowed (2 N^3 FLOP), best of 10 (burst)
    and back to back for `seconds` (sustained).  MEASURED_PEAKS.json only records bf16."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a, b = torch.randn(n, n, device=dev), torch.randn(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(5, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        sustained = e0.elapsed_time(e1) / reps
        del a, b
        return dict(burst=2. * n ** 3 / (best * 1e-3) / 1e12, sustained=2. * n ** 3 / (sustained * 1e-3) / 1e12)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


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


def config_of(wname, batch, world):
    """identical for both arms (the driver compares the two dicts)"""
    w = WORKLOADS[wname]
    n_cfg = len(BETA_RAMP) if w['kind'] == 'sweep' else 1
    return {'workload': w['text'], 'per_gpu_batch': batch, 'global_batch': batch * world,
            'parallelism': ('replicas%d' if w['kind'] == 'sweep' else 'dp%d') % world, 'configs': n_cfg}


# --------------------------------------------------------------------------------------------- reference arm (host CPU)

def reference_run(wname, batch, device, steps, warmup, budget_s=None):
    """the reference's own files (baseline/_ref, see baseline/ref_arm.py) on `device`; falls back to the oracle port when the
    files did not travel.  Returns dict(value img/s, sec_per_step, steps, cores, kind)."""
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_arm
    w = WORKLOADS[wname]
    if ref_arm.available() or ref_arm.install():
        r = ref_arm.RefRunner(wname, device, batch=batch)
        sec, n = r.time(steps, warmup, budget_s)
        return dict(value=batch / sec, sec_per_step=sec, steps=n, cores=r.threads, kind='reference')
    if device != 'cpu':
        raise RuntimeError('baseline/_ref is missing: no stock-PyTorch GPU baseline')
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import cpu_train_step as C
    import ibinn_oracle as O
    from ibinn_b200 import config
    cfg = O.cfg_from_args(config.make_args(overrides=w['overrides']))
    tr = C.CpuTrainer(cfg, batch, beta=1.0)
    fn = 'step' if w['kind'] in ('train', 'sweep') else 'forward_only'
    if budget_s is not None:
        t0 = time.perf_counter()
        getattr(tr, fn)()
        first = time.perf_counter() - t0
        steps, warmup = max(2, min(steps, int(budget_s / max(first, 1e-3)))), 0
    sec = C.time_steps(tr, steps, warmup, fn)
    return dict(value=batch / sec, sec_per_step=sec, steps=steps, cores=tr.threads, kind='port')


def reference_main(a, rank, world, warmup):
    if rank != 0:
        return
    w = WORKLOADS[a.workload]
    batch = a.batch or w['batch']
    dev = a.ref_device
    # every step is one full batch of the workload; the number of steps is bounded so the run ends within a few minutes
    r = reference_run(a.workload, batch, dev, a.steps, warmup, budget_s=a.budget if a.budget > 0 else None)
    sample = '%d full %s of batch %d on %s' % (r['steps'], 'training steps' if w['kind'] in ('train', 'sweep') else 'forward passes', batch,
                                               'the host CPU cores' if dev == 'cpu' else 'cuda:0 with stock PyTorch (cuDNN / cuBLAS, TF32 off)')
    line = {'impl': 'reference', 'metric': METRIC[w['kind']], 'value': r['value'], 'unit': 'img/s', 'n_gpus': a.gpus,
            'steps': r['steps'], 'warmup': warmup if a.budget <= 0 else min(warmup, 1), 'ms_per_step': r['sec_per_step'] * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic', 'config': config_of(a.workload, batch, world),
            'device': dev,
            'cpu_baseline': {'value': r['value'], 'unit': 'img/s', 'cores': r['cores'], 'kind': r['kind'], 'sample': sample},
            'e2e': {'value': r['value'], 'unit': 'img/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'note': 'the unmodified reference files (baseline/_ref) through oracle/ref_shims; one process, rank 0 only' if r['kind'] == 'reference'
                    else 'oracle port of the reference (baseline/_ref did not travel)'}
    print(json.dumps(line), flush=True)


def sub_reference(wname, batch, device, steps, warmup, budget):
    """run the reference arm in a fresh process (its shims monkey-patch torch) and parse its JSON line"""
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--workload', wname, '--batch', str(batch),
           '--ref-device', device, '--steps', str(steps), '--warmup', str(warmup), '--budget', str(budget)]
    env = dict(os.environ)
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                return json.loads(ln)
        return {'error': (out.stderr or out.stdout)[-400:]}
    except Exception as e:       # noqa: BLE001
        return {'error': repr(e)}


# --------------------------------------------------------------------------------------------- our arm

def tag_flops(key, B):
    m = re.search(r'\[(?:train|eval) x(\d+) C(\d+) w(\d+) @(\d+)x(\d+)', key)       # fused run of AIO blocks: 4 contractions per block
    if m:
        nblk, C, w, h, wd = (int(v) for v in m.groups())
        c1, c2 = C - C // 2, C // 2
        return 2. * B * nblk * h * wd * (9 * c1 * w + 9 * w * w + w * 2 * c2 + C * C)
    m = re.search(r'\[k(\d) s(\d) (?:[NT] )?(\d+)->(\d+) @(\d+)x(\d+)', key)
    if not m:
        return 0.
    k, _, ci, co, h, w = (int(v) for v in m.groups())
    return 2. * B * h * w * ci * co * k * k


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cifar10_train', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (default: the workload\'s)')
    ap.add_argument('--ref-device', default='cpu', choices=['cpu', 'cuda'], help='--impl reference: host cores (the arm) or stock PyTorch on the GPU')
    ap.add_argument('--budget', type=float, default=0., help='--impl reference: clip the number of steps so the run ends within about this many seconds (0 = exactly --steps)')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-torch-baseline', action='store_true')
    ap.add_argument('--profile-out', default=None, help='write the per-kernel-class device-time table here (json)')
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    warmup = max(a.warmup, 3)

    if a.impl == 'reference':
        reference_main(a, rank, world, warmup)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ROOT, 'ibinn_b200', 'csrc', 'libibinn_b200.so')):
        ge.build()
    from ibinn_b200 import _lib, config
    from ibinn_b200.engine import ScoreStep, TrainStep
    from ibinn_b200.model import GenerativeClassifier
    if not torch.cuda.is_available():
        raise SystemExit('bench.py (impl ours) needs a B200: there is no CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    W = WORKLOADS[a.workload]
    kind = W['kind']
    B = a.batch or W['batch']
    C = W['classes']
    if world > 1:
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'          # NCCL prints its version banner on stdout: keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=dev)
    args = config.make_args(overrides=W['overrides'])     # the reference's default.ini (+ the workload's keys)

    def make_model():
        torch.manual_seed(0)
        np.random.seed(0)
        return GenerativeClassifier(args).to(dev)

    g = torch.Generator().manual_seed(1 + rank)
    model = make_model()
    xh = torch.randn(B, *model.dims, generator=g).pin_memory()
    lab = torch.randint(0, C, (B,), generator=g)
    eps = float(args['data']['label_smoothing'])
    yh = (torch.zeros(B, C).scatter_(1, lab.view(-1, 1), 1.) * (1 - eps) + eps / C).pin_memory()

    # ---- the engines of this workload
    if kind == 'train':
        model.train()
        engines = [TrainStep(model, batch_size=B, beta=1.0, use_graph=not a.no_graph)]
    elif kind == 'sweep':
        # replicas only: rank r trains the configurations r, r + world, ... of the ramp, each its own model; no data-path collective
        mine = BETA_RAMP[rank::world]
        engines = []
        for i, beta in enumerate(mine):
            m = model if i == 0 else make_model()
            m.train()
            engines.append(TrainStep(m, batch_size=B, beta=beta, use_graph=not a.no_graph, process_group=None, replicas_only=True))
    elif kind == 'fwd':
        model.train()
        engines = [ScoreStep(model, batch_size=B, with_labels=True, loss_mean=True, train_mode=True, use_graph=not a.no_graph)]
    else:
        model.eval()
        engines = [ScoreStep(model, batch_size=B, with_labels=False, loss_mean=False, train_mode=False, use_graph=not a.no_graph)]
    for e in engines:
        e.load(xh, yh) if kind != 'infer' else e.load(xh)
    main_e = engines[0]
    n_cfg_total = len(BETA_RAMP) if kind == 'sweep' else 1
    imgs_per_step_global = B * (n_cfg_total if kind == 'sweep' else world)

    # ---- per-kernel-class device times: one eager pass bracketed with CUDA events (after a warm pass)
    def eager(e):
        if isinstance(e, TrainStep):
            e._eager()
        else:
            e._forward()
    side = getattr(main_e, 'side_streams', None)
    if side is not None:
        main_e.side_streams = None                        # serialised on one stream: every launch gets its own event pair
    eager(main_e)
    torch.cuda.synchronize()
    passes = []
    for _ in range(3):                                    # three profiled passes; per LAUNCH the fastest of the three, so that a one-off
        _lib.profile(True)                                # stall (a host hiccup between the two event records) cannot land in a family
        n0 = _lib.lib().ibinn_launch_count()
        eager(main_e)
        passes.append(_lib.profile(False).summary())
        eager_launches = int(_lib.lib().ibinn_launch_count() - n0)
    table = {}
    for k, v in passes[0].items():
        same = [p[k] for p in passes if k in p and len(p[k]) == len(v)]
        table[k] = (sum(min(col) for col in zip(*same)), len(v))
    if side is not None:
        main_e.side_streams = side
    total_ms = sum(t for t, _ in table.values())
    if kind in ('fwd', 'infer'):
        main_e.refresh()

    def run_all():
        for e in engines:
            e.run()

    # ---- timed region: K steps, inputs resident in HBM
    for _ in range(warmup):
        run_all()
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
        run_all()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of the step's result, every step, through the public engine API
    if kind in ('train', 'sweep'):
        res_h = torch.zeros(len(engines), 4).pin_memory()
        d2h = 16 * len(engines)
    elif kind == 'fwd':
        res_h = torch.zeros(1, 5).pin_memory()
        d2h = 20
    else:
        res_h = torch.zeros(B, 1 + C).pin_memory()           # per-sample score L_x and the class log-likelihoods (ood.py:43-46)
        d2h = B * (1 + C) * 4
    h2d = int(xh.numel() * 4 + (yh.numel() * 4 if kind != 'infer' else 0)) * len(engines)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    f0.record()
    for _ in range(a.steps):
        for i, e in enumerate(engines):
            if kind in ('train', 'sweep'):
                out = e.step(xh, yh)
                res_h[i].copy_(torch.stack([out['L_x_tr'], out['L_y_tr'], out['acc_tr'], out['grad_norm']]))   # synchronous read, like train.py:120-121
            elif kind == 'fwd':
                out = e.step(xh, yh)
                res_h[0].copy_(torch.stack([out['L_x_tr'], out['logits_tr'], out['L_cNLL_tr'], out['L_y_tr'], out['acc_tr']]))
            else:
                out = e.step(xh)
                res_h.copy_(torch.cat([out['L_x_tr'].view(-1, 1), out['logits_tr']], 1))
    f1.record()
    torch.cuda.synchronize()
    ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    loss_ok = bool(np.isfinite(res_h.numpy()).all())
    launches = sum((e.launches if e.launches is not None else eager_launches) for e in engines)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    tf32 = measure_tf32_peak(torch, dev)
    ms_step = ms / a.steps
    value = imgs_per_step_global / (ms_step * 1e-3)
    # ---- roofline of the dominant kernel.  The profiler keys carry the launch shapes
    # ('ibinn_conv_forward[k3 s1 N 32->32 @16x16 pre2 epi1]'), so the algorithmic FLOPs (2 x MAC, counted once although the
    # 3xTF32 arithmetic issues 3 MMAs per product) are exact per launch; the duration is the CUDA-event time of the same launches.
    fam = {'conv_tc_kernel (subnet conv forward + dgrad)': lambda k: k.startswith('ibinn_conv_forward[') and ' s1 ' in k,
           'wgrad_tc_kernel (subnet conv weight gradient)': lambda k: k.startswith('ibinn_conv_wgrad[') and ' s1 ' in k,
           'aio_stage_kernel (fused run of coupling blocks: conv1, conv2, conv3, soft permutation)': lambda k: k.startswith('ibinn_aio_stage_forward[')}
    stats = {}
    for name, pred in fam.items():
        keys = [k for k in table if pred(k)]
        if keys:
            stats[name] = dict(ms=sum(table[k][0] for k in keys), n=sum(table[k][1] for k in keys),
                               flops=sum(tag_flops(k, B) * table[k][1] for k in keys))
    top = max(stats, key=lambda n: stats[n]['ms'])
    st = stats[top]
    ach = st['flops'] / (st['ms'] * 1e-3) / 1e12 if st['ms'] > 0 else 0.
    traffic = None
    try:
        ncu = json.load(open(os.path.join(ROOT, 'profiles', 'r02_ncu_full.json' if os.path.exists(os.path.join(ROOT, 'profiles', 'r02_ncu_full.json')) else 'r01_ncu_full.json')))
        key = top.split(' ')[0]
        traffic = ncu[key + '_eval' if kind in ('fwd', 'infer') and key + '_eval' in ncu else key]['dram_bytes_per_launch']
    except Exception:
        pass
    # HBM-bound kernel families at this batch: algorithmic bytes / CUDA-event time of the same launches
    n_param = sum(q.numel() for q in model.parameters() if q.requires_grad)
    hbm_bytes = {'ibinn_sgd_step': 20. * n_param, 'ibinn_grad_norm_clip': 4. * n_param}
    if a.workload.startswith('cifar'):
        hbm_bytes.update({'ibinn_aio_coupling_forward': 8 * 32768 * B + 48 * 36864 * B,      # 8 blocks of 32 768 B + 48 of 36 864 B per image
                          'ibinn_aio_coupling_backward': 8 * 53248 * B + 48 * 61440 * B,
                          'ibinn_gmm_head_forward': (12288 + 4 * (C + 3)) * B + 4 * C * 3072})
    hbm_fam = {}
    for k, nbytes in hbm_bytes.items():
        if k in table and table[k][0] > 0:
            gbs = nbytes / (table[k][0] * 1e-3) / 1e9
            hbm_fam[k] = {'ms_per_step': table[k][0], 'launches': table[k][1], 'bytes_per_step': nbytes, 'gbs': gbs, 'frac_of_hbm_peak': gbs / pk['hbm']}
    roofline = {'bound': 'tensor', 'kernel': top, 'achieved': ach, 'peak': tf32['sustained'], 'unit': 'TFLOP/s', 'frac': ach / tf32['sustained'],
                'traffic': traffic,
                'peak_source': 'measured in this run: cuBLAS TF32 matmul 8192^3 back to back for 1 s (burst %.0f, sustained %.0f TFLOP/s); '
                               'MEASURED_PEAKS.json (%s) bf16 dense sustained = %.0f' % (tf32['burst'], tf32['sustained'], pk['source'], pk['bf16_sustained']),
                'frac_of_bf16_peak': ach / pk['bf16_sustained'],
                'issued_over_algorithmic': 3.0, 'frac_issued': 3.0 * ach / tf32['sustained'],
                'launches_per_step': st['n'], 'avg_launch_us': st['ms'] * 1e3 / max(st['n'], 1), 'flops_per_launch': st['flops'] / max(st['n'], 1),
                'share_of_step': st['ms'] / total_ms if total_ms else None,
                'whole_step_tflops': W['flop_per_img'] * imgs_per_step_global / world / (ms_step * 1e-3) / 1e12,
                'hbm_families': hbm_fam,
                'families': {n: {'ms_per_step': v['ms'], 'launches': v['n'], 'tflops': (v['flops'] / (v['ms'] * 1e-3) / 1e12 if v['ms'] > 0 else 0.)}
                             for n, v in stats.items()},
                'note': 'fp32-accurate 3xTF32 on tcgen05 (3 MMAs issued per algorithmic product: frac_issued = 3 x frac); time = CUDA-event '
                        'durations of these launches in one eager step, traffic = ncu dram bytes per launch (profiles/)'}
    line = {'metric': METRIC[kind], 'value': value, 'unit': 'img/s', 'n_gpus': world, 'steps': a.steps, 'warmup': warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': config_of(a.workload, B, world),
            'notes': {'cuda_graph': main_e.use_graph, 'dropout': 'as shipped (Dropout2d 0.25 stage 2, Dropout 0.5 fc)' if kind in ('train', 'sweep', 'fwd') else 'eval mode',
                      'l2': 'no explicit flush: one step streams %.0f MB of activations + parameter state, larger than the 126 MB L2'
                            % (torch.cuda.max_memory_allocated() / 2 ** 20)},
            'clocks': clocks,
            'e2e': {'value': imgs_per_step_global / (ms_e2e / a.steps * 1e-3), 'unit': 'img/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': launches * a.steps, 'gpu_launches_per_step': launches,
            'roofline': roofline, 'finite': loss_ok}
    if world > 1:
        dist.destroy_process_group()
    del engines, main_e, model
    torch.cuda.empty_cache()
    if world == 1:
        if not a.no_cpu_baseline:
            r = sub_reference(a.workload, B, 'cpu', 10, 1, 20.)
            line['cpu_baseline'] = r.get('cpu_baseline', r)
        if not a.no_torch_baseline:
            # informational: the reference's own deployment (train.py:45 `.cuda()`): the same unmodified files with stock PyTorch
            # kernels (cuDNN / cuBLAS, TF32 off) on this B200
            r = sub_reference(a.workload, B, 'cuda', 20, 3, 60.)
            line['torch_gpu_baseline'] = ({'value': r['value'], 'unit': 'img/s', 'ms_per_step': r['ms_per_step'], 'sample': r['cpu_baseline']['sample']}
                                          if 'value' in r else r)
    if a.profile_out:
        rows = sorted(((k, v[0], v[1]) for k, v in table.items()), key=lambda r: -r[1])
        json.dump({'total_ms': total_ms, 'launch_calls': eager_launches, 'rows': rows}, open(a.profile_out, 'w'), indent=1)
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
