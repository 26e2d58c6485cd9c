"""Thin Python wrappers (tensor -> pointer/shape marshalling) over the C ABI.  No autograd here;
`functions.py` builds the torch.autograd.Function layer on top.  Everything raises if the CUDA
library is missing or a tensor is not on the B200 (see _lib.py)."""
import ctypes
import os

import torch

from . import _lib
from ._lib import BnRef, ConvArgs, WgradArgs, check, ptr, stream

PRE_NONE, PRE_RELU, PRE_BN_RELU, PRE_BNBWD = 0, 1, 2, 3
EPI_STORE, EPI_BNSTATS, EPI_MASKSTATS, EPI_RELUMASK = 0, 1, 2, 3

_COUNTERS = {}
_N_COUNTERS = 4096


def counter_ptr(ref):
    """A zero-initialised uint32 slot on ref's device.  Kernels return their slot to zero, so the
    slots are handed out round-robin (kernels on one stream never overlap on a slot)."""
    key = (ref.device.type, ref.device.index)
    ent = _COUNTERS.get(key)
    if ent is None:
        ent = [torch.zeros(_N_COUNTERS, dtype=torch.int32, device=ref.device), 0]
        _COUNTERS[key] = ent
    ent[1] = (ent[1] + 1) % _N_COUNTERS
    return ent[0].data_ptr() + 4 * ent[1]


def counters_ptr(ref, n):
    """n consecutive zero-initialised uint32 slots (see counter_ptr)"""
    if n > _N_COUNTERS:
        raise ValueError('ibinn_b200: %d counters requested' % n)
    key = (ref.device.type, ref.device.index)
    if key not in _COUNTERS:
        counter_ptr(ref)
    ent = _COUNTERS[key]
    if ent[1] + 1 + n > _N_COUNTERS:
        ent[1] = -1
    first = ent[1] + 1
    ent[1] = first + n - 1
    return ent[0].data_ptr() + 4 * first


def _nchw(t):
    """(pointer, batch stride) of a (B,C,H,W) tensor that is contiguous within each image."""
    if t is None:
        return None, 0
    B, C, H, W = t.shape
    st = t.stride()
    if not (st[3] == 1 and (H == 1 or st[2] == W) and (C == 1 or st[1] == H * W)):
        raise ValueError('ibinn_b200: tensor must be NCHW-contiguous within an image, got strides %s for shape %s' % (st, tuple(t.shape)))
    return ptr(t, strided=True), (st[0] if B > 1 else C * H * W)


def _rows(t):
    """(pointer, row stride) of a (B,L) tensor with unit inner stride."""
    B, L_ = t.shape
    if t.stride(1) != 1 and L_ > 1:
        raise ValueError('ibinn_b200: inner stride must be 1')
    return ptr(t, strided=True), (t.stride(0) if B > 1 else L_)


def bn_ref(mean, scale, gamma, beta, eps, scale_is_variance):
    return BnRef(ptr(mean), ptr(scale), ptr(gamma), ptr(beta), float(eps), int(bool(scale_is_variance)))


class BN:
    """What a kernel needs to know about one BatchNorm2d layer at one call."""
    __slots__ = ('mean', 'scale', 'gamma', 'beta', 'eps', 'is_var')

    def __init__(self, mean, scale, gamma, beta, eps, is_var):
        self.mean, self.scale, self.gamma, self.beta, self.eps, self.is_var = mean, scale, gamma, beta, eps, is_var

    def ref(self):
        return bn_ref(self.mean, self.scale, self.gamma, self.beta, self.eps, self.is_var)


def conv_forward(x, w, ksize, stride=1, *, cout=None, pre=PRE_NONE, pre_bn=None, pre_mask=None, x2=None, pre_sums=None,
                 epi=EPI_STORE, bias=None, addend=None, out=None, w_transposed=False, upsample_to=None,
                 stats=None, mask=None):
    """One launch of the subnet convolution kernel (see include/ibinn_b200.h, ibinn_conv_args).

    stats (EPI_BNSTATS): dict(save_mean, save_invstd, running_mean, running_var, nbt, momentum, eps)
    mask  (EPI_MASKSTATS / EPI_RELUMASK): dict(h, bn=BN|None, drop, sums_out, dgamma, dbeta)
    """
    B, Cin, Hs, Ws = x.shape
    if cout is None:
        cout = w.shape[1] if w_transposed else w.shape[0]
    Hin, Win = (Hs, Ws) if upsample_to is None else upsample_to
    pad = ksize // 2
    Hout = (Hin + 2 * pad - ksize) // stride + 1
    Wout = (Win + 2 * pad - ksize) // stride + 1
    if out is None:
        out = torch.empty((B, cout, Hout, Wout), dtype=torch.float32, device=x.device)
    a = ConvArgs()
    a.x, a.x_bs = _nchw(x)
    a.x2, a.x2_bs = _nchw(x2)
    a.w, a.bias = ptr(w), ptr(bias)
    a.addend, a.addend_bs = _nchw(addend)
    a.y, a.y_bs = _nchw(out)
    a.B, a.Cin, a.Cout, a.Hs, a.Ws, a.Hout, a.Wout = B, Cin, cout, Hs, Ws, Hout, Wout
    a.ksize, a.stride, a.upsample2, a.Hin, a.Win = ksize, stride, int(upsample_to is not None), Hin, Win
    a.w_transposed, a.pre, a.epi = int(w_transposed), pre, epi
    if pre_bn is not None:
        a.pre_bn = pre_bn.ref()
    a.pre_mask, a.pre_sums = ptr(pre_mask), ptr(pre_sums)
    a.pre_inv_count = 1.0 / float(B * Hs * Ws)
    if stats is not None:
        a.save_mean, a.save_invstd = ptr(stats['save_mean']), ptr(stats['save_invstd'])
        a.running_mean, a.running_var = ptr(stats.get('running_mean')), ptr(stats.get('running_var'))
        a.num_batches_tracked = ptr(stats.get('nbt'), torch.int64)
        mom = stats.get('momentum')
        a.momentum = -1.0 if mom is None else float(mom)
        a.bn_eps = float(stats['eps'])
    if mask is not None:
        a.m_h, a.m_bs = _nchw(mask['h'])
        if mask.get('bn') is not None:
            a.m_bn = mask['bn'].ref()
        a.m_drop, a.sums_out = ptr(mask.get('drop')), ptr(mask.get('sums_out'))
        a.m_dgamma, a.m_dbeta = ptr(mask.get('dgamma')), ptr(mask.get('dbeta'))
    L = _lib.lib()
    keep = None
    nw = L.ibinn_conv_wprep_floats(ctypes.byref(a))
    if nw > 0:
        cached = getattr(w, '_ibinn_wprep', None)          # engine.WeightPrep: operand images prepared once per step
        if cached is not None and cached['valid'] and cached[bool(w_transposed)].numel() == nw:
            a.wprep, a.wprep_ready = ptr(cached[bool(w_transposed)]), 1
        else:
            wprep = torch.empty(int(nw), dtype=torch.float32, device=x.device)
            a.wprep = ptr(wprep)
    if epi in (EPI_BNSTATS, EPI_MASKSTATS):
        n = L.ibinn_conv_partials_floats(ctypes.byref(a))
        if n < 0:
            check(int(n), 'conv_partials_floats')
        keep = torch.empty(max(int(n), 1), dtype=torch.float32, device=x.device)
        a.partials = ptr(keep)
        a.counter = counter_ptr(x)
    if _lib._profiler is not None:
        _lib.set_tag('k%d s%d %s %d->%d @%dx%d pre%d epi%d' % (ksize, stride, 'T' if w_transposed else 'N', Cin, cout, Hout, Wout, pre, epi))
    check(L.ibinn_conv_forward(ctypes.byref(a), stream(x)), 'conv_forward')
    _lib.set_tag('')
    return out


DEFERRED_WGRAD = None       # list while engine.TrainStep collects deferred wgrad reductions, else None
WGRAD_STREAMS = None        # engine.TrainStep: side streams the weight-gradient kernels of a backward pass are forked onto
_wgrad_rr = 0


WGRAD_REDUCE = None         # engine.TrainStep: dict(stream, group, pinned, upto, keep) - record reductions in groups on a third stream


def join_wgrad_streams():
    """make the current stream wait for everything forked onto the wgrad side streams (and the reduction stream)"""
    if WGRAD_STREAMS:
        cur = torch.cuda.current_stream()
        for s in WGRAD_STREAMS:
            cur.wait_stream(s)
    if WGRAD_REDUCE is not None:
        torch.cuda.current_stream().wait_stream(WGRAD_REDUCE['stream'])


def _reduce_table(entries, device, slot):
    """descriptor table on the device, staged through the pinned buffer of `slot` (dict(buf, ev), reused across steps: no
    pinned allocation may happen under graph capture).  In eager mode the host runs ahead of the device, so the slot's last
    copy is waited for before the buffer is rewritten."""
    import numpy as np
    raw = b''.join(bytes(d) for _, d in entries)
    capturing = torch.cuda.is_current_stream_capturing()
    if slot.get('buf') is None or slot['buf'].numel() < len(raw):
        slot['buf'], slot['ev'] = torch.empty(len(raw), dtype=torch.uint8).pin_memory(), None
    if slot.get('ev') is not None and not capturing:
        slot['ev'].synchronize()
    slot['buf'][:len(raw)].copy_(torch.from_numpy(np.frombuffer(raw, dtype=np.uint8).copy()))
    table = torch.empty(len(raw), dtype=torch.uint8, device=device)
    table.copy_(slot['buf'][:len(raw)], non_blocking=True)
    if not capturing:
        slot['ev'] = torch.cuda.Event()
        slot['ev'].record()
    else:
        slot['ev'] = None
    return table


def flush_deferred_wgrad(entries, ref, pinned):
    """one launch for all deferred wgrad record reductions; `pinned` = reusable staging slot (dict) or None; returns the slot"""
    entries = [e for e in entries if e[1] is not None]
    if not entries:
        return pinned
    pinned = pinned if isinstance(pinned, dict) else {}
    table = _reduce_table(entries, ref.device, pinned)
    check(_lib.lib().ibinn_wgrad_reduce_many(ptr(table, torch.uint8), len(entries), stream(ref)), 'wgrad_reduce_many')
    return pinned


def flush_wgrad_group(final=False):
    """Sum the records written since the last group on the reduction stream: the reduction is bandwidth-bound (the records of
    a 64-channel layer are 9 MB) and runs beside the latency-bound kernels of the other two streams while the records are
    still in L2, instead of as one 230 us tail after the backward pass."""
    R = WGRAD_REDUCE
    live = [e for e in DEFERRED_WGRAD[R['upto']:] if e[1] is not None]
    if not live or (len(live) < R['group'] and not final):
        return
    R['upto'] = len(DEFERRED_WGRAD)
    rs = R['stream']
    for s in (WGRAD_STREAMS or [torch.cuda.current_stream()]):
        rs.wait_stream(s)
    k = len(R['keep'])
    if len(R['pinned']) <= k:
        R['pinned'].append({})
    with torch.cuda.stream(rs):
        table = _reduce_table(live, live[0][0].device, R['pinned'][k])
        check(_lib.lib().ibinn_wgrad_reduce_many(ptr(table, torch.uint8), len(live), rs.cuda_stream), 'wgrad_reduce_many')
    R['keep'].append(table)


def conv_wprep_floats(w, transposed, H, W):
    """size of the tensor-core operand image of filter tensor `w` (0 if that conv shape runs on the CUDA-core kernel)"""
    a = ConvArgs()
    co, ci, ks, _ = w.shape
    if transposed:
        co, ci = ci, co
    a.B, a.Cin, a.Cout, a.Hs, a.Ws, a.Hout, a.Wout, a.Hin, a.Win = 1, ci, co, H, W, H, W, H, W
    a.ksize, a.stride = ks, 1
    return int(_lib.lib().ibinn_conv_wprep_floats(ctypes.byref(a)))


def wprep_many(table, n, ref):
    check(_lib.lib().ibinn_wprep_many(ptr(table, torch.uint8), n, stream(ref)), 'wprep_many')


def conv_wgrad(g, x, dw, ksize, stride=1, *, g_pre=PRE_NONE, gh=None, g_bn=None, g_sums=None, x_pre=PRE_NONE, x_bn=None,
               x_mask=None, dbias=None):
    B, Cout, Hout, Wout = g.shape
    _, Cin, Hin, Win = x.shape
    a = WgradArgs()
    a.g, a.g_bs = _nchw(g)
    a.gh, a.gh_bs = _nchw(gh)
    a.g_pre = g_pre
    if g_bn is not None:
        a.g_bn = g_bn.ref()
    a.g_sums = ptr(g_sums)
    a.g_inv_count = 1.0 / float(B * Hout * Wout)
    a.x, a.x_bs = _nchw(x)
    a.x_pre = x_pre
    if x_bn is not None:
        a.x_bn = x_bn.ref()
    a.x_mask = ptr(x_mask)
    a.dw, a.dbias = ptr(dw), ptr(dbias)
    a.B, a.Cin, a.Cout, a.Hin, a.Win, a.Hout, a.Wout, a.ksize, a.stride = B, Cin, Cout, Hin, Win, Hout, Wout, ksize, stride
    nws = _lib.lib().ibinn_conv_wgrad_workspace_floats(ctypes.byref(a))
    if nws > 0:
        ws = torch.empty(int(nws), dtype=torch.float32, device=g.device)
        a.workspace = ptr(ws)
        if DEFERRED_WGRAD is not None:
            # engine.TrainStep: leave the per-CTA records in `ws`; one launch sums the records of all layers later
            ncta, rec = ctypes.c_int32(), ctypes.c_int32()
            check(_lib.lib().ibinn_conv_wgrad_records(ctypes.byref(a), ctypes.byref(ncta), ctypes.byref(rec)), 'conv_wgrad_records')
            a.defer_reduce = 1
            DEFERRED_WGRAD.append((ws, _lib.WreduceDesc(ptr(ws), ptr(dw), ptr(dbias), ncta.value, rec.value, Cout, Cin, ksize * ksize)))
    if _lib._profiler is not None:
        _lib.set_tag('k%d s%d %d->%d @%dx%d' % (ksize, stride, Cin, Cout, Hout, Wout))
    st = stream(g)
    if DEFERRED_WGRAD is not None and WGRAD_STREAMS and g.is_cuda:
        # The weight gradient feeds nothing downstream in the backward pass: fork it onto a side stream so it fills the
        # SMs the (latency-bound) dgrad chain leaves idle.  The operands stay referenced until the join.
        global _wgrad_rr
        side = WGRAD_STREAMS[_wgrad_rr % len(WGRAD_STREAMS)]
        _wgrad_rr += 1
        side.wait_stream(torch.cuda.current_stream())
        st = side.cuda_stream
        # every tensor the kernel reads through a pointer stays referenced until the join - including the BatchNorm statistics
        # inside g_bn / x_bn, which autograd frees right after this block's backward while the kernel may still be queued
        DEFERRED_WGRAD.append(((g, x, gh, g_sums, x_mask, dw, dbias, g_bn, x_bn), None))
    check(_lib.lib().ibinn_conv_wgrad(ctypes.byref(a), st), 'conv_wgrad')
    _lib.set_tag('')
    if DEFERRED_WGRAD is not None and WGRAD_REDUCE is not None:
        flush_wgrad_group()


# Off by default: measured on the whole step, the one CUDA-core launch (~30 us) loses to the two tcgen05 launches it replaces, which
# run concurrently on the main and the weight-gradient stream (9.40 -> 10.21 ms per step with it); kept as a tested alternative.
FUSED_1X1_BWD = os.environ.get('IBINN_FUSED_1X1_BWD', '0') == '1'


def conv1x1_backward(g, h, bn, drop, w, sums_out, dgamma, dbeta, dw, dbias):
    """Backward of the subnets' 1x1 output convolution in ONE launch (ibinn_conv1x1_backward): returns the masked gradient w.r.t.
    the BatchNorm output (what conv_forward(..., w_transposed=True, epi=EPI_MASKSTATS) returns) and accumulates dw / dbias (what
    conv_wgrad does), or returns None when the shape is not served (the caller then takes the two-launch route)."""
    if not FUSED_1X1_BWD or not g.is_cuda:      # (the simulator build has no batched record reduction: it keeps the two-launch route)
        return None
    B, Cout, H, W = g.shape
    Cin = h.shape[1]
    a = _lib.Conv1x1BwdArgs()
    a.B, a.Cin, a.Cout, a.H, a.W = B, Cin, Cout, H, W
    L = _lib.lib()
    ncta, rec = ctypes.c_int32(), ctypes.c_int32()
    if L.ibinn_conv1x1_bwd_geometry(ctypes.byref(a), ctypes.byref(ncta), ctypes.byref(rec)) != 0:
        return None
    g = g.contiguous()
    a.g = ptr(g)
    a.h, a.h_bs = _nchw(h)
    a.bn = bn.ref()
    a.drop, a.w = ptr(drop), ptr(w)
    g_in = torch.empty((B, Cin, H, W), dtype=torch.float32, device=g.device)
    part = torch.empty(ncta.value * 2 * Cin, dtype=torch.float32, device=g.device)
    wrec = torch.empty(ncta.value * rec.value, dtype=torch.float32, device=g.device)
    a.g_in, a.sums_out, a.dgamma, a.dbeta = ptr(g_in), ptr(sums_out), ptr(dgamma), ptr(dbeta)
    a.partials, a.counter, a.wrecords = ptr(part), counter_ptr(g), ptr(wrec)
    if _lib._profiler is not None:
        _lib.set_tag('k1 %d->%d @%dx%d' % (Cin, Cout, H, W))
    check(L.ibinn_conv1x1_backward(ctypes.byref(a), stream(g)), 'conv1x1_backward')
    _lib.set_tag('')
    desc = _lib.WreduceDesc(ptr(wrec), ptr(dw), ptr(dbias), ncta.value, rec.value, Cout, Cin, 1)
    if DEFERRED_WGRAD is not None:
        DEFERRED_WGRAD.append((wrec, desc))                          # engine.TrainStep: summed with every other layer's records in one launch
        DEFERRED_WGRAD.append(((part, g, h, drop), None))            # operands stay referenced until the reduction has run
    else:
        flush_deferred_wgrad([(wrec, desc)], g, None)
    return g_in


# ------------------------------------------------------------------------------------ fused run of AIO blocks

_BARRIERS = {}


def _grid_barrier_counter(ref):
    """one zero-initialised 64-bit counter per device for the grid-wide barriers of ibinn_aio_stage_forward (monotonic)"""
    key = (ref.device.type, ref.device.index)
    t = _BARRIERS.get(key)
    if t is None:
        t = _BARRIERS[key] = torch.zeros(2, dtype=torch.int64, device=ref.device)
    return t


class AioStagePlan:
    """A run of consecutive AIO_Blocks of one resolution stage that ONE ibinn_aio_stage_forward launch covers (see
    include/ibinn_b200.h).  Holds what does not change between calls: the blocks, the geometry, a pinned staging buffer for
    the descriptor table and -- when no engine manages the tensor-core filter images -- its own images."""

    def __init__(self, blocks):
        self.blocks = list(blocks)
        b0 = self.blocks[0]
        self.C = b0.in_channels
        self.width = b0.s.conv1.weight.shape[0]
        self.clamp = float(b0.clamp)
        self._pinned = {}
        self._own = None
        self._ok = {}

    def __deepcopy__(self, memo):
        import copy
        return AioStagePlan([copy.deepcopy(b, memo) for b in self.blocks])      # staging buffers / CUDA events are per instance

    def weights(self):
        return [w for blk in self.blocks for w in (blk.s.conv1.weight, blk.s.conv2.weight, blk.s.conv3.weight, blk.w)]

    def supported(self, x, train):
        if not x.is_cuda or x.dim() != 4 or x.dtype != torch.float32:
            return False
        B, C, H, W = x.shape
        key = (B, H, W, bool(train))
        ok = self._ok.get(key)
        if ok is None:
            a = _lib.AioStageArgs()
            a.nblk, a.B, a.C, a.H, a.W, a.width, a.train = len(self.blocks), B, C, H, W, self.width, int(train)
            ok = self._ok[key] = bool(_lib.lib().ibinn_aio_stage_supported(ctypes.byref(a)))
        return ok

    def weight_images(self, ref):
        """operand images (forward orientation) of conv1 / conv2 / conv3 / w of every block, as a flat list of tensors"""
        ws = self.weights()
        if all(getattr(w, '_ibinn_wprep', None) is not None and w._ibinn_wprep['valid'] for w in ws):
            return [w._ibinn_wprep[False] for w in ws]          # engine.WeightPrep refreshed them for this step
        if self._own is None or self._own['dev'] != ref.device:
            import numpy as np
            bufs, descs = [], []
            for w in ws:
                co, ci, ks, _ = w.shape
                n = conv_wprep_floats(w, False, 8, 8)
                if n <= 0:
                    raise RuntimeError('ibinn_b200: no tensor-core operand image for a filter of shape %s' % (tuple(w.shape),))
                buf = torch.empty(n, dtype=torch.float32, device=ref.device)
                bufs.append(buf)
                descs.append(bytes(_lib.WprepDesc(w.data_ptr(), buf.data_ptr(), ci, co, ks, 0)))
            raw = np.frombuffer(b''.join(descs), dtype=np.uint8).copy()
            self._own = dict(dev=ref.device, bufs=bufs, table=torch.from_numpy(raw).to(ref.device), n=len(descs),
                             ptrs=[w.data_ptr() for w in ws])
        if self._own['ptrs'] != [w.data_ptr() for w in ws]:
            self._own = None                                    # the parameters moved (e.g. engine.FlatParams): rebuild
            return self.weight_images(ref)
        wprep_many(self._own['table'], self._own['n'], ref)     # one launch; the filters may have changed since the last call
        return self._own['bufs']


def aio_stage_forward(plan, x, train, masks=None, save_intermediates=True, anchor_every=0):
    """One launch for the whole run.  Returns dict(out, jac[, outs, h1, h2, a, save]) -- in train mode the per-block tensors
    the backward entry points need (outs[k] = output of block k); with save_intermediates=False only the batch statistics
    (`save`), the last output and -- every `anchor_every` blocks counted from the end -- an anchor output (`anchors[k]`) are
    written: the invertibility-based backward recomputes the rest, re-starting from an anchor so that the reconstruction error of
    this (random-init: chaotic) network cannot grow over more than `anchor_every` blocks."""
    B, C, H, W = x.shape
    nblk, width, c2 = len(plan.blocks), plan.width, C // 2
    dev = x.device
    f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    res = dict(jac=f(B))
    keep_all = train and save_intermediates
    if keep_all:
        res.update(outs=f(nblk, B, C, H, W), h1=f(nblk, B, width, H, W), h2=f(nblk, B, width, H, W), a=f(nblk, B, 2 * c2, H, W),
                   save=f(nblk, 4, width))
        res['out'] = res['outs'][nblk - 1]
    else:
        res['out'] = f(B, C, H, W)
        if train:
            res['save'] = f(nblk, 4, width)
            res['anchors'] = {k: f(B, C, H, W) for k in range(nblk - 1) if anchor_every > 0 and (nblk - 1 - k) % anchor_every == 0}
    imgs = plan.weight_images(x)
    entries = []
    for k, blk in enumerate(plan.blocks):
        net = blk.s
        d = _lib.AioBlockDesc()
        d.wprep1, d.wprep2, d.wprep3, d.wprep_perm = (ptr(t) for t in imgs[4 * k:4 * k + 4])
        d.bias3 = ptr(net.conv3.bias)
        for name, bn in (('bn1', net.bn1), ('bn2', net.bn2)):
            if train:
                setattr(d, name, BnRef(None, None, ptr(bn.weight), ptr(bn.bias), float(bn.eps), 0))
            else:
                setattr(d, name, bn_ref(bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.eps, True))
        d.act_norm, d.act_offset = ptr(blk.act_norm), ptr(blk.act_offset)
        d.relu_first = int(net.relu_first)
        if train:
            m = masks[k] if masks is not None else None
            d.drop_mask = ptr(m)
            d.save = ptr(res['save'][k])
            if keep_all:
                d.out, d.h1, d.h2, d.a = ptr(res['outs'][k]), ptr(res['h1'][k]), ptr(res['h2'][k]), ptr(res['a'][k])
            elif k == nblk - 1:
                d.out = ptr(res['out'])
            elif k in res['anchors']:
                d.out = ptr(res['anchors'][k])
            d.running_mean1, d.running_var1, d.nbt1 = ptr(net.bn1.running_mean), ptr(net.bn1.running_var), ptr(net.bn1.num_batches_tracked, torch.int64)
            d.running_mean2, d.running_var2, d.nbt2 = ptr(net.bn2.running_mean), ptr(net.bn2.running_var), ptr(net.bn2.num_batches_tracked, torch.int64)
        elif k == nblk - 1:
            d.out = ptr(res['out'])
        entries.append((None, d))
    table = _reduce_table(entries, dev, plan._pinned)
    a = _lib.AioStageArgs()
    a.x, a.jac, a.blocks = ptr(x), ptr(res['jac']), ptr(table, torch.uint8)
    a.nblk, a.B, a.C, a.H, a.W, a.width, a.train = nblk, B, C, H, W, width, int(train)
    a.clamp, a.alpha = plan.clamp, 0.1
    mom = plan.blocks[0].s.bn1.momentum
    a.momentum = -1.0 if mom is None else float(mom)
    keep = [table, imgs]
    if train:
        L = _lib.lib()
        rec = f(max(int(L.ibinn_aio_stage_workspace_floats(ctypes.byref(a))), 1))
        a.stat_records, a.barrier = ptr(rec), ptr(_grid_barrier_counter(x), torch.int64)
        keep.append(rec)
    if _lib._profiler is not None:
        _lib.set_tag('%s x%d C%d w%d @%dx%d' % ('train' if train else 'eval', nblk, C, width, H, W))
    check(_lib.lib().ibinn_aio_stage_forward(ctypes.byref(a), stream(x)), 'aio_stage_forward')
    _lib.set_tag('')
    res['_keep'] = keep
    return res


# ------------------------------------------------------------------------------------ couplings

def aio_coupling_forward(x, a, w, act_norm, act_offset, clamp, alpha=0.1):
    B, C, H, W = x.shape
    x = x.contiguous()
    out = torch.empty_like(x)
    jac = torch.empty(B, dtype=torch.float32, device=x.device)
    check(_lib.lib().ibinn_aio_coupling_forward(ptr(x), ptr(a), ptr(w), ptr(act_norm), ptr(act_offset), ptr(out), ptr(jac),
                                                B, C, H, W, clamp, alpha, stream(x)), 'aio_coupling_forward')
    return out, jac


DEFERRED_AIO = None         # list while engine.TrainStep collects the ActNorm-gradient records of the AIO blocks, else None


def aio_coupling_backward(g_out, g_jac, x, a, out, w, act_norm, act_offset, clamp, alpha, g_act_norm, g_act_offset):
    """returns g_x, g_a; accumulates into g_act_norm / g_act_offset (C)."""
    B, C, H, W = x.shape
    g_out = g_out.contiguous()
    g_x = torch.empty_like(x)
    g_a = torch.empty_like(a)
    L = _lib.lib()
    nrec = L.ibinn_aio_backward_records(B, C, H, W)
    if nrec <= 0:
        check(nrec, 'aio_backward_records')
    part = torch.empty(nrec * 2 * C, dtype=torch.float32, device=x.device)
    defer = DEFERRED_AIO is not None
    check(L.ibinn_aio_coupling_backward(ptr(g_out), ptr(g_jac), ptr(x), ptr(a), ptr(out), ptr(w), ptr(act_norm),
                                        ptr(act_offset), ptr(g_x), ptr(g_a), ptr(g_act_norm), ptr(g_act_offset),
                                        ptr(part), None if defer else counter_ptr(x), B, C, H, W, clamp, alpha, int(defer), stream(x)),
          'aio_coupling_backward')
    if defer:
        # engine.TrainStep: one launch sums the records of all blocks after the backward pass
        DEFERRED_AIO.append((part, _lib.AioReduceDesc(ptr(part), ptr(g_act_offset), ptr(g_act_norm), nrec, C)))
    return g_x, g_a


def flush_deferred_aio(entries, ref, pinned=None):
    """one launch for the ActNorm-gradient records of every AIO block; returns the pinned staging tensor (keep it alive)"""
    pinned = pinned if isinstance(pinned, dict) else {}
    if not entries:
        return pinned
    table = _reduce_table(entries, ref.device, pinned)
    check(_lib.lib().ibinn_aio_param_reduce_many(ptr(table, torch.uint8), len(entries), stream(ref)), 'aio_param_reduce_many')
    return pinned


def aio_unpermute(y, w_inv, act_norm, act_offset):
    B, C, H, W = y.shape
    y = y.contiguous()
    u = torch.empty_like(y)
    check(_lib.lib().ibinn_aio_unpermute(ptr(y), ptr(w_inv), ptr(act_norm), ptr(act_offset), ptr(u), B, C, H, W, stream(y)),
          'aio_unpermute')
    return u


def aio_affine_inverse_(u, a, act_norm, clamp, alpha=0.1):
    B, C, H, W = u.shape
    jac = torch.empty(B, dtype=torch.float32, device=u.device)
    check(_lib.lib().ibinn_aio_affine_inverse(ptr(u), ptr(a), ptr(act_norm), ptr(jac), B, C, H, W, clamp, alpha, stream(u)),
          'aio_affine_inverse')
    return jac


def s2d_coupling_forward(x, a, out, jac, accumulate, clamp, alpha=0.2):
    B, cp, H, W = x.shape
    xp, xbs = _nchw(x)
    op, obs = _nchw(out)
    check(_lib.lib().ibinn_s2d_coupling_forward(xp, xbs, ptr(a), op, obs, ptr(jac), int(accumulate), B, cp, H, W, clamp, alpha,
                                                stream(x)), 's2d_coupling_forward')


def s2d_coupling_inverse(y, a, x, jac, accumulate, clamp, alpha=0.2):
    B, cp, H, W = x.shape
    xp, xbs = _nchw(x)
    yp, ybs = _nchw(y)
    check(_lib.lib().ibinn_s2d_coupling_inverse(yp, ybs, ptr(a), xp, xbs, ptr(jac), int(accumulate), B, cp, H, W, clamp, alpha,
                                                stream(x)), 's2d_coupling_inverse')


def s2d_coupling_backward(g_out, g_jac, x, a, g_x, clamp, alpha=0.2):
    B, cp, H, W = x.shape
    gp, gbs = _nchw(g_out)
    xp, xbs = _nchw(x)
    gxp, gxbs = _nchw(g_x)
    g_a = torch.empty_like(a)
    check(_lib.lib().ibinn_s2d_coupling_backward(gp, gbs, ptr(g_jac), xp, xbs, ptr(a), gxp, gxbs, ptr(g_a), B, cp, H, W, clamp,
                                                 alpha, stream(x)), 's2d_coupling_backward')
    return g_a


def glow_affine_forward(x, r, y, jac, accumulate, clamp):
    B, L_ = x.shape
    xp, xrs = _rows(x)
    yp, yrs = _rows(y)
    check(_lib.lib().ibinn_glow_affine_forward(xp, xrs, ptr(r), yp, yrs, ptr(jac), int(accumulate), B, L_, clamp, stream(x)),
          'glow_affine_forward')


def glow_affine_inverse(y, r, x, jac, accumulate, clamp):
    B, L_ = y.shape
    xp, xrs = _rows(x)
    yp, yrs = _rows(y)
    check(_lib.lib().ibinn_glow_affine_inverse(yp, yrs, ptr(r), xp, xrs, ptr(jac), int(accumulate), B, L_, clamp, stream(y)),
          'glow_affine_inverse')


def glow_affine_backward(g_y, g_jac, x, r, g_x, clamp):
    B, L_ = x.shape
    gp, grs = _rows(g_y)
    xp, xrs = _rows(x)
    gxp, gxrs = _rows(g_x)
    g_r = torch.empty_like(r)
    check(_lib.lib().ibinn_glow_affine_backward(gp, grs, ptr(g_jac), xp, xrs, ptr(r), gxp, gxrs, ptr(g_r), B, L_, clamp,
                                                stream(x)), 'glow_affine_backward')
    return g_r


# ------------------------------------------------------------------------------------ data movement

def dct2d_forward(x, M, scale):
    B, C, N, _ = x.shape
    x = x.contiguous()
    out = torch.empty((B, C * N * N), dtype=torch.float32, device=x.device)
    check(_lib.lib().ibinn_dct2d_forward(ptr(x), ptr(M), ptr(out), B, C, N, scale, stream(x)), 'dct2d_forward')
    return out


def dct2d_inverse(t, M, C, N, scale, transpose_M):
    B = t.shape[0]
    t = t.contiguous()
    x = torch.empty((B, C, N, N), dtype=torch.float32, device=t.device)
    check(_lib.lib().ibinn_dct2d_inverse(ptr(t), ptr(M), ptr(x), B, C, N, scale, int(transpose_M), stream(t)), 'dct2d_inverse')
    return x


def haar_forward(x, fac):
    B, C, H, W = x.shape
    x = x.contiguous()
    y = torch.empty((B, 4 * C, H // 2, W // 2), dtype=torch.float32, device=x.device)
    check(_lib.lib().ibinn_haar_forward(ptr(x), ptr(y), B, C, H, W, fac, stream(x)), 'haar_forward')
    return y


def haar_inverse(y, fac):
    B, C4, Ho, Wo = y.shape
    y = y.contiguous()
    x = torch.empty((B, C4 // 4, 2 * Ho, 2 * Wo), dtype=torch.float32, device=y.device)
    check(_lib.lib().ibinn_haar_inverse(ptr(y), ptr(x), B, C4 // 4, 2 * Ho, 2 * Wo, fac, stream(y)), 'haar_inverse')
    return x


def permute_features(x, index):
    B, D = x.shape
    x = x.contiguous()
    out = torch.empty_like(x)
    check(_lib.lib().ibinn_permute_features(ptr(x), ptr(index, torch.int32), ptr(out), B, D, stream(x)), 'permute_features')
    return out


# ------------------------------------------------------------------------------------ linear

def _linear_ws(ref, M, N, Kc):
    """(workspace tensor | None, workspace pointer, tickets pointer) of a GEMM with an (M, N) result and contraction length Kc: with them
    the split-K partial results are summed in a fixed order (bit-reproducible step) instead of with fp32 atomics"""
    L = _lib.lib()
    n = int(L.ibinn_linear_workspace_floats(M, N, Kc))
    if n < 0:
        check(n, 'linear_workspace_floats')
    if n == 0:
        return None, None, None
    ws = torch.empty(n, dtype=torch.float32, device=ref.device)
    return ws, ptr(ws), counters_ptr(ref, int(L.ibinn_linear_tickets(M, N, Kc)))


def linear_forward(x, w, bias, *, x_pre=0, x_mask=None):
    B, K = x.shape
    N = w.shape[0]
    xp, xrs = _rows(x)
    y = torch.empty((B, N), dtype=torch.float32, device=x.device)
    ws, wsp, tk = _linear_ws(x, B, N, K)
    check(_lib.lib().ibinn_linear_forward(xp, xrs, ptr(x_mask), x_pre, ptr(w), ptr(bias), ptr(y), B, K, N, wsp, tk, stream(x)),
          'linear_forward')
    return y


def linear_dgrad(g, w, dx, *, accumulate, g_pre=0, h=None, g_mask=None):
    B, N = g.shape
    K = w.shape[1]
    dp, drs = _rows(dx)
    ws, wsp, tk = _linear_ws(g, B, K, N)
    check(_lib.lib().ibinn_linear_dgrad(ptr(g), ptr(h), ptr(g_mask), g_pre, ptr(w), dp, drs, int(accumulate), B, K, N, wsp, tk, stream(g)),
          'linear_dgrad')


def linear_wgrad(g, x, dw, db, *, g_pre=0, h=None, g_mask=None, x_pre=0, x_mask=None):
    B, N = g.shape
    K = x.shape[1]
    xp, xrs = _rows(x)
    st = stream(g)
    ws, wsp, tk = _linear_ws(g, N, K, B)
    if DEFERRED_WGRAD is not None and WGRAD_STREAMS and g.is_cuda:
        # like the convolution weight gradients: nothing downstream in the backward pass reads dw, so it leaves the critical path
        global _wgrad_rr
        side = WGRAD_STREAMS[_wgrad_rr % len(WGRAD_STREAMS)]
        _wgrad_rr += 1
        side.wait_stream(torch.cuda.current_stream())
        st = side.cuda_stream
        DEFERRED_WGRAD.append(((g, x, h, g_mask, x_mask, dw, db, ws), None))      # operands stay referenced until the join
    check(_lib.lib().ibinn_linear_wgrad(ptr(g), ptr(h), ptr(g_mask), g_pre, xp, xrs, ptr(x_mask), x_pre, ptr(dw), ptr(db), B, K, N, wsp, tk, st),
          'linear_wgrad')


# ------------------------------------------------------------------------------------ head / optimiser

HEAD_GEMM = True        # scoring batches: distance GEMM on tcgen05 (ibinn_gmm_head_forward_gemm); False keeps the CUDA-core kernels


def gmm_head_forward(z, jac, mu, phi, y, head_ws=None):
    """head_ws: optional dict owned by a caller that KNOWS the class means do not change between its calls (a scoring loop): the
    tensor-core path then prepares the class-mean operand image once instead of per call."""
    B, D = z.shape
    C = phi.shape[0]
    dev = z.device
    z = z.contiguous()
    f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    out = dict(L_x=f(B), logits=f(B, C), lse=f(B), means=f(5))
    if y is not None:
        y = y.contiguous()
        out.update(L_cNLL=f(B), L_y=f(B), correct=f(B))
    part = f(B * 5)
    L = _lib.lib()
    # (the 10-class configs keep head_fwd_stream_kernel: class means resident in shared memory, 40 % of the HBM peak at B = 8192
    #  against 20 % for the GEMM formulation, whose N = 16 tiles leave the tensor pipe nothing to amortise; measured, profiles/)
    if HEAD_GEMM and z.is_cuda and (C > 10 or HEAD_GEMM == 'force') and L.ibinn_gmm_head_gemm_supported(B, C, D):
        key = (mu.data_ptr(), B, C, D)
        prepared = head_ws is not None and head_ws.get('key') == key
        if prepared:
            ws = head_ws['ws']
        else:
            ws = f(int(L.ibinn_gmm_head_gemm_workspace_floats(B, C, D)))
            if head_ws is not None:
                head_ws.update(key=key, ws=ws)
        ent = (key, ws)
        check(L.ibinn_gmm_head_forward_gemm(ptr(z), ptr(jac.contiguous()), ptr(mu), ptr(phi), ptr(y), ptr(out['L_x']), ptr(out['logits']),
                                            ptr(out.get('L_cNLL')), ptr(out.get('L_y')), ptr(out.get('correct')), ptr(out['lse']),
                                            ptr(out['means']), ptr(part), counter_ptr(z), ptr(ent[1]), int(prepared), B, C, D, stream(z)),
              'gmm_head_forward_gemm')
        out['_ws'] = ent[1]
        return out
    check(_lib.lib().ibinn_gmm_head_forward(ptr(z), ptr(jac.contiguous()), ptr(mu), ptr(phi), ptr(y), ptr(out['L_x']), ptr(out['logits']),
                                            ptr(out.get('L_cNLL')), ptr(out.get('L_y')), ptr(out.get('correct')), ptr(out['lse']),
                                            ptr(out['means']), ptr(part), counter_ptr(z), B, C, D, stream(z)), 'gmm_head_forward')
    return out


def gmm_head_backward(z, mu, phi, y, logits, lse, grads, scale_vec, scale_logits, g_mu=None, g_phi=None):
    """grads: dict with optional 'L_x','L_cNLL','L_y','logits' upstream gradients (per-sample, or 0-dim when the
    means were returned).  Returns g_z, g_jac, g_mu, g_phi (the latter two accumulated in place if given)."""
    B, D = z.shape
    C = phi.shape[0]
    dev = z.device

    def gs(name):
        g = grads.get(name)
        if g is None:
            return None, 0
        g = g.contiguous()
        return ptr(g), (0 if g.numel() == 1 else 1)

    acc_mu, acc_phi = g_mu is not None, g_phi is not None
    if g_mu is None:
        g_mu = torch.empty((C, D), dtype=torch.float32, device=dev)
    if g_phi is None:
        g_phi = torch.empty(C, dtype=torch.float32, device=dev)
    G = torch.empty((B, C), dtype=torch.float32, device=dev)
    Glw = torch.empty((B, C), dtype=torch.float32, device=dev)
    g_z = torch.empty((B, D), dtype=torch.float32, device=dev)
    g_jac = torch.empty(B, dtype=torch.float32, device=dev)
    (p1, s1), (p2, s2), (p3, s3), (p4, s4) = gs('L_x'), gs('L_cNLL'), gs('L_y'), gs('logits')
    check(_lib.lib().ibinn_gmm_head_backward(ptr(z), ptr(mu), ptr(phi), ptr(y), ptr(logits), ptr(lse), p1, s1, p2, s2, p3, s3, p4, s4,
                                             scale_vec, scale_logits, ptr(G), ptr(Glw), ptr(g_z), ptr(g_jac), ptr(g_mu), int(acc_mu),
                                             ptr(g_phi), int(acc_phi), B, C, D, stream(z)), 'gmm_head_backward')
    return g_z, g_jac, g_mu, g_phi


def grad_norm_clip(flat_grad, max_norm, extra_sumsq=None):
    """-> tensor [norm, clip_coef] on the device (no host sync)."""
    L = _lib.lib()
    dev = flat_grad.device
    part = torch.empty(int(L.ibinn_grad_norm_partials_doubles()), dtype=torch.float64, device=dev)
    out = torch.empty(2, dtype=torch.float32, device=dev)
    n_extra = 0 if extra_sumsq is None else extra_sumsq.numel()
    check(L.ibinn_grad_norm_clip(ptr(flat_grad), flat_grad.numel(), ptr(extra_sumsq), n_extra, float(max_norm),
                                 ptr(part, torch.float64), counter_ptr(flat_grad), ptr(out), stream(flat_grad)), 'grad_norm_clip')
    return out


def sgd_step_(param, grad, buf, hyper, clip_coef=None, grad_scale=1.0):
    cp = None if clip_coef is None else ptr(clip_coef)
    check(_lib.lib().ibinn_sgd_step(ptr(param), ptr(grad), ptr(buf), param.numel(), ptr(hyper), cp, float(grad_scale),
                                    stream(param)), 'sgd_step')
