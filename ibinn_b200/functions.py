"""torch.autograd.Function layer: one Function per graph node (forward = our CUDA kernels,
backward = the hand-derived CUDA backward), so autograd sees ~60 nodes per step instead of the
~3500 of the reference.  Parameter gradients are either returned to autograd, or -- when
engine.FlatParams has attached a flat gradient view to each parameter (`p._ibinn_grad`) --
accumulated straight into that buffer by the kernels (no per-tensor add kernels).
"""
import os

import torch

from . import ops
from .ops import BN, EPI_BNSTATS, EPI_MASKSTATS, EPI_RELUMASK, EPI_STORE, PRE_BN_RELU, PRE_BNBWD, PRE_NONE, PRE_RELU


# ------------------------------------------------------------------------------------ helpers

class GradSink:
    """Where the kernels of one node accumulate parameter gradients."""

    def __init__(self, params):
        self.params = list(params)
        self.direct = all(getattr(p, '_ibinn_grad', None) is not None for p in self.params)
        if self.direct:
            self.targets = [p._ibinn_grad for p in self.params]
        else:
            total = sum(p.numel() for p in self.params)
            flat = torch.zeros(total, dtype=torch.float32, device=self.params[0].device)
            self.targets, o = [], 0
            for p in self.params:
                self.targets.append(flat[o:o + p.numel()].view(p.shape))
                o += p.numel()

    def __getitem__(self, i):
        return self.targets[i]

    def returns(self):
        return [None] * len(self.params) if self.direct else list(self.targets)


def dropout_mask(shape, p, ref):
    """Bernoulli(1-p)/(1-p) mask as a float tensor (torch's Philox stream; parity tests inject masks)."""
    if p <= 0.:
        return None
    return (torch.rand(shape, device=ref.device) >= p).to(torch.float32) / (1. - p)


def pooled_mask(net, shape, p, ref):
    """engine.MaskPool draws all dropout masks of a step with three launches; without it, one mask per call."""
    m = getattr(net, '_ibinn_mask', None)
    if m is not None and getattr(net, '_ibinn_mask_fresh', False) and tuple(m.shape) == tuple(shape) and m.device == ref.device:
        net._ibinn_mask_fresh = False          # one use per redraw: a forward outside TrainStep draws its own mask
        return m
    return dropout_mask(shape, p, ref)


# ------------------------------------------------------------------------------------ conv subnet

def conv_subnet_forward(net, x, training, mask=None, stats=None):
    """basic_residual_block / strided_residual_block (reference inn_architecture.py:68-110).
    x: (B,cin,H,W) possibly a channel slice.  Returns a and the tensors the backward needs.
    stats: (4,width) = batch mean1, inv-std1, mean2, inv-std2 of an earlier training-mode forward: normalise with exactly those
    (the recomputation of the invertibility-based backward), nothing is updated."""
    bn1, bn2 = net.bn1, net.bn2
    dev = x.device
    width = net.conv1.weight.shape[0]
    stride2 = 2 if net.strided else 1
    pre0 = PRE_RELU if net.relu_first else PRE_NONE
    if stats is not None:
        h1 = ops.conv_forward(x, net.conv1.weight, 3, 1, pre=pre0)
        r1 = BN(stats[0], stats[1], bn1.weight, bn1.bias, bn1.eps, False)
        h2 = ops.conv_forward(h1, net.conv2.weight, 3, stride2, pre=PRE_BN_RELU, pre_bn=r1)
        r2 = BN(stats[2], stats[3], bn2.weight, bn2.bias, bn2.eps, False)
        training = True
    elif training:
        sv = torch.empty(4, width, dtype=torch.float32, device=dev)

        def bn_stats(bn, k):
            return dict(save_mean=sv[2 * k], save_invstd=sv[2 * k + 1], running_mean=bn.running_mean,
                        running_var=bn.running_var, nbt=bn.num_batches_tracked, momentum=bn.momentum, eps=bn.eps)
        h1 = ops.conv_forward(x, net.conv1.weight, 3, 1, pre=pre0, epi=EPI_BNSTATS, stats=bn_stats(bn1, 0))
        r1 = BN(sv[0], sv[1], bn1.weight, bn1.bias, bn1.eps, False)
        h2 = ops.conv_forward(h1, net.conv2.weight, 3, stride2, pre=PRE_BN_RELU, pre_bn=r1, epi=EPI_BNSTATS, stats=bn_stats(bn2, 1))
        r2 = BN(sv[2], sv[3], bn2.weight, bn2.bias, bn2.eps, False)
        if mask is None and net.p_drop > 0.:
            mask = pooled_mask(net, (x.shape[0], width), net.p_drop, x)
    else:
        h1 = ops.conv_forward(x, net.conv1.weight, 3, 1, pre=pre0)
        r1 = BN(bn1.running_mean, bn1.running_var, bn1.weight, bn1.bias, bn1.eps, True)
        h2 = ops.conv_forward(h1, net.conv2.weight, 3, stride2, pre=PRE_BN_RELU, pre_bn=r1)
        r2 = BN(bn2.running_mean, bn2.running_var, bn2.weight, bn2.bias, bn2.eps, True)
        mask = None
    a = ops.conv_forward(h2, net.conv3.weight, 1, 1, pre=PRE_BN_RELU, pre_bn=r2, pre_mask=mask, bias=net.conv3.bias)
    return a, (h1, h2, r1, r2, mask, training)


CONV_SUBNET_PARAMS = ('conv1.weight', 'bn1.weight', 'bn1.bias', 'conv2.weight', 'bn2.weight', 'bn2.bias', 'conv3.weight', 'conv3.bias')


def conv_subnet_params(net):
    return [net.conv1.weight, net.bn1.weight, net.bn1.bias, net.conv2.weight, net.bn2.weight, net.bn2.bias,
            net.conv3.weight, net.conv3.bias]


def conv_subnet_backward(net, saved, x, g_a, sink, s0, g_in_out, addend):
    """Backward of conv_subnet_forward.  sink[s0:s0+8] receive the parameter gradients in
    CONV_SUBNET_PARAMS order; the input gradient (+ addend, may alias) is written to g_in_out."""
    h1, h2, r1, r2, mask, training = saved
    dev = x.device
    width = net.conv1.weight.shape[0]
    cin = net.conv1.weight.shape[1]
    stride2 = 2 if net.strided else 1
    dw1, dg1, db1, dw2, dg2, db2, dw3, dbias3 = [sink[s0 + i] for i in range(8)]
    sums = torch.empty(2, 2, width, dtype=torch.float32, device=dev)
    zero_sums = None if training else torch.zeros(2, width, dtype=torch.float32, device=dev)
    g_a = g_a.contiguous()
    # 1x1 head: input gradient (+ BatchNorm-backward sums) and weight / bias gradient in one launch where the shape allows
    g2 = ops.conv1x1_backward(g_a, h2, r2, mask, net.conv3.weight, sums[1], dg2, db2, dw3, dbias3)
    if g2 is None:
        ops.conv_wgrad(g_a, h2, dw3, 1, 1, x_pre=PRE_BN_RELU, x_bn=r2, x_mask=mask, dbias=dbias3)
        g2 = ops.conv_forward(g_a, net.conv3.weight, 1, 1, cout=width, w_transposed=True, epi=EPI_MASKSTATS,
                              mask=dict(h=h2, bn=r2, drop=mask, sums_out=sums[1], dgamma=dg2, dbeta=db2))
    s2 = sums[1] if training else zero_sums
    # middle 3x3 (stride 1 or 2)
    ops.conv_wgrad(g2, h1, dw2, 3, stride2, g_pre=PRE_BNBWD, gh=h2, g_bn=r2, g_sums=s2, x_pre=PRE_BN_RELU, x_bn=r1)
    g1 = ops.conv_forward(g2, net.conv2.weight, 3, 1, cout=width, w_transposed=True, pre=PRE_BNBWD, x2=h2, pre_bn=r2, pre_sums=s2,
                          upsample_to=(h1.shape[2], h1.shape[3]) if net.strided else None, epi=EPI_MASKSTATS,
                          mask=dict(h=h1, bn=r1, sums_out=sums[0], dgamma=dg1, dbeta=db1))
    s1 = sums[0] if training else zero_sums
    # first 3x3
    ops.conv_wgrad(g1, x, dw1, 3, 1, g_pre=PRE_BNBWD, gh=h1, g_bn=r1, g_sums=s1, x_pre=PRE_RELU if net.relu_first else PRE_NONE)
    ops.conv_forward(g1, net.conv1.weight, 3, 1, cout=cin, w_transposed=True, pre=PRE_BNBWD, x2=h1, pre_bn=r1, pre_sums=s1,
                     epi=EPI_RELUMASK if net.relu_first else EPI_STORE, mask=dict(h=x) if net.relu_first else None,
                     addend=addend, out=g_in_out)


class ConvSubnetFn(torch.autograd.Function):
    """Stand-alone subnet call (`net(x)`), used by the reverse pass and by tests."""

    @staticmethod
    def forward(ctx, x, net, mask, *params):
        x = x if x.dim() == 4 else x
        a, saved = conv_subnet_forward(net, x, net.training, mask)
        ctx.net, ctx.saved = net, saved
        ctx.save_for_backward(x)
        return a

    @staticmethod
    def backward(ctx, g_a):
        (x,) = ctx.saved_tensors
        net = ctx.net
        sink = GradSink(conv_subnet_params(net))
        g_x = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        conv_subnet_backward(net, ctx.saved, x, g_a, sink, 0, g_x, None)
        return (g_x, None, None) + tuple(sink.returns())


# ------------------------------------------------------------------------------------ AIO block

class AIOBlockFn(torch.autograd.Function):
    """AIO_Block.forward, rev=False (reference all_in_one_block.py:83-114)."""

    @staticmethod
    def forward(ctx, x, blk, mask, *params):
        x = x.contiguous()
        c1 = blk.split_len1
        a, saved = conv_subnet_forward(blk.s, x[:, :c1], blk.training, mask)
        out, jac = ops.aio_coupling_forward(x, a, blk.w, blk.act_norm, blk.act_offset, blk.clamp, 0.1)
        ctx.blk, ctx.saved = blk, saved
        ctx.save_for_backward(x, a, out)
        return out, jac

    @staticmethod
    def backward(ctx, g_out, g_jac):
        x, a, out = ctx.saved_tensors
        blk = ctx.blk
        c1 = blk.split_len1
        sink = GradSink(conv_subnet_params(blk.s) + [blk.act_norm, blk.act_offset])
        if g_out is None:
            g_out = torch.zeros_like(out)
        g_jac = None if g_jac is None else g_jac.contiguous()
        g_x, g_a = ops.aio_coupling_backward(g_out, g_jac, x, a, out, blk.w, blk.act_norm, blk.act_offset, blk.clamp, 0.1,
                                             sink[8], sink[9])
        conv_subnet_backward(blk.s, ctx.saved, x[:, :c1], g_a, sink, 0, g_x[:, :c1], g_x[:, :c1])
        return (g_x, None, None) + tuple(sink.returns())


INVERTIBLE_BACKWARD = os.environ.get('IBINN_INVERTIBLE_BACKWARD', '0') == '1'
INVERTIBLE_ANCHOR_EVERY = int(os.environ.get('IBINN_INVERTIBLE_ANCHOR', '2'))      # blocks between stored outputs (0: only the run's last output)


def aio_block_recompute(blk, y, stats, mask):
    """Invert an AIO block whose OUTPUT y is known (reference all_in_one_block.py:69,93-103) with the batch statistics of the
    forward pass: returns the block input x and everything the backward kernels need (a, h1, h2, BN references), so that the
    forward pass has to keep nothing but y and the statistics."""
    u = ops.aio_unpermute(y, blk.w_inv, blk.act_norm, blk.act_offset)
    a, saved = conv_subnet_forward(blk.s, u[:, :blk.split_len1], True, mask, stats=stats)
    ops.aio_affine_inverse_(u, a, blk.act_norm, blk.clamp, 0.1)          # in place: u becomes x
    return u, a, saved


class AIOStageFn(torch.autograd.Function):
    """A run of consecutive AIO_Blocks as ONE launch (ops.aio_stage_forward).  The backward walks the run's blocks in
    reverse with the per-block backward kernels, on the tensors the fused forward saved -- or, with INVERTIBLE_BACKWARD
    (IBINN_INVERTIBLE_BACKWARD=1), on tensors recomputed block by block from the run's OUTPUT by inverting each block: the forward
    then keeps only the last output and 4 x width batch statistics per block instead of x, a, h1, h2 of every block."""

    @staticmethod
    def forward(ctx, x, plan, *params):
        x = x.contiguous()
        blocks = plan.blocks
        train = blocks[0].training
        B = x.shape[0]
        masks = None
        if train:
            masks = [pooled_mask(blk.s, (B, plan.width), blk.s.p_drop, x) if blk.s.p_drop > 0. else None for blk in blocks]
        ctx.invertible = bool(INVERTIBLE_BACKWARD and train)          # (grad mode is off inside Function.forward: cannot be consulted here)
        res = ops.aio_stage_forward(plan, x, train, masks, save_intermediates=not ctx.invertible, anchor_every=INVERTIBLE_ANCHOR_EVERY)
        ctx.plan, ctx.masks, ctx.train = plan, masks, train
        if ctx.invertible:
            ctx.anchor_keys = sorted(res['anchors'])
            ctx.save_for_backward(res['out'], res['save'], *[res['anchors'][k] for k in ctx.anchor_keys])
        elif train:
            ctx.save_for_backward(x, res['outs'], res['h1'], res['h2'], res['a'], res['save'])
        ctx.res = res if train else None
        out = res['out']
        return out, res['jac']

    @staticmethod
    def backward(ctx, g_out, g_jac):
        if not ctx.train:
            raise RuntimeError('ibinn_b200: the fused eval-mode run keeps no activations; it is only used under torch.no_grad()')
        blocks = ctx.plan.blocks
        if ctx.invertible:
            y, save = ctx.saved_tensors[:2]
            anchors = dict(zip(ctx.anchor_keys, ctx.saved_tensors[2:]))
        else:
            x, outs, h1, h2, a, save = ctx.saved_tensors
            y = outs[-1]
        if g_out is None:
            g_out = torch.zeros_like(y)
        g = g_out.contiguous()
        g_jac = None if g_jac is None else g_jac.contiguous()
        grads = [None] * len(blocks)
        for k in range(len(blocks) - 1, -1, -1):
            blk = blocks[k]
            net = blk.s
            c1 = blk.split_len1
            sink = GradSink(conv_subnet_params(net) + [blk.act_norm, blk.act_offset])
            if ctx.invertible:
                y = anchors.get(k, y)                                                    # a stored output re-anchors the reconstruction
                xk, ak, saved = aio_block_recompute(blk, y, save[k], ctx.masks[k])        # x_k = f_k^-1(y); y is the block output
                yk = y
            else:
                xk, ak, yk = (x if k == 0 else outs[k - 1]), a[k], outs[k]
                r1 = BN(save[k][0], save[k][1], net.bn1.weight, net.bn1.bias, net.bn1.eps, False)
                r2 = BN(save[k][2], save[k][3], net.bn2.weight, net.bn2.bias, net.bn2.eps, False)
                saved = (h1[k], h2[k], r1, r2, ctx.masks[k], True)
            g_x, g_a = ops.aio_coupling_backward(g, g_jac, xk, ak, yk, blk.w, blk.act_norm, blk.act_offset, blk.clamp, 0.1,
                                                 sink[8], sink[9])
            conv_subnet_backward(net, saved, xk[:, :c1], g_a, sink, 0, g_x[:, :c1], g_x[:, :c1])
            grads[k] = sink.returns()
            g, y = g_x, xk
        ctx.res = None
        return (g, None) + tuple(t for gk in grads for t in gk)


def aio_stage_params(plan):
    return [p for blk in plan.blocks for p in conv_subnet_params(blk.s) + [blk.act_norm, blk.act_offset]]


def aio_block_reverse(blk, y):
    """AIO_Block.forward, rev=True (reference all_in_one_block.py:93-103).  No autograd."""
    u = ops.aio_unpermute(y, blk.w_inv, blk.act_norm, blk.act_offset)
    a, _ = conv_subnet_forward(blk.s, u[:, :blk.split_len1], blk.training)
    jac = ops.aio_affine_inverse_(u, a, blk.act_norm, blk.clamp, 0.1)
    return u, jac


# ------------------------------------------------------------------------------------ downsampling block

class DownBlockFn(torch.autograd.Function):
    """DownsampleCouplingBlock.forward, rev=False (reference downsampling_coupling_block.py:68-94)."""

    @staticmethod
    def forward(ctx, x, blk, mask, *params):
        x = x.contiguous()
        B, C, H, W = x.shape
        c1 = blk.split_len1
        out = torch.empty((B, 4 * C, H // 2, W // 2), dtype=torch.float32, device=x.device)
        jac = torch.empty(B, dtype=torch.float32, device=x.device)
        a1, sv1 = conv_subnet_forward(blk.s_hi, x[:, :c1], blk.training)
        ops.s2d_coupling_forward(x[:, c1:], a1, out[:, 4 * c1:], jac, False, blk.clamp, 0.2)
        a2, sv2 = conv_subnet_forward(blk.s_lo, out[:, 4 * c1:], blk.training, mask)
        ops.s2d_coupling_forward(x[:, :c1], a2, out[:, :4 * c1], jac, True, blk.clamp, 0.2)
        ctx.blk, ctx.sv1, ctx.sv2 = blk, sv1, sv2
        ctx.save_for_backward(x, a1, a2, out)
        return out, jac

    @staticmethod
    def backward(ctx, g_out, g_jac):
        x, a1, a2, out = ctx.saved_tensors
        blk = ctx.blk
        c1 = blk.split_len1
        sink = GradSink(conv_subnet_params(blk.s_hi) + conv_subnet_params(blk.s_lo))
        if g_out is None:
            g_out = torch.zeros_like(out)
        g_out = g_out.contiguous()
        g_jac = None if g_jac is None else g_jac.contiguous()
        g_x = torch.empty_like(x)
        g_a2 = ops.s2d_coupling_backward(g_out[:, :4 * c1], g_jac, x[:, :c1], a2, g_x[:, :c1], blk.clamp, 0.2)
        g_y2 = torch.empty(out[:, 4 * c1:].shape, dtype=torch.float32, device=x.device)
        conv_subnet_backward(blk.s_lo, ctx.sv2, out[:, 4 * c1:], g_a2, sink, 8, g_y2, g_out[:, 4 * c1:])
        g_a1 = ops.s2d_coupling_backward(g_y2, g_jac, x[:, c1:], a1, g_x[:, c1:], blk.clamp, 0.2)
        conv_subnet_backward(blk.s_hi, ctx.sv1, x[:, :c1], g_a1, sink, 0, g_x[:, :c1], g_x[:, :c1])
        return (g_x, None, None) + tuple(sink.returns())


def down_block_reverse(blk, y):
    """DownsampleCouplingBlock.forward, rev=True (reference downsampling_coupling_block.py:84-91)."""
    y = y.contiguous()
    B, C4, Ho, Wo = y.shape
    C = C4 // 4
    c1 = blk.split_len1
    x = torch.empty((B, C, 2 * Ho, 2 * Wo), dtype=torch.float32, device=y.device)
    jac = torch.empty(B, dtype=torch.float32, device=y.device)
    a2, _ = conv_subnet_forward(blk.s_lo, y[:, 4 * c1:], blk.training)
    ops.s2d_coupling_inverse(y[:, :4 * c1], a2, x[:, :c1], jac, False, blk.clamp, 0.2)
    a1, _ = conv_subnet_forward(blk.s_hi, x[:, :c1], blk.training)
    ops.s2d_coupling_inverse(y[:, 4 * c1:], a1, x[:, c1:], jac, True, blk.clamp, 0.2)
    return x, jac


# ------------------------------------------------------------------------------------ FC subnet / GLOW block

def fc_subnet_params(net):
    return [net.lin1.weight, net.lin1.bias, net.lin2.weight, net.lin2.bias]


def fc_subnet_forward(net, x, training, mask=None):
    """fc_constr (reference inn_architecture.py:112-120): Linear -> ReLU -> Dropout -> Linear."""
    h = ops.linear_forward(x, net.lin1.weight, net.lin1.bias)
    if training and mask is None and net.p_drop > 0.:
        mask = pooled_mask(net, h.shape, net.p_drop, h)
    if not training:
        mask = None
    r = ops.linear_forward(h, net.lin2.weight, net.lin2.bias, x_pre=1, x_mask=mask)
    return r, (h, mask)


def fc_subnet_backward(net, saved, x, g_r, sink, s0, g_in_out, accumulate):
    h, mask = saved
    dW1, db1, dW2, db2 = [sink[s0 + i] for i in range(4)]
    g_r = g_r.contiguous()
    ops.linear_wgrad(g_r, h, dW2, db2, x_pre=1, x_mask=mask)
    g_h = torch.empty_like(h)
    ops.linear_dgrad(g_r, net.lin2.weight, g_h, accumulate=False)
    ops.linear_wgrad(g_h, x, dW1, db1, g_pre=2, h=h, g_mask=mask)
    ops.linear_dgrad(g_h, net.lin1.weight, g_in_out, accumulate=accumulate, g_pre=2, h=h, g_mask=mask)


class FCSubnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, net, mask, *params):
        r, saved = fc_subnet_forward(net, x, net.training, mask)
        ctx.net, ctx.saved = net, saved
        ctx.save_for_backward(x)
        return r

    @staticmethod
    def backward(ctx, g_r):
        (x,) = ctx.saved_tensors
        sink = GradSink(fc_subnet_params(ctx.net))
        g_x = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        fc_subnet_backward(ctx.net, ctx.saved, x, g_r, sink, 0, g_x, False)
        return (g_x, None, None) + tuple(sink.returns())


class GlowBlockFn(torch.autograd.Function):
    """FrEIA v0.2 GLOWCouplingBlock forward (reference call site inn_architecture.py:161); s2 acts first."""

    @staticmethod
    def forward(ctx, x, blk, masks, *params):
        x = x.contiguous()
        B, D = x.shape
        l1 = blk.split_len1
        m1, m2 = masks if masks is not None else (None, None)
        out = torch.empty_like(x)
        jac = torch.empty(B, dtype=torch.float32, device=x.device)
        r2, sv2 = fc_subnet_forward(blk.s2, x[:, l1:], blk.training, m2)
        ops.glow_affine_forward(x[:, :l1], r2, out[:, :l1], jac, False, blk.clamp)
        r1, sv1 = fc_subnet_forward(blk.s1, out[:, :l1], blk.training, m1)
        ops.glow_affine_forward(x[:, l1:], r1, out[:, l1:], jac, True, blk.clamp)
        ctx.blk, ctx.sv1, ctx.sv2 = blk, sv1, sv2
        ctx.save_for_backward(x, r1, r2, out)
        return out, jac

    @staticmethod
    def backward(ctx, g_out, g_jac):
        x, r1, r2, out = ctx.saved_tensors
        blk = ctx.blk
        l1 = blk.split_len1
        sink = GradSink(fc_subnet_params(blk.s1) + fc_subnet_params(blk.s2))
        if g_out is None:
            g_out = torch.zeros_like(out)
        g_out = g_out.contiguous()
        g_jac = None if g_jac is None else g_jac.contiguous()
        g_x = torch.empty_like(x)
        g_r1 = ops.glow_affine_backward(g_out[:, l1:], g_jac, x[:, l1:], r1, g_x[:, l1:], blk.clamp)
        g_y1 = g_out[:, :l1].contiguous().clone() if g_out[:, :l1].is_contiguous() else g_out[:, :l1].contiguous()
        fc_subnet_backward(blk.s1, ctx.sv1, out[:, :l1], g_r1, sink, 0, g_y1, True)
        g_r2 = ops.glow_affine_backward(g_y1, g_jac, x[:, :l1], r2, g_x[:, :l1], blk.clamp)
        fc_subnet_backward(blk.s2, ctx.sv2, x[:, l1:], g_r2, sink, 4, g_x[:, l1:], True)
        return (g_x, None, None) + tuple(sink.returns())


def glow_block_reverse(blk, y):
    y = y.contiguous()
    B, D = y.shape
    l1 = blk.split_len1
    x = torch.empty_like(y)
    jac = torch.empty(B, dtype=torch.float32, device=y.device)
    r1, _ = fc_subnet_forward(blk.s1, y[:, :l1], blk.training)
    ops.glow_affine_inverse(y[:, l1:], r1, x[:, l1:], jac, False, blk.clamp)
    r2, _ = fc_subnet_forward(blk.s2, x[:, l1:], blk.training)
    ops.glow_affine_inverse(y[:, :l1], r2, x[:, :l1], jac, True, blk.clamp)
    return x, jac


# ------------------------------------------------------------------------------------ data movement nodes

class DCTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, scale):
        ctx.weight, ctx.scale, ctx.shape = weight, scale, x.shape
        return ops.dct2d_forward(x, weight, scale)

    @staticmethod
    def backward(ctx, g):
        _, C, N, _ = ctx.shape
        return ops.dct2d_inverse(g, ctx.weight, C, N, ctx.scale, True), None, None


class HaarFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fac):
        ctx.fac = fac
        return ops.haar_forward(x, fac)

    @staticmethod
    def backward(ctx, g):
        return ops.haar_inverse(g, ctx.fac), None


class PermuteFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, perm, perm_inv):
        ctx.perm_inv = perm_inv
        return ops.permute_features(x, perm)

    @staticmethod
    def backward(ctx, g):
        return ops.permute_features(g, ctx.perm_inv), None, None


# ------------------------------------------------------------------------------------ GMM / IB head

class GMMHeadFn(torch.autograd.Function):
    """GenerativeClassifier.forward after the INN (reference model.py:144-163)."""

    @staticmethod
    def forward(ctx, z, jac, mu, phi, y, loss_mean):
        mu2 = mu.reshape(-1, mu.shape[-1])
        o = ops.gmm_head_forward(z, jac, mu2, phi, y)
        ctx.save_for_backward(z, mu2, phi, o['logits'], o['lse'])
        ctx.y, ctx.loss_mean, ctx.mu_shape, ctx.mu_param, ctx.phi_param = y, loss_mean, mu.shape, mu, phi
        has_y = y is not None
        if loss_mean:
            m = o['means']
            outs = (m[0], m[1], m[2] if has_y else None, m[3] if has_y else None, m[4] if has_y else None)
        else:
            outs = (o['L_x'], o['logits'], o.get('L_cNLL'), o.get('L_y'), m_acc(o) if has_y else None)
        ctx.mark_non_differentiable(*[t for t in outs[4:] if t is not None])
        return outs

    @staticmethod
    def backward(ctx, g_lx, g_logits, g_lc, g_ly, g_acc):
        z, mu2, phi, logits, lse = ctx.saved_tensors
        B, C = logits.shape
        grads = {'L_x': g_lx, 'logits': g_logits, 'L_cNLL': g_lc, 'L_y': g_ly}
        sv, sl = (1. / B, 1. / (B * C)) if ctx.loss_mean else (1., 1.)
        tm = getattr(ctx.mu_param, '_ibinn_grad', None)
        tp = getattr(ctx.phi_param, '_ibinn_grad', None)
        g_z, g_jac, g_mu, g_phi = ops.gmm_head_backward(z, mu2, phi, ctx.y, logits, lse, grads, sv, sl,
                                                        None if tm is None else tm.view(C, -1), tp)
        return g_z, g_jac, (None if tm is not None else g_mu.view(ctx.mu_shape)), (None if tp is not None else g_phi), None, None


def m_acc(o):
    return o['means'][4]
