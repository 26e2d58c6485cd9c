"""CPU oracle for the IB-INN hot path.  TEST INFRASTRUCTURE ONLY.

A *functional* restatement (plain torch ops on a flat dict of tensors, any dtype, runs in
fp64 for parity and fp32 for the CPU baseline) of the algorithm in the reference files

    /root/reference/model.py:113-165               GMM distances + IB losses
    /root/reference/inn_architecture.py:30-164     graph topology, subnets
    /root/reference/all_in_one_block.py:61-114     AIO coupling block
    /root/reference/downsampling_coupling_block.py:49-97
    /root/reference/dct_transform.py:64-104
    FrEIA v0.2 (pinned c5fe1af, NOT on disk): Reshape, HaarDownsampling, PermuteRandom,
        GLOWCouplingBlock, Flatten, graph evaluation order

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (ibinn_b200/) never does and has no CPU path.

Pinning: the reference ships NO golden vectors or tests.  This oracle is
pinned against the reference's own unmodified python files executed in the build container
through oracle/ref_shims (see oracle/make_golden.py; fixtures in tests/golden/).  The FrEIA
pieces are restated from the published v0.2 behaviour with no source or vector to check
against: **parity unpinned for the FrEIA-owned arithmetic** (Haar, PermuteRandom, GLOW
coupling, graph order); everything the reference itself owns is pinned.

Tensor names follow the reference's state_dict: `module_list.<i>.s.<j>.weight`, `...act_norm`,
`...w`, `...s_hi.<j>...`, `...s1.<j>...` etc.
"""
import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-4          # inn_architecture.py:46-48
BN_MOMENTUM = 0.999


# --------------------------------------------------------------------------- config / topology

@dataclass
class Cfg:
    dataset: str = 'CIFAR10'
    ch_pad: int = 0
    fc_width: int = 1024
    n_fc: int = 1
    use_dct: bool = True
    conv_widths: Tuple[int, ...] = (16, 32, 64)
    n_blocks: Tuple[int, ...] = (8, 24, 24)
    dropouts: Tuple[float, ...] = (0., 0., 0.25)
    dropout_fc: float = 0.5
    groups: int = 1
    clamp: float = 0.7
    act_norm: float = 0.75
    weight_init: float = 1.0
    mu_init: float = 5.0

    @property
    def dims(self):
        if self.dataset == 'MNIST':
            return (28, 28)
        return (3 + self.ch_pad, 32, 32)

    @property
    def n_classes(self):
        return 100 if self.dataset == 'CIFAR100' else 10

    @property
    def ndim_tot(self):
        return int(np.prod(self.dims))


def cfg_from_args(args) -> Cfg:
    """Same keys, same eval()/float()/int() parsing as model.py:17-38, inn_architecture.py:32-42."""
    m = args['model']
    return Cfg(dataset=args['data']['dataset'],
               ch_pad=eval(args['data']['pad_noise_channels']),
               fc_width=int(m['fc_width']), n_fc=int(m['n_coupling_blocks_fc']),
               use_dct=eval(m['dct_pooling']), conv_widths=tuple(eval(m['conv_widths'])),
               n_blocks=tuple(eval(m['n_coupling_blocks_conv'])),
               dropouts=tuple(eval(m['dropout_conv'])), dropout_fc=float(m['dropout_fc']),
               groups=int(m['n_groups']), clamp=float(m['clamp']), act_norm=float(m['act_norm']),
               weight_init=eval(m['weight_init']), mu_init=eval(m['mu_init']))


@dataclass
class NodeSpec:
    kind: str                       # reshape|haar|aio|down|dct|flatten|perm|glow
    idx: int                        # position in module_list (input node is 0)
    name: str
    dims_in: Tuple[int, ...]
    dims_out: Tuple[int, ...]
    extra: dict = field(default_factory=dict)

    @property
    def key(self):
        return 'module_list.%d' % self.idx


def graph_spec(cfg: Cfg) -> List[NodeSpec]:
    """The node chain of inn_architecture.py:122-164."""
    if cfg.groups != 1:
        raise NotImplementedError('n_groups != 1 is never used by a shipped config')
    nodes: List[NodeSpec] = []
    dims = tuple(cfg.dims)

    def add(kind, name, dims_out, **extra):
        nonlocal dims
        nodes.append(NodeSpec(kind, len(nodes) + 1, name, dims, tuple(dims_out), extra))
        dims = tuple(dims_out)

    if cfg.dataset == 'MNIST':
        add('reshape', 'reshape', (1,) + tuple(cfg.dims))
        add('haar', 'haar', (dims[0] * 4, dims[1] // 2, dims[2] // 2), rebalance=1.)
    for i, (width, nb) in enumerate(zip(cfg.conv_widths, cfg.n_blocks)):
        if cfg.dataset == 'MNIST' and i == 0:
            continue
        for k in range(nb):
            add('aio', 'CONV_%d_%d' % (i, k), dims, width=width, dropout=cfg.dropouts[i],
                relu_first=not (i == 0 and k == 0), clamp=cfg.clamp)
        if i < len(cfg.conv_widths) - 1:
            add('down', 'DOWN_%d' % i, (dims[0] * 4, dims[1] // 2, dims[2] // 2),
                width=2 * width, dropout=cfg.dropouts[i], clamp=cfg.clamp)
    if cfg.use_dct:
        add('dct', 'DCT', (dims[0] * dims[1] * dims[2],), rebalance=0.5)
    else:
        add('flatten', 'Flatten', (dims[0] * dims[1] * dims[2],))
    for k in range(cfg.n_fc):
        add('perm', 'PERM_FC_%d' % k, dims, seed=k)
        add('glow', 'FC_%d' % k, dims, clamp=2.0, width=cfg.fc_width, dropout=cfg.dropout_fc)
    return nodes


# --------------------------------------------------------------------------- fixed matrices

def dct_matrix(n: int, dtype=torch.float64) -> torch.Tensor:
    """W[k, m] = 2 cos(pi (2m+1) k / (2n)); equals `dct_1d(I).t()` of dct_transform.py:75."""
    k = torch.arange(n, dtype=torch.float64).view(-1, 1)
    m = torch.arange(n, dtype=torch.float64).view(1, -1)
    return (2. * torch.cos(math.pi * (2 * m + 1) * k / (2 * n))).to(dtype)


def permutation(d: int, seed: int) -> np.ndarray:
    """FrEIA PermuteRandom: numpy legacy global RNG, seed -> permutation(d)."""
    return np.random.RandomState(seed).permutation(d)


def haar_kernels(dtype) -> torch.Tensor:
    """(4,2,2) sign patterns, output channel 4c+j (FrEIA HaarDownsampling)."""
    k = torch.ones(4, 2, 2, dtype=dtype)
    k[1, :, 1] = -1
    k[2, 1, :] = -1
    k[3, 1, 0] = -1
    k[3, 0, 1] = -1
    return k


# --------------------------------------------------------------------------- building blocks

def space_to_depth(x):
    """out[b, 4c+2dy+dx, h, w] = x[b, c, 2h+dy, 2w+dx]  (downsampling_coupling_block.py:35-50)."""
    b, c, h, w = x.shape
    x = x.view(b, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4)
    return x.reshape(b, 4 * c, h // 2, w // 2)


def depth_to_space(x):
    b, c4, h, w = x.shape
    c = c4 // 4
    x = x.view(b, c, 2, 2, h, w).permute(0, 1, 4, 2, 5, 3)
    return x.reshape(b, c, 2 * h, 2 * w)


class _State:
    """Per-call scratch: BN mode, dropout masks, running-stat updates, activations."""

    def __init__(self, train, masks, collect, relu_masks=None):
        self.train = train
        self.masks = masks or {}
        # optional {'<subnet key>.relu<i>': 0/1 tensor}: the ReLU decisions are taken from the caller instead of from the
        # sign of the pre-activation (flip-free gradient parity, tests/test_blocks.py::test_*_injected_relu_masks)
        self.relu_masks = relu_masks
        self.new_stats: Dict[str, torch.Tensor] = {}
        self.collect = collect
        self.records: Dict[str, torch.Tensor] = {}


def _bn(h, P, key, st: _State):
    g, b = P[key + '.weight'], P[key + '.bias']
    if st.train:
        mean = h.mean(dim=(0, 2, 3))
        var = h.var(dim=(0, 2, 3), unbiased=False)
        n = h.numel() // h.shape[1]
        st.new_stats[key + '.running_mean'] = ((1 - BN_MOMENTUM) * P[key + '.running_mean']
                                               + BN_MOMENTUM * mean.detach())
        st.new_stats[key + '.running_var'] = ((1 - BN_MOMENTUM) * P[key + '.running_var']
                                              + BN_MOMENTUM * var.detach() * n / max(n - 1, 1))
    else:
        mean, var = P[key + '.running_mean'], P[key + '.running_var']
    inv = torch.rsqrt(var + BN_EPS)
    return (h - mean.view(1, -1, 1, 1)) * (inv * g).view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def _dropout(h, key, p, st: _State):
    """Mask is injected (parity tests) -- entries are 0 or 1/(1-p); absent mask == p=0."""
    m = st.masks.get(key)
    if m is None:
        return h
    return h * m.to(h.dtype).view(*m.shape, *([1] * (h.dim() - m.dim())))


def _relu(h, key, st: _State):
    m = st.relu_masks.get(key) if st.relu_masks is not None else None
    if m is None:
        return F.relu(h)
    return h * m.to(h.dtype)


def conv_subnet(h, P, key, relu_first, strided, st: _State):
    """basic_residual_block / strided_residual_block, inn_architecture.py:68-110."""
    j = 0
    if relu_first:
        h = _relu(h, key + '.relu0', st)
        j = 1
    h = F.conv2d(h, P['%s.%d.weight' % (key, j)], padding=1)
    h = _relu(_bn(h, P, '%s.%d' % (key, j + 1), st), key + '.relu1', st)
    h = F.conv2d(h, P['%s.%d.weight' % (key, j + 3)], padding=1, stride=2 if strided else 1)
    h = _relu(_bn(h, P, '%s.%d' % (key, j + 4), st), key + '.relu2', st)
    if strided:
        last = j + 6
    else:
        h = _dropout(h, '%s.%d' % (key, j + 6), None, st)
        last = j + 7
    return F.conv2d(h, P['%s.%d.weight' % (key, last)], P['%s.%d.bias' % (key, last)])


def fc_subnet(h, P, key, st: _State):
    """fc_constr, inn_architecture.py:112-120."""
    h = _relu(F.linear(h, P[key + '.0.weight'], P[key + '.0.bias']), key + '.relu1', st)
    h = _dropout(h, key + '.2', None, st)
    return F.linear(h, P[key + '.3.weight'], P[key + '.3.bias'])


def aio_forward(x, P, n: NodeSpec, st: _State):
    """all_in_one_block.py:83-114 (forward direction)."""
    C = n.dims_in[0]
    c1, c2 = C - C // 2, C // 2
    clamp = n.extra['clamp']
    x1, x2 = x[:, :c1], x[:, c1:]
    a = conv_subnet(x1, P, n.key + '.s', n.extra['relu_first'], False, st)
    sig = clamp * torch.tanh(0.1 * a[:, :c2])
    y2 = x2 * torch.exp(sig) + a[:, c2:]
    u = torch.cat((x1, y2), 1)
    v = F.conv2d(u, P[n.key + '.w'].to(x.dtype))
    an = P[n.key + '.act_norm']
    out = v * an.exp() + P[n.key + '.act_offset']
    jac = sig.sum(dim=(1, 2, 3)) + an.sum() * (n.dims_in[1] * n.dims_in[2])
    if st.collect:
        st.records[n.key + '.a'] = a
    return out, jac


def aio_reverse(y, P, n: NodeSpec, st: _State):
    C = n.dims_in[0]
    c1, c2 = C - C // 2, C // 2
    clamp = n.extra['clamp']
    an = P[n.key + '.act_norm']
    u = F.conv2d((y - P[n.key + '.act_offset']) * (-an).exp(), P[n.key + '.w_inv'].to(y.dtype))
    x1, y2 = u[:, :c1], u[:, c1:]
    a = conv_subnet(x1, P, n.key + '.s', n.extra['relu_first'], False, st)
    sig = clamp * torch.tanh(0.1 * a[:, :c2])
    x2 = (y2 - a[:, c2:]) * torch.exp(-sig)
    jac = -sig.sum(dim=(1, 2, 3)) - an.sum() * (n.dims_in[1] * n.dims_in[2])
    return torch.cat((x1, x2), 1), jac


def down_forward(x, P, n: NodeSpec, st: _State):
    """downsampling_coupling_block.py:68-97 (forward direction)."""
    C = n.dims_in[0]
    c1, c2 = C // 2, C - C // 2
    clamp = n.extra['clamp']
    x1, x2 = x[:, :c1], x[:, c1:]
    y2 = space_to_depth(x2)
    a1 = conv_subnet(x1, P, n.key + '.s_hi', False, True, st)
    s1 = clamp * torch.tanh(0.2 * a1[:, :4 * c2])
    y2 = y2 * torch.exp(s1) + a1[:, 4 * c2:]
    y1 = space_to_depth(x1)
    a2 = conv_subnet(y2, P, n.key + '.s_lo', False, False, st)
    s2 = clamp * torch.tanh(0.2 * a2[:, :4 * c1])
    y1 = y1 * torch.exp(s2) + a2[:, 4 * c1:]
    jac = s1.sum(dim=(1, 2, 3)) + s2.sum(dim=(1, 2, 3))
    if st.collect:
        st.records[n.key + '.a1'] = a1
        st.records[n.key + '.a2'] = a2
    return torch.cat((y1, y2), 1), jac


def down_reverse(y, P, n: NodeSpec, st: _State):
    C = n.dims_in[0]
    c1, c2 = C // 2, C - C // 2
    clamp = n.extra['clamp']
    y1, y2 = y[:, :4 * c1], y[:, 4 * c1:]
    a2 = conv_subnet(y2, P, n.key + '.s_lo', False, False, st)
    s2 = clamp * torch.tanh(0.2 * a2[:, :4 * c1])
    x1 = depth_to_space((y1 - a2[:, 4 * c1:]) * torch.exp(-s2))
    a1 = conv_subnet(x1, P, n.key + '.s_hi', False, True, st)
    s1 = clamp * torch.tanh(0.2 * a1[:, :4 * c2])
    x2 = depth_to_space((y2 - a1[:, 4 * c2:]) * torch.exp(-s1))
    jac = -s1.sum(dim=(1, 2, 3)) - s2.sum(dim=(1, 2, 3))
    return torch.cat((x1, x2), 1), jac


def haar_forward(x, n: NodeSpec):
    b, c, h, w = x.shape
    k = haar_kernels(x.dtype)
    blocks = x.view(b, c, h // 2, 2, w // 2, 2).permute(0, 1, 2, 4, 3, 5)      # b c h2 w2 dy dx
    out = torch.einsum('bchwyx,jyx->bcjhw', blocks, k).reshape(b, 4 * c, h // 2, w // 2)
    fac = 0.5 * n.extra['rebalance']
    jac = (c * h * w) / 4 * (math.log(16.) + 4 * math.log(fac))
    return out * fac, jac


def haar_reverse(y, n: NodeSpec):
    b, c4, h, w = y.shape
    c = c4 // 4
    k = haar_kernels(y.dtype)
    fac = 0.5 / n.extra['rebalance']
    blocks = torch.einsum('bcjhw,jyx->bchywx', y.view(b, c, 4, h, w) * fac, k)
    jac = -(c * 2 * h * 2 * w) / 4 * (math.log(16.) + 4 * math.log(0.5 * n.extra['rebalance']))
    return blocks.reshape(b, c, 2 * h, 2 * w), jac


def dct_forward(x, n: NodeSpec):
    """dct_transform.py:90-104.  t[b,c,kh,kw] = sum W[kh,h] W[kw,w] x[b,c,h,w];
    out[b, kw*N*C + kh*C + c] = t / (2(N+1)/rebalance); declared log-det N^2 C ln(rebalance)."""
    C, N = n.dims_in[0], n.dims_in[1]
    W = dct_matrix(N, x.dtype)
    t = torch.einsum('kh,bchw,lw->bckl', W, x, W)
    out = t.permute(0, 3, 2, 1).reshape(x.shape[0], -1) / (2 * (N + 1) / n.extra['rebalance'])
    return out, (N ** 2 * C) * math.log(n.extra['rebalance'])


def dct_reverse(z, n: NodeSpec):
    C, N = n.dims_in[0], n.dims_in[1]
    Wi = torch.linalg.inv(dct_matrix(N, torch.float64)).to(z.dtype)
    t = z.view(z.shape[0], N, N, C).permute(0, 3, 2, 1) * (2 * (N + 1) / n.extra['rebalance'])
    x = torch.einsum('hk,bckl,wl->bchw', Wi, t, Wi)
    return x, -(N ** 2 * C) * math.log(n.extra['rebalance'])


def glow_forward(x, P, n: NodeSpec, st: _State):
    """FrEIA GLOWCouplingBlock: s2 acts first; log_e(s) = clamp*0.636*atan(s/clamp)."""
    D = n.dims_in[0]
    l1, l2 = D // 2, D - D // 2
    cl = n.extra['clamp']
    x1, x2 = x[:, :l1], x[:, l1:]
    r2 = fc_subnet(x2, P, n.key + '.s2', st)
    e2 = cl * 0.636 * torch.atan(r2[:, :l1] / cl)
    y1 = torch.exp(e2) * x1 + r2[:, l1:]
    r1 = fc_subnet(y1, P, n.key + '.s1', st)
    e1 = cl * 0.636 * torch.atan(r1[:, :l2] / cl)
    y2 = torch.exp(e1) * x2 + r1[:, l2:]
    return torch.cat((y1, y2), 1), e1.sum(dim=1) + e2.sum(dim=1)


def glow_reverse(y, P, n: NodeSpec, st: _State):
    D = n.dims_in[0]
    l1, l2 = D // 2, D - D // 2
    cl = n.extra['clamp']
    y1, y2 = y[:, :l1], y[:, l1:]
    r1 = fc_subnet(y1, P, n.key + '.s1', st)
    e1 = cl * 0.636 * torch.atan(r1[:, :l2] / cl)
    x2 = (y2 - r1[:, l2:]) * torch.exp(-e1)
    r2 = fc_subnet(x2, P, n.key + '.s2', st)
    e2 = cl * 0.636 * torch.atan(r2[:, :l1] / cl)
    x1 = (y1 - r2[:, l1:]) * torch.exp(-e2)
    return torch.cat((x1, x2), 1), -e1.sum(dim=1) - e2.sum(dim=1)


def node_forward(x, P, n: NodeSpec, st: _State):
    if n.kind == 'aio':
        return aio_forward(x, P, n, st)
    if n.kind == 'down':
        return down_forward(x, P, n, st)
    if n.kind == 'glow':
        return glow_forward(x, P, n, st)
    if n.kind == 'dct':
        return dct_forward(x, n)
    if n.kind == 'haar':
        return haar_forward(x, n)
    if n.kind == 'perm':
        perm = torch.as_tensor(permutation(n.dims_in[0], n.extra['seed']))
        return x[:, perm], 0.
    if n.kind == 'reshape':
        return x.reshape(x.shape[0], *n.dims_out), 0.
    if n.kind == 'flatten':
        return x.reshape(x.shape[0], -1), 0.
    raise ValueError(n.kind)


def node_reverse(y, P, n: NodeSpec, st: _State):
    if n.kind == 'aio':
        return aio_reverse(y, P, n, st)
    if n.kind == 'down':
        return down_reverse(y, P, n, st)
    if n.kind == 'glow':
        return glow_reverse(y, P, n, st)
    if n.kind == 'dct':
        return dct_reverse(y, n)
    if n.kind == 'haar':
        return haar_reverse(y, n)
    if n.kind == 'perm':
        perm = permutation(n.dims_in[0], n.extra['seed'])
        inv = np.empty_like(perm)
        inv[perm] = np.arange(len(perm))
        return y[:, torch.as_tensor(inv)], 0.
    if n.kind in ('reshape', 'flatten'):
        return y.reshape(y.shape[0], *n.dims_in), 0.
    raise ValueError(n.kind)


# --------------------------------------------------------------------------- whole network

def inn_forward(spec, P, x, train=False, masks=None, collect=False, teacher=None):
    """z, jac(B,), state.  `teacher`: optional {idx: input} to force a node's input (per-block
    parity).  state.records['module_list.<i>.in' / '.out' / '.jac'] when collect=True."""
    st = _State(train, masks, collect)
    h = x
    jac = torch.zeros(x.shape[0], dtype=x.dtype)
    for n in spec:
        if teacher is not None and n.idx in teacher:
            h = teacher[n.idx].to(x.dtype)
        if collect:
            st.records[n.key + '.in'] = h
        h, j = node_forward(h, P, n, st)
        if collect:
            st.records[n.key + '.out'] = h
            st.records[n.key + '.jac'] = j if torch.is_tensor(j) else torch.full_like(jac, j)
        jac = jac + j
    return h, jac, st


def inn_reverse(spec, P, z, train=False, masks=None):
    st = _State(train, masks, False)
    h = z
    jac = torch.zeros(z.shape[0], dtype=z.dtype)
    for n in reversed(spec):
        h, j = node_reverse(h, P, n, st)
        jac = jac + j
    return h, jac, st


def gmm_head(z, jac, mu, phi, y=None):
    """model.py:113-159 with loss_mean=False: per-sample L_x, logits, L_cNLL, L_y and scalar acc."""
    D = z.shape[1]
    mu2 = mu.reshape(-1, D)
    log_wy = torch.log_softmax(phi, dim=0).view(1, -1)
    zz = -2 * z @ mu2.t() + (z ** 2).sum(dim=1, keepdim=True) + (mu2 ** 2).sum(dim=1).view(1, -1)
    out = {'L_x_tr': (-torch.logsumexp(-0.5 * zz + log_wy, dim=1) - jac) / D,
           'logits_tr': -0.5 * zz}
    if y is not None:
        lw = log_wy.detach()
        out['L_cNLL_tr'] = (0.5 * (zz * y.round()).sum(dim=1) - jac) / D
        out['L_y_tr'] = ((torch.log_softmax(-0.5 * zz + lw, dim=1) - lw) * y).sum(dim=1)
        out['acc_tr'] = (y.argmax(dim=1) == out['logits_tr'].detach().argmax(dim=1)).to(z.dtype).mean()
    return out


def classifier_forward(spec, P, x, y=None, loss_mean=True, train=False, masks=None):
    """GenerativeClassifier.forward (model.py:136-165).  P also holds 'mu' (1,C,D) and 'phi'."""
    z, jac, st = inn_forward(spec, P, x, train=train, masks=masks)
    out = gmm_head(z, jac, P['mu'], P['phi'], y)
    if loss_mean:
        out = {k: v.mean() for k, v in out.items()}
    return out, z, jac, st


def beta_weights(beta, train_nll=True):
    """train.py:88-92."""
    if not train_nll:
        return 0., 1.
    return 2. / (1 + beta), 2. * beta / (1 + beta)


# --------------------------------------------------------------------------- construction

def init_state(cfg: Cfg, seed=0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random-init state with the reference's distributions (inn_architecture.py:58-66,
    all_in_one_block.py:34-53, model.py:46-68).  NOT stream-identical to the reference
    (that needs the reference itself); used for synthetic benchmarks and self-consistency."""
    from scipy.stats import special_ortho_group
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    P: Dict[str, torch.Tensor] = {}
    wi = cfg.weight_init

    def kaiming(*shape):
        fan_in = int(np.prod(shape[1:]))
        return torch.randn(*shape, generator=g) * math.sqrt(2. / fan_in)

    def conv_subnet_init(key, cin, width, cout, relu_first, strided):
        j = 1 if relu_first else 0
        last = j + 6 if strided else j + 7
        P['%s.%d.weight' % (key, j)] = kaiming(width, cin, 3, 3) * wi
        P['%s.%d.weight' % (key, j + 3)] = kaiming(width, width, 3, 3) * wi
        for b in (j + 1, j + 4):
            P['%s.%d.weight' % (key, b)] = torch.ones(width) * wi
            P['%s.%d.bias' % (key, b)] = torch.zeros(width)
            P['%s.%d.running_mean' % (key, b)] = torch.zeros(width)
            P['%s.%d.running_var' % (key, b)] = torch.ones(width)
        P['%s.%d.weight' % (key, last)] = kaiming(cout, width, 1, 1) * wi
        bound = 1. / math.sqrt(width)
        P['%s.%d.bias' % (key, last)] = (torch.rand(cout, generator=g) * 2 - 1) * bound * wi

    def fc_init(key, cin, cout):
        for j, (o, i) in ((0, (cfg.fc_width, cin)), (3, (cout, cfg.fc_width))):
            P['%s.%d.weight' % (key, j)] = kaiming(o, i) * 0.1 * wi
            P['%s.%d.bias' % (key, j)] = (torch.rand(o, generator=g) * 2 - 1) / math.sqrt(i) * wi

    for n in graph_spec(cfg):
        if n.kind == 'aio':
            C = n.dims_in[0]
            c1, c2 = C - C // 2, C // 2
            P[n.key + '.act_norm'] = torch.full((1, C, 1, 1), math.log(cfg.act_norm)) * wi
            P[n.key + '.act_offset'] = torch.zeros(1, C, 1, 1)
            w = special_ortho_group.rvs(C, random_state=rs) if C > 1 else np.ones((1, 1))
            P[n.key + '.w'] = torch.tensor(w, dtype=torch.float32).view(C, C, 1, 1)
            P[n.key + '.w_inv'] = torch.tensor(np.linalg.inv(w), dtype=torch.float32).view(C, C, 1, 1)
            conv_subnet_init(n.key + '.s', c1, n.extra['width'], 2 * c2, n.extra['relu_first'], False)
        elif n.kind == 'down':
            C = n.dims_in[0]
            c1, c2 = C // 2, C - C // 2
            conv_subnet_init(n.key + '.s_hi', c1, n.extra['width'], 8 * c2, False, True)
            conv_subnet_init(n.key + '.s_lo', 4 * c2, n.extra['width'], 8 * c1, False, False)
        elif n.kind == 'glow':
            D = n.dims_in[0]
            l1, l2 = D // 2, D - D // 2
            fc_init(n.key + '.s1', l1, 2 * l2)
            fc_init(n.key + '.s2', l2, 2 * l1)
    D, C = cfg.ndim_tot, cfg.n_classes
    mu = torch.zeros(1, C, D)
    scale = cfg.mu_init / np.sqrt(2 * D // C)
    for k in range(D // C):
        mu[0, :, C * k:C * (k + 1)] = scale * torch.eye(C)
    P['mu'] = mu
    P['phi'] = torch.zeros(C)
    return {k: v.to(dtype) for k, v in P.items()}


def trainable_keys(P) -> List[str]:
    """Keys that model.py:58-59 collects (requires_grad parameters of the INN), then mu, phi."""
    pat = re.compile(r'^module_list\.\d+\.((s|s_hi|s_lo|s1|s2)\.\d+\.(weight|bias)|act_norm|act_offset)$')
    out = [k for k in P if pat.match(k)]
    return out + [k for k in ('mu', 'phi') if k in P]


def sgd_clip_step(P, grads, bufs, groups, max_norm):
    """torch.nn.utils.clip_grad_norm_ (train.py:109) + torch.optim.SGD.step (model.py:97-100).
    groups: list of dict(keys, lr, momentum, weight_decay).  In-place on P / bufs."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
    coef = min(1., float(max_norm / (total + 1e-6)))
    for grp in groups:
        for k in grp['keys']:
            d = grads[k] * coef + grp['weight_decay'] * P[k]
            if grp['momentum'] != 0:
                if k in bufs:
                    bufs[k].mul_(grp['momentum']).add_(d)
                else:
                    bufs[k] = d.clone()
                d = bufs[k]
            P[k].sub_(grp['lr'] * d)
    return total
