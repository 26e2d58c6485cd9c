"""Training-step engine around GenerativeClassifier: flat parameter / gradient buffers, fused
clip + SGD, whole-step CUDA-graph capture, and batch-sharded data parallelism over NCCL.

Reference semantics reproduced (train.py:88-111): loss = beta_x * L_x - beta_y * L_y with
beta_x = 2/(1+beta), beta_y = 2 beta/(1+beta); clip_grad_norm_ over *all* trainable tensors
(INN, mu, phi) at `clip_grad_norm`; torch.optim.SGD per parameter group (model.py:73-100).
Multi-GPU: every rank runs the reference's single-GPU step on its own batch
with local BatchNorm statistics; the only exchange is one all-reduce of the flat gradient.
"""
import os

import torch
import torch.distributed as dist

from . import _lib, ops


class FlatParams:
    """Moves every trainable tensor of the model into one flat fp32 buffer (segments per optimiser
    group, 16-byte aligned) with a matching flat gradient and momentum buffer.  Parameters keep
    their identity (`p.data` becomes a view), so state_dict / checkpoints / the torch optimiser
    object are unaffected; `p._ibinn_grad` tells the backward kernels to accumulate in place."""

    def __init__(self, model):
        self.model = model
        dev = model.mu.device
        groups = model.optimizer.param_groups
        in_group = set()
        segs = []
        for gi, g in enumerate(groups):
            ps = [p for p in g['params'] if p.requires_grad]
            in_group.update(id(p) for p in ps)
            segs.append((gi, ps))
        rest = [p for p in model.trainable_params if id(p) not in in_group and p.requires_grad]
        if rest:
            segs.append((None, rest))      # clipped with the others, but no optimiser group updates them
        total, layout = 0, []
        for gi, ps in segs:
            start = total
            offs = []
            for p in ps:
                offs.append(total)
                total += (p.numel() + 3) // 4 * 4
            layout.append((gi, ps, offs, start, total))
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.buf = torch.zeros(total, dtype=torch.float32, device=dev)
        self.segments = []
        for gi, ps, offs, start, end in layout:
            for p, o in zip(ps, offs):
                n = p.numel()
                self.flat[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.flat[o:o + n].view(p.shape)
                p._ibinn_grad = self.grad[o:o + n].view(p.shape)
                p.grad = p._ibinn_grad
            hyper = torch.zeros(3, dtype=torch.float32, device=dev) if gi is not None else None
            self.segments.append(dict(group=gi, start=start, end=end, hyper=hyper, host=None))
        self.sync_hyper()

    def sync_hyper(self):
        """Copy lr / momentum / weight_decay of the torch optimiser groups to the device (call after a
        scheduler step; a captured graph reads them from there)."""
        for s in self.segments:
            if s['group'] is None:
                continue
            g = self.model.optimizer.param_groups[s['group']]
            vals = (float(g['lr']), float(g.get('momentum', 0.)), float(g.get('weight_decay', 0.)))
            if vals != s['host']:
                s['hyper'].copy_(torch.tensor(vals, dtype=torch.float32))
                s['host'] = vals

    def zero_grad(self):
        self.grad.zero_()

    def _momentum_views(self):
        """(parameter, view of `buf`) for every parameter of an optimiser group with momentum"""
        for s in self.segments:
            if s['group'] is None:
                continue
            g = self.model.optimizer.param_groups[s['group']]
            if not float(g.get('momentum', 0.)):
                continue
            for p in g['params']:
                if p.requires_grad:
                    o = p.data.data_ptr() - self.flat.data_ptr()
                    yield p, self.buf[o // 4:o // 4 + p.numel()].view(p.shape)

    def export_momentum(self):
        """Make the torch optimiser's state show the fused step's momentum (views of `buf`), so that
        `optimizer.state_dict()` -- what the reference checkpoints (model.py:237-241) -- is complete."""
        for p, v in self._momentum_views():
            self.model.optimizer.state[p]['momentum_buffer'] = v

    def import_momentum(self, opt_state):
        """inverse of export_momentum from a `torch.optim.SGD.state_dict()`; entries without a buffer are left at zero"""
        order = [p for g in self.model.optimizer.param_groups for p in g['params']]
        ids = [i for g in opt_state.get('param_groups', []) for i in g['params']]
        by_param = {id(p): opt_state['state'].get(i) for p, i in zip(order, ids)}
        for p, v in self._momentum_views():
            st = by_param.get(id(p))
            if st is not None and st.get('momentum_buffer') is not None:
                v.copy_(st['momentum_buffer'])

    def clip_and_step(self, max_norm, grad_scale=1.0):
        """Fused clip_grad_norm_ + SGD.  Returns the device tensor [total_norm, clip_coef]."""
        # ||grad_scale * g|| <= max_norm  <=>  ||g|| <= max_norm / grad_scale
        nc = ops.grad_norm_clip(self.grad, max_norm / grad_scale)
        for s in self.segments:
            if s['group'] is None:
                continue
            sl = slice(s['start'], s['end'])
            ops.sgd_step_(self.flat[sl], self.grad[sl], self.buf[sl], s['hyper'], nc[1:], grad_scale)
        return nc


class WeightPrep:
    """Tensor-core operand images (hi/lo tf32 split, per-tap K-major) of every stride-1 subnet filter, in
    both orientations (forward / dgrad), refreshed by ONE launch per step instead of one per convolution."""

    def __init__(self, model):
        import ctypes
        import numpy as np
        from . import _lib
        from .subnets import ConvSubnet
        dev = model.mu.device
        descs, self.params, self._keep = [], [], []
        for m in model.inn.modules():
            if not isinstance(m, ConvSubnet):
                continue
            for conv in (m.conv1, m.conv2, m.conv3):
                w = conv.weight
                entry = {'valid': False}
                for tr in (False, True):
                    n = ops.conv_wprep_floats(w, tr, 8, 8)
                    if n <= 0:
                        continue
                    buf = torch.empty(n, dtype=torch.float32, device=dev)
                    entry[tr] = buf
                    co, ci, ks, _ = w.shape
                    d = _lib.WprepDesc(w.data_ptr(), buf.data_ptr(), co if tr else ci, ci if tr else co, ks, int(tr))
                    descs.append(bytes(d))
                if False in entry and True in entry:
                    w._ibinn_wprep = entry
                    self.params.append(w)
        from .all_in_one_block import AIO_Block
        for m in model.inn.modules():
            # the frozen soft-permutation matrices as 1x1 filters: operand of the fused block kernel (block_tc.cu)
            if isinstance(m, AIO_Block):
                w = m.w
                n = ops.conv_wprep_floats(w, False, 8, 8)
                if n > 0:
                    buf = torch.empty(n, dtype=torch.float32, device=dev)
                    co, ci, ks, _ = w.shape
                    descs.append(bytes(_lib.WprepDesc(w.data_ptr(), buf.data_ptr(), ci, co, ks, 0)))
                    w._ibinn_wprep = {'valid': False, False: buf}
                    self.params.append(w)
        self.n = len(descs)
        raw = np.frombuffer(b''.join(descs), dtype=np.uint8).copy() if descs else np.zeros(1, np.uint8)
        self.table = torch.from_numpy(raw).to(dev)

    def refresh(self):
        if self.n:
            ops.wprep_many(self.table, self.n, self.table)
            for w in self.params:
                w._ibinn_wprep['valid'] = True

    def invalidate(self):
        for w in self.params:
            w._ibinn_wprep['valid'] = False


class MaskPool:
    """All dropout masks of one step (Dropout2d (B,width) per stage-2 subnet, Dropout (B,fc_width) per FC subnet) live
    in one buffer per drop probability and are redrawn with three in-place launches per probability."""

    def __init__(self, model, batch_size):
        from .subnets import ConvSubnet, FCSubnet
        dev = model.mu.device
        by_p = {}
        for m in model.inn.modules():
            if isinstance(m, (ConvSubnet, FCSubnet)) and m.p_drop > 0.:
                width = m.conv1.weight.shape[0] if isinstance(m, ConvSubnet) else m.lin1.weight.shape[0]
                by_p.setdefault(m.p_drop, []).append((m, batch_size * width, (batch_size, width)))
        self.pools = []
        for p, items in by_p.items():
            buf = torch.zeros(sum(n for _, n, _ in items), dtype=torch.float32, device=dev)
            o = 0
            for m, n, shape in items:
                m._ibinn_mask = buf[o:o + n].view(shape)
                o += n
            self.pools.append((p, buf))
        self.modules = [m for items in by_p.values() for m, _, _ in items]

    def redraw(self):
        for p, buf in self.pools:
            buf.uniform_()
            buf.ge_(p)
            buf.mul_(1. / (1. - p))
        for m in self.modules:
            m._ibinn_mask_fresh = True


def beta_weights(beta, train_nll=True):
    """reference train.py:88-92"""
    if not train_nll:
        return 0., 1.
    return 2. / (1. + beta), 2. * beta / (1. + beta)


class TrainStep:
    """One optimisation step: zero grads -> forward -> IB loss -> backward -> [all-reduce] -> clip -> SGD.

    `step(x, y)` takes device tensors or pinned host tensors (copied into static device buffers on the
    step's stream) and returns the dict of device scalars the reference logs (no host sync).  With
    `use_graph` the whole step is captured once into CUDA graphs and replayed."""

    def __init__(self, model, batch_size, beta=None, clip=None, use_graph=True, process_group=None, class_nll=None, replicas_only=False):
        if not isinstance(model.optimizer, torch.optim.SGD):
            raise NotImplementedError('the fused step implements the reference default (SGD); use the plain torch loop otherwise')
        t = model.args['training']
        self.model = model
        self.beta = float(t['beta_IB']) if beta is None else float(beta)
        self.clip = float(t['clip_grad_norm']) if clip is None else float(clip)
        ab = model.args['ablations']
        self.class_nll = eval(ab['class_NLL']) if class_nll is None else class_nll
        self.bx, self.by = beta_weights(self.beta, not eval(ab['no_NLL_term']))
        self.pg = process_group
        # replicas_only: an independent training run per process (the beta sweep, BASELINE.json configs[4]): no gradient exchange
        self.world = dist.get_world_size(process_group) if (process_group is not None or dist.is_initialized()) and not replicas_only else 1
        self.flat = FlatParams(model)
        model._ibinn_flat = self.flat          # model.save / model.load checkpoint the momentum through it
        dev = model.mu.device
        if self.world > 1:
            self._broadcast_replicas()
        self.wprep = WeightPrep(model) if dev.type == 'cuda' else None      # after FlatParams: the filters have their final addresses
        self.masks = MaskPool(model, batch_size)
        self.x = torch.zeros((batch_size,) + tuple(model.dims), dtype=torch.float32, device=dev)
        self.y = torch.zeros((batch_size, model.n_classes), dtype=torch.float32, device=dev)
        self.use_graph = bool(use_graph) and dev.type == 'cuda'
        n_side = int(os.environ.get('IBINN_WGRAD_STREAMS', '1'))
        self.side_streams = [torch.cuda.Stream() for _ in range(n_side)] if dev.type == 'cuda' and n_side > 0 else None
        # layers per record reduction on a third stream; 0 = one reduction after the backward pass.  Measured: 8 / 16 / 32 are
        # within noise of 0 (10.76 / 10.90 / 10.81 vs 10.80 ms per step), so the simpler schedule is the default.
        self.reduce_group = int(os.environ.get('IBINN_WGRAD_REDUCE_GROUP', '0'))
        self.reduce_stream = torch.cuda.Stream() if self.side_streams and self.reduce_group > 0 else None
        self._pinned_groups = []
        # Gradient exchange: the flat gradient is laid out in parameter order, so everything from the first
        # fully connected coupling block to the end (84 % of the payload: the FC subnets, then mu and phi) is complete as soon as
        # that block's backward has run.  Its all-reduce starts there, on a side stream, and runs under the backward pass of the 56
        # convolutional blocks; only the remaining 16 % is reduced after the backward pass.
        self.early_start, self._first_glow, self.comm_stream, self._early_issued, self._glow_idx = None, None, None, False, None
        force_split = os.environ.get('IBINN_FORCE_SPLIT_BACKWARD', '0') == '1'      # one-GPU validation of the two-piece backward
        if (self.world > 1 or force_split) and os.environ.get('IBINN_OVERLAP_ALLREDUCE', '1') != '0':
            from .modules import GLOWCouplingBlock
            glows = [m for m in model.inn.modules() if isinstance(m, GLOWCouplingBlock)]
            if glows:
                first = glows[0]
                p0 = next(p for p in first.parameters() if p.requires_grad)
                off = (p0.data.data_ptr() - self.flat.flat.data_ptr()) // 4
                tail = {id(p) for gl in glows for p in gl.parameters()} | {id(model.mu), id(model.phi)}
                after = [p for g in model.optimizer.param_groups for p in g['params']
                         if p.requires_grad and (p.data.data_ptr() - self.flat.flat.data_ptr()) // 4 >= off]
                if all(id(p) in tail for p in after) and 0 < off < self.flat.grad.numel():
                    self.early_start, self._first_glow = int(off), first
                    self._glow_idx = [i for i, mm in enumerate(model.inn.module_list) if mm is first][0]
                    self.comm_stream = torch.cuda.Stream() if dev.type == 'cuda' else None
        self._g1 = self._g1b = self._g2 = None
        self._one_graph = False
        self.out = None
        self.launches = None

    def _broadcast_replicas(self):
        """Every rank starts from rank 0's model: trainable tensors (one flat buffer), the frozen soft-permutation
        matrices (drawn from numpy's global RNG, reference all_in_one_block.py:43) and the BatchNorm buffers."""
        src = dist.get_global_rank(self.pg, 0) if self.pg is not None else 0
        dist.broadcast(self.flat.flat, src=src, group=self.pg)
        with torch.no_grad():
            for p in self.model.parameters():
                if not p.requires_grad:
                    dist.broadcast(p.data, src=src, group=self.pg)
            for name, b in self.model.named_buffers():
                if 'tmp_var' not in name and b is not None and b.numel():
                    dist.broadcast(b, src=src, group=self.pg)

    def _snapshot(self):
        """everything a training step mutates besides the static input buffers"""
        bufs = [b for n, b in self.model.named_buffers() if 'tmp_var' not in n and b is not None]
        snap = dict(flat=self.flat.flat.clone(), buf=self.flat.buf.clone(), bufs=[(b, b.clone()) for b in bufs],
                    cpu_rng=torch.get_rng_state())
        if self.flat.flat.is_cuda:
            snap['cuda_rng'] = torch.cuda.get_rng_state(self.flat.flat.device)
        return snap

    def _restore(self, snap):
        with torch.no_grad():
            self.flat.flat.copy_(snap['flat'])
            self.flat.buf.copy_(snap['buf'])
            for b, v in snap['bufs']:
                b.copy_(v)
        torch.set_rng_state(snap['cpu_rng'])
        if 'cuda_rng' in snap:
            torch.cuda.set_rng_state(snap['cuda_rng'], self.flat.flat.device)
        if self.wprep is not None:
            self.wprep.invalidate()

    # -- pieces ---------------------------------------------------------------------------------
    def _fwd_bwd_a(self):
        """forward, loss, and the backward pass down to the input of the first fully connected coupling block: after this the
        tail of the flat gradient (FC subnets, mu, phi) is final"""
        self.flat.zero_grad()
        if self.wprep is not None:
            self.wprep.refresh()
        if self.model.inn.training:
            self.masks.redraw()
        self.model.inn.cut_before = self._glow_idx if self.early_start is not None else None
        try:
            losses = self.model(self.x, self.y)
        finally:
            self.model.inn.cut_before = None
        if self.class_nll:
            loss = 2. * losses['L_cNLL_tr']
        else:
            loss = self.bx * losses['L_x_tr'] - self.by * losses['L_y_tr']
        ops.DEFERRED_WGRAD = self._pending = []
        ops.DEFERRED_AIO = self._pending_aio = [] if self.x.is_cuda else None
        ops.WGRAD_STREAMS = self.side_streams
        self._grouped = self.side_streams is not None and self.reduce_stream is not None
        if self._grouped:
            ops.WGRAD_REDUCE = dict(stream=self.reduce_stream, group=self.reduce_group, pinned=self._pinned_groups, upto=0, keep=[])
        self.out = {k: losses[k].detach() for k in ('L_x_tr', 'L_y_tr', 'acc_tr', 'L_cNLL_tr')}
        self._split = None
        try:
            cut = self.model.inn._cut
            if cut:
                # the graph above the cut (head, FC block) is complete in itself: its backward leaves the gradients of the detached
                # leaves, and every parameter gradient of the tail in flat.grad
                loss.backward()
                if self.side_streams:                  # the FC weight gradients were forked onto the side stream(s): the tail of the
                    cur = torch.cuda.current_stream()  # flat gradient is final (and this piece's graph closed) only after they join
                    for s_ in self.side_streams:
                        cur.wait_stream(s_)
                self._split = ([t for t, _ in cut], [d.grad for _, d in cut])
                self.model.inn._cut = []
            else:
                self._split = ([loss], None)
        except BaseException:
            self._clear_backward_state()
            raise

    def _fwd_bwd_b(self):
        """the rest of the backward pass (DCT, the convolutional blocks) and the deferred record reductions"""
        t, g = self._split
        self._split = None
        try:
            torch.autograd.backward(t, grad_tensors=g)
            if self._grouped:
                ops.flush_wgrad_group(final=True)
                self._reduce_keep = ops.WGRAD_REDUCE['keep']
            ops.join_wgrad_streams()
        finally:
            self._clear_backward_state()
        if self._pending_aio:
            self._pinned_aio = ops.flush_deferred_aio(self._pending_aio, self.x, getattr(self, '_pinned_aio', None))
        if not self._grouped:
            self._pinned = ops.flush_deferred_wgrad(self._pending, self.x, getattr(self, '_pinned', None))
        # self._pending / self._pending_aio keep the record workspaces alive until the reductions have run

    @staticmethod
    def _clear_backward_state():
        ops.DEFERRED_WGRAD = None
        ops.DEFERRED_AIO = None
        ops.WGRAD_STREAMS = None
        ops.WGRAD_REDUCE = None

    def _early_allreduce(self):
        """between the two pieces of the backward pass: reduce the (final) tail of the flat gradient on the communication
        stream, under the backward pass of the convolutional blocks"""
        self._early_issued = False
        if self.world <= 1 or self.early_start is None:
            return
        tail = self.flat.grad[self.early_start:]
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(tail, op=dist.ReduceOp.SUM, group=self.pg)
        else:
            dist.all_reduce(tail, op=dist.ReduceOp.SUM, group=self.pg)
        self._early_issued = True

    def _fwd_bwd(self):
        self._fwd_bwd_a()
        self._early_allreduce()
        self._fwd_bwd_b()

    def _reduce(self):
        if self.world > 1:
            if self._early_issued:
                dist.all_reduce(self.flat.grad[:self.early_start], op=dist.ReduceOp.SUM, group=self.pg)
                if self.comm_stream is not None:
                    torch.cuda.current_stream().wait_stream(self.comm_stream)
            else:
                dist.all_reduce(self.flat.grad, op=dist.ReduceOp.SUM, group=self.pg)

    def _update(self):
        # the all-reduce sums; the mean gradient is grad / world (folded into the clip and the update)
        nc = self.flat.clip_and_step(self.clip, grad_scale=1.0 / self.world)
        if self.wprep is not None:
            self.wprep.invalidate()         # filters changed: images are stale until the next refresh
        self.out['grad_norm'] = nc[0] / self.world if self.world > 1 else nc[0]

    def _eager(self):
        self._fwd_bwd()
        self._reduce()
        self._update()

    def capture(self, warmup=3):
        """Warm up eagerly (lazy inits, allocator), then capture.  The all-reduce stays outside the
        graphs (two graphs around one NCCL call) so that capture never depends on the NCCL build.
        The warm-up steps are real optimisation steps on whatever batch is loaded, so parameters, momentum, BatchNorm
        buffers and the RNG state are put back afterwards: capturing has no training side effect."""
        L = _lib.lib()
        snap = self._snapshot()
        # Warm up and capture on ONE side stream: autograd remembers the stream on which a parameter's
        # gradient accumulator was created and syncs it with the backward's stream, which is illegal
        # across a capture boundary.
        self.out = None
        if hasattr(self.model.inn, 'release_graph'):
            self.model.inn.release_graph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore(snap)
        torch.cuda.synchronize()
        n0 = L.ibinn_launch_count()
        self._one_graph = False
        if self.world > 1 and os.environ.get('IBINN_NCCL_IN_GRAPH', '0') == '1':
            # opt-in: the whole step, collectives included, as ONE graph (needs an NCCL build that supports stream capture)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                self._fwd_bwd()
                self._reduce()
                self._update()
            self._g1, self._g1b, self._g2, self._one_graph = g, None, None, True
        else:
            # the collectives stay outside the graphs: [forward + head / FC backward] -> early all-reduce of the gradient tail on the
            # communication stream -> [rest of the backward pass] -> all-reduce of the rest -> [clip + SGD]
            self._g1 = torch.cuda.CUDAGraph()
            if self.early_start is None:
                with torch.cuda.graph(self._g1, stream=s):          # nothing to exchange in between: one graph for forward + backward
                    self._fwd_bwd_a()
                    self._fwd_bwd_b()
                self._g1b = None
            else:
                with torch.cuda.graph(self._g1, stream=s):
                    self._fwd_bwd_a()
                self._g1b = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._g1b, pool=self._g1.pool(), stream=s):
                    self._fwd_bwd_b()
            self._g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._g2, pool=self._g1.pool(), stream=s):
                self._update()
        n1 = L.ibinn_launch_count()
        self.launches = int(n1 - n0)
        torch.cuda.synchronize()

    # -- public ---------------------------------------------------------------------------------
    def load(self, x, y):
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)

    def run(self):
        """one step on whatever is in the static buffers"""
        self.flat.sync_hyper()
        if self.use_graph:
            if self._g1 is None:
                self.capture()
            self._g1.replay()
            if not self._one_graph:
                if self._g1b is not None:
                    self._early_allreduce()
                    self._g1b.replay()
                self._reduce()
                self._g2.replay()
        else:
            self._eager()
        return self.out

    def step(self, x, y):
        self.load(x, y)
        return self.run()


class ScoreStep:
    """Forward-only pass through the public model API under `torch.no_grad()`: log-likelihood / OoD scoring at large batch
    (reference evaluation/ood.py:43: `inn(x, y=None, loss_mean=False)` in eval mode) or the forward + IB loss of a training
    batch (`train_mode=True`, `with_labels=True`).  Static input buffers, tensor-core filter images prepared once, and the
    whole pass captured into one CUDA graph; `step(x[, y])` returns the model's dict of device tensors (no host sync)."""

    def __init__(self, model, batch_size, with_labels=False, loss_mean=False, train_mode=False, use_graph=True):
        self.model = model
        dev = model.mu.device
        self.with_labels, self.loss_mean, self.train_mode = with_labels, loss_mean, train_mode
        self.x = torch.zeros((batch_size,) + tuple(model.dims), dtype=torch.float32, device=dev)
        self.y = torch.zeros((batch_size, model.n_classes), dtype=torch.float32, device=dev) if with_labels else None
        self.use_graph = bool(use_graph) and dev.type == 'cuda'
        self.wprep = WeightPrep(model) if dev.type == 'cuda' else None
        self._graph, self.out, self.launches = None, None, None
        self.refresh()

    def refresh(self):
        """call after the parameters changed: the captured graph reads the filter images, not the raw filters"""
        if self.wprep is not None:
            self.wprep.refresh()

    def _forward(self):
        was_training = self.model.inn.training
        self.model.inn.train(self.train_mode)
        try:
            with torch.no_grad():
                out = self.model(self.x, self.y, loss_mean=self.loss_mean)
        finally:
            self.model.inn.train(was_training)
        if hasattr(self.model.inn, 'release_graph'):
            self.model.inn.release_graph()
        self.out = out

    def capture(self, warmup=2):
        L = _lib.lib()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._forward()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = L.ibinn_launch_count()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=s):
            self._forward()
        self.launches = int(L.ibinn_launch_count() - n0)
        torch.cuda.synchronize()

    def load(self, x, y=None):
        self.x.copy_(x, non_blocking=True)
        if self.with_labels:
            self.y.copy_(y, non_blocking=True)

    def run(self):
        if self.use_graph:
            if self._graph is None:
                self.capture()
            self._graph.replay()
        else:
            self._forward()
        return self.out

    def step(self, x, y=None):
        self.load(x, y)
        return self.run()
