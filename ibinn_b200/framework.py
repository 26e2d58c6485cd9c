"""Graph container with the FrEIA-v0.2 API surface the reference uses (FrEIA itself is an
un-vendored dependency, requirements.txt:2; call sites reference inn_architecture.py:122-164,
model.py:141-142,235, train.py:161-163).  Every graph the reference builds is a chain, so the
container is a sequential interpreter; node modules keep the FrEIA protocol
`__init__(dims_in, **kw)`, `forward([x], rev) -> [y]`, `jacobian([x], rev)`, `output_dims(dims)`.
"""
import os

import torch
import torch.nn as nn


class Node:
    def __init__(self, inputs, module_type, module_args, conditions=None, name=None):
        if isinstance(inputs, Node):
            inputs = [(inputs, 0)]
        elif isinstance(inputs, tuple) and len(inputs) == 2 and isinstance(inputs[0], Node):
            inputs = [inputs]
        if conditions:
            raise ValueError('conditional nodes are not used by IB-INN and not supported')
        self.inputs = list(inputs)
        self.module_type, self.module_args, self.name = module_type, module_args, name
        self.module, self.output_dims = None, None

    @property
    def out0(self):
        return (self, 0)

    def build(self):
        dims_in = [src.output_dims[idx] for src, idx in self.inputs]
        self.module = self.module_type(dims_in, **self.module_args)
        self.output_dims = [tuple(d) for d in self.module.output_dims(dims_in)]


class InputNode(Node):
    def __init__(self, *dims, name='node'):
        self.inputs, self.name, self.module = [], name, None
        self.output_dims = [tuple(dims)]

    def build(self):
        pass


class OutputNode(Node):
    def __init__(self, inputs, name='node'):
        if isinstance(inputs, Node):
            inputs = [(inputs, 0)]
        elif isinstance(inputs, tuple) and len(inputs) == 2 and isinstance(inputs[0], Node):
            inputs = [inputs]
        self.inputs, self.name, self.module, self.output_dims = list(inputs), name, None, None

    def build(self):
        self.output_dims = [src.output_dims[idx] for src, idx in self.inputs]


class ReversibleGraphNet(nn.Module):
    """`inn(x)`, `inn(z, rev=True)`, `inn.log_jacobian(run_forward=False)`; `module_list[i]` follows
    `node_list[i]` (None for the input / output nodes), `tmp_var_<i>` buffers as in FrEIA so that
    reference checkpoints and train.py:161-163 keep working."""

    def __init__(self, node_list, ind_in=None, ind_out=None, verbose=True):
        super().__init__()
        for k, n in enumerate(node_list):
            if k > 0 and not isinstance(n, InputNode) and [src for src, _ in n.inputs] != [node_list[k - 1]]:
                raise ValueError('ibinn_b200 supports chain graphs only (every IB-INN graph is one)')
            n.build()
        self.node_list = node_list
        self.module_list = nn.ModuleList([n.module for n in node_list])
        for k in range(len(node_list)):
            self.register_buffer('tmp_var_%d' % k, torch.zeros(1))
        self._jacs, self._rev = [], False
        # runs of consecutive AIO_Blocks of one resolution stage go through ONE fused launch when the kernel supports the
        # geometry (functions.AIOStageFn); `fused_stages = False` (or IBINN_FUSED=0) keeps one node = one autograd function
        self.fused_stages = os.environ.get('IBINN_FUSED', '1') != '0'
        self._runs = None
        # engine.TrainStep: cut the autograd graph in front of node `cut_before` (the first fully connected block), so that the
        # backward pass can run in two pieces with the gradient all-reduce of the tail started in between.  `_cut` then holds
        # [(tensor below the cut, detached leaf above it), ...] for the activation and for the summed log-dets of the nodes below.
        self.cut_before = None
        self._cut = []
        if verbose:
            for n in node_list:
                print(n.name, n.output_dims)

    def _find_runs(self):
        """{first index: AioStagePlan} for every maximal run (>= 1) of consecutive AIO_Blocks with the same geometry"""
        from . import ops
        from .all_in_one_block import AIO_Block
        runs, k, n = {}, 0, len(self.module_list)
        while k < n:
            m = self.module_list[k]
            if isinstance(m, AIO_Block) and not m.act_norm_trigger:
                j = k + 1
                dims = self.node_list[k].output_dims
                while (j < n and isinstance(self.module_list[j], AIO_Block) and self.node_list[j].output_dims == dims
                       and not self.module_list[j].act_norm_trigger and self.module_list[j].clamp == m.clamp
                       and self.module_list[j].s.conv1.weight.shape == m.s.conv1.weight.shape):
                    j += 1
                runs[k] = ops.AioStagePlan([self.module_list[i] for i in range(k, j)])
                k = j
            else:
                k += 1
        return runs

    def forward(self, x, c=None, rev=False, intermediate_outputs=False):
        from . import functions as Fn
        n = len(self.node_list)
        self._node_inputs, self._jacs, self._rev = {}, [], rev
        h = x
        if rev:
            for k in reversed(range(n)):
                m = self.module_list[k]
                if m is None:
                    continue
                self._node_inputs[k] = h
                h = m([h], rev=True)[0]
                self._jacs.append((k, h))
            return h
        if self.fused_stages and self._runs is None:
            self._runs = self._find_runs()
        self._cut = []
        k = 0
        while k < n:
            m = self.module_list[k]
            if m is None:
                k += 1
                continue
            if k == self.cut_before and torch.is_grad_enabled() and h.requires_grad:
                hd = h.detach().requires_grad_(True)
                self._cut.append((h, hd))
                h = hd
            plan = self._runs.get(k) if self.fused_stages else None
            if plan is not None and (m.training or not torch.is_grad_enabled()) and plan.supported(h, m.training):
                # the whole run in one launch; its log-det is reported by the first block of the run
                self._node_inputs[k] = h
                h, jac = Fn.AIOStageFn.apply(h, plan, *Fn.aio_stage_params(plan))
                m.last_jac = jac
                for blk in plan.blocks[1:]:
                    blk.last_jac = None
                k += len(plan.blocks)
                self._jacs.append((k - 1, h))
                continue
            self._node_inputs[k] = h
            h = m([h], rev=False)[0]
            self._jacs.append((k, h))
            k += 1
        return h

    def release_graph(self):
        """Drop every reference to autograd graphs of earlier passes (cached per-node log-dets, node
        inputs).  engine.TrainStep calls this before CUDA-graph capture: a live old graph keeps the
        parameters' gradient accumulators -- and the stream they were created on -- alive."""
        self._node_inputs, self._jacs, self._cut = {}, [], []
        for m in self.module_list:
            if m is not None and hasattr(m, 'last_jac') and torch.is_tensor(m.last_jac):
                m.last_jac = None

    def log_jacobian(self, x=None, c=None, rev=False, run_forward=True, intermediate_outputs=False):
        if run_forward:
            self.forward(x, rev=rev)
        const, tens, below = 0., [], []
        cutting = bool(self._cut) and not rev
        for k, h in self._node_inputs.items():
            j = self.module_list[k].jacobian([h], rev=rev)
            if torch.is_tensor(j):
                (below if cutting and k < self.cut_before else tens).append(j)
            else:
                const = const + j
        if below:
            jb = below[0] if len(below) == 1 else torch.stack(below).sum(0)
            if jb.requires_grad:
                jd = jb.detach().requires_grad_(True)
                self._cut.append((jb, jd))
                jb = jd
            tens.append(jb)
        if not tens:
            return const
        if len(tens) == 1:
            return tens[0] + const
        return torch.stack(tens).sum(0) + const          # two kernels instead of one add per node

    jacobian = log_jacobian
