"""Stand-in for FrEIA.framework (v0.2 API surface used by the reference).

Call sites that define what must exist: /root/reference/inn_architecture.py:122-164,
/root/reference/model.py:141-142,235, /root/reference/train.py:161-163.
Every graph the reference builds is a chain, so a sequential interpreter is
sufficient.  TEST INFRASTRUCTURE ONLY (see package docstring).
"""
import torch
import torch.nn as nn


class _Port(tuple):
    """`node.out0` in FrEIA is the pair (node, output index)."""


class Node:
    def __init__(self, inputs, module_type, module_args, conditions=None, name=None):
        if isinstance(inputs, Node):
            inputs = [(inputs, 0)]
        elif isinstance(inputs, tuple) and len(inputs) == 2 and isinstance(inputs[0], Node):
            inputs = [inputs]
        self.inputs = list(inputs)
        self.module_type = module_type
        self.module_args = module_args
        self.name = name
        self.module = None
        self.output_dims = None

    @property
    def out0(self):
        return _Port((self, 0))

    def build(self):
        dims_in = [src.output_dims[idx] for src, idx in self.inputs]
        self.module = self.module_type(dims_in, **self.module_args)
        self.output_dims = [tuple(d) for d in self.module.output_dims(dims_in)]


class InputNode(Node):
    def __init__(self, *dims, name='node'):
        self.inputs = []
        self.name = name
        self.module = None
        self.output_dims = [tuple(dims)]

    def build(self):
        pass


class OutputNode(Node):
    def __init__(self, inputs, name='node'):
        if isinstance(inputs, Node):
            inputs = [(inputs, 0)]
        elif isinstance(inputs, tuple) and len(inputs) == 2 and isinstance(inputs[0], Node):
            inputs = [inputs]
        self.inputs = list(inputs)
        self.name = name
        self.module = None
        self.output_dims = None

    def build(self):
        self.output_dims = [src.output_dims[idx] for src, idx in self.inputs]


class ReversibleGraphNet(nn.Module):
    def __init__(self, node_list, ind_in=None, ind_out=None, verbose=True):
        super().__init__()
        self.node_list = node_list
        for n in node_list:
            n.build()
        # positions follow node_list; Input/Output nodes carry no module
        self.module_list = nn.ModuleList([n.module for n in node_list])
        self._node_inputs = {}
        for k in range(len(node_list)):
            self.register_buffer('tmp_var_%d' % k, torch.zeros(1))

    def forward(self, x, c=None, rev=False, intermediate_outputs=False):
        order = range(len(self.node_list))
        if rev:
            order = reversed(order)
        self._node_inputs = {}
        h = x
        for k in order:
            m = self.module_list[k]
            if m is None:
                continue
            self._node_inputs[k] = h
            h = m([h], rev=rev)[0]
        return h

    def log_jacobian(self, x=None, c=None, rev=False, run_forward=True, intermediate_outputs=False):
        if run_forward:
            self.forward(x, rev=rev)
        total = 0.
        for k, h in self._node_inputs.items():
            total = total + self.module_list[k].jacobian([h], rev=rev)
        return total

    jacobian = log_jacobian
