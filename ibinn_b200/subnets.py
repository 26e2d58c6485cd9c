"""Subnet containers.  They hold the *same* nn.Module tree as the reference's subnet constructors
(reference inn_architecture.py:68-120: state_dict keys `0.weight`, `1.running_mean`, ... and real
nn.BatchNorm2d children, which evaluation/__init__.py:25-38 walks), but calling them runs the
ibinn_b200 CUDA kernels instead of the layers one by one."""
import torch.nn as nn

from . import functions as Fn


class ConvSubnet(nn.Sequential):
    """basic_residual_block (strided=False) / strided_residual_block (strided=True)."""

    def __init__(self, cin, cout, width, dropout, relu_first, strided, bn_args):
        layers = [nn.ReLU()] if relu_first else []
        layers += [nn.Conv2d(cin, width, kernel_size=3, padding=1, bias=False),
                   nn.BatchNorm2d(width, **bn_args),
                   nn.ReLU(inplace=not strided),
                   nn.Conv2d(width, width, kernel_size=3, stride=2 if strided else 1, padding=1, bias=False),
                   nn.BatchNorm2d(width, **bn_args),
                   nn.ReLU(inplace=True)]
        if not strided:
            layers.append(nn.Dropout2d(p=dropout))
        layers.append(nn.Conv2d(width, cout, 1, padding=0))
        super().__init__(*layers)
        self.relu_first, self.strided = bool(relu_first), bool(strided)
        self._j = 1 if relu_first else 0

    conv1 = property(lambda self: self[self._j])
    bn1 = property(lambda self: self[self._j + 1])
    conv2 = property(lambda self: self[self._j + 3])
    bn2 = property(lambda self: self[self._j + 4])
    conv3 = property(lambda self: self[len(self) - 1])

    @property
    def p_drop(self):
        return 0. if self.strided else float(self[self._j + 6].p)

    def forward(self, x, mask=None):
        return Fn.ConvSubnetFn.apply(x, self, mask, *Fn.conv_subnet_params(self))


class FCSubnet(nn.Sequential):
    """fc_constr: Linear -> ReLU -> Dropout -> Linear."""

    def __init__(self, c_in, c_out, width, dropout):
        super().__init__(nn.Linear(c_in, width), nn.ReLU(), nn.Dropout(p=dropout), nn.Linear(width, c_out))

    lin1 = property(lambda self: self[0])
    lin2 = property(lambda self: self[3])

    @property
    def p_drop(self):
        return float(self[2].p)

    def forward(self, x, mask=None):
        return Fn.FCSubnetFn.apply(x, self, mask, *Fn.fc_subnet_params(self))
