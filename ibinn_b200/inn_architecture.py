"""`constuct_inn` / `construct_irevnet` with the reference's signature, .ini keys, node names, module
tree and initialisation order (reference inn_architecture.py:30-164), built from ibinn_b200 nodes.
Constructed under the same torch / numpy seeds it yields the same parameters as the reference."""
from functools import partial

import torch
import torch.nn as nn

from . import framework as Ff
from . import modules as Fm
from .all_in_one_block import AIO_Block
from .dct_transform import DCTPooling2d
from .downsampling_coupling_block import DownsampleCouplingBlock
from .subnets import ConvSubnet, FCSubnet


def construct_irevnet(classifier):
    inn = constuct_inn(classifier)
    projection_layer = nn.Linear(classifier.ndim_tot, classifier.n_classes)

    class RevNetWrapper(nn.Module):
        def __init__(self, inn, projection_layer):
            super().__init__()
            self.inn = inn
            self.proj = projection_layer

        def forward(self, x):
            return self.proj(self.inn(x))

    return RevNetWrapper(inn, projection_layer)


def _weights_init(m):
    # reference inn_architecture.py:58-66
    if type(m) == nn.Conv2d:
        torch.nn.init.kaiming_normal_(m.weight)
    if type(m) == nn.BatchNorm2d:
        m.weight.data.fill_(1)
        m.bias.data.zero_()
    if type(m) == nn.Linear:
        torch.nn.init.kaiming_normal_(m.weight)
        m.weight.data *= 0.1


def constuct_inn(classifier, verbose=False):
    m = classifier.args['model']
    fc_width = int(m['fc_width'])
    n_coupling_blocks_fc = int(m['n_coupling_blocks_fc'])
    use_dct = eval(m['dct_pooling'])
    conv_widths = eval(m['conv_widths'])
    n_coupling_blocks_conv = eval(m['n_coupling_blocks_conv'])
    dropouts = eval(m['dropout_conv'])
    dropouts_fc = float(m['dropout_fc'])
    groups = int(m['n_groups'])
    clamp = float(m['clamp'])
    if groups != 1:
        raise NotImplementedError('n_groups != 1 is not used by any shipped IB-INN config and has no kernel')
    ndim_input = classifier.dims

    batchnorm_args = {'track_running_stats': True, 'momentum': 0.999, 'eps': 1e-4}
    coupling_args = {'subnet_constructor': None, 'clamp': clamp, 'act_norm': float(m['act_norm']),
                     'gin_block': False, 'permute_soft': True}

    def basic_residual_block(width, groups, dropout, relu_first, cin, cout):
        net = ConvSubnet(cin, cout, width * groups, dropout, relu_first, False, batchnorm_args)
        net.apply(_weights_init)
        return net

    def strided_residual_block(width, groups, cin, cout):
        net = ConvSubnet(cin, cout, width * groups, 0., False, True, batchnorm_args)
        net.apply(_weights_init)
        return net

    def fc_constr(c_in, c_out):
        net = FCSubnet(c_in, c_out, fc_width, dropouts_fc)
        net.apply(_weights_init)
        return net

    nodes = [Ff.InputNode(*ndim_input, name='input')]
    if classifier.dataset == 'MNIST':
        nodes.append(Ff.Node(nodes[-1].out0, Fm.Reshape, {'target_dim': (1, *classifier.dims)}))
        nodes.append(Ff.Node(nodes[-1].out0, Fm.HaarDownsampling, {'rebalance': 1.}))

    for i, (width, n_blocks) in enumerate(zip(conv_widths, n_coupling_blocks_conv)):
        if classifier.dataset == 'MNIST' and i == 0:
            continue
        drop = dropouts[i]
        conv_constr = partial(basic_residual_block, width, groups, drop, True)
        conv_strided = partial(strided_residual_block, width * 2, groups)
        conv_lowres = partial(basic_residual_block, width * 2, groups, drop, False)
        conv_first = partial(basic_residual_block, width, groups, drop, False) if i == 0 else conv_constr

        nodes.append(Ff.Node(nodes[-1], AIO_Block, dict(coupling_args, subnet_constructor=conv_first), name=f'CONV_{i}_0'))
        for k in range(1, n_blocks):
            nodes.append(Ff.Node(nodes[-1], AIO_Block, dict(coupling_args, subnet_constructor=conv_constr), name=f'CONV_{i}_{k}'))
        if i < len(conv_widths) - 1:
            nodes.append(Ff.Node(nodes[-1], DownsampleCouplingBlock,
                                 {'subnet_constructor_low_res': conv_lowres, 'subnet_constructor_strided': conv_strided,
                                  'clamp': clamp}, name=f'DOWN_{i}'))

    if use_dct:
        nodes.append(Ff.Node(nodes[-1].out0, DCTPooling2d, {'rebalance': 0.5}, name='DCT'))
    else:
        nodes.append(Ff.Node(nodes[-1].out0, Fm.Flatten, {}, name='Flatten'))

    for k in range(n_coupling_blocks_fc):
        nodes.append(Ff.Node(nodes[-1], Fm.PermuteRandom, {'seed': k}, name=f'PERM_FC_{k}'))
        nodes.append(Ff.Node(nodes[-1].out0, Fm.GLOWCouplingBlock, {'subnet_constructor': fc_constr, 'clamp': 2.0}, name=f'FC_{k}'))

    nodes.append(Ff.OutputNode(nodes[-1], name='output'))
    return Ff.ReversibleGraphNet(nodes, verbose=verbose)
