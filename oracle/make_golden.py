"""Generate tests/golden/*.pt by running the reference's UNMODIFIED files (through
oracle/ref_shims) in the build container.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py            # needs /root/reference; rewrites tests/golden/

The reference ships no known-answer vectors, so these fixtures are the pin:
they record what /root/reference/model.py::GenerativeClassifier computes, in fp64 (the truth
the CUDA path and the oracle are compared with) and in fp32 (to show how far the reference's
own fp32 arithmetic sits from that truth on the same inputs).  Dropout is disabled through the
.ini keys (`dropout_conv`, `dropout_fc`) because torch's Philox stream is not part of parity.

Per case: the merged config (as the ini overrides), inputs, the full fp32 state, and for
train-mode BN and eval-mode BN: z, log-det, the per-sample loss dict, (fp64, train mode) gradients
of beta_x*L_x - beta_y*L_y w.r.t. every trainable tensor, the BN running statistics after the
train-mode forward, every node's (input, output, log-det) from the fp64 chain, and the
reverse pass x = inn(z, rev=True).
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'ref_shims'))
sys.path.insert(0, os.path.dirname(HERE))
import load_reference as lr            # noqa: E402

NO_DROP = {'dropout_conv': '[0., 0., 0.]', 'dropout_fc': '0.'}

CASES = {
    # small, well-conditioned instances of the three shipped architectures
    'cifar10_222': dict(overrides={'model': dict(NO_DROP, n_coupling_blocks_conv='[2, 2, 2]', fc_width='8')},
                        B=6, beta=1.0, smooth=0.02),
    'mnist_22': dict(overrides={'model': dict(NO_DROP, n_coupling_blocks_conv='[8, 2, 2]', fc_width='16',
                                              act_norm='0.80'),
                                'data': {'dataset': 'MNIST', 'label_smoothing': '0.01'}},
                     B=5, beta=0.7010, smooth=0.01),
    'cifar100_111': dict(overrides={'model': dict(NO_DROP, n_coupling_blocks_conv='[1, 1, 1]', fc_width='4'),
                                    'data': {'dataset': 'CIFAR100'}},
                         B=4, beta=4.6261, smooth=0.02),
    'cifar10_pad5_noDCT_wi05': dict(overrides={'model': dict(NO_DROP, n_coupling_blocks_conv='[1, 1, 1]', fc_width='4',
                                                             dct_pooling='False', weight_init='0.5'),
                                               'data': {'pad_noise_channels': '5'}},
                                    B=3, beta=1.0, smooth=0.02),
}


def f32(t):
    return t.detach().to(torch.float32).clone() if t.is_floating_point() else t.detach().clone()


def run(model, x, y, beta, train):
    model.train(train)
    model.zero_grad()
    out = model(x, y, loss_mean=False)
    bx, by = 2. / (1 + beta), 2. * beta / (1 + beta)           # train.py:88-92
    loss = bx * out['L_x_tr'].mean() - by * out['L_y_tr'].mean()
    res = {k: f32(v) for k, v in out.items()}
    res['loss'] = f32(loss)
    return res, loss


def main():
    ref = lr.load()
    outdir = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
    os.makedirs(outdir, exist_ok=True)
    for name, case in CASES.items():
        args = lr.make_args(case['overrides'])
        torch.manual_seed(0)
        np.random.seed(0)
        m32 = ref.GenerativeClassifier(args)
        # move the GMM off its symmetric init so that every gradient path is exercised
        g = torch.Generator().manual_seed(7)
        m32.mu.data += 0.05 * torch.randn(m32.mu.shape, generator=g)
        m32.phi.data += 0.3 * torch.randn(m32.phi.shape, generator=g)
        for p_name, p in m32.inn.named_parameters():
            if p_name.endswith('act_offset') or (p_name.endswith('.bias') and p.requires_grad):
                p.data += 0.1 * torch.randn(p.shape, generator=g)
        C, B = m32.n_classes, case['B']
        x = torch.randn(B, *m32.dims, generator=g)
        lab = torch.randint(0, C, (B,), generator=g)
        y = torch.zeros(B, C).scatter_(1, lab.view(-1, 1), 1.) * (1 - case['smooth']) + case['smooth'] / C

        state = {k: f32(v) for k, v in m32.inn.state_dict().items() if 'tmp_var' not in k}
        fix = {'overrides': case['overrides'], 'beta': case['beta'], 'x': x, 'y': y,
               'state': state, 'mu': f32(m32.mu), 'phi': f32(m32.phi)}

        for dt, tag in ((torch.float64, 'ref64'), (torch.float32, 'ref32')):
            model = copy.deepcopy(m32).to(dt)
            xd, yd = x.to(dt), y.to(dt)
            for train in (True, False):
                mode = 'train' if train else 'eval'
                mm = copy.deepcopy(model)
                res, loss = run(mm, xd, yd, case['beta'], train)
                loss.backward()
                # z / jac / per-node records from a second, stats-neutral walk of the same chain
                mm2 = copy.deepcopy(model)
                mm2.train(train)
                with torch.no_grad():
                    z = mm2.inn(xd)
                    jac = mm2.inn.log_jacobian(run_forward=False)
                    nodes = {}
                    if dt == torch.float64 and train:
                        for k, h in mm2.inn._node_inputs.items():
                            probe = copy.deepcopy(model).train(train).inn.module_list[k]
                            o = probe([h])[0]
                            j = probe.jacobian([h])
                            j = f32(j) if torch.is_tensor(j) else torch.full((B,), float(j))
                            nodes[k] = {'out': f32(o), 'jac': j}     # input of node k = output of node k-1 (x for the first)
                res['z'], res['jac'] = f32(z), f32(jac)
                if train and dt == torch.float64:
                    grads = {n_: f32(p.grad) for n_, p in mm.inn.named_parameters() if p.grad is not None}
                    grads['mu'], grads['phi'] = f32(mm.mu.grad), f32(mm.phi.grad)
                    res['grads'] = grads
                if train:
                    res['running'] = {k: f32(v) for k, v in mm.inn.state_dict().items() if 'running_' in k}
                if dt == torch.float64:
                    if train:
                        res['nodes'] = nodes
                    with torch.no_grad():
                        me = copy.deepcopy(model).eval()
                        if not train:
                            res['x_rev'] = f32(me.inn(z, rev=True))
                fix['%s_%s' % (tag, mode)] = res
        path = os.path.join(outdir, name + '.pt')
        torch.save(fix, path)
        e = (fix['ref32_train']['z'] - fix['ref64_train']['z']).abs().max() / fix['ref64_train']['z'].abs().max()
        print('%-22s %6.2f MB   ref fp32-vs-fp64 z rel err (train BN) %.2e' % (
            name, os.path.getsize(path) / 2 ** 20, e))


if __name__ == '__main__':
    main()
