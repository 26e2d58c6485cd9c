"""Full-depth end-to-end parity: the DEFAULT architecture ([8, 24, 24] blocks, fc 1024, 61 nodes) at B = 32 against what the
reference's unmodified files computed (tests/golden/full/full_depth.pt, made by oracle/make_golden_full.py).

* `weight_init = 0.5` (the conditioned full-depth config of SURVEY.md Appendix D): north_star's gates -- z and per-sample
  log-det 1e-4, losses / per-class log-likelihoods 1e-3, bit-exact arg-max -- in eval-mode BatchNorm; in train-mode BatchNorm
  the reference's OWN fp32 run sits at 2.1e-4 from its fp64 run on z, so that mode is gated relative to it.
* default init: chaotic end to end (reference fp32 vs fp64: 3.6e-2 on z); the test prints err(engine) / err(reference fp32)
  and only bounds the ratio.
The parameters are rebuilt from the seeds (bit-identical constructors) and pinned by per-tensor checksums."""
import os

import numpy as np
import pytest
import torch

from util import O, ROOT, rel

from ibinn_b200 import config
from ibinn_b200.model import GenerativeClassifier

PATH = os.path.join(ROOT, 'tests', 'golden', 'full', 'full_depth.pt')


def build(rec):
    args = config.make_args(overrides=rec['overrides'])
    torch.manual_seed(0)
    np.random.seed(0)
    m = GenerativeClassifier(args)
    sd = m.inn.state_dict()
    for k, (s1, s2) in rec['checksum'].items():
        v = sd[k].double()
        tol = 1e-6 if v.dim() == 2 else 1e-9          # DCT matrices: analytic here, FFT-derived (3e-7) in the reference
        assert abs(float(v.sum()) - s1) <= tol * max(1., abs(s1)) and abs(float((v ** 2).sum()) - s2) <= tol * max(1., s2), k
    return m, args


@pytest.mark.parametrize('mode', ['train', 'eval'])
def test_oracle_full_depth_matches_reference(mode):
    """pins the fp64 oracle on the full-depth network (1e-6 on z / log-det / losses)"""
    fix = torch.load(PATH, weights_only=False)
    rec = fix['wi05']
    m, args = build(rec)
    P = {k: v.detach().double().clone() for k, v in m.inn.state_dict().items() if 'tmp_var' not in k}
    P['mu'], P['phi'] = m.mu.detach().double(), m.phi.detach().double()
    with torch.no_grad():
        out, z, jac, _ = O.classifier_forward(O.graph_spec(O.cfg_from_args(args)), P, fix['x'].double(), fix['y'].double(),
                                              loss_mean=False, train=(mode == 'train'))
    ref = rec['ref64_' + mode]
    assert rel(z, ref['z']) < 1e-6 and rel(jac, ref['jac']) < 1e-6
    for k in ('L_x_tr', 'L_cNLL_tr', 'L_y_tr', 'logits_tr'):
        assert rel(out[k], ref[k]) < 1e-5, k


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_engine_full_depth_conditioned(cuda_dev, mode):
    fix = torch.load(PATH, weights_only=False)
    rec = fix['wi05']
    ref, r32 = rec['ref64_' + mode], rec['ref32_' + mode]
    m, _ = build(rec)
    m = m.to(cuda_dev)
    m.train(mode == 'train')
    x, y = fix['x'].to(cuda_dev), fix['y'].to(cuda_dev)
    out = m(x, y, loss_mean=False)
    z = m.inn._jacs[-1][1]
    jac = m.inn.log_jacobian(run_forward=False)
    ez, e32 = rel(z, ref['z']), rel(r32['z'], ref['z'])
    print('full depth, weight_init 0.5, %s BN: z err engine %.2e, reference fp32 %.2e, ratio %.2f; log-det err %.2e'
          % (mode, ez, e32, ez / e32, rel(jac, ref['jac'])))
    gate_z = 1e-4 if mode == 'eval' else max(1e-4, 3. * e32)        # train BN: the reference's own fp32 is at 2.1e-4
    assert ez < gate_z
    assert rel(jac, ref['jac']) < 1e-4
    for k in ('L_x_tr', 'L_cNLL_tr', 'L_y_tr', 'logits_tr'):
        # train-mode BN at full depth: the reference's own fp32 L_y is 8.4e-4 from its fp64 value, so that mode is gated relative to it
        gate = 1e-3 if mode == 'eval' else max(1e-3, 3. * rel(r32[k], ref[k]))
        assert rel(out[k], ref[k]) < gate, (k, rel(out[k], ref[k]), rel(r32[k], ref[k]))
    assert torch.equal(out['logits_tr'].argmax(1).cpu(), ref['logits_tr'].argmax(1))
    assert float(out['acc_tr']) == float(ref['acc_tr'])
    if mode == 'train':
        bx, by = 2. / (1 + fix['beta']), 2. * fix['beta'] / (1 + fix['beta'])
        (bx * out['L_x_tr'].mean() - by * out['L_y_tr'].mean()).backward()
        # mu / phi sit above every ReLU: tight.  The network gradient is gated through its global norm (what clip_grad_norm_ sees)
        assert rel(m.mu.grad, ref['g_mu']) < 1e-3 and rel(m.phi.grad, ref['g_phi']) < 1e-3
        gn = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.trainable_params if p.grad is not None)))
        assert abs(gn - ref['grad_norm']) < 2e-2 * ref['grad_norm']
        sd = m.inn.state_dict()
        for k, v in ref['running'].items():
            assert rel(sd[k], v) < max(1e-4, 2. * ez), k           # statistics of activations that carry the z error above


@pytest.mark.gpu
def test_engine_full_depth_default_init_error_ratio(cuda_dev):
    """err(engine vs fp64) / err(reference fp32 vs fp64) on z at the DEFAULT init, train-mode BN."""
    fix = torch.load(PATH, weights_only=False)
    rec = fix['default']
    ref, r32 = rec['ref64_train'], rec['ref32_train']
    m, _ = build(rec)
    m = m.to(cuda_dev)
    m.train()
    with torch.no_grad():
        z = m.inn(fix['x'].to(cuda_dev))
        jac = m.inn.log_jacobian(run_forward=False)
    ez, e32 = rel(z, ref['z']), rel(r32['z'], ref['z'])
    ej, j32 = rel(jac, ref['jac']), rel(r32['jac'], ref['jac'])
    print('full depth, default init, train BN: z err engine %.2e vs reference fp32 %.2e (ratio %.2f); log-det %.2e vs %.2e'
          % (ez, e32, ez / e32, ej, j32))
    assert ez < 10. * e32 and ej < 10. * j32 + 1e-6
