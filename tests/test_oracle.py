"""The oracle (oracle/ibinn_oracle.py) pinned against the golden fixtures that oracle/make_golden.py
recorded from the reference's own unmodified files (run through oracle/ref_shims in the build
container).  CPU only."""
import glob
import os

import pytest
import torch

from util import O, ROOT, rel

from ibinn_b200 import config

FIXTURES = sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', '*.pt')))


def _load(path):
    fix = torch.load(path, weights_only=False)
    args = config.make_args(overrides=fix['overrides'])
    cfg = O.cfg_from_args(args)
    spec = O.graph_spec(cfg)
    P = {k: v.double().clone() for k, v in fix['state'].items()}
    P['mu'], P['phi'] = fix['mu'].double().clone(), fix['phi'].double().clone()
    return fix, spec, P


def test_fixtures_exist():
    assert len(FIXTURES) == 4


@pytest.mark.parametrize('path', FIXTURES, ids=[os.path.basename(p)[:-3] for p in FIXTURES])
@pytest.mark.parametrize('mode', ['train', 'eval'])
def test_oracle_matches_reference(path, mode):
    fix, spec, P = _load(path)
    ref = fix['ref64_' + mode]
    train = mode == 'train'
    keys = O.trainable_keys(P)
    for k in keys:
        P[k].requires_grad_()
    out, z, jac, st = O.classifier_forward(spec, P, fix['x'].double(), fix['y'].double(), loss_mean=False, train=train)
    assert rel(z, ref['z']) < 1e-6
    assert rel(jac, ref['jac']) < 1e-6
    for k in ('L_x_tr', 'logits_tr', 'L_cNLL_tr', 'L_y_tr', 'acc_tr'):
        assert rel(out[k], ref[k]) < 1e-6, k
    if train:
        bx, by = O.beta_weights(fix['beta'])
        (bx * out['L_x_tr'].mean() - by * out['L_y_tr'].mean()).backward()
        for k, g in ref['grads'].items():
            assert rel(P[k].grad, g) < 1e-5, k
        for k, v in ref['running'].items():
            assert rel(st.new_stats[k], v) < 1e-6, k
        for idx, rec in ref['nodes'].items():
            pass   # per-node records are exercised through the CUDA path in test_golden.py
    else:
        with torch.no_grad():
            xr, _, _ = O.inn_reverse(spec, {k: v.detach() for k, v in P.items()}, ref['z'].double())
        assert rel(xr, ref["x_rev"]) < 1e-4      # z is stored rounded to fp32; the inverse amplifies that ~100x


def test_oracle_fp32_distance_is_recorded():
    """The fixtures also hold the reference's own fp32 results: that gap is the yardstick for the CUDA path."""
    for path in FIXTURES:
        fix = torch.load(path, weights_only=False)
        assert rel(fix['ref32_train']['z'], fix['ref64_train']['z']) < 1e-4


LIVE = r"""
import os, sys
import numpy as np, torch
ROOT = sys.argv[1]
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'oracle', 'ref_shims')]
import ibinn_oracle as O
import load_reference as lr
ref = lr.load()
ov = {'model': {'n_coupling_blocks_conv': '[1, 2, 1]', 'fc_width': '8', 'dropout_conv': '[0., 0., 0.]', 'dropout_fc': '0.'}}
args = lr.make_args(ov)
torch.manual_seed(3); np.random.seed(3)
m = ref.GenerativeClassifier(args).double()
x = torch.randn(3, 3, 32, 32, dtype=torch.float64)
y = torch.softmax(torch.randn(3, 10, dtype=torch.float64), 1)
m.train()
out = m(x, y, loss_mean=False)
P = {k: v.double().clone() for k, v in m.inn.state_dict().items() if 'tmp_var' not in k}
P['mu'], P['phi'] = m.mu.detach().clone(), m.phi.detach().clone()
o, z, jac, _ = O.classifier_forward(O.graph_spec(O.cfg_from_args(args)), P, x, y, loss_mean=False, train=True)
for k in out:
    e = ((o[k] - out[k]).abs().max() / out[k].abs().max().clamp_min(1e-30)).item()
    assert e < 1e-7, (k, e)
print('LIVE-OK')
"""


@pytest.mark.skipif(not os.path.isfile('/root/reference/model.py'), reason='reference checkout not present (GPU box)')
def test_oracle_against_live_reference():
    """Build container only: import the reference's unmodified files (in a subprocess, the shims patch
    torch.cuda) and compare a fresh config with the oracle."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, '-c', LIVE, ROOT], capture_output=True, text=True, timeout=600)
    assert 'LIVE-OK' in r.stdout, r.stdout + r.stderr
