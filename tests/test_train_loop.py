"""Boundary callers: the train driver's log format, schedule, checkpoint round trip."""
import os

import numpy as np
import torch

from ibinn_b200 import config, train
from ibinn_b200.model import GenerativeClassifier

OV = {'model': {'n_coupling_blocks_conv': '[1, 1, 1]', 'fc_width': '8'},
      'data': {'batch_size': '3'}, 'training': {'n_epochs': '2', 'scheduler_milestones': '[1]'},
      'checkpoints': {'interval_log': '1'}}


def test_train_driver_and_checkpoint(dev, tmp_path):
    args = config.make_args(overrides=OV)
    args['checkpoints']['output_dir'] = str(tmp_path)
    torch.manual_seed(0)
    np.random.seed(0)
    ds = train.SyntheticDataset(args, n_batches=2, device=str(dev))
    inn = train.train(args, dataset=ds, device=str(dev), use_graph=False)
    rows = np.loadtxt(os.path.join(tmp_path, 'losses.dat'), usecols=[0] + list(range(3, 10)), skiprows=1)   # exactly as evaluation/__init__.py:124-126 parses it
    assert rows.shape[0] == 4 and np.isfinite(rows).all()
    assert inn.optimizer.param_groups[0]['lr'] == 0.07 * 0.1        # MultiStepLR milestone passed
    # checkpoint: same keys as the reference's save(), FrEIA tmp_var buffers stripped, loadable into a fresh model
    ck = torch.load(os.path.join(tmp_path, 'model.pt'), weights_only=False)
    assert set(ck) == {'inn', 'mu', 'phi', 'opt'} and not any('tmp_var' in k for k in ck['inn'])
    m2 = GenerativeClassifier(args).to(dev)
    m2.load(os.path.join(tmp_path, 'model.pt'))
    inn.eval(), m2.eval()
    x = ds.val_x
    with torch.no_grad():
        a, b = inn(x, ds.val_y, loss_mean=False), m2(x, ds.val_y, loss_mean=False)
    assert torch.equal(a['L_x_tr'], b['L_x_tr'])


def test_train_refuses_to_fall_back_to_noise(tmp_path):
    """no dataset and no reference data.py on sys.path: raise instead of silently training on N(0,1) noise"""
    import pytest
    args = config.make_args(overrides=OV)
    args['checkpoints']['output_dir'] = str(tmp_path)
    with pytest.raises(RuntimeError, match='synthetic=True'):
        train.train(args, dataset=None, device='cpu', use_graph=False)
