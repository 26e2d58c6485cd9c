"""Import the reference's UNMODIFIED python files from /root/reference in this container.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py and by the `not gpu` tests that
cross-check the oracle when /root/reference is present.  Nothing under ibinn_b200/ may
import this.  The four shims are the ones listed in SURVEY.md Appendix F:
  1. torch.rfft / torch.irfft aliases (removed in torch 1.8; dct_transform.py:20,57)
  2. FrEIA stand-in package (requirements.txt:2 is not vendored)
  3. matplotlib stub (data.py:6)
  4. CPU-only box: neutralise .cuda() (dct_transform.py:72-73, model.py:134,213,233)
"""
import configparser
import importlib
import os
import sys

import torch

REFERENCE_DIR = os.environ.get('IBINN_REFERENCE_DIR', '/root/reference')
_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, 'model.py'))


def _install_shims(force_cpu=False):
    if not hasattr(torch, 'rfft'):
        torch.rfft = lambda v, nd, onesided=False: torch.view_as_real(torch.fft.fft(v, dim=1))
        torch.irfft = lambda V, nd, onesided=False: torch.fft.ifft(
            torch.view_as_complex(V.contiguous()), dim=1).real
    if force_cpu or not torch.cuda.is_available():
        torch.cuda.is_available = lambda: True
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    for p in (REFERENCE_DIR, _HERE):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_DIR)
    sys.path.insert(0, _HERE)      # FrEIA / matplotlib stand-ins shadow nothing real


def load(force_cpu=False):
    """Returns the reference's `model` module (GenerativeClassifier lives there).  force_cpu: neutralise `.cuda()` even on
    a box that has a GPU (the CPU arm of bench.py)."""
    if not available():
        raise FileNotFoundError(REFERENCE_DIR)
    _install_shims(force_cpu)
    for name in ('model', 'inn_architecture', 'all_in_one_block',
                 'downsampling_coupling_block', 'dct_transform'):
        mod = sys.modules.get(name)
        if mod is not None and not (getattr(mod, '__file__', '') or '').startswith(REFERENCE_DIR):
            del sys.modules[name]
    return importlib.import_module('model')


def make_args(overrides=None, ini=None):
    """default.ini overlaid by an experiment ini and then by {section: {key: value}}."""
    args = configparser.ConfigParser()
    args.read(os.path.join(REFERENCE_DIR, 'default.ini'))
    if ini:
        args.read(ini)
    for sec, kv in (overrides or {}).items():
        for k, v in kv.items():
            args[sec][k] = str(v)
    return args
