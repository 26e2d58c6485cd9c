"""The C-ABI library loads and exports every symbol include/ibinn_b200.h declares; the product path
refuses to run without the CUDA library or on CPU tensors (no fallback).  No GPU needed."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from util import ROOT

HEADER = os.path.join(ROOT, 'include', 'ibinn_b200.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ibinn_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ('ibinn_conv_forward', 'ibinn_conv_wgrad', 'ibinn_aio_coupling_forward', 'ibinn_aio_coupling_backward',
                 's2d_coupling_forward', 'glow_affine_forward', 'dct2d_forward', 'haar_forward', 'permute_features',
                 'linear_forward', 'gmm_head_forward', 'gmm_head_backward', 'grad_norm_clip', 'sgd_step'):
        assert any(must in n for n in names), must


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    path = ge.build()
    lib = ctypes.CDLL(path)
    for name in declared_functions():
        assert hasattr(lib, name), name
    from ibinn_b200 import _lib
    assert set(_lib.EXPORTS) == set(declared_functions())
    assert lib.ibinn_abi_version() == 1
    out = subprocess.run(['cuobjdump', '-lelf', path], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out


def test_no_cpu_fallback():
    from ibinn_b200 import _lib, ops
    _lib._use_library_for_tests(None, False)
    x = torch.randn(2, 4, 8, 8)
    with pytest.raises(RuntimeError, match='B200 only'):
        ops.haar_forward(x, 0.5)
    if not torch.cuda.is_available():
        # and the library itself refuses a machine without a B200
        assert _lib.lib().ibinn_check_device() != 0


def test_missing_library_fails_loudly(tmp_path):
    from ibinn_b200 import _lib
    with pytest.raises(RuntimeError, match='no CPU or PyTorch fallback'):
        _lib.load(str(tmp_path / 'nope.so'))


def test_product_does_not_import_oracle_or_simulator():
    pkg = os.path.join(ROOT, 'ibinn_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'ibinn_oracle' not in src and 'ref_shims' not in src, f
                assert 'libibinn_emu' not in src and 'build_emu' not in src, f
