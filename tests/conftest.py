import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


def _emu_library():
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
    import build_emu
    return build_emu.build()


@pytest.fixture(params=['emu', pytest.param('cuda', marks=pytest.mark.gpu)])
def dev(request):
    """Device under test.  'emu' = the kernel sources compiled for the host-side SIMT simulator
    (tests/emu, logic check only, CPU tensors); 'cuda' = the real sm_100a library on a B200."""
    from ibinn_b200 import _lib
    if request.param == 'emu':
        _lib._use_library_for_tests(_emu_library(), True)
        yield torch.device('cpu')
        _lib._use_library_for_tests(None, False)
    else:
        if not torch.cuda.is_available():
            pytest.skip('no GPU')
        _lib._use_library_for_tests(None, False)
        yield torch.device('cuda:0')


@pytest.fixture
def cuda_dev():
    from ibinn_b200 import _lib
    if os.environ.get('IBINN_TEST_EMU_AS_CUDA') == '1' and not torch.cuda.is_available():
        # dry run of the GPU-only tests' host logic on the simulator (developer aid; shrink sizes with IBINN_TEST_BENCH_B)
        _lib._use_library_for_tests(_emu_library(), True)
        return torch.device('cpu')
    if not torch.cuda.is_available():
        pytest.skip('no GPU')
    _lib._use_library_for_tests(None, False)
    return torch.device('cuda:0')
