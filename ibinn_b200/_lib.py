"""ctypes binding of the C ABI in include/ibinn_b200.h (csrc/libibinn_b200.so, sm_100a only).

There is no CPU path and no fallback: if the shared library is missing, or the current device is
not a B200, every op raises.  (tests/emu builds a host-side SIMT *simulator* of the same kernel
sources; only tests may point this module at it through `_use_library_for_tests`.)
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('IBINN_B200_LIB') or os.path.join(_HERE, 'csrc', 'libibinn_b200.so')      # override: A/B runs of two builds

_lib = None
_allow_host_pointers = False      # set only by tests/emu (simulator build of the kernels)

ERRORS = {-1: 'IBINN_EINVAL (bad shape / unsupported size)', -2: 'IBINN_EALIGN (misaligned pointer or stride)',
          -3: 'IBINN_EARCH (device is not sm_100 / B200; there is no fallback path)', -4: 'IBINN_ENULL (required pointer is NULL)'}


class BnRef(Structure):
    _fields_ = [('mean', c_void_p), ('scale', c_void_p), ('gamma', c_void_p), ('beta', c_void_p),
                ('eps', c_float), ('scale_is_variance', c_int32)]


class ConvArgs(Structure):
    _fields_ = [('x', c_void_p), ('x_bs', c_int64), ('x2', c_void_p), ('x2_bs', c_int64), ('w', c_void_p), ('bias', c_void_p),
                ('addend', c_void_p), ('addend_bs', c_int64), ('y', c_void_p), ('y_bs', c_int64),
                ('B', c_int32), ('Cin', c_int32), ('Cout', c_int32), ('Hs', c_int32), ('Ws', c_int32), ('Hout', c_int32),
                ('Wout', c_int32), ('ksize', c_int32), ('stride', c_int32), ('upsample2', c_int32), ('Hin', c_int32),
                ('Win', c_int32), ('w_transposed', c_int32), ('pre', c_int32), ('epi', c_int32),
                ('pre_bn', BnRef), ('pre_mask', c_void_p), ('pre_sums', c_void_p), ('pre_inv_count', c_float),
                ('save_mean', c_void_p), ('save_invstd', c_void_p), ('running_mean', c_void_p), ('running_var', c_void_p),
                ('num_batches_tracked', c_void_p), ('momentum', c_float), ('bn_eps', c_float),
                ('m_h', c_void_p), ('m_bs', c_int64), ('m_bn', BnRef), ('m_drop', c_void_p), ('sums_out', c_void_p),
                ('m_dgamma', c_void_p), ('m_dbeta', c_void_p), ('partials', c_void_p), ('counter', c_void_p), ('wprep', c_void_p), ('wprep_ready', c_int32)]


class WprepDesc(Structure):
    _fields_ = [('w', c_void_p), ('out', c_void_p), ('Cin', c_int32), ('Cout', c_int32), ('ksize', c_int32), ('transposed', c_int32)]


class WreduceDesc(Structure):
    _fields_ = [('partials', c_void_p), ('dw', c_void_p), ('dbias', c_void_p), ('ncta', c_int32), ('rec_floats', c_int32),
                ('Cout', c_int32), ('Cin', c_int32), ('kk', c_int32)]


class AioReduceDesc(Structure):
    _fields_ = [('partials', c_void_p), ('g_act_offset', c_void_p), ('g_act_norm', c_void_p), ('nrec', c_int32), ('C', c_int32)]


class WgradArgs(Structure):
    _fields_ = [('g', c_void_p), ('g_bs', c_int64), ('gh', c_void_p), ('gh_bs', c_int64), ('g_pre', c_int32),
                ('g_bn', BnRef), ('g_sums', c_void_p), ('g_inv_count', c_float),
                ('x', c_void_p), ('x_bs', c_int64), ('x_pre', c_int32), ('x_bn', BnRef), ('x_mask', c_void_p),
                ('dw', c_void_p), ('dbias', c_void_p),
                ('B', c_int32), ('Cin', c_int32), ('Cout', c_int32), ('Hin', c_int32), ('Win', c_int32), ('Hout', c_int32),
                ('Wout', c_int32), ('ksize', c_int32), ('stride', c_int32), ('workspace', c_void_p), ('defer_reduce', c_int32)]


class Conv1x1BwdArgs(Structure):
    _fields_ = [('g', c_void_p), ('h', c_void_p), ('h_bs', c_int64), ('bn', BnRef), ('drop', c_void_p), ('w', c_void_p),
                ('g_in', c_void_p), ('sums_out', c_void_p), ('dgamma', c_void_p), ('dbeta', c_void_p),
                ('partials', c_void_p), ('counter', c_void_p), ('wrecords', c_void_p),
                ('B', c_int32), ('Cin', c_int32), ('Cout', c_int32), ('H', c_int32), ('W', c_int32)]


class AioBlockDesc(Structure):
    _fields_ = [('wprep1', c_void_p), ('wprep2', c_void_p), ('wprep3', c_void_p), ('wprep_perm', c_void_p), ('bias3', c_void_p),
                ('bn1', BnRef), ('bn2', BnRef), ('act_norm', c_void_p), ('act_offset', c_void_p), ('drop_mask', c_void_p),
                ('out', c_void_p), ('h1', c_void_p), ('h2', c_void_p), ('a', c_void_p), ('save', c_void_p),
                ('running_mean1', c_void_p), ('running_var1', c_void_p), ('nbt1', c_void_p),
                ('running_mean2', c_void_p), ('running_var2', c_void_p), ('nbt2', c_void_p),
                ('relu_first', c_int32), ('reserved', c_int32)]


class AioStageArgs(Structure):
    _fields_ = [('x', c_void_p), ('jac', c_void_p), ('blocks', c_void_p),
                ('nblk', c_int32), ('B', c_int32), ('C', c_int32), ('H', c_int32), ('W', c_int32), ('width', c_int32), ('train', c_int32),
                ('clamp', c_float), ('alpha', c_float), ('momentum', c_float),
                ('stat_records', c_void_p), ('barrier', c_void_p)]


P, I, F, L = c_void_p, c_int, c_float, c_int64
_SIGNATURES = {
    'ibinn_sizeof_conv1x1_bwd_args': ([], c_size_t),
    'ibinn_conv1x1_bwd_geometry': ([POINTER(Conv1x1BwdArgs), POINTER(c_int32), POINTER(c_int32)], c_int),
    'ibinn_conv1x1_backward': ([POINTER(Conv1x1BwdArgs), P], c_int),
    'ibinn_sizeof_aio_block_desc': ([], c_size_t),
    'ibinn_sizeof_aio_stage_args': ([], c_size_t),
    'ibinn_aio_stage_supported': ([POINTER(AioStageArgs)], c_int),
    'ibinn_aio_stage_workspace_floats': ([POINTER(AioStageArgs)], c_int64),
    'ibinn_aio_stage_forward': ([POINTER(AioStageArgs), P], c_int),
    'ibinn_abi_version': ([], c_int),
    'ibinn_launch_count': ([], c_int64),
    'ibinn_last_error': ([], ctypes.c_char_p),
    'ibinn_check_device': ([], c_int),
    'ibinn_sizeof_conv_args': ([], c_size_t),
    'ibinn_conv_partials_floats': ([POINTER(ConvArgs)], c_int64),
    'ibinn_conv_wprep_floats': ([POINTER(ConvArgs)], c_int64),
    'ibinn_conv_forward': ([POINTER(ConvArgs), P], c_int),
    'ibinn_wprep_many': ([P, I, P], c_int),
    'ibinn_sizeof_wgrad_args': ([], c_size_t),
    'ibinn_conv_wgrad_workspace_floats': ([POINTER(WgradArgs)], c_int64),
    'ibinn_conv_wgrad': ([POINTER(WgradArgs), P], c_int),
    'ibinn_conv_wgrad_records': ([POINTER(WgradArgs), POINTER(c_int32), POINTER(c_int32)], c_int),
    'ibinn_wgrad_reduce_many': ([P, I, P], c_int),
    'ibinn_aio_coupling_forward': ([P, P, P, P, P, P, P, I, I, I, I, F, F, P], c_int),
    'ibinn_aio_coupling_backward': ([P] * 14 + [I, I, I, I, F, F, I, P], c_int),
    'ibinn_aio_backward_records': ([I, I, I, I], c_int),
    'ibinn_aio_param_reduce_many': ([P, I, P], c_int),
    'ibinn_aio_unpermute': ([P, P, P, P, P, I, I, I, I, P], c_int),
    'ibinn_aio_affine_inverse': ([P, P, P, P, I, I, I, I, F, F, P], c_int),
    'ibinn_s2d_coupling_forward': ([P, L, P, P, L, P, I, I, I, I, I, F, F, P], c_int),
    'ibinn_s2d_coupling_inverse': ([P, L, P, P, L, P, I, I, I, I, I, F, F, P], c_int),
    'ibinn_s2d_coupling_backward': ([P, L, P, P, L, P, P, L, P, I, I, I, I, F, F, P], c_int),
    'ibinn_glow_affine_forward': ([P, L, P, P, L, P, I, I, I, F, P], c_int),
    'ibinn_glow_affine_inverse': ([P, L, P, P, L, P, I, I, I, F, P], c_int),
    'ibinn_glow_affine_backward': ([P, L, P, P, L, P, P, L, P, I, I, F, P], c_int),
    'ibinn_dct2d_forward': ([P, P, P, I, I, I, F, P], c_int),
    'ibinn_dct2d_inverse': ([P, P, P, I, I, I, F, I, P], c_int),
    'ibinn_haar_forward': ([P, P, I, I, I, I, F, P], c_int),
    'ibinn_haar_inverse': ([P, P, I, I, I, I, F, P], c_int),
    'ibinn_permute_features': ([P, P, P, I, I, P], c_int),
    'ibinn_linear_workspace_floats': ([I, I, I], c_int64),
    'ibinn_linear_tickets': ([I, I, I], c_int),
    'ibinn_linear_forward': ([P, L, P, I, P, P, P, I, I, I, P, P, P], c_int),
    'ibinn_linear_dgrad': ([P, P, P, I, P, P, L, I, I, I, I, P, P, P], c_int),
    'ibinn_linear_wgrad': ([P, P, P, I, P, L, P, I, P, P, I, I, I, P, P, P], c_int),
    'ibinn_gmm_head_partials_floats': ([I], c_int64),
    'ibinn_gmm_head_forward': ([P] * 14 + [I, I, I, P], c_int),
    'ibinn_gmm_head_gemm_supported': ([I, I, I], c_int),
    'ibinn_gmm_head_gemm_workspace_floats': ([I, I, I], c_int64),
    'ibinn_gmm_head_forward_gemm': ([P] * 15 + [I, I, I, I, P], c_int),
    'ibinn_gmm_head_backward': ([P] * 6 + [P, I, P, I, P, I, P, I, F, F, P, P, P, P, P, I, P, I, I, I, I, P], c_int),
    'ibinn_augment_batch': ([P, P, P, P, P, P, I, I, I, I, I, POINTER(c_float), POINTER(c_float), F, F, I, P], c_int),
    'ibinn_grad_norm_partials_doubles': ([], c_int64),
    'ibinn_grad_norm_clip': ([P, L, P, I, F, P, P, P, P], c_int),
    'ibinn_sgd_step': ([P, P, P, L, P, P, F, P], c_int),
}
EXPORTS = tuple(_SIGNATURES)


def _declare(lib):
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the library does not export it
        fn.argtypes = argtypes
        fn.restype = restype
    if lib.ibinn_sizeof_conv_args() != ctypes.sizeof(ConvArgs) or lib.ibinn_sizeof_wgrad_args() != ctypes.sizeof(WgradArgs) \
            or lib.ibinn_sizeof_conv1x1_bwd_args() != ctypes.sizeof(Conv1x1BwdArgs) or lib.ibinn_sizeof_aio_block_desc() != ctypes.sizeof(AioBlockDesc) or lib.ibinn_sizeof_aio_stage_args() != ctypes.sizeof(AioStageArgs):
        raise RuntimeError('ibinn_b200: ctypes structure layout does not match include/ibinn_b200.h')
    return lib


def load(path=None):
    """dlopen the shared library and declare every symbol of include/ibinn_b200.h (no compute call)."""
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError('ibinn_b200: %s not found.  Build it with `python -c "import __graft_entry__ as g; g.build()"` '
                           '(nvcc, sm_100a).  There is no CPU or PyTorch fallback.' % path)
    return _declare(ctypes.CDLL(path))


class _Profiler:
    """Proxy over the library that brackets every entry point with CUDA events on the current stream
    (bench.py's per-kernel-class device times; eager mode only, never active under graph capture)."""

    def __init__(self, lib):
        self._lib, self.records = lib, []

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if fn.restype is not c_int or name == 'ibinn_check_device' or name.endswith('_records'):      # queries launch nothing
            return fn

        def timed(*args):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = fn(*args)
            b.record()
            self.records.append((name + ('[%s]' % _tag if _tag else ''), a, b))
            return rc
        return timed

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.records:
            out.setdefault(name, []).append(a.elapsed_time(b))        # one duration per launch, in launch order
        return out


_profiler = None
_tag = ''


def set_tag(t):
    global _tag
    _tag = t


def profile(enable):
    """bench.py: start / stop per-entry-point device timing.  Returns the finished profiler on stop."""
    global _profiler
    old, _profiler = _profiler, (_Profiler(lib()) if enable else None)
    return old


def lib():
    global _lib
    if _lib is None:
        _lib = load()
    return _profiler if _profiler is not None else _lib


def _use_library_for_tests(path, allow_host_pointers):
    """tests/emu only: route calls to the simulator build of the kernels."""
    global _lib, _allow_host_pointers
    _lib = load(path) if path else None
    _allow_host_pointers = bool(allow_host_pointers)


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError('ibinn_b200.%s rejected the call: %s' % (what, ERRORS.get(rc, rc)))
    msg = ''
    try:
        msg = (_lib.ibinn_last_error() or b'').decode()
    except Exception:
        pass
    raise RuntimeError('ibinn_b200.%s: CUDA error %d (%s)' % (what, rc, msg))


def ptr(t, dtype=torch.float32, strided=False):
    """device pointer of a tensor (None -> NULL).  Dense row-major unless the caller handles strides."""
    if t is None:
        return None
    if not strided and not t.is_contiguous():
        raise ValueError('ibinn_b200: tensor of shape %s with strides %s must be contiguous' % (tuple(t.shape), t.stride()))
    if t.dtype != dtype:
        raise TypeError('ibinn_b200: expected %s, got %s' % (dtype, t.dtype))
    if not t.is_cuda and not _allow_host_pointers:
        raise RuntimeError('ibinn_b200 runs on a B200 only: got a %s tensor (there is no CPU path)' % t.device)
    return t.data_ptr()


def stream(t):
    if t.is_cuda:
        return torch.cuda.current_stream(t.device).cuda_stream
    return None
