"""Build libibinn_b200.so (sm_100a) in-tree with nvcc.  No torch headers are involved: the
library is a plain C ABI (include/ibinn_b200.h) loaded with ctypes."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ['runtime.cu', 'conv.cu', 'conv_tc.cu', 'conv1x1_bwd.cu', 'block_tc.cu', 'wgrad.cu', 'wgrad_tc.cu', 'coupling.cu', 'layout.cu', 'augment.cu', 'linear.cu', 'head.cu', 'head_tc.cu', 'optim.cu']
OUT = os.path.join(HERE, 'libibinn_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, s) for s in SOURCES] + [os.path.join(HERE, 'common.cuh'),
                                                        os.path.join(HERE, '..', '..', 'include', 'ibinn_b200.h')]
    deps += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith('.cuh')]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, s.replace('.cu', '.o'))
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(HERE, s), '-o', o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    ok = True
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            ok = False
            sys.stderr.write('nvcc failed on %s:\n%s\n' % (s, out))
        elif verbose:
            print(out)
    if not ok:
        raise RuntimeError('nvcc build of libibinn_b200.so failed')
    subprocess.check_call([NVCC, '-shared', '-o', OUT] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static', '-ldl', '-lrt', '-lpthread'])
    return OUT


def build_variant(name, defines, sources=('block_tc.cu',)):
    """libibinn_b200_<name>.so: the shipped objects with `sources` recompiled under extra -D flags.  Measurement builds only
    (scripts/blk_timeline.py: name 'dbg', -DIBINN_BLK_DBG; scripts/stage_rate.py with IBINN_B200_LIB: -DIBINN_EXP_NOMMA,
    -DIBINN_EXP_SKIPLO); never loaded unless IBINN_B200_LIB points at it."""
    build()
    out = os.path.join(HERE, 'libibinn_b200_%s.so' % name)
    objs = []
    for s in SOURCES:
        o = os.path.join(HERE, s.replace('.cu', '.o'))
        if s in sources:
            o = os.path.join(HERE, s.replace('.cu', '.%s.o' % name))
            subprocess.check_call([NVCC] + FLAGS + ['-D' + d for d in defines] + ['-c', os.path.join(HERE, s), '-o', o])
        objs.append(o)
    subprocess.check_call([NVCC, '-shared', '-o', out] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static', '-ldl', '-lrt', '-lpthread'])
    return out


if __name__ == '__main__':
    if '--variant' in sys.argv:          # python build.py --variant dbg IBINN_BLK_DBG
        i = sys.argv.index('--variant')
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
