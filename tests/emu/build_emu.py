"""Compile the CUDA-core kernels of ibinn_b200/csrc with g++ against the SIMT simulator in
cuda_emu.h -> tests/emu/libibinn_emu.so.  TEST INFRASTRUCTURE ONLY (see cuda_emu.h)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, '..', '..', 'ibinn_b200', 'csrc')
SOURCES = ['runtime.cu', 'conv.cu', 'conv_tc.cu', 'conv1x1_bwd.cu', 'block_tc.cu', 'wgrad.cu', 'wgrad_tc.cu', 'coupling.cu', 'layout.cu', 'augment.cu', 'linear.cu', 'head.cu', 'head_tc.cu', 'optim.cu']
OUT = os.path.join(HERE, 'libibinn_emu.so')


def build(force=False):
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, 'common.cuh'), os.path.join(CSRC, 'conv_internal.cuh'), os.path.join(HERE, 'cuda_emu.h'),
                                                        os.path.join(HERE, 'cuda_emu.cpp'),
                                                        os.path.join(CSRC, '..', '..', 'include', 'ibinn_b200.h')]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    objs = []
    procs = []
    common = ['g++', '-O2', '-g', '-std=c++17', '-fPIC', '-DIBINN_EMU=1', '-include', os.path.join(HERE, 'cuda_emu.h'),
              '-Wno-unused-variable', '-Wno-unknown-pragmas', '-Wno-unused-function']
    for s in SOURCES:
        o = os.path.join(HERE, 'emu_' + s.replace('.cu', '.o'))
        procs.append((s, subprocess.Popen(common + ['-x', 'c++', '-c', os.path.join(CSRC, s), '-o', o],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    o = os.path.join(HERE, 'emu_cuda_emu.o')
    procs.append(('cuda_emu.cpp', subprocess.Popen(common + ['-c', os.path.join(HERE, 'cuda_emu.cpp'), '-o', o],
                                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs.append(o)
    ok = True
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            ok = False
            sys.stderr.write('g++ failed on %s:\n%s\n' % (s, out))
    if not ok:
        raise RuntimeError('emulator build failed')
    subprocess.check_call(['g++', '-shared', '-o', OUT] + objs)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
