"""Builds libssasr.so (every CUDA kernel + the C ABI of include/ssasr.h) in-tree with nvcc for sm_100a."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libssasr.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--use_fast_math=false',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '-shared']
FLAGS.remove('--use_fast_math=false')   # precise math everywhere: greedy parity depends on it


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(HERE, '..', 'include', '*.h'))
    return any(os.path.getmtime(s) > t for s in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [NVCC] + FLAGS + ['-I', os.path.join(HERE, '..', 'include'), '-o', LIB] + sources()
    if verbose:
        print(' '.join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('nvcc failed building libssasr.so')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
