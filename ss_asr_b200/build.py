"""Builds libssasr.so (every CUDA kernel + the C ABI of include/ssasr.h) in-tree with nvcc for sm_100a.

One object per csrc/*.cu (compiled in parallel, only when the source or any header is newer than the object), then one link
step into a temporary file that is atomically renamed over libssasr.so.  The whole build runs under an exclusive file lock,
so concurrent ranks of a torchrun launch never see -- or write -- a half-linked library: the first rank builds, the others
wait and find it fresh."""
import fcntl
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(HERE, '..', 'include')
OBJ = os.path.join(HERE, '..', 'build_tmp', 'obj')
LIB = os.path.join(HERE, 'libssasr.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
# precise math everywhere (no --use_fast_math): greedy parity depends on it
CFLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
          '-Xcompiler', '-O2', '-I', INCLUDE]
ABI_VERSION = 4          # must equal SSASR_ABI_VERSION in include/ssasr.h (checked by _lib.load)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _headers():
    return glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(INCLUDE, '*.h'))


def _obj(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')


def _newer(path, deps):
    if not os.path.isfile(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(d) > t for d in deps)


def stale():
    return _newer(LIB, sources() + _headers())


def _run(cmd, verbose):
    if verbose:
        print(' '.join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('nvcc failed: ' + ' '.join(cmd[:1] + cmd[-3:]))


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, '.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not stale():          # another process built it while we were waiting for the lock
                return LIB
            hdrs = _headers()
            todo = [s for s in sources() if force or _newer(_obj(s), [s] + hdrs)]
            with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1) or 1) as ex:
                list(ex.map(lambda s: _run([NVCC] + CFLAGS + ['-c', s, '-o', _obj(s)], verbose), todo))
            tmp = LIB + '.tmp.%d' % os.getpid()
            _run([NVCC] + CFLAGS + ['-shared', '-o', tmp] + [_obj(s) for s in sources()], verbose)
            os.replace(tmp, LIB)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
