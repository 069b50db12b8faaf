"""Builds libmcaz.so (hand-written sm_100a CUDA + the C ABI of include/mcaz.h) in-tree with nvcc."""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
SO_PATH = os.path.join(PKG, 'libmcaz.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    # FP64 tree arithmetic must round like numpy: no FMA contraction
    '--fmad=false',
    '-shared', '-Xcompiler', '-fPIC',
    '-I', os.path.join(REPO, 'include'), '-I', CSRC,
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(SO_PATH):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(REPO, 'include', 'mcaz.h')]
    return max(os.path.getmtime(p) for p in deps) > os.path.getmtime(SO_PATH)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO_PATH
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + sources() + ['-o', SO_PATH]
    subprocess.check_call(cmd)
    return SO_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
