import ctypes
import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)')


def load_golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith('.gz'):
        with gzip.open(path, 'rt') as f:
            return json.load(f)
    if name.endswith('.npz'):
        return np.load(path)
    with open(path) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def mcaz_lib():
    """libmcaz.so, built in-tree if needed (nvcc cross-compiles without a GPU)."""
    from minitchess_alphazero_b200 import build, _lib
    build.build()
    return _lib.lib()


@pytest.fixture(scope='session')
def host_rules():
    """Host build of csrc/minitchess.cuh (test harness; see tests/host_harness/rules_host.cpp)."""
    src = os.path.join(REPO, 'tests', 'host_harness', 'rules_host.cpp')
    out = os.path.join(REPO, 'tests', 'host_harness', '_build', 'librules_host.so')
    hdr = os.path.join(REPO, 'minitchess_alphazero_b200', 'csrc', 'minitchess.cuh')
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-ffp-contract=off',
                               '-I', os.path.join(REPO, 'include'),
                               '-I', os.path.join(REPO, 'minitchess_alphazero_b200', 'csrc'), src, '-o', out])
    return ctypes.CDLL(out)


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)
