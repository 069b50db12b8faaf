"""CPU: libmcaz.so builds for sm_100a, loads, exports every symbol include/mcaz.h declares, and
refuses to compute without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from conftest import REPO


def declared_symbols():
    text = open(os.path.join(REPO, 'include', 'mcaz.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b((?:mc|az|mcaz)_[a-z_0-9]+)\s*\(', text)))


def test_exports_every_declared_symbol(mcaz_lib):
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(mcaz_lib, n)]
    assert not missing, missing


def test_host_helpers(mcaz_lib):
    from minitchess_alphazero_b200 import rules
    s = rules.state_from_fen(rules.STARTING_FEN)
    assert rules.state_to_fen(s) == rules.STARTING_FEN
    from oracle import rules_c as rc
    assert s == rc.fen_to_state(rules.STARTING_FEN)
    import ctypes
    f, t = ctypes.c_int(), ctypes.c_int()
    for code in range(554):
        for white in (0, 1):
            assert mcaz_lib.mc_code_squares(code, white, ctypes.byref(f), ctypes.byref(t)) == 0
            assert mcaz_lib.mc_squares_code(f.value, t.value, white) == code
            assert rc.code_of(f.value, t.value, white) == code
    assert mcaz_lib.mc_code_squares(554, 1, ctypes.byref(f), ctypes.byref(t)) != 0
    with pytest.raises(Exception):
        rules.state_from_fen('not a fen')


def test_no_cpu_fallback(mcaz_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from minitchess_alphazero_b200 import rules, _lib
    with pytest.raises(_lib.McazError) as e:
        rules.legal_moves(rules.state_from_fen(rules.STARTING_FEN))
    assert e.value.code == -2


def test_integration_md_stub_matches_the_abi(mcaz_lib):
    """The raw ctypes stub INTEGRATION.md shows a maintainer must describe the structs the library was built with
    (field order and sizes), or the documented binding would corrupt az_config."""
    import ctypes
    text = open(os.path.join(REPO, 'INTEGRATION.md')).read()
    block = re.search(r'## 3\. Raw ctypes stub.*?```python\n(.*?)```', text, flags=re.S).group(1)
    # only the declarations: stop before the first call that needs a device
    decl = block.split('def check(rc)')[0].replace("L = ctypes.CDLL('minitchess_alphazero_b200/libmcaz.so')", 'L = _L')
    ns = {'_L': mcaz_lib}
    exec(decl, ns)
    mcaz_lib.mcaz_struct_size.restype = ctypes.c_size_t
    assert ns['STATE'].itemsize == mcaz_lib.mcaz_struct_size(0)
    assert ctypes.sizeof(ns['McRules']) == mcaz_lib.mcaz_struct_size(1)
    assert ctypes.sizeof(ns['AzConfig']) == mcaz_lib.mcaz_struct_size(2)
    # same field names, order and types as the package's own mirror
    from minitchess_alphazero_b200 import _lib
    assert [(n, t) for n, t in ns['McRules']._fields_] == [(n, t) for n, t in _lib.Rules._fields_]
    ours = [(n, ctypes.sizeof(t)) for n, t in _lib.Config._fields_]
    theirs = [(n, ctypes.sizeof(t)) for n, t in ns['AzConfig']._fields_]
    assert ours == theirs
