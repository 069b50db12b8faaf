"""The 554-code action table of exp/moves_dict.json, regenerated from the library's own formula
(mc_code_squares) instead of shipping the reference's JSON.  `dump_json` writes a byte-identical
`moves_dict.json` for code that opens it from the cwd like exp/environment.py:16 does."""
import ctypes
import json

from . import _lib

NUM_ACTIONS = _lib.MC_NUM_ACTIONS
FILES = 'abcde'


def square_name(sq):
    return FILES[sq % 5] + str(sq // 5 + 1)


def _build():
    L = _lib.lib()
    f, t = ctypes.c_int(), ctypes.c_int()
    fwd = {True: {}, False: {}}
    for white in (True, False):
        for code in range(NUM_ACTIONS):
            _lib.check(L.mc_code_squares(code, int(white), ctypes.byref(f), ctypes.byref(t)))
            fwd[white][square_name(f.value) + square_name(t.value)] = code
    return fwd


MOVES_DICT = _build()                                                     # exp/environment.py:18
MOVES_DICT_INV = {side: {v: k for k, v in MOVES_DICT[side].items()} for side in (True, False)}   # :19


def as_json_text():
    return json.dumps({'w': MOVES_DICT[True], 'b': MOVES_DICT[False]})


def dump_json(path='moves_dict.json'):
    with open(path, 'w') as f:
        f.write(as_json_text())
