"""MinitChess rules on the GPU (mc_* entry points of include/mcaz.h).

Positions travel as numpy records of `STATE_DTYPE` (the packed `mc_state`) or as device
tensors viewed as int32[n, 5].  Mirrors what exp/environment.py obtains from python-chess.
"""
import ctypes

import numpy as np

from ._lib import (MC_MAX_MOVES, MC_TOKENS, STATE_DTYPE, RESULT_STRINGS, check, lib, ptr)

STARTING_FEN = '2nbk/2ppp/5/5/PPP2/KBN2 w 0 1'   # exp/environment.py:6


def state_from_fen(fen):
    s = np.zeros(1, dtype=STATE_DTYPE)
    check(lib().mc_state_from_fen(fen.encode(), ptr(s)))
    return s[0]


def states_from_fens(fens):
    out = np.zeros(len(fens), dtype=STATE_DTYPE)
    L = lib()
    for i, fen in enumerate(fens):
        check(L.mc_state_from_fen(fen.encode(), ctypes.c_void_p(out.ctypes.data + i * STATE_DTYPE.itemsize)))
    return out


def state_to_fen(state):
    s = np.ascontiguousarray(np.atleast_1d(state), dtype=STATE_DTYPE)
    buf = ctypes.create_string_buffer(64)
    check(lib().mc_state_to_fen(ptr(s), buf, 64))
    return buf.value.decode()


def _rules_ptr(rules):
    return ctypes.byref(rules) if rules is not None else None


def legal_moves(states, rules=None):
    """-> codes uint16[n, MC_MAX_MOVES] (sorted, first counts[i] valid), counts int32[n], results int8[n]."""
    states = np.ascontiguousarray(np.atleast_1d(states), dtype=STATE_DTYPE)
    n = len(states)
    codes = np.zeros((n, MC_MAX_MOVES), dtype=np.uint16)
    counts = np.zeros(n, dtype=np.int32)
    results = np.zeros(n, dtype=np.int8)
    check(lib().mc_legal_moves(ptr(states), n, _rules_ptr(rules), ptr(codes), ptr(counts), ptr(results)))
    return codes, counts, results


def apply(states, codes, rules=None):
    """-> (next states, status int8[n]: 0 ok, 1 illegal, 2 finished)."""
    states = np.ascontiguousarray(np.atleast_1d(states), dtype=STATE_DTYPE)
    codes = np.ascontiguousarray(np.atleast_1d(codes), dtype=np.uint16)
    n = len(states)
    out = np.zeros(n, dtype=STATE_DTYPE)
    status = np.zeros(n, dtype=np.int8)
    check(lib().mc_apply(ptr(states), ptr(codes), n, _rules_ptr(rules), ptr(out), ptr(status)))
    return out, status


def perft(states, depth, rules=None):
    states = np.ascontiguousarray(np.atleast_1d(states), dtype=STATE_DTYPE)
    nodes = np.zeros(len(states), dtype=np.uint64)
    check(lib().mc_perft(ptr(states), len(states), int(depth), _rules_ptr(rules), ptr(nodes)))
    return nodes


def tokenize(states):
    """Network.process_observation (exp/policy.py:96-105) -> tokens uint8[n, 60], clocks float32[n]."""
    states = np.ascontiguousarray(np.atleast_1d(states), dtype=STATE_DTYPE)
    n = len(states)
    tokens = np.zeros((n, MC_TOKENS), dtype=np.uint8)
    clocks = np.zeros(n, dtype=np.float32)
    check(lib().mc_tokenize(ptr(states), n, ptr(tokens), ptr(clocks)))
    return tokens, clocks


def result_string(code):
    return RESULT_STRINGS[int(code)]
