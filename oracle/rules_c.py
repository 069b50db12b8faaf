"""ctypes front-end of oracle/minitchess_oracle.c (CPU rules restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
SO_PATH = os.path.join(HERE, '_build', 'libmc_oracle.so')

STATE_DTYPE = np.dtype([('pl0', '<u4'), ('pl1', '<u4'), ('pl2', '<u4'), ('white', '<u4'), ('meta', '<u4')])
MAX_MOVES = 96
NUM_ACTIONS = 554
PIECES = '.prbnqk'   # type code = index, same numbering as the tokeniser alphabet (exp/policy.py:7)
RESULT_STR = {0: '*', 1: '1-0', 2: '0-1', 3: '1/2-1/2'}


class Rules(ctypes.Structure):
    _fields_ = [('pawn_double_step', ctypes.c_int32), ('promo_multiplicity', ctypes.c_int32),
                ('max_fullmoves', ctypes.c_int32), ('insufficient_material', ctypes.c_int32),
                ('fivefold_repetition', ctypes.c_int32)]

    @classmethod
    def default(cls):
        return cls(0, 1, 30, 1, 1)


def build(force=False):
    src = os.path.join(HERE, 'minitchess_oracle.c')
    if not force and os.path.exists(SO_PATH) and os.path.getmtime(SO_PATH) >= max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(REPO, 'include', 'mcaz.h'))):
        return SO_PATH
    os.makedirs(os.path.dirname(SO_PATH), exist_ok=True)
    subprocess.check_call(['gcc', '-O2', '-std=c11', '-shared', '-fPIC', '-ffp-contract=off',
                           '-I', os.path.join(REPO, 'include'), src, '-o', SO_PATH])
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_perft.restype = ctypes.c_uint64
        _lib.orc_perft.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _lib.orc_random_positions.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def _rp(rules):
    return ctypes.byref(rules) if rules is not None else None


def fen_to_state(fen):
    """4-field FEN (exp/environment.py:6) -> one packed mc_state record."""
    rows, turn, half, full = fen.split()
    s = np.zeros((), dtype=STATE_DTYPE)
    pl = [0, 0, 0]
    white = 0
    for i, row in enumerate(rows.split('/')):
        rank, f = 5 - i, 0
        for ch in row:
            if ch.isdigit():
                f += int(ch)
                continue
            sq = 5 * rank + f
            t = PIECES.index(ch.lower())
            for b in range(3):
                if (t >> b) & 1:
                    pl[b] |= 1 << sq
            if ch.isupper():
                white |= 1 << sq
            f += 1
    s['pl0'], s['pl1'], s['pl2'], s['white'] = pl[0], pl[1], pl[2], white
    s['meta'] = (1 if turn == 'w' else 0) | (int(half) << 8) | (int(full) << 16)
    return s


def state_to_fen(s):
    pl0, pl1, pl2, white, meta = int(s['pl0']), int(s['pl1']), int(s['pl2']), int(s['white']), int(s['meta'])
    rows = []
    for rank in range(5, -1, -1):
        row, run = '', 0
        for f in range(5):
            sq = 5 * rank + f
            t = ((pl0 >> sq) & 1) | (((pl1 >> sq) & 1) << 1) | (((pl2 >> sq) & 1) << 2)
            if t == 0:
                run += 1
                continue
            if run:
                row += str(run)
                run = 0
            ch = PIECES[t]
            row += ch.upper() if (white >> sq) & 1 else ch
        if run:
            row += str(run)
        rows.append(row)
    return '%s %s %d %d' % ('/'.join(rows), 'w' if meta & 1 else 'b', (meta >> 8) & 0xff, (meta >> 16) & 0xff)


def fens_to_states(fens):
    out = np.zeros(len(fens), dtype=STATE_DTYPE)
    for i, f in enumerate(fens):
        out[i] = fen_to_state(f)
    return out


def legal_moves(states, rules=None):
    """-> (codes [n, MAX_MOVES] u16, counts [n] i32, results [n] i8)"""
    states = np.ascontiguousarray(np.atleast_1d(states), dtype=STATE_DTYPE)
    n = len(states)
    codes = np.zeros((n, MAX_MOVES), dtype=np.uint16)
    counts = np.zeros(n, dtype=np.int32)
    results = np.zeros(n, dtype=np.int8)
    lib().orc_legal_moves_bulk(states.ctypes.data_as(ctypes.c_void_p), n, _rp(rules),
                               codes.ctypes.data_as(ctypes.c_void_p), counts.ctypes.data_as(ctypes.c_void_p),
                               results.ctypes.data_as(ctypes.c_void_p))
    return codes, counts, results


def apply(states, codes, rules=None):
    states = np.ascontiguousarray(np.atleast_1d(states), dtype=STATE_DTYPE)
    codes = np.ascontiguousarray(np.atleast_1d(codes), dtype=np.uint16)
    n = len(states)
    out = np.zeros(n, dtype=STATE_DTYPE)
    status = np.zeros(n, dtype=np.int8)
    lib().orc_apply_bulk(states.ctypes.data_as(ctypes.c_void_p), codes.ctypes.data_as(ctypes.c_void_p), n, _rp(rules),
                         out.ctypes.data_as(ctypes.c_void_p), status.ctypes.data_as(ctypes.c_void_p))
    return out, status


def perft(state, depth, rules=None):
    state = np.ascontiguousarray(np.atleast_1d(state), dtype=STATE_DTYPE)
    return int(lib().orc_perft(state.ctypes.data_as(ctypes.c_void_p), depth, _rp(rules)))


def tokenize(states):
    states = np.ascontiguousarray(np.atleast_1d(states), dtype=STATE_DTYPE)
    n = len(states)
    tokens = np.zeros((n, 60), dtype=np.uint8)
    clocks = np.zeros(n, dtype=np.float32)
    L = lib()
    for i in range(n):
        L.orc_tokenize(states[i:i + 1].ctypes.data_as(ctypes.c_void_p),
                       tokens[i].ctypes.data_as(ctypes.c_void_p), clocks[i:i + 1].ctypes.data_as(ctypes.c_void_p))
    return tokens, clocks


def random_positions(seed, n, rules=None, unique=True):
    """Positions visited by uniform-random legal playouts from STARTING_FEN."""
    out = np.zeros(n, dtype=STATE_DTYPE)
    lib().orc_random_positions(seed, n, _rp(rules), out.ctypes.data_as(ctypes.c_void_p))
    if unique:
        out = np.unique(out)
    return out


def start_state():
    s = np.zeros(1, dtype=STATE_DTYPE)
    lib().orc_start_state(s.ctypes.data_as(ctypes.c_void_p))
    return s[0]


def code_of(from_sq, to_sq, white_to_move):
    return int(lib().orc_code_of(int(from_sq), int(to_sq), int(bool(white_to_move))))
