"""ctypes binding of libmcaz.so (include/mcaz.h).  Fails loudly when the library is missing:
there is no Python or CPU fallback for any compute entry point."""
import ctypes
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(PKG, 'libmcaz.so')

MC_MAX_MOVES = 96
MC_NUM_ACTIONS = 554
MC_TOKENS = 60
AZ_NUM_PARAMS = 10693458
AZ_NUM_BN_STATS = 9734
AZ_NUM_WEIGHT_FLOATS = AZ_NUM_PARAMS + AZ_NUM_BN_STATS
AZ_NUM_COUNTERS = 18

STATE_DTYPE = np.dtype([('pl0', '<u4'), ('pl1', '<u4'), ('pl2', '<u4'), ('white', '<u4'), ('meta', '<u4')])
RESULT_STRINGS = {0: '*', 1: '1-0', 2: '0-1', 3: '1/2-1/2'}


class McazError(RuntimeError):
    def __init__(self, code, message):
        super().__init__('libmcaz error %d: %s' % (code, message))
        self.code = code


ABI_VERSION = 6      # MCAZ_ABI_VERSION of include/mcaz.h


class Rules(ctypes.Structure):
    _fields_ = [('pawn_double_step', ctypes.c_int32), ('promo_multiplicity', ctypes.c_int32),
                ('max_fullmoves', ctypes.c_int32), ('insufficient_material', ctypes.c_int32),
                ('fivefold_repetition', ctypes.c_int32)]


class Config(ctypes.Structure):
    _fields_ = [('n_games', ctypes.c_int32), ('max_sims_per_move', ctypes.c_int32),
                ('node_capacity', ctypes.c_int32), ('edge_capacity', ctypes.c_int32),
                ('cpuct', ctypes.c_float), ('tau_change', ctypes.c_int32),
                ('dirichlet_alpha', ctypes.c_float), ('dirichlet_epsilon', ctypes.c_float),
                ('numpy1_dtype_flow', ctypes.c_int32), ('device_rng', ctypes.c_int32),
                ('seed', ctypes.c_uint64), ('rules', Rules), ('network', ctypes.c_int32),
                ('leaves_per_step', ctypes.c_int32), ('own_stream', ctypes.c_int32),
                ('eval_cache_log2', ctypes.c_int32), ('free_sims', ctypes.c_int32), ('recycle', ctypes.c_int32),
                ('lookahead_rows', ctypes.c_int32), ('fp8_convolutions', ctypes.c_int32),
                ('defer_rows', ctypes.c_int32)]


_lib = None


def lib():
    """The loaded library.  Raises if it has not been built (python -m minitchess_alphazero_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError('libmcaz.so is not built: run `python -m minitchess_alphazero_b200.build` '
                              '(needs nvcc; there is no CPU fallback)')
        L = ctypes.CDLL(SO_PATH)
        L.mcaz_last_error.restype = ctypes.c_char_p
        L.mcaz_kernel_launches.restype = ctypes.c_uint64
        if L.mcaz_abi_version() != ABI_VERSION:
            raise ImportError('libmcaz.so ABI mismatch')
        L.mcaz_struct_size.restype = ctypes.c_size_t
        for which, mirror in ((0, STATE_DTYPE.itemsize), (1, ctypes.sizeof(Rules)), (2, ctypes.sizeof(Config))):
            if L.mcaz_struct_size(which) != mirror:
                raise ImportError('libmcaz.so struct %d is %d bytes, the Python mirror %d: rebuild the library'
                                  % (which, L.mcaz_struct_size(which), mirror))
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise McazError(rc, lib().mcaz_last_error().decode(errors='replace'))


def ptr(x):
    """void* of a numpy array, a torch tensor, an int address or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return ctypes.c_void_p(x)
    if isinstance(x, np.ndarray):
        assert x.flags['C_CONTIGUOUS']
        return ctypes.c_void_p(x.ctypes.data)
    if hasattr(x, 'data_ptr'):
        assert x.is_contiguous()
        return ctypes.c_void_p(x.data_ptr())
    if isinstance(x, ctypes.Structure):
        return ctypes.cast(ctypes.pointer(x), ctypes.c_void_p)
    raise TypeError(type(x))


def default_rules():
    r = Rules()
    lib().mc_default_rules(ctypes.byref(r))
    return r
