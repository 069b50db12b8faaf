"""Python handle over the az_* entry points of include/mcaz.h: the batched AlphaZero search engine.

One `Engine` owns `n_games` concurrent games with two GPU-resident trees each (one per colour's
agent, like the two `SimpleAlphaZeroAgent`s of app/base.py:113).  Search runs either with the
built-in network (`search`) or one simulation at a time around an external evaluator
(`select_expand` -> evaluate the leaf batch -> `backup`), which is also how the parity tests feed
the reference MCTS and the engine identical priors.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import (AZ_NUM_COUNTERS, AZ_NUM_WEIGHT_FLOATS, MC_MAX_MOVES, MC_NUM_ACTIONS, MC_TOKENS, STATE_DTYPE, Config,
                   check, ptr)

COUNTER_NAMES = ('simulations', 'evaluations', 'terminal_leaves', 'moves', 'games_finished', 'nodes', 'edges',
                 'kernel_launches', 'collisions', 'cached_evaluations', 'path_depth', 'path_edges', 'recycled_nodes', 'duplicate_rows',
                 'replay_dropped', 'deferred_rows', 'trimmed_batches', 'evicted_nodes')

REPLAY_DTYPE = np.dtype([('observation', STATE_DTYPE), ('n_legal', '<u2'), ('action', '<u2'), ('reward', 'i1'),
                         ('pad', 'u1', 3), ('weights_version', '<u4'), ('codes', '<u2', MC_MAX_MOVES), ('pi', '<f4', MC_MAX_MOVES)])
NODE_TERMINAL, NODE_DECISIVE = 1 << 16, 1 << 17      # MC_NODE_* bits of az_tree_dump's info words


assert REPLAY_DTYPE.itemsize == 608          # az_replay_tuple; checked against the library in Engine.__init__


class _CudaView:
    """Engine-owned device memory exposed through __cuda_array_interface__ (zero copy into torch)."""

    def __init__(self, address, shape, typestr):
        self.__cuda_array_interface__ = {'data': (int(address), False), 'shape': tuple(shape), 'typestr': typestr,
                                         'version': 2, 'strides': None}


def default_config(**overrides):
    cfg = Config()
    _lib.lib().az_default_config(ctypes.byref(cfg))
    for k, v in overrides.items():
        if k == 'rules':
            cfg.rules = v
        else:
            if not hasattr(cfg, k):
                raise TypeError('unknown engine option %r' % k)
            setattr(cfg, k, v)
    return cfg


class Engine:
    def __init__(self, n_games=1, _backend=None, **options):
        # `_backend` is a test hook: the CPU suite passes the host build of csrc/mcts_core.cuh
        # (tests/host_harness).  The product always uses libmcaz.so.
        self._L = _backend if _backend is not None else _lib.lib()
        self._host = _backend is not None
        if not self._host and self._L.mcaz_struct_size(3) != REPLAY_DTYPE.itemsize:
            raise ImportError('az_replay_tuple size mismatch between libmcaz.so and engine.REPLAY_DTYPE')
        cfg = Config()
        self._L.az_default_config(ctypes.byref(cfg))
        cfg.n_games = int(n_games)
        for k, v in options.items():
            if not hasattr(cfg, k):
                raise TypeError('unknown engine option %r' % k)
            setattr(cfg, k, v)
        self.config = cfg
        self.n_games = int(n_games)
        self.n_slots = self.n_games * max(1, int(cfg.leaves_per_step))      # rows of the leaf batch
        self._h = ctypes.c_void_p()
        self._check(self._L.az_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        self._noise_used = np.zeros(self.n_games, dtype=np.uint8)

    def _check(self, rc):
        if rc != 0:
            msg = self._L.mcaz_last_error()
            if isinstance(msg, int):
                msg = ctypes.cast(msg, ctypes.c_char_p).value
            raise _lib.McazError(rc, (msg or b'').decode(errors='replace'))

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self._L.az_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ games
    @staticmethod
    def _ids(game_ids):
        return None if game_ids is None else np.ascontiguousarray(game_ids, dtype=np.int32)

    def reset_games(self, game_ids=None, states=None):
        """MonteCarloInit.on_episode_begin (exp/callbacks.py:57-62) + env.new_episode for the listed games."""
        ids = self._ids(game_ids)
        n = self.n_games if ids is None else len(ids)
        st = None if states is None else np.ascontiguousarray(states, dtype=STATE_DTYPE)
        self._check(self._L.az_reset_games(self._h, ptr(ids), n, ptr(st)))

    def reset_trees(self, tree_ids):
        """Empty single trees (tree id = 2 * game + k): `init_mcts()` of one of two agents sharing the engine."""
        ids = np.ascontiguousarray(np.atleast_1d(tree_ids), dtype=np.int32)
        self._check(self._L.az_reset_trees(self._h, ptr(ids), len(ids)))

    def set_positions(self, states, game_ids=None, trees=None):
        ids = self._ids(game_ids)
        st = np.ascontiguousarray(np.atleast_1d(states), dtype=STATE_DTYPE)
        tr = None if trees is None else np.ascontiguousarray(trees, dtype=np.int32)
        self._check(self._L.az_set_positions(self._h, ptr(ids), len(st), ptr(st), ptr(tr)))

    def play(self, codes, game_ids=None):
        ids = self._ids(game_ids)
        codes = np.ascontiguousarray(np.atleast_1d(codes), dtype=np.uint16)
        results = np.zeros(len(codes), dtype=np.int8)
        self._check(self._L.az_play(self._h, ptr(ids), ptr(codes), len(codes), ptr(results)))
        return results

    def play_device(self):
        self._check(self._L.az_play_device(self._h))

    def game_states(self, game_ids=None):
        ids = self._ids(game_ids)
        n = self.n_games if ids is None else len(ids)
        states = np.zeros(n, dtype=STATE_DTYPE)
        results = np.zeros(n, dtype=np.int8)
        self._check(self._L.az_game_states(self._h, ptr(ids), n, ptr(states), ptr(results)))
        return states, results

    # ------------------------------------------------------------- simulation
    def select_expand(self, noise=None, want_noise_used=False):
        """One PUCT descent + expansion per active game.  `noise`: float64 [n_games, MC_MAX_MOVES] or None."""
        if noise is not None and isinstance(noise, np.ndarray):
            noise = np.ascontiguousarray(noise, dtype=np.float64)
            assert noise.shape == (self.n_games, MC_MAX_MOVES)
        used = self._noise_used if want_noise_used else None
        self._check(self._L.az_select_expand(self._h, ptr(noise), ptr(used)))
        return used

    def _leaf_pointers(self):
        tok, clk, need, st = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        n = ctypes.c_int()
        self._check(self._L.az_leaf_batch(self._h, ctypes.byref(tok), ctypes.byref(clk), ctypes.byref(need),
                                          ctypes.byref(st), ctypes.byref(n)))
        return tok.value, clk.value, need.value, st.value, n.value

    def leaf_batch_device(self):
        """Zero-copy torch views of the leaf batch: tokens uint8[G,60], clocks float32[G], needs_eval uint8[G]."""
        import torch
        tok, clk, need, _st, n = self._leaf_pointers()
        assert not self._host
        return (torch.as_tensor(_CudaView(tok, (n, MC_TOKENS), '|u1'), device='cuda'),
                torch.as_tensor(_CudaView(clk, (n,), '<f4'), device='cuda'),
                torch.as_tensor(_CudaView(need, (n,), '|u1'), device='cuda'))

    def leaf_batch(self):
        """Host copies: tokens, clocks, needs_eval, leaf_states (numpy)."""
        tok, clk, need, st, n = self._leaf_pointers()
        if self._host:
            def view(addr, dtype, shape):
                count = int(np.prod(shape))
                buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(addr)
                return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape).copy()
            return (view(tok, np.uint8, (n, MC_TOKENS)), view(clk, np.float32, (n,)), view(need, np.uint8, (n,)),
                    view(st, STATE_DTYPE, (n,)))
        import torch
        tokens = torch.as_tensor(_CudaView(tok, (n, MC_TOKENS), '|u1'), device='cuda').cpu().numpy()
        clocks = torch.as_tensor(_CudaView(clk, (n,), '<f4'), device='cuda').cpu().numpy()
        needs = torch.as_tensor(_CudaView(need, (n,), '|u1'), device='cuda').cpu().numpy()
        raw = torch.as_tensor(_CudaView(st, (n, 5), '<i4'), device='cuda').cpu().numpy()
        return tokens, clocks, needs, raw.view(np.uint32).reshape(-1).view(STATE_DTYPE)

    def backup(self, values, logits=None, priors=None):
        """exp/agent.py:67-72 + :47-52.  logits [G,554] / priors [G,MC_MAX_MOVES] / values [G], float32."""
        def prep(x, shape):
            if x is None:
                return None
            if isinstance(x, np.ndarray):
                x = np.ascontiguousarray(x, dtype=np.float32)
                assert x.shape == shape, (x.shape, shape)
            return x
        self._check(self._L.az_backup(self._h, ptr(prep(logits, (self.n_slots, MC_NUM_ACTIONS))),
                                      ptr(prep(values, (self.n_slots,))),
                                      ptr(prep(priors, (self.n_slots, MC_MAX_MOVES)))))

    def eval_backup(self):
        """Evaluate the pending leaf batch with the built-in network and back it up (one simulation's second half)."""
        self._check(self._L.az_eval_backup(self._h))

    def search(self, n_sims):
        """n_sims simulations for every active game with the built-in network."""
        self._check(self._L.az_search(self._h, int(n_sims)))

    def search_noise(self, noise):
        """len(noise) simulations with the built-in network; noise: float64 [n_sims, n_games, MC_MAX_MOVES], row s is
        the root Dirichlet sample of simulation s (used where the root is expanded by then, exp/agent.py:81-82)."""
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        assert noise.ndim == 3 and noise.shape[1:] == (self.n_games, MC_MAX_MOVES), noise.shape
        self._check(self._L.az_search_noise(self._h, int(noise.shape[0]), ptr(noise)))

    def selfplay(self, n_steps, sims_per_move):
        """Continuous self-play (az_selfplay): n_steps network batches; every game searches, moves, records and
        restarts on its own inside the search kernel, so the batch stays full whatever each move costs."""
        self._check(self._L.az_selfplay(self._h, int(n_steps), int(sims_per_move)))

    # ------------------------------------------------------------------ read-back
    def root_stats(self, game_ids=None, want_q=True, out=None):
        """-> codes uint16[n,M], visits uint32[n,M], q float64[n,M] or None, n_legal int32[n] (-1: not visited).
        `out` = (codes, visits, n_legal) arrays to fill instead of fresh ones (with want_q=False) -- page-locked buffers kept by the
        caller make the read-back an asynchronous copy instead of a staged one into freshly mapped pages."""
        ids = self._ids(game_ids)
        n = self.n_games if ids is None else len(ids)
        if out is not None:
            assert not want_q, 'out= carries no Q buffer'
            codes, visits, n_legal = out
            assert codes.shape == (n, MC_MAX_MOVES) and codes.dtype == np.uint16 and codes.flags.c_contiguous
            assert visits.shape == (n, MC_MAX_MOVES) and visits.dtype == np.uint32 and visits.flags.c_contiguous
            assert n_legal.shape == (n,) and n_legal.dtype == np.int32
            self._check(self._L.az_root_stats(self._h, ptr(ids), n, ptr(codes), ptr(visits), None, ptr(n_legal)))
            return codes, visits, None, n_legal
        codes = np.zeros((n, MC_MAX_MOVES), dtype=np.uint16)
        visits = np.zeros((n, MC_MAX_MOVES), dtype=np.uint32)
        q = np.zeros((n, MC_MAX_MOVES), dtype=np.float64) if want_q else None
        n_legal = np.zeros(n, dtype=np.int32)
        self._check(self._L.az_root_stats(self._h, ptr(ids), n, ptr(codes), ptr(visits), ptr(q), ptr(n_legal)))
        return codes, visits, q, n_legal

    def node_stats(self, game_id, tree, state):
        """MonteCarloTreeSearch.__getitem__ for one position (exp/agent.py:38-39); None if never visited."""
        st = np.ascontiguousarray(np.atleast_1d(state), dtype=STATE_DTYPE)
        found, n_legal, term = ctypes.c_int(), ctypes.c_int32(), ctypes.c_int()
        tval = ctypes.c_double()
        codes = np.zeros(MC_MAX_MOVES, dtype=np.uint16)
        visits = np.zeros(MC_MAX_MOVES, dtype=np.uint32)
        q = np.zeros(MC_MAX_MOVES, dtype=np.float64)
        pri = np.zeros(MC_MAX_MOVES, dtype=np.float32)
        self._check(self._L.az_node_stats(self._h, int(game_id), int(tree), ptr(st), ctypes.byref(found), ptr(codes),
                                          ptr(visits), ptr(q), ptr(pri), ctypes.byref(n_legal), ctypes.byref(term),
                                          ctypes.byref(tval)))
        if not found.value:
            return None
        E = n_legal.value
        return {'legal_moves': codes[:E].astype(int).tolist(), 'N': visits[:E].astype(np.float64), 'Q': q[:E].copy(),
                'P': pri[:E].copy(), 'terminal': tval.value if term.value else None}

    def tree_dump(self, game_id=0, tree=0):
        """Every node of one tree (az_tree_dump) -- the whole dicts of MonteCarloTreeSearch (exp/agent.py:25-36) in creation
        order: {'states' STATE_DTYPE[n], 'info' uint32[n] (edges | NODE_TERMINAL | NODE_DECISIVE), 'edge_off' uint32[n],
        'codes' uint16[m], 'N' uint32[m], 'Q' float64[m], 'P' float32[m]}."""
        n, m = ctypes.c_int(), ctypes.c_int()
        self._check(self._L.az_tree_dump(self._h, int(game_id), int(tree), 0, None, None, None, ctypes.byref(n), 0, None, None, None,
                                         None, ctypes.byref(m)))
        nn, mm = n.value, m.value
        states = np.zeros(nn, dtype=STATE_DTYPE)
        info, off = np.zeros(nn, dtype=np.uint32), np.zeros(nn, dtype=np.uint32)
        codes, visits = np.zeros(mm, dtype=np.uint16), np.zeros(mm, dtype=np.uint32)
        q, pri = np.zeros(mm, dtype=np.float64), np.zeros(mm, dtype=np.float32)
        self._check(self._L.az_tree_dump(self._h, int(game_id), int(tree), nn, ptr(states), ptr(info), ptr(off), ctypes.byref(n), mm,
                                         ptr(codes), ptr(visits), ptr(q), ptr(pri), ctypes.byref(m)))
        assert (n.value, m.value) == (nn, mm)
        return {'states': states, 'info': info, 'edge_off': off, 'codes': codes, 'N': visits, 'Q': q, 'P': pri}

    def counters(self):
        out = np.zeros(AZ_NUM_COUNTERS, dtype=np.uint64)
        self._check(self._L.az_counters(self._h, ptr(out)))
        return dict(zip(COUNTER_NAMES, (int(x) for x in out)))

    # ------------------------------------------------------------------ network
    def set_weights(self, flat, version=None):
        """flat: float32 [AZ_NUM_WEIGHT_FLOATS] numpy array or CUDA tensor (policy.flatten_state_dict).  `version`: the
        learner's stamp of these weights (LearnPuppet.weights_version, app/base.py:171-174) as an unsigned 32-bit number;
        default: the engine counts its uploads.  Finished games' replay tuples carry it (`weights_version`)."""
        n = flat.numel() if hasattr(flat, 'numel') else flat.size
        assert n == AZ_NUM_WEIGHT_FLOATS, n
        if hasattr(flat, 'is_cuda') and flat.is_cuda:      # produced on torch's stream, consumed on the engine's
            import torch
            torch.cuda.current_stream().synchronize()
        self._check(self._L.az_set_weights(self._h, ptr(flat), ctypes.c_size_t(n)))
        if version is not None:
            self.set_weights_version(version)

    def set_weights_version(self, version):
        self._check(self._L.az_set_weights_version(self._h, ctypes.c_uint32(int(version) & 0xffffffff)))

    def network_forward(self, tokens, clocks):
        tokens = np.ascontiguousarray(tokens, dtype=np.uint8).reshape(-1, MC_TOKENS)
        clocks = np.ascontiguousarray(clocks, dtype=np.float32).reshape(-1)
        n = len(tokens)
        logits = np.zeros((n, MC_NUM_ACTIONS), dtype=np.float32)
        values = np.zeros(n, dtype=np.float32)
        self._check(self._L.az_network_forward(self._h, ptr(tokens), ptr(clocks), n, ptr(logits), ptr(values)))
        return logits, values

    def drain_replay(self, max_tuples=None):
        cap = self.n_games * 64 if max_tuples is None else int(max_tuples)
        out = np.zeros(cap, dtype=REPLAY_DTYPE)
        n = ctypes.c_int()
        self._check(self._L.az_drain_replay(self._h, ptr(out), cap, ctypes.byref(n)))
        return out[:n.value]


def sample_root_noise(seed, alpha, n_edges, n):
    """n Dirichlet(alpha) samples over n_edges edges from the device sampler of throughput mode (az_sample_root_noise)."""
    out = np.zeros((int(n), int(n_edges)), dtype=np.float64)
    check(_lib.lib().az_sample_root_noise(ctypes.c_uint64(int(seed)), ctypes.c_float(float(alpha)), int(n_edges), int(n), ptr(out)))
    return out


def _profile_network(self, on=True, read=False):
    """bench hook: (avg ms per residual tower, forwards measured, kernel launches per tower) since the last call."""
    ms, n, lpf = ctypes.c_double(), ctypes.c_int(), ctypes.c_int()
    self._check(self._L.az_profile_network(self._h, int(bool(on)), ctypes.byref(ms) if read else None,
                                           ctypes.byref(n) if read else None, ctypes.byref(lpf)))
    return (ms.value, n.value, lpf.value) if read else None


def _profile_tree(self, on=True, read=False):
    """bench hook: (total ms in search_step_kernel, launches) since the last call."""
    ms, n = ctypes.c_double(), ctypes.c_int()
    self._check(self._L.az_profile_tree(self._h, int(bool(on)), ctypes.byref(ms) if read else None,
                                        ctypes.byref(n) if read else None))
    return (ms.value, n.value) if read else None


Engine.profile_network = _profile_network
Engine.profile_tree = _profile_tree
