"""Shared driver for the tree-logic parity tests (T1): plays games on an `Engine` exactly the way
the reference loop does (host numpy RNG, one simulation at a time, priors injected bit-exactly)
so visit counts can be compared with the reference MCTS / its restatement.  Used by the CPU suite
(host build of csrc/mcts_core.cuh) and by the GPU suite (libmcaz.so)."""
import hashlib

import numpy as np

from oracle import rules_c as rc
from minitchess_alphazero_b200._lib import MC_MAX_MOVES

ALPHA = 0.6


def synthetic_positions(seed, n):
    """Random piece placements (one king each, up to 10 other pieces, any mix): far denser in pins, double checks,
    adjacent kings and pieces en prise than reachable play -- inputs for the rules parity tests (CPU host build and GPU)."""
    rng = np.random.RandomState(seed)
    out = np.zeros(n, dtype=rc.STATE_DTYPE)
    for i in range(n):
        k = rng.randint(2, 12)
        sq = rng.permutation(30)[:k + 2]
        types = [6, 6] + list(rng.choice([1, 2, 3, 4, 5], size=k, p=[0.3, 0.2, 0.15, 0.15, 0.2]))
        white = [1, 0] + list(rng.randint(0, 2, size=k))
        pl = [0, 0, 0]
        w = 0
        for s_, t_, c_ in zip(sq, types, white):
            if t_ == 1 and (s_ < 5 or s_ >= 25):
                t_ = 4                                   # no pawns on the first or last rank
            for b in range(3):
                if (t_ >> b) & 1:
                    pl[b] |= 1 << int(s_)
            if c_:
                w |= 1 << int(s_)
        out[i] = (pl[0], pl[1], pl[2], w, int(rng.randint(0, 2)) | (int(rng.randint(0, 20)) << 8) | (int(rng.randint(1, 30)) << 16))
    return out


def host_backend():
    """CDLL of tests/host_harness/mcts_host.cpp (built on demand)."""
    import ctypes
    import os
    import subprocess
    from conftest import REPO
    src = os.path.join(REPO, 'tests', 'host_harness', 'mcts_host.cpp')
    out = os.path.join(REPO, 'tests', 'host_harness', '_build', 'libmcts_host.so')
    deps = [src] + [os.path.join(REPO, 'minitchess_alphazero_b200', 'csrc', f) for f in ('mcts_core.cuh', 'minitchess.cuh')]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(p) for p in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-ffp-contract=off',
                               '-I', os.path.join(REPO, 'include'),
                               '-I', os.path.join(REPO, 'minitchess_alphazero_b200', 'csrc'), src, '-o', out])
    L = ctypes.CDLL(out)
    L.mcaz_last_error.restype = ctypes.c_char_p
    return L


def play_games(engine, evaluate, sims, rngs, tau_change=6, max_plies=None, start_states=None, epsilon=0.25):
    """Plays engine.n_games games in lockstep.  rngs[g] is the numpy RNG of game g (np.random for the
    global legacy stream).  Returns per-game lists of ply records like oracle.ref_selfplay.play_game."""
    G = engine.n_games
    engine.reset_games(states=start_states)
    records = [[] for _ in range(G)]
    states, results = engine.game_states()
    plies = 0
    while (results == 0).any() and (max_plies is None or plies < max_plies):
        active = results == 0
        root_fens = [rc.state_to_fen(s) for s in states]
        for _ in range(sims):
            _, _, _, n_legal = engine.root_stats(want_q=False)
            noise = None
            if epsilon > 0 and (n_legal[active] > 0).any():
                noise = np.zeros((G, MC_MAX_MOVES))
                for g in np.nonzero(active & (n_legal > 0))[0]:
                    noise[g, :n_legal[g]] = rngs[g].dirichlet([ALPHA] * int(n_legal[g]))
            used = engine.select_expand(noise, want_noise_used=True)
            if noise is not None:
                assert np.array_equal(used.astype(bool), active & (n_legal > 0))
            tokens, clocks, needs, leaf_states = engine.leaf_batch()
            values = np.zeros(G, dtype=np.float32)
            priors = np.zeros((G, MC_MAX_MOVES), dtype=np.float32)
            idx = np.nonzero(needs)[0]
            if len(idx):
                codes, counts, _ = rc.legal_moves(np.ascontiguousarray(leaf_states[idx]))
                for k, g in enumerate(idx):
                    fen = rc.state_to_fen(leaf_states[g])
                    legal = codes[k, :counts[k]].astype(int).tolist()
                    p, v = evaluate(fen, legal)
                    assert np.float32(v) == v
                    priors[g, :len(legal)] = p
                    values[g] = v
            engine.backup(values, priors=priors)
        codes, visits, q, n_legal = engine.root_stats()
        actions = np.zeros(G, dtype=np.uint16)
        for g in np.nonzero(active)[0]:
            E = int(n_legal[g])
            legal = codes[g, :E].astype(int).tolist()
            N = visits[g, :E].astype(np.float64)
            pi = N / N.sum()
            if int(root_fens[g].split()[3]) < tau_change:
                action = rngs[g].choice(legal, p=pi)
            else:
                best = np.where(pi == pi.max())[0]
                action = legal[rngs[g].choice(best)]
            actions[g] = action
            records[g].append({'observation': root_fens[g], 'legal_moves': legal, 'N': N.tolist(),
                               'Q': q[g, :E].tolist(), 'pi': pi.tolist(), 'action': int(action)})
        ids = np.nonzero(active)[0].astype(np.int32)
        engine.play(actions[ids], game_ids=ids)
        states, results = engine.game_states()
        plies += 1
    return records, states, results


def tree_digest_from_engine(engine, game, tree, ref_tree):
    """sha256 over (fen, N, Q) of every node of the reference tree, values read back from the engine."""
    h = hashlib.sha256()
    for fen in sorted(ref_tree.N):
        st = engine.node_stats(game, tree, rc.fen_to_state(fen))
        assert st is not None, fen
        h.update(fen.encode()); h.update(st['N'].tobytes()); h.update(st['Q'].tobytes())
    for fen in sorted(ref_tree.terminal):
        st = engine.node_stats(game, tree, rc.fen_to_state(fen))
        assert st is not None and st['terminal'] is not None, fen
        h.update(fen.encode()); h.update(np.float64(st['terminal']).tobytes())
    return h.hexdigest()


def compare_with_tree(engine, game, tree, ref_tree):
    """Every node of the reference tree exists in the engine with bit-identical N, Q, P and codes."""
    for fen, N in ref_tree.N.items():
        st = engine.node_stats(game, tree, rc.fen_to_state(fen))
        assert st is not None, fen
        assert st['legal_moves'] == list(ref_tree.legal[fen]), fen
        assert np.array_equal(st['N'], N), (fen, st['N'], N)
        assert st['Q'].tobytes() == np.asarray(ref_tree.Q[fen]).tobytes() or np.array_equal(st['Q'], ref_tree.Q[fen]), fen
        assert np.array_equal(st['P'], ref_tree.P[fen]), fen
    for fen, val in ref_tree.terminal.items():
        st = engine.node_stats(game, tree, rc.fen_to_state(fen))
        assert st is not None and st['terminal'] is not None, fen
        assert st['terminal'] == val, fen
