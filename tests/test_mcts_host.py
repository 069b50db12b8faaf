"""CPU: tree logic of csrc/mcts_core.cuh (host build) bit-exact against the reference MCTS.

Golden side: tests/golden/mcts_hash_game*.json were produced by the reference's UNMODIFIED
exp/agent.py; the live side is its pinned restatement oracle/ref_selfplay.py."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import ref_selfplay as rs
from oracle import rules_c as rc
from oracle.hash_eval import hash_evaluate
import parity_common as pc


@pytest.fixture(scope='module')
def backend():
    return pc.host_backend()


def make_engine(backend, n_games, sims, **kw):
    from minitchess_alphazero_b200.engine import Engine
    return Engine(n_games, _backend=backend, max_sims_per_move=sims, **kw)


@pytest.mark.parametrize('name', ['mcts_hash_game.json', 'mcts_hash_game_s1.json'])
def test_golden_game_visit_counts_bit_exact(backend, name):
    g = load_golden(name)
    eng = make_engine(backend, 1, g['sims'])
    np.random.seed(g['seed'])
    records, states, results = pc.play_games(eng, hash_evaluate, g['sims'], [np.random])
    rows = records[0]
    assert len(rows) == len(g['plies'])
    for mine, ref in zip(rows, g['plies']):
        assert mine['observation'] == ref['observation']
        assert mine['legal_moves'] == ref['legal_moves']
        assert mine['N'] == ref['N']                      # visit counts: bit-exact
        assert mine['Q'] == ref['Q']                      # float64 running means: bit-exact
        assert mine['pi'] == ref['pi']
        assert mine['action'] == ref['action']
    assert rc.state_to_fen(states[0]) == g['plies'][-1]['next']
    # whole trees, through the restatement (itself pinned to the same golden by make_golden.py)
    np.random.seed(g['seed'])
    _, _, trees = rs.play_game(hash_evaluate, g['sims'])
    for t in (0, 1):
        pc.compare_with_tree(eng, 0, t, trees[t])
        assert [len(trees[t].N), len(trees[t].terminal)] == g['tree_sizes'][t]
    c = eng.counters()
    assert c['simulations'] == g['sims'] * len(rows)
    assert c['evaluations'] == trees[0].n_evals + trees[1].n_evals
    assert c['nodes'] == sum(len(t.visited) for t in trees)


def test_batched_games_match_restatement(backend):
    """Several concurrent games with independent RNG streams == independent reference games."""
    G, sims = 6, 24
    eng = make_engine(backend, G, sims)
    rngs = [np.random.RandomState(100 + g) for g in range(G)]
    records, states, results = pc.play_games(eng, hash_evaluate, sims, rngs)
    for g in range(G):
        ref_records, ep, trees = rs.play_game(hash_evaluate, sims, rng=np.random.RandomState(100 + g))
        assert len(records[g]) == len(ref_records)
        for a, b in zip(records[g], ref_records):
            assert a['observation'] == b['observation'] and a['action'] == b['action'] and a['pi'] == b['pi']
        assert rc.state_to_fen(states[g]) == ep.fen
        for t in (0, 1):
            pc.compare_with_tree(eng, g, t, trees[t])


def test_no_noise_and_numpy1_flow_switch(backend):
    sims = 30
    for kw in ({'dirichlet_epsilon': 0.0}, {'cpuct': 1.5}):
        eng = make_engine(backend, 1, sims, **kw)
        eps = kw.get('dirichlet_epsilon', 0.25)
        records, states, _ = pc.play_games(eng, hash_evaluate, sims, [np.random.RandomState(5)], max_plies=12, epsilon=eps)
        ref_records, ep, trees = rs.play_game(hash_evaluate, sims, cpuct=kw.get('cpuct', 1), max_plies=12,
                                              rng=np.random.RandomState(5), epsilon=eps)
        assert [r['action'] for r in records[0]] == [r['action'] for r in ref_records]
        assert [r['pi'] for r in records[0]] == [r['pi'] for r in ref_records]
    # the numpy-1.x dtype flow (Q6) is a different arithmetic: it must run and may differ
    eng = make_engine(backend, 1, sims, numpy1_dtype_flow=1)
    pc.play_games(eng, hash_evaluate, sims, [np.random.RandomState(5)], max_plies=6)


def test_terminal_revisit_sign_flip(backend):
    """Q1: first visit of a mate backs up -reward, every revisit +reward (exp/agent.py:59-63 vs :75-77)."""
    fen = 'k4/5/1K3/5/5/2Q2 w 0 10'      # Qc1-c6 mates at once
    sims = 60
    eng = make_engine(backend, 1, sims, dirichlet_epsilon=0.0)
    st = rc.fens_to_states([fen])
    records, _, _ = pc.play_games(eng, hash_evaluate, sims, [np.random.RandomState(0)], max_plies=1, start_states=st, epsilon=0.0)
    tree = rs.RefTree(hash_evaluate, 1, epsilon=0.0)
    tree.simulate(sims, fen)
    assert records[0][0]['N'] == tree.N[fen].tolist()
    assert records[0][0]['Q'] == tree.Q[fen].tolist()
    assert len(tree.terminal) > 0 and any(v == -1.0 for v in tree.terminal.values())
    pc.compare_with_tree(eng, 0, 0, tree)


def test_fivefold_repetition_inside_a_simulation(backend):
    """Both kings can only shuffle (a1-b1, e6-d6), every pawn is blocked: the search path is forced, so the
    17th simulation reaches the start position for the fifth time along its own path and the reference's
    board.result() calls it a draw there.  The engine must stop at the same depth."""
    fen = '2p1k/2p1p/2P1P/p1p2/P1P2/K1P2 w 0 1'
    from oracle.ref_selfplay import RefEpisode
    ep = RefEpisode(fen)
    assert len(ep.legal) == 1 and not ep.done
    sims = 30
    tree = rs.RefTree(hash_evaluate, 1, epsilon=0.0)
    tree.simulate(sims, fen)
    assert any(v == 0 for v in tree.terminal.values()) and tree.n_evals == 16
    eng = make_engine(backend, 1, sims, dirichlet_epsilon=0.0)
    records, _, _ = pc.play_games(eng, hash_evaluate, sims, [np.random.RandomState(0)], max_plies=1,
                                  start_states=rc.fens_to_states([fen]), epsilon=0.0)
    assert records[0][0]['N'] == tree.N[fen].tolist() and records[0][0]['Q'] == tree.Q[fen].tolist()
    pc.compare_with_tree(eng, 0, 0, tree)
    assert eng.counters()['evaluations'] == 16


def test_capacity_overflow_fails_loudly(backend):
    from minitchess_alphazero_b200._lib import McazError
    eng = make_engine(backend, 1, 8, node_capacity=5, edge_capacity=64)
    with pytest.raises(McazError):
        pc.play_games(eng, hash_evaluate, 8, [np.random.RandomState(0)], max_plies=3)


def test_phase_errors(backend):
    from minitchess_alphazero_b200._lib import McazError
    eng = make_engine(backend, 1, 8)
    with pytest.raises(McazError):
        eng.backup(np.zeros(1, dtype=np.float32), priors=np.zeros((1, 96), dtype=np.float32))
    eng.select_expand()
    with pytest.raises(McazError):
        eng.select_expand()
