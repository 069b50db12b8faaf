"""GPU: throughput mode replaces numpy's global RNG by device streams -- root noise Philox4x32-10 -> Marsaglia-Tsang
gamma -> Dirichlet(0.6) (csrc/philox.cuh, az::root_noise_gammas) in place of `np.random.dirichlet` (exp/agent.py:82) and
the device move choice (`play_device_one`) in place of `np.random.choice` (exp/agent.py:113-118).  The streams cannot be
bit-compared; what must hold is the distribution.  Kolmogorov-Smirnov on the Dirichlet marginals (against the exact
Beta law and against numpy's sampler) and chi-square on the move frequencies, early (fullmove < tau_change: proportional
to the visit counts) and late (uniform among the maxima).  Fixed seeds: the outcomes are deterministic."""
import numpy as np
import pytest
from scipy import stats

from oracle import rules_c as rc

pytestmark = pytest.mark.gpu
ALPHA = 0.6


@pytest.mark.parametrize('E', [2, 6, 13, 40])
def test_device_dirichlet_has_the_law_of_numpy_dirichlet(mcaz_lib, E):
    from minitchess_alphazero_b200.engine import sample_root_noise
    n = 40000
    x = sample_root_noise(seed=2026, alpha=ALPHA, n_edges=E, n=n)
    assert x.shape == (n, E) and (x > 0).all() and np.abs(x.sum(1) - 1).max() < 1e-12
    ref = np.random.RandomState(7).dirichlet([ALPHA] * E, size=n)
    for i in sorted({0, E // 2, E - 1}):
        # marginal of Dirichlet(alpha 1_E): Beta(alpha, (E - 1) alpha)
        assert stats.kstest(x[:, i], 'beta', args=(ALPHA, (E - 1) * ALPHA)).pvalue > 1e-3, i
        assert stats.ks_2samp(x[:, i], ref[:, i]).pvalue > 1e-3, i
    # second moments: Var = a(a0 - a) / (a0^2 (a0 + 1)), Cov = -a^2 / (a0^2 (a0 + 1)) with a0 = E alpha
    a0 = E * ALPHA
    var = ALPHA * (a0 - ALPHA) / (a0 * a0 * (a0 + 1))
    assert abs(x[:, 0].var() - var) < 0.05 * var + 1e-5
    assert abs(x.mean(0) - 1.0 / E).max() < 4 * np.sqrt(var / n) + 1e-4
    if E > 1:
        cov = -ALPHA * ALPHA / (a0 * a0 * (a0 + 1))
        assert abs(np.cov(x[:, 0], x[:, E - 1])[0, 1] - cov) < 0.1 * abs(cov) + 1e-4
    # samples of different game slots / simulations are independent draws
    assert abs(np.corrcoef(x[:-1, 0], x[1:, 0])[0, 1]) < 0.02
    y = sample_root_noise(seed=2027, alpha=ALPHA, n_edges=E, n=1000)
    assert not np.array_equal(x[:1000], y) and np.array_equal(x[:1000], sample_root_noise(2026, ALPHA, E, 1000))


def flat_tree_engine(fen, G, sims, seed):
    """All G games on one position, searched with uniform priors and value 0 through the external-evaluator calls: Q stays 0,
    so PUCT's first-max walks the edges round-robin (exp/agent.py:84-85) and after 1 + sims simulations the visit counts of
    every game are the same known vector."""
    from minitchess_alphazero_b200.engine import Engine
    from minitchess_alphazero_b200._lib import MC_MAX_MOVES
    st = rc.fens_to_states([fen] * G)
    E = int(rc.legal_moves(st[:1])[1][0])
    eng = Engine(G, max_sims_per_move=sims + 1, node_capacity=64, device_rng=1, dirichlet_epsilon=0.0, seed=seed)
    eng.reset_games(states=st)
    # the same prior for every edge of every node (the vector is wide enough for any position): with Q = 0 only the visit
    # counts tell the edges apart
    wide = np.full((G, MC_MAX_MOVES), 1.0 / MC_MAX_MOVES, dtype=np.float32)
    val = np.zeros(G, dtype=np.float32)
    for k in range(sims + 1):
        eng.select_expand()
        eng.backup(val, priors=wide)
    codes, visits, _, n_legal = eng.root_stats(want_q=False)
    assert (n_legal == E).all() and (visits == visits[0]).all() and visits[0].sum() == sims
    return eng, st, codes[0, :E].astype(int), visits[0, :E].astype(np.float64)


@pytest.mark.parametrize('fen,sims,early', [('2nbk/2ppp/5/5/PPP2/KBN2 w 0 1', 6 * 3 + 2, True),       # N = 4,4,3,3,3,3 -> pi
                                             ('2nbk/2ppp/5/5/PPP2/KBN2 w 0 3', 6 * 2, True),          # uniform pi
                                             ('2nbk/2ppp/5/5/PPP2/KBN2 w 0 9', 6 * 2, False),         # all six tied: uniform among maxima
                                             ('2nbk/2ppp/5/5/PPP2/KBN2 w 0 9', 6 * 2 + 2, False)])    # two maxima
def test_device_move_choice_follows_select_action(mcaz_lib, fen, sims, early):
    G = 8192
    eng, st, codes, N = flat_tree_engine(fen, G, sims, seed=11)
    E = len(codes)
    if early:
        p = N / N.sum()                                               # np.random.choice(legal, p=pi), exp/agent.py:115
    else:
        p = (N == N.max()).astype(float)
        p /= p.sum()                                                  # uniform among the maxima, exp/agent.py:117-118
    eng.play_device()
    tuples = eng.drain_replay()                                       # nothing finished: the replay queue is empty ...
    assert len(tuples) == 0
    after, results = eng.game_states()
    # ... so read the moves off the positions: apply every candidate with the oracle and match
    children = [rc.apply(st[:1], np.array([c], dtype=np.uint16))[0][0] for c in codes]
    counts = np.array([(after == ch).sum() for ch in children], dtype=np.float64)
    assert counts.sum() == G
    assert (counts[p == 0] == 0).all()                                # never a move outside the support
    live = p > 0
    chi = stats.chisquare(counts[live], G * p[live])
    assert chi.pvalue > 1e-3, (counts, G * p)
    # and it is a per-game stream: another seed gives other picks with the same law
    eng2, _, _, _ = flat_tree_engine(fen, G, sims, seed=12)
    eng2.play_device()
    after2, _ = eng2.game_states()
    assert not np.array_equal(after, after2)
    counts2 = np.array([(after2 == ch).sum() for ch in children], dtype=np.float64)
    assert stats.chisquare(counts2[live], G * p[live]).pvalue > 1e-3
