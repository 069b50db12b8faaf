"""GPU: edge cases of the batched engine -- ragged game counts, chunked network batches, zero-length calls,
determinism of the device RNG, replay bookkeeping in throughput mode."""
import numpy as np
import pytest
import torch

from oracle import rules_c as rc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def net(mcaz_lib):
    from minitchess_alphazero_b200.policy import Network
    torch.manual_seed(0)
    return Network().eval()


def make(net, n_games, sims, **kw):
    from minitchess_alphazero_b200.engine import Engine
    from minitchess_alphazero_b200.policy import flatten_state_dict
    eng = Engine(n_games, max_sims_per_move=sims, network=1, device_rng=1, **kw)
    eng.set_weights(flatten_state_dict(net.state_dict(), device='cuda'))
    return eng


@pytest.mark.parametrize('n_games', [1, 3, 257, 1000])
def test_ragged_game_counts(net, n_games):
    sims = 10
    eng = make(net, n_games, sims, seed=7)
    eng.search(0)                                      # zero simulations: a no-op
    assert eng.counters()['simulations'] == 0
    for _ in range(3):
        eng.search(sims)
        codes, visits, _, n_legal = eng.root_stats(want_q=False)
        assert (n_legal > 0).all()
        assert (visits.sum(1) >= sims - 1).all()       # a fresh root gets sims-1 edge visits, a reused one more
        eng.play_device()
    c = eng.counters()
    assert c['simulations'] == 3 * sims * n_games and c['moves'] == 3 * n_games
    assert c['simulations'] == c['evaluations'] + c['terminal_leaves']
    states, results = eng.game_states()
    fens = [rc.state_to_fen(s) for s in states]
    assert all(f.split()[1] == 'b' and f.split()[3] == '2' for f in fens)   # three plies played from the start


def test_device_rng_is_deterministic_and_seeded(net):
    def run(seed):
        eng = make(net, 64, 12, seed=seed)
        for _ in range(4):
            eng.search(12)
            eng.play_device()
        return eng.game_states()[0], eng.root_stats(want_q=False)[1]
    a, b, c = run(11), run(11), run(12)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert not np.array_equal(a[0], c[0])


def test_root_noise_changes_the_search(net):
    quiet = make(net, 32, 40, seed=1, dirichlet_epsilon=0.0)
    noisy = make(net, 32, 40, seed=1)
    quiet.search(40); noisy.search(40)
    vq, vn = quiet.root_stats(want_q=False)[1], noisy.root_stats(want_q=False)[1]
    assert (vq == vq[0]).all()                          # no noise: all games search the start position identically
    assert not (vn == vn[0]).all()                      # per-game Dirichlet draws: they differ


def test_network_batches_larger_than_one_chunk(net):
    eng = make(net, 16, 4)
    pos = rc.random_positions(21, 40000)
    pos = np.ascontiguousarray(np.tile(pos, 3)[:9000])
    tokens, clocks = rc.tokenize(pos)
    logits, values = eng.network_forward(tokens, clocks)          # 8192 + 808 boards
    l2, v2 = eng.network_forward(tokens[8100:8400], clocks[8100:8400])
    assert np.array_equal(logits[8100:8400], l2) and np.array_equal(values[8100:8400], v2)
    assert np.isfinite(logits).all() and (np.abs(values) <= 1).all()
    e0, e1 = eng.network_forward(tokens[:0], clocks[:0])
    assert e0.shape == (0, 554) and e1.shape == (0,)


def test_replay_rewards_and_restarts(net):
    from minitchess_alphazero_b200.selfplay import BatchedSelfPlay, replay_to_episode_dicts
    sp = BatchedSelfPlay(net, n_games=128, num_simulations=4, seed=2)
    sp.run(66)
    c = sp.engine.counters()
    assert c['games_finished'] >= 128
    tuples = sp.drain()
    eps = replay_to_episode_dicts(tuples[:600])
    starts = [i for i, e in enumerate(eps) if e['observation'] == '2nbk/2ppp/5/5/PPP2/KBN2 w 0 1']
    assert len(starts) >= 2
    for a, b in zip(starts[:-1], starts[1:]):
        game = eps[a:b]
        assert all(game[k]['reward'] == -game[k + 1]['reward'] for k in range(len(game) - 1))
        # the game line is consistent: each observation follows from the previous one by its action
        states = rc.fens_to_states([g['observation'] for g in game])
        nxt, st = rc.apply(states[:-1], np.array([g['action'] for g in game[:-1]], dtype=np.uint16))
        assert (st == 0).all() and [rc.state_to_fen(s) for s in nxt] == [g['observation'] for g in game[1:]]
        # ... and ends where the rules say it ends: the last move leads to a finished position whose result gives the
        # last mover's reward (a game may also end by fivefold repetition, which a single position does not show)
        last, st = rc.apply(states[-1:], np.array([game[-1]['action']], dtype=np.uint16))
        assert st[0] == 0
        result = int(rc.legal_moves(last)[2][0])
        assert result != 0 or len(game) >= 9, rc.state_to_fen(last[0])
        if result != 0:
            assert game[-1]['reward'] == (0.0 if result == 3 else 1.0)
    assert len(sp.drain()) == 0                         # drained
