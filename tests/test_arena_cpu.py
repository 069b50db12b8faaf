"""CPU (host build of csrc/mcts_core.cuh): the arena's side / turn bookkeeping and the 0.55 gate -- the reference's
WinnerRecorder (exp/callbacks.py:7-28: a decisive game is won by the side that moved last) and the commented-out gate
(exp/learner.py:97-145, app/base.py:195-196) -- on positions whose winner is known: every legal move of the side to move
mates (found by the oracle), so the agent that owns that colour wins whatever it evaluates."""
import numpy as np
import pytest

from oracle import rules_c as rc
import parity_common as pc

WHITE_MATES = ['2nQ1/1Q1p1/pk3/1qBqK/1q3/1qQ2 w 6 4', '1Q3/pR3/k1PbR/R4/K2q1/1Br2 w 12 23', '5/kQ2K/qQp2/5/4q/3nq w 15 20',
               '1qB2/2K1Q/2Q2/p2B1/1k3/3N1 w 15 22']
BLACK_MATES = ['1r1r1/1qr1r/1Qp2/Q1PrK/1kP2/5 b 4 15', 'kRqqq/4K/2N2/1q1PB/3B1/5 b 8 9', '1n2b/4p/4r/1K3/qB1p1/1nk1N b 2 18',
               'n2Kr/5/1n3/1k3/N1Np1/N1N1r b 0 10']


def uniform(sign):
    """A 'network': uniform priors, value = sign * material balance of the mover (any deterministic function will do)."""
    worth = np.array([0, 1, 5, 3, 3, 9, 0], dtype=np.float32)

    def f(tokens, clocks, leaf_states):
        t = tokens.astype(np.int64)
        v = np.clip(sign * (worth[t[:, :30]].sum(1) - worth[t[:, 30:]].sum(1)) / 10.0, -1, 1).astype(np.float32)
        return np.zeros((len(t), 554), dtype=np.float32), v
    return f


def test_forced_mates_are_what_the_oracle_says():
    for fens, win in ((WHITE_MATES, 1), (BLACK_MATES, 2)):
        st = rc.fens_to_states(fens)
        codes, counts, res = rc.legal_moves(st)
        assert (res == 0).all() and (counts > 0).all()
        for i in range(len(st)):
            for k in range(counts[i]):
                out, status = rc.apply(st[i:i + 1], codes[i, k:k + 1].copy())
                assert status[0] == 0 and rc.legal_moves(out)[2][0] == win


@pytest.mark.parametrize('second_half,want', [(BLACK_MATES, (8, 0)), (WHITE_MATES, (4, 4))])
def test_arena_counts_wins_by_side_and_applies_the_gate(second_half, want):
    from minitchess_alphazero_b200.arena import Arena, GATE_THRESHOLD, gate_result
    # games 0-3: A owns white and white mates; games 4-7: A owns black -- black mates (A wins) or white mates (B wins)
    start = rc.fens_to_states(WHITE_MATES + second_half)
    for a_eval, b_eval in ((uniform(+1), uniform(-1)), (uniform(-1), uniform(+1))):     # the winner is the colour, not the evaluator
        arena = Arena(games_per_side=4, num_simulations=8, epsilon=0.0, seed=3, evaluators=(a_eval, b_eval), _backend=pc.host_backend())
        out = arena.play(start_states=start)
        assert (out['a'], out['b']) == want and out['draws'] == 0 and out['plies'] == 1
        assert out['a_as_white'] == (4, 0)
        assert out['a_as_black'] == ((4, 0) if second_half is BLACK_MATES else (0, 4))
        assert out['gate'] == gate_result(*want)
        accepted = out['gate'] > GATE_THRESHOLD                     # app/base.py:195-196
        assert accepted == (second_half is BLACK_MATES)
    assert abs(gate_result(3, 1) - 0.75) < 1e-6 and gate_result(0, 0) == 0.0       # exp/learner.py:145 (draws do not count)


def test_arena_draws_do_not_count():
    """From the start position with these evaluators every game runs into the 30-move cap: no decisive game, gate 0."""
    from minitchess_alphazero_b200.arena import Arena
    arena = Arena(games_per_side=3, num_simulations=8, epsilon=0.0, seed=1, evaluators=(uniform(+1), uniform(-1)), _backend=pc.host_backend())
    out = arena.play()
    assert out['a'] + out['b'] + out['draws'] == 6 and out['plies'] <= 60
    assert out['gate'] == (out['a'] / (out['a'] + out['b'] + 1e-8))
