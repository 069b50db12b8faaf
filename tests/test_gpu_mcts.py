"""GPU: the same tree-logic parity checks as tests/test_mcts_host.py, but through libmcaz.so's
CUDA kernels (select_expand_kernel / backup_kernel / play_kernel) and the C ABI."""
import numpy as np
import pytest

import parity_common as pc
from oracle import ref_selfplay as rs
from oracle import rules_c as rc
from oracle.hash_eval import hash_evaluate
from test_mcts_host import (test_golden_game_visit_counts_bit_exact, test_batched_games_match_restatement,  # noqa: F401
                            test_no_noise_and_numpy1_flow_switch, test_terminal_revisit_sign_flip,
                            test_capacity_overflow_fails_loudly, test_phase_errors, test_fivefold_repetition_inside_a_simulation,
                            make_engine)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def backend(mcaz_lib):
    return None          # Engine(_backend=None) -> the real library


def test_many_concurrent_games(backend):
    """64 concurrent games (warps across several CTAs) against 64 independent reference games."""
    G, sims = 64, 16
    eng = make_engine(backend, G, sims)
    rngs = [np.random.RandomState(1000 + g) for g in range(G)]
    records, states, results = pc.play_games(eng, hash_evaluate, sims, rngs, max_plies=14)
    for g in range(0, G, 5):
        ref_records, ep, trees = rs.play_game(hash_evaluate, sims, rng=np.random.RandomState(1000 + g), max_plies=14)
        assert [r['action'] for r in records[g]] == [r['action'] for r in ref_records]
        assert [r['N'] for r in records[g]] == [(np.array(r['pi']) * sum(np.array(records[g][i]['N']))).round().tolist()
                                                for i, r in enumerate(ref_records)]
        for t in (0, 1):
            pc.compare_with_tree(eng, g, t, trees[t])
