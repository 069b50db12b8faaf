"""The pin of oracle/ref_selfplay.py (the restatement that travels to the GPU box) to the reference's own code.

Everywhere: the restatement reproduces the committed golden games -- fixtures written by tests/golden/make_golden.py from
the reference's UNMODIFIED exp/agent.py / exp/policy.py / exp/environment.py -- ply by ply (visit counts, Q bits, moves)
and as whole-tree digests.  Where /root/reference exists (the build container) the reference itself is run again and
compared with both (tests/golden/check_pin.py, in a subprocess)."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REPO, load_golden
from oracle import ref_selfplay as rs
from oracle.hash_eval import hash_evaluate
from oracle.ref_runner import reference_available


def tree_digest(tree):
    h = hashlib.sha256()
    for k in sorted(tree.N):
        h.update(k.encode()); h.update(np.asarray(tree.N[k], dtype=np.float64).tobytes())
        h.update(np.asarray(tree.Q[k], dtype=np.float64).tobytes())
    for k in sorted(tree.terminal):
        h.update(k.encode()); h.update(np.float64(tree.terminal[k]).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize('name', ['mcts_hash_game.json', 'mcts_hash_game_s1.json'])
def test_restatement_reproduces_the_reference_games(name):
    g = load_golden(name)
    np.random.seed(g['seed'])
    records, ep, trees = rs.play_game(hash_evaluate, g['sims'])
    assert len(records) == len(g['plies'])
    for mine, ref in zip(records, g['plies']):
        assert mine['observation'] == ref['observation'] and mine['legal_moves'] == ref['legal_moves']
        assert mine['pi'] == ref['pi'] and mine['action'] == ref['action']
    assert ep.fen == g['plies'][-1]['next']
    assert [tree_digest(t) for t in trees] == g['tree_sha256']
    assert [[len(t.N), len(t.terminal)] for t in trees] == g['tree_sizes']


def test_pin_record_is_committed():
    pin = load_golden('restatement_pin.json')
    assert all(v['restatement_identical'] for v in pin.values()) and len(pin) == 3


@pytest.mark.skipif(not reference_available(), reason='the reference tree (/root/reference) only exists in the build container')
def test_reference_itself_still_agrees():
    out = subprocess.run([sys.executable, os.path.join(GOLDEN, 'check_pin.py')], cwd=REPO, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
