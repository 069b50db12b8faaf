"""Re-check, from the reference's own UNMODIFIED code, what tests/golden/restatement_pin.json records: the reference
(exp/agent.py + exp/policy.py + exp/environment.py on the oracle shims) and oracle/ref_selfplay.py play the same games --
same visit counts, Q, moves and whole-tree digests -- and both equal the committed fixture.  Build container only
(needs /root/reference); run by tests/test_oracle_pinned.py in a subprocess so that its sys.path changes stay here."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference through oracle/ref_runner.py)


def main():
    bad = []
    for name, model, evaluate, sims, seed in (('mcts_hash_game.json', mg.HashModel(), mg.hash_evaluate, 36, 0),
                                              ('mcts_hash_game_s1.json', mg.HashModel(), mg.hash_evaluate, 50, 1)):
        rows, digests, sizes = mg.reference_game(model, sims, seed)
        r2, d2 = mg.restated_game(evaluate, sims, seed)
        gold = json.load(open(os.path.join(HERE, name)))
        same_ref = len(r2) == len(rows) and d2 == digests and all(
            a['action'] == b['action'] and a['pi'] == b['pi'] and a['observation'] == b['observation'] for a, b in zip(r2, rows))
        same_gold = gold['tree_sha256'] == digests and gold['tree_sizes'] == sizes and json.loads(json.dumps(rows)) == gold['plies']
        print(name, 'reference == restatement:', same_ref, '| reference == committed fixture:', same_gold)
        if not (same_ref and same_gold):
            bad.append(name)
    # the action table the reference ships is the one the restatement's formula gives
    import hashlib
    blob = open(os.path.join(mg.REFERENCE_ROOT, 'exp', 'moves_dict.json'), 'rb').read()
    want = json.load(open(os.path.join(HERE, 'moves_dict.json.sha256')))
    if hashlib.sha256(blob).hexdigest() != want['sha256']:
        bad.append('moves_dict.json')
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
