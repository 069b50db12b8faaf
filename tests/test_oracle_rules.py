"""CPU: the oracle's rules against the committed golden vectors (made by the reference's own
exp/environment.py on the shim), the C mailbox against the Python mailbox, and the host build
of the product's bitboard header against both."""
import ctypes

import numpy as np
import pytest

from conftest import load_golden, vp
from parity_common import synthetic_positions
from oracle import rules_c as rc

RES = {'*': 0, '1-0': 1, '0-1': 2, '1/2-1/2': 3}


@pytest.fixture(scope='module')
def golden_positions():
    return load_golden('rules_positions.json.gz')


def test_moves_table_matches_reference_json():
    """The 554-code table regenerates the reference's moves_dict.json byte for byte."""
    import hashlib
    import json
    from oracle.shims import chess as mchess
    d = {'w': {}, 'b': {}}
    for side, key in ((1, 'w'), (0, 'b')):
        inv = {}
        for f in range(30):
            for t in range(30):
                c = rc.code_of(f, t, side)
                if c >= 0:
                    inv[c] = mchess.square_name(f) + mchess.square_name(t)
        d[key] = {inv[c]: c for c in range(554)}
    blob = json.dumps(d).encode()
    meta = load_golden('moves_dict.json.sha256')
    assert len(blob) == meta['bytes']
    assert hashlib.sha256(blob).hexdigest() == meta['sha256']


def test_c_oracle_matches_golden(golden_positions):
    fens = [r['fen'] for r in golden_positions]
    states = rc.fens_to_states(fens)
    assert [rc.state_to_fen(s) for s in states] == fens
    codes, counts, results = rc.legal_moves(states)
    for i, r in enumerate(golden_positions):
        assert list(codes[i, :counts[i]]) == r['legal'], r['fen']
        assert results[i] == RES[r['result']], r['fen']
        assert (results[i] != 0) == r['done']
        if not r['done']:
            out, st = rc.apply(np.repeat(states[i:i + 1], counts[i]), codes[i, :counts[i]])
            assert (st == 0).all()
            assert [rc.state_to_fen(o) for o in out] == [c[0] for c in r['children']], r['fen']


def test_c_oracle_perft_matches_golden():
    g = load_golden('perft.json')
    s = rc.fen_to_state(g['fen'])
    assert [rc.perft(s, d) for d in range(1, 6)] == g['nodes']
    assert g['nodes'][0] == 6   # a2a3 b2b3 c2c3 c1b3 c1d3 c1e2


def test_python_shim_matches_c_oracle():
    from oracle.shims import chess as mchess
    from oracle.ref_selfplay import RefEpisode
    pos = rc.random_positions(11, 3000)
    codes, counts, results = rc.legal_moves(pos)
    for i in range(0, len(pos), 3):
        ep = RefEpisode(rc.state_to_fen(pos[i]))
        assert ep.legal == list(codes[i, :counts[i]])
        assert RES[ep.board.result()] == results[i]


def test_edge_cases():
    # finished position: stepping is refused (exp/environment.py:69-70) and the position is unchanged
    mate = rc.fen_to_state('k4/1Q3/2K2/5/5/5 b 3 20')
    out, st = rc.apply(mate, 0)
    assert st[0] == 2 and out[0] == mate
    # illegal code
    start = rc.start_state()
    out, st = rc.apply(start, 0)
    assert st[0] == 1 and out[0] == start
    # promotion always queens (exp/environment.py:72-74)
    promo = rc.fen_to_state('1k3/4P/5/5/5/K4 w 0 10')
    codes, counts, _ = rc.legal_moves(promo)
    e5e6 = rc.code_of(4 + 5 * 4, 4 + 5 * 5, 1)
    assert e5e6 in codes[0, :counts[0]]
    out, st = rc.apply(promo, e5e6)
    assert st[0] == 0 and rc.state_to_fen(out[0]) == '1k2Q/5/5/5/5/K4 b 0 10'
    # 30-move cap boundary
    assert rc.legal_moves(rc.fen_to_state('2nbk/2ppp/5/5/PPP2/KBN2 b 4 30'))[2][0] == 0
    assert rc.legal_moves(rc.fen_to_state('2nbk/2ppp/5/5/PPP2/KBN2 w 4 31'))[2][0] == 3
    # empty input
    c, n, r = rc.legal_moves(np.zeros(0, dtype=rc.STATE_DTYPE))
    assert c.shape == (0, rc.MAX_MOVES) and len(n) == 0


def test_rule_switches():
    r = rc.Rules.default()
    r.pawn_double_step = 1
    codes, counts, _ = rc.legal_moves(rc.start_state(), r)
    assert counts[0] == 9          # a2a4 b2b4 c2c4 added
    r = rc.Rules.default()
    r.promo_multiplicity = 4
    promo = rc.fen_to_state('1k3/4P/5/5/5/K4 w 0 10')
    codes, counts, _ = rc.legal_moves(promo, r)
    e5e6 = rc.code_of(24, 29, 1)
    assert list(codes[0, :counts[0]]).count(e5e6) == 4


def test_bitboard_header_matches_oracle(host_rules):
    """csrc/minitchess.cuh compiled for the host == mailbox oracle on ~100k reachable positions."""
    pos = rc.random_positions(5, 120000)
    c0, n0, r0 = rc.legal_moves(pos)
    c1, n1, r1 = np.zeros_like(c0), np.zeros_like(n0), np.zeros_like(r0)
    host_rules.hh_legal_moves(vp(pos), len(pos), None, vp(c1), vp(n1), vp(r1))
    assert np.array_equal(n0, n1) and np.array_equal(r0, r1) and np.array_equal(c0, c1)
    idx = np.repeat(np.arange(len(pos)), n0)
    codes = np.ascontiguousarray(c0[np.arange(c0.shape[1])[None, :] < n0[:, None]])
    src = np.ascontiguousarray(pos[idx])
    o0, s0 = rc.apply(src, codes)
    o1, s1 = np.zeros_like(o0), np.zeros_like(s0)
    host_rules.hh_apply(vp(src), vp(codes), len(codes), None, vp(o1), vp(s1))
    assert np.array_equal(o0, o1) and np.array_equal(s0, s1)
    rnd = np.random.RandomState(0).randint(0, 554, len(pos)).astype(np.uint16)
    o0, s0 = rc.apply(pos, rnd)
    o1, s1 = np.zeros_like(o0), np.zeros_like(s0)
    host_rules.hh_apply(vp(pos), vp(rnd), len(pos), None, vp(o1), vp(s1))
    assert np.array_equal(o0, o1) and np.array_equal(s0, s1)
    for rules in (rc.Rules(1, 1, 30, 1, 1), rc.Rules(0, 4, 30, 1, 1), rc.Rules(0, 1, 12, 0, 1)):
        import ctypes
        c0, n0, r0 = rc.legal_moves(pos[:20000], rules)
        c1, n1, r1 = np.zeros_like(c0), np.zeros_like(n0), np.zeros_like(r0)
        host_rules.hh_legal_moves(vp(np.ascontiguousarray(pos[:20000])), 20000, ctypes.byref(rules), vp(c1), vp(n1), vp(r1))
        assert np.array_equal(n0, n1) and np.array_equal(r0, r1) and np.array_equal(c0, c1)


def test_guard_matches_move_by_move_king_safety(host_rules):
    """mc::legal_targets (pins, check mask and king danger worked out once per position) == one king-safety test per
    candidate move, for every piece of 100 k reachable and 100 k synthetic positions, under both pawn rules."""
    import ctypes
    reach = np.ascontiguousarray(rc.random_positions(9, 100000))
    synth = synthetic_positions(3, 100000)
    for pos in (reach, synth):
        for rules in (None, rc.Rules(1, 1, 30, 1, 1)):
            a = np.zeros((len(pos), 30), dtype=np.uint32)
            b = np.zeros_like(a)
            host_rules.hh_targets_both(vp(pos), len(pos), ctypes.byref(rules) if rules is not None else None, vp(a), vp(b))
            bad = np.nonzero((a != b).any(1))[0]
            assert len(bad) == 0, rc.state_to_fen(pos[bad[0]])
            assert b.any()
    # and the full generator and the step against the mailbox oracle on the synthetic set (checks, pins and mates
    # galore), under the default rules and every switch
    for rules in (None, rc.Rules(1, 1, 30, 1, 1), rc.Rules(0, 4, 30, 1, 1), rc.Rules(0, 1, 12, 0, 1)):
        rp = ctypes.byref(rules) if rules is not None else None
        c0, n0, r0 = rc.legal_moves(synth, rules)
        c1, n1, r1 = np.zeros_like(c0), np.zeros_like(n0), np.zeros_like(r0)
        host_rules.hh_legal_moves(vp(synth), len(synth), rp, vp(c1), vp(n1), vp(r1))
        assert np.array_equal(n0, n1) and np.array_equal(r0, r1) and np.array_equal(c0, c1)
        rnd = np.random.RandomState(1).randint(0, 554, len(synth)).astype(np.uint16)
        first = np.ascontiguousarray(c0[:, 0])
        for codes in (first, rnd):
            o0, s0 = rc.apply(synth, codes, rules)
            o1, s1 = np.zeros_like(o0), np.zeros_like(s0)
            host_rules.hh_apply(vp(synth), vp(codes), len(synth), rp, vp(o1), vp(s1))
            assert np.array_equal(o0, o1) and np.array_equal(s0, s1)


def test_tokeniser_matches_golden(host_rules):
    rows = load_golden('tokens.json')
    states = rc.fens_to_states([r['fen'] for r in rows])
    t0, k0 = rc.tokenize(states)
    t1, k1 = np.zeros_like(t0), np.zeros_like(k0)
    host_rules.hh_tokenize(vp(states), len(states), vp(t1), vp(k1))
    for i, r in enumerate(rows):
        assert t0[i].tolist() == r['tokens'] and t1[i].tolist() == r['tokens'], r['fen']
        assert k0[i].tobytes().hex() == r['clock_f32_hex'] and k1[i].tobytes().hex() == r['clock_f32_hex']


def test_search_tables_match_the_arithmetic(host_rules):
    """The two tables the one-warp-per-tree search reads on the device (csrc/minitchess.cuh: CODE_VIEW, MOVE_ORDER) are constexpr
    products of the arithmetic definitions; the same objects built for the host must reproduce code_to_view for all 554 codes and
    the walk of emit_square_codes -- code by code, place by place -- for random target sets, both blocks, both colours and
    repeated promotion codes (promo_multiplicity 4)."""
    fv, tv = ctypes.c_int(), ctypes.c_int()
    for code in range(554):
        assert host_rules.hh_code_to_view(code, ctypes.byref(fv), ctypes.byref(tv)) == 0
        assert host_rules.hh_code_view_table(code) == fv.value | (tv.value << 8)
        assert host_rules.hh_view_to_code(fv.value, tv.value) == code
    rng = np.random.RandomState(11)
    walk = np.zeros(256, dtype=np.uint16)
    table = np.zeros(256, dtype=np.uint16)
    n_table = ctypes.c_int()
    checked = 0
    for square in range(30):
        for knight in (0, 1):
            for white in (0, 1):
                # every target a piece on this view square can have in its block (view squares), as a real-square set
                reach = 0
                for t in range(30):
                    c = host_rules.hh_view_to_code(square, t)
                    if c >= 0 and (c >= 430) == bool(knight):
                        reach |= 1 << (t if white else 29 - t)
                for trial in range(24):
                    tg = reach & int(rng.randint(0, 1 << 30)) if trial else reach
                    for promo_piece, rep in ((0, 1), (1, 1), (1, 4)):
                        if knight and promo_piece:
                            continue
                        walk[:] = 0xffff; table[:] = 0xfffe
                        n = host_rules.hh_emit_both(square, white, knight, ctypes.c_uint32(tg), promo_piece, rep, vp(walk), vp(table), ctypes.byref(n_table))
                        assert n == n_table.value, (square, knight, white, hex(tg), promo_piece, rep)
                        assert np.array_equal(walk[:n], table[:n]), (square, knight, white, hex(tg), promo_piece, rep)
                        checked += n
    assert checked > 30000
