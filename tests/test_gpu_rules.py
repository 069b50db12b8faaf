"""GPU: the CUDA rules kernels through the C ABI, bit-exact against the CPU oracle and the
golden vectors (BASELINE.json config 2: move-gen validation on 1M random reachable positions)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import rules_c as rc

pytestmark = pytest.mark.gpu
RES = {'*': 0, '1-0': 1, '0-1': 2, '1/2-1/2': 3}


@pytest.fixture(scope='module')
def rules(mcaz_lib):
    from minitchess_alphazero_b200 import rules
    return rules


def test_golden_positions(rules):
    recs = load_golden('rules_positions.json.gz')
    states = rules.states_from_fens([r['fen'] for r in recs])
    codes, counts, results = rules.legal_moves(states)
    for i, r in enumerate(recs):
        assert list(codes[i, :counts[i]]) == r['legal'], r['fen']
        assert results[i] == RES[r['result']], r['fen']
    for i, r in enumerate(recs):
        if r['done']:
            continue
        out, st = rules.apply(np.repeat(states[i:i + 1], counts[i]), codes[i, :counts[i]])
        assert (st == 0).all()
        assert [rules.state_to_fen(o) for o in out] == [c[0] for c in r['children']]


def test_perft(rules):
    g = load_golden('perft.json')
    s = rules.state_from_fen(g['fen'])
    for d, want in enumerate(g['nodes'], start=1):
        assert int(rules.perft(s, d)[0]) == want
    assert int(rules.perft(s, 0)[0]) == 1
    for d in (6, 7):
        assert int(rules.perft(s, d)[0]) == rc.perft(rc.fen_to_state(g['fen']), d)


def test_one_million_positions(rules):
    """Config 2 at full size: sorted legal code lists, results and every successor."""
    pos = rc.random_positions(0, 1600000)
    extra_seed = 1
    while len(pos) < (1 << 20):
        pos = np.unique(np.concatenate([pos, rc.random_positions(extra_seed, 800000)]))
        extra_seed += 1
    pos = np.ascontiguousarray(pos[:1 << 20])
    c0, n0, r0 = rc.legal_moves(pos)
    c1, n1, r1 = rules.legal_moves(pos)
    assert np.array_equal(n0, n1)
    assert np.array_equal(r0, r1)
    assert np.array_equal(c0, c1)
    mask = (np.arange(c0.shape[1])[None, :] < n0[:, None]) & (r0 == 0)[:, None]
    idx = np.repeat(np.arange(len(pos)), mask.sum(1))
    codes = np.ascontiguousarray(c0[mask])
    src = np.ascontiguousarray(pos[idx])
    o0, s0 = rc.apply(src, codes)
    o1, s1 = rules.apply(src, codes)
    assert np.array_equal(s0, s1) and (s1 == 0).all()
    assert np.array_equal(o0, o1)
    # perft(2) from 10k samples
    sample = np.ascontiguousarray(pos[:: len(pos) // 10000][:10000])
    want = np.array([rc.perft(s, 2) for s in sample], dtype=np.uint64)
    assert np.array_equal(rules.perft(sample, 2), want)


def test_illegal_and_finished(rules):
    pos = rc.random_positions(3, 200000)
    rnd = np.random.RandomState(0).randint(0, 554, len(pos)).astype(np.uint16)
    o0, s0 = rc.apply(pos, rnd)
    o1, s1 = rules.apply(pos, rnd)
    assert np.array_equal(s0, s1) and np.array_equal(o0, o1)
    assert set(np.unique(s1)) == {0, 1, 2}


def test_rule_switches(rules):
    from minitchess_alphazero_b200 import _lib
    pos = np.ascontiguousarray(rc.random_positions(9, 60000))
    for vals in ((1, 1, 30, 1, 1), (0, 4, 30, 1, 1), (0, 1, 12, 0, 1)):
        c0, n0, r0 = rc.legal_moves(pos, rc.Rules(*vals))
        c1, n1, r1 = rules.legal_moves(pos, _lib.Rules(*vals))
        assert np.array_equal(n0, n1) and np.array_equal(r0, r1) and np.array_equal(c0, c1)


def test_tokeniser(rules):
    rows = load_golden('tokens.json')
    states = rules.states_from_fens([r['fen'] for r in rows])
    tokens, clocks = rules.tokenize(states)
    for i, r in enumerate(rows):
        assert tokens[i].tolist() == r['tokens']
        assert clocks[i].tobytes().hex() == r['clock_f32_hex']
    pos = np.ascontiguousarray(rc.random_positions(2, 50000))
    t0, k0 = rc.tokenize(pos)
    t1, k1 = rules.tokenize(pos)
    assert np.array_equal(t0, t1) and np.array_equal(k0, k1)


def test_empty_and_device_buffers(rules):
    import torch
    from minitchess_alphazero_b200 import _lib
    c, n, r = rules.legal_moves(np.zeros(0, dtype=_lib.STATE_DTYPE))
    assert len(n) == 0
    pos = np.ascontiguousarray(rc.random_positions(4, 5000))
    d_states = torch.from_numpy(pos.view(np.int32).reshape(-1, 5)).cuda()
    d_codes = torch.zeros(len(pos), _lib.MC_MAX_MOVES, dtype=torch.int16, device='cuda')
    d_counts = torch.zeros(len(pos), dtype=torch.int32, device='cuda')
    d_res = torch.zeros(len(pos), dtype=torch.int8, device='cuda')
    _lib.check(_lib.lib().mc_legal_moves(_lib.ptr(d_states), len(pos), None, _lib.ptr(d_codes), _lib.ptr(d_counts), _lib.ptr(d_res)))
    c0, n0, r0 = rc.legal_moves(pos)
    assert np.array_equal(d_counts.cpu().numpy(), n0)
    assert np.array_equal(d_codes.cpu().numpy().view(np.uint16), c0)


def test_synthetic_positions(rules):
    """Random piece placements (dense in pins, double checks, mates): the CUDA kernels against the mailbox oracle."""
    from parity_common import synthetic_positions
    pos = synthetic_positions(3, 50000)
    c0, n0, r0 = rc.legal_moves(pos)
    c1, n1, r1 = rules.legal_moves(pos)
    assert np.array_equal(n0, n1) and np.array_equal(r0, r1) and np.array_equal(c0, c1)
    rnd = np.random.RandomState(1).randint(0, 554, len(pos)).astype(np.uint16)
    for codes in (np.ascontiguousarray(c0[:, 0]), rnd):
        o0, s0 = rc.apply(pos, codes)
        o1, s1 = rules.apply(pos, codes)
        assert np.array_equal(o0, o1) and np.array_equal(s0, s1)
