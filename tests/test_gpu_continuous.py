"""GPU: the three levers of SURVEY.md 7.3 point 6 that raise simulations/s above evaluations/s without changing
a single visit count -- simulations that end on terminal positions complete inside the search launch
(`free_sims`), the exact evaluation cache (`eval_cache_log2`), and continuous self-play (`az_selfplay`, every
game moves as soon as its own simulations are done).  Each is checked bit-for-bit against the plain lock-step
search: the order of a game's simulations, which is all the reference's sequential MCTS depends on
(exp/agent.py:41-45), must not change."""
import numpy as np
import pytest
import torch

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


def snapshot(eng):
    codes, visits, q, n_legal = eng.root_stats()
    states, results = eng.game_states()
    return codes, visits, q, n_legal, states, results


def same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


def test_cache_and_free_sims_leave_every_visit_count_unchanged(net):
    G, sims = 300, 24
    plain = make(net, G, sims, seed=5, free_sims=1)
    free = make(net, G, sims, seed=5, free_sims=6)
    cached = make(net, G, sims, seed=5, free_sims=4, eval_cache_log2=18)
    tiny = make(net, G, sims, seed=5, eval_cache_log2=6)            # 64 entries: constant replacement
    engines = (plain, free, cached, tiny)
    for move in range(12):
        for e in engines:
            e.search(sims)
        ref = snapshot(plain)
        for e in engines[1:]:
            assert same(ref, snapshot(e)), move                       # N, Q (float64 bits), codes, positions
        for e in engines:
            e.play_device()
    counters = [e.counters() for e in engines]
    for c in counters:
        assert c['simulations'] == 12 * sims * G
        assert c['simulations'] == c['evaluations'] + c['terminal_leaves'] + c['cached_evaluations']
        assert c['nodes'] == counters[0]['nodes'] and c['edges'] == counters[0]['edges']
    assert counters[0]['cached_evaluations'] == 0 and counters[1]['cached_evaluations'] == 0
    assert counters[2]['cached_evaluations'] > 0.2 * counters[2]['simulations']   # 300 games share the opening
    assert 0 < counters[3]['cached_evaluations'] < counters[2]['cached_evaluations']
    assert counters[2]['evaluations'] < counters[0]['evaluations']


def test_late_game_terminal_leaves_need_no_network_rows(net):
    """Near the 30-move cap most leaves are finished positions: they are backed up inside the launch."""
    from oracle import rules_c as rc
    G, sims = 64, 40
    pos = rc.random_positions(17, 60000)
    late = np.ascontiguousarray(pos[((pos['meta'] >> 16) & 0xff) >= 30][:G])
    assert len(late) == G
    got = []
    for free in (1, 8):
        eng = make(net, G, sims, seed=3, free_sims=free)
        eng.reset_games(states=late)
        eng.search(sims)
        got.append((snapshot(eng), eng.counters()))
    assert same(got[0][0], got[1][0])
    c = got[1][1]
    assert c['terminal_leaves'] > c['evaluations']
    assert c['simulations'] == c['evaluations'] + c['terminal_leaves'] == got[0][1]['simulations']


def test_cache_is_dropped_when_the_weights_change(net):
    from minitchess_alphazero_b200.policy import Network, flatten_state_dict
    G, sims = 128, 16
    eng = make(net, G, sims, seed=9, eval_cache_log2=16)
    eng.search(sims)
    assert eng.counters()['cached_evaluations'] > 0
    torch.manual_seed(123)
    other = Network().eval()
    flat = flatten_state_dict(other.state_dict(), device='cuda')
    eng.set_weights(flat)
    eng.reset_games()
    before = eng.counters()['cached_evaluations']
    eng.search(1)                                                   # the root of every game: all must be evaluated afresh
    assert eng.counters()['cached_evaluations'] == before
    eng.search(sims - 1)
    fresh = make(other, G, sims, seed=9)                            # no cache; same per-game RNG position as `eng`
    fresh.search(sims)
    fresh.reset_games()
    fresh.search(sims)
    assert same(snapshot(eng), snapshot(fresh))


def finished_games(tuples):
    """Replay stream -> list of games, each the raw bytes of its tuples (a game starts at the start position)."""
    from minitchess_alphazero_b200 import rules
    start = rules.state_from_fen(rules.STARTING_FEN)
    is_start = np.array([t['observation'] == start for t in tuples])
    cuts = list(np.nonzero(is_start)[0]) + [len(tuples)]
    return [tuples[a:b].tobytes() for a, b in zip(cuts[:-1], cuts[1:])]


def test_continuous_selfplay_plays_the_same_games_as_lockstep(net):
    from collections import Counter
    G, sims = 96, 6
    lock = make(net, G, sims, seed=21, eval_cache_log2=14)
    for _ in range(62):                                            # every slot finishes its first game (<= 60 plies)
        lock.search(sims)
        lock.play_device()
    want = Counter(finished_games(lock.drain_replay()))
    assert sum(want.values()) >= G

    cont = make(net, G, sims, seed=21, eval_cache_log2=14, free_sims=3)
    got = Counter()
    for _ in range(40):
        cont.selfplay(64, sims)
        got.update(finished_games(cont.drain_replay()))
        if all(got[k] >= n for k, n in want.items()):
            break
    missing = [k for k, n in want.items() if got[k] < n]
    assert not missing, '%d of %d lock-step games were not reproduced by az_selfplay' % (len(missing), len(want))
    c = cont.counters()
    assert c['simulations'] == c['evaluations'] + c['terminal_leaves'] + c['cached_evaluations']
    assert c['moves'] * sims <= c['simulations'] + G * sims           # every move was preceded by its simulations


def test_deferred_rows_change_nothing(net):
    """az_config.defer_rows (az_selfplay): the network pass only runs whole tile pairs (256 rows); a short last pair's leaves
    wait for the next batch.  Their games lose a launch, nothing else: continuous self-play plays the same games as the plain
    lock-step search, move for move and visit count for visit count (the replay tuples carry pi), and az_search never defers."""
    from collections import Counter
    G, sims = 600, 6                                                 # batches of 300 .. 600 rows: two or three tile pairs
    lock = make(net, G, sims, seed=21, eval_cache_log2=14, defer_rows=255)
    for _ in range(62):                                              # every slot finishes its first game (<= 60 plies)
        lock.search(sims)
        lock.play_device()
    assert lock.counters()['trimmed_batches'] == 0                   # lock-step: a game has no launch to spare
    want = Counter(finished_games(lock.drain_replay()))
    assert sum(want.values()) >= G
    for thr in (255, 160):
        cont = make(net, G, sims, seed=21, eval_cache_log2=14, free_sims=3, recycle=1, defer_rows=thr)
        got = Counter()
        for _ in range(40):
            cont.selfplay(64, sims)
            got.update(finished_games(cont.drain_replay()))
            if all(got[k] >= n for k, n in want.items()):
                break
        missing = [k for k, n in want.items() if got[k] < n]
        assert not missing, '%d of %d lock-step games were not reproduced (defer_rows %d)' % (len(missing), len(want), thr)
        c = cont.counters()
        assert c['trimmed_batches'] > 0 and c['deferred_rows'] >= c['trimmed_batches']
        assert c['simulations'] == c['evaluations'] + c['terminal_leaves'] + c['cached_evaluations']
        assert c['moves'] * sims <= c['simulations'] + G * sims


def test_selfplay_argument_checks(net):
    from minitchess_alphazero_b200._lib import McazError
    eng = make(net, 8, 8)
    with pytest.raises(McazError):
        eng.selfplay(4, 9)                                          # more simulations than the arenas were sized for
    with pytest.raises(McazError):
        eng.selfplay(-1, 4)
    eng.selfplay(0, 4)                                              # nothing to do
    assert eng.counters()['simulations'] == 0


def test_recycling_unreachable_plies_changes_nothing(net):
    """SURVEY.md 7.3 point 7: nodes are stratified by ply, so everything not ahead of the current position can be
    dropped.  Same games, visit counts and Q with 5x (and, forced, 8x) smaller arenas; whole games incl. restarts."""
    G, sims = 160, 20
    keep = make(net, G, sims, seed=13)                                           # 31 x sims + 64 nodes per tree
    small = make(net, G, sims, seed=13, recycle=1, eval_cache_log2=15)           # 6 x sims + 64
    tight = make(net, G, sims, seed=13, recycle=1, node_capacity=4 * sims)       # compacts almost every move
    for move in range(70):
        for e in (keep, small, tight):
            e.search(sims)
        if move % 5 == 0 or move > 55:
            ref = snapshot(keep)
            assert same(ref, snapshot(small)) and same(ref, snapshot(tight)), move
        for e in (keep, small, tight):
            e.play_device()
    ck, cs, ct = keep.counters(), small.counters(), tight.counters()
    assert ck['recycled_nodes'] == 0 and cs['recycled_nodes'] > 0 and ct['recycled_nodes'] > cs['recycled_nodes']
    assert cs['evicted_nodes'] == 0                                           # the exact rule was enough: nothing else was dropped
    assert ck['games_finished'] == cs['games_finished'] == ct['games_finished'] >= G
    assert ck['nodes'] == cs['nodes'] == ct['nodes']
    from collections import Counter                                            # same games (games that end in one launch
    a, b = keep.drain_replay(), tight.drain_replay()                           # reach the replay ring in any order)
    assert len(a) == len(b) and Counter(finished_games(a)) == Counter(finished_games(b))
    # continuous self-play on recycled arenas: runs whole games without overflowing the small arenas
    cont = make(net, G, sims, seed=13, recycle=1, eval_cache_log2=15)
    for _ in range(8):
        cont.selfplay(10 * sims, sims)
    assert cont.counters()['games_finished'] >= G and cont.counters()['recycled_nodes'] > 0


def test_sharp_priors_fall_back_to_edge_reachable_compaction(net):
    """A network with sharp priors grows deep trees, and a deep node stays ahead of the game for many moves: the exact rule
    keeps it, the reference's dicts simply grow, and a 6x arena overflowed in the third iteration of examples/alphazero_loop.py
    (MCAZ_ECAPACITY).  The second-stage compaction keeps what the tree's edges reach from the current position: the search
    goes on, the drop is counted (counter 17), every search still runs all its simulations and the games end."""
    import copy
    sharp = copy.deepcopy(net)
    with torch.no_grad():
        sharp.plinear.weight.mul_(80.0)
        sharp.plinear.bias.mul_(80.0)
    G, sims = 64, 50
    eng = make(sharp, G, sims, seed=5, recycle=1, eval_cache_log2=15, node_capacity=sims * 5 // 2)
    before = 0
    for move in range(70):
        eng.search(sims)
        c = eng.counters()
        states, results = eng.game_states()
        assert c['simulations'] - before == sims * int((results == 0).sum()), move     # every running game spent its whole budget
        before = c['simulations']
        codes, visits, q, n_legal = eng.root_stats()
        live = n_legal > 0
        assert (visits.sum(1)[live] >= sims - 1).all(), move
        eng.play_device()
    c = eng.counters()
    assert c['evicted_nodes'] > 0 and c['games_finished'] >= G
    # continuous self-play on the default recycled arena with the same network
    cont = make(sharp, G, sims, seed=5, recycle=1, eval_cache_log2=15)
    for _ in range(8):
        cont.selfplay(10 * sims, sims)
    assert cont.counters()['games_finished'] >= G


def test_more_games_than_one_network_pass(net):
    """More than 8192 leaf rows per batch are evaluated in equal chunks; results do not depend on the chunking:
    the first 300 games of a 9000-game engine play exactly like a 300-game engine with the same seed."""
    sims = 6
    big = make(net, 9000, sims, seed=31, recycle=1, eval_cache_log2=16)
    small = make(net, 300, sims, seed=31)
    for move in range(5):
        big.search(sims); small.search(sims)
        ids = np.arange(300, dtype=np.int32)
        a = big.root_stats(game_ids=ids) + big.game_states(game_ids=ids)
        assert same(a, snapshot(small)), move
        big.play_device(); small.play_device()
    c = big.counters()
    assert c['simulations'] == 5 * sims * 9000 and c['moves'] == 5 * 9000


def test_full_size_config3_invariants(net):
    """BASELINE.json configs[2] at full size (4096 games x 200 sims/move) through the public batched API, on a game
    population spread over all plies: size-independent properties of the search."""
    from minitchess_alphazero_b200.selfplay import BatchedSelfPlay
    G, sims = 4096, 200
    free0 = torch.cuda.mem_get_info()[0]
    sp = BatchedSelfPlay(net, n_games=G, num_simulations=sims, seed=77)
    sp.stagger()
    states, _ = sp.engine.game_states()
    plies = 2 * (((states['meta'] >> 16) & 0xff).astype(int) - 1) + (1 - (states['meta'] & 1).astype(int))
    assert len(np.unique(states)) > 0.8 * G and plies.min() == 0 and plies.max() >= 55       # all phases, distinct games
    c0 = sp.engine.counters()
    for _ in range(2):
        sp.search()
        codes, visits, q, n_legal = sp.engine.root_stats()
        assert (n_legal > 0).all()
        total = visits.sum(1)
        assert (total >= sims - 1).all()                  # sims - 1 edge visits on a fresh root, more on a reused one
        assert (np.abs(q) <= 1.0).all() and np.isfinite(q).all()
        live = np.arange(visits.shape[1])[None, :] < n_legal[:, None]
        assert (visits[~live] == 0).all() and (np.diff(codes.astype(int), axis=1)[live[:, 1:]] > 0).all()   # sorted legal codes
        sp.engine.play_device()
    c1 = sp.engine.counters()
    d = {k: c1[k] - c0[k] for k in c1}
    assert d['simulations'] == 2 * G * sims and d['moves'] == 2 * G
    assert d['simulations'] == d['evaluations'] + d['terminal_leaves'] + d['cached_evaluations']
    assert d['nodes'] <= d['simulations'] and d['cached_evaluations'] > 0.1 * d['simulations']
    assert d['path_depth'] >= d['simulations'] - G * 2    # every simulation but a root expansion descends at least one level
    used = free0 - torch.cuda.mem_get_info()[0]
    assert used < 12e9, used                              # recycled arenas (~4 GB) + cache + activations, not 18 GB of trees
