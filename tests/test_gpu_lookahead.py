"""GPU: look-ahead rows (`lookahead_rows`, the option the per-agent facade turns on).  With one tree and one leaf per
network pass the tower's tile is empty but for one row, so every new node also queues its children as rows; their
evaluations go to the exact cache only.  The search must not notice: same order of simulations, same noise rows, same
visit counts and Q bits as the plain one-pass-per-simulation loop (exp/agent.py:41-45), with far fewer passes."""
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
    eng = Engine(n_games, max_sims_per_move=sims, network=1, **kw)
    eng.set_weights(flatten_state_dict(net.state_dict(), device='cuda'))
    return eng


def snapshot(eng):
    codes, visits, q, n_legal = eng.root_stats()
    states, results = eng.game_states()
    return codes, visits, q, n_legal, states, results


def same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize('n_games', [1, 5])
def test_caller_noise_search_is_unchanged_by_lookahead_rows(net, n_games):
    from minitchess_alphazero_b200._lib import MC_MAX_MOVES
    sims = 36
    plain = make(net, n_games, sims)
    ahead = make(net, n_games, sims, eval_cache_log2=14, lookahead_rows=255)
    small = make(net, n_games, sims, eval_cache_log2=14, lookahead_rows=7)      # the batch fills up: rows are dropped
    manual = make(net, n_games, sims)                                           # the loop spelt out: select / evaluate / backup
    engines = (plain, ahead, small, manual)
    rng = np.random.RandomState(3)
    for move in range(14):
        noise = np.zeros((sims, n_games, MC_MAX_MOVES))
        noise[:, :, :40] = rng.dirichlet([0.6] * 40, size=(sims, n_games))     # any positive rows do: only the legal prefix is read
        for e in engines[:3]:
            e.search_noise(noise)
        for k in range(sims):
            manual.select_expand(noise[k])
            manual.eval_backup()
        ref = snapshot(plain)
        for e in engines[1:]:
            assert same(ref, snapshot(e)), move
        codes, visits, _, n_legal = ref[:4]
        pick = np.array([codes[g, visits[g, :max(n_legal[g], 1)].argmax()] for g in range(n_games)], dtype=np.uint16)
        live = np.nonzero(ref[5] == 0)[0].astype(np.int32)
        if len(live) == 0:
            break
        for e in engines:
            e.play(pick[live], game_ids=live)
    c = [e.counters() for e in engines]
    assert c[0]['simulations'] == c[1]['simulations'] == c[2]['simulations']
    assert c[0]['nodes'] == c[1]['nodes'] and c[0]['edges'] == c[1]['edges']
    assert c[0]['cached_evaluations'] == 0
    # most simulations found their leaf evaluated ahead of time
    assert c[1]['evaluations'] < 0.35 * c[0]['evaluations'], (c[0]['evaluations'], c[1]['evaluations'])
    assert c[1]['evaluations'] <= c[2]['evaluations'] < c[0]['evaluations']


def test_device_noise_search_is_unchanged_by_lookahead_rows(net):
    sims = 24
    plain = make(net, 3, sims, device_rng=1, seed=9)
    ahead = make(net, 3, sims, device_rng=1, seed=9, eval_cache_log2=14, lookahead_rows=200)
    for move in range(10):
        plain.search(sims)
        ahead.search(sims)
        assert same(snapshot(plain), snapshot(ahead)), move
        plain.play_device()
        ahead.play_device()
    a, b = plain.counters(), ahead.counters()
    assert a['simulations'] == b['simulations'] and b['evaluations'] < 0.5 * a['evaluations']


def test_facade_uses_lookahead_rows_and_plays_the_same_game(net):
    """The drop-in agent with and without the option: identical visit distributions and moves under one numpy seed."""
    from minitchess_alphazero_b200.agent import MonteCarloTreeSearch, SimpleAlphaZeroAgent
    from minitchess_alphazero_b200.environment import MinitChessEnvironment
    from minitchess_alphazero_b200.policy import SimpleAlphaZeroPolicy
    env = MinitChessEnvironment()
    policy = SimpleAlphaZeroPolicy(net)
    games, evals = [], []
    for options in ({}, {'share_engine': False}, {'lookahead_rows': 0, 'eval_cache_log2': 0, 'share_engine': False}):
        np.random.seed(11)
        agents = [SimpleAlphaZeroAgent(env, policy, 20) for _ in range(2)]
        for a in agents:
            a._mcts = MonteCarloTreeSearch(env, policy.model, 1, engine_options=options)
        episode, obs = env.new_episode()
        trace = []
        for ply in range(12):
            if ply == 6:
                agents[0].init_mcts()           # one agent starts a new tree; the other's, in the same engine, must survive
            act = agents[ply & 1].select_action(obs)
            trace.append((int(act.action), act.info['pi'].tolist()))
            obs, _, done = episode.step(act.action)
            if done:
                break
        games.append(trace)
        engines = {id(a._mcts.engine): a._mcts.engine for a in agents}
        assert len(engines) == (1 if options.get('share_engine', True) else 2)
        evals.append(sum(e.counters()['evaluations'] for e in engines.values()))
    assert games[0] == games[1] == games[2]
    # shared engine (one cache for both agents) <= separate engines with look-ahead rows < plain
    assert evals[0] <= evals[1] < 0.5 * evals[2], evals


def test_lookahead_rows_beyond_one_network_pass(net):
    """More rows than one pass of the tower holds (8192): the dense batch runs as chunks, look-ahead rows included."""
    sims = 12
    plain = make(net, 40, sims, device_rng=1, seed=4)
    ahead = make(net, 40, sims, device_rng=1, seed=4, eval_cache_log2=16, lookahead_rows=9000)
    for move in range(5):
        plain.search(sims)
        ahead.search(sims)
        assert same(snapshot(plain), snapshot(ahead)), move
        plain.play_device()
        ahead.play_device()
    a, b = plain.counters(), ahead.counters()
    assert a['simulations'] == b['simulations'] and b['evaluations'] < a['evaluations']


def test_dropped_agents_free_their_tree_slots(net):
    """app/base.py:113 builds two new agents per batch of episodes: they must land on the engine the previous pair used."""
    from minitchess_alphazero_b200.agent import SimpleAlphaZeroAgent
    from minitchess_alphazero_b200.environment import MinitChessEnvironment
    from minitchess_alphazero_b200.policy import SimpleAlphaZeroPolicy
    env = MinitChessEnvironment()
    policy = SimpleAlphaZeroPolicy(net)
    seen = set()
    for batch in range(3):
        agents = [SimpleAlphaZeroAgent(env, policy, 6) for _ in range(2)]
        _, obs = env.new_episode()
        for a in agents:
            a.select_action(obs)
        assert agents[0]._mcts.engine is agents[1]._mcts.engine and agents[0]._mcts._tree != agents[1]._mcts._tree
        seen.add(id(agents[0]._mcts.engine))
        del agents, a
    assert len(seen) == 1


@pytest.mark.parametrize('cache', [0, 14])
@pytest.mark.parametrize('fen', ['2nQ1/1Q1p1/pk3/1qBqK/1q3/1qQ2 w 6 4',        # every legal move mates: all leaves below the root are finished
                                 '1Q3/pR3/k1PbR/R4/K2q1/1Br2 w 12 23',
                                 'k4/5/1QK2/5/5/5 w 2 20'])                   # mates and stalemates one ply down
def test_caller_noise_search_spends_its_whole_budget_on_finished_leaves(net, fen, cache):
    """One game, no look-ahead rows: once a mate is in the tree PUCT keeps coming back to it, so whole launches end with
    nobody waiting for a network row while simulations are still unspent.  The call must run all of them (it used to stop
    at the first such launch) and build the tree of the loop spelt out call by call (exp/agent.py:41-45)."""
    from minitchess_alphazero_b200._lib import MC_MAX_MOVES
    from minitchess_alphazero_b200 import rules
    sims = 36
    start = np.array([rules.state_from_fen(fen)])
    one = make(net, 1, sims, eval_cache_log2=cache, lookahead_rows=0)
    manual = make(net, 1, sims)
    rng = np.random.RandomState(5)
    for e in (one, manual):
        e.reset_games(states=start)
    for call in range(3):                                   # the later calls start with the mates already in the tree
        noise = np.zeros((sims, 1, MC_MAX_MOVES))
        noise[:, :, :40] = rng.dirichlet([0.6] * 40, size=(sims, 1))
        before = one.counters()['simulations']
        one.search_noise(noise)
        assert one.counters()['simulations'] - before == sims
        for k in range(sims):
            manual.select_expand(noise[k])
            manual.eval_backup()
        assert same(snapshot(one), snapshot(manual)), call
    c = one.counters()
    assert c['terminal_leaves'] > c['evaluations']          # most simulations ended on finished positions
    assert c['simulations'] == c['evaluations'] + c['terminal_leaves'] + c['cached_evaluations'] == 3 * sims
