"""GPU: the drop-in Agent / Environment / Policy classes and the batched self-play driver."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import rules_c as rc

pytestmark = pytest.mark.gpu
RES = {'*': 0, '1-0': 1, '0-1': 2, '1/2-1/2': 3}


@pytest.fixture(scope='module')
def net(mcaz_lib):
    from minitchess_alphazero_b200.policy import Network
    torch.manual_seed(0)
    return Network().eval()


def test_environment_facade_matches_golden(mcaz_lib):
    from minitchess_alphazero_b200.environment import (MinitChessEnvironment, TerminatedEpisodeStepException,
                                                       IlegalMoveException, NUM_ACTIONS, MOVES_DICT, STARTING_FEN)
    env = MinitChessEnvironment()
    ep, obs = env.new_episode()
    assert obs == STARTING_FEN and ep.turn is True and NUM_ACTIONS == 554
    assert [k for k, v in sorted(MOVES_DICT[True].items(), key=lambda kv: kv[1]) if v in ep.get_legal_moves()] == \
        ['a2a3', 'b2b3', 'c2c3', 'c1e2', 'c1d3', 'c1b3']
    for r in load_golden('rules_positions.json.gz')[:120]:
        ep, obs = env.new_episode(fen=r['fen'])
        assert obs == r['fen'] and ep.get_legal_moves() == r['legal'] and ep.is_done() == r['done']
        assert ep.get_result() == r['result']
        if r['done']:
            assert ep.get_reward() == r['reward']
            with pytest.raises(TerminatedEpisodeStepException):
                ep.step(r['legal'][0] if r['legal'] else 0)
            continue
        bad = next(c for c in range(554) if c not in r['legal'])
        with pytest.raises(IlegalMoveException):
            ep.step(bad)
        st = ep.step(r['legal'][0])
        assert [st.observation, st.reward if st.done else None, st.done] == r['children'][0]


def test_golden_game_through_environment_facade(mcaz_lib):
    """Replays the reference's recorded game line (incl. the final draw by the 30-move cap)."""
    from minitchess_alphazero_b200.environment import MinitChessEnvironment
    g = load_golden('mcts_hash_game.json')
    ep, obs = MinitChessEnvironment().new_episode()
    for ply in g['plies']:
        assert obs == ply['observation'] and ep.get_legal_moves() == ply['legal_moves']
        obs, reward, done = ep.step(ply['action'])
        assert obs == ply['next'] and done == ply['done'] and (reward == ply['reward'] or not done)


def test_network_parity_torch_evaluator(net):
    """T2: fp32 within 1e-5 relative of the reference Network's output; bf16 tower within 1e-2."""
    from torch_evaluator import TorchEvaluator
    g = load_golden('network_seed0.npz')
    tok = torch.from_numpy(g['tokens'].reshape(-1, 60)).cuda()
    clk = torch.from_numpy(g['clocks'].reshape(-1)).cuda()
    scale = np.abs(g['logits']).max()
    lg, v = TorchEvaluator(net, dtype=torch.float32).forward(tok, clk)
    assert np.abs(lg.cpu().numpy() - g['logits']).max() <= 1e-5 * max(scale, 1.0) * 10
    assert np.abs(v.cpu().numpy() - g['values'].reshape(-1)).max() <= 1e-5
    lg, v = TorchEvaluator(net, dtype=torch.bfloat16).forward(tok, clk)
    p_ref = torch.from_numpy(g['logits']).softmax(-1).numpy()
    assert np.abs(lg.softmax(-1).cpu().numpy() - p_ref).max() <= 1e-2
    assert np.abs(v.cpu().numpy() - g['values'].reshape(-1)).max() <= 1e-2


@pytest.mark.parametrize('evaluator', ['builtin', 'torch_fp32'])
def test_agent_facade_first_moves_match_reference_game(net, evaluator):
    """T3: the reference's config-1 game (random-init net, seed 0, 36 sims) through the drop-in
    agent classes.  fp32 GPU evaluation differs from the CPU in the last bits (and the bf16 tower within 1e-2), so
    visit counts are compared as distributions -- at most 3 of the 36 visits may sit elsewhere, and at least one ply must
    be identical (they are checked bit-exactly with shared evaluator bits in test_gpu_mcts / test_gpu_production_pin)."""
    from minitchess_alphazero_b200.agent import SimpleAlphaZeroAgent, RoundRobinReferee, MonteCarloTreeSearch
    from minitchess_alphazero_b200.environment import MinitChessEnvironment
    from minitchess_alphazero_b200.policy import SimpleAlphaZeroPolicy
    from torch_evaluator import TorchEvaluator
    g = load_golden('mcts_net_game.json')
    env = MinitChessEnvironment()
    policy = SimpleAlphaZeroPolicy(net)
    assert policy.num_actions() == 554 and policy.model is net
    agents = [SimpleAlphaZeroAgent(env, policy, g['sims']) for _ in range(2)]
    if evaluator == 'torch_fp32':
        for a in agents:
            a._mcts._evaluator = TorchEvaluator(net, dtype=torch.float32)
    referee = RoundRobinReferee(tuple(agents))
    np.random.seed(g['seed'])
    ep, obs = env.new_episode()
    same = 0
    for ply in g['plies'][:6]:
        assert obs == ply['observation']
        action = referee.select_action(obs)
        assert action.info['legal_moves'] == ply['legal_moves']
        pi = action.info['pi']
        assert isinstance(pi, np.ndarray) and pi.dtype == np.float64 and abs(pi.sum() - 1) < 1e-12
        assert np.abs(pi - np.array(ply['pi'])).max() <= 3.0 / g['sims'], (evaluator, pi.tolist(), ply['pi'])
        same += int(np.array_equal(pi, np.array(ply['pi'])))
        if int(action.action) != ply['action']:
            break                                   # the lines diverged; later plies are not comparable
        obs, _, _ = ep.step(int(action.action))
    assert same >= 1, evaluator                     # at least one ply with the very visit counts of the reference's game
    # dict-style access like the reference's MonteCarloTreeSearch.__getitem__
    tree = agents[0]._mcts
    assert isinstance(tree, MonteCarloTreeSearch)
    root = g['plies'][0]['observation']
    assert tree['legal_moves'][root] == g['plies'][0]['legal_moves']
    assert root in tree['visited'] and tree['N'][root].sum() >= g['sims'] - 1
    assert tree['P'][root].dtype == np.float32 and abs(tree['P'][root].sum() - 1) < 1e-5
    # ... and as whole dicts (exp/agent.py:25-36): every key of N answers the point lookups with the same arrays
    keys = tree['N'].keys()
    assert root in keys and len(keys) == len(tree['Q']) == len(tree['legal_moves']) and len(tree['visited']) == len(keys) + len(tree['terminal'])
    for fen, n in tree['N'].items():
        assert np.array_equal(n, tree['N'][fen]) and len(n) == len(tree['legal_moves'][fen])
    assert sum(n.sum() for n in tree['N'].values()) >= g['sims'] - 1
    # a second game reuses the engine with empty trees (MonteCarloInit.on_episode_begin)
    eng = agents[0]._mcts.engine
    agents[0].init_mcts()
    assert agents[0]._mcts.engine is eng and agents[0]._mcts['N'].get(g['plies'][0]['observation']) is None


def test_weights_reload_is_seen(net):
    from minitchess_alphazero_b200.agent import SimpleAlphaZeroAgent
    from minitchess_alphazero_b200.environment import MinitChessEnvironment, STARTING_FEN
    from minitchess_alphazero_b200.policy import SimpleAlphaZeroPolicy, Network
    torch.manual_seed(3)
    mine = Network().eval()
    agent = SimpleAlphaZeroAgent(MinitChessEnvironment(), SimpleAlphaZeroPolicy(mine), 4)
    np.random.seed(0)
    agent.select_action(STARTING_FEN)
    p0 = agent._mcts['P'][STARTING_FEN].copy()
    mine.load_state_dict(net.state_dict())                      # app/base.py:126-129
    agent.init_mcts()
    agent.select_action(STARTING_FEN)
    p1 = agent._mcts['P'][STARTING_FEN]
    assert not np.allclose(p0, p1)


def test_batched_selfplay_replay_format(net):
    from minitchess_alphazero_b200.selfplay import BatchedSelfPlay, replay_to_episode_dicts
    sp = BatchedSelfPlay(net, n_games=96, num_simulations=6, seed=1)
    sp.run(70)                                                   # > 60 plies: every game finishes at least once
    c = sp.engine.counters()
    assert c['simulations'] > 0 and c['evaluations'] > 0 and c['games_finished'] >= 96
    assert c['simulations'] == c['evaluations'] + c['terminal_leaves'] + c['cached_evaluations']
    tuples = sp.drain()
    assert len(tuples) >= 96
    eps = replay_to_episode_dicts(tuples[:400])
    states = rc.fens_to_states([e['observation'] for e in eps])
    codes, counts, _ = rc.legal_moves(states)
    for i, e in enumerate(eps):
        assert e['legal_moves'] == codes[i, :counts[i]].tolist()
        assert e['action'] in e['legal_moves']
        assert abs(sum(e['pi']) - 1.0) < 1e-5 and e['reward'] in (-1.0, 0.0, 1.0)
    # rewards alternate along one game (exp/callbacks.py:49-53)
    first = [i for i, e in enumerate(eps) if e['observation'] == '2nbk/2ppp/5/5/PPP2/KBN2 w 0 1']
    if len(first) >= 2:
        game = eps[first[0]:first[1]]
        assert all(game[k]['reward'] == -game[k + 1]['reward'] for k in range(len(game) - 1))


def test_batched_agent_select_actions(mcaz_lib):
    """BatchedAlphaZeroAgent.select_actions = select_action for many games at once: legal moves and pi of every game
    equal the engine's root statistics, actions are legal, trees are kept per colour across moves."""
    from minitchess_alphazero_b200.agent import BatchedAlphaZeroAgent
    from minitchess_alphazero_b200.environment import MinitChessEnvironment
    from minitchess_alphazero_b200.policy import Network, SimpleAlphaZeroPolicy
    torch.manual_seed(0)
    policy = SimpleAlphaZeroPolicy(Network().eval())
    n, sims = 48, 12
    agent = BatchedAlphaZeroAgent(policy, n_games=n, num_simulations=sims, seed=4, rng=np.random.RandomState(0))
    env = MinitChessEnvironment()
    episodes, obs = zip(*[env.new_episode() for _ in range(n)])
    obs = list(obs)
    agent.init_mcts()
    for ply in range(6):
        actions = agent.select_actions(obs)
        assert len(actions) == n
        for k, (ep, a) in enumerate(zip(episodes, actions)):
            assert a.info['legal_moves'] == ep.get_legal_moves()
            assert a.action in a.info['legal_moves']
            assert abs(a.info['pi'].sum() - 1.0) < 1e-12
            visits = a.info['pi'] * (sims - 1 if ply < 2 else 1)
            if ply < 2:                                          # fresh trees: sims - 1 edge visits (exp/policy.py:118-121)
                assert np.allclose(visits, np.round(visits))
            obs[k] = ep.step(a.action).observation
    assert len(set(obs)) > n // 2                                # per-game noise and sampling: the games diverge
    c = agent.engine.counters()
    assert c['simulations'] == 6 * n * sims
    # new episodes for some games: MonteCarloInit empties their trees
    agent.init_mcts(game_ids=np.array([3, 7], dtype=np.int32))
    for i in (3, 7):
        episodes = list(episodes)
        episodes[i], obs[i] = env.new_episode()
    again = agent.select_actions(obs)
    for i in (3, 7):
        v = again[i].info['pi'] * (sims - 1)
        assert np.allclose(v, np.round(v)) and again[i].info['legal_moves'] == episodes[i].get_legal_moves()
