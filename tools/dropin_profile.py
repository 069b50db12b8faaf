"""Where a move of the per-agent drop-in (one game, sequential search) spends its time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from minitchess_alphazero_b200 import _lib
from minitchess_alphazero_b200.agent import RoundRobinReferee, SimpleAlphaZeroAgent, MonteCarloTreeSearch
from minitchess_alphazero_b200.environment import MinitChessEnvironment
from minitchess_alphazero_b200.policy import Network, SimpleAlphaZeroPolicy

sims = int(sys.argv[1]) if len(sys.argv) > 1 else 36
opts = {} if len(sys.argv) < 3 or sys.argv[2] != 'plain' else {'lookahead_rows': 0, 'eval_cache_log2': 0}
if len(sys.argv) > 2 and sys.argv[2].startswith('la='):
    opts = {'lookahead_rows': int(sys.argv[2][3:])}
torch.manual_seed(0); np.random.seed(0)
env = MinitChessEnvironment()
policy = SimpleAlphaZeroPolicy(Network().eval())
agents = [SimpleAlphaZeroAgent(env, policy, sims) for _ in range(2)]
for a in agents:
    a._mcts = MonteCarloTreeSearch(env, policy.model, 1, engine_options=opts)
T = {'simulate': 0.0, 'select_action': 0.0, 'step': 0.0}
orig = MonteCarloTreeSearch.simulate
def timed(self, n, obs):
    t = time.perf_counter(); r = orig(self, n, obs); T['simulate'] += time.perf_counter() - t; return r
MonteCarloTreeSearch.simulate = timed
plies = 0
t_all = time.perf_counter()
for ep in range(6):
    if ep == 1:   # episode 0 warms up
        for k in T: T[k] = 0.0
        plies = 0; t_all = time.perf_counter()
        l0 = _lib.lib().mcaz_kernel_launches()
        engines = list({id(a._mcts.engine): a._mcts.engine for a in agents}.values())
        c0 = [e.counters() for e in engines]
        for e in engines:
            e.profile_network(True, read=True); e.profile_tree(True, read=True)
    for a in agents: a.init_mcts()
    episode, obs = env.new_episode()
    done, turn = False, 0
    while not done:
        t = time.perf_counter(); act = agents[turn].select_action(obs); T['select_action'] += time.perf_counter() - t
        t = time.perf_counter(); obs, r, done = episode.step(act.action); T['step'] += time.perf_counter() - t
        turn ^= 1; plies += 1
dt = time.perf_counter() - t_all
net_ms = tree_ms = 0.0; n_fwd = n_tree = 0
for e in engines:
    ms, n, _ = e.profile_network(False, read=True); net_ms += ms * n; n_fwd += n
    ms, n = e.profile_tree(False, read=True); tree_ms += ms; n_tree += n
print('per move: %.1f tower launches %.3f ms, %.1f search launches %.3f ms' % (n_fwd / plies, net_ms / plies, n_tree / plies, tree_ms / plies))
c1 = [e.counters() for e in engines]
ev = sum(b['evaluations'] - a['evaluations'] for a, b in zip(c0, c1))
ch = sum(b['cached_evaluations'] - a['cached_evaluations'] for a, b in zip(c0, c1))
print('%d plies, %.2f ms/move (%.0f sims/s): simulate %.2f, rest of select_action %.2f, env.step %.2f ms; network rows/move %.1f, cache hits/move %.1f, launches/move %.1f' % (
    plies, 1e3 * dt / plies, plies * sims / dt, 1e3 * T['simulate'] / plies, 1e3 * (T['select_action'] - T['simulate']) / plies,
    1e3 * T['step'] / plies, ev / plies, ch / plies, (_lib.lib().mcaz_kernel_launches() - l0) / plies))
