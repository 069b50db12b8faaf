"""Per-phase timing of the host-buffer (e2e) step; diagnostic only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from minitchess_alphazero_b200.policy import Network
from minitchess_alphazero_b200.selfplay import BatchedSelfPlay

G, S = 4096, 200
torch.manual_seed(0)
sp = BatchedSelfPlay(Network().eval(), n_games=G, num_simulations=S, seed=1)
eng = sp.engine
for _ in range(2):
    sp.step()
torch.cuda.synchronize()
states, _ = eng.game_states()
plies = np.zeros(G, dtype=np.int32)
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.perf_counter() - t0
for it in range(3):
    t = time.perf_counter(); eng.set_positions(states, trees=plies & 1); tick('set_positions', t)
    t = time.perf_counter(); sp.search(); tick('search', t)
    t = time.perf_counter(); codes, visits, _, n_legal = eng.root_stats(want_q=False); tick('root_stats', t)
    t = time.perf_counter()
    E = np.maximum(n_legal, 1); w = visits.astype(np.float64); cum = np.cumsum(w, axis=1)
    u = np.random.random_sample(G) * cum[np.arange(G), E - 1]
    pick = np.minimum((cum <= u[:, None]).sum(1), E - 1)
    actions = codes[np.arange(G), pick]; tick('host_sample', t)
    t = time.perf_counter(); results = eng.play(actions); tick('play', t)
    t = time.perf_counter(); states, _ = eng.game_states(); tick('game_states', t)
    plies += 1
print({k: round(v / 3 * 1000, 2) for k, v in T.items()})
t = time.perf_counter()
for _ in range(3): sp.step()
torch.cuda.synchronize(); print('device step ms', (time.perf_counter() - t) / 3 * 1000)
