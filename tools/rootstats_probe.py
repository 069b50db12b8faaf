import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from minitchess_alphazero_b200.engine import Engine
from minitchess_alphazero_b200.policy import Network, flatten_state_dict
for G, sims in ((4096, 200), (4096, 8), (256, 200)):
    eng = Engine(G, max_sims_per_move=sims, network=1, device_rng=1)
    torch.manual_seed(0)
    eng.set_weights(flatten_state_dict(Network().state_dict()).numpy())
    eng.search(sims)
    torch.cuda.synchronize()
    for k in range(3):
        t = time.perf_counter(); eng.root_stats(want_q=False); torch.cuda.synchronize(); a = time.perf_counter() - t
        t = time.perf_counter(); eng.game_states(); torch.cuda.synchronize(); b = time.perf_counter() - t
        t = time.perf_counter(); eng.counters(); c = time.perf_counter() - t
        print(G, sims, 'root_stats %.2f ms  game_states %.2f ms counters %.2f ms' % (a * 1e3, b * 1e3, c * 1e3))
    eng.close()
