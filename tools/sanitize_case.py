"""Small end-to-end case for compute-sanitizer (memcheck): rules kernels, search with the built-in network,
device move choice, replay drain, collate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from minitchess_alphazero_b200 import rules
from minitchess_alphazero_b200.policy import Network
from minitchess_alphazero_b200.selfplay import BatchedSelfPlay, collate_device
from oracle import rules_c as rc
pos = np.ascontiguousarray(rc.random_positions(1, 3000))
c, n, r = rules.legal_moves(pos)
rules.apply(pos, c[:, 0]); rules.tokenize(pos); rules.perft(pos[:4], 3)
torch.manual_seed(0)
sp = BatchedSelfPlay(Network().eval(), n_games=96, num_simulations=5, seed=3)
sp.run(64)
t = sp.drain()
out = collate_device(t[:200])
torch.cuda.synchronize()
print('sanitize case ok', len(t), sp.engine.counters())
