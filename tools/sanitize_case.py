"""Small end-to-end case for compute-sanitizer (memcheck): rules kernels, search with the built-in network,
device move choice, replay drain, collate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from minitchess_alphazero_b200 import rules
from minitchess_alphazero_b200.policy import Network
from minitchess_alphazero_b200.selfplay import BatchedSelfPlay, collate_device



def random_positions(seed, n):
    """Positions visited by uniform-random legal playouts from the start position, generated with the
    library's own rules kernels (the oracle is test infrastructure and is not imported by tools)."""
    rng = np.random.RandomState(seed)
    walkers = np.repeat(rules.states_from_fens([rules.STARTING_FEN]), 256)
    out = []
    while sum(len(o) for o in out) < n:
        out.append(walkers.copy())
        codes, counts, results = rules.legal_moves(walkers)
        live = (results == 0) & (counts > 0)
        pick = codes[np.arange(len(walkers)), (rng.random_sample(len(walkers)) * np.maximum(counts, 1)).astype(np.int64)]
        nxt, status = rules.apply(walkers, pick)
        ok = live & (status == 0)
        walkers = nxt
        walkers[~ok] = rules.state_from_fen(rules.STARTING_FEN)
    return np.ascontiguousarray(np.concatenate(out)[:n])


pos = random_positions(1, 3000)
c, n, r = rules.legal_moves(pos)
rules.apply(pos, c[:, 0]); rules.tokenize(pos); rules.perft(pos[:4], 3)
torch.manual_seed(0)
sp = BatchedSelfPlay(Network().eval(), n_games=96, num_simulations=5, seed=3)
sp.run(64)
t = sp.drain()
out = collate_device(t[:200])
torch.cuda.synchronize()
print('sanitize case ok', len(t), sp.engine.counters())
