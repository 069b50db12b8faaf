"""az_config.defer_rows at bench size: how many batches a pass trims and what the tower then costs per evaluated row."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from minitchess_alphazero_b200.policy import Network
from minitchess_alphazero_b200.selfplay import BatchedSelfPlay

G, S = 4096, 200
torch.manual_seed(0)
net = Network().eval()
for mode in ('lockstep', 'continuous'):
    for d in [int(a) for a in sys.argv[1:]] or [0, 128]:
        sp = BatchedSelfPlay(net, n_games=G, num_simulations=S, seed=1234, eval_cache_log2=24, free_sims=4 if mode == 'continuous' else 0, defer_rows=d)
        sp.stagger()
        step = (lambda: sp.run_continuous(S)) if mode == 'continuous' else sp.step
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        c0 = sp.engine.counters()
        sp.engine.profile_network(True, read=True)
        t = time.perf_counter()
        for _ in range(4):
            step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        ms, n, _ = sp.engine.profile_network(False, read=True)
        c1 = sp.engine.counters()
        dl = {k: int(c1[k] - c0[k]) for k in c1}
        print('%-10s defer_rows %3d: %.3f M sims/s, %d passes, tower %.4f ms per pass, %.4f us per evaluated row, %d passes trimmed, %d rows deferred (%.1f per pass)' % (
            mode, d, dl['simulations'] / dt / 1e6, n, ms, 1e3 * ms * n / max(dl['evaluations'], 1), dl['trimmed_batches'], dl['deferred_rows'],
            dl['deferred_rows'] / max(n, 1)), flush=True)
        sp.engine.close()
        del sp
