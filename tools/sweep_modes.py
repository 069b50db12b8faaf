"""Experiment: simulations/s of 4096 games x 200 sims/move for the search schedules and cache settings
(lock-step vs continuous self-play, free_sims, evaluation cache), with the tree-kernel and tower times of each."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minitchess_alphazero_b200.policy import Network
from minitchess_alphazero_b200.selfplay import BatchedSelfPlay

torch.manual_seed(0)
net = Network().eval()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 200
configs = [tuple(int(x) for x in a.split(',')) for a in sys.argv[3:]] or \
    [(0, 0, 1), (0, 23, 4), (1, 0, 1), (1, 23, 1), (1, 23, 2), (1, 23, 4), (1, 23, 8), (1, 23, 16)]
for cont, cache, free in configs:
    sp = BatchedSelfPlay(net, n_games=G, num_simulations=S, seed=1234, eval_cache_log2=cache, free_sims=free)
    step = (lambda: sp.run_continuous(S)) if cont else sp.step
    if os.environ.get('STAGGER', '1') != '0':
        sp.stagger(sims=int(os.environ.get('STAGGER_SIMS', '8')))
        st = sp.engine.game_states()[0]
        import numpy as np
        print('distinct positions after the pre-roll: %d of %d' % (len(np.unique(st)), G), flush=True)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    c0 = sp.engine.counters()
    sp.reset_kernel_timer()
    t = time.perf_counter()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    c1 = sp.engine.counters()
    d = {k: c1[k] - c0[k] for k in c1}
    tower_ms, n_fwd, _ = sp.engine.profile_network(False, read=True)
    tree_ms, n_tree = sp.engine.profile_tree(False, read=True)
    print('%s cache=%2d free=%2d: %.3f M sims/s, %.3f M evals/s, rows/fwd %.0f, hits %.3f, terminal %.4f, tower %.3f ms, tree %.3f ms/launch, '
          'moves %d, duplicate rows %.4f' % ('continuous' if cont else 'lockstep  ', cache, free, d['simulations'] / dt / 1e6, d['evaluations'] / dt / 1e6,
                        d['evaluations'] / max(n_fwd, 1), d['cached_evaluations'] / d['simulations'], d['terminal_leaves'] / d['simulations'],
                        tower_ms, tree_ms / max(n_tree, 1), d['moves'], d['duplicate_rows'] / max(d['evaluations'], 1)), flush=True)
    sp.engine.close()
    del sp
