"""Launches every kernel of libmcaz.so a few times at working sizes -- the case `ncu --set full` walks for the per-kernel pages
under profiles/ (tools/profile_round.sh).  The hot kernels of a 200-simulation search (search_step_kernel, heads_legal_kernel,
stem_onehot_kernel, tower_tc_kernel) are captured inside bench.py as well; here they run on trees a few moves old.
    python tools/all_kernels.py [games] [all|lookahead]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from minitchess_alphazero_b200 import rules
from minitchess_alphazero_b200._lib import MC_MAX_MOVES
from minitchess_alphazero_b200.engine import Engine, sample_root_noise
from minitchess_alphazero_b200.policy import Network, flatten_state_dict
from minitchess_alphazero_b200.selfplay import BatchedSelfPlay, collate_device

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ONLY_LOOKAHEAD = len(sys.argv) > 2 and sys.argv[2] == 'lookahead'
torch.manual_seed(0)
net = Network().eval()
flat = flatten_state_dict(net.state_dict(), device='cuda')

rng = np.random.RandomState(1)


def main_sections():
    # ---- stateless rules kernels on positions from random playouts (legal_moves_kernel, apply_kernel, tokenize_kernel, perft_*)
    walkers = np.repeat(rules.states_from_fens([rules.STARTING_FEN]), 65536)
    for _ in range(3):
        codes, counts, results = rules.legal_moves(walkers)
        pick = codes[np.arange(len(walkers)), (rng.random_sample(len(walkers)) * np.maximum(counts, 1)).astype(np.int64)]
        nxt, status = rules.apply(walkers, pick)
        ok = (results == 0) & (counts > 0) & (status == 0)
        nxt[~ok] = walkers[~ok]
        walkers = nxt
    rules.tokenize(walkers)
    rules.perft(walkers[:64], 3)

    # ---- throughput mode: search_step_kernel<false>, stem_onehot_kernel, tower_tc_kernel, heads_legal_kernel<false>, play_device_kernel,
    # restart_finished_kernel, recycle_kernel (arenas of 3 x sims nodes fill up every other move), prep_* (set_weights)
    sp = BatchedSelfPlay(net, n_games=G, num_simulations=32, seed=5, node_capacity=2 * 32 + 64, eval_cache_log2=22)
    for _ in range(4):
        sp.step()
    sp.run_continuous(8)
    tuples = sp.drain()
    eng = sp.engine
    ids = np.arange(G, dtype=np.int32)
    eng.root_stats()                                   # root_stats_kernel
    states, results = eng.game_states()                # game_states_kernel
    eng.node_stats(0, 0, states[0])                    # node_stats_kernel
    eng.tree_dump(0, 0)
    eng.reset_games(game_ids=ids[:64])                 # reset_games_kernel
    eng.reset_trees(ids[:64])                          # reset_trees_kernel
    eng.set_positions(states, trees=1 - (states['meta'] & 1).astype(np.int32))     # set_positions_kernel
    eng.search(4)
    c, v, _, n = eng.root_stats(want_q=False)
    live = np.nonzero((results == 0) & (n > 0))[0].astype(np.int32)
    eng.play(c[live, 0], game_ids=live)                # play_kernel
    tok = torch.randint(0, 7, (G, 60), dtype=torch.uint8, device='cuda')
    eng.network_forward(tok.cpu().numpy(), np.random.rand(G).astype(np.float32))   # heads_kernel (all 554 logits)
    if len(tuples) < 4096:                             # few games finish in so few moves: tuples of the current positions
        from minitchess_alphazero_b200.engine import REPLAY_DTYPE
        tuples = np.zeros(4096, dtype=REPLAY_DTYPE)
        tuples['observation'] = np.resize(states, 4096)
        tuples['n_legal'] = 1
    collate_device(tuples[:4096])                      # collate_kernel
    sample_root_noise(1, 0.6, 9, 65536)                # sample_root_noise_kernel
    eng.close()

    # ---- the external-evaluator calls of the parity mode: select_expand_kernel, backup_kernel
    ext = Engine(G, max_sims_per_move=16, device_rng=1, seed=2)
    pri = np.full((G, MC_MAX_MOVES), 1.0 / MC_MAX_MOVES, dtype=np.float32)
    val = np.zeros(G, dtype=np.float32)
    for _ in range(3):
        ext.select_expand()
        ext.backup(val, priors=pri)
    ext.close()



if not ONLY_LOOKAHEAD:
    main_sections()

# ---- the per-agent drop-in's engine: search_step_kernel<true>, heads_legal_kernel<true>, untag_rows_kernel (look-ahead rows)
one = Engine(2, max_sims_per_move=36, network=1, eval_cache_log2=18, lookahead_rows=255)
one.set_weights(flat)
noise = np.zeros((36, 2, MC_MAX_MOVES))
noise[:, :, :40] = rng.dirichlet([0.6] * 40, size=(36, 2))
for _ in range(3):
    one.search_noise(noise)
    c, v, _, n = one.root_stats(want_q=False)
    one.play(np.array([c[g, v[g, :max(n[g], 1)].argmax()] for g in range(2)], dtype=np.uint16))
one.close()
torch.cuda.synchronize()
print('all kernels ok')
