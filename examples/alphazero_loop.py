#!/usr/bin/env python
"""BASELINE.json configs[4]: self-play on every GPU + learner step + NCCL weight broadcast per iteration.

    python examples/alphazero_loop.py --games 1024 --sims 50 --moves 20 --iterations 3
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/alphazero_loop.py ...
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

from minitchess_alphazero_b200 import _lib
from minitchess_alphazero_b200.loop import iteration
from minitchess_alphazero_b200.policy import Network
from minitchess_alphazero_b200.selfplay import BatchedSelfPlay


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--games', type=int, default=1024)
    ap.add_argument('--sims', type=int, default=50)
    ap.add_argument('--moves', type=int, default=20, help='self-play moves per iteration')
    ap.add_argument('--iterations', type=int, default=3)
    ap.add_argument('--batch-size', type=int, default=32)
    ap.add_argument('--lr', type=float, default=0.2)
    ap.add_argument('--eager-learner', action='store_true', help='launch the learner step kernel by kernel instead of replaying its CUDA graph')
    a = ap.parse_args()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    _lib.check(_lib.lib().mcaz_set_device(local))
    torch.manual_seed(0)
    net = Network().eval()
    sp = BatchedSelfPlay(net, n_games=a.games, num_simulations=a.sims, seed=100 + int(os.environ.get('RANK', '0')))
    for it in range(a.iterations):
        t0 = time.perf_counter()
        out = iteration(sp, net, a.moves, batch_size=a.batch_size, optim_params={'lr': a.lr}, graph=False if a.eager_learner else None)
        torch.cuda.synchronize()
        if int(os.environ.get('RANK', '0')) == 0:
            loss = sum(out['losses']) / len(out['losses']) if out['losses'] else float('nan')
            print('iteration %d: %d replay tuples, %d learner steps, mean loss %.4f, %.2f s' % (
                it, out['tuples'], len(out['losses']), loss, time.perf_counter() - t0))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
