"""Experiment: 4096 games as two engines of 2048 on their own streams, driven by two host threads, so one
engine's tree / stem / head kernels overlap the other's tower."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minitchess_alphazero_b200.policy import Network
from minitchess_alphazero_b200.selfplay import BatchedSelfPlay
torch.manual_seed(0)
net = Network().eval()
def run(n_engines, games_each, steps=4, own=1):
    sps = [BatchedSelfPlay(net, n_games=games_each, num_simulations=200, seed=10 + i, own_stream=own) for i in range(n_engines)]
    def work(sp, n):
        for _ in range(n): sp.step()
    for phase in (2, steps):
        ths = [threading.Thread(target=work, args=(sp, phase)) for sp in sps]
        torch.cuda.synchronize(); t = time.perf_counter()
        c0 = [sp.engine.counters()['simulations'] for sp in sps]
        for th in ths: th.start()
        for th in ths: th.join()
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    sims = sum(sp.engine.counters()['simulations'] - c for sp, c in zip(sps, c0))
    print('%d engine(s) x %d games, own_stream=%d: %.0f sims/s' % (n_engines, games_each, own, sims / dt))
    del sps
run(1, 4096, own=0)
run(1, 4096, own=1)
run(2, 2048, own=1)
run(4, 1024, own=1)
