"""Throughput of the stateless rules kernels (BASELINE.json configs[1]: move generation on ~1 M reachable positions) with
device-resident buffers, against the HBM copy peak: legal_moves_kernel, apply_kernel, tokenize_kernel, and perft from the
start position.  Positions come from uniform random playouts run with the same kernels."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from minitchess_alphazero_b200 import _lib, rules
from minitchess_alphazero_b200._lib import MC_MAX_MOVES, MC_TOKENS, check, lib, ptr

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
REPS = 20
L = lib()
dev = 'cuda'
torch.manual_seed(0)
start = np.frombuffer(np.asarray(rules.state_from_fen(rules.STARTING_FEN)).tobytes(), dtype=np.int32)
states = torch.from_numpy(np.tile(start, (N, 1))).to(dev).contiguous()          # int32 [N, 5]
codes = torch.zeros(N, MC_MAX_MOVES, dtype=torch.int16, device=dev)
counts = torch.zeros(N, dtype=torch.int32, device=dev)
results = torch.zeros(N, dtype=torch.int8, device=dev)
nxt = torch.zeros_like(states)
status = torch.zeros(N, dtype=torch.int8, device=dev)
tokens = torch.zeros(N, MC_TOKENS, dtype=torch.uint8, device=dev)
clocks = torch.zeros(N, dtype=torch.float32, device=dev)


def legal():
    check(L.mc_legal_moves(ptr(states), N, None, ptr(codes), ptr(counts), ptr(results)))


def advance(plies):
    """every game plays a uniformly random number of random legal moves (finished games stay where they are)"""
    global states
    want = torch.randint(0, plies, (N,), device=dev)
    for p in range(plies):
        legal()
        pick = (torch.rand(N, device=dev) * counts.clamp(min=1)).long().clamp(max=MC_MAX_MOVES - 1)
        move = codes.gather(1, pick[:, None]).squeeze(1).contiguous()
        check(L.mc_apply(ptr(states), ptr(move), N, None, ptr(nxt), ptr(status)))
        go = (status == 0) & (results == 0) & (counts > 0) & (want > p)
        states = torch.where(go[:, None], nxt, states).contiguous()


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REPS


advance(40)
legal()
mean_moves = float(counts.float().mean())
ongoing = float((results == 0).float().mean())
pick = (torch.rand(N, device=dev) * counts.clamp(min=1)).long()
move = codes.gather(1, pick[:, None]).squeeze(1).contiguous()
peak = 6547.5
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
out = {'positions': N, 'mean_legal_moves': mean_moves, 'ongoing_fraction': ongoing, 'hbm_peak_gbs': peak, 'kernels': {}}
for name, fn, byts in (
        ('legal_moves_kernel', legal, 20 + 2 * mean_moves + 5),                  # position in; sorted codes, count, result out
        ('apply_kernel', lambda: check(L.mc_apply(ptr(states), ptr(move), N, None, ptr(nxt), ptr(status))), 20 + 2 + 20 + 1),
        ('tokenize_kernel', lambda: check(L.mc_tokenize(ptr(states), N, ptr(tokens), ptr(clocks))), 20 + 60 + 4)):
    ms = timed(fn)
    gbs = N * byts / (ms / 1e3) / 1e9
    out['kernels'][name] = {'ms': ms, 'positions_per_second': N / (ms / 1e3), 'algorithmic_bytes_per_position': byts,
                            'achieved_gbs': gbs, 'frac_of_hbm_peak': gbs / peak}
root = np.asarray(rules.state_from_fen(rules.STARTING_FEN)).reshape(1)
for depth in (6, 7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rules.perft(root, depth)
    torch.cuda.synchronize()
    e0.record()
    nodes = int(rules.perft(root, depth)[0])
    e1.record()
    torch.cuda.synchronize()
    out['perft_%d' % depth] = {'nodes': nodes, 'ms': e0.elapsed_time(e1), 'nodes_per_second': nodes / (e0.elapsed_time(e1) / 1e3)}
print(json.dumps(out))
