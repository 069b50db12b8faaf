"""Where the tower's MMA issuer and TMA producer wait (MCAZ_TOWER_STATS=1): per-CTA clock-cycle counters of one launch.
    MCAZ_TOWER_STATS=1 python tools/tower_stats.py [rows]"""
import ctypes
import os
import sys

os.environ.setdefault('MCAZ_TOWER_STATS', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from minitchess_alphazero_b200 import _lib
from minitchess_alphazero_b200.engine import Engine
from minitchess_alphazero_b200.policy import Network, flatten_state_dict

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
FP8 = len(sys.argv) > 2 and sys.argv[2] == 'fp8'
eng = Engine(G, max_sims_per_move=4, network=2 if FP8 else 1)
torch.manual_seed(0)
eng.set_weights(flatten_state_dict(Network().state_dict()).numpy())
tok = torch.randint(0, 7, (G, 60), dtype=torch.uint8, device='cuda')
clk = torch.rand(G, device='cuda')
lg = torch.empty(G, 554, device='cuda')
vl = torch.empty(G, device='cuda')
L = _lib.lib()
for _ in range(20):
    _lib.check(L.az_network_forward(eng._h, _lib.ptr(tok), _lib.ptr(clk), G, _lib.ptr(lg), _lib.ptr(vl)))
torch.cuda.synchronize()
out = np.zeros(148 * 12, dtype=np.uint64)
L.az_tower_stats.restype = ctypes.c_int
n = L.az_tower_stats(eng._h, _lib.ptr(out), len(out))
assert n > 0, L.mcaz_last_error()
st = out[:n].reshape(-1, 12).astype(np.float64)
lead = st[0::2]
print('rows %d, %d CTA pairs%s' % (G, len(lead), ', e4m3 tower' if FP8 else ''))
print('MMA issuer : %.0f k cycles; waiting for operands (TMA) %.1f %%, for a free accumulator (epilogue) %.1f %%, issuing %.1f %%' % (
    lead[:, 0].mean() / 1e3, 100 * lead[:, 1].sum() / lead[:, 0].sum(), 100 * lead[:, 2].sum() / lead[:, 0].sum(),
    100 * (1 - (lead[:, 1].sum() + lead[:, 2].sum()) / lead[:, 0].sum())))
print('MMA issuer : %.1f %% of the operand wait falls on the first stage of a work item (%.0f items per pair, %.0f cycles per item there, %.0f cycles per item in all)' % (
    100 * lead[:, 6].sum() / max(lead[:, 1].sum(), 1), lead[:, 7].mean(), lead[:, 6].sum() / max(lead[:, 7].sum(), 1), lead[:, 0].sum() / max(lead[:, 7].sum(), 1)))
print('MMA issuer : cycles inside the issue of a stage (4 instructions + commit): e4m3 stages %.0f (%d stages per pair), bf16 stages %.0f (%d)' % (
    lead[:, 8].sum() / max(lead[:, 9].sum(), 1), lead[:, 9].mean(), lead[:, 10].sum() / max(lead[:, 11].sum(), 1), lead[:, 11].mean()))
print('TMA producer: %.0f k cycles; waiting for dependencies %.1f %%, for a free stage %.1f %%, issuing %.1f %%' % (
    st[:, 3].mean() / 1e3, 100 * st[:, 4].sum() / st[:, 3].sum(), 100 * st[:, 5].sum() / st[:, 3].sum(),
    100 * (1 - (st[:, 4].sum() + st[:, 5].sum()) / st[:, 3].sum())))
if G >= 512:
    print('per pair, operand-wait share: min %.1f %% max %.1f %%; dependency-wait share: min %.1f %% max %.1f %%' % (
    100 * (lead[:, 1] / lead[:, 0]).min(), 100 * (lead[:, 1] / lead[:, 0]).max(),
    100 * (st[:, 4] / st[:, 3]).min(), 100 * (st[:, 4] / st[:, 3]).max()))
