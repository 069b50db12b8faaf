"""Times the built-in network forward alone (4096 boards): whole forward and per tower-conv launch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import numpy as np, torch
from minitchess_alphazero_b200 import _lib
from minitchess_alphazero_b200.engine import Engine
from minitchess_alphazero_b200.policy import Network, flatten_state_dict
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
FP8 = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # number of e4m3 convolutions (0 = the bf16 tower)
eng = Engine(G, max_sims_per_move=4, network=2, fp8_convolutions=FP8) if FP8 else Engine(G, max_sims_per_move=4, network=1)
torch.manual_seed(0)
eng.set_weights(flatten_state_dict(Network().state_dict()).numpy())
tok = torch.randint(0, 7, (G, 60), dtype=torch.uint8, device='cuda')
clk = torch.rand(G, device='cuda')
lg = torch.empty(G, 554, device='cuda'); vl = torch.empty(G, device='cuda')
L = _lib.lib()
def fwd():
    _lib.check(L.az_network_forward(eng._h, _lib.ptr(tok), _lib.ptr(clk), G, _lib.ptr(lg), _lib.ptr(vl)))
for _ in range(5): fwd()
eng.profile_network(True, read=True)
torch.cuda.synchronize(); t = time.perf_counter()
N = int(sys.argv[2]) if len(sys.argv) > 2 else 600
for _ in range(N): fwd()
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / N
ms, n, lpf = eng.profile_network(False, read=True)
ms = ms / 18          # per convolution (the tower's time also holds the stem level and the head epilogue)
print('G=%d forward %.3f ms (%.2f M evals/s); conv launch avg %.1f us over %d forwards; nominal %.0f TFLOP/s' % (
    G, dt * 1e3, G / dt / 1e6, ms * 1e3, n, G * 2 * 17694720 / (ms / 1e3) / 1e12))
