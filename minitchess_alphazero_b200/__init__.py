"""B200-native batched AlphaZero self-play for MinitChess.

Drop-in for the self-play hot path of schouhy/minitchess-alphazero (exp/agent.py,
exp/environment.py, exp/policy.py): the same Agent / Environment / Policy surface over
hand-written sm_100a CUDA kernels behind the C ABI of include/mcaz.h.
"""
__version__ = '0.1.0'
