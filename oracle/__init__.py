"""CPU oracle for the MinitChess AlphaZero self-play hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / reference arm may
import anything from this package.  The product (`minitchess_alphazero_b200`) never does.

Parity status: MCTS / policy / tokeniser restatements are PINNED against the
reference's own unmodified `exp/agent.py` + `exp/policy.py` run in the build container
(`tests/golden/make_golden.py`).  The MinitChess *rules* are **PARITY UNPINNED**: they live in
a python-chess fork that is absent from the reference tree; see `oracle/shims/chess`.
"""
