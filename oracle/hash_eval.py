"""Deterministic, torch-free leaf evaluator for tree-logic parity (T1).  TEST INFRASTRUCTURE ONLY.

Bit-exact visit counts only make sense when both sides see identical priors and values
(SURVEY.md §7.3 point 2).  This evaluator derives both from a BLAKE2 hash of the FEN with
integer arithmetic and one float32 division, so it gives the same bits on any machine.
`HashModel` wraps it in the duck type the reference's UNMODIFIED MonteCarloTreeSearch expects
from `model` (exp/agent.py:67-69: `model(model.process_observation(fen))`,
`p[0][legal].softmax(0).data.numpy()`, `v.item()`).
"""
from hashlib import blake2b

import numpy as np


def _h(text):
    return int.from_bytes(blake2b(text.encode(), digest_size=4).digest(), 'little')


def hash_priors(fen, legal):
    w = [1 + _h('%s|%d' % (fen, c)) % 1000 for c in legal]
    return np.array(w, dtype=np.float32) / np.float32(sum(w))


def hash_value(fen):
    return float(np.float32(_h(fen + '|v') % 2001 - 1000) / np.float32(1000))


def hash_evaluate(fen, legal):
    return hash_priors(fen, legal), hash_value(fen)


class _Vec:
    def __init__(self, fen, legal=None):
        self.fen, self.legal = fen, legal

    def __getitem__(self, item):
        return self if isinstance(item, int) and self.legal is None else _Vec(self.fen, list(item))

    def softmax(self, dim):
        return self

    @property
    def data(self):
        return self

    def numpy(self):
        return hash_priors(self.fen, self.legal)


class _Val:
    def __init__(self, fen):
        self.fen = fen

    def item(self):
        return hash_value(self.fen)


class HashModel:
    @staticmethod
    def process_observation(fen):
        return fen

    def __call__(self, fen):
        return _Vec(fen), _Val(fen)

    def eval(self):
        return self

    def cpu(self):
        return self
