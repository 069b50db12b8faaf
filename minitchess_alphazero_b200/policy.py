"""Drop-in for exp/policy.py: `Network` (same state_dict keys and shapes, so `load_state_dict`
and the reference learner work unchanged), the FEN tokeniser, and `SimpleAlphaZeroPolicy`.

`Network` is an ordinary torch module: it is what the learner trains and what weights arrive in.
Self-play never runs it on the CPU; `get_distribution` drives the GPU search engine, which reads
the weights through `flatten_state_dict` (`az_set_weights`).
"""
import numpy as np
import torch
from torch import nn

from ._lib import AZ_NUM_WEIGHT_FLOATS
from .erlyx_compat import Policy

NUM_ACTIONS = 554
EMBEDDING_DIM = 4
MAX_NUM_MOVES_ALLOWED = 30           # exp/policy.py:12
PIECE_TOKENS = '0prbnqk'             # exp/policy.py:7


class ConvBlock(nn.Module):
    """conv -> batchnorm [-> relu]; parameters live under `.layers.{0,1}` (exp/policy.py:15-38)."""

    def __init__(self, nin, nout, kernel_size, stride, padding, batchnorm=True, nonlinearity=True):
        super().__init__()
        stack = [nn.Conv2d(nin, nout, kernel_size, stride, padding), nn.BatchNorm2d(nout)]
        if nonlinearity:
            stack.append(nn.ReLU())
        self.layers = nn.Sequential(*stack)

    def forward(self, x):
        return self.layers(x)


class ResidualBlock(nn.Module):
    """exp/policy.py:41-50."""

    def __init__(self, nin, nhid, nout):
        super().__init__()
        self.convblock1 = ConvBlock(nin, nhid, 3, 1, 1)
        self.convblock2 = ConvBlock(nhid, nout, 3, 1, 1, nonlinearity=False)
        self.nonl = nn.ReLU()

    def forward(self, x):
        return self.nonl(self.convblock2(self.convblock1(x)) + x)


class Network(nn.Module):
    """Embedding(7,4) -> Conv(8->256) -> 9 x Residual(256) -> policy head (554) + value head (tanh).
    10 693 458 parameters; exp/policy.py:53-105."""

    def __init__(self, num_actions=NUM_ACTIONS):
        super().__init__()
        self.emb = nn.Embedding(7, EMBEDDING_DIM)
        self.resbody = nn.Sequential(ConvBlock(2 * EMBEDDING_DIM, 256, 3, 1, 1),
                                     *[ResidualBlock(256, 256, 256) for _ in range(9)])
        self.pconv = ConvBlock(256, 2, 1, 1, 0)
        self.plinear = nn.Linear(2 * 30 + 1, num_actions)
        self.vconv = ConvBlock(256, 1, 1, 1, 0)
        self.vlinear = nn.Sequential(nn.Linear(30 + 1, 256), nn.ReLU(), nn.Linear(256, 1), nn.Tanh())

    def forward(self, input_data):
        tokens, clock = input_data
        x = self.emb(tokens).permute(0, 1, 4, 2, 3).contiguous().view(-1, 2 * EMBEDDING_DIM, 6, 5)
        x = self.resbody(x)
        p = self.plinear(torch.cat([self.pconv(x).view(-1, 60), clock], dim=1))
        v = self.vlinear(torch.cat([self.vconv(x).view(-1, 30), clock], dim=1))
        return p, v

    # ---- tokeniser (exp/policy.py:82-105) --------------------------------------------------
    @classmethod
    def tokenize(cls, bfen, color):
        """60 tokens: the mover's pieces then the opponent's, board seen from the mover's side."""
        cells = []
        for ch in bfen:
            if ch == '/':
                continue
            if ch.isdigit():
                cells.extend('0' * int(ch))
            else:
                cells.append(ch)
        if color == 'b':                       # rotate 180 degrees and swap colours
            cells = [c.swapcase() for c in reversed(cells)]
        mine = [PIECE_TOKENS.index(c.lower()) if c.isupper() else 0 for c in cells]
        theirs = [PIECE_TOKENS.index(c) if c.islower() else 0 for c in cells]
        return mine + theirs

    @classmethod
    def process_observation(cls, observation):
        bfen, color, _halfmove, fullmove = observation.split()
        channels = torch.LongTensor(cls.tokenize(bfen, color)).reshape(1, 2, 6, 5)
        clock = float(fullmove) + (0.5 if color == 'b' else 0.0)
        return channels, torch.tensor([[clock / MAX_NUM_MOVES_ALLOWED]]).float()


def flatten_state_dict(state_dict, device=None):
    """float32 vector in state_dict order without the num_batches_tracked counters: the layout
    az_set_weights expects (parameters 10 693 458 + BatchNorm running statistics 9 734)."""
    parts = [v.detach().reshape(-1).float() for k, v in state_dict.items() if not k.endswith('num_batches_tracked')]
    flat = torch.cat([p.to(device) if device is not None else p for p in parts]).contiguous()
    assert flat.numel() == AZ_NUM_WEIGHT_FLOATS, flat.numel()
    return flat


def weights_fingerprint(module):
    """Cheap staleness check: torch bumps `_version` on every in-place update / load_state_dict.  Walks the module's own
    parameter and buffer objects (same version counters as the state_dict aliases, without building the dict)."""
    return tuple(p._version for p in module.parameters()) + tuple(b._version for b in module.buffers())


class SimpleAlphaZeroPolicy(Policy):
    """exp/policy.py:107-125."""

    def __init__(self, network=None):
        self._network = network or Network()

    @property
    def model(self):
        return self._network

    def get_distribution(self, observation, mcts, num_simulations):
        with torch.no_grad():
            mcts.simulate(num_simulations, observation)
            legal_moves = mcts['legal_moves'][observation]
            N = mcts['N'][observation]
            return {'legal_moves': legal_moves, 'pi': N / N.sum()}

    def num_actions(self):
        return NUM_ACTIONS
