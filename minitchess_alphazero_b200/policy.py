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


class TorchEvaluator:
    """Batched leaf evaluation of a `Network` on the GPU with library kernels (cuDNN / cuBLAS via
    PyTorch): BatchNorm folded into the convolutions, channels-last, bf16 tower by default with
    fp32 heads.  This is the library baseline that the hand-written tcgen05 tower is measured
    against; no CPU path."""

    def __init__(self, network, dtype=torch.bfloat16, device='cuda'):
        if not torch.cuda.is_available():
            raise RuntimeError('TorchEvaluator needs a CUDA device (self-play has no CPU fallback)')
        self.device, self.dtype = torch.device(device), dtype
        self.load(network)
        self._graph = None
        torch.backends.cudnn.benchmark = True

    @staticmethod
    def _fold(block):
        conv, bn = block.layers[0], block.layers[1]
        scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
        w = conv.weight.detach() * scale.view(-1, 1, 1, 1)
        b = (conv.bias.detach() - bn.running_mean) * scale + bn.bias.detach()
        return w, b

    def load(self, network):
        dev, dt = self.device, self.dtype
        self.emb = network.emb.weight.detach().float().to(dev)

        def put(w, b, tower=True):
            w = w.float().to(dev)
            if tower:
                w = w.to(dt).contiguous(memory_format=torch.channels_last)
                return w, b.float().to(dev).to(dt)
            return w, b.float().to(dev)
        self.stem = put(*self._fold(network.resbody[0]))
        self.blocks = [(put(*self._fold(blk.convblock1)), put(*self._fold(blk.convblock2))) for blk in list(network.resbody)[1:]]
        self.pconv = put(*self._fold(network.pconv), tower=False)
        self.vconv = put(*self._fold(network.vconv), tower=False)
        self.plinear = (network.plinear.weight.detach().float().to(dev), network.plinear.bias.detach().float().to(dev))
        self.v1 = (network.vlinear[0].weight.detach().float().to(dev), network.vlinear[0].bias.detach().float().to(dev))
        self.v2 = (network.vlinear[2].weight.detach().float().to(dev), network.vlinear[2].bias.detach().float().to(dev))
        self._graph = None

    @torch.no_grad()
    def forward(self, tokens_u8, clocks):
        """tokens uint8 [B,60], clocks float32 [B] (CUDA) -> logits float32 [B,554], values float32 [B]."""
        if self.dtype == torch.float32:            # true fp32 (no TF32) when used as the parity reference
            with torch.backends.cudnn.flags(enabled=True, benchmark=True, allow_tf32=False):
                old = torch.backends.cuda.matmul.allow_tf32
                torch.backends.cuda.matmul.allow_tf32 = False
                try:
                    return self._forward(tokens_u8, clocks)
                finally:
                    torch.backends.cuda.matmul.allow_tf32 = old
        return self._forward(tokens_u8, clocks)

    def _forward(self, tokens_u8, clocks):
        F = torch.nn.functional
        B = tokens_u8.shape[0]
        x = F.embedding(tokens_u8.long().view(B, 2, 6, 5), self.emb).permute(0, 1, 4, 2, 3).reshape(B, 8, 6, 5)
        x = x.to(self.dtype).contiguous(memory_format=torch.channels_last)
        x = F.relu(F.conv2d(x, self.stem[0], self.stem[1], padding=1))
        for (w1, b1), (w2, b2) in self.blocks:
            y = F.relu(F.conv2d(x, w1, b1, padding=1))
            x = F.relu(F.conv2d(y, w2, b2, padding=1) + x)
        x = x.float()
        clk = clocks.view(B, 1).float()
        px = F.relu(F.conv2d(x, self.pconv[0], self.pconv[1])).reshape(B, 60)
        logits = F.linear(torch.cat([px, clk], 1), *self.plinear)
        vx = F.relu(F.conv2d(x, self.vconv[0], self.vconv[1])).reshape(B, 30)
        v = torch.tanh(F.linear(F.relu(F.linear(torch.cat([vx, clk], 1), *self.v1)), *self.v2))
        return logits.contiguous(), v.reshape(B).contiguous()

    def capture(self, tokens_u8, clocks):
        """CUDA-graph the forward over fixed input buffers (the engine's leaf batch)."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self.forward(tokens_u8, clocks)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.forward(tokens_u8, clocks)
        self._graph, self._graph_out = g, out
        return out

    def replay(self):
        self._graph.replay()
        return self._graph_out


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
