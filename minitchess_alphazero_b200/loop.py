"""Full AlphaZero iteration on the engine (BASELINE.json configs[4]): self-play on every rank -> replay
gather -> learner update on the learner rank -> NCCL weight broadcast -> engines reload.

The learner step restates `SimpleAlphaZeroLearner.update` (exp/learner.py:72-94): a fresh AdamW per update,
shuffled mini-batches, BatchNorm in train mode, loss = mean((v - z)^2 - sum(pi * log_softmax(p)))
(exp/learner.py:89), fed by the device collate instead of the Python `collate_fn`.  The reference's own
learner class works unchanged on the same `Network`; this one only avoids its host round trips.

Rounding: the reference subtracts a (B,) vector from a (B,1) column, so its `.mean()` runs over a B x B matrix
(exp/learner.py:89); that is mean(a) - mean(c) exactly in real arithmetic and to within float32 rounding here
(`keepdim=True`, a B x 1 column).  tests/test_learner_golden.py compares losses and weights with the reference's own
update at a stated tolerance.

Versions: every update re-stamps the weights (LearnPuppet.weights, app/base.py:171-174), the stamp travels with the
broadcast, finished games carry the stamp of the weights they were played with, and the learner drops tuples with another
stamp (app/learner.py:51-53).  Optional gate: the arena of exp/learner.py:97-145 with the 0.55 threshold of
app/base.py:195-196 (commented out in the reference, hence off by default).
"""
import copy
import time
import weakref

import torch
import torch.distributed as dist

from . import parallel
from .policy import flatten_state_dict
from .selfplay import collate_device


_GRAPHED = weakref.WeakKeyDictionary()      # network -> (key, optimizer, _GraphedStep): the captured step is kept between updates


class _GraphedStep:
    """One optimiser step (forward, loss, backward, AdamW) of `learner_update` captured as a CUDA graph over static batch buffers.
    At batch 32 the step is ~200 small kernels and bound by their launches from Python; replayed as one graph it costs its GPU time.
    The three warm-up steps PyTorch asks for before a capture run on the real parameters, so parameters, BatchNorm buffers and
    the optimiser state are put back afterwards: the update takes exactly the steps the eager loop takes."""

    def __init__(self, model, optimizer, channels, clock, pi, reward, batch_size):
        B = batch_size
        self.channels = channels.new_zeros((B,) + tuple(channels.shape[1:]))
        self.clock = clock.new_zeros((B,) + tuple(clock.shape[1:]))
        self.pi = pi.new_zeros((B,) + tuple(pi.shape[1:]))
        self.reward = reward.new_zeros((B,) + tuple(reward.shape[1:]))
        self.load(channels, clock, pi, reward, torch.arange(B, device=channels.device))
        tensors = list(model.parameters()) + list(model.buffers())
        saved = [t.detach().clone() for t in tensors]
        # parameters that took an eager step before carry gradient-accumulation nodes of the default stream: expected here
        quiet = getattr(torch.autograd.graph, 'set_warn_on_accumulate_grad_stream_mismatch', None)
        if quiet is not None:
            quiet(False)
        side = torch.cuda.Stream()                                      # warm-up and capture on one stream
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                optimizer.zero_grad(set_to_none=True)
                self._loss(model).backward()
                optimizer.step()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph, stream=side):
            self.loss = self._loss(model)
            self.loss.backward()
            optimizer.step()
        with torch.no_grad():
            for t, v in zip(tensors, saved):
                t.copy_(v)
        _reset(optimizer)

    def _loss(self, model):
        p, v = model((self.channels, self.clock))
        return ((v - self.reward) ** 2 - (self.pi * p.log_softmax(-1)).sum(1, keepdim=True)).mean()     # exp/learner.py:89

    def load(self, channels, clock, pi, reward, idx):
        torch.index_select(channels, 0, idx, out=self.channels)
        torch.index_select(clock, 0, idx, out=self.clock)
        torch.index_select(pi, 0, idx, out=self.pi)
        torch.index_select(reward, 0, idx, out=self.reward)

    def step(self, channels, clock, pi, reward, idx):
        self.load(channels, clock, pi, reward, idx)
        self.graph.replay()
        return self.loss.detach().clone()


def _reset(optimizer):
    """The optimiser as `torch.optim.AdamW(...)` leaves it before its first step (exp/learner.py:73 builds a new one per update)."""
    with torch.no_grad():
        for state in optimizer.state.values():
            for v in state.values():
                if torch.is_tensor(v):
                    v.zero_()


def learner_update(network, tuples, batch_size=32, epochs=1, optim_params=None, device='cuda', generator=None, max_batches=None,
                   order=None, graph=None):
    """One `update` over the given replay tuples (packed az_replay_tuple records, or the four tensors `collate_fn` returns).
    Returns the list of mini-batch losses.  `order`: an explicit list of index lists (one per mini-batch) instead of a fresh
    shuffle -- the parity test replays the batches the reference's DataLoader drew.  `graph`: replay the step of full
    mini-batches as one CUDA graph (default: on a CUDA device when the update has at least 16 of them); short batches and
    `graph=False` take the same step launched kernel by kernel.  The captured step and its optimiser stay with the network
    between updates (same parameters, batch size and optimiser arguments): a later update starts from a zeroed optimiser
    state, which is what a new AdamW starts from."""
    optim_params = dict(optim_params or {'lr': 0.2})                 # app/learner.py:69
    if isinstance(tuples, (list, tuple)) and len(tuples) == 4 and all(torch.is_tensor(t) for t in tuples):
        pi, channels, clock, reward = (t.to(device) for t in tuples)  # already collated (exp/learner.py:23-41 layout)
    else:
        pi, channels, clock, reward = collate_device(tuples, device=device)
    model = network.train().to(device)
    n = pi.shape[0]
    on_cuda = torch.device(device).type == 'cuda'
    planned = (len(order) if order is not None else epochs * (n // batch_size))
    if max_batches is not None:
        planned = min(planned, max_batches)
    if graph is None:
        graph = planned >= 16
    graph = bool(graph) and on_cuda
    optimizer = graphed = key = None
    if graph:
        optim_params.setdefault('capturable', True)
        key = (batch_size, repr(sorted(optim_params.items())), tuple(p.data_ptr() for p in model.parameters()))
        kept = _GRAPHED.get(network)
        if kept is not None and kept[0] == key:
            optimizer, graphed = kept[1], kept[2]
            _reset(optimizer)
    if optimizer is None:
        optimizer = torch.optim.AdamW(model.parameters(), **optim_params)  # exp/learner.py:73
    losses = []                                                       # device scalars: one read-back at the end, not one per step
    for _ in range(epochs):
        if order is None:
            perm = torch.randperm(n, device=device, generator=generator)
            batches = [perm[i:i + batch_size] for i in range(0, n, batch_size)]
        else:
            batches = [torch.as_tensor(b, device=device, dtype=torch.long) for b in order]
        for idx in batches:
            if idx.numel() < 2:                                       # BatchNorm needs more than one sample
                continue
            if max_batches is not None and len(losses) >= max_batches:
                break
            if graph and idx.numel() == batch_size:
                if graphed is None:
                    graphed = _GraphedStep(model, optimizer, channels, clock, pi, reward, batch_size)
                    _GRAPHED[network] = (key, optimizer, graphed)
                losses.append(graphed.step(channels, clock, pi, reward, idx))
                continue
            p, v = model((channels[idx], clock[idx]))
            loss = ((v - reward[idx]) ** 2 - (pi[idx] * p.log_softmax(-1)).sum(1, keepdim=True)).mean()
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            losses.append(loss.detach())
    network.eval()
    return torch.stack(losses).tolist() if losses else []


def iteration(selfplay, network, n_moves, learner_rank=0, max_tuples=None, arena_games_per_side=0, arena_simulations=None,
              **update_kwargs):
    """One loop iteration.  Works single-process (no process group) or under torchrun with NCCL.

    Returns {'tuples' gathered, 'stale' dropped for their weights version, 'used', 'losses', 'accepted', 'version',
    'seconds': {'selfplay', 'gather', 'learner', 'arena', 'broadcast'}}."""
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size() if distributed else 1
    rank = dist.get_rank() if distributed else 0
    version = int(getattr(selfplay, 'weights_version', 1))
    sec = {}

    def tick():
        torch.cuda.synchronize()
        return time.perf_counter()
    t0 = tick()
    selfplay.run(n_moves)
    t1 = tick(); sec['selfplay'] = t1 - t0
    cap = max_tuples or selfplay.n_games * 64
    gathered, counts = parallel.gather_replay(selfplay.engine, world, cap)
    t2 = tick(); sec['gather'] = t2 - t1
    losses, stale, used, accepted, arena = [], 0, 0, True, None
    if rank == learner_rank:
        rows, stale = parallel.drop_stale(parallel.valid_rows(gathered, counts), version)    # app/learner.py:51-53
        used = int(rows.shape[0])
        old = copy.deepcopy(network.state_dict()) if arena_games_per_side > 0 else None
        if used >= 2:
            losses = learner_update(network, rows, **update_kwargs)
        t3 = tick(); sec['learner'] = t3 - t2
        if arena_games_per_side > 0 and losses:
            from .arena import passes_gate
            from .policy import Network
            old_net = Network().eval()
            old_net.load_state_dict(old)
            accepted, arena = passes_gate(network, old_net, games_per_side=arena_games_per_side,
                                          num_simulations=arena_simulations or selfplay.num_simulations)
            if not accepted:
                network.load_state_dict(old)                             # app/base.py:195-197: the old weights stay
        sec['arena'] = tick() - t3
    t4 = tick()
    # LearnPuppet.weights setter (app/base.py:171-174): accepted weights get a new stamp
    new_version = version + 1 if (rank == learner_rank and accepted and losses) else version
    if distributed and world > 1:
        flat, new_version = parallel.broadcast_weights(network, src=learner_rank, version=new_version)
    else:
        flat = flatten_state_dict(network.state_dict(), device='cuda')
    selfplay.sync_weights(flat, version=new_version)                      # SimulatePuppet.load_weights
    sec['broadcast'] = tick() - t4
    return {'tuples': int(counts.sum()), 'stale': stale, 'used': used, 'losses': losses, 'accepted': bool(accepted),
            'version': new_version, 'arena': arena, 'seconds': sec}
