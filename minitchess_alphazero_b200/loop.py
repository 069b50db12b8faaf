"""Full AlphaZero iteration on the engine (BASELINE.json configs[4]): self-play on every rank -> replay
gather -> learner update on the learner rank -> NCCL weight broadcast -> engines reload.

The learner step restates `SimpleAlphaZeroLearner.update` (exp/learner.py:72-94): a fresh AdamW per update,
shuffled mini-batches, BatchNorm in train mode, loss = mean((v - z)^2 - sum(pi * log_softmax(p)))
(exp/learner.py:89), fed by the device collate instead of the Python `collate_fn`.  The reference's own
learner class works unchanged on the same `Network`; this one only avoids its host round trips.
"""
import torch
import torch.distributed as dist

from . import parallel
from .policy import flatten_state_dict
from .selfplay import collate_device


def learner_update(network, tuples, batch_size=32, epochs=1, optim_params=None, device='cuda', generator=None):
    """One `update` over the given replay tuples.  Returns the list of mini-batch losses."""
    optim_params = optim_params or {'lr': 0.2}                       # app/learner.py:69
    pi, channels, clock, reward = collate_device(tuples, device=device)
    model = network.train().to(device)
    optimizer = torch.optim.AdamW(model.parameters(), **optim_params)  # exp/learner.py:73
    n = pi.shape[0]
    losses = []
    for _ in range(epochs):
        perm = torch.randperm(n, device=device, generator=generator)
        for i in range(0, n, batch_size):
            idx = perm[i:i + batch_size]
            if idx.numel() < 2:                                       # BatchNorm needs more than one sample
                continue
            p, v = model((channels[idx], clock[idx]))
            loss = ((v - reward[idx]) ** 2 - (pi[idx] * p.log_softmax(-1)).sum(1, keepdim=True)).mean()
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            losses.append(float(loss.detach()))
    network.eval()
    return losses


def iteration(selfplay, network, n_moves, learner_rank=0, max_tuples=None, **update_kwargs):
    """One loop iteration.  Works single-process (no process group) or under torchrun with NCCL."""
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size() if distributed else 1
    rank = dist.get_rank() if distributed else 0
    selfplay.run(n_moves)
    cap = max_tuples or selfplay.n_games * 64
    gathered, counts = parallel.gather_replay(selfplay.engine, world, cap)
    losses = []
    if rank == learner_rank:
        rows = torch.cat([gathered[r, :int(counts[r])] for r in range(world)]) if world > 1 else gathered[0, :int(counts[0])]
        if rows.shape[0] >= 2:
            losses = learner_update(network, rows, **update_kwargs)
    if distributed and world > 1:
        flat, _ = parallel.broadcast_weights(network, src=learner_rank)
    else:
        flat = flatten_state_dict(network.state_dict(), device='cuda')
    if selfplay.mode == 'builtin':
        selfplay.engine.set_weights(flat)                             # SimulatePuppet.load_weights
    else:
        if rank != learner_rank:
            parallel.load_flat_weights(network, flat)
        selfplay.sync_weights()
    return {'tuples': int(counts.sum()), 'losses': losses}
