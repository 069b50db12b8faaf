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

import torch
import torch.distributed as dist

from . import parallel
from .policy import flatten_state_dict
from .selfplay import collate_device


def learner_update(network, tuples, batch_size=32, epochs=1, optim_params=None, device='cuda', generator=None, max_batches=None,
                   order=None):
    """One `update` over the given replay tuples (packed az_replay_tuple records, or the four tensors `collate_fn` returns).
    Returns the list of mini-batch losses.  `order`: an explicit list of index lists (one per mini-batch) instead of a fresh
    shuffle -- the parity test replays the batches the reference's DataLoader drew."""
    optim_params = optim_params or {'lr': 0.2}                       # app/learner.py:69
    if isinstance(tuples, (list, tuple)) and len(tuples) == 4 and all(torch.is_tensor(t) for t in tuples):
        pi, channels, clock, reward = (t.to(device) for t in tuples)  # already collated (exp/learner.py:23-41 layout)
    else:
        pi, channels, clock, reward = collate_device(tuples, device=device)
    model = network.train().to(device)
    optimizer = torch.optim.AdamW(model.parameters(), **optim_params)  # exp/learner.py:73
    n = pi.shape[0]
    losses = []
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
            p, v = model((channels[idx], clock[idx]))
            loss = ((v - reward[idx]) ** 2 - (pi[idx] * p.log_softmax(-1)).sum(1, keepdim=True)).mean()
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            losses.append(float(loss.detach()))
    network.eval()
    return losses


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
