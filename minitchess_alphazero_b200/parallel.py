"""Multi-GPU plumbing: one process per GPU, games sharded by id, `torch.distributed` (NCCL over
NVLink) only for the two exchanges the reference does over MQTT/HTTP:

  * weights:  learner -> actors.  One flat float32 buffer broadcast from the learner rank; replaces
              LearnPuppet.get_weights_dict + the flask relay + download_weights
              (app/base.py:201-203, app/web.py:15-30, app/base.py:31-39).
  * replay:   actors -> learner.  Packed tuples (each stamped with the version of the weights its game was played with)
              all-gathered per move step, counts first; replaces MQTTDataset.push -> on_message -> push_data
              (app/base.py:52-70, app/learner.py:44-62) including the learner's stale-version drop (:51-53).

Self-play itself has no data-path collective: trees never interact.
The functions take CPU tensors too, so the host logic is tested with gloo at world size 2.
"""
import numpy as np
import torch
import torch.distributed as dist

from ._lib import AZ_NUM_WEIGHT_FLOATS
from .engine import REPLAY_DTYPE
from .policy import flatten_state_dict

TUPLE_BYTES = REPLAY_DTYPE.itemsize


def shard_of_games(n_games_total, rank, world):
    """Game g lives on rank g mod world."""
    return np.arange(rank, n_games_total, world, dtype=np.int64)


def broadcast_weights(network, src=0, device=None, version=None):
    """Broadcast the learner's weights; every rank returns (flat float32 tensor, version)."""
    rank = dist.get_rank()
    dev = device if device is not None else ('cuda' if dist.get_backend() == 'nccl' else 'cpu')
    if rank == src:
        flat = flatten_state_dict(network.state_dict(), device=dev)
    else:
        flat = torch.empty(AZ_NUM_WEIGHT_FLOATS, dtype=torch.float32, device=dev)
    dist.broadcast(flat, src=src)
    stamp = torch.tensor([int(version or 0)], dtype=torch.int64, device=dev)
    dist.broadcast(stamp, src=src)
    return flat, int(stamp.item())


def load_flat_weights(network, flat):
    """Inverse of flatten_state_dict: write a flat buffer back into a torch Network (load_state_dict)."""
    sd = network.state_dict()
    out, off = {}, 0
    flat = flat.detach().cpu()
    for k, v in sd.items():
        if k.endswith('num_batches_tracked'):
            out[k] = v
            continue
        n = v.numel()
        out[k] = flat[off:off + n].view_as(v).to(v.dtype)
        off += n
    assert off == flat.numel()
    network.load_state_dict(out)


def gather_tuples(local, count, world=None):
    """all_gather of this rank's first `count` tuple rows.  local: uint8 [capacity, TUPLE_BYTES].  The counts are gathered
    first, then only max(count) rows per rank travel (a fixed-size gather of the whole buffer moved 2 * n_games rows per rank
    and step whatever had finished).  Returns (gathered uint8 [world, max(count), TUPLE_BYTES], counts int64 [world])."""
    world = world or (dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1)
    cnt = torch.tensor([int(count)], dtype=torch.int64, device=local.device)
    if world == 1:
        return local[:int(count)].unsqueeze(0), cnt
    nccl = dist.get_backend() == 'nccl'
    counts = torch.empty(world, dtype=torch.int64, device=local.device)
    if nccl:
        dist.all_gather_into_tensor(counts, cnt)
    else:
        dist.all_gather(list(counts.split(1)), cnt)
    rows = int(counts.max())
    assert rows <= local.shape[0], (rows, local.shape)
    gathered = torch.empty((world, rows) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if rows:
        part = local[:rows].contiguous()
        if nccl:
            dist.all_gather_into_tensor(gathered.view(-1), part.reshape(-1))
        else:
            dist.all_gather(list(gathered.unbind(0)), part)
    return gathered, counts


_replay_buffers = {}


def gather_replay(engine, world, max_tuples):
    """Drain up to `max_tuples` of this rank's finished-game tuples into a device buffer (what does not fit stays queued in
    the engine for the next call) and all_gather them."""
    import ctypes
    from ._lib import ptr
    key = (id(engine), max_tuples)
    if key not in _replay_buffers:
        _replay_buffers.clear()                       # one engine at a time per process: drop a closed engine's buffer
        _replay_buffers[key] = torch.zeros(max_tuples, TUPLE_BYTES, dtype=torch.uint8, device='cuda')
    buf = _replay_buffers[key]
    n = ctypes.c_int()
    engine._check(engine._L.az_drain_replay(engine._h, ptr(buf), max_tuples, ctypes.byref(n)))
    return gather_tuples(buf, n.value, world)


def valid_rows(gathered, counts):
    """The valid rows of a gather as one uint8 tensor [sum(counts), TUPLE_BYTES] (device of `gathered`)."""
    c = [int(x) for x in counts.tolist()]
    parts = [gathered[r, :c[r]] for r in range(len(c)) if c[r]]
    return torch.cat(parts) if parts else gathered.new_zeros((0, TUPLE_BYTES))


VERSION_OFFSET = REPLAY_DTYPE.fields['weights_version'][1]


def drop_stale(rows, version):
    """app/learner.py:51-53: episodes played with other weights than the learner's current ones are not used.  rows: uint8
    [n, TUPLE_BYTES] (torch, any device) or a REPLAY_DTYPE array.  Returns (kept rows, number dropped)."""
    version = int(version) & 0xffffffff
    if isinstance(rows, np.ndarray):
        keep = rows['weights_version'] == version
        return rows[keep], int((~keep).sum())
    if rows.shape[0] == 0:
        return rows, 0
    stamp = rows[:, VERSION_OFFSET:VERSION_OFFSET + 4].contiguous().view(torch.int32).view(-1)
    want = version - (1 << 32) if version >= (1 << 31) else version
    keep = stamp == want
    return rows[keep], int((~keep).sum())


def unpack_gathered(gathered, counts):
    """-> numpy REPLAY_DTYPE array of all valid tuples (learner side)."""
    g = gathered.cpu().numpy()
    c = counts.cpu().numpy()
    parts = [g[r, :c[r]].reshape(-1).view(REPLAY_DTYPE) for r in range(len(c))]
    return np.concatenate(parts) if parts else np.zeros(0, dtype=REPLAY_DTYPE)
