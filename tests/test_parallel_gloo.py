"""CPU, world size 2 over gloo: game sharding, weight broadcast and replay gather host logic."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from minitchess_alphazero_b200 import parallel
    from minitchess_alphazero_b200.engine import REPLAY_DTYPE
    from minitchess_alphazero_b200.policy import Network, flatten_state_dict
    out = {}
    # sharding: disjoint cover
    out['shard'] = parallel.shard_of_games(11, rank, world).tolist()
    # weights: rank 0 is the learner
    torch.manual_seed(100 + rank)
    net = Network()
    flat, version = parallel.broadcast_weights(net, src=0, device='cpu', version=20261018)
    parallel.load_flat_weights(net, flat)
    out['w_sum'] = float(flatten_state_dict(net.state_dict()).double().sum())
    out['version'] = version
    # replay: different counts per rank
    tuples = np.zeros(8, dtype=REPLAY_DTYPE)
    n = 3 + 2 * rank
    tuples['action'][:n] = np.arange(n) + 100 * rank
    tuples['reward'][:n] = 1 - 2 * (np.arange(n) % 2)
    tuples['pi'][:n, 0] = 0.5 + rank
    local = torch.from_numpy(tuples.view(np.uint8).reshape(8, -1).copy())
    gathered, counts = parallel.gather_tuples(local, n)
    allt = parallel.unpack_gathered(gathered, counts)
    out['n_all'] = len(allt)
    out['actions'] = allt['action'].tolist()
    out['pi0'] = allt['pi'][:, 0].tolist()
    q.put((rank, out))
    dist.destroy_process_group()


def test_world2_gloo():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res[0]['shard'] + res[1]['shard']) == list(range(11))
    assert res[0]['w_sum'] == res[1]['w_sum'] and res[0]['version'] == res[1]['version'] == 20261018
    for r in (0, 1):
        assert res[r]['n_all'] == 3 + 5
        assert res[r]['actions'] == [0, 1, 2, 100, 101, 102, 103, 104]
        assert res[r]['pi0'] == [0.5] * 3 + [1.5] * 5
