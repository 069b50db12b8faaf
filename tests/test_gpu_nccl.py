"""Two GPUs, NCCL: the two exchanges of row (f)-2 / §8(e) -- the learner's weights broadcast with their version stamp
(app/base.py:171-174, :201-203 -> :31-39, :126-129) and the replay gather with the learner's stale-version drop
(app/base.py:63-70 -> app/learner.py:44-62, :51-53).  Skipped on boxes with fewer than two GPUs (the gloo test covers the
host logic on the CPU)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REPO

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from minitchess_alphazero_b200 import _lib, parallel
    from minitchess_alphazero_b200.policy import Network, flatten_state_dict
    from minitchess_alphazero_b200.selfplay import BatchedSelfPlay
    _lib.check(_lib.lib().mcaz_set_device(rank))
    out = {}
    torch.manual_seed(100 + rank)                       # different weights per rank until the learner's arrive
    net = Network().eval()
    flat, version = parallel.broadcast_weights(net, src=0, version=5)
    out['w_sum'] = float(flat.double().sum())
    out['version'] = version
    sp = BatchedSelfPlay(net, n_games=64, num_simulations=6, seed=rank)
    sp.sync_weights(flat, version=version)
    sp.run(35)
    if rank == 1:
        sp.sync_weights(flat, version=6)                # this actor already runs on other weights than the learner's
    sp.run(35)
    gathered, counts = parallel.gather_replay(sp.engine, world, 64 * 64)
    out['counts'] = counts.tolist()
    rows = parallel.valid_rows(gathered, counts)
    kept, stale = parallel.drop_stale(rows, 5)          # app/learner.py:51-53 on the learner
    allt = parallel.unpack_gathered(gathered, counts)
    out['n'] = len(allt)
    out['versions'] = {int(v): int((allt['weights_version'] == v).sum()) for v in np.unique(allt['weights_version'])}
    out['stale'], out['kept'] = stale, int(kept.shape[0])
    out['kept_versions'] = np.unique(kept.cpu().numpy().reshape(-1).view(parallel.REPLAY_DTYPE)['weights_version']).tolist()
    out['rank0_versions'] = np.unique(allt[:counts[0]]['weights_version']).tolist()
    out['rank1_versions'] = np.unique(allt[counts[0]:]['weights_version']).tolist()
    out['dropped_counter'] = sp.engine.counters()['replay_dropped']
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_nccl_weight_broadcast_and_versioned_replay_gather(mcaz_lib):
    import torch.multiprocessing as mp
    world, port = 2, 29600 + os.getpid() % 2000
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    a, b = res[0], res[1]
    assert a['w_sum'] == b['w_sum'] and a['version'] == b['version'] == 5        # everybody holds the learner's weights
    assert a['counts'] == b['counts'] and a['n'] == b['n'] == sum(a['counts']) and min(a['counts']) > 0
    assert a['versions'] == b['versions'] and set(a['versions']) == {5, 6}
    assert a['rank0_versions'] == [5] and a['rank1_versions'][-1] == 6           # rank 1's games that ended after its switch
    assert a['stale'] == a['versions'][6] > 0 and a['kept'] == a['versions'][5] and a['kept_versions'] == [5]
    assert a['dropped_counter'] == b['dropped_counter'] == 0
