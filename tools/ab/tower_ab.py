"""A/B of libmcaz.so variants on one box: tower time per forward at several batch sizes, wait statistics, and a digest of the
network outputs (variants that only reschedule work must be bit-identical).  Each variant runs in its own process.
    python tools/ab/tower_ab.py tools/ab/libmcaz_a.so tools/ab/libmcaz_b.so ... [--rounds 2]
"""
import hashlib
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(so):
    sys.path.insert(0, REPO)
    import numpy as np
    import torch
    from minitchess_alphazero_b200 import _lib
    _lib.SO_PATH = so
    from minitchess_alphazero_b200.engine import Engine
    from minitchess_alphazero_b200.policy import Network, flatten_state_dict
    torch.manual_seed(0)
    flat = flatten_state_dict(Network().state_dict()).numpy()
    out = {'so': os.path.basename(so)}
    L = _lib.lib()
    rng = np.random.RandomState(0)
    tok_all = rng.randint(0, 7, size=(4096, 60)).astype(np.uint8)
    clk_all = rng.rand(4096).astype(np.float32)
    for rows in (2816, 4096):
        fp8 = int(os.environ.get('AB_FP8', '0'))
        eng = Engine(rows, max_sims_per_move=4, network=2, fp8_convolutions=fp8) if fp8 else Engine(rows, max_sims_per_move=4, network=1)
        eng.set_weights(flat)
        tok = torch.from_numpy(tok_all[:rows]).cuda(); clk = torch.from_numpy(clk_all[:rows]).cuda()
        lg = torch.empty(rows, 554, device='cuda'); vl = torch.empty(rows, device='cuda')

        def fwd():
            _lib.check(L.az_network_forward(eng._h, _lib.ptr(tok), _lib.ptr(clk), rows, _lib.ptr(lg), _lib.ptr(vl)))
        for _ in range(10):
            fwd()
        torch.cuda.synchronize()
        if rows == 2816:
            out['digest'] = hashlib.sha256(lg.cpu().numpy().tobytes() + vl.cpu().numpy().tobytes()).hexdigest()[:16]
        eng.profile_network(True, read=True)
        n = 400 if rows >= 2048 else 800
        for _ in range(n):
            fwd()
        ms, cnt, _ = eng.profile_network(False, read=True)
        out['tower_ms_%d' % rows] = round(ms, 4)
        out['fp8'] = fp8
        if os.environ.get('MCAZ_TOWER_STATS') and rows in (2816, 4096):
            import ctypes
            buf = np.zeros(148 * 12, dtype=np.uint64)
            L.az_tower_stats.restype = ctypes.c_int
            k = L.az_tower_stats(eng._h, _lib.ptr(buf), len(buf))
            if k > 0:
                st = buf[:k].reshape(-1, 12).astype(np.float64)
                lead = st[0::2]
                out['wait_%d' % rows] = {'operands': round(100 * lead[:, 1].sum() / lead[:, 0].sum(), 1),
                                         'accumulator': round(100 * lead[:, 2].sum() / lead[:, 0].sum(), 1),
                                         'producer_deps': round(100 * st[:, 4].sum() / st[:, 3].sum(), 1),
                                         'producer_slot': round(100 * st[:, 5].sum() / st[:, 3].sum(), 1),
                                         'kcycles': round(lead[:, 0].mean() / 1e3), 'cycles_per_item': round(lead[:, 0].sum() / max(lead[:, 7].sum(), 1))}
        eng.close()
    print('RESULT ' + json.dumps(out))


if __name__ == '__main__':
    if sys.argv[1] == '--worker':
        worker(sys.argv[2])
        sys.exit(0)
    rounds = 2
    sos = [a for a in sys.argv[1:] if not a.startswith('--')]
    if '--rounds' in sys.argv:
        rounds = int(sys.argv[sys.argv.index('--rounds') + 1])
        sos = [a for a in sos if a != str(rounds)]
    for r in range(rounds):                      # alternate the variants: the boxes drift with temperature
        for so in sos:
            env = dict(os.environ)
            if r == rounds - 1:
                env['MCAZ_TOWER_STATS'] = '1'    # the last round also reads the wait counters (they cost a little time)
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), '--worker', os.path.abspath(so)], capture_output=True, text=True,
                                   env=env, timeout=240)
            except subprocess.TimeoutExpired:
                print(json.dumps({'so': so, 'failed': 'timeout'}), flush=True)
                continue
            lines = [l for l in p.stdout.splitlines() if l.startswith('RESULT ')]
            print(lines[0][7:] if lines else json.dumps({'so': so, 'failed': p.returncode, 'stderr': p.stderr[-1500:]}), flush=True)
