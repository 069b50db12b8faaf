#!/usr/bin/env python
"""bench.py -- MCTS simulations/s of batched MinitChess AlphaZero self-play on B200.

Contract:  python bench.py --gpus N --steps K --warmup W
  * workload = BASELINE.json configs[2]: 4096 concurrent games x 200 simulations/move per GPU,
    random-init reference-size network (seed 0), synthetic self-play.  The games are spread uniformly
    over the plies of a game by a low-simulation pre-roll before the warm-up (the population a
    long-running actor holds), not started together from STARTING_FEN.
  * one "step" = one move in every game: 200 simulations (select/expand -> network -> backup)
    for each of the 4096 active trees, then move choice, replay recording, play, restarts.
    A simulation whose leaf is a finished position, or a position the exact evaluation cache has
    seen (same board, side, fullmove number = everything the network reads), needs no network row;
    `evals_per_second` and `sims_breakdown` say how many did, `without_cache_lockstep` is the same
    workload with the cache off, and `roofline` is computed from the rows actually evaluated.
  * `value` = simulations/s of the whole job, everything resident in HBM (device RNG, device move
    choice).  `e2e` = the same metric through the host-buffer API: per step the positions go up
    from pinned host memory, root statistics come back, the host samples the moves.
  * --impl reference : the reference's CPU self-play path (oracle port of exp/agent.py +
    exp/policy.py + exp/environment.py) on the host cores, same sims/move.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

FLOP_PER_EVAL = 638245892           # 2 x 319 122 946 MAC per leaf evaluation
METRIC = 'mcts_simulations_per_second'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--games', type=int, default=4096, help='concurrent games per GPU')
    ap.add_argument('--sims', type=int, default=200, help='simulations per move')
    ap.add_argument('--cpu-seconds', type=float, default=15.0, help='budget of the CPU baseline sample')
    ap.add_argument('--mode', default='lockstep', choices=['continuous', 'lockstep'],
                    help='continuous: az_selfplay, every game moves as soon as its own simulations are done (a step = '
                         'sims-per-move network batches); lockstep: az_search + az_play_device, one move in every game per step')
    ap.add_argument('--eval-cache', type=int, default=24, help='log2 entries of the exact evaluation cache (0 = off)')
    ap.add_argument('--free-sims', type=int, default=0, help='descents per game and launch (0 = library default)')
    ap.add_argument('--defer-rows', type=int, default=-1, help='az_config.defer_rows (a short last tile pair of a batch waits for the '
                    'next batch); -1 = the default of BatchedSelfPlay (192), 0 = off')
    ap.add_argument('--no-stagger', action='store_true', help='start all games from the start position instead of spreading '
                    'them over the plies of a game (see BatchedSelfPlay.stagger)')
    ap.add_argument('--no-plain', action='store_true', help='skip the comparison pass without cache / continuous mode')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-fp8', action='store_true', help='skip the leg that runs the same workload on the opt-in e4m3 tower')
    ap.add_argument('--no-legs', action='store_true', help='N > 1 only: skip the config4 (32768 games x 800 sims/move sharded) and '
                    'loop (self-play + gather + learner step + weight broadcast) legs')
    ap.add_argument('--loop-moves', type=int, default=6, help='moves of self-play per loop iteration in the loop leg')
    return ap.parse_args()


def peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'tflops': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'which': 'measured (sustained bf16)'}
    return {'hbm_gbs': 6650.0, 'tflops': 1400.0, 'which': 'fallback'}


def tower_traffic(rows):
    """DRAM bytes of one tower launch from the committed ncu captures: the entry whose batch size is nearest to `rows`."""
    path = os.path.join(REPO, 'profiles', 'tower_traffic.json')
    if not os.path.exists(path):
        return None, None
    table = json.load(open(path)).get('fused', [])
    if not table:
        return None, None
    best = min(table, key=lambda e: abs(e['rows'] - rows))
    return best['read_bytes'] + best['write_bytes'], 'ncu dram__bytes_read.sum + dram__bytes_write.sum of a %d-row launch (%s); this run averaged %.0f rows per launch' % (
        best['rows'], best['source'], rows)


# --------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            parts = [x.strip() for x in r.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), parts[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------- CPU baseline (oracle port)
def _cpu_worker(args):
    seed, sims, seconds, threads = args
    import numpy as np
    import torch
    torch.set_num_threads(threads)
    from minitchess_alphazero_b200.policy import Network
    from oracle import ref_selfplay as rs
    torch.manual_seed(0)
    net = rs.RefNetwork(Network().eval().state_dict())
    rng = np.random.RandomState(seed)
    trees = [rs.RefTree(net.evaluate, 1, rng=rng) for _ in range(2)]
    ep = rs.RefEpisode(rs.STARTING_FEN)
    t0 = time.perf_counter()
    n_sims = n_moves = 0
    turn = 0
    while time.perf_counter() - t0 < seconds:
        if ep.done:
            trees = [rs.RefTree(net.evaluate, 1, rng=rng) for _ in range(2)]
            ep = rs.RefEpisode(rs.STARTING_FEN)
            turn = 0
        action, _ = rs.ref_select_action(trees[turn], ep.fen, sims, rng=rng)
        ep.step(action)
        n_sims += sims
        n_moves += 1
        turn ^= 1
    dt = time.perf_counter() - t0
    return n_sims, n_moves, dt


def cpu_reference_sample(sims, seconds, processes, threads, pool=None):
    """Runs the oracle port of the reference self-play on `processes` host processes for ~`seconds`.  `pool`: a pool of that many
    spawned workers kept between samples (the reference arm: its workers import torch once, not once per step)."""
    import multiprocessing as mp
    jobs = [(i, sims, seconds, threads) for i in range(processes)]
    if pool is not None:
        res = pool.map(_cpu_worker, jobs, chunksize=1)
    elif processes == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context('spawn').Pool(processes) as own:
            res = own.map(_cpu_worker, jobs, chunksize=1)
    total_sims = sum(r[0] for r in res)
    wall = max(r[2] for r in res)
    return total_sims / wall, sum(r[1] for r in res) / wall, wall


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    # about two moves per process per step, shortened so that warm-up + steps stay within ~3 minutes of samples
    per_step = min(max(2.0, 2.0 * args.sims * 0.017), 20.0)
    per_step = max(2.0, min(per_step, 170.0 / max(1, args.steps + args.warmup)))
    import multiprocessing as mp
    pool = mp.get_context('spawn').Pool(procs) if procs > 1 else None
    try:
        for _ in range(args.warmup):
            cpu_reference_sample(args.sims, min(per_step, 3.0), procs, 1, pool)
        t0 = time.perf_counter()
        sims_s = []
        for _ in range(args.steps):
            s, _, _ = cpu_reference_sample(args.sims, per_step, procs, 1, pool)
            sims_s.append(s)
        dt = time.perf_counter() - t0
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    value = sum(sims_s) / len(sims_s)
    sample = '%d processes x 1 torch thread, each playing self-play moves at %d sims/move for %.1f s per step' % (procs, args.sims, per_step)
    emit(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'sims/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1000 * dt / max(args.steps, 1), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'reference self-play path (oracle port of exp/agent.py + exp/policy.py + exp/environment.py), '
                               '%d sims/move, random-init net' % args.sims, 'sims_per_move': args.sims},
        'cpu_baseline': {'value': value, 'unit': 'sims/s', 'cores': procs, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'sims/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the engine has no CPU fallback)')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from minitchess_alphazero_b200 import build, _lib
    build.build()
    _lib.check(_lib.lib().mcaz_set_device(local))
    from minitchess_alphazero_b200.policy import Network
    from minitchess_alphazero_b200.selfplay import BatchedSelfPlay
    from minitchess_alphazero_b200.parallel import gather_replay

    G, S = args.games, args.sims
    torch.manual_seed(0)
    net = Network().eval()
    builtin = True
    opts = {'eval_cache_log2': args.eval_cache, 'free_sims': args.free_sims}
    if args.defer_rows >= 0:
        opts['defer_rows'] = args.defer_rows
    sp = BatchedSelfPlay(net, n_games=G, num_simulations=S, seed=1234 + rank, **opts)
    eng = sp.engine
    continuous = args.mode == 'continuous'

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        if continuous:
            sp.run_continuous(S)        # S network batches: one move's worth of simulations for a game that hits nothing
        else:
            sp.step()
        if world > 1:
            gather_replay(eng, world, max_tuples=2 * G)     # replay gather of config 4 (NCCL all_gather)

    # population: games spread uniformly over plies 0..59 (what a long-running actor holds), not 4096 copies of the
    # start position marching through the opening together -- with the evaluation cache the latter would measure
    # mostly cache hits.  Low-simulation pre-roll, outside the timed region.
    if builtin and not args.no_stagger:
        sp.stagger()
    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    c0 = eng.counters()
    l0 = _lib.lib().mcaz_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    net_ms = sp.reset_kernel_timer() if hasattr(sp, 'reset_kernel_timer') else None
    ev0.record()
    for _ in range(args.steps):
        one_step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    c1 = eng.counters()
    launches = int(_lib.lib().mcaz_kernel_launches() - l0)
    sims = c1['simulations'] - c0['simulations']
    evals = c1['evaluations'] - c0['evaluations']
    moves = c1['moves'] - c0['moves']
    cached = c1['cached_evaluations'] - c0['cached_evaluations']
    terminal = c1['terminal_leaves'] - c0['terminal_leaves']
    dup = c1['duplicate_rows'] - c0['duplicate_rows']
    evicted = c1['evicted_nodes']             # since creation: nodes the second-stage compaction dropped (0 = every tree is the reference's)
    tot = torch.tensor([sims, evals, moves, cached, terminal, dup, evicted], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tot)
    sims_all, evals_all, moves_all, cached_all, terminal_all, dup_all, evicted_all = (float(x) for x in tot.tolist())
    value = sims_all / (ms / 1000.0)

    # roofline of the dominant kernel: the network tower (tensor-bound), from the rows it actually evaluated
    pk = peaks()
    prof = sp.kernel_profile(evaluations=evals) if builtin else None
    tree = sp.tree_profile(c0, c1) if builtin else None
    if tree is not None:
        tree.update({'peak': pk['hbm_gbs'], 'frac': tree['achieved'] / pk['hbm_gbs'],
                     'note': 'latency-bound pointer chasing: one warp walks one tree; see profiles/ for warp efficiency'})
    if prof is None:
        roof = None
    else:
        roof = prof
        roof['traffic'], roof['traffic_source'] = tower_traffic(roof['rows_per_launch'])
        roof.update({'peak': pk['tflops'], 'frac': roof['achieved'] / pk['tflops'], 'mma_frac': roof['achieved_mma'] / pk['tflops'],
                     'peak_source': pk['which']})

    # end to end through the host-buffer API
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(sp, args, world)

    # the same workload without the evaluation cache, in lock-step (every simulation that is not terminal takes a
    # network row): what the cache and the continuous schedule buy, reported beside the headline
    plain = None
    if builtin and not args.no_plain and (args.eval_cache > 0 or continuous):
        eng.close()                     # frees the first engine's arenas
        sp2 = BatchedSelfPlay(net, n_games=G, num_simulations=S, seed=1234 + rank, eval_cache_log2=0)
        if not args.no_stagger:
            sp2.stagger()
        for _ in range(min(args.warmup, 3)):
            sp2.step()
        barrier()
        p0 = sp2.engine.counters()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        n_plain = max(1, min(args.steps, 2))
        for _ in range(n_plain):
            sp2.step()
        t1.record()
        barrier()
        p1 = sp2.engine.counters()
        pms = t0.elapsed_time(t1)
        pt = torch.tensor([pms, float(p1['simulations'] - p0['simulations']), float(p1['evaluations'] - p0['evaluations'])],
                          dtype=torch.float64, device='cuda')
        if world > 1:
            pmax = pt.clone()
            dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(pt)
            pt[0] = pmax[0]
        plain = {'value': float(pt[1]) / (float(pt[0]) / 1000.0), 'unit': 'sims/s', 'evals_per_second': float(pt[2]) / (float(pt[0]) / 1000.0),
                 'steps': n_plain, 'mode': 'lockstep, eval_cache off'}
        sp2.engine.close()

    # the engine's own production loop beside the lock-step headline: continuous self-play (az_selfplay: every game searches,
    # moves, records and restarts on its own inside the search kernel) with up to four chained simulations per launch
    cont = None
    if builtin and not args.no_plain and not continuous:
        sp3 = BatchedSelfPlay(net, n_games=G, num_simulations=S, seed=1234 + rank, eval_cache_log2=args.eval_cache,
                              free_sims=4)
        if not args.no_stagger:
            sp3.stagger()
        for _ in range(2):
            sp3.run_continuous(S)
        barrier()
        q0 = sp3.engine.counters()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        n_cont = max(1, min(args.steps, 2))
        for _ in range(n_cont):
            sp3.run_continuous(S)
        t1.record()
        barrier()
        q1 = sp3.engine.counters()
        ct = torch.tensor([t0.elapsed_time(t1), float(q1['simulations'] - q0['simulations']), float(q1['evaluations'] - q0['evaluations'])],
                          dtype=torch.float64, device='cuda')
        if world > 1:
            cmax = ct.clone()
            dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(ct)
            ct[0] = cmax[0]
        cont = {'value': float(ct[1]) / (float(ct[0]) / 1000.0), 'unit': 'sims/s', 'evals_per_second': float(ct[2]) / (float(ct[0]) / 1000.0),
                'steps': n_cont, 'mode': 'continuous (az_selfplay), free_sims 4, eval_cache as the headline; a step = %d network batches' % S}
        sp3.engine.close()

    # the opt-in e4m3 tower (az_config.network = 2) on the same workload: reported beside the bf16 headline, never as it
    fp8 = None
    if not args.no_plain and not args.no_fp8:
        sp4 = BatchedSelfPlay(net, n_games=G, num_simulations=S, seed=1234 + rank, precision='fp8', eval_cache_log2=args.eval_cache,
                              free_sims=args.free_sims)
        if not args.no_stagger:
            sp4.stagger()
        for _ in range(2):
            sp4.step()
        barrier()
        f0 = sp4.engine.counters()
        sp4.reset_kernel_timer()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        n_fp8 = max(1, min(args.steps, 3))
        for _ in range(n_fp8):
            sp4.step()
        t1.record()
        barrier()
        f1 = sp4.engine.counters()
        ft = torch.tensor([t0.elapsed_time(t1), float(f1['simulations'] - f0['simulations']), float(f1['evaluations'] - f0['evaluations'])],
                          dtype=torch.float64, device='cuda')
        fprof = sp4.kernel_profile(evaluations=f1['evaluations'] - f0['evaluations'])
        if world > 1:
            fmax = ft.clone()
            dist.all_reduce(fmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(ft)
            ft[0] = fmax[0]
        fp8_peak = 2.0 * pk['tflops']
        fp8 = {'value': float(ft[1]) / (float(ft[0]) / 1000.0), 'unit': 'sims/s', 'evals_per_second': float(ft[2]) / (float(ft[0]) / 1000.0),
               'steps': n_fp8, 'dtype': 'e4m3 operands in the first 12 of the 18 tower convolutions, bf16 in the last 6 (fp32 accumulate, bf16 residual stream, fp32 heads) / f64 tree statistics',
               'tolerance': 'priors and values within 1e-2 of the fp32 reference network (tests/test_gpu_fp8.py)',
               'roofline': None if fprof is None else {
                   'bound': 'tensor', 'kernel': 'tower_tc_kernel<FP8> (tcgen05.mma kind::f8f6f4)', 'achieved': fprof['achieved'], 'unit': 'TFLOP/s',
                   'ms_per_launch': fprof['ms_per_launch'], 'rows_per_launch': fprof['rows_per_launch'], 'peak': fp8_peak,
                   'frac': fprof['achieved'] / fp8_peak, 'mma_frac': fprof['achieved_mma'] / fp8_peak,
                   'peak_source': 'fallback: 2 x the measured sustained bf16 rate (MEASURED_PEAKS.json holds no fp8 figure; nominal dense fp8 is twice bf16)'}}
        sp4.engine.close()
        # all 18 convolutions on e4m3 (looser on the BatchNorm-perturbed stress network: tests/test_gpu_fp8.py): one more number
        sp5 = BatchedSelfPlay(net, n_games=G, num_simulations=S, seed=1234 + rank, precision='fp8', fp8_convolutions=18,
                              eval_cache_log2=args.eval_cache, free_sims=args.free_sims)
        if not args.no_stagger:
            sp5.stagger()
        sp5.step()
        barrier()
        h0 = sp5.engine.counters()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(2):
            sp5.step()
        t1.record()
        barrier()
        h1 = sp5.engine.counters()
        ht = torch.tensor([t0.elapsed_time(t1), float(h1['simulations'] - h0['simulations']), float(h1['evaluations'] - h0['evaluations'])],
                          dtype=torch.float64, device='cuda')
        if world > 1:
            hmax = ht.clone()
            dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(ht)
            ht[0] = hmax[0]
        fp8['fp8_convolutions'] = 12
        fp8['all_18_convolutions'] = {'value': float(ht[1]) / (float(ht[0]) / 1000.0), 'unit': 'sims/s',
                                      'evals_per_second': float(ht[2]) / (float(ht[0]) / 1000.0), 'steps': 2}
        sp5.engine.close()

    # N > 1: the two multi-GPU configurations of BASELINE.json beside the headline -- configs[3] at its full size sharded over
    # the ranks, and one iteration of the full loop of configs[4]
    config4 = loop_leg = None
    if world > 1 and not args.no_legs:
        eng.close()
        config4 = leg_config4(net, rank, world, args, barrier)
        loop_leg = leg_loop(net, rank, world, args, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = torch.get_num_threads()
        s, m, wall = cpu_reference_sample(36, args.cpu_seconds, 1, threads)
        cpu = {'value': s, 'unit': 'sims/s', 'cores': threads, 'kind': 'port',
               'sample': 'BASELINE.json configs[0]: one process, %d torch threads, self-play from STARTING_FEN at 36 sims/move for %.1f s '
                         '(%.2f positions/s); oracle/ref_selfplay.py' % (threads, wall, m)}
    dropin = None
    if rank == 0 and world == 1 and builtin and not args.no_cpu_baseline:
        dropin = measure_dropin()
    if rank == 0:
        out = {
            'metric': METRIC, 'value': value, 'unit': 'sims/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / max(args.steps, 1), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16 network / f64 tree statistics', 'data': 'synthetic',
            'config': {'workload': '%s: batched self-play, %d concurrent games x %d sims/move per GPU, random-init reference-size net' % (
                           'BASELINE.json configs[3] (32768 games x 800 sims/move sharded over the GPUs)' if (G * world == 32768 and S == 800)
                           else 'BASELINE.json configs[2]' if (G == 4096 and S == 200) else 'custom size', G, S),
                       'games_per_gpu': G, 'sims_per_move': S, 'evaluator': 'builtin (tower_tc_kernel)',
                       'mode': ('continuous (az_selfplay): a step = %d network batches; games move on their own' % S) if continuous
                               else 'lockstep (az_search + az_play_device): a step = one move in every game',
                       'eval_cache': ('exact, 2^%d entries' % args.eval_cache) if builtin and args.eval_cache > 0 else 'off',
                       'population': 'all games from the start position' if (args.no_stagger or not builtin) else
                                     'games spread uniformly over plies 0..59 by an 8-simulation pre-roll before the warm-up',
                       'parallelism': 'games sharded over %d GPU(s), no data-path collective except replay all_gather' % world,
                       'l2': 'inputs larger than L2 (tree arenas ~%d MB, activations %d MB per layer)' % (
                           int(c1['edges'] * 22 / 1e6), int(G * 30 * 256 * 2 * 2 / 1e6))},
            'positions_per_second': moves_all / (ms / 1000.0), 'evals_per_second': evals_all / (ms / 1000.0),
            'sims_breakdown': {'network_rows': evals_all / max(sims_all, 1), 'cache_hits': cached_all / max(sims_all, 1),
                               'terminal': terminal_all / max(sims_all, 1),
                               'rows_evaluated_twice_in_one_batch': dup_all / max(sims_all, 1),
                               'tree_nodes_evicted_by_the_inexact_compaction': int(evicted_all)},
            'roofline': roof, 'tree_roofline': tree, 'without_cache_lockstep': plain, 'continuous_selfplay': cont,
            'fp8_tower': fp8, 'config4': config4, 'loop': loop_leg,
            'cpu_baseline': cpu, 'dropin_config1': dropin, 'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks,
        }
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def leg_config4(net, rank, world, args, barrier, total_games=32768, sims=800, steps=2):
    """BASELINE.json configs[3]: 32768 concurrent games x 800 simulations per move sharded over the ranks (game g on rank
    g mod world: 16384 / 8192 / 4096 games per GPU at 2 / 4 / 8), the finished games' replay tuples all-gathered every step."""
    import torch
    import torch.distributed as dist
    from minitchess_alphazero_b200.parallel import gather_replay
    from minitchess_alphazero_b200.selfplay import BatchedSelfPlay
    G = total_games // world
    sp = BatchedSelfPlay(net, n_games=G, num_simulations=sims, seed=4321 + rank, eval_cache_log2=args.eval_cache)
    sp.stagger()
    gathered_tuples = 0

    def one_step():
        nonlocal gathered_tuples
        sp.step()
        _, counts = gather_replay(sp.engine, world, max_tuples=2 * G)
        gathered_tuples += int(counts.sum())
    one_step()
    barrier()
    gathered_tuples = 0
    c0 = sp.engine.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one_step()
    e1.record()
    barrier()
    c1 = sp.engine.counters()
    d = {k: c1[k] - c0[k] for k in c1}
    t = torch.tensor([e0.elapsed_time(e1), float(d['simulations']), float(d['evaluations']), float(d['cached_evaluations']),
                      float(d['terminal_leaves']), float(d['moves'])], dtype=torch.float64, device='cuda')
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t)
    ms, simsum = float(tmax[0]), float(t[1])
    sp.engine.close()
    return {'workload': 'BASELINE.json configs[3]: %d concurrent games x %d sims/move sharded over %d GPUs (%d per GPU), replay all_gather '
                        'every step (counts first, then max(count) rows)' % (total_games, sims, world, G),
            'value': simsum / (ms / 1e3), 'unit': 'sims/s', 'steps': steps, 'ms_per_step': ms / steps,
            'evals_per_second': float(t[2]) / (ms / 1e3), 'positions_per_second': float(t[5]) / (ms / 1e3),
            'sims_breakdown': {'network_rows': float(t[2]) / max(simsum, 1), 'cache_hits': float(t[3]) / max(simsum, 1),
                               'terminal': float(t[4]) / max(simsum, 1)},
            'replay_tuples_gathered': gathered_tuples, 'games_total': total_games, 'games_per_gpu': G, 'sims_per_move': sims}


def leg_loop(net, rank, world, args, barrier, learner_batches=256):
    """BASELINE.json configs[4]: one iteration of the full loop -- self-play on every GPU (4096 games x 200 sims/move each,
    `--loop-moves` moves), replay all_gather, the learner's update on rank 0 (exp/learner.py:72-94: AdamW lr 0.2, batch 32,
    app/learner.py:65-69; bounded to `learner_batches` mini-batches of the gathered tuples so that the bench stays short),
    NCCL weight broadcast with the version stamp, engines reload.  Seconds per iteration, max over ranks."""
    import torch
    import torch.distributed as dist
    from minitchess_alphazero_b200.loop import iteration
    from minitchess_alphazero_b200.selfplay import BatchedSelfPlay
    G, S = args.games, args.sims
    sp = BatchedSelfPlay(net, n_games=G, num_simulations=S, seed=777 + rank, eval_cache_log2=args.eval_cache)
    sp.stagger()
    kw = dict(batch_size=32, optim_params={'lr': 0.2}, max_batches=learner_batches)
    iteration(sp, net, 2, **dict(kw, max_batches=16))                # warm-up: every phase once (16 mini-batches: the learner's step is captured as a CUDA graph here and kept)
    barrier()
    c0 = sp.engine.counters()
    t0 = time.perf_counter()
    out = iteration(sp, net, args.loop_moves, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    c1 = sp.engine.counters()
    phases = ['selfplay', 'gather', 'learner', 'arena', 'broadcast']
    # the iteration's time is the slowest rank's; the phases are the learner rank's (the other ranks spend the learner's
    # update waiting in the broadcast)
    t = torch.tensor([dt] + [float(out['seconds'].get(k, 0.0)) for k in phases], dtype=torch.float64, device='cuda')
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.broadcast(t, src=0)
    t[0] = tmax[0]
    s = torch.tensor([float(c1['simulations'] - c0['simulations'])], dtype=torch.float64, device='cuda')
    dist.all_reduce(s)
    info = torch.tensor([float(out['tuples']), float(out['used']), float(out['stale']), float(len(out['losses'])), float(out['version'])],
                        dtype=torch.float64, device='cuda')
    dist.broadcast(info, src=0)
    sp.engine.close()
    sec = {k: float(v) for k, v in zip(phases, t[1:].tolist())}
    return {'workload': 'BASELINE.json configs[4]: self-play on %d GPUs (%d games x %d sims/move each, %d moves) -> replay all_gather -> '
                        'learner update on rank 0 (AdamW lr 0.2, batch 32, first %d mini-batches, each step one CUDA graph replay) -> NCCL weight broadcast -> engines reload'
                        % (world, G, S, args.loop_moves, learner_batches),
            'seconds_per_iteration': float(t[0]), 'seconds': sec, 'selfplay_sims_per_second': float(s[0]) / max(sec['selfplay'], 1e-9),
            'tuples_gathered': int(info[0]), 'tuples_used': int(info[1]), 'stale_dropped': int(info[2]), 'learner_steps': int(info[3]),
            'weights_version': int(info[4]), 'moves_per_iteration': args.loop_moves}


def measure_dropin(seconds=4.0, sims=36):
    """BASELINE.json configs[0] through the drop-in classes exactly as app/base.py:113-120 wires the reference's: one game at a
    time, two SimpleAlphaZeroAgent sharing one policy, RoundRobinReferee, per-episode init_mcts, 36 simulations per move,
    numpy RNG on the host.  Strictly sequential search (one tree, one leaf per network pass): this is the latency-bound
    use of the engine, reported beside the CPU baseline of the same configuration."""
    import numpy as np
    import torch
    from minitchess_alphazero_b200.agent import RoundRobinReferee, SimpleAlphaZeroAgent
    from minitchess_alphazero_b200.environment import MinitChessEnvironment
    from minitchess_alphazero_b200.erlyx_compat import BaseCallback, run_episodes
    from minitchess_alphazero_b200.policy import Network, SimpleAlphaZeroPolicy

    class Init(BaseCallback):                                   # exp/callbacks.py:57-62
        def __init__(self, agent):
            self.agent = agent

        def on_episode_begin(self, initial_observation):
            self.agent.init_mcts()

    class Count(BaseCallback):
        plies = 0

        def on_step_end(self, action, observation, reward, done):
            Count.plies += 1

    torch.manual_seed(0)
    np.random.seed(0)
    env = MinitChessEnvironment()
    policy = SimpleAlphaZeroPolicy(Network().eval())
    agents = [SimpleAlphaZeroAgent(environment=env, policy=policy, num_simulations=sims) for _ in range(2)]
    referee = RoundRobinReferee(agent_tuple=tuple(agents))
    cbs = [Count(), Init(agents[0]), Init(agents[1])]
    with torch.no_grad():
        run_episodes(env, referee, 1, callbacks=cbs, use_tqdm=False)          # warm-up episode
        Count.plies, episodes = 0, 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            referee.reset()
            run_episodes(env, referee, 1, callbacks=cbs, use_tqdm=False)
            episodes += 1
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return {'value': Count.plies * sims / dt, 'unit': 'sims/s', 'positions_per_second': Count.plies / dt, 'sims_per_move': sims,
            'episodes': episodes, 'seconds': dt,
            'api': 'SimpleAlphaZeroAgent.select_action x 2 + RoundRobinReferee + MinitChessEnvironment + run_episodes, one game at a time'}


def measure_e2e(sp, args, world):
    """Same metric through the reference-facing call: BatchedAlphaZeroAgent.select_actions_packed, the batched
    `SimpleAlphaZeroAgent.select_action`.  Every step uploads the positions from pinned host memory, searches,
    downloads the root statistics, picks the moves on the host like exp/agent.py:113-118 and plays them through
    the environment call (`az_play`), reading the results and the new positions back."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from minitchess_alphazero_b200 import rules
    from minitchess_alphazero_b200._lib import MC_MAX_MOVES, STATE_DTYPE
    from minitchess_alphazero_b200.agent import BatchedAlphaZeroAgent
    from minitchess_alphazero_b200.policy import SimpleAlphaZeroPolicy
    eng, G, S = sp.engine, sp.n_games, sp.num_simulations
    # the agent adopts the engine of the timed run: its games carry on mid-game with their trees
    agent = BatchedAlphaZeroAgent(SimpleAlphaZeroPolicy(sp.network), G, S, rng=np.random.RandomState(7), engine=eng)
    pinned = torch.empty(G * 5, dtype=torch.int32).pin_memory()
    host_states = pinned.numpy().view(np.uint32).view(STATE_DTYPE)
    # restarting all games from the start position would replay an opening the evaluation cache has just seen
    start = np.full(G, rules.state_from_fen(rules.STARTING_FEN), dtype=STATE_DTYPE)
    host_states[:] = eng.game_states()[0]

    def step():
        actions, _codes, _pi, _n = agent.select_actions_packed(host_states)    # H2D: positions; D2H: root statistics
        results = eng.play(actions)                                         # H2D: moves, D2H: results
        new_states, _ = eng.game_states()                                   # D2H: positions
        host_states[:] = new_states
        done = np.nonzero(results != 0)[0].astype(np.int32)
        if len(done):
            agent.init_mcts(game_ids=done)                                  # new episodes
            host_states[done] = start[done]
        return G * 20 + G * 4 + G * 2, G * MC_MAX_MOVES * 6 + G * 4 + G + G * 21

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    steps = max(2, min(args.steps, 4))
    c0 = eng.counters()
    t0 = time.perf_counter()
    for _ in range(steps):
        h2d, d2h = step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    c1 = eng.counters()
    t = torch.tensor([dt, float(c1['simulations'] - c0['simulations'])], dtype=torch.float64, device='cuda')
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
        dt, sims = float(tmax[0]), float(t[1])
    else:
        sims = float(t[1])
    return {'value': sims / dt, 'unit': 'sims/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h), 'steps': steps,
            'api': 'BatchedAlphaZeroAgent.select_actions_packed (set_positions / search / root_stats, host move choice) + Engine.play, '
                   'host (pinned) buffers'}


def emit(line):
    """The one JSON line goes to the process's original stdout."""
    os.write(_REAL_STDOUT, (line + '\n').encode())


if __name__ == '__main__':
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner under torchrun, ...)
    # is sent to stderr instead
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
