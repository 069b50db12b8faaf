"""Batched self-play: thousands of concurrent MinitChess games on one GPU, the superset API the
reference does not have (its actors play one game, one simulation, one batch-1 evaluation at a
time: app/base.py:108-124).  Replay tuples come back in the reference's episode format
(exp/callbacks.py:31-54) so `SimpleAlphaZeroDataset.push` / `collate_fn` run unchanged.
"""
import numpy as np
import torch

from . import rules
from ._lib import MC_MAX_MOVES
from .engine import Engine
from .policy import flatten_state_dict


def replay_to_episode_dicts(tuples):
    """Packed az_replay_tuple records -> the dicts InfoRecorder emits (exp/callbacks.py:40-53)."""
    out = []
    for t in tuples:
        E = int(t['n_legal'])
        out.append({'observation': rules.state_to_fen(t['observation']),
                    'legal_moves': t['codes'][:E].astype(int).tolist(),
                    'pi': t['pi'][:E].astype(np.float64).tolist(),
                    'action': int(t['action']), 'reward': float(t['reward'])})
    return out


def collate_device(tuples, device='cuda'):
    """exp/learner.py:23-41 `collate_fn` for packed replay tuples, on the GPU (az_collate): returns
    [pi float32 (n,554), channels int64 (n,2,6,5), clock float32 (n,1), reward float32 (n,1)] as CUDA tensors.
    `tuples` is a REPLAY_DTYPE numpy array or a uint8 CUDA tensor [n, 604]."""
    import ctypes
    from . import _lib
    from .engine import REPLAY_DTYPE
    if isinstance(tuples, np.ndarray):
        n = len(tuples)
        src = np.ascontiguousarray(tuples, dtype=REPLAY_DTYPE)
    else:
        n = tuples.shape[0]
        src = tuples.contiguous()
    pi = torch.empty(n, 554, dtype=torch.float32, device=device)
    channels = torch.empty(n, 2, 6, 5, dtype=torch.int64, device=device)
    clock = torch.empty(n, 1, dtype=torch.float32, device=device)
    reward = torch.empty(n, 1, dtype=torch.float32, device=device)
    _lib.check(_lib.lib().az_collate(_lib.ptr(src), n, _lib.ptr(pi), _lib.ptr(channels), _lib.ptr(clock), _lib.ptr(reward)))
    return [pi, channels, clock, reward]


class BatchedSelfPlay:
    """`n_games` concurrent games x `num_simulations` per move in throughput mode: Philox Dirichlet
    noise, move sampling, replay recording and game restarts all on the device, leaves evaluated by the
    hand-written sm_100a network inside libmcaz.so (`az_search` / `az_selfplay`)."""

    def __init__(self, network, n_games, num_simulations, cpuct=1.0, tau_change=6, epsilon=0.25, alpha=0.6, seed=0,
                 precision='bf16', **engine_options):
        assert precision in ('bf16', 'fp8'), precision      # 'fp8': the e4m3 tower (az_config.network = 2), opt-in
        self.precision = precision
        self.n_games, self.num_simulations = int(n_games), int(num_simulations)
        # with leaves_per_step = K every search step runs K descents per tree: size the arenas for all of them
        leaves = int(engine_options.get('leaves_per_step', 1) or 1)
        engine_options.setdefault('recycle', 1)       # games only move forward here: plies behind them can be dropped
        if leaves == 1 and not engine_options.get('lookahead_rows'):
            # continuous self-play: a short last tile pair of a batch (<= 224 of 256 rows) waits for the next batch
            engine_options.setdefault('defer_rows', 224)
        if leaves == 1:
            # exact evaluation cache sized for a few moves' worth of evaluations (192 B per entry; 4096 x 200 -> 2^23 = 1.6 GB)
            want = max(1, self.n_games * self.num_simulations * 8)
            engine_options.setdefault('eval_cache_log2', min(24, max(12, (want - 1).bit_length())))
        self.engine = Engine(n_games, max_sims_per_move=num_simulations * leaves, cpuct=float(cpuct), tau_change=int(tau_change),
                             dirichlet_epsilon=float(epsilon), dirichlet_alpha=float(alpha), seed=int(seed),
                             device_rng=1, network=2 if precision == 'fp8' else 1, **engine_options)
        self.network = network
        self.weights_version = 0
        self.sync_weights()

    def sync_weights(self, flat=None, version=None):
        """Push the torch Network's weights to the engine (SimulatePuppet.load_weights, app/base.py:126-129); `version` is
        the learner's stamp that finished games' replay tuples will carry (app/base.py:63-68; default: previous + 1)."""
        self.weights_version = int(version) if version is not None else self.weights_version + 1
        self.engine.set_weights(flatten_state_dict(self.network.state_dict(), device='cuda') if flat is None else flat,
                                version=self.weights_version)

    def search(self):
        self.engine.search(self.num_simulations)

    def step(self):
        """One move in every game: search, choose, record, play, restart finished games."""
        self.search()
        self.engine.play_device()

    def run(self, n_moves):
        for _ in range(int(n_moves)):
            self.step()

    def stagger(self, max_ply=60, sims=8):
        """Spread the games uniformly over the plies 0..max_ply-1 of a game (game g ends up at ply g*max_ply//n_games)
        with a cheap pre-roll of `sims` simulations per move: the stationary population of a long-running actor,
        instead of n_games copies of the start position marching through the opening together.  `sims` must be
        large enough for the visit counts to spread over the root's moves (with 2 simulations every game would play
        edge 0, exp/agent.py:84-85 on an all-zero u), or all games of one ply would share one position."""
        target = (np.arange(self.n_games, dtype=np.int64) * max_ply) // self.n_games
        for step in range(max_ply):
            ids = np.nonzero(target == max_ply - step)[0].astype(np.int32)
            if len(ids):
                self.engine.reset_games(game_ids=ids)
            self.engine.search(min(sims, self.num_simulations))
            self.engine.play_device()
        ids = np.nonzero(target == 0)[0].astype(np.int32)
        if len(ids):
            self.engine.reset_games(game_ids=ids)
        self.engine.drain_replay()          # the pre-roll's games are not training data

    def run_continuous(self, n_batches):
        """Continuous self-play (az_selfplay): `n_batches` network batches; every game searches, chooses, records,
        plays and restarts on its own inside the search kernel, so a game whose move needed fewer network rows
        (terminal or cached leaves) simply moves earlier and the batch stays full."""
        self.engine.selfplay(int(n_batches), self.num_simulations)

    def drain(self):
        return self.engine.drain_replay()

    # ---- measurement hooks used by bench.py ---------------------------------------------------
    CONV_FLOP_PER_EVAL = 2 * 17694720          # one 3x3 256->256 convolution on one 6x5 board

    def reset_kernel_timer(self):
        """Start bracketing the tower-convolution launches with CUDA events on the engine's stream."""
        self.engine.profile_network(True, read=True)
        self.engine.profile_tree(True, read=True)
        return True

    def kernel_profile(self, evaluations=None):
        """Roofline record of the dominant kernel (tower_tc_kernel) from the events recorded since
        reset_kernel_timer(): algorithmic FLOP of the rows it evaluated / the time it ran.  `evaluations` = rows
        evaluated in that window (the engine's evaluation counter); default: every launch had a full batch."""
        ms, n, per_forward = self.engine.profile_network(False, read=True)
        if n == 0 or ms <= 0:
            return None
        rows = self.engine.n_slots if evaluations is None else evaluations / n     # mean rows per forward
        # algorithmic FLOP of one launch: 18 convolutions 256->256, plus the stem (8->256) and the heads' three 1x1
        # convolutions (256->3) that run as level 0 and in the last epilogue of the same kernel
        flop = rows * (self.CONV_FLOP_PER_EVAL * 18 + 2 * 552960 + 2 * 23040) / per_forward
        ms_launch = ms / per_forward
        fused = per_forward == 1
        return {'bound': 'tensor',
                'kernel': ('tower_tc_kernel (stem + 18 x [3x3 conv 256->256] + head 1x1 convs in one data-flow ordered launch, tcgen05 cta_group::2)' if fused
                           else 'tower_tc_kernel (tcgen05 cta_group::2 3x3 conv 256->256, launched once per layer: MCAZ_TOWER=layers)'),
                'achieved': flop / (ms_launch / 1e3) / 1e12, 'unit': 'TFLOP/s',
                'traffic': None,          # bench.py fills this in from the committed ncu pages (profiles/tower_traffic.json)
                'ms_per_launch': ms_launch, 'launches_timed': n * per_forward,
                'rows_per_launch': rows, 'flop_per_launch': flop,
                # the kernel skips the 62 of 270 tap-positions that multiply zero padding: MMAs actually issued
                'achieved_mma': flop * 208 / 270 / (ms_launch / 1e3) / 1e12,
                'note': 'achieved counts the algorithmic FLOPs of SURVEY.md 8(d) (all 9 taps at all 30 squares) of the rows '
                        'actually evaluated; taps on zero padding are skipped, so issued MMA work is 208/270 of it (achieved_mma)'}

    def tree_profile(self, counters0, counters1):
        """HBM-side record of search_step_kernel (backup + select + expand) between two counter snapshots:
        algorithmic bytes per simulation after SURVEY.md 8(d) -- select d x (32 B node header + E x 20 B of Q, N, P),
        expand 32 B + E' x 22 B written, backup d x 32 B read-modify-write -- over the kernel's measured time."""
        ms, n = self.engine.profile_tree(False, read=True)
        d = {k: counters1[k] - counters0[k] for k in counters1}
        sims = d['simulations']
        if n == 0 or ms <= 0 or sims == 0:
            return None
        select_b = d['path_depth'] * 32 + d['path_edges'] * 20
        expand_b = d['nodes'] * 32 + d['edges'] * 22
        backup_b = d['path_depth'] * 32
        total = select_b + expand_b + backup_b
        return {'bound': 'hbm', 'kernel': 'search_step_kernel (backup + PUCT select + expand, warp per game)',
                'achieved': total / (ms / 1e3) / 1e9, 'unit': 'GB/s', 'bytes_per_sim': total / sims,
                'mean_depth': d['path_depth'] / sims, 'mean_edges_per_level': d['path_edges'] / max(d['path_depth'], 1),
                'ms_per_launch': ms / n, 'launches_timed': n, 'sims_per_launch': sims / n}
