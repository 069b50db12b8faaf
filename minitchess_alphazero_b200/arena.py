"""Batched arena: network A against network B over many concurrent games on one engine.

This is the evaluation gate the reference sketches but leaves commented out (exp/learner.py:97-145:
`ARENA_GAME_NUMBER_PER_SIDE` games per colour between the new and the old network, counted by
`WinnerRecorder`, exp/callbacks.py:7-28; the result `new_wins / (new_wins + old_wins + 1e-8)` was to be
compared with 0.55, app/base.py:195-196).  Every game already owns two trees, one per colour's agent; here the
tree of the side to move is evaluated with that side's network.
"""
import numpy as np

from . import _lib
from .engine import Engine

GATE_THRESHOLD = 0.55                  # app/base.py:196 (`if result > 0.55:`)
ARENA_GAME_NUMBER_PER_SIDE = 3         # exp/learner.py:20


def gate_result(new_wins, old_wins):
    """exp/learner.py:145: the share of decisive games the new network won (draws do not count)."""
    return new_wins / (new_wins + old_wins + 1e-8)


class _NetworkSide:
    """One weight set held by a forward-only engine (its own trees are unused)."""

    def __init__(self, network, n_rows):
        import torch
        from .policy import flatten_state_dict
        self.engine = Engine(1, max_sims_per_move=1, network=1)
        self.engine.set_weights(flatten_state_dict(network.state_dict(), device='cuda'))
        self.n_rows = n_rows
        self.logits = torch.empty(n_rows, 554, device='cuda')
        self.values = torch.empty(n_rows, device='cuda')

    def __call__(self, tokens, clocks, leaf_states=None):
        e = self.engine
        _lib.check(_lib.lib().az_network_forward(e._h, _lib.ptr(tokens), _lib.ptr(clocks), self.n_rows, _lib.ptr(self.logits),
                                                 _lib.ptr(self.values)))
        return self.logits, self.values


class Arena:
    """`Arena(net_a, net_b).play()` -- games_per_side games with A as white, as many with A as black.

    `evaluators=(fa, fb)` replaces the two networks by callables `f(tokens, clocks, leaf_states) -> (logits [n, 554],
    values [n])` over the engine's leaf batch (numpy arrays with `_backend`, the CPU test harness; CUDA tensors otherwise).
    """

    def __init__(self, network_a=None, network_b=None, games_per_side=64, num_simulations=36, cpuct=1.0, tau_change=6, epsilon=0.25,
                 alpha=0.6, seed=0, evaluators=None, _backend=None):
        self.games_per_side = int(games_per_side)
        self.n_games = 2 * self.games_per_side
        self.num_simulations = int(num_simulations)
        self.tau_change = int(tau_change)
        self._host = _backend is not None
        self.engine = Engine(self.n_games, _backend=_backend, max_sims_per_move=num_simulations, cpuct=float(cpuct),
                             tau_change=int(tau_change), dirichlet_epsilon=float(epsilon), dirichlet_alpha=float(alpha), seed=int(seed),
                             device_rng=1)
        if evaluators is None:
            evaluators = (_NetworkSide(network_a, self.n_games), _NetworkSide(network_b, self.n_games))
        self._sides = tuple(evaluators)
        self.a_is_white = np.arange(self.n_games) < self.games_per_side     # first half: A has the white pieces
        self._rng = np.random.RandomState(seed)

    def _simulate(self, a_moves):
        """num_simulations simulations in every running game; the leaf of game g is evaluated by the side to move there."""
        eng = self.engine
        if self._host:
            for _ in range(self.num_simulations):
                eng.select_expand()
                tokens, clocks, _needs, leaf_states = eng.leaf_batch()
                (la, va), (lb, vb) = (f(tokens, clocks, leaf_states) for f in self._sides)
                eng.backup(np.where(a_moves, va, vb).astype(np.float32), logits=np.where(a_moves[:, None], la, lb).astype(np.float32))
            return
        import torch
        tokens, clocks, _needs = eng.leaf_batch_device()
        mask = torch.from_numpy(a_moves).cuda()
        for _ in range(self.num_simulations):
            eng.select_expand()
            (la, va), (lb, vb) = (f(tokens, clocks, None) for f in self._sides)
            eng.backup(torch.where(mask, va, vb).contiguous(), logits=torch.where(mask[:, None], la, lb).contiguous())

    def play(self, start_states=None):
        """Plays all games to the end (from STARTING_FEN, or from `start_states`, one packed position per game).  Returns
        {'a': wins of A, 'b': wins of B, 'draws', 'a_score': (wins + draws/2)/n, 'a_as_white': (wins, losses),
        'a_as_black': (wins, losses), 'gate': gate_result(a, b), 'plies'}."""
        eng, G = self.engine, self.n_games
        eng.reset_games(states=start_states)
        states, results = eng.game_states()
        winners = np.zeros(G, dtype=np.int8)                  # +1: A won, -1: B won, 0: draw / running
        ply = 0
        while (results == 0).any():
            white_to_move = (states['meta'] & 1).astype(bool)
            a_moves = self.a_is_white == white_to_move        # per game: is it A's agent that searches and moves now?
            self._simulate(a_moves)
            codes, visits, _, n_legal = eng.root_stats(want_q=False)
            active = np.nonzero(results == 0)[0]
            actions = np.zeros(len(active), dtype=np.uint16)
            fullmove = (states['meta'] >> 16) & 0xff
            for k, g in enumerate(active):
                E = int(n_legal[g])
                n = visits[g, :E].astype(np.float64)
                if fullmove[g] < self.tau_change:              # exp/agent.py:113-118
                    pick = self._rng.choice(E, p=n / n.sum())
                else:
                    best = np.nonzero(n == n.max())[0]
                    pick = self._rng.choice(best)
                actions[k] = codes[g, pick]
            res = eng.play(actions, game_ids=active.astype(np.int32))
            for k, g in enumerate(active):
                if res[k] in (1, 2):                           # decisive: the side that just moved won (WinnerRecorder:
                    winners[g] = 1 if a_moves[g] else -1       # `winner = not referee.turn`, exp/callbacks.py:23)
            states, results = eng.game_states()
            ply += 1
        a, b = int((winners == 1).sum()), int((winners == -1).sum())
        w = self.a_is_white
        return {'a': a, 'b': b, 'draws': G - a - b, 'a_score': (a + 0.5 * (G - a - b)) / G, 'plies': ply,
                'a_as_white': (int((winners[w] == 1).sum()), int((winners[w] == -1).sum())),
                'a_as_black': (int((winners[~w] == 1).sum()), int((winners[~w] == -1).sum())),
                'gate': gate_result(a, b)}


def passes_gate(new_network, old_network, games_per_side=ARENA_GAME_NUMBER_PER_SIDE, threshold=GATE_THRESHOLD, **arena_options):
    """The gate of app/base.py:195-196: the new network replaces the old one only if it wins more than `threshold` of the
    decisive arena games (exp/learner.py:97-145; ARENA_GAME_NUMBER_PER_SIDE = 3 there).  Returns (accepted, arena result)."""
    out = Arena(new_network, old_network, games_per_side=games_per_side, **arena_options).play()
    return out['gate'] > threshold, out
