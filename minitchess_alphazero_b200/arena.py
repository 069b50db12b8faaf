"""Batched arena: network A against network B over many concurrent games on one engine.

This is the evaluation gate the reference sketches but leaves commented out (exp/learner.py:97-145:
`ARENA_GAME_NUMBER_PER_SIDE` games per colour between the new and the old network, counted by
`WinnerRecorder`, exp/callbacks.py:7-28).  Every game already owns two trees, one per colour's agent; here the
tree of the side to move is evaluated with that side's network.
"""
import numpy as np
import torch

from . import _lib
from .engine import Engine
from .policy import flatten_state_dict


class Arena:
    def __init__(self, network_a, network_b, games_per_side=64, num_simulations=36, cpuct=1.0, tau_change=6, epsilon=0.25,
                 alpha=0.6, seed=0):
        self.games_per_side = int(games_per_side)
        self.n_games = 2 * self.games_per_side
        self.num_simulations = int(num_simulations)
        self.tau_change = int(tau_change)
        self.engine = Engine(self.n_games, max_sims_per_move=num_simulations, cpuct=float(cpuct), tau_change=int(tau_change),
                             dirichlet_epsilon=float(epsilon), dirichlet_alpha=float(alpha), seed=int(seed), device_rng=1)
        # two forward-only engines hold the two weight sets (their own trees are unused)
        self._nets = []
        for net in (network_a, network_b):
            e = Engine(1, max_sims_per_move=1, network=1)
            e.set_weights(flatten_state_dict(net.state_dict(), device='cuda'))
            self._nets.append(e)
        self.a_is_white = np.arange(self.n_games) < self.games_per_side     # first half: A has the white pieces
        self._rng = np.random.RandomState(seed)

    def _forward(self, which, tokens, clocks, logits, values):
        e = self._nets[which]
        _lib.check(_lib.lib().az_network_forward(e._h, _lib.ptr(tokens), _lib.ptr(clocks), self.n_games, _lib.ptr(logits),
                                                 _lib.ptr(values)))

    def play(self):
        """Plays all games to the end.  Returns {'a': wins of A, 'b': wins of B, 'draws': n, 'a_score': (wins + draws/2)/n}."""
        eng, G = self.engine, self.n_games
        eng.reset_games()
        tokens, clocks, _needs = eng.leaf_batch_device()
        la = torch.empty(G, 554, device='cuda'); lb = torch.empty_like(la)
        va = torch.empty(G, device='cuda'); vb = torch.empty_like(va)
        a_white = torch.from_numpy(self.a_is_white).cuda()
        states, results = eng.game_states()
        winners = np.zeros(G, dtype=np.int8)                  # +1: A won, -1: B won, 0: draw / running
        ply = 0
        while (results == 0).any():
            white_to_move = (ply % 2 == 0)
            a_moves = a_white if white_to_move else ~a_white
            for _ in range(self.num_simulations):
                eng.select_expand()
                self._forward(0, tokens, clocks, la, va)
                self._forward(1, tokens, clocks, lb, vb)
                eng.backup(torch.where(a_moves, va, vb).contiguous(), logits=torch.where(a_moves[:, None], la, lb).contiguous())
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
                if res[k] in (1, 2):                           # decisive: the side that just moved won
                    mover_is_a = bool(self.a_is_white[g]) == white_to_move
                    winners[g] = 1 if mover_is_a else -1
            states, results = eng.game_states()
            ply += 1
        a, b = int((winners == 1).sum()), int((winners == -1).sum())
        return {'a': a, 'b': b, 'draws': G - a - b, 'a_score': (a + 0.5 * (G - a - b)) / G, 'plies': ply}
