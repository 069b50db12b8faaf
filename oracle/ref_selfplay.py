"""CPU restatement of the reference self-play hot path.  TEST INFRASTRUCTURE ONLY.

Follows, symbol by symbol (reference paths):
  RefEpisode      exp/environment.py:23-85   (MinitChessEpisode over the `chess` shim)
  RefTree         exp/agent.py:24-88         (MonteCarloTreeSearch; iterative, same arithmetic)
  RefNetwork      exp/policy.py:53-105       (Network forward + tokeniser, functional form)
  ref_get_distribution / ref_select_action   exp/policy.py:115-122, exp/agent.py:110-119
  play_game       app/base.py:113-120 + the erlyx loop (two agents sharing one policy)

It exists because `/root/reference` (pure Python) cannot travel to the GPU box, while the
parity checks and the CPU baseline must run there.  PINNED: tests/golden/make_golden.py runs
the reference's own UNMODIFIED exp/agent.py + exp/policy.py + exp/environment.py (on the
shims) and this restatement from the same seeds and records that they agree exactly
(tests/test_oracle_pinned.py re-checks it whenever /root/reference is present).
The numpy global RNG is consumed in the reference's order: one dirichlet(E) per simulation
whose root is already expanded (exp/agent.py:81-82), then one choice per move (:115,118).
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle.shims import chess as mchess
from oracle import rules_c

STARTING_FEN = '2nbk/2ppp/5/5/PPP2/KBN2 w 0 1'
NUM_ACTIONS = 554


def _uci_tables():
    """code <-> 4-char uci per side, from the table formula (exp/generate_moves_list.py)."""
    inv = {True: {}, False: {}}
    fwd = {True: {}, False: {}}
    for f in range(30):
        for t in range(30):
            for side in (True, False):
                c = rules_c.code_of(f, t, side)
                if c >= 0:
                    u = mchess.square_name(f) + mchess.square_name(t)
                    fwd[side][u] = c
                    inv[side][c] = u
    return fwd, inv


MOVES, MOVES_INV = _uci_tables()


class RefEpisode:
    """exp/environment.py:23-85."""

    def __init__(self, fen):
        self.board = mchess.Board(fen)
        self._refresh()

    def _refresh(self):                                   # :34-50
        b = self.board
        self.fen = b.fen()
        res = b.result()
        self.done = res != '*'
        self.reward = None if not self.done else (0.0 if res == '1/2-1/2' else 1.0)
        self.moves = b.legal_moves
        self.legal = sorted(MOVES[b.turn][m.uci()[:4]] for m in self.moves)

    def step(self, code):                                  # :68-82
        if self.done:
            raise RuntimeError('episode finished')
        uci = MOVES_INV[self.board.turn][code]
        mv = mchess.Move.from_uci(uci)
        if mv not in self.moves:
            mv = mchess.Move.from_uci(uci + 'q')
        if mv not in self.moves:
            raise ValueError('illegal move %d (%s)' % (code, uci))
        self.board.push(mv)
        self._refresh()
        return self.fen, self.reward, self.done


class RefNetwork:
    """exp/policy.py:53-105 in functional form over a reference-layout state_dict (eval mode)."""

    def __init__(self, state_dict):
        self.sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}

    def _convblock(self, x, prefix, pad, relu=True):       # :15-38, BatchNorm in eval mode
        sd = self.sd
        x = F.conv2d(x, sd[prefix + '.layers.0.weight'], sd[prefix + '.layers.0.bias'], padding=pad)
        x = F.batch_norm(x, sd[prefix + '.layers.1.running_mean'], sd[prefix + '.layers.1.running_var'],
                         sd[prefix + '.layers.1.weight'], sd[prefix + '.layers.1.bias'], False, 0.0, 1e-5)
        return F.relu(x) if relu else x

    def forward(self, tokens, clock):                      # :71-80
        sd = self.sd
        x = F.embedding(tokens, sd['emb.weight']).permute(0, 1, 4, 2, 3).contiguous().view(-1, 8, 6, 5)
        x = self._convblock(x, 'resbody.0', 1)
        for i in range(1, 10):                             # :41-50
            y = self._convblock(x, 'resbody.%d.convblock1' % i, 1)
            y = self._convblock(y, 'resbody.%d.convblock2' % i, 1, relu=False)
            x = F.relu(y + x)
        px = self._convblock(x, 'pconv', 0).view(-1, 60)
        p = F.linear(torch.cat([px, clock], 1), sd['plinear.weight'], sd['plinear.bias'])
        vx = self._convblock(x, 'vconv', 0).view(-1, 30)
        v = F.relu(F.linear(torch.cat([vx, clock], 1), sd['vlinear.0.weight'], sd['vlinear.0.bias']))
        v = torch.tanh(F.linear(v, sd['vlinear.2.weight'], sd['vlinear.2.bias']))
        return p, v

    @staticmethod
    def tokenize_fen(fen):                                 # :82-105
        rows, color, _half, full = fen.split()
        if color == 'b':
            rows = rows[::-1].swapcase()
        cells = []
        for ch in rows:
            if ch == '/':
                continue
            cells.extend('0' * int(ch) if ch.isdigit() else ch)
        mine = ['0prbnqk'.index(c.lower()) if c.isupper() else 0 for c in cells]
        theirs = ['0prbnqk'.index(c) if c.islower() else 0 for c in cells]
        clock = float(full) + (0.5 if color == 'b' else 0.0)
        tokens = torch.tensor(mine + theirs, dtype=torch.long).reshape(1, 2, 6, 5)
        return tokens, torch.tensor([[clock / 30]]).float()

    def evaluate(self, fen, legal):
        """(priors over `legal` as float32 ndarray, value float) -- exp/agent.py:67-69."""
        with torch.no_grad():
            p, v = self.forward(*self.tokenize_fen(fen))
            return p[0][legal].softmax(0).numpy(), v.item()


class RefTree:
    """exp/agent.py:24-88.  `evaluate(fen, legal) -> (P float32[E], v float)`."""

    def __init__(self, evaluate, cpuct=1, epsilon=0.25, alpha=0.6, rng=None):
        self.evaluate = evaluate
        self.cpuct = cpuct
        self.rng = rng or np.random                        # the reference uses the global legacy RNG
        self.epsilon, self.alpha = epsilon, alpha
        self.Q, self.N, self.P, self.legal, self.terminal = {}, {}, {}, {}, {}
        self.visited = set()
        self.n_evals = 0

    def simulate(self, num_simulations, fen):              # :41-45
        for _ in range(num_simulations):
            self._one_simulation(fen)

    def _one_simulation(self, root_fen):                   # :54-88 unrolled into a loop
        ep = RefEpisode(root_fen)
        path = []
        while True:
            node = ep.fen
            if node not in self.visited:                   # :57-73 expand + evaluate
                self.visited.add(node)
                if ep.done:
                    value = -ep.reward
                    self.terminal[node] = value
                else:
                    E = len(ep.legal)
                    self.Q[node], self.N[node] = np.zeros(E), np.zeros(E)
                    prior, value = self.evaluate(node, ep.legal)
                    self.n_evals += 1
                    self.P[node], self.legal[node] = prior, ep.legal
                break
            if node in self.terminal:                      # :75-77 (Q1: sign flip on revisit)
                value = -self.terminal[node]
                break
            Q, N, P = self.Q[node], self.N[node], self.P[node]
            if not path and self.epsilon > 0:              # :81-82 root noise, every simulation
                P = (1 - self.epsilon) * P + self.epsilon * self.rng.dirichlet([self.alpha] * len(P))
            u = Q + self.cpuct * P * np.sqrt(N.sum()) / (1 + N)   # :84
            a = int(u.argmax())
            ep.step(self.legal[node][a])
            path.append((node, a))
        for node, a in reversed(path):                     # :47-52
            value = -value
            Q, N = self.Q[node], self.N[node]
            Q[a] = (N[a] * Q[a] + value) / (N[a] + 1)
            N[a] += 1


def ref_get_distribution(tree, fen, num_simulations):      # exp/policy.py:115-122
    tree.simulate(num_simulations, fen)
    N = tree.N[fen]
    return {'legal_moves': tree.legal[fen], 'pi': N / N.sum()}


def ref_select_action(tree, fen, num_simulations, tau_change=6, rng=None):   # exp/agent.py:110-119
    rng = rng or np.random
    info = ref_get_distribution(tree, fen, num_simulations)
    if int(fen.split()[3]) < tau_change:
        action = rng.choice(info['legal_moves'], p=info['pi'])
    else:
        best = np.where(info['pi'] == info['pi'].max())[0]
        action = info['legal_moves'][rng.choice(best)]
    return int(action), info


def play_game(evaluate, num_simulations=36, cpuct=1, tau_change=6, fen=None, max_plies=None, trees=None,
              rng=None, epsilon=0.25):
    """One self-play game: two trees (one per agent) sharing one evaluator (app/base.py:113-120).

    Returns a list of per-ply records {'observation','legal_moves','pi','action'} plus the final
    (reward, fen); rewards are back-filled like exp/callbacks.py:49-53.
    """
    trees = trees or [RefTree(evaluate, cpuct, epsilon=epsilon, rng=rng) for _ in range(2)]
    ep = RefEpisode(fen or STARTING_FEN)
    records, turn = [], 0
    while not ep.done and (max_plies is None or len(records) < max_plies):
        obs = ep.fen
        action, info = ref_select_action(trees[turn], obs, num_simulations, tau_change, rng)
        records.append({'observation': obs, 'legal_moves': list(info['legal_moves']),
                        'pi': info['pi'].tolist(), 'action': action})
        ep.step(action)
        turn ^= 1
    if ep.done:
        reward = ep.reward
        for rec in reversed(records):
            rec['reward'] = reward
            reward = -reward
    return records, ep, trees
