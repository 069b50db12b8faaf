"""Drop-in for exp/agent.py: `SimpleAlphaZeroAgent`, `MonteCarloTreeSearch`, `RoundRobinReferee`
with the reference's signatures, the search itself running on the GPU engine.

Per-agent use (one game at a time, like app/base.py:113-120) keeps the reference's semantics
exactly: one tree per agent kept for the whole game, Dirichlet noise drawn from numpy's global
RNG once per simulation whose root is expanded (exp/agent.py:81-82), move choice by
`np.random.choice` (exp/agent.py:113-118).  For thousands of concurrent games use
`selfplay.BatchedSelfPlay`.
"""
import weakref

import numpy as np

from . import rules
from ._lib import MC_MAX_MOVES
from .engine import Engine
from .erlyx_compat import ActionData, BaseAgent, PolicyAgent
from .policy import flatten_state_dict, weights_fingerprint


class RoundRobinReferee(BaseAgent):                         # exp/agent.py:6-21
    def __init__(self, agent_tuple):
        self._agent_tuple = tuple(agent_tuple)
        self._turn = False

    def select_action(self, observation):
        action = self._agent_tuple[int(self._turn)].select_action(observation)
        self._turn = not self._turn
        return action

    def reset(self):
        self._turn = False

    @property
    def turn(self):
        return self._turn


class _NodeField:
    """`mcts['N'][fen]`-style access to one per-node array of the GPU tree."""

    def __init__(self, tree, field):
        # weak: the search owns its fields, not the other way round -- no reference cycle, so a dropped agent frees
        # its tree slot (and, with the last user, the engine) at once instead of at the next garbage collection
        self._tree, self._field = weakref.proxy(tree), field

    def _fetch(self, fen):
        st = self._tree._node(fen)
        if st is None:
            return None
        if self._field == 'terminal':
            return st['terminal']
        if st['terminal'] is not None:
            return None
        return st[self._field]

    def __getitem__(self, fen):
        val = self._fetch(fen)
        if val is None:
            raise KeyError(fen)
        return val

    def get(self, fen, default=None):
        val = self._fetch(fen)
        return default if val is None else val

    def __contains__(self, fen):
        return self._fetch(fen) is not None

    # the whole dict, like the reference's (exp/agent.py:25-36): one read-back of the tree (az_tree_dump)
    def keys(self):
        term = self._field == 'terminal'
        return [fen for fen, rec in self._tree._all_nodes().items() if (rec['terminal'] is not None) == term]

    def items(self):
        term = self._field == 'terminal'
        return [(fen, rec[self._field]) for fen, rec in self._tree._all_nodes().items() if (rec['terminal'] is not None) == term]

    def values(self):
        return [v for _, v in self.items()]

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())


class _Visited:
    def __init__(self, tree):
        self._tree = weakref.proxy(tree)

    def __contains__(self, fen):
        return self._tree._node(fen) is not None

    def __iter__(self):
        return iter(self._tree._all_nodes())

    def __len__(self):
        return len(self._tree._all_nodes())


class _SharedEngine:
    """One engine (one game slot, two trees) serving up to two per-agent searches that evaluate with the same `Network`
    object -- the reference's actor wiring: two agents, one policy (app/base.py:113).  Each search owns one tree; the
    uploaded weights and the exact evaluation cache are common, so what one agent's search (or its look-ahead rows)
    evaluated, the other's finds in the cache.  Trees never mix: results are those of two separate engines."""

    def __init__(self, engine, capacity_sims, model):
        self.engine, self.capacity_sims, self.fingerprint = engine, capacity_sims, None
        self.model = weakref.ref(model)                # the pool key holds id(model): make sure it is still that object
        self.owners = [None, None]                     # weak references to the searches holding tree 0 / tree 1

    def free_tree(self):
        for k, ref in enumerate(self.owners):
            if ref is None or ref() is None:
                return k
        return None


_ENGINE_POOL = {}      # (id(model), simulations, search parameters) -> [_SharedEngine, ...]


class MonteCarloTreeSearch:
    """exp/agent.py:24-88 over one GPU-resident tree.  `model` is the torch `Network` (`policy.model`); its
    weights are mirrored into the engine's built-in sm_100a network (az_set_weights) and re-read whenever
    they change.  `evaluator` may instead be any callable object with `.forward(tokens_u8, clocks)` ->
    (logits, values) CUDA tensors (e.g. `policy.TorchEvaluator(net, dtype=torch.float32)` for fp32 checks)."""

    def __init__(self, environment, model, cpuct, epsilon=0.25, alpha=0.6, evaluator=None, rules_switches=None, _reuse=None,
                 engine_options=None):
        self._engine_options = dict(engine_options or {})      # overrides for the built-in evaluator's engine (tests)
        self._environment = environment
        self._model = model
        self._cpuct = cpuct
        self._epsilon, self._alpha = epsilon, alpha
        self._rules = rules_switches
        self._shared, self._tree = None, 0
        self._last_node = None                                  # (fen, node statistics) read since the last search
        self._dump = None                                       # the whole tree, read since the last search
        self._evaluator = evaluator
        if _reuse is not None and _reuse._shared is not None and _reuse._evaluator is evaluator:
            # a new game of the same agent: keep the engine (arenas, uploaded weights, cache), empty this agent's tree
            self._shared, self._tree = _reuse._shared, _reuse._tree
            self._shared.owners[self._tree] = weakref.ref(self)
            self._engine.reset_trees([self._tree])
        self._fields = {k: _NodeField(self, k) for k in ('Q', 'N', 'P', 'legal_moves', 'terminal')}
        self._fields['visited'] = _Visited(self)

    def __getitem__(self, item):
        return self._fields.get(item, None)

    @property
    def _engine(self):
        return self._shared.engine if self._shared is not None else None

    def _node(self, fen):
        if self._engine is None:
            return None
        # policy.get_distribution reads two fields of the root right after a search: one read-back serves both
        if self._last_node is not None and self._last_node[0] == fen:
            return self._last_node[1]
        stats = self._engine.node_stats(0, self._tree, rules.state_from_fen(fen))
        self._last_node = (fen, stats)
        return stats

    def _all_nodes(self):
        """{fen: {'legal_moves', 'N', 'Q', 'P', 'terminal'}} of every node of this agent's tree, in creation order."""
        if self._engine is None:
            return {}
        if self._dump is None:
            from .engine import NODE_DECISIVE, NODE_TERMINAL
            d = self._engine.tree_dump(0, self._tree)
            out = {}
            for i, st in enumerate(d['states']):
                info, off = int(d['info'][i]), int(d['edge_off'][i])
                E = 0 if info & NODE_TERMINAL else info & 0xffff
                out[rules.state_to_fen(st)] = {
                    'legal_moves': d['codes'][off:off + E].astype(int).tolist(), 'N': d['N'][off:off + E].astype(np.float64),
                    'Q': d['Q'][off:off + E].copy(), 'P': d['P'][off:off + E].copy(),
                    'terminal': ((-1.0 if info & NODE_DECISIVE else -0.0) if info & NODE_TERMINAL else None)}
            self._dump = out
        return self._dump

    def _ensure(self, num_simulations):
        if self._shared is None:
            opts = dict(max_sims_per_move=max(int(num_simulations), 1), cpuct=float(self._cpuct),
                        dirichlet_epsilon=float(self._epsilon), dirichlet_alpha=float(self._alpha),
                        network=0 if self._evaluator is not None else 1)
            share = False
            if self._evaluator is None:
                # one tree, one leaf per network pass: let every pass also evaluate the children of the new nodes into
                # the exact cache (a pass costs the same for 1 row as for 256) -- same trees, far fewer passes
                # (one tile of 256 rows at the reference's 36 simulations per move; two tiles cost 3 % more per pass and
                # halve the passes of long searches: 2.29 / 2.44 ms per move at 36 simulations, 8.06 / 6.95 ms at 200)
                opts.update(eval_cache_log2=18, lookahead_rows=255 if int(num_simulations) <= 64 else 511)
                opts.update(self._engine_options)
                share = bool(opts.pop('share_engine', True))
            if self._rules is not None:
                opts['rules'] = self._rules
            # a second search over the same Network object (the reference's two agents share one policy) takes the free
            # tree of the first one's engine: one copy of the weights, one evaluation cache
            key = (id(self._model), repr(sorted((k, repr(v)) for k, v in opts.items() if k != 'rules')), id(self._rules))
            holder = None
            if share:
                pool = _ENGINE_POOL.setdefault(key, [])
                pool[:] = [h for h in pool if h.model() is not None]         # engines of models that are gone
                for h in pool:
                    if h.model() is self._model and h.free_tree() is not None:
                        holder = h
                        break
            if holder is None:
                holder = _SharedEngine(Engine(1, **opts), int(num_simulations), self._model)
                if share:
                    _ENGINE_POOL.setdefault(key, []).append(holder)
            self._shared, self._tree = holder, holder.free_tree()
            holder.owners[self._tree] = weakref.ref(self)
            holder.engine.reset_trees([self._tree])
        elif int(num_simulations) > self._shared.capacity_sims:
            raise ValueError('num_simulations grew from %d to %d: the tree arenas were sized for the first value'
                             % (self._shared.capacity_sims, num_simulations))
        fp = weights_fingerprint(self._model)
        if fp != self._shared.fingerprint:          # first use, load_state_dict or an optimiser step
            if self._evaluator is None:
                self._engine.set_weights(flatten_state_dict(self._model.state_dict(), device='cuda'))
            elif hasattr(self._evaluator, 'load'):
                self._evaluator.load(self._model)
            self._shared.fingerprint = fp

    def simulate(self, num_simulations, observation):       # exp/agent.py:41-45
        self._last_node = self._dump = None
        self._ensure(num_simulations)
        eng, ev = self._engine, self._evaluator
        eng.set_positions(rules.state_from_fen(observation), trees=[self._tree])
        if ev is not None:
            tokens, clocks, _needs = eng.leaf_batch_device()
        if ev is None:
            return self._simulate_builtin(num_simulations)
        root_edges = -1                                     # unknown until the root is seen expanded
        for _ in range(num_simulations):
            noise = None
            if self._epsilon > 0:
                if root_edges < 0:
                    root_edges = int(eng.root_stats(want_q=False)[3][0])
                if root_edges > 0:                          # root already expanded: exp/agent.py:81-82
                    noise = np.zeros((1, MC_MAX_MOVES))
                    noise[0, :root_edges] = np.random.dirichlet([self._alpha] * root_edges)
            eng.select_expand(noise)
            if ev is None:
                eng.eval_backup()                           # built-in tcgen05 network
            else:
                logits, values = ev.forward(tokens, clocks)
                eng.backup(values, logits=logits)
        return self

    def _simulate_builtin(self, num_simulations):
        """Built-in network: one library call per move.  The reference draws one Dirichlet sample per simulation
        whose root is already expanded (exp/agent.py:81-82); `np.random.dirichlet(alpha, size=k)` consumes numpy's
        global stream exactly like k successive calls, so the block is drawn up front and handed over whole."""
        eng, left = self._engine, int(num_simulations)
        if self._epsilon <= 0:
            eng.search(left)                                # no noise (the engine was created with device_rng = 0)
            return self
        root_edges = int(eng.root_stats(want_q=False)[3][0])
        if root_edges <= 0 and left > 0:                    # unseen root: the first simulation only expands it, no noise
            eng.select_expand(None)
            eng.eval_backup()
            left -= 1
            root_edges = int(eng.root_stats(want_q=False)[3][0])
        if left > 0 and root_edges > 0:
            noise = np.zeros((left, 1, MC_MAX_MOVES))
            noise[:, 0, :root_edges] = np.random.dirichlet([self._alpha] * root_edges, size=left)
            eng.search_noise(noise)
        elif left > 0:                                      # finished position: nothing to mix, simulations still count
            eng.search_noise(np.zeros((left, 1, MC_MAX_MOVES)))
        return self

    @property
    def engine(self):
        return self._engine


class SimpleAlphaZeroAgent(PolicyAgent):                    # exp/agent.py:91-119
    def __init__(self, environment, policy, num_simulations, cpuct=1, tau_change=6):
        super().__init__(policy)
        self._environment = environment
        self._num_simulations = num_simulations
        self._cpuct = cpuct
        self._tau_change = tau_change
        self.init_mcts()

    def init_mcts(self):
        evaluator = getattr(getattr(self, '_mcts', None), '_evaluator', None)    # an injected evaluator survives resets
        old = getattr(self, '_mcts', None)
        self._mcts = MonteCarloTreeSearch(self._environment, self.policy.model, self._cpuct, evaluator=evaluator,
                                          _reuse=old, engine_options=getattr(old, '_engine_options', None))
        self._count = 0

    def select_action(self, observation):
        info = self.policy.get_distribution(observation, self._mcts, self._num_simulations)
        num_moves = int(observation.split()[3])
        if num_moves < self._tau_change:
            action = np.random.choice(info['legal_moves'], p=info['pi'])
        else:
            maxima = np.where(info['pi'] == info['pi'].max())[0]
            action = info['legal_moves'][np.random.choice(maxima)]
        return ActionData(action=action, info=info)


class BatchedAlphaZeroAgent:
    """The batched superset of `SimpleAlphaZeroAgent`: `n_games` concurrent games, each with
    the two per-colour trees the reference's two agents keep (app/base.py:113), searched together on one engine.

        agent = BatchedAlphaZeroAgent(policy, n_games=4096, num_simulations=200)
        actions = agent.select_actions(observations)          # list of FEN strings, one per game
        # ... step the environments with [a.action for a in actions] ...

    `select_actions` is `select_action` (exp/agent.py:110-119) for every game at once: the position of each game is
    handed to the tree of its side to move (kept from that side's previous move), `num_simulations` simulations run
    with the built-in network, and the move is chosen on the host exactly like the reference does -- sampled from
    pi while fullmove < tau_change, else uniformly among the maxima.  Root Dirichlet noise comes from the engine's
    Philox streams (one per game) instead of numpy's global RNG, which cannot serve thousands of games in a defined order.
    """

    def __init__(self, policy, n_games, num_simulations, cpuct=1, tau_change=6, seed=0, rng=None, engine=None, **engine_options):
        self._policy = policy
        self.n_games, self._num_simulations, self._tau_change = int(n_games), int(num_simulations), int(tau_change)
        if engine is not None:                                  # adopt a running engine (its games, trees and weights)
            assert engine.n_games == self.n_games and not engine_options
            self._engine = engine
        else:
            engine_options.setdefault('recycle', 1)             # positions only move forward within a game
            engine_options.setdefault('eval_cache_log2', 20)
            if self.n_games <= 128:                             # the network tile is mostly empty: fill it with look-ahead rows
                engine_options.setdefault('lookahead_rows', 256 - self.n_games)
            self._engine = Engine(self.n_games, max_sims_per_move=self._num_simulations, cpuct=float(cpuct), tau_change=self._tau_change,
                                  device_rng=1, network=1, seed=int(seed), **engine_options)
        self._rng = rng if rng is not None else np.random
        self._fingerprint = None
        self._all = np.arange(self.n_games, dtype=np.int32)
        self._stats_buffers = None                              # page-locked read-back buffers of select_actions_packed

    @property
    def policy(self):
        return self._policy

    @property
    def engine(self):
        return self._engine

    def init_mcts(self, game_ids=None):
        """MonteCarloInit.on_episode_begin (exp/callbacks.py:57-62) for the listed games (default: all)."""
        self._engine.reset_games(game_ids=game_ids)

    def _sync_weights(self):
        fp = weights_fingerprint(self._policy.model)
        if fp != self._fingerprint:                 # first use, load_state_dict or an optimiser step
            self._engine.set_weights(flatten_state_dict(self._policy.model.state_dict(), device='cuda'))
            self._fingerprint = fp

    def select_actions_packed(self, states):
        """`select_actions` on packed positions (STATE_DTYPE array, one per game).  Returns (actions uint16 [n],
        codes uint16 [n, M], pi float64 [n, M], n_legal int32 [n]).  A game whose position is finished gets n_legal 0."""
        eng = self._engine
        states = np.ascontiguousarray(states)
        assert len(states) == self.n_games, 'one position per game (the search advances every game of the engine)'
        self._sync_weights()
        eng.set_positions(states, trees=1 - (states['meta'] & 1).astype(np.int32))   # white to move -> tree 0
        eng.search(self._num_simulations)
        if self._stats_buffers is None:                         # read the root statistics back into page-locked memory, kept between calls
            import torch
            shape = (self.n_games, MC_MAX_MOVES)
            self._stats_pinned = (torch.empty(shape, dtype=torch.int16).pin_memory(), torch.empty(shape, dtype=torch.int32).pin_memory(),
                                  torch.empty(self.n_games, dtype=torch.int32).pin_memory())
            self._stats_buffers = (self._stats_pinned[0].numpy().view(np.uint16), self._stats_pinned[1].numpy().view(np.uint32),
                                   self._stats_pinned[2].numpy())
        codes, visits, _, n_legal = eng.root_stats(want_q=False, out=self._stats_buffers)
        codes = codes.copy()                                    # returned to the caller: not a view of the reused buffer
        n_legal = np.maximum(n_legal, 0)
        n = self.n_games
        E = np.maximum(n_legal, 1)
        width = int(E.max())                                    # the arrays are MC_MAX_MOVES wide; few positions need that
        w = visits[:, :width].astype(np.float64)
        pi = np.zeros(visits.shape, dtype=np.float64)
        pi[:, :width] = w / np.maximum(w.sum(1, keepdims=True), 1.0)
        # exp/agent.py:113-118, vectorised: sample from pi early in the game, else pick uniformly among the maxima
        rows = np.arange(n)
        choice = w.argmax(1)                                    # the only maximum in most positions
        is_max = w == w[rows, choice][:, None]
        is_max &= np.arange(width)[None, :] < E[:, None]
        tied = np.nonzero(is_max.sum(1) > 1)[0]
        if len(tied):
            choice[tied] = np.where(is_max[tied], self._rng.random_sample((len(tied), width)), -1.0).argmax(1)
        early = np.nonzero(((states['meta'] >> 16) & 0xff) < self._tau_change)[0]
        if len(early):
            cum = np.cumsum(w[early], axis=1)
            u = self._rng.random_sample(len(early)) * cum[np.arange(len(early)), E[early] - 1]
            choice[early] = np.minimum((cum <= u[:, None]).sum(1), E[early] - 1)
        return codes[np.arange(n), choice], codes, pi, n_legal

    def select_actions(self, observations):
        """observations: one FEN string per game -> list of ActionData(action, info={'legal_moves', 'pi'})."""
        actions, codes, pi, n_legal = self.select_actions_packed(rules.states_from_fens(list(observations)))
        return [ActionData(action=int(actions[k]),
                           info={'legal_moves': codes[k, :n_legal[k]].astype(int).tolist(), 'pi': pi[k, :n_legal[k]].copy()})
                for k in range(len(actions))]
