"""GPU, T1 pin of the PRODUCTION path at bench size: `az_search` / `az_search_noise` with the built-in network, the exact
evaluation cache on, `free_sims` 4 and recycling on -- the configuration `bench.py` times (`search_step_kernel`,
`heads_legal_kernel`, the cache, `recycle_kernel`, `tower_tc_kernel`) -- against the reference MCTS of exp/agent.py:41-88
(its pinned restatement oracle/ref_selfplay.RefTree) on 64 concurrent games x 200 simulations per move x 10 plies.

Bit-exact visit counts need identical evaluator outputs on both sides (SURVEY.md 7.3 point 2), so the reference tree is fed
the engine's own bits: the priors the engine stored in each node (`az_tree_dump`) and the value `az_network_forward` gives for
that position (same head arithmetic, row-independent, deterministic: tests/test_gpu_network.py).  Then every node the
reference search holds must exist in the engine's tree with the same legal codes, the same N and the same float64 Q bits,
the engine must hold no node the reference lacks, the numpy RNG streams of both sides must have advanced identically
(one dirichlet per simulation whose root is expanded, one choice per move: exp/agent.py:81-82,115,118) and the games played
move by move must be the same."""
import numpy as np
import pytest
import torch

from oracle import ref_selfplay as rs
from oracle import rules_c as rc

pytestmark = pytest.mark.gpu
ALPHA = 0.6


@pytest.fixture(scope='module')
def net(mcaz_lib):
    from minitchess_alphazero_b200.policy import Network
    torch.manual_seed(0)
    return Network().eval()


def ply_of(meta):
    return 2 * int((meta >> 16) & 0xff) + (0 if (meta & 1) else 1)


def reachable(fen_or_meta_ply, fen, root_fen, root_ply):
    """What recycling keeps and the search can still reach: the root and everything at a later ply (the reference never prunes
    its dicts; the engine drops plies <= the root's, other than the root, when an arena fills up -- SURVEY.md 7.3 point 7)."""
    return fen == root_fen or fen_or_meta_ply > root_ply


def dump_by_fen(eng, g, tree, values, root_fen, root_ply):
    """{fen: node record} of the root and the nodes beyond its ply, straight from the arenas."""
    from minitchess_alphazero_b200.engine import NODE_DECISIVE, NODE_TERMINAL
    d = eng.tree_dump(g, tree)
    out, need = {}, []
    for i, s in enumerate(d['states']):
        fen = rc.state_to_fen(s)
        if not reachable(ply_of(int(s['meta'])), fen, root_fen, root_ply):
            continue                                          # an older ply the recycler has not dropped yet
        info, off = int(d['info'][i]), int(d['edge_off'][i])
        E = info & 0xffff
        term = bool(info & NODE_TERMINAL)
        rec = {'codes': d['codes'][off:off + E].astype(int).tolist(), 'N': d['N'][off:off + E].astype(np.float64),
               'Q': d['Q'][off:off + E].copy(), 'P': d['P'][off:off + E].copy(),
               'terminal': (-1.0 if info & NODE_DECISIVE else -0.0) if term else None, 'state': s}
        assert fen not in out, 'one position, two nodes: ' + fen
        out[fen] = rec
        if not term and fen not in values:
            need.append(fen)
    return out, need


@pytest.mark.parametrize('mode,G,sims,plies', [('caller_noise', 64, 200, 10), ('no_noise', 16, 200, 6)])
def test_production_search_equals_reference_mcts_at_bench_size(net, mode, G, sims, plies):
    from minitchess_alphazero_b200._lib import MC_MAX_MOVES
    from minitchess_alphazero_b200.engine import Engine
    from minitchess_alphazero_b200.policy import flatten_state_dict
    eps = 0.25 if mode == 'caller_noise' else 0.0
    # bench.py's engine options; arenas of 3 x sims nodes so that recycle_kernel compacts trees several times in 10 plies
    eng = Engine(G, max_sims_per_move=sims, network=1, eval_cache_log2=20, free_sims=4, recycle=1, node_capacity=3 * sims + 64,
                 dirichlet_epsilon=eps, device_rng=0)
    eng.set_weights(flatten_state_dict(net.state_dict(), device='cuda'))
    rng_eng = [np.random.RandomState(1000 + g) for g in range(G)]        # the engine side's host RNG, one stream per game
    rng_ref = [np.random.RandomState(1000 + g) for g in range(G)]        # the reference agents' `np.random`
    values = {}                                                          # fen -> float32 value the engine's network gives
    nodes = [{} for _ in range(G)]                                       # per game: what the engine's tree says, by fen

    def evaluator(g):
        def evaluate(fen, legal):                                         # exp/agent.py:67-69 with the engine's own bits
            rec = nodes[g][fen]                                           # KeyError = the engine never expanded this position
            assert rec['codes'] == list(legal), fen
            return rec['P'], float(values[fen])
        return evaluate
    trees = [[rs.RefTree(evaluator(g), 1, epsilon=eps, rng=rng_ref[g]) for _ in range(2)] for g in range(G)]
    episodes = [rs.RefEpisode(rs.STARTING_FEN) for _ in range(G)]
    compared = 0
    for ply in range(plies):
        states, results = eng.game_states()
        assert (results == 0).all()                                       # 10 plies from the start: nobody is done yet
        _, counts, _ = rc.legal_moves(states)
        _, _, _, seen = eng.root_stats(want_q=False)                      # -1: this agent's tree has not met the root yet
        if eps > 0:
            noise = np.zeros((sims, G, MC_MAX_MOVES))
            for g in range(G):
                E, first = int(counts[g]), (0 if seen[g] >= 0 else 1)      # an unseen root is only expanded by simulation 0
                noise[first:, g, :E] = rng_eng[g].dirichlet([ALPHA] * E, size=sims - first)
            eng.search_noise(noise)
        else:
            eng.search(sims)
        # what the engine built: trees of the side to move, values of every new position in one bulk forward
        t = ply & 1
        need = []
        for g in range(G):
            nodes[g], more = dump_by_fen(eng, g, t, values, rc.state_to_fen(states[g]), ply_of(int(states[g]['meta'])))
            need += more
        need = sorted(set(need))
        if need:
            st = rc.fens_to_states(need)
            tok, clk = rc.tokenize(st)
            _, v = eng.network_forward(tok, clk)
            values.update(zip(need, v))
        # the reference search on the same bits, then node by node
        codes, visits, q, n_legal = eng.root_stats()
        actions = np.zeros(G, dtype=np.uint16)
        for g in range(G):
            fen = rc.state_to_fen(states[g])
            assert fen == episodes[g].fen
            tree = trees[g][t]
            action, info = rs.ref_select_action(tree, fen, sims, rng=rng_ref[g])
            root_ply = ply_of(int(states[g]['meta']))
            mine = nodes[g]
            ref_fens = [f for f in list(tree.N) + list(tree.terminal)
                        if reachable(2 * int(f.split()[3]) + (0 if f.split()[1] == 'w' else 1), f, fen, root_ply)]
            assert set(ref_fens) == set(mine), (ply, g, len(ref_fens), len(mine))          # same positions, no more, no fewer
            for f in ref_fens:
                rec = mine[f]
                if f in tree.terminal:
                    assert rec['terminal'] is not None and np.float64(rec['terminal']).tobytes() == np.float64(tree.terminal[f]).tobytes()
                    continue
                assert rec['codes'] == list(tree.legal[f]), f
                assert np.array_equal(rec['N'], tree.N[f]), (ply, g, f)                    # visit counts: exact
                assert rec['Q'].tobytes() == np.asarray(tree.Q[f], dtype=np.float64).tobytes(), (ply, g, f)   # Q: float64 bits
                assert rec['P'].tobytes() == np.asarray(tree.P[f], dtype=np.float32).tobytes()
                compared += 1
            # the move: same RNG state on both sides, same choice (exp/agent.py:113-118)
            E = int(n_legal[g])
            legal = codes[g, :E].astype(int).tolist()
            N = visits[g, :E].astype(np.float64)
            pi = N / N.sum()
            assert legal == list(info['legal_moves']) and pi.tobytes() == np.asarray(info['pi']).tobytes()
            if int(fen.split()[3]) < 6:
                mine_action = rng_eng[g].choice(legal, p=pi)
            else:
                best = np.where(pi == pi.max())[0]
                mine_action = legal[rng_eng[g].choice(best)]
            assert int(mine_action) == action
            a, b = rng_eng[g].get_state(), rng_ref[g].get_state()
            assert a[2] == b[2] and np.array_equal(a[1], b[1])          # both streams consumed the same number of draws
            actions[g] = action
            episodes[g].step(action)
        eng.play(actions)
    c = eng.counters()
    assert c['simulations'] == G * sims * plies
    assert c['simulations'] == c['evaluations'] + c['terminal_leaves'] + c['cached_evaluations']
    assert c['cached_evaluations'] > 0.05 * c['simulations']             # the exact cache answered leaves (two agents per game, transpositions)
    assert c['recycled_nodes'] > 0                                       # recycle_kernel compacted trees in between
    assert compared > 0.5 * c['nodes']
