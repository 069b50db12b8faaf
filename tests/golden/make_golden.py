"""Generate the committed golden fixtures from the reference's own UNMODIFIED code.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py

What is recorded, and from what:
  moves_dict.json.sha256   sha256 of /root/reference/exp/moves_dict.json (byte-compat check)
  rules_positions.json.gz  reference exp/environment.py (MinitChessEpisode) on oracle/shims/chess:
                           legal codes, result, reward/done, FEN after every legal step
  perft.json               perft(1..6) from STARTING_FEN through the same episode class
  tokens.json              reference Network.process_observation (exp/policy.py:96-105)
  network_seed0.npz        reference Network() under torch.manual_seed(0): forward on 24 positions
  mcts_hash_game.json      reference MonteCarloTreeSearch/SimpleAlphaZeroAgent (exp/agent.py) self-play
                           game with the torch-free hash evaluator, np.random.seed(0), 36 sims
  mcts_net_game.json       same with the real random-init Network (config 1 of BASELINE.json)
  restatement_pin.json     record that oracle/ref_selfplay.py reproduced both games exactly
  learner_golden.json.gz   reference SimpleAlphaZeroLearner.update (exp/learner.py:72-94) on a fixed 256-tuple dataset,
                           CPU (`.cuda()` patched to a no-op IN THIS GENERATOR ONLY), torch.manual_seed(0): the batches
                           the DataLoader drew, every mini-batch loss, per-tensor sums of the trained state_dict --
                           for lr 0.2 (app/learner.py:69) and lr 1e-3

The rules fixtures pin the *oracle stack* (shim + reference wrapper), not the absent
python-chess fork: rules parity with the fork stays UNPINNED.
"""
import gzip
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle.ref_runner import load_reference, REFERENCE_ROOT  # noqa: E402
from oracle import rules_c, ref_selfplay  # noqa: E402
from oracle.hash_eval import HashModel, hash_evaluate  # noqa: E402

agent_mod, env_mod, policy_mod = load_reference()
import erlyx  # noqa: E402  (shim, on sys.path after load_reference)
import erlyx.callbacks  # noqa: E402

HAND_PICKED = [
    '2nbk/2ppp/5/5/PPP2/KBN2 w 0 1',      # start
    '2nbk/2ppp/5/5/PPP2/KBN2 b 0 1',      # start, black to move (tokeniser symmetry)
    '1k3/4P/5/5/5/K4 w 0 10',             # promotion available (e5e6 -> queen)
    '5/5/5/5/p4/K3k b 0 12',              # black promotion by capture impossible, push a2a1
    'k4/5/1QK2/5/5/5 b 3 20',             # stalemate? (checked by the oracle, recorded as is)
    'k4/1Q3/2K2/5/5/5 b 3 20',            # checkmate
    '4k/5/5/5/5/K4 w 10 15',              # K v K insufficient material
    '4k/5/5/5/5/KN3 w 10 15',             # K+N v K
    '4k/5/5/5/5/KB3 w 10 15',             # K+B v K
    '3bk/5/5/5/5/KB3 w 10 15',            # K+B v K+B
    '2nbk/2ppp/5/5/PPP2/KBN2 b 4 30',     # 30-move boundary: black's 30th move
    '2nbk/2ppp/5/5/PPP2/KBN2 w 4 31',     # beyond the cap
    '2nbk/2ppp/5/5/PPP2/KBN2 w 4 30',
    'q3k/5/5/5/5/K3Q w 0 8',
    'r3k/5/5/5/P4/K3R w 0 8',
]


def dump(name, obj, gz=False):
    path = os.path.join(HERE, name)
    text = json.dumps(obj, separators=(',', ':'))
    if gz:
        with gzip.GzipFile(path, 'wb', mtime=0) as f:
            f.write(text.encode())
    else:
        with open(path, 'w') as f:
            f.write(text)
    print('wrote', name, os.path.getsize(path), 'bytes')


def episode_record(fen):
    ep = env_mod.MinitChessEpisode(fen)
    rec = {'fen': ep.get_observation(), 'legal': list(ep.get_legal_moves()), 'result': ep._board.result(),
           'done': bool(ep.is_done()), 'reward': ep.get_reward() if ep.is_done() else None}
    kids = []
    if not ep.is_done():
        for code in rec['legal']:
            e2 = env_mod.MinitChessEpisode(fen)
            st = e2.step(code)
            kids.append([st.observation, st.reward if st.done else None, bool(st.done)])
    rec['children'] = kids
    return rec


def gen_rules():
    pos = rules_c.random_positions(20240601, 400000)
    rng = np.random.RandomState(7)
    # a spread of ordinary positions plus every finished position type we can find
    codes, counts, results = rules_c.legal_moves(pos)
    picks = list(rng.choice(len(pos), 1200, replace=False))
    for r in (1, 2, 3):
        idx = np.nonzero(results == r)[0]
        picks += list(idx[:60])
    picks += list(np.argsort(-counts)[:40])           # the widest positions
    fens = HAND_PICKED + [rules_c.state_to_fen(pos[i]) for i in sorted(set(int(i) for i in picks))]
    recs = [episode_record(f) for f in fens]
    dump('rules_positions.json.gz', recs, gz=True)
    return [r['fen'] for r in recs]


def perft(fen, depth):
    ep = env_mod.MinitChessEpisode(fen)
    if depth == 0:
        return 1
    if ep.is_done():
        return 0
    if depth == 1:
        return len(ep.get_legal_moves())
    total = 0
    for code in ep.get_legal_moves():
        e2 = env_mod.MinitChessEpisode(fen)
        total += perft(e2.step(code).observation, depth - 1)
    return total


def gen_perft():
    out = {'fen': env_mod.STARTING_FEN, 'nodes': [perft(env_mod.STARTING_FEN, d) for d in range(1, 6)]}
    dump('perft.json', out)


def gen_tokens(fens):
    rows = []
    for fen in fens[:160]:
        ch, clk = policy_mod.Network.process_observation(fen)
        rows.append({'fen': fen, 'tokens': ch.reshape(-1).tolist(), 'clock_f32_hex': np.float32(clk.item()).tobytes().hex()})
    dump('tokens.json', rows)


def state_dict_digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def gen_network(fens):
    torch.manual_seed(0)
    net = policy_mod.Network().eval()
    sd = net.state_dict()
    sel = fens[:24]
    with torch.no_grad():
        ins = [policy_mod.Network.process_observation(f) for f in sel]
        ch = torch.cat([c for c, _ in ins]); clk = torch.cat([k for _, k in ins])
        p, v = net((ch, clk))
        # a second network with non-trivial BatchNorm statistics / affine (exercises BN folding)
        torch.manual_seed(1)
        net2 = policy_mod.Network()
        for m in net2.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5)
                m.weight.data.uniform_(0.7, 1.3); m.bias.data.normal_(0, 0.1)
        net2.eval()
        p2, v2 = net2((ch, clk))
    np.savez_compressed(os.path.join(HERE, 'network_seed0.npz'), fens=np.array(sel), tokens=ch.numpy().astype(np.uint8),
                        clocks=clk.numpy(), logits=p.numpy(), values=v.numpy(), logits_bn=p2.numpy(), values_bn=v2.numpy())
    meta = {'seed0_state_dict_sha256': state_dict_digest(sd), 'n_params': sum(p.numel() for p in net.parameters()),
            'keys': [[k, list(v.shape)] for k, v in sd.items()]}
    dump('network_meta.json', meta)
    print('wrote network_seed0.npz')
    return net


class Recorder(erlyx.callbacks.BaseCallback):
    def __init__(self, agents):
        self.agents, self.rows, self.obs, self.turn = agents, [], None, 0

    def on_episode_begin(self, obs):
        self.obs = obs

    def on_step_end(self, action, observation, reward, done):
        tree = self.agents[self.turn]._mcts
        self.rows.append({'observation': self.obs, 'legal_moves': list(action.info['legal_moves']),
                          'N': tree['N'][self.obs].tolist(), 'Q': tree['Q'][self.obs].tolist(),
                          'pi': action.info['pi'].tolist(), 'action': int(action.action),
                          'next': observation, 'reward': reward, 'done': bool(done)})
        self.obs, self.turn = observation, self.turn ^ 1


def tree_digest(tree_N, tree_Q, terminal):
    h = hashlib.sha256()
    for k in sorted(tree_N):
        h.update(k.encode()); h.update(np.asarray(tree_N[k], dtype=np.float64).tobytes())
        h.update(np.asarray(tree_Q[k], dtype=np.float64).tobytes())
    for k in sorted(terminal):
        h.update(k.encode()); h.update(np.float64(terminal[k]).tobytes())
    return h.hexdigest()


def reference_game(model, sims, seed):
    np.random.seed(seed)
    env = env_mod.MinitChessEnvironment()
    pol = policy_mod.SimpleAlphaZeroPolicy(model)
    agents = [agent_mod.SimpleAlphaZeroAgent(env, pol, sims) for _ in range(2)]
    rec = Recorder(agents)
    with torch.no_grad():
        erlyx.run_episodes(env, agent_mod.RoundRobinReferee(tuple(agents)), 1,
                           callbacks=[rec, *(__import__('exp.callbacks').callbacks.MonteCarloInit(a) for a in agents)], use_tqdm=False)
    digests = [tree_digest(a._mcts['N'], a._mcts['Q'], a._mcts['terminal']) for a in agents]
    sizes = [[len(a._mcts['N']), len(a._mcts['terminal'])] for a in agents]
    return rec.rows, digests, sizes


def restated_game(evaluate, sims, seed):
    np.random.seed(seed)
    records, ep, trees = ref_selfplay.play_game(evaluate, sims)
    digests = [tree_digest(t.N, t.Q, t.terminal) for t in trees]
    return records, digests


def gen_mcts(net):
    pins = {}
    for name, model, evaluate, sims, seed in (
            ('mcts_hash_game.json', HashModel(), hash_evaluate, 36, 0),
            ('mcts_hash_game_s1.json', HashModel(), hash_evaluate, 50, 1),
            ('mcts_net_game.json', net, ref_selfplay.RefNetwork(net.state_dict()).evaluate, 36, 0)):
        rows, digests, sizes = reference_game(model, sims, seed)
        dump(name, {'sims': sims, 'seed': seed, 'plies': rows, 'tree_sha256': digests, 'tree_sizes': sizes})
        r2, d2 = restated_game(evaluate, sims, seed)
        same = (len(r2) == len(rows) and d2 == digests and
                all(a['action'] == b['action'] and a['pi'] == b['pi'] and a['observation'] == b['observation']
                    for a, b in zip(r2, rows)))
        pins[name] = {'restatement_identical': bool(same), 'plies': len(rows)}
        assert same, name
    dump('restatement_pin.json', pins)


class LoggingDataset:
    """What SimpleAlphaZeroDataset is to the DataLoader (exp/dataset.py:6-20), remembering the order of the reads."""

    def __init__(self, items):
        self.items, self.reads = items, []

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        self.reads.append(int(i))
        return self.items[i]


def learner_dataset(fens, n=256, seed=5):
    """Replay tuples in InfoRecorder's format (exp/callbacks.py:40-53) on golden positions: pi a seeded Dirichlet over the
    legal moves, rewards from {-1, 0, 1}.  The learner does not care where tuples come from."""
    rng = np.random.RandomState(seed)
    items = []
    for fen in fens:
        ep = env_mod.MinitChessEpisode(fen)
        legal = list(ep.get_legal_moves())
        if ep.is_done() or not legal:
            continue
        pi = rng.dirichlet([0.6] * len(legal))
        items.append({'observation': fen, 'legal_moves': legal, 'pi': pi.tolist(), 'action': int(rng.choice(legal)),
                      'reward': float(rng.choice([-1.0, 0.0, 1.0]))})
        if len(items) == n:
            break
    assert len(items) == n
    return items


def gen_learner(fens):
    import importlib
    cwd = os.getcwd()
    os.chdir('/tmp')                               # exp/learner.py:3-6 opens a log file in the cwd
    try:
        learner_mod = importlib.import_module('exp.learner')
    finally:
        os.chdir(cwd)
    items = learner_dataset(fens)
    runs = {}
    orig_module_cuda, orig_tensor_cuda = torch.nn.Module.cuda, torch.Tensor.cuda
    torch.nn.Module.cuda = lambda self, *a, **k: self          # exp/learner.py:79,86 -- no GPU in the build container
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for lr in (0.2, 1e-3):
            torch.manual_seed(0)
            net = policy_mod.Network()
            data = LoggingDataset(items)
            learner = learner_mod.SimpleAlphaZeroLearner(env_mod.MinitChessEnvironment(), 36, net, batch_size=32, epochs=1,
                                                         optim_params={'lr': lr})
            losses = []
            orig_acc = learner_mod.AvgSmoothLoss.accumulate

            def acc(self, v, _l=losses, _o=orig_acc):
                _l.append(float(v))
                return _o(self, v)
            learner_mod.AvgSmoothLoss.accumulate = acc
            try:
                learner.update(data)
            finally:
                learner_mod.AvgSmoothLoss.accumulate = orig_acc
            sd = net.state_dict()
            runs[repr(lr)] = {'lr': lr, 'batches': [data.reads[i:i + 32] for i in range(0, len(data.reads), 32)], 'losses': losses,
                              'sums': {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()
                                       if not k.endswith('num_batches_tracked')}}
    finally:
        torch.nn.Module.cuda, torch.Tensor.cuda = orig_module_cuda, orig_tensor_cuda
    dump('learner_golden.json.gz', {'items': items, 'seed': 0, 'batch_size': 32, 'runs': runs}, gz=True)


def main():
    if '--only-learner' in sys.argv:
        recs = json.load(gzip.open(os.path.join(HERE, 'rules_positions.json.gz'), 'rt'))
        gen_learner([r['fen'] for r in recs])
        return
    with open(os.path.join(REFERENCE_ROOT, 'exp', 'moves_dict.json'), 'rb') as f:
        blob = f.read()
    dump('moves_dict.json.sha256', {'sha256': hashlib.sha256(blob).hexdigest(), 'bytes': len(blob)})
    fens = gen_rules()
    gen_perft()
    gen_tokens(fens)
    net = gen_network(fens)
    gen_mcts(net)
    gen_learner(fens)


if __name__ == '__main__':
    main()
