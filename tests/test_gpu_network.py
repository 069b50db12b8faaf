"""GPU: the hand-written sm_100a network (stem table kernel, tcgen05 conv tower, heads) against the
reference Network's fp32 outputs (golden, from exp/policy.py) -- T2 parity, bf16 tolerance 1e-2."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import rules_c as rc
from oracle import ref_selfplay as rs

pytestmark = pytest.mark.gpu
TOL = 1e-2


def make_engine(net, n_games=256):
    from minitchess_alphazero_b200.engine import Engine
    from minitchess_alphazero_b200.policy import flatten_state_dict
    eng = Engine(n_games, max_sims_per_move=4, network=1)
    eng.set_weights(flatten_state_dict(net.state_dict()).numpy())
    return eng


def bn_perturbed_net():
    from minitchess_alphazero_b200.policy import Network
    torch.manual_seed(1)
    net = Network()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5)
                m.weight.data.uniform_(0.7, 1.3); m.bias.data.normal_(0, 0.1)
    return net.eval()


def test_fused_tower_kernel_matches_per_layer(mcaz_lib, monkeypatch):
    """The default runs the 18 convolutions as one data-flow-ordered persistent kernel; MCAZ_TOWER=layers
    launches one kernel per layer.  Results must be bit-identical (same MMAs, same order of accumulation)."""
    from minitchess_alphazero_b200.policy import Network
    torch.manual_seed(0)
    net = Network().eval()
    pos = np.ascontiguousarray(rc.random_positions(5, 3000)[:1000])
    tokens, clocks = rc.tokenize(pos)
    monkeypatch.setenv('MCAZ_TOWER', 'layers')
    a = make_engine(net, n_games=1000)
    la, va = a.network_forward(tokens, clocks)
    monkeypatch.delenv('MCAZ_TOWER')
    b = make_engine(net, n_games=1000)
    for _ in range(3):                      # repeated launches: epoch-stamped dependency flags
        lb, vb = b.network_forward(tokens, clocks)
        assert np.array_equal(la, lb) and np.array_equal(va, vb)
    lb, vb = b.network_forward(tokens[:300], clocks[:300])
    assert np.array_equal(la[:300], lb) and np.array_equal(va[:300], vb)


def test_golden_forward_seed0(mcaz_lib):
    from minitchess_alphazero_b200.policy import Network
    g = load_golden('network_seed0.npz')
    torch.manual_seed(0)
    eng = make_engine(Network().eval())
    logits, values = eng.network_forward(g['tokens'].reshape(-1, 60), g['clocks'].reshape(-1))
    p = torch.from_numpy(logits).softmax(-1).numpy()
    p_ref = torch.from_numpy(g['logits']).softmax(-1).numpy()
    assert np.abs(logits - g['logits']).max() < 2e-2
    assert np.abs(p - p_ref).max() < TOL and np.abs(p / p_ref - 1).max() < 5e-2
    assert np.abs(values - g['values'].reshape(-1)).max() < TOL


def test_golden_forward_with_batchnorm_statistics(mcaz_lib):
    """Non-trivial running stats / affine: exercises the BatchNorm folding of every layer."""
    g = load_golden('network_seed0.npz')
    eng = make_engine(bn_perturbed_net())
    logits, values = eng.network_forward(g['tokens'].reshape(-1, 60), g['clocks'].reshape(-1))
    p = torch.from_numpy(logits).softmax(-1).numpy()
    p_ref = torch.from_numpy(g['logits_bn']).softmax(-1).numpy()
    assert np.abs(p - p_ref).max() < TOL
    assert np.abs(values - g['values_bn'].reshape(-1)).max() < 2 * TOL


def test_batch_sizes_and_row_independence(mcaz_lib):
    """Ragged batches (not multiples of the 128-board tile) and the same position at different rows."""
    from minitchess_alphazero_b200.policy import Network
    torch.manual_seed(0)
    net = Network().eval()
    eng = make_engine(net, n_games=700)
    pos = np.ascontiguousarray(rc.random_positions(77, 3000)[:700])
    tokens, clocks = rc.tokenize(pos)
    with torch.no_grad():
        p_ref, v_ref = rs.RefNetwork(net.state_dict()).forward(torch.from_numpy(tokens.astype(np.int64)).view(-1, 2, 6, 5),
                                                               torch.from_numpy(clocks).view(-1, 1))
    p_ref = p_ref.softmax(-1).numpy()
    full_l, full_v = eng.network_forward(tokens, clocks)
    assert np.abs(torch.from_numpy(full_l).softmax(-1).numpy() - p_ref).max() < TOL
    assert np.abs(full_v - v_ref.numpy().reshape(-1)).max() < TOL
    for n in (1, 5, 127, 129, 300):
        l, v = eng.network_forward(tokens[:n], clocks[:n])
        assert np.array_equal(l, full_l[:n]) and np.array_equal(v, full_v[:n])     # deterministic, row independent
    l, v = eng.network_forward(tokens[::-1].copy(), clocks[::-1].copy())
    assert np.array_equal(l[::-1], full_l) and np.array_equal(v[::-1], full_v)


def test_search_with_builtin_network_matches_torch_evaluator(mcaz_lib):
    """az_search (built-in net) and the external-evaluator path build the same kind of tree: root visit
    distributions agree closely on a no-noise search from the start position."""
    from minitchess_alphazero_b200.engine import Engine
    from minitchess_alphazero_b200.policy import Network, flatten_state_dict
    from torch_evaluator import TorchEvaluator
    torch.manual_seed(0)
    net = Network().eval()
    sims = 64
    a = Engine(4, max_sims_per_move=sims, network=1, dirichlet_epsilon=0.0)
    a.set_weights(flatten_state_dict(net.state_dict()).numpy())
    a.search(sims)
    ca, va, _, na = a.root_stats()
    b = Engine(4, max_sims_per_move=sims, dirichlet_epsilon=0.0)
    ev = TorchEvaluator(net, dtype=torch.float32)
    tokens, clocks, _ = b.leaf_batch_device()
    for _ in range(sims):
        b.select_expand()
        lg, vl = ev.forward(tokens, clocks)
        b.backup(vl, logits=lg)
    cb, vb, _, nb = b.root_stats()
    assert np.array_equal(na, nb) and np.array_equal(ca, cb)
    assert va[0, :na[0]].sum() == sims - 1
    assert np.abs(va[0, :6].astype(float) - vb[0, :6].astype(float)).max() <= 6
    assert a.counters()['evaluations'] > 0
