"""CPU: the facade's torch Network, tokeniser, action table and weight flattening against the
golden vectors made from the reference's own exp/policy.py / moves_dict.json."""
import hashlib

import numpy as np
import torch

from conftest import load_golden


def test_network_seed0_identical_to_reference(mcaz_lib):
    from minitchess_alphazero_b200.policy import Network, flatten_state_dict
    meta = load_golden('network_meta.json')
    torch.manual_seed(0)
    net = Network().eval()
    sd = net.state_dict()
    assert [[k, list(v.shape)] for k, v in sd.items()] == meta['keys']
    assert sum(p.numel() for p in net.parameters()) == meta['n_params'] == 10693458
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode()); h.update(v.numpy().tobytes())
    assert h.hexdigest() == meta['seed0_state_dict_sha256']
    g = load_golden('network_seed0.npz')
    with torch.no_grad():
        p, v = net((torch.from_numpy(g['tokens'].astype(np.int64)), torch.from_numpy(g['clocks'])))
    assert np.array_equal(p.numpy(), g['logits']) and np.array_equal(v.numpy(), g['values'])
    flat = flatten_state_dict(sd)
    assert flat.numel() == 10693458 + 9734


def test_oracle_network_restatement_matches_golden():
    from oracle.ref_selfplay import RefNetwork
    from minitchess_alphazero_b200.policy import Network
    g = load_golden('network_seed0.npz')
    torch.manual_seed(0)
    rn = RefNetwork(Network().state_dict())
    with torch.no_grad():
        p, v = rn.forward(torch.from_numpy(g['tokens'].astype(np.int64)), torch.from_numpy(g['clocks']))
    assert np.allclose(p.numpy(), g['logits'], rtol=1e-5, atol=1e-6)
    assert np.allclose(v.numpy(), g['values'], rtol=1e-5, atol=1e-6)
    for fen, tok in zip(g['fens'][:8], g['tokens'][:8]):
        t, c = RefNetwork.tokenize_fen(str(fen))
        assert np.array_equal(t.numpy().astype(np.uint8), tok[None])


def test_tokeniser_and_moves_table(mcaz_lib):
    from minitchess_alphazero_b200.policy import Network
    from minitchess_alphazero_b200 import moves
    for r in load_golden('tokens.json'):
        ch, clk = Network.process_observation(r['fen'])
        assert ch.shape == (1, 2, 6, 5) and ch.dtype == torch.int64
        assert ch.reshape(-1).tolist() == r['tokens']
        assert np.float32(clk.item()).tobytes().hex() == r['clock_f32_hex']
    meta = load_golden('moves_dict.json.sha256')
    text = moves.as_json_text().encode()
    assert len(text) == meta['bytes'] and hashlib.sha256(text).hexdigest() == meta['sha256']
    assert moves.NUM_ACTIONS == 554 and len(moves.MOVES_DICT_INV[False]) == 554
