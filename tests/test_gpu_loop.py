"""GPU: replay tuples -> device collate -> learner step -> weights back into the engine (rows (f) of the
scope table: replay format, weight exchange, one full loop iteration)."""
import numpy as np
import pytest
import torch

from oracle import rules_c as rc
from oracle import ref_selfplay as rs

pytestmark = pytest.mark.gpu


def python_collate(batch):
    """exp/learner.py:23-41 restated (the reference's collate_fn imports erlyx, which is absent here)."""
    pib, chb, clb, rwb = [], [], [], []
    for item in batch:
        pi = torch.zeros(554).float()
        pi[item['legal_moves']] = torch.FloatTensor(item['pi'])
        pib.append(pi)
        ch, clk = rs.RefNetwork.tokenize_fen(item['observation'])
        chb.append(ch); clb.append(clk); rwb.append(item['reward'])
    return [torch.vstack(pib), torch.cat(chb, 0), torch.FloatTensor(clb).reshape(-1, 1), torch.FloatTensor(rwb).reshape(-1, 1)]


def test_collate_device_matches_reference_collate(mcaz_lib):
    from minitchess_alphazero_b200.policy import Network
    from minitchess_alphazero_b200.selfplay import BatchedSelfPlay, collate_device, replay_to_episode_dicts
    torch.manual_seed(0)
    sp = BatchedSelfPlay(Network().eval(), n_games=64, num_simulations=6, seed=3)
    sp.run(64)
    tuples = sp.drain()
    assert len(tuples) > 200
    got = collate_device(tuples[:300])
    want = python_collate(replay_to_episode_dicts(tuples[:300]))
    assert torch.equal(got[0].cpu(), want[0])                    # dense pi, float32
    assert torch.equal(got[1].cpu(), want[1]) and got[1].dtype == torch.int64 and tuple(got[1].shape[1:]) == (2, 6, 5)
    assert torch.equal(got[2].cpu(), want[2])
    assert torch.equal(got[3].cpu(), want[3])
    empty = collate_device(tuples[:0])
    assert empty[0].shape == (0, 554)


def test_one_loop_iteration_updates_engine_weights(mcaz_lib):
    from minitchess_alphazero_b200.loop import iteration
    from minitchess_alphazero_b200.policy import Network
    from minitchess_alphazero_b200.selfplay import BatchedSelfPlay
    torch.manual_seed(0)
    net = Network().eval()
    sp = BatchedSelfPlay(net, n_games=64, num_simulations=6, seed=5)
    pos = np.ascontiguousarray(rc.random_positions(9, 400)[:64])
    tokens, clocks = rc.tokenize(pos)
    before, _ = sp.engine.network_forward(tokens, clocks)
    out = iteration(sp, net, 62, batch_size=32, optim_params={'lr': 1e-3})
    assert out['tuples'] > 0 and len(out['losses']) > 0 and all(np.isfinite(out['losses']))
    after, v_after = sp.engine.network_forward(tokens, clocks)
    assert not np.allclose(before, after)                        # the engine runs on the new weights
    with torch.no_grad():                                        # ... and they are the learner's weights
        p_ref, v_ref = rs.RefNetwork({k: v.cpu() for k, v in net.state_dict().items()}).forward(
            torch.from_numpy(tokens.astype(np.int64)).view(-1, 2, 6, 5), torch.from_numpy(clocks).view(-1, 1))
    assert np.abs(torch.from_numpy(after).softmax(-1).numpy() - p_ref.softmax(-1).numpy()).max() < 1e-2
    assert np.abs(v_after - v_ref.numpy().reshape(-1)).max() < 2e-2
    sp.run(2)                                                    # self-play continues on the updated network


def test_arena_bookkeeping(mcaz_lib):
    """Row (f)-4: new-vs-old arena on one engine; each side's tree is evaluated with its own network."""
    from minitchess_alphazero_b200.arena import Arena
    from minitchess_alphazero_b200.policy import Network
    torch.manual_seed(0)
    net_a = Network().eval()
    torch.manual_seed(1)
    net_b = Network().eval()
    arena = Arena(net_a, net_b, games_per_side=16, num_simulations=6, seed=3)
    out = arena.play()
    assert out['a'] + out['b'] + out['draws'] == 32 and 0.0 <= out['a_score'] <= 1.0 and out['plies'] <= 61
    c = arena.engine.counters()
    assert c['games_finished'] == 32 and c['simulations'] > 0
    # both networks were really used: their evaluations of the start position differ
    from oracle import rules_c as rc
    tok, clk = rc.tokenize(rc.start_state())
    l0, _ = arena._sides[0].engine.network_forward(tok, clk)
    l1, _ = arena._sides[1].engine.network_forward(tok, clk)
    assert not np.allclose(l0, l1)
