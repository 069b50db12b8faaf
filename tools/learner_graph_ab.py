"""loop.learner_update: the optimiser step launched kernel by kernel against the same step replayed as one CUDA graph.
Same synthetic replay tensors, same mini-batches, same seed: ms per step and the largest difference of the losses.
    python tools/learner_graph_ab.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minitchess_alphazero_b200.loop import learner_update
from minitchess_alphazero_b200.policy import Network

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = 32
n = steps * B
g = torch.Generator().manual_seed(0)
pi = torch.rand(n, 554, generator=g); pi /= pi.sum(1, keepdim=True)
channels = torch.stack([torch.randint(0, 7, (n, 6, 5), generator=g), torch.randint(0, 2, (n, 6, 5), generator=g)], 1)
clock = torch.rand(n, 1, generator=g)
reward = torch.randint(-1, 2, (n, 1), generator=g).float()
data = [t.cuda() for t in (pi, channels, clock, reward)]
order = [list(range(i, i + B)) for i in range(0, n, B)]
out = {}
for lr in (1e-3, 0.2):
    net_kept = None
    for graph in (False, True, False, True, True):
        torch.manual_seed(0)
        net = Network()
        if graph:            # the second and third graphed updates run on the first one's network: the captured step is reused
            if net_kept is None:
                net_kept = net.cuda()
            else:
                net_kept.load_state_dict(net.state_dict()); net = net_kept
        torch.cuda.synchronize(); t = time.perf_counter()
        losses = learner_update(net, data, batch_size=B, optim_params={'lr': lr}, order=order, graph=graph)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        out[(lr, graph)] = losses
        print('lr %g graph %-5s: %d steps, %.3f ms per step (%.2f s), first losses %s' % (lr, graph, len(losses), 1e3 * dt / len(losses), dt,
                                                                                 ['%.5f' % x for x in losses[:3]]))
    a, b = out[(lr, False)], out[(lr, True)]
    print('lr %g: largest |loss difference| over the first 8 steps %.3g, over all %.3g' % (
        lr, max(abs(x - y) for x, y in zip(a[:8], b[:8])), max(abs(x - y) for x, y in zip(a, b))))
