"""CPU probe for a round-2 decision: would an fp8 (e4m3) tower hold the 1e-2 tolerance BASELINE.json states for a
reduced-precision network?  Emulates the tower's arithmetic in torch on the CPU -- operands rounded to the storage type,
products accumulated in fp32 (what `tcgen05.mma` does), BatchNorm folded, residual path kept in bf16 -- and compares the
legal-move softmax priors and the values of the golden positions (tests/golden/rules_positions.json.gz, unfinished ones)
with the fp32 `Network`.  The bf16 emulation is the control: it must land where the real kernel does (tests: < 1e-2,
measured on the GPU ≈ 2e-3).  Nothing here runs on the product path.

  python tools/fp8_probe.py [--bn-stats] [--n 512]
"""
import argparse
import gzip
import json
import os
import sys

import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from minitchess_alphazero_b200.policy import Network  # noqa: E402

E4M3_MAX = 448.0


def fold(block):
    conv, bn = block.layers[0], block.layers[1]
    scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
    return conv.weight.detach() * scale.view(-1, 1, 1, 1), (conv.bias.detach() - bn.running_mean) * scale + bn.bias.detach()


def q_bf16(x):
    return x.to(torch.bfloat16).float()


def q_e4m3_weight(w):
    """per-output-channel scale (folded back into the epilogue's fp32 multiply)"""
    s = w.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-12) / E4M3_MAX
    return (w / s).to(torch.float8_e4m3fn).float() * s


def q_e4m3_act(x, scale):
    """per-layer static scale (calibrated), saturating"""
    return (x / scale).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).float() * scale


def tower(net, tokens, clocks, mode, act_scales=None, calibrate=None):
    """mode: 'fp32' | 'bf16' | 'fp8' (e4m3 conv operands, bf16 residual stream) | 'fp8w' (e4m3 weights, bf16 activations)"""
    qw = {'fp32': lambda w: w, 'bf16': q_bf16, 'fp8': q_e4m3_weight, 'fp8w': q_e4m3_weight}[mode]
    store = (lambda x: x) if mode == 'fp32' else q_bf16          # what the epilogue writes (residual stream)
    layer = [0]

    def operand(x):
        i = layer[0]
        layer[0] += 1
        if calibrate is not None:
            calibrate.append(float(x.abs().max()))
        if mode == 'fp8':
            return q_e4m3_act(x, act_scales[i] / E4M3_MAX)
        return x

    x = net.emb(tokens).permute(0, 1, 4, 2, 3).contiguous().view(-1, 8, 6, 5)
    w, b = fold(net.resbody[0])
    x = store(F.relu(F.conv2d(x, w, b, padding=1)))             # stem: a table of fp32 sums in the kernel
    for blk in list(net.resbody)[1:]:
        w1, b1 = fold(blk.convblock1)
        w2, b2 = fold(blk.convblock2)
        h = store(F.relu(F.conv2d(operand(x), qw(w1), b1, padding=1)))
        x = store(F.relu(F.conv2d(operand(h), qw(w2), b2, padding=1) + x))
    # heads in fp32 on the stored activations (the kernel dots the fp32 accumulator; this is the pessimistic side)
    wp, bp = fold(net.pconv)
    wv, bv = fold(net.vconv)
    p = net.plinear(torch.cat([F.relu(F.conv2d(x, wp, bp)).view(-1, 60), clocks], dim=1))
    v = net.vlinear(torch.cat([F.relu(F.conv2d(x, wv, bv)).view(-1, 30), clocks], dim=1))
    return p, v


def legal_softmax(logits, legal):
    out = []
    for row, codes in zip(logits, legal):
        out.append(torch.softmax(row[codes], dim=0))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=512)
    ap.add_argument('--bn-stats', action='store_true', help='non-trivial BatchNorm statistics and gains (a trained-looking net)')
    ap.add_argument('--seed', type=int, default=0)
    a = ap.parse_args()
    torch.manual_seed(a.seed)
    net = Network().eval()
    if a.bn_stats:
        g = torch.Generator().manual_seed(1)
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(0.2 * torch.randn(m.num_features, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
                m.weight.data.copy_(0.7 + 0.6 * torch.rand(m.num_features, generator=g))
                m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=g))
    rows = [r for r in json.load(gzip.open(os.path.join(REPO, 'tests', 'golden', 'rules_positions.json.gz'), 'rt'))
            if not r['done'] and r['legal']][:a.n]
    obs = [Network.process_observation(r['fen']) for r in rows]
    tokens = torch.cat([o[0] for o in obs])
    clocks = torch.cat([o[1] for o in obs])
    legal = [torch.tensor(r['legal']) for r in rows]
    out = {}
    with torch.no_grad():
        p0, v0 = tower(net, tokens, clocks, 'fp32')
        pr, vr = net((tokens, clocks))
        assert (p0 - pr).abs().max() < 1e-4 and (v0 - vr).abs().max() < 1e-5, 'emulation does not reproduce Network.forward'
        half = len(rows) // 2
        cal = []
        tower(net, tokens[:half], clocks[:half], 'fp32', calibrate=cal)            # activation scales from the first half
        ref = legal_softmax(p0, legal)
        for mode in ('bf16', 'fp8w', 'fp8'):
            p, v = tower(net, tokens, clocks, mode, act_scales=[1.25 * c for c in cal])
            sm = legal_softmax(p, legal)
            dp = torch.stack([(x - y).abs().max() for x, y in zip(sm[half:], ref[half:])])
            dv = (v - v0).abs().view(-1)[half:]
            out[mode] = {'prior_max': float(dp.max()), 'prior_mean': float(dp.mean()), 'value_max': float(dv.max()),
                         'value_mean': float(dv.mean()), 'logit_max': float((p - p0).abs().max())}
    print(json.dumps({'positions': len(rows) - half, 'bn_stats': a.bn_stats, 'seed': a.seed, 'tolerance': 1e-2, **out}, indent=1))


if __name__ == '__main__':
    main()
