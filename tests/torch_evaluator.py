"""Library-kernel evaluation of a `Network` (cuDNN / cuBLAS through PyTorch).  TEST INFRASTRUCTURE ONLY: the fp32 form is
the checker the hand-written sm_100a network and the drop-in agent are compared with on the GPU box (tests/test_gpu_network.py,
tests/test_gpu_facade.py); it plugs into `MonteCarloTreeSearch(evaluator=...)`, the external-evaluator seam of the C ABI
(az_select_expand / az_leaf_batch / az_backup).  Nothing in the package or bench.py imports it."""
import torch


class TorchEvaluator:
    """Batched leaf evaluation of a `Network` on the GPU with library kernels (cuDNN / cuBLAS via
    PyTorch): BatchNorm folded into the convolutions, channels-last, bf16 tower by default with
    fp32 heads.  This is the library baseline that the hand-written tcgen05 tower is measured
    against; no CPU path."""

    def __init__(self, network, dtype=torch.bfloat16, device='cuda'):
        if not torch.cuda.is_available():
            raise RuntimeError('TorchEvaluator needs a CUDA device (self-play has no CPU fallback)')
        self.device, self.dtype = torch.device(device), dtype
        self.load(network)
        self._graph = None
        torch.backends.cudnn.benchmark = True

    @staticmethod
    def _fold(block):
        conv, bn = block.layers[0], block.layers[1]
        scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
        w = conv.weight.detach() * scale.view(-1, 1, 1, 1)
        b = (conv.bias.detach() - bn.running_mean) * scale + bn.bias.detach()
        return w, b

    def load(self, network):
        dev, dt = self.device, self.dtype
        self.emb = network.emb.weight.detach().float().to(dev)

        def put(w, b, tower=True):
            w = w.float().to(dev)
            if tower:
                w = w.to(dt).contiguous(memory_format=torch.channels_last)
                return w, b.float().to(dev).to(dt)
            return w, b.float().to(dev)
        self.stem = put(*self._fold(network.resbody[0]))
        self.blocks = [(put(*self._fold(blk.convblock1)), put(*self._fold(blk.convblock2))) for blk in list(network.resbody)[1:]]
        self.pconv = put(*self._fold(network.pconv), tower=False)
        self.vconv = put(*self._fold(network.vconv), tower=False)
        self.plinear = (network.plinear.weight.detach().float().to(dev), network.plinear.bias.detach().float().to(dev))
        self.v1 = (network.vlinear[0].weight.detach().float().to(dev), network.vlinear[0].bias.detach().float().to(dev))
        self.v2 = (network.vlinear[2].weight.detach().float().to(dev), network.vlinear[2].bias.detach().float().to(dev))
        self._graph = None

    @torch.no_grad()
    def forward(self, tokens_u8, clocks):
        """tokens uint8 [B,60], clocks float32 [B] (CUDA) -> logits float32 [B,554], values float32 [B]."""
        if self.dtype == torch.float32:            # true fp32 (no TF32) when used as the parity reference
            with torch.backends.cudnn.flags(enabled=True, benchmark=True, allow_tf32=False):
                old = torch.backends.cuda.matmul.allow_tf32
                torch.backends.cuda.matmul.allow_tf32 = False
                try:
                    return self._forward(tokens_u8, clocks)
                finally:
                    torch.backends.cuda.matmul.allow_tf32 = old
        return self._forward(tokens_u8, clocks)

    def _forward(self, tokens_u8, clocks):
        F = torch.nn.functional
        B = tokens_u8.shape[0]
        x = F.embedding(tokens_u8.long().view(B, 2, 6, 5), self.emb).permute(0, 1, 4, 2, 3).reshape(B, 8, 6, 5)
        x = x.to(self.dtype).contiguous(memory_format=torch.channels_last)
        x = F.relu(F.conv2d(x, self.stem[0], self.stem[1], padding=1))
        for (w1, b1), (w2, b2) in self.blocks:
            y = F.relu(F.conv2d(x, w1, b1, padding=1))
            x = F.relu(F.conv2d(y, w2, b2, padding=1) + x)
        x = x.float()
        clk = clocks.view(B, 1).float()
        px = F.relu(F.conv2d(x, self.pconv[0], self.pconv[1])).reshape(B, 60)
        logits = F.linear(torch.cat([px, clk], 1), *self.plinear)
        vx = F.relu(F.conv2d(x, self.vconv[0], self.vconv[1])).reshape(B, 30)
        v = torch.tanh(F.linear(F.relu(F.linear(torch.cat([vx, clk], 1), *self.v1)), *self.v2))
        return logits.contiguous(), v.reshape(B).contiguous()

    def capture(self, tokens_u8, clocks):
        """CUDA-graph the forward over fixed input buffers (the engine's leaf batch)."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self.forward(tokens_u8, clocks)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.forward(tokens_u8, clocks)
        self._graph, self._graph_out = g, out
        return out

    def replay(self):
        self._graph.replay()
        return self._graph_out
