"""Fused ActorCritic MLP forward on the tcgen05 tensor cores (csrc/mlp.cu): packing of nn.Sequential(Linear, ELU, ...)
weights into the kernel's operand layout and the call wrapper.  Used for the no-grad forward passes of the rollout
(`PPO.act`) and of `get_inference_policy` when the whole weight set fits in shared memory (`FusedMLP.fits`); larger nets
and the training step run layer by layer through the grouped GEMM (legged_gym_dev_b200/train_mlp.py)."""
import torch
import torch.nn as nn

from . import _lib


def _pad(n, m):
    return (n + m - 1) // m * m


class FusedMLP:
    @staticmethod
    def fits(seq: nn.Sequential) -> bool:
        """True when the weights-resident one-launch kernel can take this Linear/ELU stack."""
        lin = [m for m in seq if isinstance(m, nn.Linear)]
        acts = [m for m in seq if not isinstance(m, nn.Linear)]
        if not lin or len(lin) > 6 or len(acts) != len(lin) - 1 or not all(isinstance(a, nn.ELU) for a in acts):
            return False
        dims = [_pad(lin[0].in_features, 16)] + [_pad(l.out_features, 16) for l in lin]
        wtot = sum(dims[i] * dims[i + 1] for i in range(len(lin)))
        smem = (128 * max(dims[:-1]) + wtot + sum(dims[1:])) * 4 + 64
        return max(dims[1:]) <= 256 and smem <= 227 * 1024

    def __init__(self, seq: nn.Sequential):
        self.linears = [m for m in seq if isinstance(m, nn.Linear)]
        acts = [m for m in seq if not isinstance(m, nn.Linear)]
        if not self.linears or not all(isinstance(a, nn.ELU) for a in acts) or len(acts) != len(self.linears) - 1:
            raise ValueError("FusedMLP supports Linear/ELU stacks with a linear last layer")
        if len(self.linears) > 6:
            raise ValueError("at most 6 layers")
        self.in_dim, self.out_dim = self.linears[0].in_features, self.linears[-1].out_features
        self.dims = [_pad(self.in_dim, 16)] + [_pad(l.out_features, 16) for l in self.linears]
        dev = self.linears[0].weight.device
        _lib.require_current_device(dev)
        wtot = sum(self.dims[i] * self.dims[i + 1] for i in range(len(self.linears)))
        smem = (128 * max(self.dims[:-1]) + wtot + sum(self.dims[1:])) * 4 + 64
        if max(self.dims[1:]) > 256 or smem > 227 * 1024:
            raise ValueError("net too large for the weights-resident fused kernel")
        self.wtot = wtot
        self.wpacked = torch.zeros(wtot + (wtot + 1) // 2, device=dev)   # fp32 section, then the fp16 section (2 halves per float)
        self.bias = torch.zeros(_pad(sum(self.dims[1:]), 4), device=dev)
        self.lib = _lib.lib()
        self.repack()

    @torch.no_grad()
    def repack(self):
        """W_l [N,K] -> fp32 [K/4][N][4] and fp16 [K/8][N][8] (zero padded), biases back to back.  Call after every optimiser step."""
        woff = boff = 0
        w16 = self.wpacked[self.wtot:].view(torch.float16)
        for i, lin in enumerate(self.linears):
            K, N = self.dims[i], self.dims[i + 1]
            W = torch.zeros(N, K, device=self.wpacked.device)
            W[:lin.out_features, :lin.in_features] = lin.weight
            self.wpacked[woff:woff + K * N].copy_(W.view(N, K // 4, 4).permute(1, 0, 2).reshape(-1))
            w16[woff:woff + K * N].copy_(W.view(N, K // 8, 8).permute(1, 0, 2).reshape(-1))
            self.bias[boff:boff + N].zero_()
            self.bias[boff:boff + lin.out_features].copy_(lin.bias)
            woff += K * N
            boff += N

    def _params(self, x):
        _lib.require_cuda(x, "x")
        if x.dtype != torch.float32 or x.dim() != 2 or x.shape[1] != self.in_dim or x.stride(1) != 1:
            raise ValueError(f"expected a float32 [batch, {self.in_dim}] tensor with unit inner stride")
        p = _lib.MlpParamsPOD()
        p.batch, p.num_layers, p.in_dim, p.in_stride, p.out_dim = x.shape[0], len(self.linears), self.in_dim, x.stride(0), self.out_dim
        for i, d in enumerate(self.dims):
            p.dims[i] = d
        return p

    def __call__(self, x):
        p = self._params(x)
        out = torch.empty(x.shape[0], self.out_dim, device=x.device)
        _lib.check(self.lib.b200gym_mlp_forward(p, x.data_ptr(), self.wpacked.data_ptr(), self.bias.data_ptr(), out.data_ptr(),
                                                torch.cuda.current_stream(x.device).cuda_stream), "mlp_forward")
        return out

    @property
    def pairable(self) -> bool:
        """True when this net runs on the fp16 four-slot kernel, the one b200gym_mlp_forward_pair launches for two nets at once."""
        return (max(self.dims[1:]) <= 128 and all(d % 16 == 0 for d in self.dims[:-1]) and self.in_dim % 8 == 0 and self.wtot % 8 == 0)

    @staticmethod
    def forward_pair(a: "FusedMLP", xa, b: "FusedMLP", xb):
        """a(xa), b(xb) in one launch (csrc/mlp.cu, mlp_forward_h4_pair_kernel)."""
        pa, pb = a._params(xa), b._params(xb)
        oa = torch.empty(xa.shape[0], a.out_dim, device=xa.device)
        ob = torch.empty(xb.shape[0], b.out_dim, device=xb.device)
        _lib.check(a.lib.b200gym_mlp_forward_pair(pa, xa.data_ptr(), a.wpacked.data_ptr(), a.bias.data_ptr(), oa.data_ptr(),
                                                  pb, xb.data_ptr(), b.wpacked.data_ptr(), b.bias.data_ptr(), ob.data_ptr(),
                                                  torch.cuda.current_stream(xa.device).cuda_stream), "mlp_forward_pair")
        return oa, ob
