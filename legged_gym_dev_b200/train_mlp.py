"""Tensor-core training step of rsl_rl's ActorCritic: forward, input gradients and weight gradients of
the actor and critic MLPs as launches of ONE grouped tcgen05 GEMM kernel (csrc/gemm.cu) — no autograd, no cuBLAS.

Per minibatch (PPO.update of rsl_rl/algorithms/ppo.py: `loss.backward()`), with L layers per net:
    rows_to_f16            observation gather through the minibatch indices -> fp16 A operand of layer 0
    L   x gemm FWD         actor layer l and critic layer l share a launch; bias + ELU epilogue, fp16 activations H_l kept
    ppo_loss_gathered      loss + d(loss)/d(mu, value) -> fp16 dZ_{L-1} (unscaled), d_std, KL / loss sums
    L-1 x gemm DGRAD       dZ_{l-1} = (dZ_l . W_l) * ELU'(H_{l-1}); W_l is read MN-major from the same packed copy
    1   x gemm WGRAD       every dW_l = dZ_l^T H_{l-1} (+ db_l) of both nets in one launch, split over the batch rows,
                           accumulated in fp32 straight into ActorCritic.flat_grad with the 1/batch factor
fp16 operands, fp32 accumulation (TMEM) and fp32 master weights; `pack()` refreshes the fp16 weight copies after Adam.
The same FWD launches serve the no-grad rollout / inference forward of nets that do not fit the weights-resident
fused kernel (the 512-256-128 rough nets).
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from . import _lib


def _pad16(n):
    return (n + 15) // 16 * 16


class _Net:
    def __init__(self, seq, slices_by_id):
        mods = list(seq)
        self.linears = [m for m in mods if isinstance(m, nn.Linear)]
        acts = [m for m in mods if not isinstance(m, nn.Linear)]
        if not self.linears or len(acts) != len(self.linears) - 1 or not all(isinstance(a, nn.ELU) and a.alpha == 1.0 for a in acts):
            raise ValueError("the tensor-core PPO step supports Linear / ELU(alpha=1) stacks with a linear last layer")
        self.L = len(self.linears)
        self.K = [l.in_features for l in self.linears]
        self.N = [l.out_features for l in self.linears]
        self.Kp = [_pad16(k) for k in self.K]
        self.Np = [_pad16(n) for n in self.N]
        if self.Np[-1] != 16:
            raise ValueError("at most 16 network outputs (the loss kernel's dZ rows are 16 wide)")
        self.w_off = [slices_by_id[id(l.weight)][0] for l in self.linears]
        self.b_off = [slices_by_id[id(l.bias)][0] for l in self.linears]
        self.w16_off = []


class TensorCoreTrainer:
    """Owns the fp16 weight copies and the per-batch activation / gradient buffers of an ActorCritic whose parameters
    live in one flat fp32 buffer (ActorCritic.flatten_parameters)."""

    def __init__(self, ac):
        self.ac = ac
        self.dev = ac.flat_param.device
        _lib.require_current_device(self.dev)
        self.lib = _lib.lib()
        by_id = {id(p): ac._slices[name] for name, p in ac.named_parameters()}
        self.actor, self.critic = _Net(ac.actor, by_id), _Net(ac.critic, by_id)
        self.nets = (self.actor, self.critic)
        off = 0
        for net in self.nets:
            for l in range(net.L):
                net.w16_off.append(off)
                off += net.Np[l] * net.Kp[l]
        self.w16 = torch.zeros(off, dtype=torch.float16, device=self.dev)
        if sum(net.L for net in self.nets) > _lib.PACK_MAX:
            raise ValueError("too many layers for one pack launch")
        self._bufs = {}
        # nets whose 128-row activation tiles fit in shared memory run forward + loss + input gradients as ONE launch (csrc/ppo_chain.cu)
        def chain_fits(net):
            smem = (2 + sum(k // 8 for k in net.Kp)) * (128 * 16 + 16) + 2 * 16 * (64 * 16 + 16) + 256 + 4 * sum(net.Np) + 320
            return 2 <= net.L <= _lib.CHAIN_MAX_LAYERS and max(net.Np) <= 128 and max(net.Kp[1:]) <= 128 and smem <= 227 * 1024
        self._use_chain = all(chain_fits(n) for n in self.nets) and os.environ.get("B200GYM_PPO_CHAIN", "1") != "0"
        # ... and the weight / bias gradients and the observation gather in the same launch (B200GYM_CHAIN_WGRAD=0: A/B against the
        # round-2a form, chain + row-gather launch + grouped weight-gradient GEMM)
        self.fuse_wgrad = self._use_chain and os.environ.get("B200GYM_CHAIN_WGRAD", "1") != "0"
        # the chain kernel takes its weight tiles by TMA bulk copies from a CHUNK-MAJOR fp16 copy ([K/8][N][8]: a 64-column stage is 8
        # contiguous pieces instead of 1024 16-byte ones); the layered GEMM path reads a row-major copy (B200GYM_CHAIN_TMA=0: row-major
        # + cp.async in the chain kernel as well, for A/B)
        self.chain_tma = os.environ.get("B200GYM_CHAIN_TMA", "1") != "0"
        self._build_table()
        self.pack()

    @property
    def use_chain(self):
        return self._use_chain

    @use_chain.setter
    def use_chain(self, on):
        on = bool(on)
        if on != self._use_chain:
            self._use_chain = on
            self._build_table()       # the two paths read different layouts of the fp16 copy
            self.w16.zero_()
            self.pack()

    def _build_table(self):
        self.w16_layout = 1 if (self._use_chain and self.chain_tma) else 0
        tab = _lib.PackTablePOD()
        e, end = 0, 0
        for net in self.nets:
            for l in range(net.L):
                end += net.N[l] * net.K[l]
                t = tab.e[e]
                t.src_off, t.dst_off, t.elem_end = net.w_off[l], net.w16_off[l], end
                t.rows, t.cols, t.layout = net.N[l], net.K[l], self.w16_layout
                t.ld = net.Np[l] if self.w16_layout else net.Kp[l]     # chunk-major: rows per chunk; row-major: leading dimension
                e += 1
        tab.n, tab.total = e, end
        self._tab = tab
        for b in self._bufs.values():
            for c in b.get("chain", ()):
                c.pad = self.w16_layout

    # ------------------------------------------------------------------------------------------------------------
    def pack(self):
        """fp32 master weights -> fp16 operand copies (after every optimiser step / load_state_dict)."""
        _lib.check(self.lib.b200gym_pack_params_f16(_lib.ptr(self.ac.flat_param), self._tab, _lib.ptr(self.w16), _lib.stream_ptr(self.dev)),
                   "pack_params_f16")

    def _w16(self, net, l):
        return self.w16[net.w16_off[l]: net.w16_off[l] + net.Np[l] * net.Kp[l]].view(net.Np[l], net.Kp[l])

    def _buffers(self, B, shared_obs, train):
        key = (B, shared_obs, train)
        b = self._bufs.get(key)
        if b is not None:
            return b
        dev, h16 = self.dev, torch.float16
        b = {"x": [torch.zeros(B, self.actor.Kp[0], dtype=h16, device=dev)]}
        b["x"].append(b["x"][0] if shared_obs else torch.zeros(B, self.critic.Kp[0], dtype=h16, device=dev))
        b["h"] = [[torch.empty(B, net.Np[l], dtype=h16, device=dev) for l in range(net.L - 1)] for net in self.nets]
        b["out"] = [torch.empty(B, 16, device=dev) for _ in self.nets]
        if train:
            b["dz"] = [[torch.empty(B, net.Np[l], dtype=h16, device=dev) for l in range(net.L)] for net in self.nets]
        b["fwd"] = self._fwd_launches(b, B)
        if train:
            b["dgrad"] = self._dgrad_launches(b, B)
            if self.use_chain:
                b["chain"] = [self._chain_net(b, i) for i in range(2)]
        if len(self._bufs) > 8:
            self._bufs.clear()
        self._bufs[key] = b
        return b

    def _problem(self, **kw):
        p = _lib.GemmProblemPOD()
        for k, v in kw.items():
            setattr(p, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
        return p

    def _fwd_launches(self, b, B):
        """(grouped launches: layer l of every net that has one, per-net launches: one problem each)."""
        fp = self.ac.flat_param
        per_net = []
        for i, net in enumerate(self.nets):
            ps = []
            for l in range(net.L):
                a = b["x"][i] if l == 0 else b["h"][i][l - 1]
                last = l == net.L - 1
                out = b["out"][i] if last else b["h"][i][l]
                ps.append(self._problem(a=a, b=self._w16(net, l), out=out, bias=fp.data_ptr() + 4 * net.b_off[l], mode=_lib.GEMM_FWD,
                                        flags=2 if last else 1, m=B, n=net.Np[l], k=net.Kp[l], lda=a.stride(0), ldb=net.Kp[l],
                                        ldo=out.stride(0), m_real=B, n_real=net.N[l], splits=1, scale=1.0))
            per_net.append(ps)
        b["fwd_net"] = [[((_lib.GemmProblemPOD * 1)(p), 1) for p in ps] for ps in per_net]
        launches = []
        for l in range(max(n.L for n in self.nets)):
            ps = [per_net[i][l] for i, net in enumerate(self.nets) if l < net.L]
            launches.append(((_lib.GemmProblemPOD * len(ps))(*ps), len(ps)))
        return launches

    def _chain_net(self, b, i):
        net, c = self.nets[i], _lib.ChainNetPOD()
        c.x, c.w16, c.flat_param, c.out = b["x"][i].data_ptr(), self.w16.data_ptr(), self.ac.flat_param.data_ptr(), b["out"][i].data_ptr()
        c.num_layers, c.ldx = net.L, b["x"][i].stride(0)
        for l in range(net.L):
            c.kp[l], c.np[l], c.n_real[l], c.w_off[l], c.b_off[l] = net.Kp[l], net.Np[l], net.N[l], net.w16_off[l], net.b_off[l]
            c.dz[l] = b["dz"][i][l].data_ptr()
            if l < net.L - 1:
                c.h[l] = b["h"][i][l].data_ptr()
            c.k_real[l], c.w32_off[l] = net.K[l], net.w_off[l]
        c.pad = self.w16_layout      # B200ChainNet.w_layout
        return c

    def _dgrad_launches(self, b, B):
        launches = []
        for l in range(max(n.L for n in self.nets) - 1, 0, -1):
            ps = []
            for i, net in enumerate(self.nets):
                if l >= net.L:
                    continue
                dz, h, out = b["dz"][i][l], b["h"][i][l - 1], b["dz"][i][l - 1]
                ps.append(self._problem(a=dz, b=self._w16(net, l), out=out, aux=h, mode=_lib.GEMM_DGRAD, flags=0, m=B, n=net.Kp[l],
                                        k=net.Np[l], lda=dz.stride(0), ldb=net.Kp[l], ldo=out.stride(0), ldaux=h.stride(0), m_real=B,
                                        n_real=net.K[l], splits=1, scale=1.0))
            launches.append(((_lib.GemmProblemPOD * len(ps))(*ps), len(ps)))
        return launches

    def _wgrad_launches(self, b, B, scale):
        fg = self.ac.flat_grad
        tiles = sum(((net.Np[l] + 127) // 128) * ((net.Kp[l] + 127) // 128) for net in self.nets for l in range(net.L))
        # A/B at 24 576 rows (B200): 8 tiles (flat nets, unfused form): 24 splits 2.35 ms / 37: 2.41 / 16: 2.50 / 48: 2.45; 38 tiles (rough
        # nets): 8 splits = 304 CTAs = one wave of 2 CTAs per SM 11.8 ms / 5: 13.0 / 6: 12.4 / 10: 13.0 / 12: 12.7 / 16: 12.4
        splits = max(1, min((B + 255) // 256, round(192 / tiles) if tiles <= 8 else round(296 / tiles)))
        if os.environ.get("B200GYM_WGRAD_SPLITS"):      # A/B knob (profiles/): batch-row splits of every weight-gradient problem
            splits = max(1, min((B + 127) // 128, int(os.environ["B200GYM_WGRAD_SPLITS"])))
        ps = []
        for i, net in enumerate(self.nets):
            for l in range(net.L):
                h = b["x"][i] if l == 0 else b["h"][i][l - 1]
                dz = b["dz"][i][l]
                ps.append(self._problem(a=dz, b=h, out=fg.data_ptr() + 4 * net.w_off[l], bias=fg.data_ptr() + 4 * net.b_off[l],
                                        mode=_lib.GEMM_WGRAD, flags=0, m=net.Np[l], n=net.Kp[l], k=B, lda=dz.stride(0), ldb=h.stride(0),
                                        ldo=net.K[l], m_real=net.N[l], n_real=net.K[l], splits=splits, scale=scale))
        mx = 8
        return [((_lib.GemmProblemPOD * len(ps[j:j + mx]))(*ps[j:j + mx]), len(ps[j:j + mx])) for j in range(0, len(ps), mx)]

    def _run(self, launches):
        st = _lib.stream_ptr(self.dev)
        for arr, n in launches:
            _lib.check(self.lib.b200gym_gemm_f16(arr, n, st), "gemm_f16")

    def _convert(self, src, idx, dst, B):
        _lib.check(self.lib.b200gym_rows_to_f16(_lib.ptr(src), src.stride(0), src.shape[1], _lib.ptr(idx), _lib.ptr(dst), dst.stride(0), B,
                                                _lib.stream_ptr(self.dev)), "rows_to_f16")

    # ------------------------------------------------------------------------------------------------------------
    def forward_net(self, which, x):
        """No-grad forward of one net (0 actor, 1 critic) for a float32 [B, in] tensor: L FWD launches."""
        B = x.shape[0]
        b = self._buffers(B, False, False)
        self._convert(x, None, b["x"][which], B)
        self._run(b["fwd_net"][which])
        return b["out"][which][:, :self.nets[which].N[-1]].clone()

    def forward_both(self, obs, critic_obs):
        """No-grad forward of actor and critic in shared launches; returns (mu [B, A], value [B, 1])."""
        B = obs.shape[0]
        shared = critic_obs is obs or (critic_obs.data_ptr() == obs.data_ptr() and critic_obs.shape == obs.shape)
        b = self._buffers(B, shared, False)
        self._convert(obs, None, b["x"][0], B)
        if not shared:
            self._convert(critic_obs, None, b["x"][1], B)
        self._run(b["fwd"])
        return b["out"][0][:, :self.actor.N[-1]].clone(), b["out"][1][:, :self.critic.N[-1]].clone()

    def minibatch_forward_backward(self, storage, idx, lp, std, d_std_ptr, scalars, shared_obs):
        """Forward + loss + backward of one minibatch whose rows are `idx` (int64 [B]) of the flattened storage tensors.
        Weight / bias gradients are ACCUMULATED into ActorCritic.flat_grad (zero it first)."""
        B = idx.numel()
        b = self._buffers(B, shared_obs, True)
        ptr, st = _lib.ptr, _lib.stream_ptr(self.dev)
        flat = lambda t: t.flatten(0, 1)
        obs = flat(storage.observations)
        fuse = self.use_chain and self.fuse_wgrad
        if not fuse:
            self._convert(obs, idx, b["x"][0], B)
            if not shared_obs:
                self._convert(flat(storage.privileged_observations), idx, b["x"][1], B)
        if self.use_chain:
            srcs = (obs, obs if shared_obs else flat(storage.privileged_observations))
            for c, src in zip(b["chain"], srcs):
                # fused: rows gathered in the kernel from the fp32 storage, gradients added to the flat buffer, no operand copies in HBM
                c.x32, c.ldx32 = (src.data_ptr(), src.stride(0)) if fuse else (None, 0)
                c.flat_grad = self.ac.flat_grad.data_ptr() if fuse else None
            for i, (c, net) in enumerate(zip(b["chain"], self.nets)):
                for l in range(net.L):      # the fp16 H_l / dZ_l copies in HBM are the operands of the separate weight-gradient GEMM only
                    c.dz[l] = None if fuse else b["dz"][i][l].data_ptr()
                    if l < net.L - 1:
                        c.h[l] = None if fuse else b["h"][i][l].data_ptr()
            _lib.check(self.lib.b200gym_ppo_chain(
                b["chain"][0], b["chain"][1], lp, ptr(idx), ptr(std), ptr(storage.actions), ptr(storage.actions_log_prob),
                ptr(storage.advantages), ptr(storage.returns), ptr(storage.values), ptr(storage.mu), ptr(storage.sigma), d_std_ptr,
                ptr(scalars), st), "ppo_chain")
        else:
            self._run(b["fwd"])
            _lib.check(self.lib.b200gym_ppo_loss_gathered(
                lp, ptr(idx), ptr(b["out"][0]), 16, ptr(b["out"][1]), 16, ptr(std), ptr(storage.actions), ptr(storage.actions_log_prob),
                ptr(storage.advantages), ptr(storage.returns), ptr(storage.values), ptr(storage.mu), ptr(storage.sigma),
                ptr(b["dz"][0][-1]), ptr(b["dz"][1][-1]), d_std_ptr, ptr(scalars), st), "ppo_loss_gathered")
            self._run(b["dgrad"])
        if fuse:
            return
        key = ("wgrad", lp.inv_global_batch)
        if key not in b:
            b[key] = self._wgrad_launches(b, B, lp.inv_global_batch)
        self._run(b[key])
