"""rsl_rl-shaped rollout storage / PPO / runner whose storage arithmetic runs in the fused kernels (csrc/ppo.cu).

Mirrors the surface of rsl_rl v1.0.2 the reference touches: `OnPolicyRunner(env, train_cfg_dict,
log_dir, device=, wandb_callback=)`, `.learn(num_learning_iterations, init_at_random_ep_len)`, `.load/.save`,
`.get_inference_policy(device)`, `.alg.actor_critic`, `.alg.optimizer`, checkpoint keys
{'model_state_dict','optimizer_state_dict','iter','infos'} (deep_tube_learning/utils.py:305-308).

What is fused: the time-out bootstrap + GAE reverse scan + advantage statistics (one launch), normalisation (one
launch), the minibatch gather of all storage tensors (one launch), the whole PPO loss forward+gradient w.r.t. the
network outputs (one launch), grad-norm + clip + Adam over ONE flat parameter buffer (two launches) and the
KL-adaptive learning rate kept on the device — no host sync inside update().  Across GPUs (envs sharded, one process
per GPU) the only collectives are one all-reduce of the 3 advantage statistics per iteration and one all-reduce of the
flat gradient buffer (+ KL sum/count piggy-backed) per minibatch.
Every contraction runs on the tcgen05 tensor cores in hand-written kernels: the no-grad forward passes (rollout `act` /
`evaluate`, `get_inference_policy`) as one fused launch per net when the weights fit in shared memory
(legged_gym_dev_b200/mlp.py, csrc/mlp.cu: fp16 operands, fp32 accumulation in TMEM) or layer by layer through the grouped
GEMM otherwise (the 512-256-128 rough nets), and the training forward / backward / weight gradients of PPO.update through
the grouped GEMM kernel (legged_gym_dev_b200/train_mlp.py, csrc/gemm.cu).  There is no autograd, no cuBLAS and no eager
fallback on these paths: a net the kernels cannot take raises.
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from . import _lib


def _dist_ready():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class RolloutStorage:
    """rsl_rl/storage/rollout_storage.py: [T, N, ...] tensors, same names."""

    class Transition:
        def __init__(self):
            self.clear()

        def clear(self):
            self.observations = self.critic_observations = self.actions = self.rewards = self.dones = None
            self.values = self.actions_log_prob = self.action_mean = self.action_sigma = self.hidden_states = None

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the b200gym RolloutStorage runs on CUDA devices only (no CPU fallback)")
        T, N, dev = num_transitions_per_env, num_envs, self.device
        self.num_envs, self.num_transitions_per_env = N, T
        self.obs_shape, self.privileged_obs_shape, self.actions_shape = obs_shape, privileged_obs_shape, actions_shape
        z = lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype, device=dev)
        self.observations = z(T, N, *obs_shape)
        self.privileged_observations = z(T, N, *privileged_obs_shape) if privileged_obs_shape[0] is not None else None
        self.rewards, self.values, self.returns, self.advantages = z(T, N, 1), z(T, N, 1), z(T, N, 1), z(T, N, 1)
        self.actions, self.mu, self.sigma = z(T, N, *actions_shape), z(T, N, *actions_shape), z(T, N, *actions_shape)
        self.actions_log_prob = z(T, N, 1)
        self.dones = z(T, N, 1, dtype=torch.uint8)
        self.time_outs = z(T, N, dtype=torch.uint8)
        self._stats = torch.zeros(4, dtype=torch.double, device=dev)
        self.step = 0
        self.lib = _lib.lib()

    def add_transitions(self, tr):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s = self.step
        self.observations[s].copy_(tr.observations)
        if self.privileged_observations is not None:
            self.privileged_observations[s].copy_(tr.critic_observations)
        self.actions[s].copy_(tr.actions)
        self.rewards[s].copy_(tr.rewards.view(-1, 1))
        self.dones[s].copy_(tr.dones.view(-1, 1))
        self.values[s].copy_(tr.values)
        self.actions_log_prob[s].copy_(tr.actions_log_prob.view(-1, 1))
        self.mu[s].copy_(tr.action_mean)
        self.sigma[s].copy_(tr.action_sigma)
        if getattr(tr, "time_outs", None) is not None:
            self.time_outs[s].copy_(tr.time_outs)
        else:
            self.time_outs[s].zero_()
        self.step += 1

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam, bootstrap_time_outs=True):
        """PPO.process_env_step's time-out bootstrap + RolloutStorage.compute_returns, fused; statistics all-reduced
        across env shards before the normalisation."""
        T, N, ptr, st = self.num_transitions_per_env, self.num_envs, _lib.ptr, _lib.stream_ptr(self.device)
        self._stats.zero_()
        last_values = last_values.contiguous()
        _lib.check(self.lib.b200gym_gae_returns(ptr(self.rewards), ptr(self.values), ptr(self.dones),
                                                ptr(self.time_outs) if bootstrap_time_outs else None, ptr(last_values),
                                                ptr(self.returns), ptr(self.advantages), ptr(self._stats), T, N, gamma, lam, st),
                   "gae_returns")
        if _dist_ready():
            import torch.distributed as dist
            dist.all_reduce(self._stats)
        _lib.check(self.lib.b200gym_adv_normalize(ptr(self.advantages), ptr(self._stats), T * N, st), "adv_normalize")

    def get_statistics(self):
        done = self.dones.clone()
        done[-1] = 1
        flat = done.permute(1, 0, 2).reshape(-1, 1)
        ids = torch.cat((flat.new_tensor([-1], dtype=torch.int64), flat.nonzero(as_tuple=False)[:, 0]))
        lengths = ids[1:] - ids[:-1]
        return lengths.float().mean(), self.rewards.mean()

    def mini_batch_generator(self, num_mini_batches, num_epochs=8, generator=None, plan=None):
        """Yields the same 11-tuple as rsl_rl; each minibatch is gathered by ONE kernel launch."""
        T, N = self.num_transitions_per_env, self.num_envs
        batch = T * N
        mb = batch // num_mini_batches
        if plan is None:
            perm = torch.randperm(num_mini_batches * mb, device=self.device, generator=generator)
            plan = [perm[i * mb:(i + 1) * mb] for _ in range(num_epochs) for i in range(num_mini_batches)]
        crit = self.privileged_observations if self.privileged_observations is not None else self.observations
        srcs = [self.observations, crit, self.actions, self.values, self.returns, self.actions_log_prob, self.advantages, self.mu,
                self.sigma]
        flat = [t.flatten(0, 1) for t in srcs]
        n = len(flat)
        row_bytes = (C.c_int32 * n)(*[f.shape[1] * f.element_size() if f.dim() > 1 else f.element_size() for f in flat])
        src_ptrs = (C.c_void_p * n)(*[f.data_ptr() for f in flat])
        for idx in plan:
            idx = idx.to(self.device, torch.int64).contiguous()
            outs = [torch.empty(idx.numel(), *f.shape[1:], dtype=f.dtype, device=self.device) for f in flat]
            dst_ptrs = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
            _lib.check(self.lib.b200gym_gather_rows(dst_ptrs, src_ptrs, row_bytes, n, _lib.ptr(idx), idx.numel(),
                                                    _lib.stream_ptr(self.device)), "gather_rows")
            obs, cobs, act, val, ret, logp, adv, mu, sig = outs
            yield obs, cobs, act, val, adv, ret, logp, mu, sig, (None, None), None


class ActorCritic(nn.Module):
    """rsl_rl/modules/actor_critic.py (non-recurrent): MLP actor/critic, ELU by default, learned std.  All parameters are
    views of ONE flat fp32 buffer (and one flat gradient buffer) so that all-reduce / clip / Adam are single passes."""
    is_recurrent = False

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=(256, 256, 256),
                 critic_hidden_dims=(256, 256, 256), activation="elu", init_noise_std=1.0, fused_inference=True, **kwargs):
        super().__init__()
        self._want_fused = activation == "elu"
        self._fused_actor = self._fused_critic = self._trainer = None
        acts = {"elu": nn.ELU, "selu": nn.SELU, "relu": nn.ReLU, "lrelu": nn.LeakyReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid}
        if activation not in acts:
            raise ValueError(f"invalid activation function {activation}")

        def mlp(i, hs, o):
            layers, d = [], i
            for h in hs:
                layers += [nn.Linear(d, h), acts[activation]()]
                d = h
            return nn.Sequential(*layers, nn.Linear(d, o))
        self.actor = mlp(num_actor_obs, list(actor_hidden_dims), num_actions)
        self.critic = mlp(num_critic_obs, list(critic_hidden_dims), 1)
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        self.distribution = None
        self.flat_param = self.flat_grad = None

    def flatten_parameters(self, grad_buffer=None):
        """Re-homes every parameter (and its .grad) into one contiguous buffer; 8 spare floats carry piggy-backed scalars.
        grad_buffer: optional pre-allocated [n + 8] tensor for the gradients (peer-mapped memory in the multi-GPU update)."""
        ps = list(self.parameters())
        dev = ps[0].device
        n = sum(p.numel() for p in ps)
        self.flat_param = torch.empty(n, device=dev)
        self.flat_grad = torch.zeros(n + 8, device=dev) if grad_buffer is None else grad_buffer
        off = 0
        self._slices = {}
        for name, p in self.named_parameters():
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + k].view_as(p)
            p.grad = self.flat_grad[off:off + k].view_as(p)
            self._slices[name] = (off, k)
            off += k
        self.num_flat = n
        self.enable_fused_inference()
        return self

    def enable_fused_inference(self):
        """Builds the tensor-core paths: `TensorCoreTrainer` (grouped GEMM: training step, and the forward of nets too large
        for shared memory) and, for nets whose weights fit in shared memory, the one-launch fused forward (csrc/mlp.cu).
        Raises for nets the kernels cannot take (non-ELU activations, > 16 outputs): there is no eager fallback."""
        self._fused_actor = self._fused_critic = self._trainer = None
        if next(self.parameters()).device.type != "cuda":
            return
        if not self._want_fused:
            raise ValueError("the b200gym ActorCritic runs Linear/ELU stacks on its tensor-core kernels only (activation='elu')")
        from .mlp import FusedMLP
        from .train_mlp import TensorCoreTrainer
        self._trainer = TensorCoreTrainer(self)
        for which, net in (("_fused_actor", self.actor), ("_fused_critic", self.critic)):
            if FusedMLP.fits(net):
                setattr(self, which, FusedMLP(net))

    def repack_fused(self):
        """Refreshes every fp16 / packed weight copy from the fp32 master parameters (after an optimiser step or a load)."""
        if self._trainer is not None:
            self._trainer.pack()
        for f in (self._fused_actor, self._fused_critic):
            if f is not None:
                f.repack()

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        if self.flat_param is not None:
            self.repack_fused()
        return out

    def _forward(self, which, x):
        """Forward of the actor (0) / critic (1) on the tensor-core kernels; the result carries no autograd graph (the
        training gradients are produced by PPO.update's kernels, not by autograd)."""
        if self._trainer is None:
            raise RuntimeError("ActorCritic: call flatten_parameters() on a CUDA device first (the b200gym path has no CPU / eager fallback)")
        if not (x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1):
            raise RuntimeError("ActorCritic: expected a float32 CUDA tensor of shape [batch, num_obs] with unit inner stride")
        fused = self._fused_actor if which == 0 else self._fused_critic
        if fused is not None:
            return fused(x)
        return self._trainer.forward_net(which, x)

    def forward_both(self, obs, critic_obs):
        """(actor(obs), critic(critic_obs)) — ONE launch when both nets run on the fp16 weights-resident kernel (flat nets), else
        two launches of the same kernels."""
        fa, fc = self._fused_actor, self._fused_critic
        if (fa is not None and fc is not None and fa.pairable and fc.pairable and obs.stride(0) % 4 == 0 and critic_obs.stride(0) % 4 == 0
                and os.environ.get("B200GYM_ACT_PAIR", "1") != "0"):
            from .mlp import FusedMLP
            return FusedMLP.forward_pair(fa, obs, fc, critic_obs)
        return self._forward(0, obs), self._forward(1, critic_obs)

    def reset(self, dones=None):
        pass

    @property
    def action_mean(self):
        return self.distribution.mean

    @property
    def action_std(self):
        return self.distribution.stddev

    @property
    def entropy(self):
        return self.distribution.entropy().sum(dim=-1)

    def update_distribution(self, observations):
        mean = self._forward(0, observations)
        self.distribution = torch.distributions.Normal(mean, mean * 0.0 + self.std)

    def act(self, observations, **kwargs):
        self.update_distribution(observations)
        return self.distribution.sample()

    def get_actions_log_prob(self, actions):
        return self.distribution.log_prob(actions).sum(dim=-1)

    def act_inference(self, observations):
        return self._forward(0, observations)

    def evaluate(self, critic_observations, **kwargs):
        return self._forward(1, critic_observations)


class PPO:
    """rsl_rl/algorithms/ppo.py with the storage / loss / optimiser arithmetic in fused kernels."""

    def __init__(self, actor_critic, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95,
                 value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True,
                 schedule="fixed", desired_kl=0.01, device="cuda"):
        self.device = torch.device(device)
        self.desired_kl, self.schedule, self.learning_rate = desired_kl, schedule, learning_rate
        self.actor_critic = actor_critic.to(self.device)
        if self.actor_critic.flat_param is None or self.actor_critic.flat_param.device != self.device:
            self.actor_critic.flatten_parameters()
        self.storage = None
        # multi-GPU gradient exchange (B200GYM_GRAD_EXCHANGE): "fused" (default) = push + flags + rank-ordered sum inside the optimiser
        # kernel (peer_reduce.PeerExchange); "peer" = round-1 barrier / reduce kernel / barrier; "nccl" = all-reduce of the flat buffer.
        # The last two are A/B baselines selected explicitly — nothing falls back on its own: a failed rendezvous raises.
        self._peer = self._xchg = None
        self.exchange = "none"
        if _dist_ready():
            self.exchange = os.environ.get("B200GYM_GRAD_EXCHANGE", "nccl" if os.environ.get("B200GYM_P2P_GRADS") == "0" else "fused")
            if self.exchange == "fused":
                from .peer_reduce import PeerExchange
                self._xchg = PeerExchange(self.actor_critic.num_flat, self.device)
            elif self.exchange == "peer":
                from .peer_reduce import PeerGradReducer
                self._peer = PeerGradReducer(self.actor_critic.num_flat + 8, self.device)
                self.actor_critic.flatten_parameters(grad_buffer=self._peer.buf)
            elif self.exchange != "nccl":
                raise ValueError(f"B200GYM_GRAD_EXCHANGE={self.exchange!r}: expected fused, peer or nccl")
        self.optimizer = FlatAdam(self.actor_critic, lr=learning_rate)
        self.transition = RolloutStorage.Transition()
        self.clip_param, self.num_learning_epochs, self.num_mini_batches = clip_param, num_learning_epochs, num_mini_batches
        self.value_loss_coef, self.entropy_coef, self.gamma, self.lam = value_loss_coef, entropy_coef, gamma, lam
        self.max_grad_norm, self.use_clipped_value_loss = max_grad_norm, use_clipped_value_loss
        self.lib = _lib.lib()
        # [0:4] sums over the update {kl, surrogate, value loss, entropy}, [4] last squared gradient norm, [8:12] sums of the current minibatch
        self._scalars = torch.zeros(16, dtype=torch.double, device=self.device)
        self._sumsq = torch.zeros(1, dtype=torch.double, device=self.device)
        self._klsum = torch.zeros(2, dtype=torch.double, device=self.device)
        self._mb_cache = {}
        # PPO.act's sample: Philox stream keyed by (seed, global env id, act counter); torch.manual_seed controls it like rsl_rl's sampling
        self.seed, self.env_id_offset, self._act_event = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF, 0, 0
        self._act_event_dev = None
        self.use_graph = os.environ.get("B200GYM_PPO_GRAPH", "1") != "0"

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.storage = RolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape, self.device)

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    def act(self, obs, critic_obs):
        """rsl_rl PPO.act: the actor and critic forwards (tcgen05; one launch for flat nets) + ONE launch for Normal(mu, std).sample(), its log-prob and the transition
        written straight into row `storage.step` (csrc/ppo_rollout.cu).  rsl_rl keeps references to obs in the transition because
        its env returns a fresh tensor every step (legged_robot.py:211); the fused env rewrites ONE persistent obs_buf in place, so
        the observations are stored here, at act time.  The transition fields are views of the storage row."""
        tr, ac, st = self.transition, self.actor_critic, self.storage
        if st is None:
            raise RuntimeError("PPO.act: call init_storage() first (the transition is written straight into the rollout storage)")
        if st.step >= st.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s, ptr = st.step, _lib.ptr
        mu, val = ac.forward_both(obs, critic_obs)
        self._act_event += 1
        priv = st.privileged_observations
        _lib.check(self.lib.b200gym_ppo_act_store(
            obs.shape[0], ac.std.numel(), obs.shape[1], critic_obs.shape[1], ptr(mu), mu.stride(0), ptr(val), val.stride(0), ptr(ac.std),
            ptr(obs), obs.stride(0), ptr(critic_obs), critic_obs.stride(0), self.seed, self._act_event,
            ptr(self._act_event_dev) if self._act_event_dev is not None else None, int(self.env_id_offset),
            ptr(st.observations[s]), ptr(priv[s]) if priv is not None else None, ptr(st.actions[s]), ptr(st.values[s]),
            ptr(st.actions_log_prob[s]), ptr(st.mu[s]), ptr(st.sigma[s]), _lib.stream_ptr(self.device)), "ppo_act_store")
        tr.observations = st.observations[s]
        tr.critic_observations = priv[s] if priv is not None else tr.observations
        tr.actions, tr.values, tr.actions_log_prob = st.actions[s], st.values[s], st.actions_log_prob[s, :, 0]
        tr.action_mean, tr.action_sigma = st.mu[s], st.sigma[s]
        return tr.actions

    def use_device_act_counter(self):
        """Keeps the act counter (Philox event of PPO.act's sample) in device memory, advanced by the store kernel, so that a sequence of
        act / env.step / process_env_step calls can be captured in a CUDA graph and replayed (legged_gym_dev_b200.graphs.GraphedRollout)."""
        if self._act_event_dev is None:
            self._act_event_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._act_event_dev.fill_(self._act_event + 1)
        return self._act_event_dev

    def process_env_step(self, rewards, dones, infos):
        """rsl_rl PPO.process_env_step + RolloutStorage.add_transitions for the fields the env step produced: ONE launch.  The
        time-out bootstrap (rewards += gamma * values * time_outs) is deferred into the GAE kernel: the storage keeps the raw reward
        plus the time-out flag, and compute_returns applies it in place before the scan."""
        tr, st = self.transition, self.storage
        if st.step >= st.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s, ptr = st.step, _lib.ptr
        to = infos["time_outs"] if "time_outs" in infos else None
        as_u8 = lambda t: t.view(torch.uint8) if t.dtype == torch.bool else (t if t.dtype == torch.uint8 else (t != 0).view(torch.uint8))
        rewards = rewards if (rewards.dtype == torch.float32 and rewards.is_contiguous()) else rewards.float().contiguous()
        _lib.check(self.lib.b200gym_ppo_store_step(rewards.numel(), ptr(rewards), ptr(as_u8(dones).contiguous()),
                                                   ptr(as_u8(to).contiguous()) if to is not None else None, ptr(st.rewards[s]),
                                                   ptr(st.dones[s]), ptr(st.time_outs[s]),
                                                   ptr(self._act_event_dev) if self._act_event_dev is not None else None,
                                                   _lib.stream_ptr(self.device)), "ppo_store_step")
        st.step += 1
        tr.clear()
        self.actor_critic.reset(dones)

    def compute_returns(self, last_critic_obs):
        last_values = self.actor_critic.evaluate(last_critic_obs).detach()
        self.storage.compute_returns(last_values, self.gamma, self.lam)

    def _minibatch_step(self, idx, world):
        """One minibatch of PPO.update for the row indices in the static tensor `idx`: observation gather + forward (tcgen05
        GEMMs), fused loss + gradient w.r.t. the network outputs, input- and weight-gradient GEMMs, gradient exchange,
        KL-adaptive LR, clip + Adam, fp16 weight refresh.  No host synchronisation: capturable in a CUDA graph."""
        ac, ptr, st = self.actor_critic, _lib.ptr, _lib.stream_ptr(self.device)
        B = idx.numel()
        std_off, _ = ac._slices["std"]
        lp = _lib.PpoLossParamsPOD()
        lp.num_actions, lp.use_clipped_value_loss = ac.std.numel(), int(self.use_clipped_value_loss)
        lp.clip_param, lp.value_loss_coef, lp.entropy_coef = self.clip_param, self.value_loss_coef, self.entropy_coef
        lp.batch, lp.inv_global_batch = B, 1.0 / (B * world)
        sc = self._scalars[8:12]
        adaptive_kl = self.desired_kl if (self.desired_kl is not None and self.schedule == "adaptive") else None
        if world == 1 or self._xchg is not None:
            # the gradient buffer and `sc` were cleared by the previous optimiser launch (or by update()); across ranks the same ONE
            # launch also carries the exchange (push over NVLink, flags, rank-ordered sum): csrc/ppo_train.cu
            ac._trainer.minibatch_forward_backward(self.storage, idx, lp, ac.std, C.c_void_p(ac.flat_grad.data_ptr() + 4 * std_off), sc,
                                                   self.storage.privileged_observations is None)
            self.optimizer.fused_step(self.max_grad_norm, sc, self._scalars, float(B), adaptive_kl, xchg=self._xchg)
            return
        ac.flat_grad.zero_()
        sc.zero_()
        ac._trainer.minibatch_forward_backward(self.storage, idx, lp, ac.std, C.c_void_p(ac.flat_grad.data_ptr() + 4 * std_off), sc,
                                               self.storage.privileged_observations is None)
        self._scalars[0:4] += sc
        # one exchange: gradients + (sum kl, count) piggy-backed in the spare tail of the flat buffer
        tail = ac.flat_grad[ac.num_flat:ac.num_flat + 2]
        tail[0:1].copy_(sc[0:1])
        tail[1:2].fill_(float(B))
        grad = None
        if self._peer is not None:
            # round-1 form: barrier, ONE reduce kernel over peer-mapped memory (rank-ordered sum + squared norm), barrier
            self.optimizer.prepare()
            grad = self._peer.reduce(ac.num_flat, self.optimizer._sumsq)
            tail = grad[ac.num_flat:ac.num_flat + 2]
        else:
            import torch.distributed as dist
            dist.all_reduce(ac.flat_grad)
        if adaptive_kl is not None:
            self._klsum.copy_(tail)
            _lib.check(self.lib.b200gym_adaptive_lr(ptr(self._klsum), float(B * world), self.desired_kl, ptr(self.optimizer.lr), st),
                       "adaptive_lr")
        self.optimizer.step(self.max_grad_norm, grad=grad)
        ac._trainer.pack()

    def _static_plan(self, B, n_mb, n_epochs):
        """Static permutation buffer of one update + the captured CUDA graph of ALL its minibatch steps."""
        key = (B, n_mb, n_epochs, self.storage.observations.data_ptr())
        sp = self._mb_cache.get(key)
        if sp is None:
            sp = dict(perm=torch.zeros(n_mb * B, dtype=torch.int64, device=self.device), graph=None)
            sp["slices"] = [sp["perm"][i * B:(i + 1) * B] for _ in range(n_epochs) for i in range(n_mb)]
            self._mb_cache = {key: sp}
        return sp

    def update(self, plan=None):
        """rsl_rl PPO.update.  The bodies of ALL minibatches of the update (gather -> forward -> loss -> backward -> gradient
        exchange -> Adam -> weight repack, epochs x minibatches times) are captured ONCE as a CUDA graph over a static permutation
        buffer and replayed with one launch per update; only the permutation changes (copied into the static buffer).  A `plan`
        (explicit list of index tensors: tests) or B200GYM_PPO_GRAPH=0 runs the same bodies eagerly."""
        ac = self.actor_critic
        world = 1
        if _dist_ready():
            import torch.distributed as dist
            world = dist.get_world_size()
        T, N = self.storage.num_transitions_per_env, self.storage.num_envs
        B = T * N // self.num_mini_batches
        self._scalars.zero_()
        ac.flat_grad.zero_()
        # The NCCL baseline runs eagerly: keeping the process-group collective out of stream capture avoids depending on capture
        # support in the process group.
        use_graph = self.use_graph and (world == 1 or self._xchg is not None or self._peer is not None)
        sp = None
        if plan is None:
            # rsl_rl: ONE permutation per update, re-used by every epoch (rollout_storage.py mini_batch_generator)
            sp = self._static_plan(B, self.num_mini_batches, self.num_learning_epochs)
            torch.randperm(self.num_mini_batches * B, device=self.device, out=sp["perm"])
            plan = sp["slices"]
        elif use_graph and len(plan) == self.num_mini_batches * self.num_learning_epochs and all(p.numel() == B for p in plan):
            # an explicit plan of rsl_rl's shape (tests): epochs x the same minibatches -> the static buffer takes the first epoch
            M = self.num_mini_batches
            if all(plan[k] is plan[k % M] or torch.equal(plan[k], plan[k % M]) for k in range(M, len(plan))):
                sp = self._static_plan(B, M, self.num_learning_epochs)
                sp["perm"].copy_(torch.cat([p.to(self.device, torch.int64) for p in plan[:M]]))
        use_graph = use_graph and sp is not None
        n_updates = len(plan)
        if use_graph:
            if sp["graph"] is None:
                self._capture(sp, world)
            sp["graph"].replay()
            self.optimizer.steps += n_updates
        else:
            for idx in plan:
                self._minibatch_step(idx.to(self.device, torch.int64).contiguous(), world)
        self.storage.clear()
        ac.repack_fused()
        s = (self._scalars[0:4] / (n_updates * B))
        self.learning_rate = self.optimizer.lr   # device scalar; float(self.learning_rate) syncs on demand
        return s[2], s[1]   # mean_value_loss, mean_surrogate_loss (device scalars)

    def _capture(self, sp, world):
        """Warm-up on a side stream (buffer allocation, kernel attributes), restore the optimiser state, then capture."""
        ac, opt = self.actor_critic, self.optimizer
        state = (ac.flat_param, opt.exp_avg, opt.exp_avg_sq, opt.lr, opt.step_dev, self._scalars, ac.flat_grad)
        keep = [t.clone() for t in state]
        steps = opt.steps
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for idx in sp["slices"][:2]:
                self._minibatch_step(idx, world)
        torch.cuda.current_stream(self.device).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for idx in sp["slices"]:
                self._minibatch_step(idx, world)
        for t, k in zip(state, keep):
            t.copy_(k)
        opt.steps = steps
        ac._trainer.pack()
        sp["graph"] = g


class FlatAdam:
    """torch.optim.Adam semantics (betas 0.9/0.999, eps 1e-8) over ActorCritic.flat_param, fused with clip_grad_norm_."""

    def __init__(self, ac, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.ac, self.betas, self.eps = ac, betas, eps
        dev = ac.flat_param.device
        self.lr = torch.tensor([lr], dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(ac.flat_param)
        self.exp_avg_sq = torch.zeros_like(ac.flat_param)
        self.steps = 0
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.double, device=dev)
        self._ws = torch.zeros(2, dtype=torch.double, device=dev)   # fused_step workspace: squared norm + the two barrier counters
        self._last_sumsq = None
        self.lib = _lib.lib()

    def prepare(self):
        """Advances the device-resident step count and zeroes the squared-norm accumulator."""
        _lib.check(self.lib.b200gym_adam_prepare(_lib.ptr(self.step_dev), _lib.ptr(self._sumsq), _lib.stream_ptr(self.ac.flat_param.device)),
                   "adam_prepare")

    def step(self, max_grad_norm, grad=None):
        """clip_grad_norm_ + Adam; the step count is device-resident (CUDA-graph replayable), `steps` mirrors it on the host.
        grad: already reduced gradients whose squared norm sits in `_sumsq` (peer-memory path: `prepare()` was called before the
        reduction); default = the local flat gradient buffer."""
        ac, ptr, st = self.ac, _lib.ptr, _lib.stream_ptr(self.ac.flat_param.device)
        self.steps += 1
        if grad is None:
            self.prepare()
            _lib.check(self.lib.b200gym_grad_sumsq(ptr(ac.flat_grad), ac.num_flat, 1.0, ptr(self._sumsq), st), "grad_sumsq")
            grad = ac.flat_grad
        _lib.check(self.lib.b200gym_clip_adam_dev(ptr(ac.flat_param), ptr(grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                                  ac.num_flat, 1.0, ptr(self._sumsq), max_grad_norm, ptr(self.lr), self.betas[0],
                                                  self.betas[1], self.eps, ptr(self.step_dev), st), "clip_adam_dev")

    def fused_step(self, max_grad_norm, mb_scalars, totals, count, desired_kl, xchg=None):
        """Minibatch tail in ONE launch (csrc/ppo_train.cu ppo_optimizer_step[_peers]_kernel): [gradient exchange over peer memory,]
        squared gradient norm, KL-adaptive learning rate, clip + Adam, fp16 operand copies of the new weights, gradient buffer and
        minibatch sums cleared."""
        ac, ptr = self.ac, _lib.ptr
        self.steps += 1
        p = _lib.OptParamsPOD()
        p.n, p.count, p.adaptive = ac.num_flat, count, int(desired_kl is not None)
        p.desired_kl, p.max_grad_norm = (desired_kl or 0.0), max_grad_norm
        p.beta1, p.beta2, p.eps = self.betas[0], self.betas[1], self.eps
        tr = ac._trainer
        st = _lib.stream_ptr(ac.flat_param.device)
        if xchg is None:
            _lib.check(self.lib.b200gym_ppo_optimizer_step(p, ptr(ac.flat_param), ptr(ac.flat_grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                                           ptr(self.lr), ptr(self.step_dev), ptr(mb_scalars), ptr(totals), ptr(self._ws),
                                                           tr._tab, ptr(tr.w16), st), "ppo_optimizer_step")
        else:
            _lib.check(self.lib.b200gym_ppo_optimizer_step_peers(p, xchg.peers, xchg.world, xchg.rank, xchg.n_pad, ptr(ac.flat_param),
                                                                 ptr(ac.flat_grad), ptr(xchg.grad_sum), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                                                 ptr(self.lr), ptr(self.step_dev), ptr(mb_scalars), ptr(totals), ptr(xchg.ws),
                                                                 tr._tab, ptr(tr.w16), int(os.environ.get("B200GYM_OPT_CTAS", "0")), st),
                       "ppo_optimizer_step_peers")
        self._last_sumsq = totals[4:5]

    def grad_norm(self):
        return torch.sqrt(self._last_sumsq[0] if self._last_sumsq is not None else self._sumsq[0])

    def state_dict(self):
        return dict(lr=self.lr.clone(), exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone(), steps=self.steps)

    def load_state_dict(self, sd):
        self.lr.copy_(sd["lr"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.steps = sd["steps"]
        self.step_dev.fill_(self.steps)


class OnPolicyRunner:
    """rsl_rl/runners/on_policy_runner.py as the reference calls it (task_registry.py:148, scripts/train.py:44)."""

    def __init__(self, env, train_cfg, log_dir=None, device="cuda", wandb_callback=None):
        self.cfg, self.alg_cfg, self.policy_cfg = train_cfg["runner"], train_cfg["algorithm"], train_cfg["policy"]
        self.device, self.env, self.log_dir, self.wandb_callback = torch.device(device), env, log_dir, wandb_callback
        num_critic_obs = env.num_privileged_obs if env.num_privileged_obs is not None else env.num_obs
        ac = ActorCritic(env.num_obs, num_critic_obs, env.num_actions, **self.policy_cfg).to(self.device)
        self.alg = PPO(ac, device=self.device, **self.alg_cfg)
        self.num_steps_per_env, self.save_interval = self.cfg["num_steps_per_env"], self.cfg["save_interval"]
        self.alg.init_storage(env.num_envs, self.num_steps_per_env, [env.num_obs], [env.num_privileged_obs], [env.num_actions])
        self.alg.env_id_offset = int(getattr(env, "env_id_offset", 0))   # env shards draw distinct action noise
        self.tot_timesteps, self.tot_time, self.current_learning_iteration = 0, 0, 0
        self.log_episode_stats = False
        # opt-in: replay the T-step rollout (act / env.step / process_env_step) from ONE captured CUDA graph — for envs whose step is
        # graph-capturable (a replayed tape, a GPU-resident simulator); the first iteration runs eagerly and is the capture warm-up
        self.graph_rollout = False
        self._rollout_graph = None

    def learn(self, num_learning_iterations, init_at_random_ep_len=False):
        env = self.env
        if init_at_random_ep_len:
            env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length))
        obs = env.get_observations()
        priv = env.get_privileged_observations()
        critic_obs = priv if priv is not None else obs
        self.alg.actor_critic.train()
        infos_out = []
        # reward statistics (rsl_rl's ep_infos, logged when a log_dir / callback is given): the RAW (sum, count) rows of every env step
        # are kept on the device and summed over the env shards by ONE all-reduce per iteration, so every rank logs the means a single
        # process over all envs would (legged_robot.py:175-182; SURVEY.md 8e "reward statistics")
        raw = getattr(env, "_extras_raw", None)
        log_stats = raw is not None and (self.log_dir is not None or self.wandb_callback is not None or self.log_episode_stats)
        hist = torch.zeros(self.num_steps_per_env, raw.numel(), dtype=torch.double, device=self.device) if log_stats else None
        def rollout():
            nonlocal obs, critic_obs
            for t in range(self.num_steps_per_env):
                actions = self.alg.act(obs, critic_obs)
                obs, priv, rewards, dones, infos = env.step(actions)
                critic_obs = priv if priv is not None else obs
                self.alg.process_env_step(rewards, dones, infos)
                if log_stats:
                    hist[t].copy_(raw)

        for it in range(self.current_learning_iteration, self.current_learning_iteration + num_learning_iterations):
            with torch.inference_mode():
                if self.graph_rollout and self._rollout_graph is not None:
                    self._rollout_graph.replay()
                elif self.graph_rollout and it > self.current_learning_iteration:
                    from .graphs import GraphedRollout
                    self._rollout_graph = GraphedRollout(self, rollout)   # captures (nothing runs), then replays once
                    self._rollout_graph.replay()
                else:
                    rollout()
                self.alg.compute_returns(critic_obs)
            mean_value_loss, mean_surrogate_loss = self.alg.update()
            infos_out.append(dict(it=it, mean_value_loss=mean_value_loss, mean_surrogate_loss=mean_surrogate_loss))
            if log_stats:
                infos_out[-1]["episode"] = self._episode_stats(hist)
            if self.wandb_callback is not None:
                self.wandb_callback(dict(mean_value_loss=float(mean_value_loss), mean_surrogate_loss=float(mean_surrogate_loss), it=it),
                                    float(self.alg.learning_rate), self.alg.actor_critic.std.mean().item(),
                                    self.alg.actor_critic.state_dict(), self.alg.optimizer.state_dict(), self.device, None)
            if self.log_dir is not None and it % self.save_interval == 0:
                self.save(os.path.join(self.log_dir, f"model_{it}.pt"))
        self.current_learning_iteration += num_learning_iterations
        return infos_out

    def _episode_stats(self, hist):
        """Mean over the env steps of the rollout that saw a reset of the per-step means over the reset envs (rsl_rl's logger:
        torch.mean over the collected ep_infos), with the (sum, count) rows all-reduced over the env shards first."""
        total = self.env.num_envs
        if _dist_ready():
            import torch.distributed as dist
            hist = hist.clone()
            dist.all_reduce(hist)
            total *= dist.get_world_size()
        stats = self.env.episode_stats_from_raw(hist, total_envs=total)
        seen = hist[:, -1] > 0
        return {k: (v[seen].mean() if k != "terrain_level" else v.mean()) for k, v in stats.items()}

    def save(self, path, infos=None):
        torch.save({"model_state_dict": self.alg.actor_critic.state_dict(), "optimizer_state_dict": self.alg.optimizer.state_dict(),
                    "iter": self.current_learning_iteration, "infos": infos}, path)

    def load(self, path, load_optimizer=True):
        d = torch.load(path, map_location=self.device)
        self.alg.actor_critic.load_state_dict(d["model_state_dict"])
        self.alg.actor_critic.repack_fused()
        if load_optimizer:
            self.alg.optimizer.load_state_dict(d["optimizer_state_dict"])
        self.current_learning_iteration = d["iter"]
        return d["infos"]

    def get_inference_policy(self, device=None):
        self.alg.actor_critic.eval()
        if device is not None:
            self.alg.actor_critic.to(device)
        return self.alg.actor_critic.act_inference
