"""Group G parity on the GPU: storage / PPO kernels vs the restated rsl_rl arithmetic (oracle/port_ppo.py; parity
unpinned by the reference, see that file's header).  returns / advantages: 1e-5 (S = 1); update: losses / KL / gradient norm
to 1e-3 as per SURVEY.md §8d cfg 5 (the contractions run in fp16 on the tcgen05 tensor cores with fp32 accumulation; the
oracle is fp32 torch), identical LR-schedule decisions."""
import copy
import os

import pytest
import torch

from oracle import port_ppo as O
from oracle.compare import assert_close, assert_exact

pytestmark = pytest.mark.gpu


def make_storage_inputs(T, N, seed=1):
    g = torch.Generator().manual_seed(seed)
    rewards = 0.02 + 0.05 * torch.randn(T, N, 1, generator=g)
    values = 0.5 + 0.3 * torch.randn(T, N, 1, generator=g)
    dones = torch.rand(T, N, 1, generator=g) < 0.005
    time_outs = dones.squeeze(-1) & (torch.rand(T, N, generator=g) < 0.5)
    last_values = 0.5 + 0.3 * torch.randn(N, 1, generator=g)
    return rewards, values, dones, time_outs, last_values


@pytest.mark.parametrize("T,N", [(24, 4096), (24, 1000), (1, 7), (50, 333)])
def test_gae_returns_and_normalisation(T, N):
    from legged_gym_dev_b200.ppo import RolloutStorage
    rewards, values, dones, time_outs, last_values = make_storage_inputs(T, N)
    st = RolloutStorage(N, T, [48], [None], [12])
    st.rewards.copy_(rewards), st.values.copy_(values), st.dones.copy_(dones), st.time_outs.copy_(time_outs)
    st.compute_returns(last_values.cuda(), 0.99, 0.95)
    boot = torch.stack([O.process_env_step_bootstrap(rewards[t, :, 0], values[t], time_outs[t], 0.99) for t in range(T)]).unsqueeze(-1)
    ret, adv_raw, adv = O.compute_returns(boot, values, dones, last_values, 0.99, 0.95)
    assert_close(st.rewards.cpu(), boot, 1.0, "bootstrapped rewards")
    assert_close(st.returns.cpu(), ret, 1.0, "returns")
    if T * N > 1:
        assert_close(st.advantages.cpu(), adv, 1.0, "normalised advantages", rtol=2e-5)


def test_gather_rows_matches_indexing():
    from legged_gym_dev_b200.ppo import RolloutStorage
    T, N = 24, 512
    st = RolloutStorage(N, T, [235], [None], [12])
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in (st.observations, st.actions, st.values, st.returns, st.actions_log_prob, st.advantages, st.mu, st.sigma):
        t.copy_(torch.randn(t.shape, device="cuda", generator=g))
    plan = [torch.randperm(T * N, device="cuda", generator=g)[:3000], torch.arange(5, device="cuda")]
    flat = lambda t: t.flatten(0, 1)
    for idx, mb in zip(plan, st.mini_batch_generator(4, 1, plan=plan)):
        obs, cobs, act, val, adv, ret, logp, mu, sig, _, _ = mb
        assert torch.equal(obs, flat(st.observations)[idx]) and torch.equal(cobs, obs)
        assert torch.equal(act, flat(st.actions)[idx]) and torch.equal(val, flat(st.values)[idx])
        assert torch.equal(adv, flat(st.advantages)[idx]) and torch.equal(ret, flat(st.returns)[idx])
        assert torch.equal(logp, flat(st.actions_log_prob)[idx]) and torch.equal(mu, flat(st.mu)[idx])
        assert torch.equal(sig, flat(st.sigma)[idx])


def _oracle_and_fused(num_obs=48, hidden=(128, 64, 32), seed=1, num_actions=12):
    from legged_gym_dev_b200.ppo import ActorCritic, PPO
    torch.manual_seed(seed)
    ref = O.ActorCritic(num_obs, num_obs, num_actions, list(hidden), list(hidden), init_noise_std=1.0)
    ac = ActorCritic(num_obs, num_obs, num_actions, actor_hidden_dims=hidden, critic_hidden_dims=hidden, init_noise_std=1.0)
    ac.load_state_dict(copy.deepcopy(ref.state_dict()))
    alg = PPO(ac, num_learning_epochs=2, num_mini_batches=4, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0,
              entropy_coef=0.01, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive",
              desired_kl=0.01, device="cuda")
    return ref, alg


def _fill(alg, ref, T, N, num_obs, seed=2):
    """A rollout generated with the (shared) initial policy so that old_log_prob / mu / sigma are consistent."""
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(T, N, num_obs, generator=g)
    with torch.no_grad():
        d = ref.dist(obs)
        A = ref.std.numel()
        actions = d.mean + d.stddev * torch.randn(T, N, A, generator=g)
        logp = d.log_prob(actions).sum(-1, keepdim=True)
        values = ref.critic(obs)
    rewards, _, dones, time_outs, last_values = make_storage_inputs(T, N, seed)
    if alg.storage is None or (alg.storage.num_envs, alg.storage.num_transitions_per_env) != (N, T):
        alg.init_storage(N, T, [num_obs], [None], [ref.std.numel()])
    st = alg.storage
    for name, t in dict(observations=obs, actions=actions, actions_log_prob=logp, values=values, mu=d.mean, sigma=d.stddev,
                        rewards=rewards, dones=dones, time_outs=time_outs).items():
        getattr(st, name).copy_(t)
    st.compute_returns(last_values.cuda(), 0.99, 0.95)
    boot = torch.stack([O.process_env_step_bootstrap(rewards[t, :, 0], values[t], time_outs[t], 0.99) for t in range(T)]).unsqueeze(-1)
    ret, _, adv = O.compute_returns(boot, values, dones, last_values, 0.99, 0.95)
    f = lambda t: t.flatten(0, 1)
    store = dict(obs=f(obs), critic_obs=f(obs), actions=f(actions), values=f(values), returns=f(ret), old_log_prob=f(logp),
                 advantages=f(adv), old_mu=f(d.mean), old_sigma=f(d.stddev))
    return store


def test_ppo_loss_gradients_match_autograd():
    from legged_gym_dev_b200 import _lib
    ref, alg = _oracle_and_fused()
    store = _fill(alg, ref, 8, 256, 48)
    idx = torch.arange(0, 2048, 2)
    batch = {k: v[idx] for k, v in store.items()}
    # perturb the policy so ratios leave the clip range and the value clip engages
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(0.05 * torch.randn_like(p))
    mu = ref.actor(batch["obs"]).detach().requires_grad_(True)
    value = ref.critic(batch["critic_obs"]).detach().requires_grad_(True)
    std = ref.std.detach().clone().requires_grad_(True)
    d = torch.distributions.Normal(mu, mu * 0.0 + std)
    logp = d.log_prob(batch["actions"]).sum(-1)
    ent = d.entropy().sum(-1)
    ratio = torch.exp(logp - batch["old_log_prob"].squeeze())
    adv = batch["advantages"].squeeze()
    surr = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 0.8, 1.2)).mean()
    vc = batch["values"] + (value - batch["values"]).clamp(-0.2, 0.2)
    vl = torch.max((value - batch["returns"]).pow(2), (vc - batch["returns"]).pow(2)).mean()
    loss = surr + 1.0 * vl - 0.01 * ent.mean()
    loss.backward()
    B = len(idx)
    lp = _lib.PpoLossParamsPOD()
    lp.batch, lp.num_actions, lp.use_clipped_value_loss = B, 12, 1
    lp.clip_param, lp.value_loss_coef, lp.entropy_coef, lp.inv_global_batch = 0.2, 1.0, 0.01, 1.0 / B
    c = lambda t: t.detach().cuda().contiguous()
    d_mu, d_v, d_std = torch.empty(B, 12, device="cuda"), torch.empty(B, device="cuda"), torch.zeros(12, device="cuda")
    sc = torch.zeros(4, dtype=torch.double, device="cuda")
    args = [c(mu), c(std), c(value), c(batch["actions"]), c(batch["old_log_prob"]), c(batch["advantages"]), c(batch["returns"]),
            c(batch["values"]), c(batch["old_mu"]), c(batch["old_sigma"])]
    _lib.check(_lib.lib().b200gym_ppo_loss(lp, *[_lib.ptr(a) for a in args], _lib.ptr(d_mu), _lib.ptr(d_v), _lib.ptr(d_std), _lib.ptr(sc),
                                           _lib.stream_ptr()), "ppo_loss")
    assert_close(d_mu.cpu() * B, mu.grad * B, 1.0, "d loss / d mu", rtol=1e-4)
    assert_close(d_v.cpu() * B, value.grad.squeeze() * B, 1.0, "d loss / d value", rtol=1e-4)
    assert_close(d_std.cpu(), std.grad, 1.0, "d loss / d std", rtol=1e-4)
    assert_close(sc[1].cpu() / B, surr.detach().double(), 1.0, "surrogate", rtol=1e-4)
    assert_close(sc[2].cpu() / B, vl.detach().double(), 1.0, "value loss", rtol=1e-4)
    assert_close(sc[3].cpu() / B, ent.mean().detach().double(), 1.0, "entropy", rtol=1e-4)


def _update_pair(num_obs, hidden, T, N, epochs=2, graph=True, seed=1):
    ref, alg = _oracle_and_fused(num_obs, hidden, seed=seed)
    alg.use_graph, alg.num_learning_epochs = graph, epochs
    store = _fill(alg, ref, T, N, num_obs)
    plan = O.mini_batch_indices(T, N, 4, epochs, generator=torch.Generator().manual_seed(3))
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    lr, stats = O.ppo_update(ref, opt, store, plan, 1e-3, desired_kl=0.01, max_grad_norm=1.0, schedule="adaptive", clip_param=0.2,
                             value_loss_coef=1.0, entropy_coef=0.01, use_clipped_value_loss=True)
    v_loss, s_loss = alg.update(plan=[p.cuda() for p in plan])
    return ref, alg, plan, lr, stats, float(v_loss), float(s_loss)


def _check_update(ref, alg, plan, lr, stats, v_loss, s_loss):
    """SURVEY.md §8d cfg 5 contract for the tensor-core path: losses to 1e-3 relative and identical LR-schedule decisions.  Adam's
    sign-like first steps make a parameter-level 1e-3 bound meaningless at 10-bit operand mantissas (a flipped 1e-7 gradient moves
    a weight by 2 lr): parameters are bounded by 10 lr in the worst element and 0.5 lr on average."""
    assert alg.optimizer.steps == len(plan) and int(alg.optimizer.step_dev) == len(plan)
    assert abs(float(alg.optimizer.lr) - lr) <= 1e-9 + 1e-6 * lr, (float(alg.optimizer.lr), lr, [s["lr"] for s in stats])
    want_v = sum(s["value_loss"] for s in stats) / len(stats)
    want_s = sum(s["surrogate"] for s in stats) / len(stats)
    assert abs(v_loss - want_v) <= 1e-3 * abs(want_v) + 1e-6, (v_loss, want_v)
    assert abs(s_loss - want_s) <= 1e-3 * max(abs(want_s), 0.05), (s_loss, want_s)
    sd = alg.actor_critic.state_dict()
    lr_max = max(s["lr"] for s in stats)
    worst = max((sd[k].cpu() - v).abs().max().item() for k, v in ref.state_dict().items())
    mean = sum((sd[k].cpu() - v).abs().sum().item() for k, v in ref.state_dict().items()) / sum(v.numel() for v in ref.state_dict().values())
    assert worst <= 10 * lr_max and mean <= 0.5 * lr_max, (worst, mean, lr_max)


@pytest.mark.parametrize("graph", [True, False])
def test_ppo_update_tracks_oracle(graph):
    """cfg 5 shape: flat nets [128,64,32], 4 minibatches.  graph=True: the minibatch step replayed from one captured CUDA graph
    (the default), graph=False: the same body launched eagerly."""
    _check_update(*_update_pair(48, (128, 64, 32), 24, 512, graph=graph))


def test_ppo_update_cfg5_full_size():
    """BASELINE.json configs[4] at its stated size: 4096 envs x 24 steps, 4 minibatches of 24 576 samples."""
    _check_update(*_update_pair(48, (128, 64, 32), 24, 4096, epochs=2))


def test_ppo_update_rough_nets():
    """The base LeggedRobotCfgPPO policy (legged_robot_config.py:242-243): 235 -> 512 -> 256 -> 128 -> 12 / 1, weights streamed
    through shared memory by the grouped GEMM (they do not fit the weights-resident forward kernel)."""
    _check_update(*_update_pair(235, (512, 256, 128), 24, 256, epochs=1))


@pytest.mark.parametrize("num_obs,hidden,B,chain,A", [(48, (128, 64, 32), 6144, True, 12), (48, (128, 64, 32), 6144, False, 12),
                                                       (235, (512, 256, 128), 3000, False, 12), (48, (128, 64, 32), 100, True, 12),
                                                       (235, (128, 64), 1000, True, 12), (40, (64, 128, 32, 32), 777, True, 12),
                                                       (38, (128, 64, 32), 2000, True, 4), (48, (128, 64, 32), 1500, True, 16),
                                                       (38, (128, 64, 32), 2000, False, 4)])
def test_minibatch_gradients_match_autograd(num_obs, hidden, B, chain, A):
    """flat_grad after one forward/backward of the tcgen05 path vs torch autograd (fp32) of the restated loss on the same
    minibatch: gradient norm to 1e-3 relative, gradient vector to 5e-3 of its norm (fp16 operands, fp32 accumulation)."""
    from legged_gym_dev_b200 import _lib
    import ctypes as C
    ref, alg = _oracle_and_fused(num_obs, hidden, num_actions=A)   # A = 4: the Hopper's action width; 16: the second reduction pass
    T, N = 8, 1024
    store = _fill(alg, ref, T, N, num_obs)
    with torch.no_grad():                       # move the policy away from the rollout policy: ratios != 1, clips engage
        for p in ref.parameters():
            p.add_(0.03 * torch.randn_like(p))
    ac = alg.actor_critic
    ac.load_state_dict(copy.deepcopy(ref.state_dict()))
    assert ac._trainer.use_chain or not chain      # flat-sized nets take the one-launch chain kernel by default
    ac._trainer.use_chain = chain                  # chain=False: the layered GEMM path (what the rough nets use)
    idx = torch.randperm(T * N, generator=torch.Generator().manual_seed(4))[:B]
    batch = {k: v[idx] for k, v in store.items()}
    loss, info = O.ppo_loss(ref, batch, clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.01, use_clipped_value_loss=True)
    loss.backward()
    lp = _lib.PpoLossParamsPOD()
    lp.batch, lp.num_actions, lp.use_clipped_value_loss = B, A, 1
    lp.clip_param, lp.value_loss_coef, lp.entropy_coef, lp.inv_global_batch = 0.2, 1.0, 0.01, 1.0 / B
    ac.flat_grad.zero_()
    sc = torch.zeros(4, dtype=torch.double, device="cuda")
    std_off, _ = ac._slices["std"]
    ac._trainer.minibatch_forward_backward(alg.storage, idx.cuda(), lp, ac.std, C.c_void_p(ac.flat_grad.data_ptr() + 4 * std_off), sc, True)
    torch.cuda.synchronize()
    got = {name: ac.flat_grad[o:o + k].cpu().view_as(p) for (name, p), (o, k) in
           ((it, ac._slices[it[0]]) for it in ac.named_parameters())}
    want = {name: p.grad for name, p in ref.named_parameters()}
    gn_got = sum(float((g.double() ** 2).sum()) for g in got.values()) ** 0.5
    gn_want = sum(float((g.double() ** 2).sum()) for g in want.values()) ** 0.5
    assert abs(gn_got - gn_want) <= 1e-3 * gn_want, (gn_got, gn_want)
    diff = sum(float(((got[k] - want[k]).double() ** 2).sum()) for k in want) ** 0.5
    per = {k: float((got[k] - want[k]).abs().max() / (want[k].abs().max() + 1e-12)) for k in want}
    assert diff <= 5e-3 * gn_want, (diff, gn_want, per)
    assert abs(float(sc[0]) / B - float(info["kl_mean"])) <= 1e-3 * float(info["kl_mean"]) + 1e-6
    assert abs(float(sc[1]) / B - float(info["surrogate_loss"])) <= 1e-3 * max(abs(float(info["surrogate_loss"])), 0.05)
    assert abs(float(sc[2]) / B - float(info["value_loss"])) <= 1e-3 * float(info["value_loss"]) + 1e-6


def test_graph_and_eager_updates_agree_over_two_iterations():
    """Second update() reuses the captured graph (new indices, advanced Adam step): must match the eager path closely (the
    weight-gradient split-K sums are fp32 red.add in arrival order, so the two runs are not bit-identical)."""
    outs = []
    for graph in (True, False):
        ref, alg = _oracle_and_fused(seed=5)
        alg.use_graph = graph
        T, N = 8, 256
        for it in range(2):
            _fill(alg, ref, T, N, 48, seed=7 + it)
            plan = O.mini_batch_indices(T, N, 4, 2, generator=torch.Generator().manual_seed(11 + it))
            alg.storage.step = T
            alg.update(plan=[p.cuda() for p in plan])
        outs.append((alg.actor_critic.flat_param.clone(), float(alg.optimizer.lr), alg.optimizer.steps))
    assert outs[0][2] == outs[1][2] == 16
    assert outs[0][1] == outs[1][1]
    d = (outs[0][0] - outs[1][0]).abs()
    assert d.max().item() <= 5e-3 and d.mean().item() <= 2e-5, (d.max().item(), d.mean().item())


def test_update_launches_no_library_gemm():
    """The update path contains no cuBLAS / cutlass / autograd kernel: every launch of one update() is one of ours."""
    from torch.profiler import profile, ProfilerActivity
    ref, alg = _oracle_and_fused()
    alg.use_graph = False
    _fill(alg, ref, 8, 256, 48)
    alg.update()
    alg.storage.step = 8
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        alg.update()
        torch.cuda.synchronize()
    names = {e.key for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA}
    bad = [n for n in names if "gemm_f16_kernel" not in n and
           any(t in n.lower() for t in ("sgemm", "cutlass", "cublas", "gemv", "ampere", "sm90", "xmma", "nvjet"))]
    assert not bad, bad
    if os.environ.get("B200GYM_CHAIN_WGRAD") == "0" or os.environ.get("B200GYM_PPO_CHAIN") == "0":
        return      # A/B forms: chain + grouped weight-gradient GEMM, or the layered GEMM path
    # flat nets: forward + loss + backward (weight gradients included) is the chain kernel, then the optimiser kernel
    assert any("ppo_chain_kernel" in n for n in names), names
    ours = [n for n in names if "ppo_chain_kernel" in n or "ppo_optimizer_step" in n]
    mb = [e for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA and e.key in ours]
    assert sorted(e.count for e in mb) == [alg.num_learning_epochs * alg.num_mini_batches] * 2, [(e.key, e.count) for e in mb]
    assert not any("gemm_f16_kernel" in n or "rows_to_f16" in n for n in names), names


@pytest.mark.parametrize("world", [1, 2, 5, 8])
def test_grad_reduce_peers_kernel(world):
    """b200gym_grad_reduce_peers on ONE GPU with local tensors standing in for the peer-mapped buffers: rank-ordered fp32 sum of
    all buffers + squared norm of the parameter part (the multi-rank wiring is checked by tools/ddp_train_check.py on 2+ GPUs)."""
    from legged_gym_dev_b200 import _lib
    L = _lib.lib()
    n, n_params = 33657 + 8, 33657
    g = torch.Generator(device="cuda").manual_seed(world)
    bufs = [torch.randn(n, generator=g, device="cuda") for _ in range(world)]
    peers = _lib.PeerPtrsPOD()
    for r, b in enumerate(bufs):
        peers.ptr[r] = b.data_ptr()
    out = torch.empty(n, device="cuda")
    sumsq = torch.zeros(1, dtype=torch.double, device="cuda")
    _lib.check(L.b200gym_grad_reduce_peers(peers, world, _lib.ptr(out), n, n_params, _lib.ptr(sumsq), _lib.stream_ptr("cuda")))
    want = torch.zeros(n, device="cuda")
    for b in bufs:                                   # same order, same fp32 rounding
        want = want + b
    assert torch.equal(out, want)
    ref = float((want[:n_params].double() ** 2).sum())
    assert abs(float(sumsq) - ref) <= 1e-12 * ref
    with pytest.raises(RuntimeError):
        _lib.check(L.b200gym_grad_reduce_peers(peers, 17, _lib.ptr(out), n, n_params, _lib.ptr(sumsq), _lib.stream_ptr("cuda")))


def test_act_store_matches_oracle():
    """PPO.act (two tcgen05 forwards + ONE act/store launch) against the sampling specification oracle/port_ppo.py
    sample_actions: actions and summed log-prob to 1e-5 (S = max(1, |value|)), the transition row of the storage exact."""
    from legged_gym_dev_b200.ppo import ActorCritic, PPO
    torch.manual_seed(3)
    N, T, A = 1000, 4, 12
    ac = ActorCritic(48, 235, A, actor_hidden_dims=(128, 64, 32), critic_hidden_dims=(128, 64, 32), init_noise_std=0.7)
    alg = PPO(ac, device="cuda")
    alg.init_storage(N, T, [48], [235], [A])
    alg.env_id_offset = 5000
    g = torch.Generator(device="cuda").manual_seed(0)
    zs = []
    for s in range(3):
        obs = torch.randn(N, 48, device="cuda", generator=g)
        cobs = torch.randn(N, 235, device="cuda", generator=g)
        act = alg.act(obs, cobs)
        st = alg.storage
        assert act.data_ptr() == st.actions[s].data_ptr()
        assert torch.equal(st.observations[s], obs) and torch.equal(st.privileged_observations[s], cobs)
        mu, val = ac.act_inference(obs), ac.evaluate(cobs)
        assert torch.equal(st.mu[s], mu) and torch.equal(st.values[s], val)
        assert torch.equal(st.sigma[s], ac.std.detach().expand(N, A))
        want_a, want_lp = O.sample_actions(mu.cpu(), ac.std.detach().cpu(), alg.seed, 5000 + torch.arange(N).numpy(), s + 1)
        assert_close(st.actions[s].cpu(), want_a, 1.0, f"actions step {s}", rtol=1e-5)
        assert_close(st.actions_log_prob[s, :, 0].cpu(), want_lp, 1.0, f"log-prob step {s}", rtol=1e-5)
        assert torch.equal(alg.transition.actions_log_prob, st.actions_log_prob[s, :, 0])
        zs.append(((st.actions[s] - mu) / ac.std.detach()).flatten())
        alg.process_env_step(torch.full((N,), 0.25 + s, device="cuda"), torch.arange(N, device="cuda") % 7 == s,
                             {"time_outs": torch.arange(N, device="cuda") % 14 == s})
        assert st.step == s + 1
        assert torch.equal(st.rewards[s, :, 0], torch.full((N,), 0.25 + s, device="cuda"))
        assert torch.equal(st.dones[s, :, 0].bool(), torch.arange(N, device="cuda") % 7 == s)
        assert torch.equal(st.time_outs[s].bool(), torch.arange(N, device="cuda") % 14 == s)
    z = torch.cat(zs)
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02, "the sample is not a standard normal"
    assert not torch.equal(zs[0], zs[1])
    # long `dones` (the reference's reset_buf dtype at allocation, base_task.py:72) and no time-out entry
    alg.act(obs, cobs)
    alg.process_env_step(torch.ones(N, device="cuda"), torch.ones(N, dtype=torch.long, device="cuda"), {})
    assert bool(alg.storage.dones[3].all()) and not bool(alg.storage.time_outs[3].any())
    with pytest.raises(AssertionError):
        alg.act(obs, cobs)


@pytest.mark.parametrize("world,n", [(2, 33657), (4, 33657), (3, 1001), (2, 590000)])
def test_optimizer_step_peers_emulated_ranks(world, n):
    """b200gym_ppo_optimizer_step_peers with `world` emulated ranks on ONE GPU: every rank's launch runs on its own stream with
    local tensors standing in for the symmetric buffers, so the push / flag / wait / rank-ordered-sum protocol runs for real
    (the kernels wait for each other).  Two consecutive exchanges (both slot parities).  All ranks must hold bit-identical
    parameters equal to clip_grad_norm_ + Adam on the rank-ordered gradient sum; the 2+ GPU wiring is tools/ddp_train_check.py."""
    from legged_gym_dev_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda")
    n_pad = (n + 2 + 3) // 4 * 4
    g = torch.Generator(device="cuda").manual_seed(world)
    p0 = torch.randn(n, generator=g, device=dev) * 0.1
    sym = [torch.zeros(2 * world * n_pad + 32, device=dev) for _ in range(world)]
    peers = _lib.PeerBasesPOD()
    for r in range(world):
        peers.base[r] = sym[r].data_ptr()
    tab = _lib.PackTablePOD()
    tab.n, tab.total = 1, n
    tab.e[0].src_off, tab.e[0].dst_off, tab.e[0].elem_end, tab.e[0].rows, tab.e[0].cols, tab.e[0].ld, tab.e[0].layout = 0, 0, n, 1, n, n, 0
    R = [dict(param=p0.clone(), grad=torch.zeros(n + 8, device=dev), gsum=torch.zeros(n_pad, device=dev), m=torch.zeros(n, device=dev),
              v=torch.zeros(n, device=dev), lr=torch.tensor([1e-3], device=dev), step=torch.zeros(1, dtype=torch.int32, device=dev),
              mb=torch.zeros(4, dtype=torch.double, device=dev), tot=torch.zeros(8, dtype=torch.double, device=dev),
              ws=torch.zeros(_lib.PEER_WS_BYTES // 4, dtype=torch.int32, device=dev), w16=torch.zeros(n, dtype=torch.float16, device=dev),
              stream=torch.cuda.Stream()) for _ in range(world)]
    ref_p, ref_m, ref_v, lr = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev), 1e-3
    op = _lib.OptParamsPOD()
    op.n, op.count, op.adaptive, op.desired_kl, op.max_grad_norm, op.beta1, op.beta2, op.eps = n, 100.0, 1, 0.01, 1.0, 0.9, 0.999, 1e-8
    for it in range(2):
        grads = [torch.randn(n, generator=g, device=dev) * (0.01 if it == 0 else 1e-4) for _ in range(world)]
        kls = [0.05 * (r + 1) * 100.0 if it == 0 else 0.0001 * 100.0 for r in range(world)]
        for r in range(world):
            R[r]["grad"][:n].copy_(grads[r])
            R[r]["mb"][0] = kls[r]
        torch.cuda.synchronize()
        for r in range(world):
            d = R[r]
            with torch.cuda.stream(d["stream"]):
                _lib.check(L.b200gym_ppo_optimizer_step_peers(op, peers, world, r, n_pad, *[_lib.ptr(d[k]) for k in
                                                              ("param", "grad", "gsum", "m", "v", "lr", "step", "mb", "tot", "ws")],
                                                              tab, _lib.ptr(d["w16"]), 0, d["stream"].cuda_stream), "ppo_optimizer_step_peers")
        torch.cuda.synchronize()
        # reference: rank-ordered fp32 sum, KL schedule on the global mean, clip_grad_norm_, torch Adam arithmetic
        tot = grads[0].clone()
        for r in range(1, world):
            tot = tot + grads[r]
        kl_mean = sum(kls) / (100.0 * world)
        if kl_mean > 0.02:
            lr = max(1e-5, lr / 1.5)
        elif 0.0 < kl_mean < 0.005:
            lr = min(1e-2, lr * 1.5)
        norm = float(torch.sqrt((tot.double() ** 2).sum()))
        gc = tot * min(1.0 / (norm + 1e-6), 1.0)
        ref_m = 0.9 * ref_m + 0.1 * gc
        ref_v = 0.999 * ref_v + 0.001 * gc * gc
        t = it + 1
        ref_p = ref_p - (lr / (1 - 0.9 ** t)) * ref_m / (ref_v.sqrt() / (1 - 0.999 ** t) ** 0.5 + 1e-8)
        for r in range(world):
            d = R[r]
            assert int(d["ws"][4]) == 0, "a rank timed out waiting for its peers"
            assert torch.equal(d["gsum"][:n], tot), f"rank {r}: summed gradients differ from the rank-ordered sum"
            assert torch.equal(d["param"], R[0]["param"]) and torch.equal(d["m"], R[0]["m"]), f"rank {r} diverged from rank 0"
            assert float(d["grad"].abs().max()) == 0.0 and int(d["step"]) == t and float(d["mb"][0]) == 0.0
            assert abs(float(d["lr"]) - lr) <= 1e-9
            assert abs(float(d["tot"][4]) - norm * norm) <= 1e-9 * norm * norm
            assert torch.equal(d["w16"], d["param"].half())
        assert_close(R[0]["param"].cpu(), ref_p.cpu(), 1.0, f"parameters after exchange {it}", rtol=2e-6)
