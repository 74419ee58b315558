"""CPU restatement (torch, fp32) of the rsl_rl v1.0.2 arithmetic on the hot path.  TEST INFRASTRUCTURE.

PARITY UNPINNED: rsl_rl is a third-party dependency (leggedrobotics/rsl_rl tag v1.0.2, reference README.md:33-35;
the reference imports a *fork* of it at legged_gym/utils/task_registry.py:37-38) whose source is absent from
/root/reference and from this image, and the reference holds no test or golden vector for it.  What is restated is
the published algorithm of that tag:
    rsl_rl/storage/rollout_storage.py   RolloutStorage.compute_returns, mini_batch_generator
    rsl_rl/algorithms/ppo.py            PPO.process_env_step (time-out bootstrap), PPO.update
    rsl_rl/modules/actor_critic.py      ActorCritic (MLP actor/critic, ELU, learned std, diagonal Normal)
anchored on the reference's call sites: OnPolicyRunner(env, train_cfg_dict, log_dir, device=, wandb_callback=)
(task_registry.py:148), runner.learn(num_learning_iterations=, init_at_random_ep_len=) (scripts/train.py:44), the
algorithm/policy/runner config keys (legged_robot_config.py:236-279) and extras["time_outs"] (legged_robot.py:186-187).
"""
import torch
import torch.nn as nn


def process_env_step_bootstrap(rewards, values, time_outs, gamma):
    """PPO.process_env_step: rewards += gamma * squeeze(values * time_outs.unsqueeze(1), 1)."""
    return rewards + gamma * torch.squeeze(values * time_outs.unsqueeze(1).to(values.dtype), 1)


def compute_returns(rewards, values, dones, last_values, gamma, lam):
    """RolloutStorage.compute_returns.  rewards/values/dones: [T, N, 1]; last_values: [N, 1].
    Returns (returns, un-normalised advantages, normalised advantages)."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    advantage = 0
    for step in reversed(range(T)):
        next_values = last_values if step == T - 1 else values[step + 1]
        next_is_not_terminal = 1.0 - dones[step].float()
        delta = rewards[step] + next_is_not_terminal * gamma * next_values - values[step]
        advantage = delta + next_is_not_terminal * gamma * lam * advantage
        returns[step] = advantage + values[step]
    adv = returns - values
    return returns, adv, (adv - adv.mean()) / (adv.std() + 1e-8)


def mini_batch_indices(T, N, num_mini_batches, num_epochs, generator=None):
    """RolloutStorage.mini_batch_generator index plan: ONE randperm(T*N), reused by every epoch."""
    batch = T * N
    mb = batch // num_mini_batches
    perm = torch.randperm(num_mini_batches * mb, generator=generator)
    return [perm[i * mb:(i + 1) * mb] for _ in range(num_epochs) for i in range(num_mini_batches)]


class ActorCritic(nn.Module):
    """rsl_rl/modules/actor_critic.py: MLPs with ELU, learned per-action std, Normal(mean, mean*0 + std)."""

    def __init__(self, num_obs, num_critic_obs, num_actions, actor_hidden_dims, critic_hidden_dims, init_noise_std=1.0):
        super().__init__()

        def mlp(i, hs, o):
            layers, d = [], i
            for h in hs:
                layers += [nn.Linear(d, h), nn.ELU()]
                d = h
            return nn.Sequential(*layers, nn.Linear(d, o))
        self.actor = mlp(num_obs, actor_hidden_dims, num_actions)
        self.critic = mlp(num_critic_obs, critic_hidden_dims, 1)
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))

    def dist(self, obs):
        mean = self.actor(obs)
        return torch.distributions.Normal(mean, mean * 0.0 + self.std)


def ppo_loss(ac, batch, clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.01, use_clipped_value_loss=True):
    """The loss of PPO.update for one minibatch; also returns kl_mean for the adaptive schedule."""
    d = ac.dist(batch["obs"])
    logp = d.log_prob(batch["actions"]).sum(dim=-1)
    value = ac.critic(batch["critic_obs"])
    mu, sigma, entropy = d.mean, d.stddev, d.entropy().sum(dim=-1)
    with torch.no_grad():
        kl = torch.sum(torch.log(sigma / batch["old_sigma"] + 1.e-5) + (torch.square(batch["old_sigma"]) +
                       torch.square(batch["old_mu"] - mu)) / (2.0 * torch.square(sigma)) - 0.5, axis=-1)
        kl_mean = torch.mean(kl)
    ratio = torch.exp(logp - torch.squeeze(batch["old_log_prob"]))
    adv = torch.squeeze(batch["advantages"])
    surrogate = -adv * ratio
    surrogate_clipped = -adv * torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param)
    surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
    if use_clipped_value_loss:
        value_clipped = batch["values"] + (value - batch["values"]).clamp(-clip_param, clip_param)
        value_loss = torch.max((value - batch["returns"]).pow(2), (value_clipped - batch["returns"]).pow(2)).mean()
    else:
        value_loss = (batch["returns"] - value).pow(2).mean()
    loss = surrogate_loss + value_loss_coef * value_loss - entropy_coef * entropy.mean()
    return loss, dict(kl_mean=kl_mean, surrogate_loss=surrogate_loss, value_loss=value_loss, entropy=entropy.mean(),
                      mu=mu, value=value)


def adaptive_lr(lr, kl_mean, desired_kl):
    if kl_mean > desired_kl * 2.0:
        return max(1e-5, lr / 1.5)
    if kl_mean < desired_kl / 2.0 and kl_mean > 0.0:
        return min(1e-2, lr * 1.5)
    return lr


def ppo_update(ac, optimizer, storage, plan, lr, desired_kl=0.01, max_grad_norm=1.0, schedule="adaptive", **loss_kw):
    """PPO.update over a precomputed minibatch index plan.  storage: dict of flattened [T*N, ...] tensors."""
    stats = []
    for idx in plan:
        batch = {k: v[idx] for k, v in storage.items()}
        loss, info = ppo_loss(ac, batch, **loss_kw)
        if desired_kl is not None and schedule == "adaptive":
            lr = adaptive_lr(lr, float(info["kl_mean"]), desired_kl)
            for g in optimizer.param_groups:
                g["lr"] = lr
        optimizer.zero_grad()
        loss.backward()
        gn = nn.utils.clip_grad_norm_(ac.parameters(), max_grad_norm)
        optimizer.step()
        stats.append(dict(loss=float(loss), kl=float(info["kl_mean"]), grad_norm=float(gn), lr=lr,
                          surrogate=float(info["surrogate_loss"]), value_loss=float(info["value_loss"])))
    return lr, stats


def sample_actions(mu, std, seed, env_ids, event):
    """Specification of the action sample of the fused PPO.act (legged_gym_dev_b200/csrc/ppo_rollout.cu): rsl_rl draws
    Normal(mu, std).sample() from torch's global generator; here the standard normals come from the library's Philox stream
    (site POLICY_SAMPLE, counter = (global env id, act event)) through Box-Muller — per Philox block (w0, w1, w2, w3):
    columns 4b + {0, 1} = sqrt(-2 ln u(w0)) * {cos, sin}(2 pi u(w1)), columns 4b + {2, 3} likewise from (w2, w3), with
    u(w) = ((w >> 8) + 0.5) 2^-24 under the logarithm and (w >> 8) 2^-24 in the angle.  Returns (actions, summed log-prob) as
    float32 tensors, computed with torch's Normal.log_prob formula."""
    import numpy as np
    from . import philox
    A = mu.shape[1]
    nb = (A + 3) // 4
    w = philox._words(seed, env_ids, event, philox.SITE_POLICY_SAMPLE, nb * 4).reshape(-1, nb, 4)
    f = np.float32
    ulog = ((w[..., 0::2] >> np.uint32(8)).astype(f) + f(0.5)) * f(2.0 ** -24)
    uang = (w[..., 1::2] >> np.uint32(8)).astype(f) * f(2.0 ** -24)
    r = np.sqrt(f(-2.0) * np.log(ulog)).astype(f)
    ang = (f(6.283185307179586) * uang).astype(f)
    z = np.stack((r * np.cos(ang), r * np.sin(ang)), axis=-1).astype(f).reshape(-1, nb * 4)[:, :A]
    z = torch.from_numpy(z)
    actions = mu + std * z
    logp = torch.distributions.Normal(mu, mu * 0.0 + std).log_prob(actions).sum(dim=-1)
    return actions, logp
