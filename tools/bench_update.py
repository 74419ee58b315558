"""PPO.update timing (BASELINE cfg 5: 4096 envs x 24 steps, 5 epochs x 4 minibatches) for the flat and the rough nets,
plus the launch list of one update (torch profiler) — written as JSON to stdout.  `python tools/bench_update.py [--rough-only]`"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(num_obs, hidden, N=4096, T=24, reps=5, profile=True):
    from legged_gym_dev_b200.ppo import ActorCritic, PPO
    torch.manual_seed(0)
    ac = ActorCritic(num_obs, num_obs, 12, actor_hidden_dims=hidden, critic_hidden_dims=hidden)
    alg = PPO(ac, num_learning_epochs=5, num_mini_batches=4, schedule="adaptive", desired_kl=0.01, entropy_coef=0.01, device="cuda")
    alg.init_storage(N, T, [num_obs], [None], [12])
    st = alg.storage
    g = torch.Generator(device="cuda").manual_seed(1)
    st.observations.copy_(torch.randn(st.observations.shape, device="cuda", generator=g))
    with torch.no_grad():
        flat = st.observations.flatten(0, 1)
        mu = ac.act_inference(flat).view(T, N, 12)
        st.mu.copy_(mu)
        st.sigma.fill_(1.0)
        st.actions.copy_(mu + torch.randn(mu.shape, device="cuda", generator=g))
        st.actions_log_prob.copy_((-0.5 * (st.actions - mu) ** 2 - 0.9189385).sum(-1, keepdim=True))
        st.values.copy_(ac.evaluate(flat).view(T, N, 1))
    st.rewards.normal_(0.02, 0.05)
    last = torch.randn(N, 1, device="cuda")
    st.compute_returns(last, 0.99, 0.95)
    alg.update()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for _ in range(reps):
        st.step = T
        torch.cuda.synchronize()
        ev[0].record()
        alg.update()
        ev[1].record()
        torch.cuda.synchronize()
        times.append(ev[0].elapsed_time(ev[1]))
    out = dict(nets=f"{num_obs}-" + "-".join(map(str, hidden)), N=N, T=T, update_ms=min(times), update_ms_all=times,
               per_minibatch_us=min(times) / 20 * 1e3)
    params = sum(p.numel() for p in ac.parameters())
    flops = 3 * 2 * (params - 12) * T * N * 5
    out["tflops"] = flops / (min(times) * 1e-3) / 1e12
    if profile:
        from torch.profiler import profile as prof_, ProfilerActivity
        st.step = T
        with prof_(activities=[ProfilerActivity.CUDA]) as pr:
            alg.update()
            torch.cuda.synchronize()
        rows = sorted(((e.key, e.count, e.device_time_total) for e in pr.key_averages() if e.device_time_total > 0), key=lambda r: -r[2])
        out["kernels"] = [dict(name=k[:90], count=c, total_us=round(t, 1)) for k, c, t in rows[:20]]
    return out


if __name__ == "__main__":
    res = []
    prof = "--no-profile" not in sys.argv
    reps = 1 if "--once" in sys.argv else 5
    if "--rough-only" not in sys.argv:
        res.append(run(48, (128, 64, 32), profile=prof, reps=reps))
    if "--flat-only" not in sys.argv:
        res.append(run(235, (512, 256, 128), profile=prof, reps=reps))
    print(json.dumps(res, indent=1))
