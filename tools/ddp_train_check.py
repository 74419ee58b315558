"""2+ GPU check of the data-parallel training path (run under torchrun): envs sharded by rank, NCCL all-reduce of the
advantage statistics and of the flat gradient buffer; parameters must stay bit-identical across ranks."""
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from legged_gym_dev_b200 import synthetic as S                                   # noqa: E402
from legged_gym_dev_b200.physics import ReplayPhysics                            # noqa: E402
from legged_gym_dev_b200.sharding import env_shard, init_distributed             # noqa: E402
from legged_gym_dev_b200.task_registry import task_registry                      # noqa: E402

rank, local_rank, world = init_distributed()
dev = f"cuda:{local_rank}"
torch.cuda.set_device(dev)
TOTAL = 4096
lo, hi = env_shard(rank, world, TOTAL)
n = hi - lo
tape = S.make_state_tape(n, frames=4, seed=10 + rank, device=dev)
args = SimpleNamespace(num_envs=n, sim_device=dev, headless=True, physics_engine=None)
env, _ = task_registry.make_env("anymal_c_flat_b200", args=args, physics=ReplayPhysics(tape, device=dev), env_id_offset=lo)
torch.manual_seed(1)                                                            # same initial policy on every rank
runner, _ = task_registry.make_alg_runner(env, name="anymal_c_flat_b200", args=args)
p0 = runner.alg.actor_critic.flat_param.clone()
alg = runner.alg
mode = alg.exchange
if world > 1 and getattr(alg, "_xchg", None) is not None:
    # unit check of the in-kernel exchange against NCCL on random gradients: after one fused step grad_sum holds the rank-ordered sum
    ac, x = alg.actor_critic, alg._xchg
    state = [t.clone() for t in (ac.flat_param, alg.optimizer.exp_avg, alg.optimizer.exp_avg_sq, alg.optimizer.lr, alg.optimizer.step_dev, alg._scalars)]
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    ac.flat_grad[:ac.num_flat].copy_(torch.randn(ac.num_flat, generator=g, device=dev))
    want = ac.flat_grad[:ac.num_flat].clone()
    dist.all_reduce(want)
    alg._scalars[8] = 0.25 * (rank + 1)
    alg.optimizer.fused_step(1.0, alg._scalars[8:12], alg._scalars, 100.0, 0.01, xchg=x)
    torch.cuda.synchronize()
    got = x.grad_sum[:ac.num_flat].clone()
    err = float((got - want).abs().max())
    tail = x.grad_sum[ac.num_flat:ac.num_flat + 2].tolist()
    allg = [torch.empty_like(x.grad_sum) for _ in range(world)]
    dist.all_gather(allg, x.grad_sum)
    same_sum = all(torch.equal(allg[0], t) for t in allg)
    cleared = float(ac.flat_grad.abs().max()) == 0.0
    if rank == 0:
        print(f"fused exchange vs NCCL: max |diff| {err:.2e}, tail {tail} (want {[0.25 * world * (world + 1) / 2, 100.0 * world]}), "
              f"bit-identical across ranks: {same_sum}, local grads cleared: {cleared}, error flag: {x.error()}")
    assert err < 1e-5 and same_sum and cleared and not x.error()
    assert abs(tail[0] - 0.25 * world * (world + 1) / 2) < 1e-5 and tail[1] == 100.0 * world
    for t, k in zip((ac.flat_param, alg.optimizer.exp_avg, alg.optimizer.exp_avg_sq, alg.optimizer.lr, alg.optimizer.step_dev, alg._scalars), state):
        t.copy_(k)
    alg.optimizer.steps = 0
    ac.repack_fused()
if world > 1 and getattr(alg, "_peer", None) is not None:
    # unit check of the peer-memory reduction against NCCL on random buffers
    pr = alg._peer
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    pr.buf.copy_(torch.randn(pr.numel, generator=g, device=dev))
    want = pr.buf.clone()
    dist.all_reduce(want)
    sumsq = torch.zeros(1, dtype=torch.double, device=dev)
    got = pr.reduce(alg.actor_critic.num_flat, sumsq).clone()
    err = float((got - want).abs().max())
    nerr = abs(float(sumsq) - float((got[:alg.actor_critic.num_flat].double() ** 2).sum())) / float(sumsq)
    allg = [torch.empty_like(got) for _ in range(world)]
    dist.all_gather(allg, got)
    same_sum = all(torch.equal(allg[0], t) for t in allg)
    if rank == 0:
        print(f"peer reduce vs NCCL: max |diff| {err:.2e}, sumsq rel err {nerr:.1e}, bit-identical across ranks: {same_sum}")
    assert err < 1e-5 and nerr < 1e-12 and same_sum
    pr.buf.zero_()
runner.log_episode_stats = True
infos = runner.learn(num_learning_iterations=2, init_at_random_ep_len=True)
if world > 1:
    # reward statistics: every rank must log the same (all-reduced) means
    ep = {k: float(v) for k, v in infos[-1]["episode"].items()}
    allep = [None] * world
    dist.all_gather_object(allep, ep)
    if rank == 0:
        same_stats = all(all(a[k] == allep[0][k] or (a[k] != a[k] and allep[0][k] != allep[0][k]) for k in ep) for a in allep)
        print("episode statistics identical on all ranks:", same_stats, {k: round(v, 6) for k, v in list(ep.items())[:3]})
        assert same_stats
# time the update alone (rollout data of the last iteration is gone: refill the storage with one more rollout inside learn)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
infos += runner.learn(num_learning_iterations=3, init_at_random_ep_len=False)
torch.cuda.synchronize()
ms_iter = (time.perf_counter() - t0) / 3 * 1e3
p = runner.alg.actor_critic.flat_param
ref = p.clone()
if world > 1:
    dist.broadcast(ref, src=0)
same = bool(torch.equal(ref, p))
changed = not bool(torch.equal(p0, p))
lr = float(runner.alg.optimizer.lr)
gathered = [None] * world
if world > 1:
    dist.all_gather_object(gathered, (rank, same, changed, lr, float(infos[-1]["mean_value_loss"])))
else:
    gathered = [(rank, same, changed, lr, float(infos[-1]["mean_value_loss"]))]
if rank == 0:
    print("ddp_train_check:", gathered)
    assert all(g[1] and g[2] for g in gathered), "parameters diverged across ranks or did not change"
    assert len({g[3] for g in gathered}) == 1, "learning-rate decisions differ across ranks"
    print("OK: parameters bit-identical across", world, "ranks; lr", lr, "| gradient exchange:", mode, "| %.1f ms per training iteration (24 env steps + update)" % ms_iter)
if world > 1:
    dist.destroy_process_group()
