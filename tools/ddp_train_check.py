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
mode = "peer-memory kernel" if getattr(alg, "_peer", None) is not None else "NCCL all-reduce"
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
infos = runner.learn(num_learning_iterations=2, init_at_random_ep_len=True)
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
