"""The multi-GPU configurations of BASELINE.json, timed on every rank under `bench.py --gpus N` (N = 1 included, so that the
driver's N = 1, 2, 4, 8 runs form three scaling curves next to the weak-scaling headline):

  cfg4_rom_rollout_sharded   configs[3]: ROM tube data collection, 1 048 576 envs TOTAL sharded over the N ranks (strong scaling)
  cfg2_step_total_1m         configs[1] at 1 048 576 envs TOTAL = 131 072 x 8 / N per GPU (strong scaling of the headline step)
  cfg5_train_iteration       configs[4]: 4096 envs TOTAL x 24 steps rollout + GAE + PPO update (5 epochs x 4 minibatches), the
                             gradient exchange inside the optimiser kernel over NVLink peer memory

Every number is device-timed (CUDA events) between barriers and reduced with MAX over the ranks.  All ranks must call
run_all() (the functions contain collectives); rank 0 attaches the result under "multi_gpu" of the bench line.
"""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _max_over_ranks(ms, device, world):
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def _timed(fn, device, world, reps=1):
    """barrier + sync | `reps` x fn() between CUDA events | sync + barrier; max over ranks, per repetition."""
    import torch.distributed as dist
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    return _max_over_ranks(a.elapsed_time(b), device, world) / reps


def rom_rollout_sharded(rank, world, device, total=1 << 20, T=200):
    from legged_gym_dev_b200 import configs
    from legged_gym_dev_b200.rom import CustomSim
    from legged_gym_dev_b200.sharding import env_shard
    lo, hi = env_shard(rank, world, total)
    n = hi - lo
    env = CustomSim(configs.double_single_int_cfg(n, seed=0), device=device, env_id_offset=lo)
    obs = torch.zeros(n, 8, device=device)
    env.collect_epoch(obs, 8)
    warm = env.collect_epoch(obs, T)       # same log sizes: the timed call reuses the allocator's cached blocks
    del warm
    ms = _timed(lambda: env.collect_epoch(obs, T), device, world)
    loop_steps = 2 * T + 11
    return dict(config="ROM tube data collection epoch (reset + 200 ROM steps), persistent rollout kernel, envs sharded by global env id",
                total_envs=total, envs_per_gpu=n, rom_steps=T, ms_epoch=ms, env_loop_steps_per_s=total * loop_steps / (ms * 1e-3),
                scaling="strong", collectives_in_data_path=0)


def step_total_1m(rank, world, device, total=1 << 20, steps=96, frames=4):
    sys.path.insert(0, ROOT)
    import bench
    from legged_gym_dev_b200.graphs import GraphedReplay
    n = total // world
    env, tape = bench.build_env(n, frames, device, rank)
    acts = [tape.actions[f].to(device) for f in range(frames)]
    g = GraphedReplay(env, acts)
    for _ in range(3):
        g.replay()
    ms = _timed(g.replay, device, world, reps=steps // frames) / frames
    ab = bench.algorithmic_bytes(len(env.params.active_terms))
    peak, _ = bench.measured_peak()
    out = dict(config="anymal_c_flat env.step, upstream reward table; 1 048 576 envs TOTAL split over the ranks; CUDA-graph replay of a "
                      f"{frames}-step tape cycle", total_envs=total, envs_per_gpu=n, ms_per_step=ms, env_steps_per_s=total / (ms * 1e-3),
               step_frac_of_hbm_peak=ab["step"] * n / (ms * 1e-3) / 1e9 / peak, scaling="strong", collectives_in_data_path=0)
    del env, tape, g
    torch.cuda.empty_cache()
    return out


def train_iteration(rank, world, device, total=4096, iters=10):
    from legged_gym_dev_b200 import synthetic as S
    from legged_gym_dev_b200.physics import ReplayPhysics
    from legged_gym_dev_b200.sharding import env_shard
    from legged_gym_dev_b200.task_registry import task_registry
    lo, hi = env_shard(rank, world, total)
    n = hi - lo
    tape = S.make_state_tape(n, frames=4, seed=10 + rank, device=device)
    args = SimpleNamespace(num_envs=n, sim_device=str(device), headless=True, physics_engine=None)
    env, _ = task_registry.make_env("anymal_c_flat_b200", args=args, physics=ReplayPhysics(tape, device=device), env_id_offset=lo)
    torch.manual_seed(1)                                  # the same initial policy on every rank
    runner, _ = task_registry.make_alg_runner(env, name="anymal_c_flat_b200", args=args)
    alg = runner.alg
    runner.learn(num_learning_iterations=2, init_at_random_ep_len=True)      # warm-up: allocations, graph capture
    ms_iter = _timed(lambda: runner.learn(num_learning_iterations=1), device, world, reps=iters)
    # the same iteration with the rollout replayed from one CUDA graph (the replay tape makes the env step capturable)
    runner.graph_rollout = True
    runner.learn(num_learning_iterations=2)
    ms_iter_g = _timed(lambda: runner.learn(num_learning_iterations=1), device, world, reps=iters)
    runner.graph_rollout = False
    # the update alone, on the storage of the last rollout (update() only resets the write cursor)
    T = alg.storage.num_transitions_per_env

    def upd():
        alg.storage.step = T
        alg.update()
    upd()
    ms_upd = _timed(upd, device, world, reps=iters)
    p = alg.actor_critic.flat_param
    same = True
    if world > 1:
        import torch.distributed as dist
        ref = p.clone()
        dist.broadcast(ref, src=0)
        ok = torch.tensor([int(torch.equal(ref, p))], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())
        if getattr(alg, "_xchg", None) is not None and alg._xchg.error():
            same = False
    return dict(config="anymal_c_flat PPO training iteration: 24 env steps (act + env.step + storage) + GAE + update of 5 epochs x 4 "
                       "minibatches, nets 48-128-64-32; 4096 envs TOTAL sharded over the ranks",
                total_envs=total, envs_per_gpu=n,
                # headline: rollout replayed from one CUDA graph (GPU-bound); the eager-rollout figure (~10 Python-issued launches per env
                # step, max over ranks) is host-launch bound and jitters with the host
                ms_per_iteration=ms_iter_g, update_ms=ms_upd, rollout_ms=ms_iter_g - ms_upd,
                ms_per_iteration_eager_rollout=ms_iter, rollout_ms_eager=ms_iter - ms_upd,
                samples_per_s=total * T / (ms_iter_g * 1e-3), gradient_exchange=alg.exchange, params_bit_identical_across_ranks=same,
                launches_per_minibatch=2, scaling="strong")


def run_all(rank, world, device):
    out = {}
    for name, fn in (("cfg4_rom_rollout_sharded", rom_rollout_sharded), ("cfg2_step_total_1m", step_total_1m),
                     ("cfg5_train_iteration", train_iteration)):
        try:
            out[name] = fn(rank, world, device)
        except Exception as e:   # an extra must never take the headline line down (all ranks raise alike: the collectives stay matched)
            out[name] = dict(error=f"{type(e).__name__}: {e}")
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    import json
    from legged_gym_dev_b200.sharding import init_distributed
    rank, local_rank, world = init_distributed()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    res = run_all(rank, world, dev)
    if rank == 0:
        print(json.dumps(res, indent=1))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
