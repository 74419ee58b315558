"""End-to-end drop-in check on the GPU: task_registry.make_env -> make_alg_runner -> learn(), the reference's train()
call sequence (legged_gym/scripts/train.py:41-44) with the replay physics standing in for Isaac Gym."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_call_sequence():
    from types import SimpleNamespace
    from legged_gym_dev_b200 import synthetic as S
    from legged_gym_dev_b200.physics import ReplayPhysics
    from legged_gym_dev_b200.task_registry import task_registry
    N = 512
    tape = S.make_state_tape(N, frames=8, seed=0, device="cuda")
    args = SimpleNamespace(num_envs=N, sim_device="cuda", headless=True, physics_engine=None)
    env, env_cfg = task_registry.make_env("anymal_c_flat_b200", args=args, physics=ReplayPhysics(tape, device="cuda"))
    assert env.num_envs == N and env.num_obs == 48 and env.obs_buf.shape == (N, 48)
    runner, train_cfg = task_registry.make_alg_runner(env, name="anymal_c_flat_b200", args=args)
    p0 = runner.alg.actor_critic.flat_param.clone()
    infos = runner.learn(num_learning_iterations=2, init_at_random_ep_len=True)
    assert len(infos) == 2
    for i in infos:
        assert torch.isfinite(i["mean_value_loss"]) and torch.isfinite(i["mean_surrogate_loss"])
    assert not torch.equal(p0, runner.alg.actor_critic.flat_param), "the update did not change the parameters"
    assert env.common_step_counter == 2 * 24
    policy = runner.get_inference_policy(device=env.device)
    assert policy(env.get_observations()).shape == (N, 12)
    assert set(env.extras) >= {"episode", "time_outs"}


def test_storage_holds_the_observation_each_action_was_computed_from():
    """rsl_rl stores references to the obs tensors its env returns; the fused env rewrites ONE obs_buf in place, so PPO.act must
    store the observation at act time: storage.observations[t] == obs fed to act at step t, and observations[t + 1] == the obs
    returned by step t."""
    from types import SimpleNamespace
    from legged_gym_dev_b200 import synthetic as S
    from legged_gym_dev_b200.physics import ReplayPhysics
    from legged_gym_dev_b200.task_registry import task_registry
    N = 256
    tape = S.make_state_tape(N, frames=8, seed=1, device="cuda")
    args = SimpleNamespace(num_envs=N, sim_device="cuda", headless=True, physics_engine=None)
    env, _ = task_registry.make_env("anymal_c_flat_b200", args=args, physics=ReplayPhysics(tape, device="cuda"))
    runner, _ = task_registry.make_alg_runner(env, name="anymal_c_flat_b200", args=args)
    alg = runner.alg
    obs = env.get_observations()
    fed, returned = [], []
    for t in range(6):
        fed.append(obs.clone())
        actions = alg.act(obs, obs)
        mu = alg.actor_critic.act_inference(fed[-1])
        assert torch.equal(alg.storage.mu[t], mu), "mu was not computed from the stored observation"
        obs, _, rew, dones, infos = env.step(actions)
        returned.append(obs.clone())
        alg.process_env_step(rew, dones, infos)
        assert torch.equal(alg.storage.rewards[t, :, 0], rew) and torch.equal(alg.storage.dones[t, :, 0].bool(), dones)
    for t in range(6):
        assert torch.equal(alg.storage.observations[t], fed[t])
        if t + 1 < 6:
            assert torch.equal(alg.storage.observations[t + 1], returned[t])


def test_runner_episode_statistics_match_the_per_step_extras():
    """OnPolicyRunner's logged reward statistics (rsl_rl: mean over the rollout's ep_infos) are rebuilt from the RAW (sum, count) rows
    the step kernel leaves behind — the quantity the env shards all-reduce.  Single rank: they must equal the mean over the steps that
    saw a reset of extras["episode"] itself."""
    from types import SimpleNamespace
    from legged_gym_dev_b200 import synthetic as S
    from legged_gym_dev_b200.physics import ReplayPhysics
    from legged_gym_dev_b200.task_registry import task_registry
    N = 512
    tape = S.make_state_tape(N, frames=8, seed=3, device="cuda")
    args = SimpleNamespace(num_envs=N, sim_device="cuda", headless=True, physics_engine=None)
    env, _ = task_registry.make_env("anymal_c_flat_b200", args=args, physics=ReplayPhysics(tape, device="cuda"))
    env.episode_length_buf.copy_(S.make_episode_lengths(N, seed=1, device="cuda"))
    runner, _ = task_registry.make_alg_runner(env, name="anymal_c_flat_b200", args=args)
    runner.log_episode_stats = True
    seen = []
    orig_step = env.step

    def step(a):
        out = orig_step(a)
        if float(env.extras["num_resets"]) > 0:
            seen.append({k: float(v) for k, v in env.extras["episode"].items()})
        return out
    env.step = step
    infos = runner.learn(num_learning_iterations=1)
    assert seen, "no env reset during the rollout: the test tape must produce resets"
    ep = infos[0]["episode"]
    assert set(ep) == set(seen[0])
    for k in ep:
        want = sum(s[k] for s in seen) / len(seen)
        assert abs(float(ep[k]) - want) <= 1e-5 * max(1.0, abs(want)), (k, float(ep[k]), want)


def test_graphed_rollout_matches_eager_rollout():
    """GraphedRollout (T x [act, env.step, process_env_step] replayed from one CUDA graph, step and act counters in device memory) fills
    the rollout storage bit-identically to the eager loop, replay after replay, and OnPolicyRunner.learn runs with it."""
    from types import SimpleNamespace
    from legged_gym_dev_b200 import synthetic as S
    from legged_gym_dev_b200.graphs import GraphedRollout
    from legged_gym_dev_b200.physics import ReplayPhysics
    from legged_gym_dev_b200.task_registry import task_registry
    N = 384

    def collect(graph):
        tape = S.make_state_tape(N, frames=4, seed=2, device="cuda")
        args = SimpleNamespace(num_envs=N, sim_device="cuda", headless=True, physics_engine=None)
        env, _ = task_registry.make_env("anymal_c_flat_b200", args=args, physics=ReplayPhysics(tape, device="cuda"))
        env.episode_length_buf.copy_(S.make_episode_lengths(N, seed=1, device="cuda"))
        runner, _ = task_registry.make_alg_runner(env, name="anymal_c_flat_b200", args=args)
        alg, obs = runner.alg, env.get_observations()

        def rollout():
            for _ in range(runner.num_steps_per_env):
                actions = alg.act(obs, obs)
                _, _, rew, dones, infos = env.step(actions)
                alg.process_env_step(rew, dones, infos)
        out = []
        with torch.inference_mode():
            rollout()
            alg.storage.clear()
            g = GraphedRollout(runner, rollout) if graph else None
            for _ in range(2):
                g.replay() if graph else rollout()
                st = alg.storage
                out.append({k: getattr(st, k).clone() for k in ("observations", "actions", "rewards", "dones", "values", "actions_log_prob", "time_outs")})
                assert st.step == runner.num_steps_per_env
                st.clear()
        return out, env.common_step_counter, alg._act_event

    a, ca, ea = collect(False)
    b, cb, eb = collect(True)
    assert ca == cb == 72 and ea == eb == 72
    for ra, rb in zip(a, b):
        for k in ra:
            assert torch.equal(ra[k], rb[k]), f"storage.{k} differs between the eager and the graph-replayed rollout"
    assert not torch.equal(a[0]["actions"], a[1]["actions"])
    # end to end through the runner
    tape = S.make_state_tape(N, frames=4, seed=2, device="cuda")
    args = SimpleNamespace(num_envs=N, sim_device="cuda", headless=True, physics_engine=None)
    env, _ = task_registry.make_env("anymal_c_flat_b200", args=args, physics=ReplayPhysics(tape, device="cuda"))
    runner, _ = task_registry.make_alg_runner(env, name="anymal_c_flat_b200", args=args)
    runner.graph_rollout = True
    infos = runner.learn(num_learning_iterations=4)
    assert runner._rollout_graph is not None and env.common_step_counter == 4 * 24
    assert all(torch.isfinite(i["mean_value_loss"]) for i in infos)
