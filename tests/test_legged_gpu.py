"""Group R parity on the GPU: fused CUDA step pipeline (through the C ABI) vs the CPU oracle port.

Same seeded inputs, same counter-based random numbers; compared after EVERY step so that stateful
quantities (feet_air_time, LSTM h/c, episode sums, in-kernel resets) are checked along a trajectory.
Tolerances: oracle/compare.py (bit-exact flags/indices; 1e-5 relative with stated scale otherwise)."""
import pytest
import torch

import legged_case as LC

pytestmark = pytest.mark.gpu


def _run(name, num_envs, steps, frames=8, big_action_step=3):
    case = LC.build_case(name, num_envs, frames=frames)
    port, phys = LC.make_port(case)
    env = LC.make_fused(case)
    if case.traj:   # as reset() does: generators reset at the robots' positions (a never-reset generator divides 0/0 in the reference)
        port.gen.reset_traj(torch.arange(num_envs), port.proj_z())
        env.reset_traj(torch.arange(num_envs, device="cuda"))
    worst = {}
    resets = 0
    for s in range(steps):
        a = case.tape.actions[s % frames] * (150.0 if s == big_action_step else 1.0)   # exercises the action clip
        port.step(a.clone(), phys)
        env.step(a.cuda())
        w = LC.compare_snapshots(LC.snapshot_fused(env), LC.snapshot_port(port), tag=f"{name} step {s}: ")
        resets += int(port.reset_buf.sum())
        for k, v in w.items():
            worst[k] = max(worst.get(k, 0.0), v)
    return worst, resets


@pytest.mark.parametrize("name", list(LC.CASES))
def test_step_parity_small(name):
    worst, resets = _run(name, 256, 30)
    assert resets > 0, "case never exercised the reset path"
    print(name, "resets", resets, {k: f"{v:.1e}" for k, v in worst.items() if v > 0})


@pytest.mark.parametrize("num_envs", [4, 60, 68, 1000])
def test_ragged_sizes(num_envs):
    """Tail tiles (num_envs not a multiple of the 64-env tile) take the non-TMA path."""
    _run("flat_allterms_v", num_envs, 12)
    _run("rough_lstm_allterms", num_envs, 6)
    _run("traj_flat_allterms", num_envs, 8)
    _run("traj_rough_lstm_allterms", num_envs, 6)


def test_cfg2_size_4096():
    """BASELINE config 2 size: anymal_c_flat, 4096 envs."""
    worst, resets = _run("flat_pd_upstream", 4096, 12)
    assert resets > 0


def test_long_trajectory_timeouts():
    """Episode counters start near the limit so time-outs (1001st step, legged_robot.py:144) fire."""
    case = LC.build_case("flat_pd_upstream", 512, base_contact_prob=0.0)
    case.ep[:] = torch.randint(990, 1001, (512,), generator=torch.Generator().manual_seed(3))
    port, phys = LC.make_port(case)
    env = LC.make_fused(case)
    touts = 0
    for s in range(16):
        a = case.tape.actions[s % 8]
        port.step(a.clone(), phys)
        env.step(a.cuda())
        LC.compare_snapshots(LC.snapshot_fused(env), LC.snapshot_port(port), tag=f"timeouts step {s}: ")
        touts += int(port.time_out_buf.sum())
    assert touts > 0


def test_shard_invariance():
    """H6: per-env results do not depend on how envs are split over ranks (RNG keyed by global env id)."""
    N = 512
    case = LC.build_case("flat_pd_upstream", N)
    env = LC.make_fused(case)
    half = N // 2
    shard = LC.build_case("flat_pd_upstream", half)
    for k in ("root", "dof", "contact", "actions"):
        t = getattr(case.tape, k)
        if k == "dof":
            setattr(shard.tape, k, t.view(t.shape[0], t.shape[1], N, 12, 2)[:, :, half:].reshape(t.shape[0], t.shape[1], half * 12, 2).contiguous())
        else:
            setattr(shard.tape, k, t[:, half:].contiguous())
    shard.ep = case.ep[half:].clone()
    env2 = LC.make_fused(shard, env_id_offset=half)
    for s in range(10):
        env.step(case.tape.actions[s % 8].cuda())
        env2.step(shard.tape.actions[s % 8].cuda())
        assert torch.equal(env.obs_buf[half:], env2.obs_buf)
        assert torch.equal(env.rew_buf[half:], env2.rew_buf)
        assert torch.equal(env.reset_buf[half:], env2.reset_buf)
        assert torch.equal(env.commands[half:], env2.commands)


def test_errors_are_loud():
    from legged_gym_dev_b200 import _lib
    case = LC.build_case("flat_pd_upstream", 64)
    env = LC.make_fused(case)
    with pytest.raises(RuntimeError):
        env.step(torch.zeros(64, 12))          # CPU tensor: no fallback
    env._pod.control_type = 7
    with pytest.raises(RuntimeError, match="Unknown controller type"):
        env.step(torch.zeros(64, 12, device="cuda"))


def test_graph_replay_matches_eager():
    """CUDA-graph replay of tape cycles (device-resident step counter) == the eager API, bit for bit, incl. fresh RNG."""
    from legged_gym_dev_b200.graphs import GraphedReplay
    F = 4
    case = LC.build_case("flat_pd_upstream", 1024, frames=F)
    a = LC.make_fused(case, copy=False)
    case_b = LC.build_case("flat_pd_upstream", 1024, frames=F)
    b = LC.make_fused(case_b, copy=False)
    acts_a = [case.tape.actions[f].cuda() for f in range(F)]
    acts_b = [case_b.tape.actions[f].cuda() for f in range(F)]
    g = GraphedReplay(a, acts_a)           # runs one warm-up cycle eagerly
    for f in range(F):
        b.step(acts_b[f])
    obs_prev = None
    for cycle in range(3):
        g.replay()
        for f in range(F):
            b.step(acts_b[f])
        torch.cuda.synchronize()
        assert a.common_step_counter == b.common_step_counter
        for k in ("obs_buf", "rew_buf", "reset_buf", "commands", "episode_length_buf", "feet_air_time", "last_dof_vel"):
            assert torch.equal(getattr(a, k), getattr(b, k)), f"cycle {cycle}: {k}"
        assert torch.equal(a._sums, b._sums)
        if obs_prev is not None:
            assert not torch.equal(obs_prev, a.obs_buf), "replays must draw fresh noise"
        obs_prev = a.obs_buf.clone()


def test_host_replay_physics_matches_device_replay():
    """The pinned-host, double-buffered physics of bench.py's e2e arm feeds the same frames as the device replay."""
    from types import SimpleNamespace
    from legged_gym_dev_b200.legged_robot import Anymal
    from legged_gym_dev_b200.physics import HostReplayPhysics
    case = LC.build_case("flat_pd_upstream", 512, frames=4)
    a = LC.make_fused(case)
    lim = case.limits
    b = Anymal(case.cfg, SimpleNamespace(dt=case.cfg.sim.dt), None, "cuda", True, physics=HostReplayPhysics(case.tape, device="cuda"),
               asset=dict(dof_pos_limits=lim["dof_pos_limits"], dof_vel_limits=lim["dof_vel_limits"], torque_limits=lim["torque_limits"]),
               seed=case.seed)
    b.episode_length_buf = case.ep.cuda()
    for s in range(10):
        act = case.tape.actions[s % 4].cuda()
        a.step(act)
        b.step(act)
        torch.cuda.synchronize()
        if s == 0:
            continue   # torque evaluation 0 of the very first step sees the initial (unspecified) dof_state of each backend
        for k in ("obs_buf", "rew_buf", "reset_buf", "commands", "torques"):
            assert torch.equal(getattr(a, k), getattr(b, k)), f"step {s}: {k}"
        # the last sub-step's torques, read back to pinned host memory (what a host-side physics consumes)
        assert torch.equal(b.physics.torques_host[3], b.torques.cpu()), f"step {s}: torques read back to the host"


def test_height_cells_exact_at_scale():
    """3.7 M sample points: the truncated cell index (un-fused fp32 chain + exact division by horizontal_scale,
    SURVEY.md fact 10) must agree with the reference arithmetic for every point."""
    N = 20000
    case = LC.build_case("rough_pd_shipped", N, frames=2)
    # put a share of the robots exactly on cell boundaries and outside the field (clip path)
    r = case.tape.root
    r[:, : N // 4, 0] = torch.round(r[:, : N // 4, 0] * 10) / 10
    r[:, N // 4: N // 2, 1] = torch.round(r[:, N // 4: N // 2, 1] * 10) / 10
    r[:, -50:, 0] = -40.0
    r[:, -100:-50, 1] = 400.0
    port, phys = LC.make_port(case)
    env = LC.make_fused(case)
    for s in range(2):
        a = case.tape.actions[s]
        port.step(a.clone(), phys)
        env.step(a.cuda())
        got, want = env.measured_heights.cpu(), port.measured_heights
        bad = (got != want)
        assert not bad.any(), f"step {s}: {int(bad.sum())} of {bad.numel()} height samples differ"


@pytest.mark.parametrize("name,prob,ep0", [("flat_allterms_v", 1.0, None), ("rough_lstm_allterms", 1.0, None),
                                            ("flat_pd_upstream", 0.0, 0), ("rough_lstm_allterms", 0.0, 0)])
def test_all_reset_and_no_reset(name, prob, ep0):
    """Edge cases of SURVEY.md §4: every env resets every step (base contact on all envs) / no env ever resets
    (extras['episode'] must then keep its previous values, legged_robot.py:156-157)."""
    N = 200
    case = LC.build_case(name, N, base_contact_prob=prob)
    if prob == 1.0:
        case.tape.contact[:, :, 0, 2] = 50.0     # make sure the base force exceeds the threshold everywhere
    if ep0 is not None:
        case.ep[:] = ep0
    port, phys = LC.make_port(case)
    env = LC.make_fused(case)
    for s in range(8):
        a = case.tape.actions[s % 8]
        port.step(a.clone(), phys)
        env.step(a.cuda())
        LC.compare_snapshots(LC.snapshot_fused(env), LC.snapshot_port(port), tag=f"{name} p={prob} step {s}: ")
        n = int(port.reset_buf.sum())
        assert n == (N if prob == 1.0 else 0)
        assert float(env.extras["num_resets"]) == (N if prob == 1.0 else 0.0)
    if prob == 0.0:
        assert all(float(v) == 0.0 for v in env.extras["episode"].values())


def test_extreme_inputs_stay_finite():
    """Huge actions / velocities / forces: clips must hold and nothing may turn into NaN (compared against the oracle)."""
    N = 128
    case = LC.build_case("flat_allterms_v", N)
    case.tape.actions *= 1e4
    case.tape.root[..., 7:13] *= 50.0
    case.tape.contact *= 100.0
    case.tape.dof[..., 1] *= 30.0
    port, phys = LC.make_port(case)
    env = LC.make_fused(case)
    for s in range(6):
        a = case.tape.actions[s % 8]
        port.step(a.clone(), phys)
        env.step(a.cuda())
        snap = LC.snapshot_fused(env)
        assert all(torch.isfinite(v.float()).all() for v in snap.values())
        assert snap["obs"].abs().max() <= 100.0 and snap["torques"].abs().max() <= 80.0
        LC.compare_snapshots(snap, LC.snapshot_port(port), tag=f"extreme step {s}: ")


def test_tensor_core_lstm_variant_meets_the_torque_contract():
    """The actuator LSTM's gate mat-vecs on tcgen05 (3xTF32 operand split, fp32 accumulation): same 1e-5 (S = 80) parity
    against the oracle as the default FFMA2 kernel."""
    from legged_gym_dev_b200 import _lib
    L = _lib.lib()
    case = LC.build_case("flat_lstm_shipped", 1000)
    port, phys = LC.make_port(case)
    env = LC.make_fused(case)
    assert L.b200gym_debug_set_lstm_variant(1) == 0
    try:
        for s in range(12):
            a = case.tape.actions[s % 8]
            port.step(a.clone(), phys)
            env.step(a.cuda())
            LC.compare_snapshots(LC.snapshot_fused(env), LC.snapshot_port(port), tag=f"tc lstm step {s}: ")
    finally:
        L.b200gym_debug_set_lstm_variant(0)


@pytest.mark.parametrize("name,N", [("flat_pd_upstream", 1000), ("rough_lstm_allterms", 516), ("flat_heading_nonoise", 68), ("rough_pd_shipped", 20)])
def test_external_reset_matches_oracle(name, N):
    """reset_idx(env_ids) / reset() called from outside step() (legged_robot.py:147-187, base_task.py:111-119) against the port
    (pinned to the reference's own methods by tests/test_oracle_cpu.py::test_port_external_reset_tracks_unmodified_reference):
    the masked reset launch acts immediately, leaves time_out_buf alone and keys its draws with the external-reset event."""
    case = LC.build_case(name, N)
    port, phys = LC.make_port(case)
    env = LC.make_fused(case)
    for s in range(4):
        a = case.tape.actions[s % case.tape.frames]
        port.step(a.clone(), phys)
        env.step(a.cuda())
    ids = torch.arange(1, N, 3)
    port.reset_idx(ids)
    env.reset_idx(ids.cuda())
    LC.compare_snapshots(LC.snapshot_fused(env), LC.snapshot_port(port), tag=f"{name} partial reset: ")
    assert bool(env.reset_buf[ids.cuda()].all()) and int(env.episode_length_buf[ids.cuda()].abs().sum()) == 0
    for s in range(4, 7):
        a = case.tape.actions[s % case.tape.frames]
        port.step(a.clone(), phys)
        env.step(a.cuda())
        LC.compare_snapshots(LC.snapshot_fused(env), LC.snapshot_port(port), tag=f"{name} step {s} after the partial reset: ")
    o_port, _ = port.reset(phys)
    o_env, _ = env.reset()
    LC.compare_snapshots(LC.snapshot_fused(env), LC.snapshot_port(port), tag=f"{name} reset(): ")
    assert not bool(env.time_out_buf.any()), "BaseTask.reset must not leave time-out flags behind"
    env.reset_idx(torch.arange(0, device="cuda"))
    for s in range(7, 10):
        a = case.tape.actions[s % case.tape.frames]
        port.step(a.clone(), phys)
        env.step(a.cuda())
        LC.compare_snapshots(LC.snapshot_fused(env), LC.snapshot_port(port), tag=f"{name} step {s} after reset(): ")


@pytest.mark.parametrize("name,N", [("traj_flat_allterms", 512), ("traj_rough_lstm_allterms", 260)])
def test_trajectory_env_external_reset_matches_oracle(name, N):
    """AnymalTrajectory.reset() / reset_idx(env_ids) from outside step() (legged_robot_trajectory.py:204-246 + base_task.py:111-119) against
    the trajectory port (pinned to the reference by test_trajectory_port_external_reset_tracks_unmodified_reference): masked reset launch +
    generator reset from the new roots, immediate, generator state included; reset() starts the generators, so nothing is ever NaN."""
    case = LC.build_case(name, N)
    port, phys = LC.make_port(case)
    env = LC.make_fused(case)
    port.reset(phys)
    obs, _ = env.reset()
    assert torch.isfinite(obs).all()
    LC.compare_snapshots(LC.snapshot_fused(env), LC.snapshot_port(port), tag=f"{name} reset(): ")
    for s in range(4):
        a = case.tape.actions[s % case.tape.frames]
        port.step(a.clone(), phys)
        env.step(a.cuda())
        LC.compare_snapshots(LC.snapshot_fused(env), LC.snapshot_port(port), tag=f"{name} step {s} after reset(): ")
    ids = torch.arange(2, N, 3)
    port.reset_idx(ids)
    env.reset_idx(ids.cuda())
    LC.compare_snapshots(LC.snapshot_fused(env), LC.snapshot_port(port), tag=f"{name} partial reset: ")
    for s in range(4, 12):
        a = case.tape.actions[s % case.tape.frames]
        port.step(a.clone(), phys)
        obs, _, rew, _, _ = env.step(a.cuda())
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all() and torch.isfinite(env.prev_error).all()
        LC.compare_snapshots(LC.snapshot_fused(env), LC.snapshot_port(port), tag=f"{name} step {s} after the partial reset: ")
    assert float(env.traj_gen.k.min()) >= 0


@pytest.mark.parametrize("name,N", [("traj_flat_allterms", 1000), ("traj_rough_lstm_allterms", 132)])
def test_trajectory_env_generator_step_kernels_are_bit_identical(name, N, monkeypatch):
    """The trajectory env's generator step through the window-staging kernel (state window and the env's interpolated view moved as
    bulk tiles, csrc/rom_family.cu) against b200gym_rom_step (B200GYM_ENV_GEN_TILE=0): every buffer of the env and of the generator
    agrees bit for bit over a run with resets and new knots (N = 132: a 4-env tail CTA)."""
    case = LC.build_case(name, N)
    runs = {}
    for tile in ("1", "0"):
        monkeypatch.setenv("B200GYM_ENV_GEN_TILE", tile)
        env = LC.make_fused(case)
        env.reset_traj(torch.arange(N, device="cuda"))
        snaps = []
        for s in range(14):
            env.step(case.tape.actions[s % case.tape.frames].cuda())
            snap = LC.snapshot_fused(env)
            snap["gen_v"], snap["gen_t"], snap["gen_k"] = env.traj_gen.v.clone(), env.traj_gen.t.clone(), env.traj_gen.k.clone()
            snap["gen_traj"], snap["gen_vtraj"] = env.traj_gen.trajectory.clone(), env.traj_gen.v_trajectory.clone()
            snaps.append(snap)
        runs[tile] = snaps
    for s in range(14):
        for k, v in runs["1"][s].items():
            w = runs["0"][s][k]
            if torch.is_tensor(v):
                assert torch.equal(v, w), (s, k)


def test_trajectory_env_shard_invariance():
    """Per-env results of the trajectory env do not depend on the split over ranks (env and generator draws keyed by global id)."""
    N, half = 256, 128
    case = LC.build_case("traj_flat_allterms", N)
    env = LC.make_fused(case)
    shard = LC.build_case("traj_flat_allterms", half)
    for k in ("root", "dof", "contact", "actions"):
        t = getattr(case.tape, k)
        if k == "dof":
            setattr(shard.tape, k, t.view(t.shape[0], t.shape[1], N, 12, 2)[:, :, half:].reshape(t.shape[0], t.shape[1], half * 12, 2).contiguous())
        else:
            setattr(shard.tape, k, t[:, half:].contiguous())
    shard.ep, shard.tpush = case.ep[half:].clone(), case.tpush[half:].clone()
    env2 = LC.make_fused(shard, env_id_offset=half)
    # the flat env origins are a grid over the LOCAL env count: give the shard the rows of the full grid
    env2.env_origins.copy_(env.env_origins[half:])
    env.reset_traj(torch.arange(N, device="cuda"))
    env2.reset_traj(torch.arange(half, device="cuda"))
    for s in range(16):
        env.step(case.tape.actions[s % 8].cuda())
        env2.step(shard.tape.actions[s % 8].cuda())
        for name in ("obs_buf", "rew_buf", "reset_buf", "prev_error", "time_until_next_push", "trajectory"):
            a, b = getattr(env, name)[half:], getattr(env2, name)
            assert torch.equal(a, b), f"step {s}: {name} depends on the sharding"
        assert torch.equal(env.traj_gen.trajectory[half:], env2.traj_gen.trajectory)


def test_trajectory_env_graph_replay_matches_eager():
    """The three-kernel trajectory step (generator step, fused post-physics, gated generator reset) has no host branch: a captured
    tape cycle replays bit-identically to the eager API, generator state included."""
    from legged_gym_dev_b200.graphs import GraphedReplay
    F = 4
    ca, cb = LC.build_case("traj_flat_allterms", 1024, frames=F), LC.build_case("traj_flat_allterms", 1024, frames=F)
    a, b = LC.make_fused(ca, copy=False), LC.make_fused(cb, copy=False)
    ids = torch.arange(1024, device="cuda")
    a.reset_traj(ids)
    b.reset_traj(ids)
    acts_a = [ca.tape.actions[f].cuda() for f in range(F)]
    acts_b = [cb.tape.actions[f].cuda() for f in range(F)]
    g = GraphedReplay(a, acts_a)           # runs one warm-up cycle eagerly
    for f in range(F):
        b.step(acts_b[f])
    for cycle in range(3):
        g.replay()
        for f in range(F):
            b.step(acts_b[f])
        torch.cuda.synchronize()
        for k in ("obs_buf", "rew_buf", "reset_buf", "prev_error", "time_until_next_push", "trajectory", "episode_length_buf"):
            assert torch.equal(getattr(a, k), getattr(b, k)), f"cycle {cycle}: {k}"
        for k in ("trajectory", "k", "t", "t_final", "weights", "rng_ctr"):
            assert torch.equal(getattr(a.traj_gen, k), getattr(b.traj_gen, k)), f"cycle {cycle}: traj_gen.{k}"


def test_trajectory_env_error_paths():
    """Same failures as the reference class: the terrain curriculum reads self.commands (legged_robot_trajectory.py:508) ->
    AttributeError; rom / generator classes outside the fused set are refused loudly; a wrong num_observations is refused."""
    case = LC.build_case("traj_flat_allterms", 64)
    case.cfg.terrain.mesh_type, case.cfg.terrain.curriculum = "trimesh", True
    with pytest.raises(AttributeError, match="commands"):
        LC.make_fused(case)
    case = LC.build_case("traj_flat_allterms", 64)
    case.cfg.rom.cls = "Unicycle"
    with pytest.raises(NotImplementedError):
        LC.make_fused(case)
    case = LC.build_case("traj_flat_allterms", 64)
    case.cfg.trajectory_generator.cls = "SpiralTrajectoryGenerator"        # not a class of rom_dynamics.py
    with pytest.raises(NotImplementedError):
        LC.make_fused(case)
    case = LC.build_case("traj_flat_allterms", 64)
    case.cfg.env.num_observations = 48
    with pytest.raises(ValueError, match="num_observations"):
        LC.make_fused(case)
    env = LC.make_fused(LC.build_case("traj_flat_allterms", 64))
    assert not hasattr(env, "commands") and env.trajectory.shape == (64, 10, 2) and env.obs_buf.shape == (64, 65)
    with pytest.raises(RuntimeError):
        env.step(torch.zeros(64, 12))          # CPU tensor: no fallback


def test_trajectory_env_curriculum_rewrites_the_kernel_parameters():
    """update_command_curriculum (legged_robot_trajectory.py:519-555) is a host-side rewrite of the numbers the kernels read: an env
    moved to curriculum stage 1 must equal, bit for bit, an env built from the stage-1 values directly; the stage advances when
    common_step_counter hits curriculum_steps[stage] (:414-417)."""
    from legged_gym_dev_b200.configs import Cfg
    ca, cb = LC.build_case("traj_flat_allterms", 512), LC.build_case("traj_flat_allterms", 512)
    keys = [k for k, v in vars(ca.cfg.rewards.scales).items() if v != 0]
    ca.cfg.curriculum = Cfg(use_curriculum=True, curriculum_steps=[10 ** 9, 10 ** 9], push=Cfg(magnitude=[0.1, 0.5], time=[3, 2]),
                            max_rom_distance=[0.5, 0.8], zero_rom_distance_likelihood=[1.0, 1.0], rom=Cfg(z=[1, 1], v=[0.5, 0.75]),
                            trajectory_generator=Cfg(t_low=[3, 2], t_high=[3, 2]), sigma=Cfg(tracking_rom=[1.0, 0.8]),
                            rewards=Cfg(**{k: [1.0, 0.8] for k in keys}))
    b = cb.cfg
    b.rom.v_min, b.rom.v_max = [v * 0.75 for v in b.rom.v_min], [v * 0.75 for v in b.rom.v_max]
    b.trajectory_generator.t_low, b.trajectory_generator.t_high = b.trajectory_generator.t_low * 2, b.trajectory_generator.t_high * 2
    b.rewards.tracking_sigma = b.rewards.tracking_sigma * 0.8
    b.domain_rand.max_rom_dist = [v * 0.8 for v in b.domain_rand.max_rom_dist]
    for k in keys:
        setattr(b.rewards.scales, k, getattr(b.rewards.scales, k) * 0.8)
    ea, eb = LC.make_fused(ca), LC.make_fused(cb)
    assert ea.curriculum_state == 0 and abs(ea._pod.tracking_sigma - 0.25) < 1e-7
    ea.curriculum_state = 1
    ea.update_command_curriculum()
    assert abs(ea._pod.tracking_sigma - 0.2) < 1e-7
    ids = torch.arange(512, device="cuda")
    # the construction-time generator draw used the stage-0 bounds in A: give both the same ramp end points
    ea.traj_gen.ramp_v_end.copy_(eb.traj_gen.ramp_v_end)
    ea.reset_traj(ids)
    eb.reset_traj(ids)
    for s in range(10):
        a = ca.tape.actions[s % 8].cuda()
        ea.step(a)
        eb.step(a)
        for name in ("obs_buf", "rew_buf", "reset_buf", "prev_error", "trajectory"):
            assert torch.equal(getattr(ea, name), getattr(eb, name)), f"step {s}: {name}"
        assert torch.equal(ea.traj_gen.trajectory, eb.traj_gen.trajectory) and torch.equal(ea._sums, eb._sums)
    # the trigger itself
    cc = LC.build_case("traj_flat_allterms", 64)
    cc.cfg.curriculum = ca.cfg.curriculum.clone()
    cc.cfg.curriculum.curriculum_steps = [3, 10 ** 9]
    ec = LC.make_fused(cc)
    ec.reset_traj(torch.arange(64, device="cuda"))
    for s in range(4):
        ec.step(cc.tape.actions[s].cuda())
        assert ec.curriculum_state == (1 if s >= 2 else 0)
    assert abs(ec._pod.tracking_sigma - 0.2) < 1e-7
