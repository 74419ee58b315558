"""CPU-side pinning of the oracle (runs without a GPU):
  * Philox known-answer vectors (Random123 distribution),
  * restated isaacgym.torch_utils vs scipy rotations,
  * the oracle port vs the UNMODIFIED reference, step for step (only where /root/reference exists),
  * the C-ABI library loads and exports every symbol include/b200gym.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import legged_case as LC
from oracle import philox
from oracle.compare import assert_close, assert_exact

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = tuple(int(x) for x in philox.philox4x32(*c, *k))
        assert got == want


def test_philox_uniform_range_and_columns():
    u = philox.uniform01(7, np.arange(1000), 3, philox.SITE_OBS_NOISE, 48)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.01
    # column j is independent of how many columns are requested
    assert np.array_equal(u[:, :5], philox.uniform01(7, np.arange(1000), 3, philox.SITE_OBS_NOISE, 5))
    r = philox.randint(7, np.arange(1000), 3, philox.SITE_TERRAIN, 1, 10)
    assert r.min() >= 0 and r.max() <= 9


def test_restated_quaternion_math_against_scipy():
    from scipy.spatial.transform import Rotation
    from oracle.isaacgym_restated import quat_apply, quat_rotate_inverse
    g = torch.Generator().manual_seed(0)
    q = torch.randn(200, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    v = torch.randn(200, 3, generator=g, dtype=torch.float64)
    R = Rotation.from_quat(q.numpy())
    assert np.allclose(quat_apply(q, v).numpy(), R.apply(v.numpy()), atol=1e-12)
    assert np.allclose(quat_rotate_inverse(q, v).numpy(), R.inv().apply(v.numpy()), atol=1e-12)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200gym.h")).read()
    names = set(re.findall(r"\b(b200gym_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 7
    from legged_gym_dev_b200 import _lib
    L = _lib.lib()                       # also checks struct sizes against the binding
    for n in names:
        assert hasattr(L, n), f"libb200gym.so does not export {n}"
    assert L.b200gym_version() == 100


def test_cfg_tables_match_reference_objects():
    pytest.importorskip("torch")
    if not os.path.isdir("/root/reference/legged_gym"):
        pytest.skip("/root/reference not present")
    import dataclasses
    from oracle import ref_harness as H
    from legged_gym_dev_b200 import configs, synthetic as S
    from legged_gym_dev_b200.params import flatten_legged_cfg
    ref = H.import_reference()
    for mine, theirs in ((configs.anymal_c_flat_cfg(), ref.envs.AnymalCFlatCfg()),
                         (configs.anymal_c_rough_cfg(), ref.envs.AnymalCRoughCfg())):
        theirs.curriculum.use_curriculum = False
        theirs.domain_rand.max_push_vel = 1.0
        a = dataclasses.asdict(flatten_legged_cfg(theirs, 0.005, S.DOF_NAMES, terrain_rows=9, terrain_cols=9))
        b = dataclasses.asdict(flatten_legged_cfg(mine, 0.005, S.DOF_NAMES, terrain_rows=9, terrain_cols=9))
        assert a == b


@pytest.mark.reference
@pytest.mark.parametrize("name", ["flat_pd_upstream", "flat_lstm_shipped", "flat_allterms_v", "flat_heading_nonoise",
                                  "rough_lstm_allterms", "rough_pd_shipped", "flat_cmd_curriculum"])
def test_port_tracks_unmodified_reference(name):
    """The travelling restatement vs the reference's own code, every step, same Philox stream."""
    from oracle import ref_harness as H
    N, steps = 128, 24
    case = LC.build_case(name, N)
    task, rs, cr, lstm, over = LC.CASES[name]
    env = H.make_reference_anymal(task, N, case.tape, seed=case.seed, reward_scales=rs, command_ranges=cr,
                                  use_actuator_network=lstm,
                                  heightfield=case.terrain["height_samples"] if case.rough else None,
                                  terrain_origins=case.terrain["terrain_origins"] if case.rough else None,
                                  episode_lengths=case.ep, overrides=over)
    if case.rough:
        env.terrain_levels[:] = case.terrain["terrain_levels"]
        env.terrain_types[:] = case.terrain["terrain_types"]
        env.env_origins[:] = case.terrain["env_origins"]
    port, phys = LC.make_port(case)
    resets = 0
    for s in range(steps):
        a = case.tape.actions[s % case.tape.frames] * (150.0 if s == 3 else 1.0)
        o1, _, r1, d1, x1 = env.step(a.clone())
        o2, _, r2, d2, x2 = port.step(a.clone(), phys)
        resets += int(d1.sum())
        tag = f"{name} step {s}: "
        assert_exact(d2, d1, tag + "reset")
        assert_exact(port.time_out_buf, env.time_out_buf, tag + "time_out")
        assert_exact(port.episode_length_buf, env.episode_length_buf, tag + "ep_len")
        assert_exact(port.last_contacts, env.last_contacts, tag + "last_contacts")
        assert_close(o2, o1, 1.0, tag + "obs")
        assert_close(r2, r1, 1.0, tag + "rew")
        assert_close(port.torques, env.torques, 80.0, tag + "torques")
        assert_close(port.commands, env.commands, 1.0, tag + "commands")
        assert_close(port.root_states, env.root_states, 1.0, tag + "root")
        assert_close(port.dof_state, env.dof_state, 1.0, tag + "dof")
        assert_close(port.feet_air_time, env.feet_air_time, 1.0, tag + "feet_air_time")
        for k in env.episode_sums:
            assert_close(port.episode_sums[k], env.episode_sums[k], 1.0, tag + "sum_" + k)
        if "episode" in x1:
            assert list(x1["episode"]) == list(x2["episode"])
            for k in x1["episode"]:
                assert_close(x2["episode"][k], x1["episode"][k], 1.0, tag + "extras " + k)
        if case.rough:
            assert_exact(port.measured_heights, env.measured_heights, tag + "heights")
            assert_exact(port.terrain_levels, env.terrain_levels, tag + "levels")
        if lstm:
            assert_close(port.sea_hidden_state, env.sea_hidden_state, 1.0, tag + "h")
            assert_close(port.sea_cell_state, env.sea_cell_state, 1.0, tag + "c")
    assert resets > 0


@pytest.mark.reference
@pytest.mark.parametrize("name", ["flat_pd_upstream", "rough_lstm_allterms"])
def test_port_external_reset_tracks_unmodified_reference(name):
    """LeggedRobot.reset_idx / BaseTask.reset called from outside step() (legged_robot.py:147-187, base_task.py:111-119): the
    port's external reset against the reference's own methods — state right after the reset and over the steps that follow."""
    from oracle import ref_harness as H
    N = 96
    case = LC.build_case(name, N)
    task, rs, cr, lstm, over = LC.CASES[name]
    env = H.make_reference_anymal(task, N, case.tape, seed=case.seed, reward_scales=rs, command_ranges=cr, use_actuator_network=lstm,
                                  heightfield=case.terrain["height_samples"] if case.rough else None,
                                  terrain_origins=case.terrain["terrain_origins"] if case.rough else None,
                                  episode_lengths=case.ep, overrides=over)
    if case.rough:
        env.terrain_levels[:] = case.terrain["terrain_levels"]
        env.terrain_types[:] = case.terrain["terrain_types"]
        env.env_origins[:] = case.terrain["env_origins"]
    port, phys = LC.make_port(case)

    def check(tag):
        assert_exact(port.reset_buf, env.reset_buf.bool(), tag + "reset")
        assert_exact(port.time_out_buf, env.time_out_buf.bool(), tag + "time_out")
        assert_exact(port.episode_length_buf, env.episode_length_buf, tag + "ep_len")
        for k in ("commands", "root_states", "dof_state", "feet_air_time", "last_actions", "last_dof_vel", "obs_buf", "rew_buf"):
            assert_close(getattr(port, k), getattr(env, k), 1.0, tag + k)
        for k in env.episode_sums:
            assert_close(port.episode_sums[k], env.episode_sums[k], 1.0, tag + "sum_" + k)
        assert list(env.extras["episode"]) == list(port.extras["episode"])
        for k in env.extras["episode"]:
            assert_close(port.extras["episode"][k], env.extras["episode"][k], 1.0, tag + "extras " + k)
        if case.rough:
            assert_exact(port.terrain_levels, env.terrain_levels, tag + "levels")
            assert_close(port.env_origins, env.env_origins, 1.0, tag + "origins")
        if lstm:
            assert_close(port.sea_hidden_state, env.sea_hidden_state, 1.0, tag + "h")

    for s in range(5):
        a = case.tape.actions[s % case.tape.frames]
        env.step(a.clone())
        port.step(a.clone(), phys)
    ids = torch.arange(0, N, 3)
    env.reset_idx(ids)
    port.reset_idx(ids)
    check("partial reset: ")
    assert bool(port.reset_buf[ids].all())
    for s in range(5, 8):
        a = case.tape.actions[s % case.tape.frames]
        env.step(a.clone())
        port.step(a.clone(), phys)
        check(f"step {s} after the partial reset: ")
    o1, _ = env.reset()
    o2, _ = port.reset(phys)
    assert_close(o2, o1, 1.0, "reset() obs")
    check("reset(): ")
    assert not bool(env.time_out_buf.any()), "BaseTask.reset leaves no time-out flag behind"
    env.reset_idx(torch.arange(0))
    port.reset_idx(torch.arange(0))
    check("empty reset: ")


@pytest.mark.reference
@pytest.mark.parametrize("over", [{}, dict(prob_stationary=0.05, t_low=0.2, t_high=0.5), dict(randomize_rom_distance=False)])
def test_rom_port_tracks_unmodified_reference(over):
    """oracle/port_rom.py vs the reference's CustomSim/TrajectoryGenerator/DoubleSingleTracking, every step."""
    from oracle import ref_harness as H
    from oracle.port_rom import RomPort, rom_params
    N, steps = 96, 260
    env, policy, cfg = H.make_reference_custom_sim(N, seed=3, **over)
    port = RomPort(rom_params(N, seed=3, **over))
    assert_exact(port.ramp_v_end, env.traj_gen.ramp_v_end, "ramp_v_end at construction")
    env.reset()
    o1 = env.get_observations()
    o2, _ = port.reset()
    for s in range(steps):
        if s == 130:
            ids = torch.arange(0, N, 3)
            env.reset_idx(ids)
            o1 = env.get_observations()
            o2, _ = port.reset_idx(ids)
        a1, a2 = policy(o1), port.policy(o2)
        o1, _, _, d1, _ = env.step(a1)
        o2, d2 = port.step(a2)
        tg, tag = env.traj_gen, f"rom {over} step {s}: "
        assert_exact(port.t, tg.t, tag + "t")
        assert_exact(port.k, tg.k, tag + "k")
        assert_exact(port.stationary, tg.stationary_inds, tag + "stationary")
        assert_close(a2, a1, 1.0, tag + "action")
        assert_close(o2, o1, 1.0, tag + "obs")
        assert_close(port.traj, tg.trajectory, 1.0, tag + "trajectory")
        assert_close(port.v_traj, tg.v_trajectory, 1.0, tag + "v_trajectory")
        assert_close(port.v, tg.v, 1.0, tag + "v")
        assert_close(port.weights, tg.weights, 1.0, tag + "weights")
        assert_close(port.t_final, tg.t_final, 1.0, tag + "t_final")
        assert_close(port.trajectory, env.trajectory, 1.0, tag + "CustomSim.trajectory")
    assert np.array_equal(port.ctr, env._shim.ctr)


@pytest.mark.reference
@pytest.mark.parametrize("name", ["traj_flat_allterms", "traj_flat_lstm_shipped", "traj_flat_nonoise_norand", "traj_rough_lstm_allterms",
                                  "traj_flat_zero_gen", "traj_flat_square_gen", "traj_flat_circle_gen"])
def test_trajectory_port_tracks_unmodified_reference(name):
    """oracle/port_legged_traj.py vs the reference's own AnymalTrajectory / LeggedRobotTrajectory
    (legged_robot_trajectory.py), every step, same Philox stream (env draws keyed by step, generator draws by event counter)."""
    from oracle import ref_harness as H
    N, steps = 96, 40
    case = LC.build_case(name, N)
    task, rs, _, lstm, over = LC.CASES[name]
    env = H.make_reference_anymal_trajectory(task, N, case.tape, seed=case.seed, reward_scales=rs, use_actuator_network=lstm,
                                             heightfield=case.terrain["height_samples"] if case.rough else None,
                                             terrain_origins=case.terrain["terrain_origins"] if case.rough else None,
                                             episode_lengths=case.ep, time_until_next_push=case.tpush, overrides=over)
    if case.rough:
        env.env_origins[:] = case.terrain["env_origins"]
    port, phys = LC.make_port(case)
    assert_exact(port.gen.ramp_v_end, env.traj_gen.ramp_v_end, "ramp_v_end at construction")
    assert_close(port.noise_scale_vec, env.noise_scale_vec, 1.0, "noise_scale_vec")
    assert_close(port.reward_weighting, env.reward_weighting, 1.0, "reward_weighting")
    assert sorted(port.episode_sums) == sorted(env.episode_sums)
    # a never-reset generator divides 0/0 (rom_dynamics.py:552 with t_final == ramp_t_start == 0): start, as reset() does, from
    # generators reset at the robots' positions
    env.reset_traj(torch.arange(N))
    port.gen.reset_traj(torch.arange(N), port.proj_z())
    resets = pushes = 0
    for s in range(steps):
        a = case.tape.actions[s % case.tape.frames] * (150.0 if s == 3 else 1.0)
        t_before = env.time_until_next_push.clone()
        o1, _, r1, d1, x1 = env.step(a.clone())
        o2, _, r2, d2, x2 = port.step(a.clone(), phys)
        resets += int(d1.sum())
        pushes += int((t_before - 0.02 <= 0).sum())
        tag = f"{name} step {s}: "
        tg, g = env.traj_gen, port.gen
        assert_exact(d2, d1, tag + "reset")
        assert_exact(port.time_out_buf, env.time_out_buf, tag + "time_out")
        assert_exact(port.episode_length_buf, env.episode_length_buf, tag + "ep_len")
        assert_exact(port.last_contacts, env.last_contacts, tag + "last_contacts")
        assert_exact(g.k, tg.k, tag + "gen k")
        assert_exact(g.t, tg.t, tag + "gen t")
        assert_exact(g.stationary, tg.stationary_inds, tag + "gen stationary")
        assert_close(o2, o1, 1.0, tag + "obs")
        assert_close(r2, r1, 1.0, tag + "rew")
        assert_close(port.torques, env.torques, 80.0, tag + "torques")
        assert_close(port.root_states, env.root_states, 1.0, tag + "root")
        assert_close(port.dof_state, env.dof_state, 1.0, tag + "dof")
        assert_close(port.feet_air_time, env.feet_air_time, 1.0, tag + "feet_air_time")
        assert_close(port.trajectory, env.trajectory, 1.0, tag + "trajectory")
        assert_close(port.prev_error, env.prev_error, 1.0, tag + "prev_error")
        assert_close(port.time_until_next_push, env.time_until_next_push, 1.0, tag + "time_until_next_push")
        assert_close(g.traj, tg.trajectory, 1.0, tag + "gen trajectory")
        assert_close(g.v_traj, tg.v_trajectory, 1.0, tag + "gen v_trajectory")
        assert_close(g.weights, tg.weights, 1.0, tag + "gen weights")
        assert_close(g.t_final, tg.t_final, 1.0, tag + "gen t_final")
        assert_close(g.v, tg.v, 1.0, tag + "gen v")
        if hasattr(tg, "center"):
            assert_close(g.center, tg.center, 1.0, tag + "gen center")
        for k in env.episode_sums:
            assert_close(port.episode_sums[k], env.episode_sums[k], 1.0, tag + "sum_" + k)
        if "episode" in x1:
            assert list(x1["episode"]) == list(x2["episode"])
            for k in x1["episode"]:
                assert_close(x2["episode"][k], x1["episode"][k], 1.0, tag + "extras " + k)
        if case.rough:
            assert_exact(port.measured_heights, env.measured_heights, tag + "heights")
        if lstm:
            assert_close(port.sea_hidden_state, env.sea_hidden_state, 1.0, tag + "h")
            assert_close(port.sea_cell_state, env.sea_cell_state, 1.0, tag + "c")
    assert np.array_equal(port.gen.ctr, env._traj_shim.ctr)
    assert resets > N // 2 and pushes > 0


@pytest.mark.reference
@pytest.mark.parametrize("name", ["traj_flat_allterms", "traj_rough_lstm_allterms"])
def test_trajectory_port_external_reset_tracks_unmodified_reference(name):
    """LeggedRobotTrajectory.reset_idx / BaseTask.reset called from outside step() (legged_robot_trajectory.py:204-246, base_task.py:111-119):
    the trajectory port's external reset against the reference's own methods, generator state included."""
    from oracle import ref_harness as H
    N = 96
    case = LC.build_case(name, N)
    task, rs, _, lstm, over = LC.CASES[name]
    env = H.make_reference_anymal_trajectory(task, N, case.tape, seed=case.seed, reward_scales=rs, use_actuator_network=lstm,
                                             heightfield=case.terrain["height_samples"] if case.rough else None,
                                             terrain_origins=case.terrain["terrain_origins"] if case.rough else None,
                                             episode_lengths=case.ep, time_until_next_push=case.tpush, overrides=over)
    if case.rough:
        env.env_origins[:] = case.terrain["env_origins"]
    port, phys = LC.make_port(case)

    def check(tag):
        tg, g = env.traj_gen, port.gen
        assert_exact(port.reset_buf, env.reset_buf.bool(), tag + "reset")
        assert_exact(port.time_out_buf, env.time_out_buf.bool(), tag + "time_out")
        assert_exact(port.episode_length_buf, env.episode_length_buf, tag + "ep_len")
        assert_exact(g.k, tg.k, tag + "gen k")
        for k in ("root_states", "dof_state", "feet_air_time", "last_actions", "last_dof_vel", "obs_buf", "rew_buf", "trajectory", "prev_error"):
            assert_close(getattr(port, k), getattr(env, k), 1.0, tag + k)
        assert_close(g.traj, tg.trajectory, 1.0, tag + "gen trajectory")
        assert_close(g.v, tg.v, 1.0, tag + "gen v")
        for k in env.episode_sums:
            assert_close(port.episode_sums[k], env.episode_sums[k], 1.0, tag + "sum_" + k)
        if "episode" in env.extras:
            for k in env.extras["episode"]:
                assert_close(port.extras["episode"][k], env.extras["episode"][k], 1.0, tag + "extras " + k)

    o1, _ = env.reset()
    o2, _ = port.reset(phys)
    assert_close(o2, o1, 1.0, "reset() obs")
    check("reset(): ")
    for s in range(4):
        a = case.tape.actions[s % case.tape.frames]
        env.step(a.clone())
        port.step(a.clone(), phys)
        check(f"step {s} after reset(): ")
    ids = torch.arange(0, N, 3)
    env.reset_idx(ids)
    port.reset_idx(ids)
    check("partial reset: ")
    for s in range(4, 8):
        a = case.tape.actions[s % case.tape.frames]
        env.step(a.clone())
        port.step(a.clone(), phys)
        check(f"step {s} after the partial reset: ")
    assert np.array_equal(port.gen.ctr, env._traj_shim.ctr)


def test_trajectory_cfg_tables_match_reference_objects():
    """configs.py's anymal trajectory tables == the reference's cfg classes (+ the documented completion) after flattening."""
    if not os.path.isdir("/root/reference/legged_gym"):
        pytest.skip("/root/reference not present")
    import dataclasses
    from oracle import ref_harness as H
    from legged_gym_dev_b200 import configs, synthetic as S
    from legged_gym_dev_b200.params import flatten_legged_cfg
    ref = H.import_reference()
    for mine, theirs in ((configs.anymal_c_flat_trajectory_cfg(), ref.envs.AnymalCFlatTrajectoryCfg()),
                         (configs.anymal_c_rough_trajectory_cfg(), ref.envs.AnymalCRoughTrajectoryCfg())):
        H.complete_trajectory_cfg(theirs)
        theirs.env.num_observations = mine.env.num_observations      # the shipped 240 matches no rom (legged_robot_trajectory_config.py:36)
        a = dataclasses.asdict(flatten_legged_cfg(theirs, 0.005, S.DOF_NAMES, terrain_rows=9, terrain_cols=9, trajectory=True))
        b = dataclasses.asdict(flatten_legged_cfg(mine, 0.005, S.DOF_NAMES, terrain_rows=9, terrain_cols=9, trajectory=True))
        assert a == b
        for sec in ("rom", "trajectory_generator"):
            ta, tb = getattr(theirs, sec), getattr(mine, sec)
            for k in ("cls", "dt", "z_min", "z_max", "v_min", "v_max") if sec == "rom" else \
                    ("cls", "t_samp_cls", "weight_samp_cls", "N", "dN", "t_low", "t_high", "freq_low", "freq_high", "seed", "prob_stationary"):
                assert getattr(ta, k) == getattr(tb, k), (sec, k)
        for k in ("randomize_rom_distance", "max_rom_dist", "zero_rom_distance_likelihood", "time_between_pushes", "max_push_vel_xy"):
            assert getattr(theirs.domain_rand, k) == getattr(mine.domain_rand, k), k


def test_trajectory_env_rejects_what_the_reference_class_cannot_run():
    """LeggedRobotTrajectory has no commands (legged_robot_trajectory.py:621-622): the terms that need them raise AttributeError in the
    reference (missing _reward_* / missing self.commands) and in the flattening; tracking_rom / differential_error need the class."""
    from legged_gym_dev_b200 import configs, synthetic as S
    from legged_gym_dev_b200.params import flatten_legged_cfg
    for term in ("tracking_lin_vel", "tracking_ang_vel", "stand_still"):
        cfg = configs.anymal_c_flat_trajectory_cfg()
        setattr(cfg.rewards.scales, term, 1.0)
        with pytest.raises(AttributeError):
            flatten_legged_cfg(cfg, 0.005, S.DOF_NAMES, trajectory=True)
    for term in ("tracking_rom", "differential_error"):
        cfg = configs.anymal_c_flat_cfg()
        setattr(cfg.rewards.scales, term, 1.0)
        with pytest.raises(AttributeError):
            flatten_legged_cfg(cfg, 0.005, S.DOF_NAMES)
    p = flatten_legged_cfg(configs.anymal_c_flat_trajectory_cfg(), 0.005, S.DOF_NAMES, trajectory=True)
    assert p.traj_mode and p.time_between_pushes == [0.5, 10.0] and p.max_push_vel == 1.0
