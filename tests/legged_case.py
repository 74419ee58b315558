"""Builders shared by the Group-R parity tests: one *case* = task + reward table + control path."""
import os
from types import SimpleNamespace

import numpy as np
import torch

from legged_gym_dev_b200 import configs, synthetic as S
from legged_gym_dev_b200.params import flatten_legged_cfg

ALL_REWARD_SCALES = dict(
    action_rate=-0.01, ang_vel_xy=-0.05, base_height=-1.0, collision=-1.0, dof_acc=-2.5e-7, dof_pos_limits=-10.0,
    dof_vel=-1e-4, dof_vel_limits=-0.5, feet_air_time=1.0, feet_contact_forces=-0.01, lin_vel_z=-2.0, orientation=-5.0,
    stand_still=-0.1, stumble=-0.3, termination=-3.0, torque_limits=-0.02, torques=-1e-5, tracking_ang_vel=0.5,
    tracking_lin_vel=1.0)

# name -> (task, reward table, command table, use_actuator_network, extra overrides)
CASES = {
    "flat_pd_upstream": ("anymal_c_flat", configs.UPSTREAM_REWARD_SCALES, configs.UPSTREAM_COMMAND_RANGES, False, {}),
    "flat_lstm_shipped": ("anymal_c_flat", None, None, True, {}),
    "flat_allterms_v": ("anymal_c_flat", ALL_REWARD_SCALES, configs.UPSTREAM_COMMAND_RANGES, False,
                        {"rewards.only_positive_rewards": False, "control.control_type": "V"}),
    "flat_heading_nonoise": ("anymal_c_flat", configs.UPSTREAM_REWARD_SCALES, configs.UPSTREAM_COMMAND_RANGES, False,
                             {"commands.heading_command": True, "noise.add_noise": False,
                              "domain_rand.push_interval_s": 0.1}),
    "rough_lstm_allterms": ("anymal_c_rough", ALL_REWARD_SCALES, configs.UPSTREAM_COMMAND_RANGES, True,
                            {"rewards.only_positive_rewards": False, "domain_rand.push_interval_s": 0.1}),
    "rough_pd_shipped": ("anymal_c_rough", None, None, False, {}),
    # the fork's command curriculum of the base env (legged_robot.py:360-363, 488-506): advances at steps 5 and 9; pushes off (the
    # reference's _push_robots cannot run with it, :459 negates a list)
    "flat_cmd_curriculum": ("anymal_c_flat", configs.UPSTREAM_REWARD_SCALES, configs.UPSTREAM_COMMAND_RANGES, False,
                            {"curriculum.use_curriculum": True, "curriculum.curriculum_steps": [5, 9], "domain_rand.push_robots": False,
                             "domain_rand.max_push_vel": [1.0, 1.0], "commands.resampling_time": 0.1}),
    # AnymalTrajectory (legged_robot_trajectory.py); frequent per-env pushes so the timers fire in a short test
    "traj_flat_allterms": ("anymal_c_flat_trajectory", configs.TRAJECTORY_ALL_REWARD_SCALES, None, False,
                           {"domain_rand.time_between_pushes": [0.05, 0.3]}),
    "traj_flat_lstm_shipped": ("anymal_c_flat_trajectory", None, None, True, {}),
    "traj_flat_nonoise_norand": ("anymal_c_flat_trajectory", configs.TRAJECTORY_ALL_REWARD_SCALES, None, False,
                                 {"noise.add_noise": False, "domain_rand.randomize_rom_distance": False,
                                  "trajectory_generator.weight_samp_cls": "UniformWeightSamplerNoRamp",
                                  "trajectory_generator.t_low": 0.1, "trajectory_generator.t_high": 0.3}),
    # SURVEY 8f row 4 (generators): the deterministic TrajectoryGenerator subclasses (rom_dynamics.py:618-698)
    "traj_flat_zero_gen": ("anymal_c_flat_trajectory", configs.TRAJECTORY_ALL_REWARD_SCALES, None, False,
                           {"trajectory_generator.cls": "ZeroTrajectoryGenerator"}),
    "traj_flat_square_gen": ("anymal_c_flat_trajectory", configs.TRAJECTORY_ALL_REWARD_SCALES, None, False,
                             {"trajectory_generator.cls": "SquareTrajectoryGenerator", "domain_rand.time_between_pushes": [0.05, 0.3]}),
    "traj_flat_circle_gen": ("anymal_c_flat_trajectory", configs.TRAJECTORY_ALL_REWARD_SCALES, None, False,
                             {"trajectory_generator.cls": "CircleTrajectoryGenerator"}),
    "traj_rough_lstm_allterms": ("anymal_c_rough_trajectory", configs.TRAJECTORY_ALL_REWARD_SCALES, None, True,
                                 {"domain_rand.time_between_pushes": [0.05, 0.3]}),
}
TRAJ_TASKS = ("anymal_c_flat_trajectory", "anymal_c_rough_trajectory")


def dof_limits():
    return S.anymal_dof_limits()


def apply_overrides(cfg, reward_scales, command_ranges, lstm, overrides):
    if reward_scales is not None:
        for k, v in reward_scales.items():
            setattr(cfg.rewards.scales, k, v)
    if command_ranges is not None:
        for k, v in command_ranges.items():
            setattr(cfg.commands.ranges, k, list(v))
    cfg.control.use_actuator_network = lstm
    for path, v in overrides.items():
        obj = cfg
        parts = path.split(".")
        for q in parts[:-1]:
            obj = getattr(obj, q)
        setattr(obj, parts[-1], v)
    return cfg


def build_case(name, num_envs, frames=8, seed=5, tape_seed=1, base_contact_prob=0.02):
    """Returns cfg + all inputs (CPU tensors) for the case."""
    task, rs, cr, lstm, over = CASES[name]
    traj = task in TRAJ_TASKS
    rough = task in ("anymal_c_rough", "anymal_c_rough_trajectory")
    if traj:
        cfg = configs.anymal_c_rough_trajectory_cfg() if rough else configs.anymal_c_flat_trajectory_cfg()
    else:
        cfg = configs.anymal_c_rough_cfg() if rough else configs.anymal_c_flat_cfg()
    cfg = apply_overrides(cfg, rs, cr, lstm, over)
    cfg.env.num_envs = num_envs
    tape = S.make_state_tape(num_envs, frames=frames, seed=tape_seed, rough=rough, base_contact_prob=base_contact_prob)
    ep = S.make_episode_lengths(num_envs, seed=tape_seed)
    tpush = None
    if traj:
        keep = ep[::8].clone()
        ep[:] = 1001          # most envs time out on the first step: their generators are reset next to the robots
        ep[::8] = keep
        g = torch.Generator().manual_seed(tape_seed + 7)
        tpush = 0.4 * torch.rand(num_envs, 1, generator=g)   # stands in for the construction-time draw (legged_robot_trajectory.py:85-88)
    terrain = None
    if rough:
        hf = S.make_heightfield(seed=tape_seed)
        to = S.make_terrain_origins(seed=tape_seed)
        g = torch.Generator().manual_seed(tape_seed + 99)
        levels = torch.randint(0, cfg.terrain.max_init_terrain_level + 1, (num_envs,), generator=g)
        types = torch.div(torch.arange(num_envs), (num_envs / cfg.terrain.num_cols), rounding_mode="floor").to(torch.long)
        terrain = dict(height_samples=hf, terrain_origins=to, terrain_levels=levels, terrain_types=types,
                       env_origins=to[levels, types].clone())
    if traj:   # keep the replayed robots near their env origins, where the generators start (tracking errors of O(0.3 m))
        if terrain is not None:
            o = terrain["env_origins"]
        else:
            cols = np.floor(np.sqrt(num_envs))
            rows = np.ceil(num_envs / cols)
            xx, yy = torch.meshgrid(torch.arange(rows), torch.arange(cols), indexing="ij")
            o = torch.stack([cfg.env.env_spacing * xx.flatten()[:num_envs], cfg.env.env_spacing * yy.flatten()[:num_envs]], dim=1)
        g = torch.Generator().manual_seed(tape_seed + 8)
        tape.root[:, :, 0:2] = o[None, :, :2].float() + 0.3 * torch.randn(tape.root.shape[0], num_envs, 2, generator=g)
    return SimpleNamespace(name=name, task=task, rough=rough, lstm=lstm, cfg=cfg, tape=tape, ep=ep, terrain=terrain,
                           seed=seed, num_envs=num_envs, limits=dof_limits(), traj=traj, tpush=tpush)


def make_params(case):
    lim, t = case.limits, case.terrain
    return flatten_legged_cfg(case.cfg, case.cfg.sim.dt, S.DOF_NAMES, num_envs=case.num_envs,
                              dof_pos_limits=lim["dof_pos_limits"].tolist(), dof_vel_limits=lim["dof_vel_limits"].tolist(),
                              torque_limits=lim["torque_limits"].tolist(),
                              terrain_rows=t["height_samples"].shape[0] if t else 0,
                              terrain_cols=t["height_samples"].shape[1] if t else 0, seed=case.seed, trajectory=case.traj)


def generator_cfg(cfg):
    """trajectory_generator / rom / domain_rand numbers of a trajectory cfg, as oracle.port_legged_traj wants them."""
    tg, rom, d = cfg.trajectory_generator, cfg.rom, cfg.domain_rand
    return dict(rom_dt=rom.dt, vel_max_rom=rom.v_max[0], N=tg.N, dN=tg.dN, t_low=tg.t_low, t_high=tg.t_high, freq_low=tg.freq_low,
                freq_high=tg.freq_high, prob_stationary=tg.prob_stationary, weight_sampler=tg.weight_samp_cls, seed=tg.seed,
                generator=tg.cls,
                randomize_rom_distance=d.randomize_rom_distance, max_rom_distance=d.max_rom_dist,
                zero_rom_dist_llh=d.zero_rom_distance_likelihood)


def make_port(case, rng="philox", env_id_offset=0):
    from oracle.port_legged import LeggedPort, TapePhysics
    p = make_params(case)
    if case.traj:
        p.traj_n, p.traj_horizon = 2, case.cfg.trajectory_generator.N
        rw = case.cfg.rewards.reward_weighting
        p.traj_weight = [float(rw.position), float(rw.position), 0.0, 0.0]      # SingleInt2D.get_weighting_vector
    t = case.terrain or {}
    w = None
    if case.lstm:
        import legged_gym_dev_b200
        path = os.path.join(os.path.dirname(legged_gym_dev_b200.__file__), "resources", "anydrive_v3_lstm.npz")
        w = dict(np.load(path))
    clone = lambda x: x.clone() if x is not None else None
    origins = clone(t.get("env_origins"))
    if origins is None:                                               # legged_robot.py:808-817 (grid of robots)
        N = case.num_envs
        cols = np.floor(np.sqrt(N))
        rows = np.ceil(N / cols)
        xx, yy = torch.meshgrid(torch.arange(rows), torch.arange(cols), indexing="ij")
        origins = torch.zeros(N, 3)
        origins[:, 0] = case.cfg.env.env_spacing * xx.flatten()[:N]
        origins[:, 1] = case.cfg.env.env_spacing * yy.flatten()[:N]
    make = LeggedPort
    if case.traj:
        from oracle.port_legged_traj import LeggedTrajPort
        make = lambda p_, r, d, c, **k: LeggedTrajPort(p_, generator_cfg(case.cfg), r, d, c, case.tpush, **k)
    port = make(p, case.tape.root[0].clone(), case.tape.dof[0, 0].clone(), case.tape.contact[0].clone(),
                      env_origins=origins, height_samples=t.get("height_samples"),
                      terrain_levels=clone(t.get("terrain_levels")), terrain_types=clone(t.get("terrain_types")),
                      terrain_origins=clone(t.get("terrain_origins")), lstm=w, episode_length_buf=case.ep, rng=rng,
                      env_id_offset=env_id_offset)
    return port, TapePhysics(case.tape)


def make_fused(case, device="cuda", copy=True, env_id_offset=0):
    from legged_gym_dev_b200.legged_robot import Anymal
    from legged_gym_dev_b200.physics import ReplayPhysics
    if case.traj:
        from legged_gym_dev_b200.legged_robot_trajectory import AnymalTrajectory as Anymal   # noqa: F811
    phys = ReplayPhysics(case.tape, device=device, copy=copy)
    lim = case.limits
    asset = dict(dof_pos_limits=lim["dof_pos_limits"], dof_vel_limits=lim["dof_vel_limits"], torque_limits=lim["torque_limits"])
    env = Anymal(case.cfg, SimpleNamespace(dt=case.cfg.sim.dt), None, device, True, physics=phys, asset=asset,
                 seed=case.seed, terrain=case.terrain, env_id_offset=env_id_offset)
    env.episode_length_buf.copy_(case.ep.to(device))
    if case.traj:
        env.time_until_next_push.copy_(case.tpush.to(device))
    return env


# quantities compared after every step: name -> (getter(port), getter(fused), kind, scale)
def _gen_port(d, g):
    d.update(gen_trajectory=g.traj, gen_v_trajectory=g.v_traj, gen_t=g.t, gen_k=g.k, gen_t_final=g.t_final, gen_weights=g.weights,
             gen_stationary=g.stationary, gen_ramp_v_end=g.ramp_v_end, gen_sin_freq=g.sin_freq,
             gen_ctr=torch.from_numpy(g.ctr.astype("int64")))
    if hasattr(g, "center"):
        d["gen_center"] = g.center


def _gen_fused(d, g):
    d.update(gen_trajectory=g.trajectory, gen_v_trajectory=g.v_trajectory, gen_t=g.t, gen_k=g.k, gen_t_final=g.t_final,
             gen_weights=g.weights, gen_stationary=g.stationary_inds, gen_ramp_v_end=g.ramp_v_end, gen_sin_freq=g.sin_freq,
             gen_ctr=g.rng_ctr.long())
    if g.center is not None:
        d["gen_center"] = g.center


def snapshot_port(port):
    d = dict(obs=port.obs_buf, rew=port.rew_buf, reset=port.reset_buf, time_out=port.time_out_buf, torques=port.torques,
             ep_len=port.episode_length_buf, feet_air_time=port.feet_air_time,
             last_contacts=port.last_contacts, root=port.root_states, dof=port.dof_state, last_actions=port.last_actions,
             last_dof_vel=port.last_dof_vel, last_root_vel=port.last_root_vel, base_lin_vel=port.base_lin_vel,
             base_ang_vel=port.base_ang_vel, projected_gravity=port.projected_gravity)
    if hasattr(port, "gen"):
        d.update(trajectory=port.trajectory, prev_error=port.prev_error, time_until_next_push=port.time_until_next_push)
        _gen_port(d, port.gen)
    else:
        d["commands"] = port.commands
    for k, v in port.episode_sums.items():
        d["sum_" + k] = v
    if port.p.measure_heights:
        d["heights"] = port.measured_heights
    if port.p.terrain_curriculum:
        d["terrain_levels"] = port.terrain_levels
        d["env_origins"] = port.env_origins
    if port.p.use_actuator_network:
        d["lstm_h"], d["lstm_c"] = port.sea_hidden_state, port.sea_cell_state
    for k, v in port.extras.get("episode", {}).items():
        d["extras_" + k] = v
    return {k: (v.detach().clone() if torch.is_tensor(v) else torch.as_tensor(v)) for k, v in d.items()}


def snapshot_fused(env):
    d = dict(obs=env.obs_buf, rew=env.rew_buf, reset=env.reset_buf, time_out=env.time_out_buf, torques=env.torques,
             ep_len=env.episode_length_buf, feet_air_time=env.feet_air_time,
             last_contacts=env.last_contacts, root=env.root_states, dof=env.dof_state, last_actions=env.last_actions,
             last_dof_vel=env.last_dof_vel, last_root_vel=env.last_root_vel, base_lin_vel=env.base_lin_vel,
             base_ang_vel=env.base_ang_vel, projected_gravity=env.projected_gravity)
    if hasattr(env, "traj_gen"):
        d.update(trajectory=env.trajectory, prev_error=env.prev_error, time_until_next_push=env.time_until_next_push)
        _gen_fused(d, env.traj_gen)
    else:
        d["commands"] = env.commands
    for k, v in env.episode_sums.items():
        d["sum_" + k] = v
    if env.params.measure_heights:
        d["heights"] = env.measured_heights
    if env.params.terrain_curriculum:
        d["terrain_levels"] = env.terrain_levels
        d["env_origins"] = env.env_origins
    if env.params.use_actuator_network:
        d["lstm_h"], d["lstm_c"] = env.sea_hidden_state, env.sea_cell_state
    for k, v in env.extras["episode"].items():
        d["extras_" + k] = v
    return {k: v.detach().cpu().clone() for k, v in d.items()}


EXACT = {"reset", "time_out", "ep_len", "last_contacts", "terrain_levels", "gen_k", "gen_stationary", "gen_ctr"}
SCALES = {"torques": 80.0, "heights": 1.0}


def compare_snapshots(got, want, tag=""):
    """Bit-exact for flags / counters / indices; |a-b| <= 1e-5*(|b| + S) otherwise (S=80 for torques, else 1).
    Height cells are index work: compared exactly up to the (stated) vertical quantum."""
    from oracle.compare import assert_close, assert_exact
    worst = {}
    for k, w in want.items():
        if k not in got:
            raise AssertionError(f"{tag}: fused path has no `{k}`")
        g = got[k]
        if k in EXACT:
            assert_exact(g.to(w.dtype) if g.dtype != w.dtype else g, w, f"{tag}{k}")
        elif k == "obs" and "trajectory" in want:
            # columns 9..28 are (trajectory - root_xy) * scale (legged_robot_trajectory.py:277-283): a difference of two positions
            # of O(|origin|) metres.  The 1e-5 contract holds for the operands (trajectory is compared on its own below), so
            # the difference is compared at the operands' magnitude; every other column at S = 1.
            tw = want["trajectory"].shape[1] * want["trajectory"].shape[2]
            mag = float(want["trajectory"].abs().max()) + 1.0
            cols = torch.ones(w.shape[1], dtype=torch.bool)
            cols[9:9 + tw] = False
            worst[k] = assert_close(g[:, cols], w[:, cols].to(g.dtype), 1.0, f"{tag}{k}")
            worst[k + "_traj"] = assert_close(g[:, ~cols], w[:, ~cols].to(g.dtype), mag, f"{tag}{k}[trajectory block]")
        elif k.startswith("extras_"):
            # extras only change on steps with a reset; the port keeps stale values too
            worst[k] = assert_close(g, w, 1.0, f"{tag}{k}")
        else:
            worst[k] = assert_close(g, w.to(g.dtype), SCALES.get(k, 1.0), f"{tag}{k}")
    return worst
