"""CPU restatement (torch, fp32) of the reference's HopperTrajectory env step.  TEST INFRASTRUCTURE.

Follows legged_gym/envs/hopper/hopper_trajectory.py — step :102-133, post_physics_step :135-182, reset :286-296, _reset_dofs /
_reset_root_states / _push_robots :298-370, compute_observations :255-282, _reward_torque_limits / _reward_dof_acc /
_reward_unit_quat / _reward_raibert :470-505 — on top of what the class inherits from LeggedRobotTrajectory
(legged_gym/envs/base/legged_robot_trajectory.py: check_termination :194-202, reset_idx :204-246, reset_traj :248-253,
compute_reward :255-272, _post_physics_step_callback :405-417, the shared _reward_* terms :1000-1110).  The pieces pinned
earlier are reused as they are (oracle/port_hopper.py: torque law, observations, reset / push draws, Raibert term), the trajectory
generator is oracle/port_rom.py's.  Pinned as a whole, step by step, against the UNMODIFIED reference class by
tests/test_hopper_cpu.py::test_hopper_env_port_tracks_unmodified_reference (oracle/ref_harness.make_reference_hopper_trajectory).

Reference behaviours the parity depends on:
  * `base_quat` is a VIEW of root_states (legged_robot_trajectory.py:601; the `base_quat[:] = ...` of :124 copies it onto itself), but
    `base_ang_vel` / `base_lin_vel` are buffers refreshed in the sub-step loop (:126) / at the top of post_physics_step (:145): reset_idx
    rewrites root_states only, so the observation of a reset env carries its post-reset height and quaternion next to its pre-reset
    velocities, and the first torque evaluation of the next step uses the pre-reset angular velocity;
  * `time_until_next_push` is [N, 1] (legged_robot_trajectory.py:85-88) and HopperTrajectory does not flatten the mask (:149,152):
    `need_push.nonzero().flatten()` interleaves the env ids with the (all-zero) column ids, so env 0 is pushed whenever ANY env is;
  * `_reset_dofs` overwrites `actions[env_ids]` with `zero_action` (:314): the observation's action block and `last_actions` of a reset
    env hold (1, 0, 0, 0), not the policy's action;
  * `prev_error` is recomputed for ALL envs at the end of the step from the callback's (pre-reset) trajectory clone and the final
    root (:178);
  * `self.torques` is re-bound to the limit-clipped return value of `_compute_torques` (:112): the reward terms see clipped torques;
  * pushes rewrite all six root velocities (:362-367); the timer redraw multiplies its bounds by curriculum.push.time only with
    curriculum.use_curriculum (off here: the curriculum rewrites reward tables mid-run, legged_robot_trajectory.py:519-555).
"""
from types import SimpleNamespace

import numpy as np
import torch

from . import philox as P
from . import port_hopper as PH
from .isaacgym_restated import quat_rotate_inverse
from .port_legged_traj import generator_params
from .port_rom import RomPort

# reward terms the Hopper env can run, in the order _prepare_reward_function gives them (alphabetical, `termination` outside the loop)
TERMS = ("action_rate", "ang_vel_xy", "base_height", "collision", "differential_error", "dof_acc", "dof_pos_limits", "dof_vel",
         "dof_vel_limits", "feet_air_time", "feet_contact_forces", "lin_vel_z", "orientation", "raibert", "stumble", "torque_limits",
         "torques", "tracking_rom", "unit_quat", "termination")

# hopper_trajectory_config.py (values the env reads); `scales` = deep_tube_learning/configs/rl/hopper_single_int.yaml's table
DEFAULT = dict(
    dt=0.02, sim_dt=0.005, decimation=4, clip_actions=100.0, episode_length_s=20.0, control_type="orientation", action_scale=1.0,
    p_gains=[400.0, 15.0, 15.0, 15.0], d_gains=[40.0, 3.0, 3.0, 3.0], kd_spindown=[0.1, 0.1, 0.1], wheel_speed_limits=[600.0, 600.0, 600.0],
    torque_speed_bound_ratio=6.0, rot_actuator=[[-0.8165, 0.2511, 0.2511], [-0.0, -0.7643, 0.7643], [-0.5773, -0.5939, -0.5939]],
    torque_limits=[300.0, 1.5, 1.5, 1.5], dof_pos_limits=[[-0.03, 0.1], [-1e4, 1e4], [-1e4, 1e4], [-1e4, 1e4]], dof_vel_limits=[5.0, 600.0, 600.0, 600.0],
    num_bodies=5, foot_body=4, termination_bodies=[0, 1, 2, 3], penalised_bodies=[1, 2, 3],
    obs=dict(PH.OBS_CFG), trajectory_scale=[1.0, 1.0],
    reset=dict(PH.RESET_CFG, base_init_state=[0.0, 0.0, 0.3, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0], default_dof_pos=[0.0, 0.0, 0.0, 0.0],
               dof_vel_noise=([-0.1, -100.0, -100.0, -100.0], [0.1, 100.0, 100.0, 100.0]), max_push_vel=[0.25, 0.25, 0.25, 0.75, 0.75, 0.75]),
    push_robots=True, time_between_pushes=[0.5, 10.0],
    scales=dict(termination=-500.0, tracking_rom=6.0, ang_vel_xy=-0.01, orientation=-80.0, torques=-1e-6, dof_acc=-2.5e-8, unit_quat=-0.01,
                collision=-1.0, action_rate=-0.01, differential_error=10.0, raibert=-0.1),
    only_positive_rewards=False, tracking_sigma=0.25, soft_dof_vel_limit=1.0, soft_torque_limit=1.0, base_height_target=0.55,
    max_contact_force=100.0, reward_weighting=[1.0, 1.0], diff_neg_slope=-1.0, diff_pos_slope=-4.0,
    raibert=dict(Kp=-0.3, Kv=-0.9, K_ff=0.0, clip_pos=0.5, clip_vel=1.0, clip_ang=0.2), send_timeouts=True, seed=0,
    generator=dict(rom_dt=0.1, vel_max_rom=0.35, N=10, dN=1, t_low=1.0, t_high=2.0, freq_low=0.01, freq_high=2.0, prob_stationary=0.0001,
                   weight_sampler="UniformWeightSamplerNoRamp", seed=42, randomize_rom_distance=False, max_rom_distance=[1.0, 1.0],
                   zero_rom_dist_llh=0.25))


def hopper_env_params(num_envs, **over):
    d = {k: (dict(v) if isinstance(v, dict) else v) for k, v in DEFAULT.items()}
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(d.get(k), dict):
            d[k].update(v)
        else:
            d[k] = v
    unknown = set(d["scales"]) - set(TERMS)
    if unknown:
        raise AttributeError(f"'HopperTrajectory' object has no attribute '_reward_{sorted(unknown)[0]}'")
    d["num_envs"] = int(num_envs)
    d["max_episode_length_s"] = d["episode_length_s"]
    d["max_episode_length"] = float(np.ceil(d["episode_length_s"] / d["dt"]))
    return SimpleNamespace(**d)


def apply_params_to_cfg(cfg, hp):
    """Writes the flat parameter namespace `hp` into a HopperRoughTrajectoryCfg-shaped attribute tree: the reference's own config class
    (oracle/ref_harness.make_reference_hopper_trajectory) or the product's mirror of it (legged_gym_dev_b200.configs.hopper_flat_trajectory_cfg),
    so that both sides of a parity test are configured by the same assignments."""
    N = hp.num_envs
    cfg.env.num_envs, cfg.env.num_observations, cfg.env.episode_length_s = N, 14 + hp.generator["N"] * 2 + 4, hp.episode_length_s
    for name in TERMS:
        setattr(cfg.rewards.scales, name, hp.scales.get(name, 0.0))
    rw = cfg.rewards
    rw.only_positive_rewards, rw.tracking_sigma, rw.base_height_target = hp.only_positive_rewards, hp.tracking_sigma, hp.base_height_target
    rw.soft_dof_vel_limit, rw.soft_torque_limit, rw.max_contact_force = hp.soft_dof_vel_limit, hp.soft_torque_limit, hp.max_contact_force
    rw.reward_weighting = SimpleNamespace(position=hp.reward_weighting[0])
    rw.differential_error.neg_slope, rw.differential_error.pos_slope = hp.diff_neg_slope, hp.diff_pos_slope
    g = hp.raibert
    rw.raibert.Kp, rw.raibert.Kv, rw.raibert.Kff = g["Kp"], g["Kv"], g["K_ff"]
    rw.raibert.clip_pos, rw.raibert.clip_vel, rw.raibert.clip_ang = g["clip_pos"], g["clip_vel"], g["clip_ang"]
    cfg.control.control_type, cfg.control.action_scale, cfg.control.decimation = hp.control_type, hp.action_scale, hp.decimation
    cfg.control.zero_action = list(hp.reset["zero_action"])
    cfg.normalization.clip_actions, cfg.normalization.clip_observations = hp.clip_actions, hp.obs["clip_observations"]
    cfg.normalization.obs_scales.trajectory = list(hp.trajectory_scale)
    cfg.noise.add_noise = hp.obs["add_noise"]
    d = cfg.domain_rand
    d.push_robots, d.time_between_pushes, d.max_push_vel = hp.push_robots, list(hp.time_between_pushes), list(hp.reset["max_push_vel"])
    tgc = hp.generator
    d.randomize_rom_distance, d.max_rom_dist, d.zero_rom_distance_likelihood = tgc["randomize_rom_distance"], list(tgc["max_rom_distance"]), tgc["zero_rom_dist_llh"]
    cfg.curriculum.use_curriculum = False
    cfg.rom.dt, cfg.rom.v_min, cfg.rom.v_max = tgc["rom_dt"], [-tgc["vel_max_rom"]] * 2, [tgc["vel_max_rom"]] * 2
    tg = cfg.trajectory_generator
    tg.weight_samp_cls, tg.N, tg.dN, tg.DN = tgc["weight_sampler"], tgc["N"], tgc["dN"], tgc["dN"]
    tg.t_low, tg.t_high, tg.freq_low, tg.freq_high = tgc["t_low"], tgc["t_high"], tgc["freq_low"], tgc["freq_high"]
    tg.prob_stationary, tg.seed = tgc["prob_stationary"], tgc["seed"]
    isl, rc = cfg.init_state, hp.reset
    isl.pos, isl.rot = list(rc["base_init_state"][:3]), list(rc["base_init_state"][3:7])
    isl.lin_vel, isl.ang_vel = list(rc["base_init_state"][7:10]), list(rc["base_init_state"][10:13])
    isl.randomize_yaw = rc["randomize_yaw"]
    isl.default_dof_pos_noise_lower, isl.default_dof_pos_noise_upper = list(rc["dof_pos_noise"][0]), list(rc["dof_pos_noise"][1])
    isl.default_dof_vel_noise_lower, isl.default_dof_vel_noise_upper = list(rc["dof_vel_noise"][0]), list(rc["dof_vel_noise"][1])
    isl.default_root_pos_noise_lower, isl.default_root_pos_noise_upper = list(rc["root_pos_noise"][0]), list(rc["root_pos_noise"][1])
    isl.default_root_vel_noise_lower, isl.default_root_vel_noise_upper = list(rc["root_vel_noise"][0]), list(rc["root_vel_noise"][1])
    dof_names = ["foot_slide", "wheel1_rotation", "wheel2_rotation", "wheel3_rotation"]
    isl.default_joint_angles = dict(zip(dof_names, rc["default_dof_pos"]))
    cfg.control.stiffness = dict(zip(dof_names, hp.p_gains))
    cfg.control.damping = dict(zip(dof_names, hp.d_gains))
    cfg.control.wheel_spindown = dict(zip(dof_names[1:], hp.kd_spindown))
    cfg.asset.wheel_speed_bounds = dict(zip(dof_names[1:], hp.wheel_speed_limits))
    cfg.asset.rot_actuator, cfg.asset.torque_speed_bound_ratio = [list(r) for r in hp.rot_actuator], hp.torque_speed_bound_ratio

    return cfg


def grid_origins(num_envs, spacing=3.0):
    """env_origins of a plane terrain (legged_robot_trajectory.py:866-875): the replayed robots stay near them, so that the tracking error
    (a difference of positions) is O(0.3 m) as in a real run instead of O(|origin|)."""
    cols = np.floor(np.sqrt(num_envs))
    rows = np.ceil(num_envs / cols)
    xx, yy = torch.meshgrid(torch.arange(rows), torch.arange(cols), indexing="ij")
    o = torch.zeros(num_envs, 3)
    o[:, 0], o[:, 1] = spacing * xx.flatten()[:num_envs], spacing * yy.flatten()[:num_envs]
    return o


def make_hopper_tape(num_envs, frames=8, seed=0, decimation=4, num_bodies=5, foot_body=4, term_prob=0.01, origins=None):
    """Synthetic physics frames of the Hopper's shape, one (dof, root, contact) triple per SUB-step (the Hopper refreshes root state
    and contacts inside the decimation loop): dof [F, D, N, 4, 2], root [F, D, N, 13], contact [F, D, N, B, 3], actions [F, N, 4]."""
    g = torch.Generator().manual_seed(seed)
    F, D, N, B = frames, decimation, num_envs, num_bodies
    rn = lambda *s: torch.randn(*s, generator=g)
    ru = lambda *s: torch.rand(*s, generator=g)
    root = torch.zeros(F, D, N, 13)
    base = (origins if origins is not None else torch.zeros(N, 3))[:, :2]
    root[..., 0:2] = base + 0.3 * rn(F, D, N, 2)
    root[..., 2] = 0.45 + 0.1 * rn(F, D, N)
    q = torch.cat([0.15 * rn(F, D, N, 3), torch.ones(F, D, N, 1)], dim=-1)
    root[..., 3:7] = q / q.norm(dim=-1, keepdim=True)
    root[..., 7:10] = 0.5 * rn(F, D, N, 3)
    root[..., 10:13] = 1.5 * rn(F, D, N, 3)
    dof = torch.zeros(F, D, N, 4, 2)
    dof[..., 0, 0] = 0.01 + 0.03 * rn(F, D, N)
    dof[..., 1:, 0] = 3.0 * rn(F, D, N, 3)
    dof[..., 0, 1] = 0.5 * rn(F, D, N)
    dof[..., 1:, 1] = 150.0 * rn(F, D, N, 3)
    contact = torch.zeros(F, D, N, B, 3)
    fz = torch.clamp(150.0 + 100.0 * rn(F, D, N), min=0.0) * (ru(F, D, N) > 0.5)
    contact[..., foot_body, 2] = fz
    contact[..., foot_body, 0:2] = 20.0 * rn(F, D, N, 2) * (fz > 0).unsqueeze(-1)
    hit = ru(F, D, N, B, 1) < term_prob
    hit[..., foot_body, :] = False
    contact += 8.0 * rn(F, D, N, B, 3) * hit
    acts = rn(F, N, 4) * torch.tensor([1.0, 0.2, 0.2, 0.2]) + torch.tensor([1.0, 0.0, 0.0, 0.0])
    acts *= 0.5 + ru(F, N, 1)                         # un-normalised quaternion actions, as a policy emits them
    return SimpleNamespace(root=root, dof=dof, contact=contact, actions=acts, num_envs=N, frames=F, decimation=D)


def make_domain_rand(num_envs, hp, seed=0):
    """The per-env multipliers HopperTrajectory._update_envs draws at construction (:372-413) — inputs of the parity contract."""
    g = torch.Generator().manual_seed(seed)
    N = num_envs
    ru = lambda lo, hi, *s: (hi - lo) * torch.rand(*s, generator=g) + lo
    return dict(p_gain_random=ru(0.9, 1.1, N, 4), d_gain_random=ru(0.9, 1.1, N, 4), spring_stiffness=ru(0.9, 1.1, N, 1) * 11732.0,
                spring_damping=ru(0.9, 1.1, N, 1) * 50.0, foot_pos_des=ru(0.75, 1.25, N, 1) * 0.03,
                torque_speed_bound_ratio_random=ru(0.9, 1.1, N, 1), torque_limit_random=ru(0.95, 1.05, N, 4), wheel_limit_random=ru(0.9, 1.1, N, 3))


class HopperTapePhysics:
    def __init__(self, tape):
        self.tape, self.frame, self.sub = tape, 0, 0

    def substep(self, env):
        """gym.simulate + refresh_dof_state / refresh_actor_root_state / refresh_net_contact_force of one sub-step (:113-120)."""
        t, f = self.tape, self.frame % self.tape.frames
        env.dof_state.copy_(t.dof[f, self.sub])
        env.root_states.copy_(t.root[f, self.sub])
        env.contact_forces.copy_(t.contact[f, self.sub])
        self.sub += 1
        if self.sub == t.decimation:
            self.sub, self.frame = 0, self.frame + 1


class HopperTrajPort:
    def __init__(self, hp, dr, tape, time_until_next_push, episode_length_buf=None, env_origins=None, rng="philox", env_id_offset=0):
        N = hp.num_envs
        f32 = dict(dtype=torch.float32)
        self.p, self.N, self.rng, self.off = hp, N, rng, env_id_offset
        self.root_states, self.dof_state = tape.root[0, 0].clone(), tape.dof[0, 0].clone()
        self.contact_forces = tape.contact[0, 0].clone()
        self.dof_pos, self.dof_vel = self.dof_state[..., 0], self.dof_state[..., 1]
        self.base_quat = self.root_states[:, 3:7]                         # a view (legged_robot_trajectory.py:601)
        self.gravity_vec = torch.tensor([0.0, 0.0, -1.0]).repeat(N, 1)
        self.base_lin_vel = quat_rotate_inverse(self.base_quat, self.root_states[:, 7:10])
        self.base_ang_vel = quat_rotate_inverse(self.base_quat, self.root_states[:, 10:13])
        self.projected_gravity = quat_rotate_inverse(self.base_quat, self.gravity_vec)
        self.common_step_counter = 0
        self.obs_buf = torch.zeros(N, 14 + hp.generator["N"] * 2 + 4, **f32)
        self.rew_buf = torch.zeros(N, **f32)
        self.reset_buf = torch.ones(N, dtype=torch.bool)
        self.time_out_buf = torch.zeros(N, dtype=torch.bool)
        self.episode_length_buf = torch.zeros(N, dtype=torch.long) if episode_length_buf is None else episode_length_buf.clone()
        self.zero_action = torch.tensor(hp.reset["zero_action"], **f32).repeat(N, 1)
        self.actions = self.zero_action.clone()                           # :417
        self.torques = torch.zeros(N, 4, **f32)
        self.last_actions, self.last_dof_vel, self.last_root_vel = torch.zeros(N, 4, **f32), torch.zeros(N, 4, **f32), torch.zeros(N, 6, **f32)
        self.feet_air_time = torch.zeros(N, 1, **f32)
        self.last_contacts = torch.zeros(N, 1, dtype=torch.bool)
        self.env_origins = torch.zeros(N, 3, **f32) if env_origins is None else env_origins.clone()
        self.dr = {k: v.clone() for k, v in dr.items()}
        self.active = [n for n in TERMS if hp.scales.get(n, 0.0) != 0.0]
        self.scale = {n: hp.scales[n] * hp.dt for n in self.active}       # _prepare_reward_function: scale * dt
        self.episode_sums = {n: torch.zeros(N, **f32) for n in sorted(self.active)}   # reward_scales.keys(): alphabetical, termination included
        self.extras = {}
        gp = SimpleNamespace(num_envs=N, dt=hp.dt)
        self.gen = RomPort(generator_params(gp, hp.generator), rng=rng, env_id_offset=env_id_offset)
        W = hp.generator["N"] // hp.generator["dN"]
        self.trajectory = torch.zeros(N, W, 2, **f32)
        self.trajectory_scale = torch.tensor(hp.trajectory_scale, **f32)[None, :].repeat(W, 1)
        self.prev_error = torch.zeros(N, 2, **f32)
        self.time_until_next_push = time_until_next_push.clone().reshape(N)   # :609 of the base class re-creates it as [N]
        self.reward_weighting = torch.tensor(hp.reward_weighting, **f32)
        self.dof_pos_limits = torch.tensor(hp.dof_pos_limits, **f32)
        self.dof_vel_limits, self.torque_limits = torch.tensor(hp.dof_vel_limits, **f32), torch.tensor(hp.torque_limits, **f32)
        self.default_dof_pos = torch.tensor(hp.reset["default_dof_pos"], **f32).unsqueeze(0)

    # ---- randomness / helpers --------------------------------------------------------------------------------------
    def _u(self, site, ids, ncols):
        if self.rng == "torch":
            return torch.rand(len(ids), ncols)
        return torch.from_numpy(P.uniform01(self.p.seed, ids.numpy() + self.off, self.common_step_counter, site, ncols))

    def proj_z(self):
        return self.root_states[:, :2].clone()

    def _case(self):
        hp = self.p
        c = dict(num_envs=self.N, control_type=hp.control_type, foot_body=hp.foot_body, action_scale=hp.action_scale,
                 torque_speed_bound_ratio=hp.torque_speed_bound_ratio, rot_actuator=hp.rot_actuator, dof_state=self.dof_state,
                 contact_forces=self.contact_forces, root_states=self.root_states, base_ang_vel=self.base_ang_vel,
                 base_lin_vel=self.base_lin_vel, actions=self.actions, p_gains=torch.tensor(hp.p_gains), d_gains=torch.tensor(hp.d_gains),
                 kd_spindown=torch.tensor(hp.kd_spindown), torque_limits=self.torque_limits,
                 wheel_speed_limits=torch.tensor(hp.wheel_speed_limits))
        c.update(self.dr)
        return c

    # ---- step (:102-133) -------------------------------------------------------------------------------------------
    def step(self, actions, physics):
        hp = self.p
        self.actions = torch.clip(actions, -hp.clip_actions, hp.clip_actions)
        for _ in range(hp.decimation):
            clipped, _raw = PH.hopper_torques(self._case(), self.actions)
            self.torques = clipped                                        # :112 re-binds self.torques to the clipped return value
            physics.substep(self)
            self.base_ang_vel[:] = quat_rotate_inverse(self.base_quat, self.root_states[:, 10:13])   # :124-126 (base_quat: self-copy of a view)
        self.post_physics_step()
        return self.obs_buf, None, self.rew_buf, self.reset_buf, self.extras

    # ---- post physics (:135-182) -----------------------------------------------------------------------------------
    def post_physics_step(self):
        hp = self.p
        self.episode_length_buf += 1
        self.common_step_counter += 1
        self.base_lin_vel = quat_rotate_inverse(self.base_quat, self.root_states[:, 7:10])
        self.projected_gravity = quat_rotate_inverse(self.base_quat, self.gravity_vec)
        self.gen.gen_step_idx(torch.arange(self.N))                      # callback (legged_robot_trajectory.py:405-417)
        self.trajectory = self.gen.get_trajectory().clone()
        self.time_until_next_push -= hp.decimation * hp.sim_dt
        need = self.time_until_next_push <= 0
        if hp.push_robots and torch.any(need):
            ids = need.nonzero().flatten()
            pushed = torch.unique(torch.cat((ids, torch.zeros(1, dtype=torch.long))))   # the [N, 1] mask quirk: env 0 rides along (:152)
            mv = torch.tensor(hp.reset["max_push_vel"], dtype=torch.float32)
            self.root_states[pushed, 7:13] = (mv - (-mv)) * self._u(P.SITE_HOP_PUSH, pushed, 6) + (-mv)        # :362-367
            lo, hi = hp.time_between_pushes
            self.time_until_next_push[need] = ((hi - lo) * self._u(P.SITE_PUSH_TIMER, ids, 1) + lo).flatten()
        f = torch.norm(self.contact_forces[:, hp.termination_bodies, :], dim=-1)                                # check_termination
        self.reset_buf = torch.any(f > 1.0, dim=1)
        self.time_out_buf = self.episode_length_buf > hp.max_episode_length
        self.reset_buf |= self.time_out_buf
        self._compute_reward()
        self._reset_idx(self.reset_buf.nonzero().flatten())
        self._compute_observations()
        self.last_actions[:] = self.actions
        self.last_dof_vel[:] = self.dof_vel
        self.last_root_vel[:] = self.root_states[:, 7:13]
        self.prev_error[:] = torch.square(self.trajectory[:, 0, :] - self.proj_z())                             # :178

    def _term_values(self):
        hp, s = self.p, self
        F, feet, pen = s.contact_forces, [hp.foot_body], hp.penalised_bodies
        wheels = slice(1, 4)

        def feet_air_time():                                              # legged_robot_trajectory.py:1071-1083
            contact = F[:, feet, 2] > 1.0
            filt = torch.logical_or(contact, s.last_contacts)
            s.last_contacts = contact
            first = (s.feet_air_time > 0.0) * filt
            s.feet_air_time += hp.dt
            r = torch.sum((s.feet_air_time - 0.5) * first, dim=1)
            s.feet_air_time *= ~filt
            return r

        def dof_pos_limits():
            o = -(s.dof_pos - s.dof_pos_limits[:, 0]).clip(max=0.0)
            o += (s.dof_pos - s.dof_pos_limits[:, 1]).clip(min=0.0)
            return torch.sum(o, dim=1)

        def tracking_rom():
            err = torch.inner(torch.square(s.proj_z() - s.trajectory[:, 0, :]), s.reward_weighting)
            return torch.exp(-err / hp.tracking_sigma)

        def differential_error():
            err = torch.norm(torch.square(s.proj_z() - s.trajectory[:, 0, :]), dim=-1)
            de = err - torch.norm(s.prev_error, dim=-1)
            return ((de < 0) * hp.diff_neg_slope + (de >= 0) * hp.diff_pos_slope) * de

        def raibert():                                                    # hopper_trajectory.py:482-505
            g = hp.raibert
            return PH.hopper_reward_raibert(dict(root_states=s.root_states, actions=s.actions), s.gen.get_trajectory()[:, 0], s.gen.v, g)

        return {
            "action_rate": lambda: torch.sum(torch.square(s.last_actions - s.actions), dim=1),
            "ang_vel_xy": lambda: torch.sum(torch.square(s.base_ang_vel[:, :2]), dim=1),
            "base_height": lambda: torch.square(torch.mean(s.root_states[:, 2].unsqueeze(1) - 0, dim=1) - hp.base_height_target),
            "collision": lambda: torch.sum(1.0 * (torch.norm(F[:, pen, :], dim=-1) > 0.1), dim=1),
            "differential_error": differential_error,
            "dof_acc": lambda: torch.sum(torch.square((s.last_dof_vel[:, wheels] - s.dof_vel[:, wheels]) / hp.dt), dim=1),          # :474-476
            "dof_pos_limits": dof_pos_limits,
            "dof_vel": lambda: torch.sum(torch.square(s.dof_vel), dim=1),
            "dof_vel_limits": lambda: torch.sum((torch.abs(s.dof_vel) - s.dof_vel_limits * hp.soft_dof_vel_limit).clip(min=0.0, max=1.0), dim=1),
            "feet_air_time": feet_air_time,
            "feet_contact_forces": lambda: torch.sum((torch.norm(F[:, feet, :], dim=-1) - hp.max_contact_force).clip(min=0.0), dim=1),
            "lin_vel_z": lambda: torch.square(s.base_lin_vel[:, 2]),
            "orientation": lambda: torch.sum(torch.square(s.projected_gravity[:, :2]), dim=1),
            "raibert": raibert,
            "stumble": lambda: torch.any(torch.norm(F[:, feet, :2], dim=2) > 5 * torch.abs(F[:, feet, 2]), dim=1),
            "torque_limits": lambda: torch.sum(torch.abs(s.torques[:, wheels]), dim=1),                                              # :470-472
            "torques": lambda: torch.sum(torch.square(s.torques), dim=1),
            "tracking_rom": tracking_rom,
            "unit_quat": lambda: torch.square(1 - torch.linalg.norm(s.actions, dim=-1)),                                             # :478-480
        }

    def _compute_reward(self):                                            # legged_robot_trajectory.py:255-272
        hp = self.p
        fn = self._term_values()
        self.rew_buf = torch.zeros(self.N)
        for name in self.active:
            if name == "termination":
                continue
            r = fn[name]() * self.scale[name]
            self.rew_buf += r
            self.episode_sums[name] += r
        if hp.only_positive_rewards:
            self.rew_buf[:] = torch.clip(self.rew_buf, min=0.0)
        if "termination" in self.scale:
            r = (self.reset_buf * ~self.time_out_buf) * self.scale["termination"]
            self.rew_buf += r
            self.episode_sums["termination"] += r

    def _reset_idx(self, ids):                                            # legged_robot_trajectory.py:204-246 with the Hopper's _reset_* (:298-358)
        hp = self.p
        if len(ids) == 0:
            return
        state = dict(dof_state=self.dof_state, root_states=self.root_states, actions=self.actions)
        if self.rng == "torch":
            raise NotImplementedError("the Hopper env port draws from the Philox streams only")
        PH.hopper_traj_reset(state, ids, self.env_origins, cfg=hp.reset, seed=hp.seed, event=self.common_step_counter, env_id_offset=self.off)
        self.gen.reset_traj(ids, self.proj_z())                           # :225, :248-253
        self.last_actions[ids] = 0.0
        self.last_dof_vel[ids] = 0.0
        self.feet_air_time[ids] = 0.0
        self.episode_length_buf[ids] = 0
        self.prev_error[ids] = torch.square(self.trajectory[ids, 0, :] - self.proj_z()[ids])
        self.extras["episode"] = {}
        for k in self.episode_sums:
            self.extras["episode"]["rew_" + k] = torch.mean(self.episode_sums[k][ids]) / hp.max_episode_length_s
            self.episode_sums[k][ids] = 0.0
        if hp.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf

    def _compute_observations(self):                                      # hopper_trajectory.py:255-282 + the clip of step (:128-129)
        hp = self.p
        c = dict(num_envs=self.N, actions=self.actions, root_states=self.root_states, base_lin_vel=self.base_lin_vel,
                 base_ang_vel=self.base_ang_vel, dof_state=self.dof_state)
        self.obs_buf = PH.hopper_traj_observations(c, self.trajectory, self.trajectory_scale, cfg=hp.obs, seed=hp.seed,
                                                   event=self.common_step_counter, env_id_offset=self.off)

    def reset(self, physics):                                             # hopper_trajectory.py:286-296
        ids = torch.arange(self.N)
        self.reset_idx(ids)
        self.step(self.zero_action.clone(), physics)
        self.reset_idx(ids)
        obs, priv, _, _, _ = self.step(self.zero_action.clone(), physics)
        return obs, priv

    def reset_idx(self, ids):
        """External call (the draws of an in-step reset and of an external one share common_step_counter as their event, as in the
        reference's own call sequence; the generator's per-env event counters keep the trajectory draws distinct)."""
        self._reset_idx(torch.as_tensor(ids, dtype=torch.long))
        self.reset_buf = self.reset_buf.clone()
        self.reset_buf[ids] = True
