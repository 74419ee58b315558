"""HopperTrajectory behind the reference's class surface, the Hopper class of the fork that can run a full step
(legged_gym/envs/hopper/hopper_trajectory.py; `Hopper` itself cannot reset as shipped).

`HopperTrajectory(cfg, sim_params, physics_engine, sim_device, headless)` takes the reference's constructor arguments plus what Isaac Gym
would have produced: `physics` (a source of per-sub-step dof / root / contact tensors, `HopperReplayPhysics` here), `asset` (body indices
and URDF limits, configs.HOPPER_ASSET) and optionally `domain_rand_values` (the per-env multipliers `_update_envs` draws at
construction, hopper_trajectory.py:372-413; drawn from a seeded torch generator when absent).  `step(actions)` = the decimation
loop — 4 x `_compute_torques` (b200gym_hopper_torques: the torque law with its refresh of base_ang_vel folded into the next sub-step's launch)
— then `post_physics_step`: generator step (b200gym_rom_step), prologue + fused post-physics + extras finaliser
(b200gym_hopper_post_physics, csrc/hopper_env.cu) and the generator reset of the envs that reset (b200gym_rom_reset_from_root).  Same
buffer names as the reference: obs_buf [N, 38], rew_buf, reset_buf, time_out_buf, extras, root_states, dof_pos / dof_vel, torques,
actions, last_*, base_lin_vel / base_ang_vel / projected_gravity, trajectory, prev_error, time_until_next_push, episode_sums, and the
randomised per-env properties.

Not built: `curriculum.use_curriculum` (update_command_curriculum rewrites reward tables, generator classes and rom bounds mid-run,
legged_robot_trajectory.py:519-555) and measured heights (hopper_trajectory_config.py:13 ships measure_heights = False): both raise.
"""
import numpy as np
import torch

from . import _lib
from . import rom as R
from .legged_robot_trajectory import _T_SAMPLERS, _W_SAMPLERS

TERMS = ("action_rate", "ang_vel_xy", "base_height", "collision", "differential_error", "dof_acc", "dof_pos_limits", "dof_vel",
         "dof_vel_limits", "feet_air_time", "feet_contact_forces", "lin_vel_z", "orientation", "raibert", "stumble", "torque_limits",
         "torques", "tracking_rom", "unit_quat", "termination")


class HopperReplayPhysics:
    """Replays per-SUB-step frames (the Hopper refreshes root state and contacts inside the decimation loop, hopper_trajectory.py:113-120):
    tape.dof [F, D, N, 4, 2], tape.root [F, D, N, 13], tape.contact [F, D, N, B, 3].  copy=False hands out views of the tape."""

    def __init__(self, tape, device="cuda", copy=True):
        dev = torch.device(device)
        self.dof_frames, self.root_frames, self.contact_frames = tape.dof.to(dev), tape.root.to(dev), tape.contact.to(dev)
        self.frames, self.decimation, self.copy = tape.frames, tape.decimation, copy
        self.frame = self.sub = 0
        self.dof_state, self.root_states = self.dof_frames[0, 0].clone(), self.root_frames[0, 0].clone()
        self.contact_forces = self.contact_frames[0, 0].clone()

    def simulate(self, torques):
        f = self.frame % self.frames
        self.dof_state.copy_(self.dof_frames[f, self.sub])
        self.root_states.copy_(self.root_frames[f, self.sub])
        self.contact_forces.copy_(self.contact_frames[f, self.sub])
        self.sub += 1
        if self.sub == self.decimation:
            self.sub, self.frame = 0, self.frame + 1

    def commit_resets(self, reset_buf):
        return None


class HopperTrajectory:
    num_dof = num_actions = 4

    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda", headless=True, physics=None, asset=None,
                 domain_rand_values=None, seed=0, env_id_offset=0):
        if physics is None:
            raise RuntimeError("HopperTrajectory needs a `physics` object (the PhysX call is out of scope: SURVEY.md §0 fact 2)")
        self.cfg, self.sim_params, self.physics = cfg, sim_params, physics
        self.device = torch.device(sim_device)
        if self.device.type != "cuda":
            raise RuntimeError("the b200gym Hopper env runs on CUDA devices only (no CPU fallback)")
        _lib.require_current_device(self.device)
        self.lib = _lib.lib()
        from .configs import HOPPER_ASSET, HOPPER_DOF_NAMES
        a = dict(HOPPER_ASSET)
        a.update(asset or {})
        self.asset, self.dof_names = a, list(HOPPER_DOF_NAMES)
        self.num_envs, self.num_obs, self.num_privileged_obs = int(cfg.env.num_envs), int(cfg.env.num_observations), None
        self.num_bodies = int(a["num_bodies"])
        self.seed, self.env_id_offset, self.headless = int(seed), int(env_id_offset), headless
        if getattr(cfg.curriculum, "use_curriculum", False):
            raise NotImplementedError("HopperTrajectory: curriculum.use_curriculum rewrites reward tables / generator classes mid-run "
                                      "(legged_robot_trajectory.py:519-555) — not built")
        if cfg.terrain.measure_heights:
            raise NotImplementedError("HopperTrajectory: measured heights are not built (hopper_trajectory_config.py:13 ships measure_heights = False)")
        if cfg.control.control_type not in ("orientation", "orientation_spindown"):
            raise NameError(f"Unknown controller type: {cfg.control.control_type} (the reference method runs 'orientation' and 'orientation_spindown')")
        self._parse_cfg(cfg)
        self._init_buffers(domain_rand_values)
        self._prepare_reward_function()
        self.init_done = True

    # ------------------------------------------------------------------ configuration (legged_robot_trajectory.py:877-902)
    def _parse_cfg(self, cfg):
        sim_dt = getattr(self.sim_params, "dt", None) or cfg.sim.dt
        self.sim_dt, self.decimation = float(sim_dt), int(cfg.control.decimation)
        self.dt = self.decimation * self.sim_dt
        self.obs_scales = cfg.normalization.obs_scales
        self.max_episode_length_s = cfg.env.episode_length_s
        self.max_episode_length = float(np.ceil(self.max_episode_length_s / self.dt))
        self.tracking_sigma = float(cfg.rewards.tracking_sigma)
        self.custom_origins = cfg.terrain.mesh_type in ("heightfield", "trimesh")
        if self.custom_origins:
            raise NotImplementedError("HopperTrajectory._reset_root_states adds a [n, 2] draw to 7 columns with custom origins (:331-337): the "
                                      "reference cannot run that branch")
        rc, tc, dev = cfg.rom, cfg.trajectory_generator, self.device
        if rc.cls != "SingleInt2D":
            raise ValueError(f"Raibert Heuristic not implemented for RoM {rc.cls}")          # hopper_trajectory.py:499
        self.rom = R.SingleInt2D(dt=rc.dt, z_min=rc.z_min, z_max=rc.z_max, v_min=rc.v_min, v_max=rc.v_max, n_robots=self.num_envs,
                                 backend="torch", device=dev)
        if tc.cls != "TrajectoryGenerator":
            raise NotImplementedError(f"{tc.cls}: the Hopper env drives the sampled TrajectoryGenerator")
        self.traj_gen = R.TrajectoryGenerator(
            self.rom, _T_SAMPLERS[tc.t_samp_cls](tc.t_low, tc.t_high, backend="torch", device=dev), _W_SAMPLERS[tc.weight_samp_cls](),
            dt_loop=self.dt, N=tc.N, freq_low=tc.freq_low, freq_high=tc.freq_high, seed=tc.seed, backend="torch", device=dev,
            prob_stationary=tc.prob_stationary, dN=tc.dN, env_id_offset=self.env_id_offset, generic_kernels=False)
        if self.traj_gen.N * self.rom.n != 20 or self.num_obs != 14 + 20 + 4:
            raise ValueError("the fused Hopper observation is 14 + 10 x 2 trajectory columns + 4 (num_observations = 38)")
        self.reward_weighting = self.rom.get_weighting_vector(cfg.rewards.reward_weighting)

    # ------------------------------------------------------------------ buffers (hopper_trajectory.py:59-100, 415-434; base :584-661)
    def _init_buffers(self, drv):
        cfg, N, dev, a = self.cfg, self.num_envs, self.device, self.asset
        z = lambda *s, dtype=torch.float: torch.zeros(*s, dtype=dtype, device=dev)
        f = lambda v: torch.tensor(v, dtype=torch.float, device=dev)
        self.common_step_counter, self.extras = 0, {}
        self.obs_buf, self.rew_buf, self.privileged_obs_buf = z(N, self.num_obs), z(N), None
        self.reset_buf, self.time_out_buf = torch.ones(N, dtype=torch.bool, device=dev), z(N, dtype=torch.bool)
        self.episode_length_buf = z(N, dtype=torch.long)
        self.zero_action = f(cfg.control.zero_action).reshape(1, -1).repeat(N, 1)
        self.actions = self.zero_action.clone()                           # :417
        self.torques, self._torques_raw = z(N, 4), z(N, 4)
        self.last_actions, self.last_dof_vel, self.last_root_vel = z(N, 4), z(N, 4), z(N, 6)
        self.base_lin_vel, self.base_ang_vel, self.projected_gravity = z(N, 3), z(N, 3), z(N, 3)
        self.gravity_vec = f([0.0, 0.0, -1.0]).repeat(N, 1)
        self.feet_air_time, self.last_contacts = z(N, 1), z(N, 1, dtype=torch.bool)
        self.feet_indices = torch.tensor([a["foot_body"]], device=dev)
        self.termination_contact_indices = torch.tensor(a["termination_bodies"], dtype=torch.long, device=dev)
        self.penalised_contact_indices = torch.tensor(a["penalised_bodies"], dtype=torch.long, device=dev)
        self.dof_pos_limits, self.dof_vel_limits, self.torque_limits = f(a["dof_pos_limits"]), f(a["dof_vel_limits"]), f(a["torque_limits"])
        isl = cfg.init_state
        self.base_init_state = f(list(isl.pos) + list(isl.rot) + list(isl.lin_vel) + list(isl.ang_vel))
        self.default_dof_pos = f([isl.default_joint_angles[n] for n in self.dof_names]).unsqueeze(0)
        self.p_gains = f([cfg.control.stiffness[n] for n in self.dof_names])
        self.d_gains = f([cfg.control.damping[n] for n in self.dof_names])
        self.kd_spindown = f([cfg.control.wheel_spindown.get(n, 0.0) for n in self.dof_names[1:]])
        self.wheel_speed_limits = f([cfg.asset.wheel_speed_bounds.get(n, float("inf")) for n in self.dof_names[1:]])
        self.torque_speed_bound_ratio = float(cfg.asset.torque_speed_bound_ratio)
        self.foot_joint_index, self.wheel_joint_indices = torch.tensor([0]), torch.tensor([1, 2, 3])
        self.trajectory = z(N, self.traj_gen.N, self.rom.n)
        self.trajectory_scale = f(list(self.obs_scales.trajectory))[None, :].repeat(self.traj_gen.N // self.traj_gen.dN, 1)
        self.prev_error = z(N, self.rom.n)
        d = cfg.domain_rand
        g = torch.Generator(device=dev).manual_seed(self.seed + 977)
        ru = lambda lo, hi, *s: lo + (hi - lo) * torch.rand(*s, generator=g, device=dev)
        lo, hi = d.time_between_pushes
        self.time_until_next_push = ru(lo, hi, N, 1)                      # legged_robot_trajectory.py:85-88 (initialisation-only draw)
        # plane terrain: a grid of origins (legged_robot_trajectory.py:866-875)
        cols = np.floor(np.sqrt(N))
        rows = np.ceil(N / cols)
        xx, yy = torch.meshgrid(torch.arange(rows), torch.arange(cols), indexing="ij")
        self.env_origins = z(N, 3)
        self.env_origins[:, 0] = (cfg.env.env_spacing * xx.flatten()[:N]).to(dev)
        self.env_origins[:, 1] = (cfg.env.env_spacing * yy.flatten()[:N]).to(dev)
        # randomised per-env properties (_update_envs, hopper_trajectory.py:372-413)
        sp, pd, ts = d.spring_properties, d.pd_gain_properties, d.torque_speed_properties
        one = lambda *s: torch.ones(*s, device=dev)
        self.spring_stiffness = (ru(*sp.stiffness_range, N, 1) if sp.randomize_stiffness else one(N, 1)) * cfg.asset.spring_stiffness
        self.spring_damping = (ru(*sp.damping_range, N, 1) if sp.randomize_damping else one(N, 1)) * cfg.asset.spring_damping
        self.foot_pos_des = (ru(*sp.setpoint_range, N, 1) if sp.randomize_setpoint else one(N, 1)) * cfg.control.foot_pos_des
        self.p_gain_random = ru(*pd.p_gain_range, N, 4) if pd.randomize_p_gain else one(N, 4)
        self.d_gain_random = ru(*pd.d_gain_range, N, 4) if pd.randomize_d_gain else one(N, 4)
        self.torque_speed_bound_ratio_random = ru(*ts.slope_range, N, 1) if ts.randomize_slope else one(N, 1)
        self.torque_limit_random = ru(*ts.max_torque_range, N, 4) if ts.randomize_max_torque else one(N, 4)
        self.wheel_limit_random = ru(*ts.max_speed_range, N, 3) if ts.randomize_max_speed else one(N, 3)
        for k, v in (drv or {}).items():
            dst = getattr(self, k)
            dst.copy_(torch.as_tensor(v, dtype=torch.float).reshape(dst.shape))
        ns, lv, osc = cfg.noise.noise_scales, cfg.noise.noise_level, self.obs_scales
        nv = z(self.num_obs)                                              # _get_noise_scale_vec, :439-468
        nv[0], nv[1:5], nv[5:8] = ns.z_pos * lv * osc.z_pos, ns.quat * lv, ns.lin_vel * lv * osc.lin_vel
        nv[8:11], nv[11:14] = ns.ang_vel * lv * osc.ang_vel, ns.dof_vel * lv * osc.dof_vel
        self.noise_scale_vec, self.add_noise = nv, bool(cfg.noise.add_noise)
        self.max_vel = f(list(d.max_push_vel))
        self.raibert_Kp, self.raibert_Kv, self.raibert_Kff = cfg.rewards.raibert.Kp, cfg.rewards.raibert.Kv, cfg.rewards.raibert.Kff
        g_ = self.traj_gen                                                # the generator writes its window straight into self.trajectory
        g_._s.env_trajectory = self.trajectory.data_ptr()
        g_._p.randomize_rom_distance = int(bool(d.randomize_rom_distance))
        g_._p.max_rom_distance[:] = R._pad4(list(d.max_rom_dist))
        g_._p.zero_rom_dist_llh = float(d.zero_rom_distance_likelihood)
        # torque-law POD
        tp = _lib.HopperTorqueParamsPOD()
        tp.num_envs, tp.num_bodies, tp.foot_body = N, self.num_bodies, int(a["foot_body"])
        tp.spindown = int("spindown" in cfg.control.control_type)
        tp.action_scale, tp.torque_speed_bound_ratio = float(cfg.control.action_scale), self.torque_speed_bound_ratio
        tp.p_gains[:], tp.d_gains[:], tp.kd_spindown[:] = self.p_gains.tolist(), self.d_gains.tolist(), self.kd_spindown.tolist()
        tp.wheel_speed_limits[:], tp.torque_limits[:] = self.wheel_speed_limits.tolist(), self.torque_limits.tolist()
        tp.rot_actuator[:] = [float(x) for row in cfg.asset.rot_actuator for x in row]
        self._tp = tp

    # ------------------------------------------------------------------ reward table (legged_robot_trajectory.py:663-690)
    def _prepare_reward_function(self):
        cfg, N, dev = self.cfg, self.num_envs, self.device
        scales = {k: v for k, v in vars(cfg.rewards.scales).items() if not k.startswith("_")}
        for name, v in scales.items():
            if name not in TERMS and v != 0:
                raise AttributeError(f"'HopperTrajectory' object has no attribute '_reward_{name}'")
        self.reward_scales = {k: float(v) * self.dt for k, v in sorted(scales.items()) if v != 0 and k in TERMS}
        names = list(self.reward_scales)                                   # alphabetical, termination in its place: the episode_sums keys
        K = len(names)
        self._sums = torch.zeros(max(K, 1), N, device=dev)
        self.episode_sums = {n: self._sums[i] for i, n in enumerate(names)}
        self._extras_out = torch.zeros(K + 2, device=dev)
        self._ws_sums = torch.zeros(K + 2, dtype=torch.double, device=dev)
        self._push_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.extras["episode"] = {"rew_" + n: self._extras_out[i] for i, n in enumerate(names)}
        self.extras["num_resets"] = self._extras_out[K + 1]
        if cfg.env.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf
        rw, isl, d, a = cfg.rewards, cfg.init_state, cfg.domain_rand, self.asset
        p = _lib.HopperEnvParamsPOD()
        p.num_envs, p.num_bodies, p.foot_body = N, self.num_bodies, int(a["foot_body"])
        p.num_term, p.num_pen, p.num_sum_rows = len(a["termination_bodies"]), len(a["penalised_bodies"]), K
        p.term_idx[:p.num_term], p.pen_idx[:p.num_pen] = a["termination_bodies"], a["penalised_bodies"]
        p.push_robots, p.only_positive = int(bool(d.push_robots)), int(bool(rw.only_positive_rewards))
        p.add_noise, p.randomize_yaw = int(self.add_noise), int(bool(isl.randomize_yaw))
        p.dt, p.push_dt = self.dt, self.decimation * self.sim_dt
        p.max_episode_length, p.max_episode_length_s = self.max_episode_length, float(self.max_episode_length_s)
        lo, hi = d.time_between_pushes
        p.push_t_lo, p.push_t_span = lo, hi - lo
        p.max_push_vel[:] = list(d.max_push_vel)
        for i, n in enumerate(TERMS):
            p.reward_scale[i] = self.reward_scales.get(n, 0.0)
            p.sum_row[i] = names.index(n) if n in self.reward_scales else -1
        p.tracking_sigma, p.soft_dof_vel_limit = self.tracking_sigma, float(rw.soft_dof_vel_limit)
        p.base_height_target, p.max_contact_force = float(rw.base_height_target), float(rw.max_contact_force)
        p.traj_weight[:] = [float(v) for v in self.reward_weighting.tolist()][:2]
        p.diff_neg_slope, p.diff_pos_slope = float(rw.differential_error.neg_slope), float(rw.differential_error.pos_slope)
        r = rw.raibert
        p.raibert[:] = [r.Kp, r.Kv, r.Kff, r.clip_pos, r.clip_vel, r.clip_ang]
        p.dof_pos_lo[:], p.dof_pos_hi[:] = [v[0] for v in a["dof_pos_limits"]], [v[1] for v in a["dof_pos_limits"]]
        p.dof_vel_limits[:] = a["dof_vel_limits"]
        f32 = lambda v: np.asarray(v, dtype=np.float32)
        span = lambda lo_, hi_: (f32(hi_) - f32(lo_)).tolist()            # torch_rand_vec_float: fp32(upper) - fp32(lower)
        p.default_dof_pos[:] = self.default_dof_pos[0].tolist()
        p.dof_pos_noise_lo[:], p.dof_pos_noise_span[:] = list(isl.default_dof_pos_noise_lower), span(isl.default_dof_pos_noise_lower, isl.default_dof_pos_noise_upper)
        p.dof_vel_noise_lo[:], p.dof_vel_noise_span[:] = list(isl.default_dof_vel_noise_lower), span(isl.default_dof_vel_noise_lower, isl.default_dof_vel_noise_upper)
        p.base_init_state[:] = self.base_init_state.tolist()
        p.root_pos_noise_lo[:] = list(isl.default_root_pos_noise_lower)[2:]
        p.root_pos_noise_span[:] = span(isl.default_root_pos_noise_lower, isl.default_root_pos_noise_upper)[2:]
        p.root_vel_noise_lo[:], p.root_vel_noise_span[:] = list(isl.default_root_vel_noise_lower), span(isl.default_root_vel_noise_lower, isl.default_root_vel_noise_upper)
        p.zero_action[:] = list(cfg.control.zero_action)
        osc = self.obs_scales
        p.z_pos_scale, p.lin_vel_scale, p.ang_vel_scale, p.dof_vel_scale = osc.z_pos, osc.lin_vel, osc.ang_vel, osc.dof_vel
        p.clip_obs = float(cfg.normalization.clip_observations)
        p.traj_scale[:] = list(osc.trajectory)[:2]
        p.noise_scale_vec[:] = self.noise_scale_vec[:14].tolist()
        p.seed_lo, p.seed_hi = self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF
        self._pod = p

    # ------------------------------------------------------------------ aliased physics tensors
    @property
    def root_states(self):
        return self.physics.root_states

    @property
    def dof_state(self):
        return self.physics.dof_state

    @property
    def contact_forces(self):
        return self.physics.contact_forces

    @property
    def dof_pos(self):
        return self.physics.dof_state.view(self.num_envs, 4, 2)[..., 0]

    @property
    def dof_vel(self):
        return self.physics.dof_state.view(self.num_envs, 4, 2)[..., 1]

    @property
    def base_quat(self):
        return self.physics.root_states[:, 3:7]

    # ------------------------------------------------------------------ the step (hopper_trajectory.py:102-133)
    def _compute_torques(self, actions, first=False):
        ph, ptr = self.physics, _lib.ptr
        b = _lib.HopperTorqueBuffersPOD()
        b.actions = actions.data_ptr()
        for name in ("p_gain_random", "d_gain_random", "torque_limit_random", "wheel_limit_random", "spring_stiffness", "spring_damping",
                     "foot_pos_des", "torque_speed_bound_ratio_random", "base_ang_vel"):
            setattr(b, name, getattr(self, name).data_ptr())
        b.dof_state, b.contact_forces, b.root_states = ptr(ph.dof_state), ptr(ph.contact_forces), ptr(ph.root_states)
        b.torques, b.torques_clipped = self._torques_raw.data_ptr(), self.torques.data_ptr()   # self.torques = the clipped return value (:112)
        self._tp.ang_vel_from_root = 0 if first else 1
        _lib.check(self.lib.b200gym_hopper_torques(self._tp, b, _lib.stream_ptr(self.device)), "hopper_torques")
        return self.torques

    def step(self, actions):
        _lib.require_cuda(actions, "actions")
        c = float(self.cfg.normalization.clip_actions)
        torch.clamp(actions.float(), -c, c, out=self.actions)
        for i in range(self.decimation):
            self._compute_torques(self.actions, first=(i == 0))
            self.physics.simulate(self.torques)
        self.post_physics_step()
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def _buffers(self):
        ph = self.physics
        b = getattr(self, "_buf_pod", None)
        if b is None:
            b = _lib.HopperEnvBuffersPOD()
            t = dict(actions=self.actions, torques=self.torques, last_actions=self.last_actions, last_dof_vel=self.last_dof_vel,
                     last_root_vel=self.last_root_vel, base_lin_vel=self.base_lin_vel, base_ang_vel=self.base_ang_vel,
                     projected_gravity=self.projected_gravity, feet_air_time=self.feet_air_time, last_contacts=self.last_contacts,
                     episode_length_buf=self.episode_length_buf, reset_buf=self.reset_buf, time_out_buf=self.time_out_buf, rew_buf=self.rew_buf,
                     episode_sums=self._sums, obs_buf=self.obs_buf, trajectory=self.trajectory, gen_v=self.traj_gen.v, prev_error=self.prev_error,
                     time_until_next_push=self.time_until_next_push, env_origins=self.env_origins, extras_out=self._extras_out,
                     ws_sums=self._ws_sums, push_flag=self._push_flag)
            for k, v in t.items():
                _lib.require_cuda(v, k)
                setattr(b, k, v.data_ptr())
            self._buf_pod = b
        b.root_states, b.dof_state, b.contact_forces = ph.root_states.data_ptr(), ph.dof_state.data_ptr(), ph.contact_forces.data_ptr()
        return b

    def post_physics_step(self):                                          # hopper_trajectory.py:135-182
        self.common_step_counter += 1
        g, st = self.traj_gen, _lib.stream_ptr(self.device)
        _lib.check(g.env_step(st), "rom_step")   # callback (base :409-410)
        _lib.check(self.lib.b200gym_hopper_post_physics(self._pod, self._buffers(), self.common_step_counter, self.env_id_offset, st),
                   "hopper_post_physics")
        K = self._extras_out.numel() - 2
        _lib.check(self.lib.b200gym_rom_reset_from_root(g._p, g._s, self.reset_buf.data_ptr(), self.physics.root_states.data_ptr(), 13,
                                                        self._extras_out[K + 1:].data_ptr(), self.env_id_offset, st), "rom_reset_from_root")
        self.physics.commit_resets(self.reset_buf)

    # ------------------------------------------------------------------ BaseTask surface
    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def reset_idx(self, env_ids):
        """LeggedRobotTrajectory.reset_idx called from outside step() (legged_robot_trajectory.py:204-246 with the Hopper's _reset_* methods):
        ONE masked launch (dof / action / root redraw, buffers cleared, reset_buf set, prev_error, extras) + the generator reset from the new
        roots; immediate, time_out_buf untouched.  Draws keyed by common_step_counter, as in the reference's own call sequence."""
        if len(env_ids) == 0:
            return
        mask = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        mask[env_ids] = 1
        g, st = self.traj_gen, _lib.stream_ptr(self.device)
        _lib.check(self.lib.b200gym_hopper_reset_idx(self._pod, self._buffers(), mask.data_ptr(), int(self.common_step_counter), self.env_id_offset, st),
                   "hopper_reset_idx")
        _lib.check(self.lib.b200gym_rom_reset_from_root(g._p, g._s, mask.data_ptr(), self.physics.root_states.data_ptr(), 13, None,
                                                        self.env_id_offset, st), "rom_reset_from_root")
        self.physics.commit_resets(mask.view(torch.bool))

    def reset_traj_all(self):
        """reset_traj (legged_robot_trajectory.py:248-253) of every env at the robots' current positions: a never-reset generator evaluates
        0/0 in its ramp input (rom_dynamics.py:552), a state the reference never steps from (its reset() resets the generators first)."""
        g = self.traj_gen
        mask = torch.ones(self.num_envs, dtype=torch.bool, device=self.device)
        _lib.check(self.lib.b200gym_rom_reset_from_root(g._p, g._s, mask.data_ptr(), self.physics.root_states.data_ptr(), 13, None,
                                                        self.env_id_offset, _lib.stream_ptr(self.device)), "rom_reset_from_root")

    def reset(self):                                                      # hopper_trajectory.py:286-296
        ids = torch.arange(self.num_envs, device=self.device)
        self.reset_idx(ids)
        self.step(self.zero_action.clone())
        self.reset_idx(ids)
        obs, priv, _, _, _ = self.step(self.zero_action.clone())
        return obs, priv
