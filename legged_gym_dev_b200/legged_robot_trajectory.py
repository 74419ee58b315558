"""Drop-in `LeggedRobotTrajectory` / `AnymalTrajectory`: the legged step pipeline composed with the
reduced-order-model trajectory generator.

Mirrors legged_gym/envs/base/legged_robot_trajectory.py:51-1110 and legged_gym/envs/anymal_c/anymal_trajectory.py:46-80:
same constructor arguments, attribute names (`trajectory`, `traj_gen`, `rom`, `prev_error`, `time_until_next_push`,
`trajectory_scale`, `reward_weighting`, ...) and step semantics.  One env step is

    4 x torques  ->  TrajectoryGenerator.step (+ get_trajectory)  ->  fused post-physics (traj_mode)  ->  generator reset
    b200gym_pd_torques / _lstm_torques   b200gym_rom_step            b200gym_post_physics               b200gym_rom_reset_from_root

with no host synchronisation: the reference's `torch.any(need_push)`, `nonzero` and `len(env_ids) == 0` host branches
(:172, :184, :217) are device-side predicates (the generator reset reads the reset count the fused kernel leaves in
`extras_out`).  What the reference does that parity depends on:
  * `self.trajectory` is the clone taken in `_post_physics_step_callback` (:410): prev_error (:233) and the observation
    (:277-283) of an env that resets this step still use it, together with the NEW root position;
  * commands do not exist (:621-622): no resampling, feet_air_time is not gated (:1082), the terrain curriculum (:508) and
    stand_still (:1092) would raise AttributeError, and do here;
  * pushes are per-env timers decremented by dt every step (:169-178).
There is no CPU path.
"""
import numpy as np
import torch

from . import _lib
from . import rom as R
from .legged_robot import LeggedRobot, ActuatorNetMixin
from .params import flatten_legged_cfg, TERM_ID

_ROM_CLASSES = {"SingleInt2D": R.SingleInt2D, "DoubleInt2D": R.DoubleInt2D}
_T_SAMPLERS = {"UniformSampleHoldDT": R.UniformSampleHoldDT}
_W_SAMPLERS = {"UniformWeightSampler": R.UniformWeightSampler, "UniformWeightSamplerNoExtreme": R.UniformWeightSamplerNoExtreme,
               "UniformWeightSamplerNoRamp": R.UniformWeightSamplerNoRamp}


class LeggedRobotTrajectory(LeggedRobot):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.cfg.curriculum.use_curriculum:                           # legged_robot_trajectory.py:82-83
            self.update_command_curriculum()
        lo, hi = self.params.time_between_pushes                          # :85-88 (initialisation-only draw)
        gen = torch.Generator(device=self.device).manual_seed(self.seed + 977)
        self.time_until_next_push.copy_(lo + (hi - lo) * torch.rand(self.num_envs, 1, generator=gen, device=self.device))

    # ------------------------------------------------------------------ configuration (legged_robot_trajectory.py:877-902)
    def _parse_cfg(self, cfg):
        t = self._terrain_in or {}
        hs = t.get("height_samples")
        sim_dt = getattr(self.sim_params, "dt", None) or cfg.sim.dt
        self.params = flatten_legged_cfg(
            cfg, sim_dt, self.dof_names, num_envs=self.num_envs, feet_indices=self.feet_indices.tolist(),
            penalised_indices=self.penalised_contact_indices.tolist(),
            termination_indices=self.termination_contact_indices.tolist(),
            dof_pos_limits=self.dof_pos_limits.tolist(), dof_vel_limits=self.dof_vel_limits.tolist(),
            torque_limits=self.torque_limits.tolist(),
            terrain_rows=hs.shape[0] if hs is not None else 0, terrain_cols=hs.shape[1] if hs is not None else 0,
            seed=self.seed, trajectory=True)
        p = self.params
        p.num_bodies = self.num_bodies
        if p.terrain_curriculum:
            # _update_terrain_curriculum reads self.commands (:508), which this class never creates (:621-622)
            raise AttributeError("'LeggedRobotTrajectory' object has no attribute 'commands'")
        if p.measure_heights and p.mesh_type == "none":
            raise NameError("Can't measure height with terrain mesh type 'none'")
        if p.measure_heights and p.mesh_type != "plane" and hs is None:
            raise RuntimeError("terrain.measure_heights needs terrain['height_samples']")
        self.dt = p.dt
        self.obs_scales = cfg.normalization.obs_scales
        self.max_episode_length_s = p.max_episode_length_s
        self.max_episode_length = p.max_episode_length
        self.custom_origins = p.custom_origins
        d, rc = cfg.domain_rand, cfg.rom
        self.nominal_rom_z_max, self.nominal_rom_z_min = list(rc.z_max), list(rc.z_min)
        self.nominal_rom_v_max, self.nominal_rom_v_min = list(rc.v_max), list(rc.v_min)
        self.nominal_max_rom_distance = list(d.max_rom_dist)
        self.nominal_zero_rom_distance_likelihood = float(d.zero_rom_distance_likelihood)
        self.curriculum_state = 0
        self.nominal_tracking_sigma = self.tracking_sigma = p.tracking_sigma
        self.nominal_push_time = float(np.ceil(d.push_interval_s / p.dt))
        self.nominal_max_push_vel = getattr(d, "max_push_vel", None)
        self.push_time, self.max_push_vel = self.nominal_push_time, self.nominal_max_push_vel
        # rom + generator (:90-123); their parameters become part of the POD, so they are created before the buffers
        self._init_rom()
        self._init_trajectory_generator()
        p.traj_n, p.traj_horizon = self.rom.n, self.traj_gen.N
        if p.num_obs != p.obs_width + p.num_height_points:
            raise ValueError(f"num_observations {p.num_obs} != {p.obs_width} + {p.num_height_points} height points")
        w = self.rom.get_weighting_vector(cfg.rewards.reward_weighting)
        self.reward_weighting = w
        p.traj_weight = [float(v) for v in w.tolist()] + [0.0] * (4 - w.numel())

    def _init_rom(self):                                                  # legged_robot_trajectory.py:90-103
        rc = self.cfg.rom
        if rc.cls not in _ROM_CLASSES:
            raise NotImplementedError(f"rom class {rc.cls}: SingleInt2D / DoubleInt2D are fused (SURVEY.md §8f-4)")
        self.rom = _ROM_CLASSES[rc.cls](dt=rc.dt, z_min=rc.z_min, z_max=rc.z_max, v_min=rc.v_min, v_max=rc.v_max,
                                        n_robots=self.num_envs, backend="torch", device=self.device)
        if self.rom.n != 2:
            raise NotImplementedError("the fused trajectory observation block is built for a 2-state rom (SingleInt2D)")

    def _init_trajectory_generator(self):                                 # legged_robot_trajectory.py:105-123
        tc = self.cfg.trajectory_generator
        classes = {"TrajectoryGenerator": R.TrajectoryGenerator, "ZeroTrajectoryGenerator": R.ZeroTrajectoryGenerator,
                   "SquareTrajectoryGenerator": R.SquareTrajectoryGenerator, "CircleTrajectoryGenerator": R.CircleTrajectoryGenerator}
        if tc.cls not in classes:
            raise NotImplementedError(f"{tc.cls}: not a fused trajectory generator class")
        self.traj_gen = classes[tc.cls](
            self.rom, _T_SAMPLERS[tc.t_samp_cls](tc.t_low, tc.t_high, backend="torch", device=self.device),
            _W_SAMPLERS[tc.weight_samp_cls](), dt_loop=self.dt, N=tc.N, freq_low=tc.freq_low, freq_high=tc.freq_high,
            seed=tc.seed, backend="torch", device=self.device, prob_stationary=tc.prob_stationary, dN=tc.dN,
            env_id_offset=self.env_id_offset, generic_kernels=False)   # the env drives b200gym_rom_step itself (views, gen_kind)

    # ------------------------------------------------------------------ buffers (legged_robot_trajectory.py:584-661)
    def _init_buffers(self):
        super()._init_buffers()
        p, N, dev = self.params, self.num_envs, self.device
        self.trajectory = torch.zeros(N, self.traj_gen.N, self.rom.n, dtype=torch.float, device=dev)
        self.trajectory_scale = torch.tensor(p.traj_scale[:self.rom.n], device=dev)[None, :].repeat(self.traj_gen.N // self.traj_gen.dN, 1)
        self.prev_error = torch.zeros(N, self.rom.n, dtype=torch.float, device=dev)
        self.time_until_next_push = torch.zeros(N, 1, dtype=torch.float, device=dev)   # [N,1] after __init__ (:85-88)
        self.max_rom_distance = torch.tensor(self.nominal_max_rom_distance, dtype=torch.float, device=dev)
        self.zero_rom_dist_llh = self.nominal_zero_rom_distance_likelihood
        nv = torch.zeros(self.num_obs, device=dev)                         # :557-581
        ts = self.traj_gen.N * self.rom.n
        nv[0:3], nv[3:6], nv[6:9] = p.noise_lin_vel, p.noise_ang_vel, p.noise_gravity
        nv[9 + ts:21 + ts], nv[21 + ts:33 + ts] = p.noise_dof_pos, p.noise_dof_vel
        if p.measure_heights:
            nv[45 + ts:] = p.noise_height
        self.noise_scale_vec = nv
        # :621-622: the class has no commands; the fused launcher still wants a valid pointer in that (ignored) slot
        self._commands_unused = self.commands
        del self.commands, self.commands_scale
        # the generator writes its interpolated window (get_trajectory, :410) straight into self.trajectory
        g = self.traj_gen
        g._s.env_trajectory = self.trajectory.data_ptr()
        g._p.randomize_rom_distance = int(bool(self.cfg.domain_rand.randomize_rom_distance))
        g._p.max_rom_distance[:] = R._pad4(self.nominal_max_rom_distance)
        g._p.zero_rom_dist_llh = self.zero_rom_dist_llh

    def _buffers(self):
        first = getattr(self, "_buf_pod", None) is None
        b = super()._buffers()
        if first:
            b.trajectory, b.prev_error = self.trajectory.data_ptr(), self.prev_error.data_ptr()
            b.time_until_next_push = self.time_until_next_push.data_ptr()
        return b

    # ------------------------------------------------------------------ the step (legged_robot_trajectory.py:150-192)
    def post_physics_step(self):
        self.physics.refresh()
        self.common_step_counter += 1
        cur = self.cfg.curriculum
        if cur.use_curriculum and self.curriculum_state < len(cur.curriculum_steps) and \
                self.common_step_counter % cur.curriculum_steps[self.curriculum_state] == 0:          # :414-417
            self.curriculum_state += 1
            self.update_command_curriculum()
        ev = self._event_start()
        st = self._stream if self._stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        g = self.traj_gen
        rc = g.env_step(st)                                                                      # :409-410
        if rc:
            _lib.check(rc, "rom_step")
        rc = self.lib.b200gym_post_physics(self._pod, self._buffers(), self.common_step_counter, self.env_id_offset, st)
        if rc:
            _lib.check(rc, "post_physics")
        K = self._extras_out.numel() - 3
        rc = self.lib.b200gym_rom_reset_from_root(g._p, g._s, self.reset_buf.data_ptr(), self.physics.root_states.data_ptr(), 13,
                                                  self._extras_out[K + 1:].data_ptr(), self.env_id_offset, st)   # :224, :248-253
        if rc:
            _lib.check(rc, "rom_reset_from_root")
        self._event_end("post_physics", ev)
        self.physics.commit_resets(self.reset_buf)

    def reset_traj(self, env_ids):                                        # :248-253 (stand-alone use)
        mask = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        mask[env_ids] = True
        g = self.traj_gen
        _lib.check(self.lib.b200gym_rom_reset_from_root(g._p, g._s, mask.data_ptr(), self.physics.root_states.data_ptr(), 13, None,
                                                        self.env_id_offset, _lib.stream_ptr(self.device)), "rom_reset_from_root")

    def update_command_curriculum(self):                                  # :519-555: host-side parameter rewrite
        cur, ind = self.cfg.curriculum, self.curriculum_state
        p, g = self.params, self.traj_gen
        p.max_push_vel = float(self.cfg.domain_rand.max_push_vel_xy)      # _push_robots reads the cfg, not max_push_vel (:489)
        self.push_time = self.nominal_push_time * cur.push.time[ind]
        self.max_rom_distance = torch.tensor(self.nominal_max_rom_distance, dtype=torch.float, device=self.device) * cur.max_rom_distance[ind]
        self.zero_rom_distance_likelihood = self.nominal_zero_rom_distance_likelihood * cur.zero_rom_distance_likelihood[ind]
        rom = self.rom
        t = lambda v, m: torch.tensor([x * m for x in v], dtype=torch.float, device=self.device)
        rom.z_max, rom.z_min = t(self.nominal_rom_z_max, cur.rom.z[ind]), t(self.nominal_rom_z_min, cur.rom.z[ind])
        rom.v_max, rom.v_min = t(self.nominal_rom_v_max, cur.rom.v[ind]), t(self.nominal_rom_v_min, cur.rom.v[ind])
        tc = self.cfg.trajectory_generator
        g.t_sampler = R.UniformSampleHoldDT(tc.t_low * cur.trajectory_generator.t_low[ind], tc.t_high * cur.trajectory_generator.t_high[ind])
        self.tracking_sigma = p.tracking_sigma = self.nominal_tracking_sigma * cur.sigma.tracking_rom[ind]
        for key in list(self.reward_scales.keys()):
            self.reward_scales[key] = getattr(self.cfg.rewards.scales, key) * getattr(cur.rewards, key)[ind] * self.dt
        # push the new numbers into the POD structs the kernels read
        for key, v in self.reward_scales.items():
            self._pod.reward_scale[TERM_ID[key]] = v
        self._pod.tracking_sigma = p.tracking_sigma
        g._p.rom_z_min[:], g._p.rom_z_max[:] = R._pad4(rom.z_min.tolist()), R._pad4(rom.z_max.tolist())
        g._p.rom_v_min[:], g._p.rom_v_max[:] = rom.v_min.tolist(), rom.v_max.tolist()
        g._p.t_low, g._p.t_span = g.t_sampler.t_low, g.t_sampler.t_high - g.t_sampler.t_low
        g._p.max_rom_distance[:] = R._pad4(self.max_rom_distance.tolist())


class AnymalTrajectory(ActuatorNetMixin, LeggedRobotTrajectory):
    """legged_gym/envs/anymal_c/anymal_trajectory.py:46-80: the trajectory env with the actuator-network torque path."""
