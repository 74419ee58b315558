"""CPU restatement (torch, fp32) of the reference's LeggedRobotTrajectory / AnymalTrajectory step.

TEST INFRASTRUCTURE — the checker, never the product (same rules as oracle/port_legged.py).  Follows
legged_gym/envs/base/legged_robot_trajectory.py (post_physics_step :150-192, reset_idx :204-246, reset_traj :248-253,
compute_observations :274-295, _post_physics_step_callback :405-417, _push_robots :486-492, _get_noise_scale_vec :557-581,
_reward_tracking_rom :1060-1069, _reward_feet_air_time :1071-1083, _reward_differential_error :1100-1110) on top of the
LeggedRobot restatement, with the trajectory generator of oracle/port_rom.py.  Pinned by tests/test_oracle_cpu.py against the
unmodified reference class (oracle/ref_harness.make_reference_anymal_trajectory) and by tests/golden/legged_traj_*.npz.
"""
from types import SimpleNamespace

import torch

from . import philox as P
from .port_legged import LeggedPort
from .port_rom import RomPort


def generator_params(p, tg):
    """Numbers the generator port needs, from LeggedParams `p` and the trajectory_generator / rom / domain_rand cfg values
    in `tg` (a dict: rom_dt, vel_max_rom, N, dN, t_low, t_high, freq_low, freq_high, prob_stationary, weight_sampler, seed,
    randomize_rom_distance, max_rom_distance, zero_rom_dist_llh)."""
    d = dict(num_envs=p.num_envs, model_cls="SingleInt2D", rom_cls="SingleInt2D", model_dt=p.dt, rom_dt=tg["rom_dt"], pos_max=1e9,
             vel_max=1.0, acc_max=1.0, vel_max_rom=tg["vel_max_rom"], N=tg["N"], dN=tg["dN"], t_low=float(tg["t_low"]),
             t_high=float(tg["t_high"]), freq_low=tg["freq_low"], freq_high=tg["freq_high"], prob_stationary=tg["prob_stationary"],
             weight_sampler=tg["weight_sampler"], randomize_rom_distance=tg["randomize_rom_distance"],
             max_rom_distance=list(tg["max_rom_distance"]), zero_rom_dist_llh=tg["zero_rom_dist_llh"],
             noise_lower=[0.0, 0.0], noise_upper=[0.0, 0.0], Kp=0.0, Kd=0.0, seed=tg["seed"], episode_length_s=20,
             generator=tg.get("generator", "TrajectoryGenerator"))
    return SimpleNamespace(**d)


class LeggedTrajPort(LeggedPort):
    def __init__(self, p, tg, root_states, dof_state, contact_forces, time_until_next_push, **kw):
        super().__init__(p, root_states, dof_state, contact_forces, **kw)
        N = p.num_envs
        f32 = dict(dtype=torch.float32)
        # the generator (dt_loop = env dt: legged_robot_trajectory.py:113); only its TrajectoryGenerator half is used
        self.gen = RomPort(generator_params(p, tg), rng=self.rng, env_id_offset=self.off)
        self.trajectory = torch.zeros(N, tg["N"], 2, **f32)
        self.trajectory_scale = torch.tensor(p.traj_scale[:2], **f32)[None, :].repeat(tg["N"] // tg["dN"], 1)
        self.prev_error = torch.zeros(N, 2, **f32)
        self.time_until_next_push = time_until_next_push.clone().reshape(N, 1)
        self.reward_weighting = torch.tensor(p.traj_weight[:2], **f32)
        ts = tg["N"] * 2
        nv = torch.zeros(p.num_obs, **f32)                                # :557-581
        nv[0:3], nv[3:6], nv[6:9] = p.noise_lin_vel, p.noise_ang_vel, p.noise_gravity
        nv[9 + ts:21 + ts], nv[21 + ts:33 + ts] = p.noise_dof_pos, p.noise_dof_vel
        if p.measure_heights:
            nv[45 + ts:] = p.noise_height
        self.noise_scale_vec = nv
        del self.commands

    def proj_z(self):
        return self.root_states[:, :2].clone()                            # SingleInt2D.proj_z, rom_dynamics.py:195-196

    # ---- post physics (legged_robot_trajectory.py:150-192) ---------------------------------------
    def post_physics_step(self):
        from .isaacgym_restated import quat_rotate_inverse
        p = self.p
        self.episode_length_buf += 1
        self.common_step_counter += 1
        self.base_lin_vel = quat_rotate_inverse(self.base_quat, self.root_states[:, 7:10])
        self.base_ang_vel = quat_rotate_inverse(self.base_quat, self.root_states[:, 10:13])
        self.projected_gravity = quat_rotate_inverse(self.base_quat, self.gravity_vec)
        # callback (:405-417)
        self.gen.gen_step_idx(torch.arange(self.N))
        self.trajectory = self.gen.get_trajectory().clone()
        if p.measure_heights:
            self.measured_heights = self.get_heights()
        # per-env push timers (:169-178)
        self.time_until_next_push -= p.decimation * p.sim_dt
        need = (self.time_until_next_push <= 0).reshape(-1)
        if torch.any(need):
            ids = need.nonzero().flatten()
            self.root_states[need, 7:9] = self._uniform(-p.max_push_vel, p.max_push_vel, P.SITE_PUSH, ids, 2)
            lo, hi = p.time_between_pushes
            self.time_until_next_push[need] = self._uniform(lo, hi, P.SITE_PUSH_TIMER, ids, 1)
        # termination (:194-202)
        f = torch.norm(self.contact_forces[:, self.term, :], dim=-1)
        self.reset_buf = torch.any(f > 1.0, dim=1)
        self.time_out_buf = self.episode_length_buf > p.max_episode_length
        self.reset_buf |= self.time_out_buf
        self._compute_reward()
        self._reset_idx(self.reset_buf.nonzero().flatten())
        self._compute_observations()
        self.last_actions[:] = self.actions
        self.last_dof_vel[:] = self.dof_vel
        self.last_root_vel[:] = self.root_states[:, 7:13]

    def _term_values(self):
        p, s = self.p, self
        fn = super()._term_values()
        F = s.contact_forces

        def feet_air_time():                                              # :1071-1083 (no command gate)
            contact = F[:, s.feet, 2] > 1.0
            filt = torch.logical_or(contact, s.last_contacts)
            s.last_contacts = contact
            first = (s.feet_air_time > 0.0) * filt
            s.feet_air_time += p.dt
            r = torch.sum((s.feet_air_time - 0.5) * first, dim=1)
            s.feet_air_time *= ~filt
            return r

        def tracking_rom():                                               # :1060-1069
            err = torch.inner(torch.square(s.proj_z() - s.trajectory[:, 0, :]), s.reward_weighting)
            return torch.exp(-err / p.tracking_sigma)

        def differential_error():                                         # :1100-1110
            err = torch.norm(torch.square(s.proj_z() - s.trajectory[:, 0, :]), dim=-1)
            de = err - torch.norm(s.prev_error, dim=-1)
            return ((de < 0) * p.diff_neg_slope + (de >= 0) * p.diff_pos_slope) * de

        fn.update(feet_air_time=feet_air_time, tracking_rom=tracking_rom, differential_error=differential_error)
        for k in ("stand_still", "tracking_ang_vel", "tracking_lin_vel"):   # no self.commands / no such _reward_* in the class
            fn.pop(k)
        return fn

    def _reset_idx(self, ids):                                            # :204-246 + anymal_trajectory.py:56-60
        p = self.p
        if len(ids) == 0:
            return
        self.dof_pos[ids] = self.default_dof_pos * self._uniform(0.5, 1.5, P.SITE_RESET_DOF, ids, p.num_dof)
        self.dof_vel[ids] = 0.0
        self.root_states[ids] = self.base_init_state
        self.root_states[ids, :3] += self.env_origins[ids]
        if p.custom_origins:
            self.root_states[ids, :2] += self._uniform(-1.0, 1.0, P.SITE_RESET_XY, ids, 2)
        self.root_states[ids, 7:13] = self._uniform(-0.5, 0.5, P.SITE_RESET_VEL, ids, 6)
        self.gen.reset_traj(ids, self.proj_z())                           # :248-253
        self.last_actions[ids] = 0.0
        self.last_dof_vel[ids] = 0.0
        self.feet_air_time[ids] = 0.0
        self.episode_length_buf[ids] = 0
        self.prev_error[ids] = torch.square(self.trajectory[ids, 0, :] - self.proj_z()[ids])   # :233 (stale trajectory, new root)
        self.extras["episode"] = {}
        for k in self.episode_sums:
            self.extras["episode"]["rew_" + k] = torch.mean(self.episode_sums[k][ids]) / p.max_episode_length_s
            self.episode_sums[k][ids] = 0.0
        if p.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf
        if p.use_actuator_network:
            D = p.num_dof
            self.sea_hidden_state.view(2, self.N, D, 8)[:, ids] = 0.0
            self.sea_cell_state.view(2, self.N, D, 8)[:, ids] = 0.0

    def _compute_observations(self):                                      # :274-295
        p = self.p
        mod = self.trajectory.clone()
        mod -= self.proj_z()[:, None, :2]
        obs = torch.cat((self.base_lin_vel * p.obs_lin_vel, self.base_ang_vel * p.obs_ang_vel, self.projected_gravity,
                         (mod * self.trajectory_scale).reshape(self.N, -1),
                         (self.dof_pos - self.default_dof_pos) * p.obs_dof_pos, self.dof_vel * p.obs_dof_vel,
                         self.actions), dim=-1)
        if p.measure_heights:
            h = torch.clip(self.root_states[:, 2].unsqueeze(1) - 0.5 - self.measured_heights, -1, 1.0) * p.obs_height
            obs = torch.cat((obs, h), dim=-1)
        if p.add_noise:
            u = self._u(P.SITE_OBS_NOISE, torch.arange(self.N), p.num_obs)
            obs += (2 * u - 1) * self.noise_scale_vec
        self.obs_buf = obs
