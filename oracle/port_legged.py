"""CPU restatement (torch, fp32) of the reference's LeggedRobot / Anymal per-step pipeline.

TEST INFRASTRUCTURE — the checker, never the product.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this.  It travels to the GPU box (the
reference itself cannot), so every function cites the reference lines it follows; it is *pinned* by
tests/test_oracle_cpu.py, which runs it step for step against the unmodified reference
(oracle/ref_harness.py) in the build container, and by the golden vectors in tests/golden/ written by
oracle/make_golden.py from that same reference run.

State lives in a flat namespace of tensors named as in the reference (legged_robot.py:533-603,
anymal.py:62-69).  Randomness: `rng="philox"` draws the counter-based stream of oracle/philox.py (parity
mode); `rng="torch"` uses torch.rand like the reference does (timing mode for the CPU baseline).
"""
import math
from types import SimpleNamespace

import numpy as np
import torch

from . import philox as P
from .isaacgym_restated import quat_apply, quat_rotate_inverse, normalize

# union of the LeggedRobot and LeggedRobotTrajectory terms, alphabetical, `termination` last (== params.REWARD_TERMS)
TERMS = ("action_rate", "ang_vel_xy", "base_height", "collision", "differential_error", "dof_acc", "dof_pos_limits", "dof_vel",
         "dof_vel_limits", "feet_air_time", "feet_contact_forces", "lin_vel_z", "orientation", "stand_still",
         "stumble", "torque_limits", "torques", "tracking_ang_vel", "tracking_lin_vel", "tracking_rom", "termination")


class LeggedPort:
    def __init__(self, p, root_states, dof_state, contact_forces, env_origins=None, height_samples=None,
                 terrain_levels=None, terrain_types=None, terrain_origins=None, lstm=None,
                 episode_length_buf=None, rng="philox", env_id_offset=0):
        """p: legged_gym_dev_b200.params.LeggedParams (plain numbers).  The three physics tensors are the
        aliased PhysX buffers (legged_robot.py:545-551) and are mutated in place on reset."""
        N, D = p.num_envs, p.num_dof
        f32 = dict(dtype=torch.float32)
        self.p, self.N, self.rng, self.off = p, N, rng, env_id_offset
        self.root_states, self.dof_state = root_states, dof_state
        self.contact_forces = contact_forces.view(N, -1, 3)
        self.dof_pos = self.dof_state.view(N, D, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, D, 2)[..., 1]
        self.base_quat = self.root_states[:, 3:7]
        self.common_step_counter = 0
        self.obs_buf = torch.zeros(N, p.num_obs, **f32)
        self.rew_buf = torch.zeros(N, **f32)
        self.reset_buf = torch.ones(N, dtype=torch.bool)
        self.time_out_buf = torch.zeros(N, dtype=torch.bool)
        self.episode_length_buf = (torch.zeros(N, dtype=torch.long) if episode_length_buf is None
                                   else episode_length_buf.clone())
        self.torques = torch.zeros(N, D, **f32)
        self.actions = torch.zeros(N, D, **f32)
        self.last_actions = torch.zeros(N, D, **f32)
        self.last_dof_vel = torch.zeros(N, D, **f32)
        self.last_root_vel = torch.zeros(N, 6, **f32)
        self.commands = torch.zeros(N, 4, **f32)
        self.feet_air_time = torch.zeros(N, len(p.feet_indices), **f32)
        self.last_contacts = torch.zeros(N, len(p.feet_indices), dtype=torch.bool)
        self.gravity_vec = torch.tensor([0.0, 0.0, -1.0]).repeat(N, 1)
        self.forward_vec = torch.tensor([1.0, 0.0, 0.0]).repeat(N, 1)
        self.base_lin_vel = quat_rotate_inverse(self.base_quat, self.root_states[:, 7:10])
        self.base_ang_vel = quat_rotate_inverse(self.base_quat, self.root_states[:, 10:13])
        self.projected_gravity = quat_rotate_inverse(self.base_quat, self.gravity_vec)
        self.measured_heights = 0
        self.episode_sums = {n: torch.zeros(N, **f32) for n in p.active_terms}
        self.extras = {}
        t = lambda v: torch.tensor(v, **f32)
        self.p_gains, self.d_gains = t(p.p_gains), t(p.d_gains)
        self.default_dof_pos = t(p.default_dof_pos).unsqueeze(0)
        self.torque_limits, self.dof_vel_limits = t(p.torque_limits), t(p.dof_vel_limits)
        self.dof_pos_limits = t(p.dof_pos_limits)
        self.base_init_state = t(p.base_init_state)
        self.feet = torch.tensor(p.feet_indices)
        self.pen = torch.tensor(p.penalised_indices)
        self.term = torch.tensor(p.termination_indices)
        self.env_origins = torch.zeros(N, 3, **f32) if env_origins is None else env_origins
        self.height_samples = height_samples
        self.terrain_levels, self.terrain_types, self.terrain_origins = terrain_levels, terrain_types, terrain_origins
        nv = torch.zeros(p.num_obs, **f32)                                # legged_robot.py:517-529
        nv[0:3], nv[3:6], nv[6:9] = p.noise_lin_vel, p.noise_ang_vel, p.noise_gravity
        nv[12:24], nv[24:36] = p.noise_dof_pos, p.noise_dof_vel
        if p.measure_heights:
            nv[48:] = p.noise_height
            gx, gy = torch.meshgrid(t(p.measured_points_x), t(p.measured_points_y), indexing="ij")   # :867-875
            self.height_points = torch.zeros(N, gx.numel(), 3, **f32)
            self.height_points[:, :, 0] = gx.flatten()
            self.height_points[:, :, 1] = gy.flatten()
        self.noise_scale_vec = nv
        self.commands_scale = t([p.obs_lin_vel, p.obs_lin_vel, p.obs_ang_vel])
        # command curriculum of the fork (legged_robot.py:822-833, 488-506)
        self.nominal_command_ranges = dict(lin_vel_x=list(p.cmd_lin_vel_x), lin_vel_y=list(p.cmd_lin_vel_y),
                                           ang_vel_yaw=list(p.cmd_ang_vel_yaw), heading=list(p.cmd_heading))
        self.command_ranges, self.curriculum_state = self.nominal_command_ranges, 0
        if getattr(p, "use_curriculum", False):
            self.update_command_curriculum()
        if lstm is not None:                                              # anymal.py:62-69
            self.lstm = {k: torch.as_tensor(v, **f32) for k, v in lstm.items()}
            self.sea_hidden_state = torch.zeros(2, N * D, 8, **f32)
            self.sea_cell_state = torch.zeros(2, N * D, 8, **f32)

    def update_command_curriculum(self):                                  # legged_robot.py:488-506
        c = self.p.curriculum_commands[self.curriculum_state]
        self.command_ranges = {k: [v * c if k != "heading" else v for v in val] for k, val in self.nominal_command_ranges.items()}

    # ---- randomness -----------------------------------------------------------------------------
    def _u(self, site, env_ids, ncols):
        n = len(env_ids)
        if self.rng == "torch":
            return torch.rand(n, ncols)
        ids = env_ids.numpy() + self.off
        return torch.from_numpy(P.uniform01(self.p.seed, ids, self._event(), site, ncols))

    def _event(self):
        """Philox event of the current draw: common_step_counter inside a step, or the external-reset event (reset_idx)."""
        ev = getattr(self, "_ext_event", None)
        return self.common_step_counter if ev is None else ev

    def _uniform(self, lo, hi, site, env_ids, ncols, col=None):
        u = self._u(site, env_ids, ncols)
        if col is not None:
            u = u[:, col:col + 1]
        return (hi - lo) * u + lo                                         # torch_rand_float, helpers.py:129-130

    # ---- torques (R2, R3) -----------------------------------------------------------------------
    def pd_torques(self, actions):
        """legged_robot.py:389-413."""
        p = self.p
        a = actions * p.action_scale
        if p.control_type == 0:
            tq = self.p_gains * (a + self.default_dof_pos - self.dof_pos) - self.d_gains * self.dof_vel
        elif p.control_type == 1:
            tq = self.p_gains * (a - self.dof_vel) - self.d_gains * (self.dof_vel - self.last_dof_vel) / p.sim_dt
        else:
            tq = a
        return torch.clip(tq, -self.torque_limits, self.torque_limits)

    def lstm_torques(self, actions):
        """anymal.py:71-78 with the TorchScript LSTMsea written out: 2-layer LSTM(2->8->8),
        gate order i,f,g,o, then out_scale * Linear(8->1).  Updates h,c in place."""
        p, w = self.p, self.lstm
        x = torch.stack([(actions * p.action_scale + self.default_dof_pos - self.dof_pos).flatten() * 2.0,
                         self.dof_vel.flatten() * 0.25], dim=-1)
        inp = x
        for l in range(2):
            g = (inp @ w[f"w_ih{l}"].T + w[f"b_ih{l}"]) + (self.sea_hidden_state[l] @ w[f"w_hh{l}"].T + w[f"b_hh{l}"])
            i, f, gg, o = g.split(8, dim=-1)
            c = torch.sigmoid(f) * self.sea_cell_state[l] + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            self.sea_cell_state[l], self.sea_hidden_state[l] = c, h
            inp = h
        return (20.0 * (inp @ w["w_lin"].T + w["b_lin"]).squeeze(-1)).view(self.N, p.num_dof)

    def compute_torques(self, actions):
        tq = self.lstm_torques(actions) if self.p.use_actuator_network else self.pd_torques(actions)
        self.torques = tq.view(self.torques.shape)
        return self.torques

    # ---- post physics (R4-R12) ------------------------------------------------------------------
    def post_physics_step(self):
        """legged_robot.py:106-134 (order is load-bearing)."""
        p = self.p
        self.episode_length_buf += 1
        self.common_step_counter += 1
        self.base_lin_vel = quat_rotate_inverse(self.base_quat, self.root_states[:, 7:10])
        self.base_ang_vel = quat_rotate_inverse(self.base_quat, self.root_states[:, 10:13])
        self.projected_gravity = quat_rotate_inverse(self.base_quat, self.gravity_vec)
        # callback (:343-359)
        ids = (self.episode_length_buf % p.resample_steps == 0).nonzero().flatten()
        self._resample_commands(ids, P.SITE_CMD_PERIODIC)
        if p.heading_command:
            fwd = quat_apply(self.base_quat, self.forward_vec)
            heading = torch.atan2(fwd[:, 1], fwd[:, 0])
            self.commands[:, 2] = torch.clip(0.5 * self._wrap_to_pi(self.commands[:, 3] - heading), -1.0, 1.0)
        if p.measure_heights:
            self.measured_heights = self.get_heights()
        if p.push_robots and self.common_step_counter % p.push_time == 0:
            all_ids = torch.arange(self.N)
            self.root_states[:, 7:9] = self._uniform(-p.max_push_vel, p.max_push_vel, P.SITE_PUSH, all_ids, 2)
        if getattr(p, "use_curriculum", False) and self.curriculum_state < len(p.curriculum_steps) and \
                self.common_step_counter % p.curriculum_steps[self.curriculum_state] == 0:                   # :360-363
            self.curriculum_state += 1
            self.update_command_curriculum()
        # termination (:139-145)
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

    @staticmethod
    def _wrap_to_pi(a):                                                   # math.py:45-48
        a = a % (2 * np.pi)
        a = a - 2 * np.pi * (a > np.pi)
        return a

    def _resample_commands(self, ids, site):                              # legged_robot.py:365-387
        p = self.p
        if len(ids) == 0 and self.rng == "philox":
            return
        cr = self.command_ranges
        rx, ry = cr["lin_vel_x"], cr["lin_vel_y"]
        self.commands[ids, 0] = self._uniform(rx[0], rx[1], site, ids, 3, 0).squeeze(1)
        self.commands[ids, 1] = self._uniform(ry[0], ry[1], site, ids, 3, 1).squeeze(1)
        if p.heading_command:
            r = cr["heading"]
            self.commands[ids, 3] = self._uniform(r[0], r[1], site, ids, 3, 2).squeeze(1)
        else:
            r = cr["ang_vel_yaw"]
            self.commands[ids, 2] = self._uniform(r[0], r[1], site, ids, 3, 2).squeeze(1)
        self.commands[ids, :2] *= (torch.norm(self.commands[ids, :2], dim=1) > 0.2).unsqueeze(1)

    def get_heights(self):
        """legged_robot.py:877-915 + math.py:38-42: yaw-only rotation, trunc to cell, min of 3 int16 samples."""
        p = self.p
        if p.mesh_type == "plane":
            return torch.zeros(self.N, p.num_height_points)
        if p.mesh_type == "none":
            raise NameError("Can't measure height with terrain mesh type 'none'")
        H = p.num_height_points
        qy = self.base_quat.repeat(1, H).clone().view(-1, 4)
        qy[:, :2] = 0.0
        qy = normalize(qy)
        pts = quat_apply(qy, self.height_points) + self.root_states[:, :3].unsqueeze(1)
        pts = pts + p.border_size
        pts = (pts / p.horizontal_scale).long()
        px = torch.clip(pts[:, :, 0].reshape(-1), 0, self.height_samples.shape[0] - 2)
        py = torch.clip(pts[:, :, 1].reshape(-1), 0, self.height_samples.shape[1] - 2)
        h = torch.min(torch.min(self.height_samples[px, py], self.height_samples[px + 1, py]),
                      self.height_samples[px, py + 1])
        return h.view(self.N, -1) * p.vertical_scale

    def _term_values(self):
        """All 18 loop terms of legged_robot.py:918-1015, lazily, keyed by name."""
        p, s = self.p, self
        F = s.contact_forces
        cmd_norm = lambda: torch.norm(s.commands[:, :2], dim=1)

        def feet_air_time():                                              # :988-1000 (stateful)
            contact = F[:, s.feet, 2] > 1.0
            filt = torch.logical_or(contact, s.last_contacts)
            s.last_contacts = contact
            first = (s.feet_air_time > 0.0) * filt
            s.feet_air_time += p.dt
            r = torch.sum((s.feet_air_time - 0.5) * first, dim=1)
            r *= cmd_norm() > 0.1
            s.feet_air_time *= ~filt
            return r

        def dof_pos_limits():
            o = -(s.dof_pos - s.dof_pos_limits[:, 0]).clip(max=0.0)
            o += (s.dof_pos - s.dof_pos_limits[:, 1]).clip(min=0.0)
            return torch.sum(o, dim=1)

        return {
            "action_rate": lambda: torch.sum(torch.square(s.last_actions - s.actions), dim=1),
            "ang_vel_xy": lambda: torch.sum(torch.square(s.base_ang_vel[:, :2]), dim=1),
            "base_height": lambda: torch.square(
                torch.mean(s.root_states[:, 2].unsqueeze(1) - s.measured_heights, dim=1) - p.base_height_target),
            "collision": lambda: torch.sum(1.0 * (torch.norm(F[:, s.pen, :], dim=-1) > 0.1), dim=1),
            "dof_acc": lambda: torch.sum(torch.square((s.last_dof_vel - s.dof_vel) / p.dt), dim=1),
            "dof_pos_limits": dof_pos_limits,
            "dof_vel": lambda: torch.sum(torch.square(s.dof_vel), dim=1),
            "dof_vel_limits": lambda: torch.sum(
                (torch.abs(s.dof_vel) - s.dof_vel_limits * p.soft_dof_vel_limit).clip(min=0.0, max=1.0), dim=1),
            "feet_air_time": feet_air_time,
            "feet_contact_forces": lambda: torch.sum(
                (torch.norm(F[:, s.feet, :], dim=-1) - p.max_contact_force).clip(min=0.0), dim=1),
            "lin_vel_z": lambda: torch.square(s.base_lin_vel[:, 2]),
            "orientation": lambda: torch.sum(torch.square(s.projected_gravity[:, :2]), dim=1),
            "stand_still": lambda: torch.sum(torch.abs(s.dof_pos - s.default_dof_pos), dim=1) * (cmd_norm() < 0.1),
            "stumble": lambda: torch.any(
                torch.norm(F[:, s.feet, :2], dim=2) > 5 * torch.abs(F[:, s.feet, 2]), dim=1),
            "torque_limits": lambda: torch.sum(
                (torch.abs(s.torques) - s.torque_limits * p.soft_torque_limit).clip(min=0.0), dim=1),
            "torques": lambda: torch.sum(torch.square(s.torques), dim=1),
            "tracking_ang_vel": lambda: torch.exp(
                -torch.square(s.commands[:, 2] - s.base_ang_vel[:, 2]) / p.tracking_sigma),
            "tracking_lin_vel": lambda: torch.exp(
                -torch.sum(torch.square(s.commands[:, :2] - s.base_lin_vel[:, :2]), dim=1) / p.tracking_sigma),
        }

    def _compute_reward(self):                                            # legged_robot.py:189-206
        p = self.p
        fn = self._term_values()
        self.rew_buf[:] = 0.0
        for i, name in enumerate(TERMS[:-1]):
            sc = p.reward_scales[i]
            if sc == 0.0:
                continue
            r = fn[name]() * sc
            self.rew_buf += r
            self.episode_sums[name] += r
        if p.only_positive_rewards:
            self.rew_buf[:] = torch.clip(self.rew_buf, min=0.0)
        sc = p.reward_scales[len(TERMS) - 1]
        if sc != 0.0:
            r = (self.reset_buf * ~self.time_out_buf) * sc
            self.rew_buf += r
            self.episode_sums["termination"] += r

    def _reset_idx(self, ids):                                            # legged_robot.py:147-187, anymal.py:56-60
        p = self.p
        if len(ids) == 0:
            return
        if p.terrain_curriculum:                                          # :463-486
            dist = torch.norm(self.root_states[ids, :2] - self.env_origins[ids, :2], dim=1)
            up = dist > p.terrain_length / 2
            down = (dist < torch.norm(self.commands[ids, :2], dim=1) * p.max_episode_length_s * 0.5) * ~up
            self.terrain_levels[ids] += 1 * up - 1 * down
            if self.rng == "torch":
                rnd = torch.randint_like(self.terrain_levels[ids], p.max_terrain_level)
            else:
                rnd = torch.from_numpy(P.randint(p.seed, ids.numpy() + self.off, self._event(),
                                                 P.SITE_TERRAIN, 1, p.max_terrain_level)).squeeze(1)
            self.terrain_levels[ids] = torch.where(self.terrain_levels[ids] >= p.max_terrain_level, rnd,
                                                   torch.clip(self.terrain_levels[ids], 0))
            self.env_origins[ids] = self.terrain_origins[self.terrain_levels[ids], self.terrain_types[ids]]
        self.dof_pos[ids] = self.default_dof_pos * self._uniform(0.5, 1.5, P.SITE_RESET_DOF, ids, p.num_dof)   # :423
        self.dof_vel[ids] = 0.0
        self.root_states[ids] = self.base_init_state                     # :441-449
        self.root_states[ids, :3] += self.env_origins[ids]
        if p.custom_origins:
            self.root_states[ids, :2] += self._uniform(-1.0, 1.0, P.SITE_RESET_XY, ids, 2)
        self.root_states[ids, 7:13] = self._uniform(-0.5, 0.5, P.SITE_RESET_VEL, ids, 6)
        self._resample_commands(ids, P.SITE_CMD_RESET)
        self.last_actions[ids] = 0.0
        self.last_dof_vel[ids] = 0.0
        self.feet_air_time[ids] = 0.0
        self.episode_length_buf[ids] = 0
        self.extras["episode"] = {}
        for k in self.episode_sums:
            self.extras["episode"]["rew_" + k] = torch.mean(self.episode_sums[k][ids]) / p.max_episode_length_s
            self.episode_sums[k][ids] = 0.0
        if p.terrain_curriculum:
            self.extras["episode"]["terrain_level"] = torch.mean(self.terrain_levels.float())
        if getattr(p, "use_curriculum", False):
            self.extras["episode"]["max_command_x"] = self.command_ranges["lin_vel_x"][1]                     # :183-184
        if p.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf
        if p.use_actuator_network:
            D = p.num_dof
            self.sea_hidden_state.view(2, self.N, D, 8)[:, ids] = 0.0
            self.sea_cell_state.view(2, self.N, D, 8)[:, ids] = 0.0

    def reset_idx(self, ids):
        """LeggedRobot.reset_idx called from OUTSIDE step() (legged_robot.py:147-187; BaseTask.reset base_task.py:111-119).  Same
        body as the in-step reset plus `reset_buf[env_ids] = 1` (:170); its draws are keyed by the event
        (external reset count << 40) | common_step_counter, which no env step uses."""
        ids = torch.as_tensor(ids, dtype=torch.long)
        if len(ids) == 0:
            return
        self._ext_resets = getattr(self, "_ext_resets", 0) + 1
        self._ext_event = (self._ext_resets << 40) | self.common_step_counter
        try:
            self._reset_idx(ids)
        finally:
            self._ext_event = None
        self.reset_buf = self.reset_buf.clone()
        self.reset_buf[ids] = True

    def reset(self, physics):
        """BaseTask.reset of the fork (base_task.py:111-119): reset_idx(all) + zero-action step, twice."""
        ids, zero = torch.arange(self.N), torch.zeros(self.N, self.p.num_dof)
        self.reset_idx(ids)
        self.step(zero, physics)
        self.reset_idx(ids)
        obs, priv, _, _, _ = self.step(zero.clone(), physics)
        return obs, priv

    def _compute_observations(self):                                      # legged_robot.py:208-226
        p = self.p
        obs = torch.cat((self.base_lin_vel * p.obs_lin_vel, self.base_ang_vel * p.obs_ang_vel,
                         self.projected_gravity, self.commands[:, :3] * self.commands_scale,
                         (self.dof_pos - self.default_dof_pos) * p.obs_dof_pos, self.dof_vel * p.obs_dof_vel,
                         self.actions), dim=-1)
        if p.measure_heights:
            h = torch.clip(self.root_states[:, 2].unsqueeze(1) - 0.5 - self.measured_heights, -1, 1.0) * p.obs_height
            obs = torch.cat((obs, h), dim=-1)
        if p.add_noise:
            u = self._u(P.SITE_OBS_NOISE, torch.arange(self.N), p.num_obs)
            obs += (2 * u - 1) * self.noise_scale_vec
        self.obs_buf = obs

    # ---- one env step against a physics stand-in (legged_robot.py:80-104) ----------------------
    def step(self, actions, physics):
        """physics.simulate(port) loads the next dof_state sub-frame; physics.refresh(port) loads root+contacts."""
        p = self.p
        self.actions = torch.clip(actions, -p.clip_actions, p.clip_actions)
        for _ in range(p.decimation):
            self.compute_torques(self.actions)
            physics.simulate(self)
        physics.refresh(self)
        self.post_physics_step()
        self.obs_buf = torch.clip(self.obs_buf, -p.clip_observations, p.clip_observations)
        return self.obs_buf, None, self.rew_buf, self.reset_buf, self.extras


class TapePhysics:
    """Replays a legged_gym_dev_b200.synthetic tape into the port's aliased tensors."""

    def __init__(self, tape):
        self.tape, self.frame, self.sub = tape, 0, 0

    def simulate(self, env):
        env.dof_state.copy_(self.tape.dof[self.frame % self.tape.frames, self.sub])
        self.sub += 1

    def refresh(self, env):
        f = self.frame % self.tape.frames
        env.root_states.copy_(self.tape.root[f])
        env.contact_forces.copy_(self.tape.contact[f])
        self.frame += 1
        self.sub = 0
