"""Drop-in `LeggedRobot` / `Anymal` whose per-step work runs in the fused sm_100a kernels.

Mirrors the public surface of the reference classes (legged_gym/envs/base/legged_robot.py:52,
legged_gym/envs/base/base_task.py:38, legged_gym/envs/anymal_c/anymal.py:46): same constructor
arguments, same attribute names and shapes, same `step / reset / reset_idx / get_observations /
get_privileged_observations`, same template-method hook (`_compute_torques`).  What differs is where the
work happens: `step()` is 4 torque launches + ONE fused post-physics launch through the C ABI
(include/b200gym.h), with no host synchronisation (the reference syncs at legged_robot.py:128,156).

There is no CPU path: tensors must live on a CUDA device and the shared library must be built.
"""
import os

import numpy as np
import torch

from . import _lib
from .params import LeggedParams, flatten_legged_cfg, REWARD_TERMS, TERM_ID
from . import synthetic as S

_RES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resources")


class LeggedRobot:
    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless, physics=None, asset=None, seed=0,
                 env_id_offset=0, terrain=None):
        """cfg: LeggedRobotCfg-shaped object (the reference's or legged_gym_dev_b200.configs').
        physics: object with root_states/dof_state/contact_forces + simulate/refresh (see physics.py).
        asset: optional dict of what the asset loader yields (dof_names, feet/penalised/termination body
               indices, dof_pos_limits, dof_vel_limits, torque_limits); defaults to the ANYmal-C layout.
        terrain: optional dict(height_samples int16 [rows,cols], terrain_origins [rows,cols,3],
                 terrain_levels [N], terrain_types [N], env_origins [N,3])."""
        if physics is None:
            raise RuntimeError("LeggedRobot needs a physics backend (legged_gym_dev_b200.physics.ReplayPhysics or an "
                               "Isaac Gym adapter, see INTEGRATION.md)")
        self.cfg, self.sim_params, self.physics_engine = cfg, sim_params, physics_engine
        self.sim_device, self.headless, self.physics = sim_device, headless, physics
        self.device = torch.device(sim_device)
        if self.device.type != "cuda":
            raise RuntimeError("the b200gym step pipeline runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.lib()
        self.height_samples = None
        self.debug_viz = False
        self.init_done = False
        self.viewer = None
        self._timing = None          # bench.py: list of (kind, start_event, end_event) when enabled
        asset = dict(asset or {})
        self.dof_names = list(asset.get("dof_names", S.DOF_NAMES))
        self.num_envs = int(cfg.env.num_envs)
        self.num_obs = int(cfg.env.num_observations)
        self.num_privileged_obs = cfg.env.num_privileged_obs
        self.num_actions = int(cfg.env.num_actions)
        self.num_dof = self.num_dofs = len(self.dof_names)
        self.num_bodies = int(asset.get("num_bodies", S.NUM_BODIES))
        dev = self.device
        self.feet_indices = torch.tensor(asset.get("feet_indices", S.FEET_INDICES), dtype=torch.long, device=dev)
        self.penalised_contact_indices = torch.tensor(asset.get("penalised_contact_indices", S.PENALISED_INDICES),
                                                      dtype=torch.long, device=dev)
        self.termination_contact_indices = torch.tensor(asset.get("termination_contact_indices", S.TERMINATION_INDICES),
                                                        dtype=torch.long, device=dev)
        lim = asset.get("dof_pos_limits")
        self.dof_pos_limits = (torch.as_tensor(lim, dtype=torch.float, device=dev) if lim is not None
                               else torch.tensor([[-3.14, 3.14]] * self.num_dof, dtype=torch.float, device=dev))
        self.dof_vel_limits = torch.as_tensor(asset.get("dof_vel_limits", [20.0] * self.num_dof), dtype=torch.float, device=dev)
        self.torque_limits = torch.as_tensor(asset.get("torque_limits", [80.0] * self.num_dof), dtype=torch.float, device=dev)
        self.env_id_offset = int(env_id_offset)
        self.seed = int(seed)
        self._terrain_in = terrain
        _lib.require_current_device(self.device)
        self._parse_cfg(cfg)
        self._init_buffers()
        self._prepare_reward_function()
        self.init_done = True

    # ------------------------------------------------------------------ configuration (legged_robot.py:819-837)
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
            seed=self.seed)
        p = self.params
        p.num_bodies = self.num_bodies
        if p.measure_heights and p.mesh_type == "none":
            raise NameError("Can't measure height with terrain mesh type 'none'")          # legged_robot.py:892-893
        if p.measure_heights and p.mesh_type != "plane" and hs is None:
            raise RuntimeError("terrain.measure_heights needs terrain['height_samples']")
        self.dt = p.dt
        self.obs_scales = cfg.normalization.obs_scales
        self.max_episode_length_s = p.max_episode_length_s
        self.max_episode_length = p.max_episode_length
        self.push_time, self.max_push_vel = p.push_time, p.max_push_vel
        self.custom_origins = p.custom_origins
        self.command_ranges = dict(lin_vel_x=p.cmd_lin_vel_x, lin_vel_y=p.cmd_lin_vel_y, ang_vel_yaw=p.cmd_ang_vel_yaw,
                                   heading=p.cmd_heading)
        # the fork's command curriculum (legged_robot.py:822-833): nominal ranges scaled by curriculum.commands[state]
        self.nominal_command_ranges = {k: list(v) for k, v in self.command_ranges.items()}
        self.nominal_push_time, self.nominal_max_push_vel = p.push_time, p.max_push_vel
        self.curriculum_state = 0
        if p.use_curriculum:
            self.update_command_curriculum()

    def _curriculum_ranges(self, ind):
        c = self.params.curriculum_commands[ind]
        return {k: [v * c if k != "heading" else v for v in val] for k, val in self.nominal_command_ranges.items()}   # :502-503

    def update_command_curriculum(self):                                  # legged_robot.py:488-506
        p, ind = self.params, self.curriculum_state
        self.command_ranges = self._curriculum_ranges(ind)
        r = self.command_ranges
        p.cmd_lin_vel_x, p.cmd_lin_vel_y, p.cmd_ang_vel_yaw, p.cmd_heading = r["lin_vel_x"], r["lin_vel_y"], r["ang_vel_yaw"], r["heading"]
        self.push_time = self.nominal_push_time * p.curriculum_push_time[ind]                       # :505 (pushes cannot run: params.py)
        pod = getattr(self, "_pod", None)
        if pod is not None:
            self._write_command_ranges(pod, r, reset_only=False)

    @staticmethod
    def _write_command_ranges(pod, r, reset_only):
        for i, k in enumerate(("lin_vel_x", "lin_vel_y", "ang_vel_yaw", "heading")):
            pod.cmd_lo_reset[i], pod.cmd_span_reset[i] = r[k][0], r[k][1] - r[k][0]
            if not reset_only:
                pod.cmd_lo[i], pod.cmd_span[i] = r[k][0], r[k][1] - r[k][0]
        pod.max_command_x = r["lin_vel_x"][1]

    # ------------------------------------------------------------------ buffers (base_task.py:70-79, legged_robot.py:533-603)
    def _init_buffers(self):
        p, N, D, dev = self.params, self.num_envs, self.num_dof, self.device
        z = lambda *s, dtype=torch.float: torch.zeros(*s, dtype=dtype, device=dev)
        self.obs_buf = z(N, self.num_obs)
        self.rew_buf = z(N)
        self.reset_buf = torch.ones(N, dtype=torch.bool, device=dev)
        self._episode_length_buf = z(N, dtype=torch.long)
        self.time_out_buf = z(N, dtype=torch.bool)
        self.privileged_obs_buf = z(N, self.num_privileged_obs) if self.num_privileged_obs is not None else None
        self.common_step_counter = 0
        self.extras = {}
        self.torques = z(N, self.num_actions)
        self.actions = z(N, self.num_actions)
        self.last_actions = z(N, self.num_actions)
        self.last_dof_vel = z(N, D)
        self.last_root_vel = z(N, 6)
        self.commands = z(N, cfg_num_commands(self.cfg))
        self.commands_scale = torch.tensor([p.obs_lin_vel, p.obs_lin_vel, p.obs_ang_vel], device=dev)
        self.feet_air_time = z(N, len(p.feet_indices))
        self.last_contacts = z(N, len(p.feet_indices), dtype=torch.bool)
        self.base_lin_vel, self.base_ang_vel, self.projected_gravity = z(N, 3), z(N, 3), z(N, 3)
        self.gravity_vec = torch.tensor([0.0, 0.0, -1.0], device=dev).repeat(N, 1)
        self.forward_vec = torch.tensor([1.0, 0.0, 0.0], device=dev).repeat(N, 1)
        self.p_gains = torch.tensor(p.p_gains, device=dev)
        self.d_gains = torch.tensor(p.d_gains, device=dev)
        self.default_dof_pos = torch.tensor(p.default_dof_pos, device=dev).unsqueeze(0)
        self.base_init_state = torch.tensor(p.base_init_state, device=dev)
        nv = z(self.num_obs)                                              # legged_robot.py:517-529
        nv[0:3], nv[3:6], nv[6:9] = p.noise_lin_vel, p.noise_ang_vel, p.noise_gravity
        nv[12:24], nv[24:36] = p.noise_dof_pos, p.noise_dof_vel
        if p.measure_heights:
            nv[48:] = p.noise_height
        self.noise_scale_vec, self.add_noise = nv, p.add_noise
        t = self._terrain_in or {}
        self.measured_heights = z(N, p.num_height_points) if p.measure_heights else 0
        self.num_height_points = p.num_height_points
        if t.get("height_samples") is not None:
            self.height_samples = t["height_samples"].to(dev).contiguous()
        eo = t.get("env_origins")
        self.env_origins = eo.to(dev).float().contiguous().clone() if eo is not None else self._grid_origins()
        self.terrain_levels = t["terrain_levels"].to(dev).long().clone() if t.get("terrain_levels") is not None else None
        self.terrain_types = t["terrain_types"].to(dev).long().clone() if t.get("terrain_types") is not None else None
        self.terrain_origins = t["terrain_origins"].to(dev).float().contiguous() if t.get("terrain_origins") is not None else None
        self.max_terrain_level = p.max_terrain_level
        if p.terrain_curriculum and (self.terrain_levels is None or self.terrain_origins is None):
            raise RuntimeError("terrain.curriculum needs terrain_levels / terrain_types / terrain_origins")

    def _grid_origins(self):                                              # legged_robot.py:808-817
        N = self.num_envs
        cols = np.floor(np.sqrt(N))
        rows = np.ceil(N / cols)
        xx, yy = torch.meshgrid(torch.arange(rows), torch.arange(cols), indexing="ij")
        o = torch.zeros(N, 3, device=self.device)
        sp = self.cfg.env.env_spacing
        o[:, 0] = (sp * xx.flatten()[:N]).to(self.device)
        o[:, 1] = (sp * yy.flatten()[:N]).to(self.device)
        return o

    def _prepare_reward_function(self):                                   # legged_robot.py:605-629
        p = self.params
        self.reward_scales = {n: p.reward_scales[TERM_ID[n]] for n in p.active_terms}
        self.reward_names = [n for n in p.active_terms if n != "termination"]
        K = len(p.active_terms)
        self._sums = torch.zeros(max(K, 1), self.num_envs, dtype=torch.float, device=self.device)
        self.episode_sums = {n: self._sums[i] for i, n in enumerate(p.active_terms)}
        sum_row = [-1] * len(REWARD_TERMS)
        for i, n in enumerate(p.active_terms):
            sum_row[TERM_ID[n]] = i
        self._extras_out = torch.zeros(K + 3, dtype=torch.float, device=self.device)
        self._ws_sums = torch.zeros(K + 2, dtype=torch.double, device=self.device)
        self._extras_raw = torch.zeros(K + 2, dtype=torch.double, device=self.device)   # (sums over reset envs, level sum, reset count) of the last step
        self._ws_counter = torch.zeros(4, dtype=torch.int32, device=self.device)
        self._pod = _lib.fill_params(p, K, sum_row, zero_lstm_on_reset=self._has_actuator_state())
        self._ptr_actions, self._ptr_torques = self.actions.data_ptr(), self.torques.data_ptr()
        self._ptr_last_dof_vel = self.last_dof_vel.data_ptr()
        self._stream = None
        # extras["episode"]: 0-d VIEWS of one persistent device buffer that every step rewrites in place (no allocation, no host sync).  A
        # consumer that keeps them across steps must clone them — the runner logs from the per-step raw rows instead (ppo.py, _episode_stats).
        ep = {"rew_" + n: self._extras_out[i] for i, n in enumerate(p.active_terms)}
        if p.terrain_curriculum:
            ep["terrain_level"] = self._extras_out[K]
        if p.use_curriculum:
            ep["max_command_x"] = self._extras_out[K + 2]                 # legged_robot.py:183-184
        self.extras["episode"] = ep
        self.extras["num_resets"] = self._extras_out[K + 1]
        if p.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf

    def _has_actuator_state(self):
        return False

    # ------------------------------------------------------------------ aliased physics tensors (legged_robot.py:545-551)
    @property
    def episode_length_buf(self):
        return self._episode_length_buf

    @episode_length_buf.setter
    def episode_length_buf(self, value):
        # rsl_rl's runner REBINDS this attribute (init_at_random_ep_len); the kernels hold the buffer's address, so copy in place
        self._episode_length_buf.copy_(value)

    @property
    def root_states(self):
        return self.physics.root_states

    @property
    def dof_state(self):
        return self.physics.dof_state

    @property
    def contact_forces(self):
        return self.physics.contact_forces.view(self.num_envs, -1, 3)

    @property
    def dof_pos(self):
        return self.physics.dof_state.view(self.num_envs, self.num_dof, 2)[..., 0]

    @property
    def dof_vel(self):
        return self.physics.dof_state.view(self.num_envs, self.num_dof, 2)[..., 1]

    @property
    def base_quat(self):
        return self.physics.root_states[:, 3:7]

    # ------------------------------------------------------------------ the step (legged_robot.py:80-104)
    def step(self, actions):
        _lib.require_cuda(actions, "actions")
        if actions.dtype != torch.float32:
            actions = actions.float()
        self._stream = torch.cuda.current_stream(self.device).cuda_stream    # looked up once per step
        ph = self.physics
        for i in range(self.params.decimation):
            # the first evaluation also writes the clipped actions (fuses legged_robot.py:86-87)
            ev = self._event_start()
            self._compute_torques(actions, write_clipped=(i == 0))
            self._event_end("torques", ev)
            ph.simulate(self.torques)
        self.post_physics_step()
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def _compute_torques(self, actions, write_clipped=False):             # legged_robot.py:389-413
        rc = self.lib.b200gym_pd_torques(self._pod, actions.data_ptr(), self._ptr_actions if write_clipped else None,
                                         self.physics.dof_state.data_ptr(), self._ptr_last_dof_vel, self._ptr_torques,
                                         self._stream if self._stream is not None else torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            _lib.check(rc, "pd_torques")
        return self.torques

    def _buffers(self):
        """POD of raw device pointers; built once, only the (re-pointable) physics tensors are refreshed per call."""
        ph = self.physics
        b = getattr(self, "_buf_pod", None)
        if b is None:
            b = _lib.LeggedBuffersPOD()
            t = dict(actions=self.actions, torques=self.torques, last_actions=self.last_actions,
                     last_dof_vel=self.last_dof_vel, last_root_vel=self.last_root_vel,
                     commands=getattr(self, "commands", getattr(self, "_commands_unused", None)),
                     feet_air_time=self.feet_air_time, last_contacts=self.last_contacts,
                     episode_length_buf=self.episode_length_buf, reset_buf=self.reset_buf, time_out_buf=self.time_out_buf,
                     rew_buf=self.rew_buf, episode_sums=self._sums, obs_buf=self.obs_buf, base_lin_vel=self.base_lin_vel,
                     base_ang_vel=self.base_ang_vel, projected_gravity=self.projected_gravity,
                     measured_heights=self.measured_heights if torch.is_tensor(self.measured_heights) else None,
                     height_samples=self.height_samples, env_origins=self.env_origins, terrain_levels=self.terrain_levels,
                     terrain_types=self.terrain_types, terrain_origins=self.terrain_origins,
                     lstm_h=getattr(self, "sea_hidden_state", None), lstm_c=getattr(self, "sea_cell_state", None),
                     extras_out=self._extras_out, ws_sums=self._ws_sums, ws_counter=self._ws_counter, extras_raw=self._extras_raw)
            for k, v in t.items():
                if v is not None:
                    _lib.require_cuda(v, k)
                    setattr(b, k, v.data_ptr())
            for k in ("root_states", "dof_state", "contact_forces"):
                _lib.require_cuda(getattr(ph, k), k)
            self._buf_pod = b
        b.root_states, b.dof_state, b.contact_forces = ph.root_states.data_ptr(), ph.dof_state.data_ptr(), ph.contact_forces.data_ptr()
        return b

    def post_physics_step(self):                                          # legged_robot.py:106-134, fused
        self.physics.refresh()
        self.common_step_counter += 1
        p = self.params
        advance = (p.use_curriculum and self.curriculum_state < len(p.curriculum_steps) and
                   self.common_step_counter % p.curriculum_steps[self.curriculum_state] == 0)                 # legged_robot.py:360-363
        if advance:
            # the callback advances the curriculum AFTER this step's periodic resample and BEFORE its reset_idx: inside the fused
            # launch only the reset resample (and extras["episode"]["max_command_x"]) sees the new ranges
            self.curriculum_state += 1
            self._write_command_ranges(self._pod, self._curriculum_ranges(self.curriculum_state), reset_only=True)
        ev = self._event_start()
        st = self._stream if self._stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.b200gym_post_physics(self._pod, self._buffers(), self.common_step_counter, self.env_id_offset, st)
        if rc:
            _lib.check(rc, "post_physics")
        self._event_end("post_physics", ev)
        if advance:
            self.update_command_curriculum()
        self.physics.commit_resets(self.reset_buf)

    def use_device_step_counter(self):
        """Keeps the step counter the kernels see in device memory (advanced by the kernels themselves), so that a
        sequence of env steps can be captured in a CUDA graph and replayed (legged_gym_dev_b200.graphs)."""
        if self.params.use_curriculum:
            raise RuntimeError("the command curriculum advances on the host (legged_robot.py:360-363): its steps cannot be replayed from a CUDA graph")
        if getattr(self, "_step_dev", None) is None:
            self._step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._step_dev.fill_(self.common_step_counter + 1)
        self._buffers().step_counter = self._step_dev.data_ptr()
        return self._step_dev

    def _event_start(self):
        if self._timing is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.device))
        return ev

    def _event_end(self, kind, start):
        if start is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.device))
        self._timing.append((kind, start, ev))

    # ------------------------------------------------------------------ API kept from BaseTask (base_task.py:101-119)
    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def reset_idx(self, env_ids):
        """LeggedRobot.reset_idx called from outside step() (legged_robot.py:147-187): ONE masked launch that redraws the dof / root
        state, resamples commands, clears the per-episode buffers, folds episode_sums into extras["episode"] and sets reset_buf —
        immediately, with time_out_buf untouched, as the reference does.  (The per-step resets live in the fused step kernel.)
        Draws are keyed by an event no env step uses: (external reset count << 40) | common_step_counter."""
        if len(env_ids) == 0:
            return
        if getattr(self, "_reset_mask", None) is None:
            self._reset_mask = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
            self._ext_resets = 0
        self._reset_mask.zero_()
        self._reset_mask[env_ids] = 1
        self._ext_resets += 1
        step = self.common_step_counter if getattr(self, "_step_dev", None) is None else int(self._step_dev.item()) - 1
        rc = self.lib.b200gym_legged_reset_idx(self._pod, self._buffers(), self._reset_mask.data_ptr(), (self._ext_resets << 40) | int(step),
                                               self.env_id_offset, torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            _lib.check(rc, "legged_reset_idx")
        if self.params.traj_mode:                                          # legged_robot_trajectory.py:225 (after the root redraw)
            self.reset_traj(env_ids)
        self.physics.commit_resets(self._reset_mask.view(torch.bool))

    def reset(self):
        """BaseTask.reset of the fork (base_task.py:111-119): reset_idx(all), a zero-action step, and both once more."""
        ids = torch.arange(self.num_envs, device=self.device)
        zero = torch.zeros(self.num_envs, self.num_actions, device=self.device)
        self.reset_idx(ids)
        self.step(zero)
        self.reset_idx(ids)
        obs, priv, _, _, _ = self.step(zero)
        return obs, priv

    def render(self, sync_frame_time=True):
        return None

    # ------------------------------------------------------------------ reward statistics across env shards
    def episode_stats_from_raw(self, raw, total_envs=None):
        """extras["episode"] (legged_robot.py:175-182) from RAW per-step statistics `raw` [..., K + 2] = (per-term sums over the envs
        that reset, sum of terrain levels, number of resets), e.g. `_extras_raw` rows summed over the ranks by ONE all-reduce: the
        means a single process over all envs would log.  Rows without a reset give NaN (nothing to average)."""
        from .sharding import combine_episode_stats
        p = self.params
        return combine_episode_stats(raw, p.active_terms, p.max_episode_length_s, total_envs if total_envs is not None else self.num_envs,
                                     terrain_level=p.terrain_curriculum)


def cfg_num_commands(cfg):
    return int(getattr(getattr(cfg, "commands", None), "num_commands", 4))


class ActuatorNetMixin:
    """anymal.py:46-80 / anymal_trajectory.py:46-80: the actuator-network torque path and its per-(env,dof) LSTM state."""

    def _has_actuator_state(self):
        return bool(self.params.use_actuator_network)

    def _init_buffers(self):
        super()._init_buffers()
        if not self.params.use_actuator_network:
            return
        N, A, dev = self.num_envs, self.num_actions, self.device
        self.sea_hidden_state = torch.zeros(2, N * A, 8, device=dev)
        self.sea_cell_state = torch.zeros(2, N * A, 8, device=dev)
        self.sea_hidden_state_per_env = self.sea_hidden_state.view(2, N, A, 8)
        self.sea_cell_state_per_env = self.sea_cell_state.view(2, N, A, 8)
        self.load_actuator_network(os.path.join(_RES, "anydrive_v3_lstm.npz"))

    def load_actuator_network(self, path):
        """Weights of resources/actuator_nets/anydrive_v3_lstm.pt (anymal.py:52-54) as extracted by
        tools/extract_actuator_net.py; uploaded once into constant memory."""
        w = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in np.load(path).items()}
        self.actuator_weights = w
        order = ["w_ih0", "w_hh0", "b_ih0", "b_hh0", "w_ih1", "w_hh1", "b_ih1", "b_hh1", "w_lin", "b_lin"]
        args = [w[k].ctypes.data for k in order]
        _lib.check(self.lib.b200gym_set_actuator_net(*args, float(w["in_scale"][0]), float(w["in_scale"][1]),
                                                     float(w["out_scale"][0])), "set_actuator_net")

    def _compute_torques(self, actions, write_clipped=False):             # anymal.py:71-78
        if not self.params.use_actuator_network:
            return super()._compute_torques(actions, write_clipped)
        rc = self.lib.b200gym_lstm_torques(self._pod, actions.data_ptr(), self._ptr_actions if write_clipped else None,
                                           self.physics.dof_state.data_ptr(), self.sea_hidden_state.data_ptr(),
                                           self.sea_cell_state.data_ptr(), self._ptr_torques,
                                           self._stream if self._stream is not None else torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            _lib.check(rc, "lstm_torques")
        return self.torques


class Anymal(ActuatorNetMixin, LeggedRobot):
    """legged_gym/envs/anymal_c/anymal.py:46-80."""
