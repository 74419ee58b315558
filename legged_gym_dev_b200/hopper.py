"""The Hopper's torque law behind the reference's own method name (first part).

`HopperActuation` carries the tensors Hopper._compute_torques, compute_observations and the Hopper's own reward terms read, under the names the reference class gives them
(legged_gym/envs/hopper/hopper.py:44-69,343-403): dof_pos / dof_vel views of dof_state, contact_forces, root_states, base_ang_vel,
p_gains / d_gains and their per-env random multipliers, spring_stiffness / spring_damping / foot_pos_des, kd_spindown, the torque and
wheel-speed limits with their multipliers, torques.  `_compute_torques(actions)` (hopper.py:168-237) is ONE launch of
b200gym_hopper_torques and returns the limit-clipped torques, leaving `self.torques` as the reference does (clipped to the
torque-speed envelope only).  `compute_observations()` (hopper.py:239-258 + the clip of step, :116-117; noise uniforms = Philox keyed by
(seed, global env id, common_step_counter)) and `_reward_torque_limits / _reward_dof_acc / _reward_unit_quat` (:448-458) are one launch
each.  The rest of the Hopper env (post_physics_step's callback / push timers, the inherited reward terms and reward assembly, reset_idx with
its domain-randomisation draws, hopper_trajectory.py) is not built yet.

Control types: "orientation_spindown" (shipped, hopper_config.py:63) and "orientation".  The reference's "w_foot" branch (:195-196)
fails on a [num_envs, 1] vs [num_envs] broadcast and its "V" / "T" branches (:223-227) on a [num_envs] vs [3] index broadcast, so
they are rejected here rather than given behaviour the reference never had.
"""
import torch

from . import _lib

# hopper_config.py:35-55,62-63,76-89; asset effort limits from the hopper URDF are an input (torque_limits)
# hopper_config.py:92-112 (normalization + noise); 21 observations (hopper_config.py:6), measure_heights False (:13)
DEFAULT_OBS_CFG = dict(z_pos=1.0, lin_vel=0.5, ang_vel=0.25, dof_vel=0.01, clip_observations=100.0, add_noise=True, noise_level=1.0,
                       noise_scales=dict(z_pos=0.02, quat=0.05, lin_vel=0.1, ang_vel=0.2, dof_vel=1.5))
DEFAULT_CFG = dict(control_type="orientation_spindown", action_scale=1.0, p_gains=[900.0, 15.0, 15.0, 0.0], d_gains=[60.0, 3.0, 3.0, 0.0],
                   kd_spindown=[0.1, 0.1, 0.1], foot_pos_des=0.021, spring_stiffness=7000.0, spring_damping=4.0, torque_speed_bound_ratio=6.0,
                   rot_actuator=[[-0.8165, 0.2511, 0.2511], [-0.0, -0.7643, 0.7643], [-0.5773, -0.5939, -0.5939]],
                   wheel_speed_limits=[600.0, 600.0, 600.0], torque_limits=[300.0, 1.5, 1.5, 1.5])


class HopperActuation:
    num_dof = num_actions = 4

    def __init__(self, num_envs, num_bodies=5, foot_body=4, device="cuda", obs_cfg=None, seed=0, env_id_offset=0, dt=0.02, **cfg):
        c = dict(DEFAULT_CFG)
        unknown = set(cfg) - set(c)
        if unknown:
            raise TypeError(f"unknown Hopper torque settings: {sorted(unknown)}")
        c.update(cfg)
        if c["control_type"] not in ("orientation", "orientation_spindown"):
            raise NameError(f"Unknown controller type: {c['control_type']} (the reference method runs 'orientation' and 'orientation_spindown')")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the b200gym Hopper torque law runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.lib()
        self.num_envs, self.num_bodies, self.control_type = int(num_envs), int(num_bodies), c["control_type"]
        self.feet_indices = torch.tensor([foot_body], device=self.device)
        N, dev = self.num_envs, self.device
        f = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)
        ones = lambda *s: torch.ones(*s, dtype=torch.float32, device=dev)
        zeros = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.p_gains, self.d_gains, self.kd_spindown = f(c["p_gains"]), f(c["d_gains"]), f(c["kd_spindown"])
        self.torque_limits, self.wheel_speed_limits = f(c["torque_limits"]), f(c["wheel_speed_limits"])
        self.torque_speed_bound_ratio, self.action_scale, self.rot_actuator = c["torque_speed_bound_ratio"], c["action_scale"], c["rot_actuator"]
        # state the physics side refreshes (hopper.py reads them as views of the gym tensors)
        self.dof_state = zeros(N, 4, 2)
        self.dof_pos, self.dof_vel = self.dof_state[..., 0], self.dof_state[..., 1]
        self.contact_forces, self.root_states, self.base_ang_vel = zeros(N, self.num_bodies, 3), zeros(N, 13), zeros(N, 3)
        # randomised properties, nominal until randomised (hopper.py:65-69,343-382)
        self.spring_stiffness, self.spring_damping = ones(N, 1) * c["spring_stiffness"], ones(N, 1) * c["spring_damping"]
        self.foot_pos_des = ones(N, 1) * c["foot_pos_des"]
        self.p_gain_random, self.d_gain_random, self.torque_limit_random = ones(N, 4), ones(N, 4), ones(N, 4)
        self.wheel_limit_random, self.torque_speed_bound_ratio_random = ones(N, 3), ones(N, 1)
        self.torques = zeros(N, 4)
        self._clipped = zeros(N, 4)
        p = _lib.HopperTorqueParamsPOD()
        p.num_envs, p.num_bodies, p.foot_body, p.spindown = N, self.num_bodies, int(foot_body), int("spindown" in self.control_type)
        p.action_scale, p.torque_speed_bound_ratio = float(self.action_scale), float(self.torque_speed_bound_ratio)
        p.p_gains[:], p.d_gains[:], p.kd_spindown[:] = c["p_gains"], c["d_gains"], c["kd_spindown"]
        p.wheel_speed_limits[:], p.torque_limits[:] = c["wheel_speed_limits"], c["torque_limits"]
        p.rot_actuator[:] = [float(x) for row in c["rot_actuator"] for x in row]
        self._p = p
        # observations + reward terms (hopper.py:239-258,407-430,448-458)
        oc = dict(DEFAULT_OBS_CFG)
        oc.update(obs_cfg or {})
        self.obs_cfg, self.dt, self.seed, self.env_id_offset, self.common_step_counter = oc, float(dt), int(seed), int(env_id_offset), 0
        self.add_noise = bool(oc["add_noise"])
        self.base_lin_vel, self.commands, self.actions, self.last_dof_vel = zeros(N, 3), zeros(N, 4), zeros(N, 4), zeros(N, 4)
        self.obs_buf, self._terms = zeros(N, 21), zeros(N, 3)
        self.commands_scale = f([oc["lin_vel"], oc["lin_vel"], oc["ang_vel"]])
        ns, lv = oc["noise_scales"], oc["noise_level"]
        nv = [ns["z_pos"] * lv * oc["z_pos"]] + [ns["quat"] * lv] * 4 + [ns["lin_vel"] * lv * oc["lin_vel"]] * 3 + [ns["ang_vel"] * lv * oc["ang_vel"]] * 3 \
            + [ns["dof_vel"] * lv * oc["dof_vel"]] * 3 + [0.0] * 7                                       # _get_noise_scale_vec, :407-430
        self.noise_scale_vec = f(nv)
        op = _lib.HopperObsParamsPOD()
        op.num_envs, op.add_noise = N, int(self.add_noise)
        op.z_pos_scale, op.lin_vel_scale, op.ang_vel_scale, op.dof_vel_scale = oc["z_pos"], oc["lin_vel"], oc["ang_vel"], oc["dof_vel"]
        op.clip_observations = oc["clip_observations"]
        op.commands_scale[:] = [oc["lin_vel"], oc["lin_vel"], oc["ang_vel"]]
        op.noise_scale_vec[:] = self.noise_scale_vec.tolist()
        op.seed_lo, op.seed_hi = self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF
        self._op = op

    def compute_observations(self):
        """hopper.py:239-258 (+ the clip of :116-117) into self.obs_buf; the noise event is self.common_step_counter, as in the legged env."""
        _lib.check(self.lib.b200gym_hopper_observations(self._op, _lib.ptr(self.root_states), _lib.ptr(self.base_lin_vel), _lib.ptr(self.base_ang_vel),
                                                        _lib.ptr(self.dof_state), _lib.ptr(self.commands), _lib.ptr(self.actions), _lib.ptr(self.obs_buf),
                                                        int(self.common_step_counter), self.env_id_offset, _lib.stream_ptr(self.device)),
                   "hopper_observations")
        return self.obs_buf

    def _reward_terms(self):
        _lib.check(self.lib.b200gym_hopper_reward_terms(self.num_envs, self.dt, _lib.ptr(self.torques), _lib.ptr(self.dof_state), _lib.ptr(self.last_dof_vel),
                                                        _lib.ptr(self.actions), _lib.ptr(self._terms), _lib.stream_ptr(self.device)), "hopper_reward_terms")
        return self._terms

    def _reward_torque_limits(self):                                      # hopper.py:448-450
        return self._reward_terms()[:, 0]

    def _reward_dof_acc(self):                                            # :452-454
        return self._reward_terms()[:, 1]

    def _reward_unit_quat(self):                                          # :456-458
        return self._reward_terms()[:, 2]

    def load(self, **tensors):
        """Copies state / randomised-property tensors in by the reference's attribute names (shapes checked)."""
        for name, t in tensors.items():
            dst = getattr(self, name)
            if not torch.is_tensor(dst) or name in ("p_gains", "d_gains", "kd_spindown", "torque_limits", "wheel_speed_limits"):
                raise AttributeError(f"{name} is not a per-env tensor of the Hopper torque law")
            dst.copy_(torch.as_tensor(t, dtype=torch.float32).reshape(dst.shape))

    def _compute_torques(self, actions):
        _lib.require_cuda(actions, "actions")
        if tuple(actions.shape) != (self.num_envs, 4) or actions.dtype != torch.float32:
            raise ValueError(f"actions must be a float32 [{self.num_envs}, 4] tensor (got {tuple(actions.shape)}, {actions.dtype})")
        actions = actions.contiguous()
        b = _lib.HopperTorqueBuffersPOD()
        b.actions = actions.data_ptr()
        for name in ("dof_state", "contact_forces", "root_states", "base_ang_vel", "p_gain_random", "d_gain_random", "torque_limit_random",
                     "wheel_limit_random", "spring_stiffness", "spring_damping", "foot_pos_des", "torque_speed_bound_ratio_random", "torques"):
            t = getattr(self, name)
            if not t.is_contiguous():
                raise RuntimeError(f"{name} must stay contiguous")
            setattr(b, name, t.data_ptr())
        b.torques_clipped = self._clipped.data_ptr()
        _lib.check(self.lib.b200gym_hopper_torques(self._p, b, _lib.stream_ptr(self.device)), "hopper_torques")
        return self._clipped
