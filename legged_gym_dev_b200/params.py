"""Host-side flattening of the reference's nested-class configs into plain numbers.

The reference reads its cfg objects all over the step pipeline (legged_robot.py:819-837 `_parse_cfg`,
:507-530 `_get_noise_scale_vec`, :605-629 `_prepare_reward_function`, :586-603 gains).  The fused kernels
take one POD struct per launch instead; this module builds it once from the *same cfg objects*
(duck-typed attribute access, so the reference's own `AnymalCFlatCfg()` instances work unchanged).
"""
import math
from dataclasses import dataclass, field
from typing import List

import numpy as np

# Reward terms in the order the reference evaluates them: `class_to_dict` walks `dir()` (helpers.py:111-126),
# i.e. alphabetical; `termination` is kept out of the loop and added after clipping (legged_robot.py:203-206).
# The table is the union of LeggedRobot's terms (legged_robot.py:918-1015) and LeggedRobotTrajectory's
# (legged_robot_trajectory.py:1000-1110: + differential_error, tracking_rom; - tracking_lin_vel/ang_vel).
REWARD_TERMS = ("action_rate", "ang_vel_xy", "base_height", "collision", "differential_error", "dof_acc", "dof_pos_limits",
                "dof_vel", "dof_vel_limits", "feet_air_time", "feet_contact_forces", "lin_vel_z", "orientation",
                "stand_still", "stumble", "torque_limits", "torques", "tracking_ang_vel", "tracking_lin_vel",
                "tracking_rom", "termination")
TRAJ_ONLY_TERMS = ("differential_error", "tracking_rom")
# _reward_tracking_lin_vel/_ang_vel do not exist in LeggedRobotTrajectory; _reward_stand_still reads self.commands,
# which that class never creates (legged_robot_trajectory.py:621-622, :1090-1093)
NON_TRAJ_TERMS = ("tracking_lin_vel", "tracking_ang_vel", "stand_still")
TERM_ID = {n: i for i, n in enumerate(REWARD_TERMS)}
NUM_TERMS = len(REWARD_TERMS)
CONTROL_TYPES = {"P": 0, "V": 1, "T": 2}


def cfg_to_dict(obj):
    """Same traversal contract as the reference's class_to_dict (helpers.py:111-126): public attributes
    in dir() order, nested classes/instances recursed, lists element-wise."""
    if not hasattr(obj, "__dict__"):
        return obj
    out = {}
    for key in dir(obj):
        if key.startswith("_"):
            continue
        val = getattr(obj, key)
        if callable(val) and not isinstance(val, type):
            continue
        out[key] = [cfg_to_dict(v) for v in val] if isinstance(val, list) else cfg_to_dict(val)
    return out


@dataclass
class LeggedParams:
    num_envs: int = 0
    num_obs: int = 48
    num_dof: int = 12
    num_bodies: int = 17
    feet_indices: List[int] = field(default_factory=lambda: [4, 8, 12, 16])
    penalised_indices: List[int] = field(default_factory=lambda: [3, 7, 11, 15, 2, 6, 10, 14])
    termination_indices: List[int] = field(default_factory=lambda: [0])
    # timing
    sim_dt: float = 0.005
    decimation: int = 4
    dt: float = 0.02
    max_episode_length: float = 1000.0
    max_episode_length_s: float = 20.0
    # control
    control_type: int = 0
    use_actuator_network: bool = False
    action_scale: float = 0.5
    p_gains: List[float] = field(default_factory=lambda: [0.0] * 12)
    d_gains: List[float] = field(default_factory=lambda: [0.0] * 12)
    default_dof_pos: List[float] = field(default_factory=lambda: [0.0] * 12)
    torque_limits: List[float] = field(default_factory=lambda: [80.0] * 12)
    dof_pos_limits: List[List[float]] = field(default_factory=lambda: [[-1.0, 1.0]] * 12)
    dof_vel_limits: List[float] = field(default_factory=lambda: [20.0] * 12)
    clip_actions: float = 100.0
    clip_observations: float = 100.0
    # observation scales / noise
    obs_lin_vel: float = 2.0
    obs_ang_vel: float = 0.25
    obs_dof_pos: float = 1.0
    obs_dof_vel: float = 0.05
    obs_height: float = 5.0
    add_noise: bool = True
    noise_lin_vel: float = 0.0
    noise_ang_vel: float = 0.0
    noise_gravity: float = 0.0
    noise_dof_pos: float = 0.0
    noise_dof_vel: float = 0.0
    noise_height: float = 0.0
    # commands
    heading_command: bool = False
    resample_steps: int = 200
    cmd_lin_vel_x: List[float] = field(default_factory=lambda: [0.0, 0.0])
    cmd_lin_vel_y: List[float] = field(default_factory=lambda: [0.0, 0.0])
    cmd_ang_vel_yaw: List[float] = field(default_factory=lambda: [0.0, 0.0])
    cmd_heading: List[float] = field(default_factory=lambda: [0.0, 0.0])
    # command / push curriculum of the fork's base env (legged_robot.py:360-363, 488-506; legged_robot_config.py:178-185)
    use_curriculum: bool = False
    curriculum_steps: List[int] = field(default_factory=list)
    curriculum_commands: List[float] = field(default_factory=list)
    curriculum_push_magnitude: List[float] = field(default_factory=list)
    curriculum_push_time: List[float] = field(default_factory=list)
    # pushes
    push_robots: bool = True
    push_time: int = 750
    max_push_vel: float = 1.0
    # rewards
    reward_scales: List[float] = field(default_factory=lambda: [0.0] * NUM_TERMS)   # already * dt
    only_positive_rewards: bool = True
    tracking_sigma: float = 0.25
    soft_dof_vel_limit: float = 1.0
    soft_torque_limit: float = 1.0
    base_height_target: float = 1.0
    max_contact_force: float = 100.0
    # terrain
    mesh_type: str = "plane"
    measure_heights: bool = False
    terrain_curriculum: bool = False
    border_size: float = 25.0
    horizontal_scale: float = 0.1
    vertical_scale: float = 0.005
    terrain_rows: int = 0
    terrain_cols: int = 0
    measured_points_x: List[float] = field(default_factory=list)
    measured_points_y: List[float] = field(default_factory=list)
    max_terrain_level: int = 10
    terrain_num_cols: int = 20
    terrain_length: float = 8.0
    custom_origins: bool = False
    base_init_state: List[float] = field(default_factory=lambda: [0, 0, 1.0, 0, 0, 0, 1.0, 0, 0, 0, 0, 0, 0])
    send_timeouts: bool = True
    seed: int = 0
    # LeggedRobotTrajectory (legged_robot_trajectory.py)
    traj_mode: bool = False
    traj_n: int = 0
    traj_horizon: int = 0
    traj_scale: List[float] = field(default_factory=lambda: [1.0, 1.0, 1.0, 1.0])
    traj_weight: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0, 0.0])
    diff_neg_slope: float = 1.0
    diff_pos_slope: float = 4.0
    time_between_pushes: List[float] = field(default_factory=lambda: [0.5, 10.0])

    @property
    def obs_width(self):
        """Columns before the height block: 48, or 45 + N * rom.n in the trajectory env (legged_robot_trajectory.py:277-287)."""
        return 45 + self.traj_horizon * self.traj_n if self.traj_mode else 48

    @property
    def num_height_points(self):
        return len(self.measured_points_x) * len(self.measured_points_y) if self.measure_heights else 0

    @property
    def active_terms(self):
        """Reward names with non-zero scale, in evaluation order (termination included, as in episode_sums)."""
        return sorted(n for i, n in enumerate(REWARD_TERMS) if self.reward_scales[i] != 0.0)


def flatten_legged_cfg(cfg, sim_dt, dof_names, num_envs=None, feet_indices=None, penalised_indices=None,
                       termination_indices=None, dof_pos_limits=None, dof_vel_limits=None, torque_limits=None,
                       custom_origins=None, terrain_rows=0, terrain_cols=0, seed=0, trajectory=False) -> LeggedParams:
    """cfg: an (instantiated) LeggedRobotCfg-shaped object.  Mirrors _parse_cfg / _init_buffers /
    _prepare_reward_function / _get_noise_scale_vec of the reference (citations in the module docstring)."""
    p = LeggedParams()
    p.num_envs = int(num_envs if num_envs is not None else cfg.env.num_envs)
    p.num_obs = int(cfg.env.num_observations)
    p.num_dof = len(dof_names)
    p.seed = int(seed)
    if feet_indices is not None:
        p.feet_indices = [int(i) for i in feet_indices]
    if penalised_indices is not None:
        p.penalised_indices = [int(i) for i in penalised_indices]
    if termination_indices is not None:
        p.termination_indices = [int(i) for i in termination_indices]
    p.sim_dt = float(sim_dt)
    p.decimation = int(cfg.control.decimation)
    p.dt = p.decimation * p.sim_dt                                    # legged_robot.py:820
    p.max_episode_length_s = cfg.env.episode_length_s
    p.max_episode_length = float(np.ceil(p.max_episode_length_s / p.dt))   # :837
    ctype = cfg.control.control_type
    if ctype not in CONTROL_TYPES:
        raise NameError(f"Unknown controller type: {ctype}")           # legged_robot.py:412
    p.control_type = CONTROL_TYPES[ctype]
    p.use_actuator_network = bool(getattr(cfg.control, "use_actuator_network", False))
    p.action_scale = float(cfg.control.action_scale)
    p.p_gains, p.d_gains, p.default_dof_pos = [], [], []
    for name in dof_names:                                            # :588-602 (substring match, last wins)
        p.default_dof_pos.append(float(cfg.init_state.default_joint_angles[name]))
        kp = kd = 0.0
        for key in cfg.control.stiffness.keys():
            if key in name:
                kp, kd = float(cfg.control.stiffness[key]), float(cfg.control.damping[key])
        p.p_gains.append(kp)
        p.d_gains.append(kd)
    if torque_limits is not None:
        p.torque_limits = [float(v) for v in torque_limits]
    if dof_vel_limits is not None:
        p.dof_vel_limits = [float(v) for v in dof_vel_limits]
    if dof_pos_limits is not None:
        p.dof_pos_limits = [[float(a), float(b)] for a, b in dof_pos_limits]
    p.clip_actions = float(cfg.normalization.clip_actions)
    p.clip_observations = float(cfg.normalization.clip_observations)
    s = cfg.normalization.obs_scales
    p.obs_lin_vel, p.obs_ang_vel, p.obs_dof_pos = float(s.lin_vel), float(s.ang_vel), float(s.dof_pos)
    p.obs_dof_vel, p.obs_height = float(s.dof_vel), float(s.height_measurements)
    n, lvl = cfg.noise.noise_scales, cfg.noise.noise_level            # :517-529
    p.add_noise = bool(cfg.noise.add_noise)
    p.noise_lin_vel = n.lin_vel * lvl * s.lin_vel
    p.noise_ang_vel = n.ang_vel * lvl * s.ang_vel
    p.noise_gravity = n.gravity * lvl
    p.noise_dof_pos = n.dof_pos * lvl * s.dof_pos
    p.noise_dof_vel = n.dof_vel * lvl * s.dof_vel
    p.noise_height = n.height_measurements * lvl * s.height_measurements
    d = cfg.domain_rand
    p.push_robots = bool(d.push_robots)
    p.push_time = int(np.ceil(d.push_interval_s / p.dt))              # :826
    if not trajectory:
        c = cfg.commands
        p.heading_command = bool(c.heading_command)
        p.resample_steps = int(c.resampling_time / p.dt)                  # :348
        r = c.ranges
        p.cmd_lin_vel_x, p.cmd_lin_vel_y = [float(v) for v in r.lin_vel_x], [float(v) for v in r.lin_vel_y]
        p.cmd_ang_vel_yaw, p.cmd_heading = [float(v) for v in r.ang_vel_yaw], [float(v) for v in r.heading]
        mv = getattr(d, "max_push_vel", getattr(d, "max_push_vel_xy", 1.0))
        p.max_push_vel = float(mv[0] if isinstance(mv, (list, tuple)) else mv)   # a list only with the curriculum (:504), where pushes cannot run
        cur = getattr(cfg, "curriculum", None)
        p.use_curriculum = bool(getattr(cur, "use_curriculum", False))   # annotation-only in the fork's base cfg (:179): absent = off
        if p.use_curriculum:
            p.curriculum_steps = [int(v) for v in cur.curriculum_steps]
            p.curriculum_commands = [float(v) for v in cur.commands]
            p.curriculum_push_magnitude = [float(v) for v in cur.push.magnitude]
            p.curriculum_push_time = [float(v) for v in cur.push.time]
            if p.push_robots:
                # update_command_curriculum turns max_push_vel into a list (:504) and _push_robots negates it (:459): the reference
                # raises this TypeError on its first push; raised here at construction instead of mid-run
                raise TypeError("bad operand type for unary -: 'list' (legged_robot.py:459: curriculum.use_curriculum with "
                                "domain_rand.push_robots cannot run in the reference; set push_robots = False)")
    else:   # LeggedRobotTrajectoryCfg has no `commands`; pushes use max_push_vel_xy and per-env timers
        p.traj_mode = True
        p.heading_command, p.resample_steps = False, 1
        p.max_push_vel = float(d.max_push_vel_xy)                      # legged_robot_trajectory.py:489
        p.time_between_pushes = [float(v) for v in d.time_between_pushes]   # :85-88, :175-178
    rw = cfg.rewards
    scales = cfg_to_dict(rw.scales)
    p.reward_scales = [0.0] * NUM_TERMS
    cls_name = "LeggedRobotTrajectory" if trajectory else "LeggedRobot"
    for name, v in scales.items():                                    # :610-615
        missing = name not in TERM_ID or name in (NON_TRAJ_TERMS[:2] if trajectory else TRAJ_ONLY_TERMS)
        if missing:
            if v != 0:
                raise AttributeError(f"'{cls_name}' object has no attribute '_reward_{name}'")
            continue
        if trajectory and name == "stand_still" and v != 0:
            raise AttributeError(f"'{cls_name}' object has no attribute 'commands'")
        p.reward_scales[TERM_ID[name]] = float(v) * p.dt if v != 0 else 0.0
    p.only_positive_rewards = bool(rw.only_positive_rewards)
    p.tracking_sigma = float(rw.tracking_sigma)
    if trajectory:
        de = rw.differential_error
        p.diff_neg_slope, p.diff_pos_slope = float(de.neg_slope), float(de.pos_slope)       # :1107-1108
        p.traj_scale = [float(v) for v in s.trajectory] + [1.0] * (4 - len(s.trajectory))   # :626-630
    p.soft_dof_vel_limit, p.soft_torque_limit = float(rw.soft_dof_vel_limit), float(rw.soft_torque_limit)
    p.base_height_target, p.max_contact_force = float(rw.base_height_target), float(rw.max_contact_force)
    t = cfg.terrain
    p.mesh_type = t.mesh_type
    p.measure_heights = bool(t.measure_heights)
    p.terrain_curriculum = bool(t.curriculum) and t.mesh_type in ("heightfield", "trimesh")   # :834-835
    p.border_size, p.horizontal_scale, p.vertical_scale = float(t.border_size), float(t.horizontal_scale), float(t.vertical_scale)
    p.terrain_rows, p.terrain_cols = int(terrain_rows), int(terrain_cols)
    p.measured_points_x = [float(v) for v in t.measured_points_x]
    p.measured_points_y = [float(v) for v in t.measured_points_y]
    p.max_terrain_level = int(t.num_rows)                             # :804
    p.terrain_num_cols = int(t.num_cols)
    p.terrain_length = float(t.terrain_length)
    p.custom_origins = bool(custom_origins if custom_origins is not None
                            else t.mesh_type in ("heightfield", "trimesh"))          # :794-808
    isl = cfg.init_state
    p.base_init_state = [float(v) for v in (list(isl.pos) + list(isl.rot) + list(isl.lin_vel) + list(isl.ang_vel))]
    p.send_timeouts = bool(cfg.env.send_timeouts)
    if not trajectory and p.measure_heights and p.num_obs != 48 + p.num_height_points:
        raise ValueError(f"num_observations {p.num_obs} != 48 + {p.num_height_points} height points")
    return p
