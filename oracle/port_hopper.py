"""CPU restatement (torch, fp32) of the Hopper's torque law, Hopper._compute_torques (legged_gym/envs/hopper/hopper.py:168-237):
contact-switched foot spring / PD, quaternion-log orientation control of the three reaction wheels through the actuator rotation,
spin-down in contact, torque-speed and torque-limit clipping, with the per-env randomised gain / limit multipliers (:343-382).
TEST INFRASTRUCTURE.  Pinned by tests/test_hopper_cpu.py against the unmodified method (oracle/ref_harness.reference_hopper_torques,
with pytorch3d.transforms = oracle/pytorch3d_restated — see that module's header for what that pin does and does not cover) and by
tests/golden/hopper_torques_reference.npz.

Control types: "orientation_spindown" (shipped, hopper_config.py:63) and "orientation" — the two the reference method can run.
Its "w_foot" branch (:195-196) subtracts a [num_envs] from a [num_envs, 1] tensor and fails to assign for num_envs > 1; its "V" / "T"
branches (:223-227) index actions with a [num_envs] against a [3] index tensor and only run when num_envs == 3: not restated
(tests/test_hopper_cpu.py shows the reference raising on them).
"""
import torch

from . import pytorch3d_restated as P3

# hopper_config.py:35-55,76-89 (+ LeggedRobot._init_buffers' per-dof gain vectors; wheel3 has no stiffness / damping entry -> 0)
DEFAULTS = dict(action_scale=1.0, p_gains=[900.0, 15.0, 15.0, 0.0], d_gains=[60.0, 3.0, 3.0, 0.0], kd_spindown=[0.1, 0.1, 0.1],
                foot_pos_des=0.021, spring_stiffness=7000.0, spring_damping=4.0, torque_speed_bound_ratio=6.0,
                rot_actuator=[[-0.8165, 0.2511, 0.2511], [-0.0, -0.7643, 0.7643], [-0.5773, -0.5939, -0.5939]],
                wheel_speed_limits=[600.0, 600.0, 600.0], torque_limits=[300.0, 1.5, 1.5, 1.5])


def hopper_case(num_envs, seed=0, control_type="orientation_spindown", max_angle=2.6, num_bodies=5, foot_body=4, **over):
    """Seeded synthetic state of the Hopper's shape: 4 DOF (foot slide + 3 wheels), quaternion actions, contact on ~half of the envs,
    randomised multipliers in the ranges of hopper_config.py:144-175.  The orientation error angle is drawn in [0, max_angle]."""
    g = torch.Generator().manual_seed(seed)
    N = num_envs
    rn = lambda *s: torch.randn(*s, generator=g)
    ru = lambda lo, hi, *s: (hi - lo) * torch.rand(*s, generator=g) + lo
    d = dict(DEFAULTS)
    d.update(over)
    unit = lambda q: q / torch.linalg.norm(q, dim=-1, keepdim=True)
    q_act = unit(rn(N, 4))                                        # wxyz
    axis = unit(rn(N, 3))
    ang = ru(0.0, max_angle, N, 1)
    dq = torch.cat((torch.cos(ang / 2), torch.sin(ang / 2) * axis), dim=1)
    q_des = P3.quaternion_raw_multiply(q_act, dq) * ru(0.5, 2.0, N, 1)      # the policy's raw (un-normalised) quaternion action
    root = rn(N, 13)
    root[:, 3:7] = q_act[:, [1, 2, 3, 0]]                          # xyzw in root_states
    dof_state = torch.stack((torch.cat((ru(-0.03, 0.05, N, 1), rn(N, 3)), 1), torch.cat((rn(N, 1) * 0.5, rn(N, 3) * 200.0), 1)), dim=-1)
    cf = rn(N, num_bodies, 3)
    cf[:, foot_body, 2] = torch.where(torch.rand(N, generator=g) < 0.5, ru(1.0, 300.0, N), torch.zeros(N))
    case = dict(num_envs=N, control_type=control_type, foot_body=foot_body, action_scale=d["action_scale"],
                torque_speed_bound_ratio=d["torque_speed_bound_ratio"], rot_actuator=d["rot_actuator"],
                dof_state=dof_state, contact_forces=cf, root_states=root, base_ang_vel=rn(N, 3) * 2.0,
                p_gains=torch.tensor(d["p_gains"]), d_gains=torch.tensor(d["d_gains"]), kd_spindown=torch.tensor(d["kd_spindown"]),
                torque_limits=torch.tensor(d["torque_limits"]), wheel_speed_limits=torch.tensor(d["wheel_speed_limits"]),
                p_gain_random=ru(0.9, 1.1, N, 4), d_gain_random=ru(0.9, 1.1, N, 4),
                spring_stiffness=ru(0.9, 1.1, N, 1) * d["spring_stiffness"], spring_damping=ru(0.9, 1.1, N, 1) * d["spring_damping"],
                foot_pos_des=ru(0.75, 1.25, N, 1) * d["foot_pos_des"], torque_speed_bound_ratio_random=ru(0.9, 1.1, N, 1),
                torque_limit_random=ru(0.9, 1.1, N, 4), wheel_limit_random=ru(0.9, 1.1, N, 3))
    return case, q_des


def hopper_torques(case, actions):
    """hopper.py:179-237 per env, masks instead of index lists."""
    c = case
    ct = c["control_type"]
    if ct not in ("orientation", "orientation_spindown"):
        raise ValueError(f"control_type {ct!r}: the reference method runs 'orientation' and 'orientation_spindown' only")
    a = actions * c["action_scale"]
    dof_pos, dof_vel = c["dof_state"][..., 0], c["dof_state"][..., 1]
    foot_pos, foot_vel, wheel_vel = dof_pos[:, 0], dof_vel[:, 0], dof_vel[:, 1:4]
    contact = c["contact_forces"][:, c["foot_body"], 2] > 0.1
    p_gains = c["p_gains"] * c["p_gain_random"]
    d_gains = c["d_gains"] * c["d_gain_random"]
    kd_spin = c["kd_spindown"] * c["d_gain_random"][:, 1:4]
    tq = torch.zeros(c["num_envs"], 4)
    foot = torch.where(contact, torch.zeros_like(foot_pos),                                       # :198-199
                       -p_gains[:, 0] * (foot_pos - c["foot_pos_des"][:, 0]) - d_gains[:, 0] * foot_vel)
    spring = -c["spring_stiffness"][:, 0] * foot_pos - c["spring_damping"][:, 0] * foot_vel      # :201
    tq[:, 0] = torch.where(contact, foot + spring, foot)
    quat_des = a / torch.linalg.norm(a, dim=-1, keepdim=True)                                     # :213
    quat_act = c["root_states"][:, [6, 3, 4, 5]]
    err = P3.quaternion_multiply(P3.quaternion_invert(quat_des), quat_act)
    log_err = P3.so3_log_map(P3.quaternion_to_matrix(err))
    local_tau = -p_gains[:, 1:4] * log_err - d_gains[:, 1:4] * c["base_ang_vel"]                  # :219
    tau = P3.Rotate(torch.tensor(c["rot_actuator"])).transform_points(local_tau)                  # :221
    if "spindown" in ct:                                                                          # :204-206
        tau = torch.where(contact[:, None], -kd_spin * wheel_vel, tau)
    tq[:, 1:4] = tau
    ts_ratio = c["torque_speed_bound_ratio"] * c["torque_speed_bound_ratio_random"]               # :231-237
    t_bound = c["torque_limits"] * c["torque_limit_random"]
    w_bound = c["wheel_speed_limits"] * c["wheel_limit_random"]
    upper = -ts_ratio * t_bound[:, 1:4] / w_bound * (wheel_vel - w_bound)
    lower = -ts_ratio * t_bound[:, 1:4] / w_bound * (wheel_vel + w_bound)
    tq[:, 1:4] = torch.clip(tq[:, 1:4], lower, upper)
    return torch.clip(tq, -t_bound, t_bound), tq


# ---- observations + the Hopper's own reward terms (hopper.py:239-258, 407-430, 448-458) -------------------------------------------
OBS_CFG = dict(z_pos=1.0, lin_vel=0.5, ang_vel=0.25, dof_vel=0.01, clip_observations=100.0, add_noise=True, noise_level=1.0,
               noise_scales=dict(z_pos=0.02, quat=0.05, lin_vel=0.1, ang_vel=0.2, dof_vel=1.5))      # hopper_config.py:92-112


def noise_scale_vec(cfg=OBS_CFG):
    """Hopper._get_noise_scale_vec, :407-430 (measure_heights False)."""
    ns, lv = cfg["noise_scales"], cfg["noise_level"]
    v = torch.zeros(21)
    v[0] = ns["z_pos"] * lv * cfg["z_pos"]
    v[1:5] = ns["quat"] * lv
    v[5:8] = ns["lin_vel"] * lv * cfg["lin_vel"]
    v[8:11] = ns["ang_vel"] * lv * cfg["ang_vel"]
    v[11:14] = ns["dof_vel"] * lv * cfg["dof_vel"]
    return v


def obs_case(num_envs, seed=0):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    root = rn(num_envs, 13)
    root[:, 3:7] = root[:, 3:7] / torch.linalg.norm(root[:, 3:7], dim=1, keepdim=True)
    return dict(num_envs=num_envs, root_states=root, base_lin_vel=rn(num_envs, 3), base_ang_vel=rn(num_envs, 3) * 3.0,
                dof_state=torch.stack((rn(num_envs, 4), rn(num_envs, 4) * 150.0), dim=-1), commands=rn(num_envs, 4) * 0.3,
                actions=rn(num_envs, 4) * torch.rand(num_envs, 1, generator=g) * 2.0, last_dof_vel=rn(num_envs, 4) * 150.0, torques=rn(num_envs, 4) * 2.0)


def hopper_observations(case, cfg=OBS_CFG, seed=0, event=1, env_id_offset=0):
    """compute_observations :239-258 followed by the clip of step :116-117; noise uniforms = Philox (seed, env, event, OBS_NOISE, column)."""
    from . import philox as PH
    c = case
    a = c["actions"].clone()
    a /= torch.linalg.norm(a, dim=-1, keepdim=True)
    a[a[:, 0] < 0, :] *= -1
    scale = torch.tensor([cfg["lin_vel"], cfg["lin_vel"], cfg["ang_vel"]])
    obs = torch.cat((c["root_states"][:, 2][:, None] * cfg["z_pos"], c["root_states"][:, 3:7], c["base_lin_vel"] * cfg["lin_vel"],
                     c["base_ang_vel"] * cfg["ang_vel"], c["dof_state"][:, 1:4, 1] * cfg["dof_vel"], c["commands"][:, :3] * scale, a), dim=-1)
    if cfg["add_noise"]:
        import numpy as np
        u = torch.from_numpy(PH.uniform01(seed, np.arange(c["num_envs"]) + env_id_offset, event, PH.SITE_OBS_NOISE, 21))
        obs += (2 * u - 1) * noise_scale_vec(cfg)
    return torch.clip(obs, -cfg["clip_observations"], cfg["clip_observations"])


def hopper_reward_terms(case, dt):
    """_reward_torque_limits, _reward_dof_acc, _reward_unit_quat (:448-458), raw."""
    c = case
    tl = torch.sum(torch.abs(c["torques"][:, 1:4]), dim=1)
    acc = torch.sum(torch.square((c["last_dof_vel"][:, 1:4] - c["dof_state"][:, 1:4, 1]) / dt), dim=1)
    uq = torch.square(1 - torch.linalg.norm(c["actions"], dim=-1))
    return torch.stack((tl, acc, uq), dim=1)


# ---- HopperTrajectory (hopper_trajectory.py), groundwork for the remaining Hopper rows: oracle only, no kernel yet ---------------------
def hopper_traj_noise_scale_vec(traj_size, cfg=OBS_CFG):
    """HopperTrajectory._get_noise_scale_vec, hopper_trajectory.py:439-468 (measure_heights False): the trajectory block replaces the 3 command
    columns; everything after the wheel velocities carries zero noise."""
    v = torch.zeros(14 + traj_size + 4)
    v[:14] = noise_scale_vec(cfg)[:14]
    return v


def hopper_traj_observations(case, trajectory, trajectory_scale, cfg=OBS_CFG, seed=0, event=1, env_id_offset=0):
    """HopperTrajectory.compute_observations, hopper_trajectory.py:255-282, for a SingleInt2D rom (proj_z = root xy) + the clip of step."""
    from . import philox as PH
    import numpy as np
    c = case
    a = c["actions"].clone()
    a /= torch.linalg.norm(a, dim=-1, keepdim=True)
    a[a[:, 0] < 0, :] *= -1
    mod = trajectory.clone()
    mod[:, :, :2] -= c["root_states"][:, :2][:, None, :2]
    obs = torch.cat((c["root_states"][:, 2][:, None] * cfg["z_pos"], c["root_states"][:, 3:7], c["base_lin_vel"] * cfg["lin_vel"],
                     c["base_ang_vel"] * cfg["ang_vel"], c["dof_state"][:, 1:4, 1] * cfg["dof_vel"],
                     (mod * trajectory_scale).reshape(c["num_envs"], -1), a), dim=-1)
    if cfg["add_noise"]:
        u = torch.from_numpy(PH.uniform01(seed, np.arange(c["num_envs"]) + env_id_offset, event, PH.SITE_OBS_NOISE, obs.shape[1]))
        obs += (2 * u - 1) * hopper_traj_noise_scale_vec(trajectory.shape[1] * trajectory.shape[2], cfg)
    return torch.clip(obs, -cfg["clip_observations"], cfg["clip_observations"])


def hopper_reward_raibert(case, desired_position, desired_velocity, gains):
    """HopperTrajectory._reward_raibert, hopper_trajectory.py:482-505 (SingleInt2D rom): squared distance between the policy's quaternion action
    and the Raibert heuristic's (deep_tube_learning/controllers.py:38-73, oracle/port_controllers.py)."""
    from .isaacgym_restated import quat_rotate_inverse
    from .port_controllers import raibert_policy
    rs = case["root_states"]
    cur_v = quat_rotate_inverse(rs[:, 3:7], rs[:, 7:10])[:, :2]
    rh_obs = torch.cat((desired_position - rs[:, :2], cur_v, desired_velocity, rs[:, 3:7]), dim=1)
    rh = raibert_policy(rh_obs, gains["Kp"], gains["Kv"], gains["K_ff"], gains["clip_pos"], gains["clip_vel"], gains["clip_ang"])
    return torch.sum(torch.square(case["actions"] - rh), dim=1)


# hopper_config.py:15-31 / hopper_trajectory_config (init_state)
RESET_CFG = dict(base_init_state=[0.0, 0.0, 0.5, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0], default_dof_pos=[0.021, 0.0, 0.0, 0.0],
                 dof_pos_noise=([-0.02, 0, 0, 0], [0.02, 0, 0, 0]), dof_vel_noise=([-0.1, 100, 100, 100], [0.1, 100, 100, 100]),
                 root_pos_noise=([-0.0, -0.0, -0.05, -0.03, -0.03, -0.03, -0.03], [0.0, 0.0, 0.05, 0.03, 0.03, 0.03, 0.03]),
                 root_vel_noise=([-0.05, -0.05, -0.05, -0.2, -0.2, -0.2], [0.05, 0.05, 0.05, 0.2, 0.2, 0.2]), randomize_yaw=True,
                 zero_action=[1.0, 0.0, 0.0, 0.0], max_push_vel=[0.5, 0.5, 0.2, 0.3, 0.3, 0.3])


def hopper_traj_reset(state, env_ids, env_origins, cfg=RESET_CFG, seed=0, event=1, env_id_offset=0):
    """HopperTrajectory._reset_dofs + _reset_root_states (hopper_trajectory.py:298-358, the custom_origins == False branch — the other one
    adds a [n, 2] draw to 7 columns and cannot run).  `state` = dict(dof_state [N,4,2], root_states [N,13], actions [N,4]), modified in place."""
    import math
    import numpy as np
    from . import philox as PH
    t = lambda v: torch.tensor(v, dtype=torch.float32)
    ids = np.asarray(env_ids) + env_id_offset
    u = lambda site, ncols: torch.from_numpy(PH.uniform01(seed, ids, event, site, ncols))
    lo, hi = (t(v) for v in cfg["dof_pos_noise"])
    state["dof_state"][env_ids, :, 0] = t(cfg["default_dof_pos"]) + ((hi - lo) * u(PH.SITE_HOP_DOF_POS, 4) + lo)
    lo, hi = (t(v) for v in cfg["dof_vel_noise"])
    state["dof_state"][env_ids, :, 1] = (hi - lo) * u(PH.SITE_HOP_DOF_VEL, 4) + lo
    state["actions"][env_ids, :] = t(cfg["zero_action"])
    rs = state["root_states"]
    rs[env_ids] = t(cfg["base_init_state"])
    rs[env_ids, :3] += env_origins[env_ids]
    lo, hi = (t(v)[2:] for v in cfg["root_pos_noise"])
    rs[env_ids, 2:7] += (hi - lo) * u(PH.SITE_HOP_ROOT_POS, 5) + lo
    rs[env_ids, 3:7] /= torch.linalg.norm(rs[env_ids, 3:7], dim=-1, keepdim=True)
    if cfg["randomize_yaw"]:
        yaw = (math.pi - (-math.pi)) * u(PH.SITE_HOP_YAW, 1) + (-math.pi)
        quat_yaw = P3.matrix_to_quaternion(P3.euler_angles_to_matrix(torch.cat([torch.zeros(len(env_ids), 2), yaw], dim=-1), "XYZ"))
        q_new = P3.quaternion_multiply(rs[:, [6, 3, 4, 5]][env_ids, :], quat_yaw)
        rs[env_ids, 3:7] = q_new[:, [1, 2, 3, 0]]
    lo, hi = (t(v) for v in cfg["root_vel_noise"])
    rs[env_ids, 7:13] = (hi - lo) * u(PH.SITE_HOP_ROOT_VEL, 6) + lo


def hopper_traj_push(state, push_idx, cfg=RESET_CFG, seed=0, event=1, env_id_offset=0):
    """HopperTrajectory._push_robots, hopper_trajectory.py:362-367."""
    import numpy as np
    from . import philox as PH
    mv = torch.tensor(cfg["max_push_vel"], dtype=torch.float32)
    u = torch.from_numpy(PH.uniform01(seed, np.asarray(push_idx) + env_id_offset, event, PH.SITE_HOP_PUSH, 6))
    state["root_states"][push_idx, 7:13] = (mv - (-mv)) * u + (-mv)
