"""Config objects with the attribute surface the step pipeline reads.

The reference keeps nested-class configs (legged_gym/envs/base/legged_robot_config.py:34-279 and the
anymal_c overrides).  The fused path is duck-typed: it accepts those very objects.  Because the reference
cannot be imported on a box without Isaac Gym, this module builds equivalent attribute trees from compact
tables, so `task_registry.make_env("anymal_c_flat")` works standalone.  Values follow the reference files
cited per block; where the fork's shipped config is broken the two missing attributes
are filled in and flagged.
"""
import copy


class Cfg:
    """Attribute tree; dict-valued leaves stay dicts (e.g. stiffness / default_joint_angles)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def clone(self):
        return copy.deepcopy(self)

    def __repr__(self):
        return "Cfg(" + ", ".join(f"{k}={v!r}" for k, v in vars(self).items()) + ")"


def _tree(d):
    return Cfg(**{k: _tree(v) if isinstance(v, dict) and k not in _DICT_LEAVES else copy.deepcopy(v) for k, v in d.items()})


_DICT_LEAVES = {"default_joint_angles", "stiffness", "damping", "terrain_kwargs", "wheel_spindown", "wheel_speed_bounds"}

_GRID_X = [round(-0.8 + 0.1 * i, 1) for i in range(17)]      # legged_robot_config.py:55-57
_GRID_Y = [round(-0.5 + 0.1 * i, 1) for i in range(11)]

_BASE = {
    "env": dict(num_envs=4096, num_observations=235, num_privileged_obs=None, num_actions=12, env_spacing=3.0,
                send_timeouts=True, episode_length_s=20),                                       # :35-42
    "terrain": dict(mesh_type="trimesh", horizontal_scale=0.1, vertical_scale=0.005, border_size=25, curriculum=True,
                    static_friction=1.0, dynamic_friction=1.0, restitution=0.0, measure_heights=True,
                    measured_points_x=_GRID_X, measured_points_y=_GRID_Y, selected=False, terrain_kwargs=None,
                    max_init_terrain_level=5, terrain_length=8.0, terrain_width=8.0, num_rows=10, num_cols=20,
                    terrain_proportions=[0.1, 0.1, 0.35, 0.25, 0.2], slope_treshold=0.75),      # :44-69
    "commands": dict(num_commands=4, resampling_time=10.0, heading_command=True,
                     ranges=dict(lin_vel_x=[-0.0, 0.0], lin_vel_y=[-0.0, 0.0], ang_vel_yaw=[-0, 0],
                                 heading=[-0.0, 0.0])),                                         # :71-83
    "init_state": dict(pos=[0.0, 0.0, 1.0], rot=[0.0, 0.0, 0.0, 1.0], lin_vel=[0.0, 0.0, 0.0], ang_vel=[0.0, 0.0, 0.0],
                       default_joint_angles={"joint_a": 0.0, "joint_b": 0.0}),
    "control": dict(control_type="P", stiffness={"joint_a": 10.0, "joint_b": 15.0}, damping={"joint_a": 1.0, "joint_b": 1.5},
                    action_scale=0.5, decimation=4),
    "asset": dict(file="", name="legged_robot", foot_name="None", penalize_contacts_on=[], terminate_after_contacts_on=[],
                  self_collisions=0),
    "domain_rand": dict(randomize_friction=True, friction_range=[0.5, 1.25], randomize_base_mass=False,
                        added_mass_range=[-1.0, 1.0], push_robots=True, push_interval_s=15, max_push_vel_xy=1.0,
                        max_push_vel=1.0),   # max_push_vel: read at legged_robot.py:827, missing in the fork
    "rewards": dict(scales=dict(termination=-0.0), only_positive_rewards=True, tracking_sigma=0.25, soft_dof_pos_limit=1.0,
                    soft_dof_vel_limit=1.0, soft_torque_limit=1.0, base_height_target=1.0, max_contact_force=100.0),
    "curriculum": dict(use_curriculum=False, curriculum_steps=[100, 200], commands=[0.5, 0.75, 1],
                       push=dict(magnitude=[0.1, 0.5, 1], time=[3, 2, 1])),   # use_curriculum: annotation-only in the fork
    "normalization": dict(obs_scales=dict(lin_vel=2.0, ang_vel=0.25, dof_pos=1.0, dof_vel=0.05, height_measurements=5.0),
                          clip_observations=100.0, clip_actions=100.0),
    "noise": dict(add_noise=True, noise_level=1.0,
                  noise_scales=dict(dof_pos=0.01, dof_vel=1.5, lin_vel=0.1, ang_vel=0.2, gravity=0.05,
                                    height_measurements=0.1)),
    "sim": dict(dt=0.005, substeps=1, gravity=[0.0, 0.0, -9.81], up_axis=1),
}

_PPO = {
    "seed": 1, "runner_class_name": "OnPolicyRunner",
    "policy": dict(init_noise_std=1.0, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128], activation="elu"),
    "algorithm": dict(value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.01,
                      num_learning_epochs=5, num_mini_batches=4, learning_rate=1.0e-3, schedule="adaptive", gamma=0.99,
                      lam=0.95, desired_kl=0.01, max_grad_norm=1.0),                              # :249-261
    "runner": dict(policy_class_name="ActorCritic", algorithm_class_name="PPO", num_steps_per_env=24, max_iterations=1500,
                   save_interval=50, experiment_name="test", run_name="", resume=False, load_run=-1, checkpoint=-1,
                   resume_path=None),
}


def _merge(base, over):
    out = copy.deepcopy(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and k not in _DICT_LEAVES:
            out[k] = _merge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


_ANYMAL_ROUGH = _merge(_BASE, {      # anymal_c/mixed_terrains/anymal_c_rough_config.py:32-86
    "env": dict(num_envs=4096, num_actions=12),
    "init_state": dict(pos=[0.0, 0.0, 0.6], default_joint_angles={
        "LF_HAA": 0.0, "LH_HAA": 0.0, "RF_HAA": -0.0, "RH_HAA": -0.0, "LF_HFE": 0.4, "LH_HFE": -0.4, "RF_HFE": 0.4,
        "RH_HFE": -0.4, "LF_KFE": -0.8, "LH_KFE": 0.8, "RF_KFE": -0.8, "RH_KFE": 0.8}),
    "control": dict(stiffness={"HAA": 80.0, "HFE": 80.0, "KFE": 80.0}, damping={"HAA": 2.0, "HFE": 2.0, "KFE": 2.0},
                    use_actuator_network=True,
                    actuator_net_file="{LEGGED_GYM_ROOT_DIR}/resources/actuator_nets/anydrive_v3_lstm.pt"),
    "asset": dict(file="{LEGGED_GYM_ROOT_DIR}/resources/robots/anymal_c/urdf/anymal_c.urdf", name="anymal_c", foot_name="FOOT",
                  penalize_contacts_on=["SHANK", "THIGH"], terminate_after_contacts_on=["base"], self_collisions=1),
    "domain_rand": dict(randomize_base_mass=True, added_mass_range=[-5.0, 5.0]),
    "rewards": dict(base_height_target=0.5, max_contact_force=500.0, only_positive_rewards=True),
})

_ANYMAL_FLAT = _merge(_ANYMAL_ROUGH, {   # anymal_c/flat/anymal_c_flat_config.py:32-60
    "env": dict(num_observations=48),
    "terrain": dict(mesh_type="plane", measure_heights=False),
    "asset": dict(self_collisions=0),
    "rewards": dict(max_contact_force=350.0, scales=dict(orientation=-5.0, torques=-0.000025, feet_air_time=2.0)),
    "commands": dict(heading_command=False, resampling_time=4.0, ranges=dict(ang_vel_yaw=[-1.5, 1.5])),
    "domain_rand": dict(friction_range=[0.0, 1.5]),
})

# The upstream defaults the fork commented out (legged_robot_config.py:76-79,155-168): the throughput config
# of SURVEY.md §8d cfg 2(b).
UPSTREAM_REWARD_SCALES = dict(tracking_lin_vel=1.0, tracking_ang_vel=0.5, lin_vel_z=-2.0, ang_vel_xy=-0.05, torques=-0.00001,
                              dof_acc=-2.5e-7, feet_air_time=1.0, collision=-1.0, action_rate=-0.01, termination=-0.0)
UPSTREAM_COMMAND_RANGES = dict(lin_vel_x=[-1.0, 1.0], lin_vel_y=[-1.0, 1.0], ang_vel_yaw=[-1, 1], heading=[-3.14, 3.14])


# ---- LeggedRobotTrajectoryCfg family (legged_gym/envs/base/legged_robot_trajectory_config.py:33-231 and the anymal_c
# *_trajectory overrides).  The fork's anymal trajectory configs do not carry every attribute LeggedRobotTrajectory reads
# (they were moved on to the hopper configs): `rewards.tracking_sigma`, `domain_rand.{randomize_rom_distance, max_rom_dist,
# zero_rom_distance_likelihood}`, `curriculum.*`, `trajectory_generator.{dN, prob_stationary}` and an existing weight-sampler
# class are filled in from legged_gym/envs/hopper/flat_trajectory/hopper_trajectory_config.py:113-260 and flagged here.
_TRAJ_BASE = {
    "env": dict(num_envs=4096, num_observations=240, num_privileged_obs=None, num_actions=12, env_spacing=3.0,
                send_timeouts=True, episode_length_s=20),
    "terrain": dict(_BASE["terrain"], slope_threshold=0.75),
    "rom": dict(cls="SingleInt2D", dt=0.1, z_min=[-1e9, -1e9], z_max=[1e9, 1e9], v_min=[-0.35, -0.35], v_max=[0.35, 0.35],
                prob_stationary=1e-4, stationary_duration=1.0),                                    # :71-88
    "trajectory_generator": dict(cls="TrajectoryGenerator", t_samp_cls="UniformSampleHoldDT",
                                 weight_samp_cls="UniformWeightSampler",   # shipped name 'WeightSamplerSampleAndHold' does not exist
                                 N=10, t_low=1, t_high=2, freq_low=0.01, freq_high=2, seed=42, DN=1,
                                 dN=1, prob_stationary=1e-4),                                      # :90-100 (+ dN, prob_stationary)
    "init_state": _BASE["init_state"],
    "control": _BASE["control"],
    "asset": _BASE["asset"],
    "domain_rand": dict(randomize_friction=True, friction_range=[0.5, 1.25], randomize_base_mass=False,
                        added_mass_range=[-1.0, 1.0], push_robots=True, push_interval_s=15, max_push_vel_xy=1.0,
                        max_push_vel=[0.25, 0.25, 0.25, 0.75, 0.75, 0.75], time_between_pushes=[0.5, 10.0],   # :149-158
                        randomize_rom_distance=True, max_rom_dist=[1.0, 1.0], zero_rom_distance_likelihood=0.25),  # filled in
    "rewards": dict(scales=dict(termination=-0.5), only_positive_rewards=False, soft_dof_pos_limit=1.0, soft_dof_vel_limit=1.0,
                    soft_torque_limit=1.0, base_height_target=1.0, max_contact_force=100.0,
                    differential_error=dict(neg_slope=1.0, pos_slope=4.0),
                    reward_weighting=dict(position=1.0, velocity=1.0, orientation=0.3, angular_velocity=0.2, v_perp=0.4),
                    tracking_sigma=0.25),                                                          # :160-194 (+ tracking_sigma)
    "curriculum": dict(use_curriculum=False, curriculum_steps=[2500, 5000]),                       # filled in
    "normalization": dict(obs_scales=dict(lin_vel=2.0, ang_vel=0.25, dof_pos=1.0, dof_vel=0.05, height_measurements=5.0,
                                          trajectory=[1.0, 1.0]), clip_observations=100.0, clip_actions=100.0),
    "noise": _BASE["noise"],
    "sim": _BASE["sim"],
}

_ANYMAL_ROUGH_TRAJ = _merge(_TRAJ_BASE, {   # anymal_c/mixed_terrains_trajectory/anymal_c_rough_trajectory_config.py:32-86
    "env": dict(num_envs=4096, num_actions=12, num_observations=65 + 187),
    "terrain": dict(mesh_type="trimesh", curriculum=False),   # the terrain curriculum needs self.commands (:508): off
    "init_state": _ANYMAL_ROUGH["init_state"],
    "control": _ANYMAL_ROUGH["control"],
    "asset": _ANYMAL_ROUGH["asset"],
    "domain_rand": dict(randomize_base_mass=True, added_mass_range=[-5.0, 5.0]),
    "rewards": dict(base_height_target=0.5, max_contact_force=500.0, only_positive_rewards=False),
})

_ANYMAL_FLAT_TRAJ = _merge(_ANYMAL_ROUGH_TRAJ, {   # anymal_c/flat_trajectory/anymal_c_flat_trajectory_config.py:32-52
    "env": dict(num_observations=65),
    "terrain": dict(mesh_type="plane", measure_heights=False),
    "asset": dict(self_collisions=0),
    "rewards": dict(max_contact_force=350.0, scales=dict(orientation=-5.0, torques=-0.000025, feet_air_time=0.5)),
    "domain_rand": dict(friction_range=[0.0, 1.5]),
})

# every reward term LeggedRobotTrajectory defines (legged_robot_trajectory.py:1000-1110) except stand_still (reads self.commands)
TRAJECTORY_ALL_REWARD_SCALES = dict(
    action_rate=-0.01, ang_vel_xy=-0.05, base_height=-1.0, collision=-1.0, differential_error=2.0, dof_acc=-2.5e-7,
    dof_pos_limits=-10.0, dof_vel=-1e-4, dof_vel_limits=-0.5, feet_air_time=1.0, feet_contact_forces=-0.01, lin_vel_z=-2.0,
    orientation=-5.0, stumble=-0.3, termination=-3.0, torque_limits=-0.02, torques=-1e-5, tracking_rom=6.0)


def anymal_c_rough_trajectory_cfg():
    return _tree(_ANYMAL_ROUGH_TRAJ)


def anymal_c_flat_trajectory_cfg():
    return _tree(_ANYMAL_FLAT_TRAJ)


def anymal_c_flat_trajectory_cfg_ppo():
    return _tree(_merge(_PPO, {"policy": dict(actor_hidden_dims=[128, 64, 32], critic_hidden_dims=[128, 64, 32]),
                               "runner": dict(experiment_name="flat_anymal_c_trajectory", max_iterations=300)}))


def anymal_c_rough_trajectory_cfg_ppo():
    return _tree(_merge(_PPO, {"runner": dict(experiment_name="rough_anymal_trajectory_c")}))


# ---- HopperRoughTrajectoryCfg (legged_gym/envs/hopper/flat_trajectory/hopper_trajectory_config.py:3-260), the configuration of the only
# Hopper class of the fork that can run a full step (HopperTrajectory).  Values the shipped class leaves undefined or unusable are taken from
# the training yaml (deep_tube_learning/configs/rl/hopper_single_int.yaml) and flagged.
_HOPPER_TRAJ = _merge(_TRAJ_BASE, {
    "env": dict(num_envs=4096 * 4, num_observations=38, num_actions=4),                            # :4-7
    "terrain": dict(mesh_type="plane", measure_heights=False, curriculum=False),                   # :9-13
    "init_state": dict(pos=[0.0, 0.0, 0.3], rot=[0.0, 0.0, 0.0, 1.0], lin_vel=[0.0, 0.0, 0.0], ang_vel=[0.0, 0.0, 0.0],
                       default_joint_angles={"foot_slide": 0.0, "wheel1_rotation": 0.0, "wheel2_rotation": 0.0, "wheel3_rotation": 0.0},
                       randomize_yaw=True, default_dof_pos_noise_lower=[-0.02, 0, 0, 0], default_dof_pos_noise_upper=[0.02, 0, 0, 0],
                       default_dof_vel_noise_lower=[-0.1, -100.0, -100.0, -100.0], default_dof_vel_noise_upper=[0.1, 100.0, 100.0, 100.0],
                       default_root_pos_noise_lower=[-0.0, -0.0, -0.05, -0.03, -0.03, -0.03, -0.03],
                       default_root_pos_noise_upper=[0.0, 0.0, 0.05, 0.03, 0.03, 0.03, 0.03],
                       default_root_vel_noise_lower=[-0.05, -0.05, -0.05, -0.2, -0.2, -0.2],
                       default_root_vel_noise_upper=[0.05, 0.05, 0.05, 0.2, 0.2, 0.2]),               # :15-31
    "control": dict(stiffness={"foot_slide": 400.0, "wheel1_rotation": 15.0, "wheel2_rotation": 15.0, "wheel3_rotation": 15.0},
                    damping={"foot_slide": 40.0, "wheel1_rotation": 3.0, "wheel2_rotation": 3.0, "wheel3_rotation": 3.0},
                    wheel_spindown={"wheel1_rotation": 0.1, "wheel2_rotation": 0.1, "wheel3_rotation": 0.1}, foot_pos_des=0.03, action_scale=1,
                    decimation=4, control_type="orientation", zero_action=[1.0, 0.0, 0.0, 0.0], use_actuator_network=False),   # :33-58
    "asset": dict(name="hopper_flat_trajectory", foot_name="foot", penalize_contacts_on=[], terminate_after_contacts_on=["wheel", "torso"],
                  spring_stiffness=11732, spring_damping=50,
                  rot_actuator=[[-0.8165, 0.2511, 0.2511], [-0.0, -0.7643, 0.7643], [-0.5773, -0.5939, -0.5939]],
                  wheel_speed_bounds={"wheel1_rotation": 600, "wheel2_rotation": 600, "wheel3_rotation": 600}, torque_speed_bound_ratio=6),   # :60-90
    "normalization": dict(obs_scales=dict(lin_vel=0.5, ang_vel=0.25, dof_vel=0.01, z_pos=1.0, trajectory=[1.0, 1.0], height_measurements=5.0),
                          clip_observations=100.0, clip_actions=100.0),                             # :92-103
    "noise": dict(add_noise=True, noise_level=1.0, noise_scales=dict(dof_vel=1.5, lin_vel=0.1, ang_vel=0.2, gravity=0.05, z_pos=0.02, quat=0.05,
                                                                     height_measurements=0.1)),     # :105-116
    "domain_rand": dict(push_robots=True, randomize_rom_distance=False, max_rom_dist=[1.0, 1.0],    # max_rom_dist: None as shipped (:131), yaml value
                        zero_rom_distance_likelihood=0.25,
                        spring_properties=dict(randomize_stiffness=True, stiffness_range=[0.9, 1.1], randomize_damping=True, damping_range=[0.9, 1.1],
                                               randomize_setpoint=True, setpoint_range=[0.75, 1.25]),
                        pd_gain_properties=dict(randomize_p_gain=True, p_gain_range=[0.9, 1.1], randomize_d_gain=True, d_gain_range=[0.9, 1.1]),
                        torque_speed_properties=dict(randomize_max_torque=True, max_torque_range=[0.95, 1.05], randomize_max_speed=True,
                                                     max_speed_range=[0.9, 1.1], randomize_slope=True, slope_range=[0.9, 1.1])),   # :118-170
    "rewards": dict(scales=dict(termination=-0.5, collision=-1.0), only_positive_rewards=False, base_height_target=0.55, max_contact_force=100.0,
                    tracking_sigma=0.25, reward_weighting=dict(position=1.0),                        # tracking_sigma / weighting: yaml
                    raibert=dict(Kp=-0.3, Kv=-0.9, Kff=0.0, clip_pos=0.5, clip_vel=1.0, clip_ang=0.2),
                    differential_error=dict(pos_slope=4, neg_slope=1)),                            # :172-190
    "trajectory_generator": dict(weight_samp_cls="UniformWeightSamplerNoRamp"),                    # yaml; the shipped class name does not exist
    "curriculum": dict(use_curriculum=False, curriculum_steps=[2500, 5000]),                       # :214-216
})
# the reward table the Hopper is trained with (deep_tube_learning/configs/rl/hopper_single_int.yaml)
HOPPER_YAML_REWARD_SCALES = dict(termination=-500.0, tracking_rom=6.0, ang_vel_xy=-0.01, orientation=-80.0, torques=-1e-6, dof_acc=-2.5e-8,
                                 unit_quat=-0.01, collision=-1.0, action_rate=-0.01, differential_error=10.0, raibert=-0.1)
HOPPER_DOF_NAMES = ["foot_slide", "wheel1_rotation", "wheel2_rotation", "wheel3_rotation"]
# what the asset loader would report for resources/robots/hopper/urdf/hopper.urdf (bodies: torso, wheel1-3, foot)
HOPPER_ASSET = dict(num_bodies=5, foot_body=4, termination_bodies=[0, 1, 2, 3], penalised_bodies=[], torque_limits=[300.0, 1.5, 1.5, 1.5],
                    dof_pos_limits=[[-0.03, 0.1], [-1e4, 1e4], [-1e4, 1e4], [-1e4, 1e4]], dof_vel_limits=[5.0, 600.0, 600.0, 600.0])


def hopper_flat_trajectory_cfg():
    return _tree(_HOPPER_TRAJ)


def hopper_flat_trajectory_cfg_ppo():
    return _tree(_merge(_PPO, {"policy": dict(actor_hidden_dims=[128, 64, 32], critic_hidden_dims=[128, 64, 32]),
                               "algorithm": dict(entropy_coef=0.01), "runner": dict(experiment_name="hopper_flat_trajectory", max_iterations=1500)}))


def anymal_c_rough_cfg():
    return _tree(_ANYMAL_ROUGH)


def anymal_c_flat_cfg():
    return _tree(_ANYMAL_FLAT)


def anymal_c_rough_cfg_ppo():
    return _tree(_merge(_PPO, {"runner": dict(experiment_name="rough_anymal_c")}))


def anymal_c_flat_cfg_ppo():
    return _tree(_merge(_PPO, {"policy": dict(actor_hidden_dims=[128, 64, 32], critic_hidden_dims=[128, 64, 32]),
                               "runner": dict(experiment_name="flat_anymal_c", max_iterations=300)}))


def with_upstream_rewards(cfg, pd_control=True):
    """cfg 2(b): upstream reward/command tables; PD branch of _compute_torques."""
    cfg = cfg.clone()
    for k, v in UPSTREAM_REWARD_SCALES.items():
        setattr(cfg.rewards.scales, k, v)
    for k, v in UPSTREAM_COMMAND_RANGES.items():
        setattr(cfg.commands.ranges, k, list(v))
    if pd_control:
        cfg.control.use_actuator_network = False
    return cfg


def double_single_int_cfg(num_envs=8192, seed=0, **over):
    """deep_tube_learning/configs/data_generation/double_single_int.yaml:29-87 (+ default_custom.yaml:24-25 pos_max=1e9):
    DoubleInt2D model (dt 0.05) tracked against a SingleInt2D ROM (dt 0.1), horizon N=10."""
    d = dict(pos_max=1e9, vel_max=0.3, acc_max=0.5, vel_max_rom=0.2, model_dt=0.05, rom_dt=0.1, N=10, dN=1, t_low=1, t_high=2,
             freq_low=0.01, freq_high=2, prob_stationary=0.0005, weight_samp_cls="UniformWeightSamplerNoRamp",
             randomize_rom_distance=True, max_rom_distance=[1.0, 1.0], zero_rom_dist_llh=0.25,
             noise_lower=[0.0, 0.0, -0.1, -0.1], noise_upper=[0.0, 0.0, 0.1, 0.1], Kp=10, Kd=10, episode_length_s=20)
    d.update(over)
    pm, vm, am, vr = d["pos_max"], d["vel_max"], d["acc_max"], d["vel_max_rom"]
    return _tree({
        "env": dict(num_envs=num_envs, episode_length_s=d["episode_length_s"], type="custom",
                    model=dict(dt=d["model_dt"], cls="DoubleInt2D", z_min=[-pm, -pm, -vm, -vm], z_max=[pm, pm, vm, vm],
                               v_min=[-am, -am], v_max=[am, am])),
        "rom": dict(cls="SingleInt2D", dt=d["rom_dt"], z_min=[-pm, -pm], z_max=[pm, pm], v_min=[-vr, -vr], v_max=[vr, vr]),
        "trajectory_generator": dict(cls="TrajectoryGenerator", t_samp_cls="UniformSampleHoldDT",
                                     weight_samp_cls=d["weight_samp_cls"], N=d["N"], t_low=d["t_low"], t_high=d["t_high"],
                                     freq_low=d["freq_low"], freq_high=d["freq_high"], seed=seed,
                                     prob_stationary=d["prob_stationary"], dN=d["dN"]),
        "noise": dict(add_noise=False),
        "domain_rand": dict(randomize_rom_distance=d["randomize_rom_distance"], max_rom_distance=d["max_rom_distance"],
                            zero_rom_dist_llh=d["zero_rom_dist_llh"]),
        "init_state": dict(default_noise_lower=d["noise_lower"], default_noise_upper=d["noise_upper"]),
        "controller": dict(Kp=d["Kp"], Kd=d["Kd"]),
    })
