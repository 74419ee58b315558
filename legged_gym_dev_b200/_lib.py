"""ctypes binding of libb200gym.so (the C ABI declared in include/b200gym.h).

The library is built in-tree by `make` / `__graft_entry__.build()`.  There is NO fallback: if the shared
object is missing or a call fails, a RuntimeError is raised (BASELINE.json north_star: "no CPU fallback").
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200GYM_LIB") or os.path.join(_HERE, "libb200gym.so")   # env override: A/B of kernel variants

NUM_DOF, NUM_FEET, NUM_PEN, MAX_TERM, NUM_TERMS, MAX_POINTS = 12, 4, 8, 4, 21, 32

f32, i32, u32 = C.c_float, C.c_int32, C.c_uint32
vp = C.c_void_p


class LeggedParamsPOD(C.Structure):
    _fields_ = [
        ("num_envs", i32), ("num_obs", i32), ("num_bodies", i32), ("num_heights", i32),
        ("feet_idx", i32 * NUM_FEET), ("pen_idx", i32 * NUM_PEN), ("term_idx", i32 * MAX_TERM), ("num_term", i32),
        ("control_type", i32),
        ("action_scale", f32), ("sim_dt", f32), ("dt", f32), ("clip_actions", f32), ("clip_obs", f32),
        ("p_gains", f32 * NUM_DOF), ("d_gains", f32 * NUM_DOF), ("default_dof_pos", f32 * NUM_DOF),
        ("torque_limits", f32 * NUM_DOF), ("dof_pos_lo", f32 * NUM_DOF), ("dof_pos_hi", f32 * NUM_DOF),
        ("dof_vel_limits", f32 * NUM_DOF),
        ("obs_lin_vel", f32), ("obs_ang_vel", f32), ("obs_dof_pos", f32), ("obs_dof_vel", f32), ("obs_height", f32),
        ("add_noise", i32),
        ("noise_lin_vel", f32), ("noise_ang_vel", f32), ("noise_gravity", f32), ("noise_dof_pos", f32),
        ("noise_dof_vel", f32), ("noise_height", f32),
        ("heading_command", i32), ("resample_steps", i32),
        ("cmd_lo", f32 * 4), ("cmd_span", f32 * 4), ("cmd_lo_reset", f32 * 4), ("cmd_span_reset", f32 * 4), ("max_command_x", f32),
        ("push_robots", i32), ("push_time", i32), ("push_lo", f32), ("push_span", f32),
        ("max_episode_length", f32), ("max_episode_length_s", f32),
        ("reward_scale", f32 * NUM_TERMS), ("sum_row", i32 * NUM_TERMS),
        ("num_sum_rows", i32), ("only_positive", i32),
        ("tracking_sigma", f32), ("soft_dof_vel_limit", f32), ("soft_torque_limit", f32),
        ("base_height_target", f32), ("max_contact_force", f32),
        ("mesh_plane", i32), ("terrain_curriculum", i32), ("terrain_rows", i32), ("terrain_cols", i32),
        ("max_terrain_level", i32), ("terrain_num_cols", i32),
        ("border_size", f32), ("horizontal_scale", f32), ("vertical_scale", f32), ("half_env_length", f32),
        ("n_px", i32), ("n_py", i32), ("points_x", f32 * MAX_POINTS), ("points_y", f32 * MAX_POINTS),
        ("custom_origins", i32), ("zero_lstm_on_reset", i32),
        ("base_init_state", f32 * 13),
        ("seed_lo", u32), ("seed_hi", u32),
        ("traj_mode", i32), ("traj_n", i32), ("traj_horizon", i32),
        ("traj_scale", f32 * 4), ("traj_weight", f32 * 4),
        ("diff_neg_slope", f32), ("diff_pos_slope", f32), ("push_t_lo", f32), ("push_t_span", f32),
    ]


_BUF_FIELDS = ["root_states", "dof_state", "contact_forces", "actions", "torques", "last_actions", "last_dof_vel",
               "last_root_vel", "commands", "feet_air_time", "last_contacts", "episode_length_buf", "reset_buf",
               "time_out_buf", "rew_buf", "episode_sums", "obs_buf", "base_lin_vel", "base_ang_vel",
               "projected_gravity", "measured_heights", "height_samples", "env_origins", "terrain_levels",
               "terrain_types", "terrain_origins", "lstm_h", "lstm_c", "extras_out", "ws_sums", "ws_counter", "step_counter",
               "trajectory", "prev_error", "time_until_next_push", "extras_raw"]


class LeggedBuffersPOD(C.Structure):
    _fields_ = [(n, vp) for n in _BUF_FIELDS]


class RomParamsPOD(C.Structure):
    _fields_ = [
        ("num_envs", i32), ("model_type", i32), ("rom_type", i32), ("window", i32), ("dN", i32), ("horizon", i32),
        ("model_dt", f32), ("rom_dt", f32), ("dt_loop", f32),
        ("model_z_min", f32 * 4), ("model_z_max", f32 * 4), ("model_v_min", f32 * 2), ("model_v_max", f32 * 2),
        ("rom_z_min", f32 * 4), ("rom_z_max", f32 * 4), ("rom_v_min", f32 * 2), ("rom_v_max", f32 * 2),
        ("t_low", f32), ("t_span", f32), ("freq_low", f32), ("freq_high", f32), ("prob_stationary", f32),
        ("weight_zero_col", i32), ("randomize_rom_distance", i32),
        ("max_rom_distance", f32 * 4), ("zero_rom_dist_llh", f32), ("noise_lower", f32 * 4), ("noise_upper", f32 * 4),
        ("Kp", f32), ("Kd", f32), ("seed_lo", u32), ("seed_hi", u32),
        ("gen_kind", i32), ("gen_c", f32 * 4), ("gen_v", f32 * 4),
    ]


class RomFamilyParamsPOD(C.Structure):
    _fields_ = [
        ("num_envs", i32), ("rom_type", i32), ("window", i32), ("dN", i32), ("rom_dt", f32), ("dt_loop", f32),
        ("z_min", f32 * 8), ("z_max", f32 * 8), ("v_min", f32 * 4), ("v_max", f32 * 4),
        ("t_low", f32), ("t_span", f32), ("freq_low", f32), ("freq_high", f32), ("prob_stationary", f32),
        ("weight_zero_col", i32), ("seed_lo", u32), ("seed_hi", u32),
    ]


class HopperTorqueParamsPOD(C.Structure):
    _fields_ = [("num_envs", i32), ("num_bodies", i32), ("foot_body", i32), ("spindown", i32), ("action_scale", f32),
                ("torque_speed_bound_ratio", f32), ("p_gains", f32 * 4), ("d_gains", f32 * 4), ("kd_spindown", f32 * 3),
                ("wheel_speed_limits", f32 * 3), ("torque_limits", f32 * 4), ("rot_actuator", f32 * 9), ("ang_vel_from_root", i32), ("pad", i32)]


class HopperObsParamsPOD(C.Structure):
    _fields_ = [("num_envs", i32), ("add_noise", i32), ("z_pos_scale", f32), ("lin_vel_scale", f32), ("ang_vel_scale", f32), ("dof_vel_scale", f32),
                ("clip_observations", f32), ("commands_scale", f32 * 3), ("noise_scale_vec", f32 * 21), ("seed_lo", u32), ("seed_hi", u32)]


HOPPER_NUM_TERMS = 20


class HopperEnvParamsPOD(C.Structure):
    _fields_ = [("num_envs", i32), ("num_bodies", i32), ("foot_body", i32), ("num_term", i32), ("num_pen", i32), ("num_sum_rows", i32),
                ("term_idx", i32 * 8), ("pen_idx", i32 * 8), ("push_robots", i32), ("only_positive", i32), ("add_noise", i32), ("randomize_yaw", i32),
                ("dt", f32), ("push_dt", f32), ("max_episode_length", f32), ("max_episode_length_s", f32),
                ("push_t_lo", f32), ("push_t_span", f32), ("max_push_vel", f32 * 6),
                ("reward_scale", f32 * HOPPER_NUM_TERMS), ("sum_row", i32 * HOPPER_NUM_TERMS),
                ("tracking_sigma", f32), ("soft_dof_vel_limit", f32), ("base_height_target", f32), ("max_contact_force", f32), ("traj_weight", f32 * 2),
                ("diff_neg_slope", f32), ("diff_pos_slope", f32), ("raibert", f32 * 6),
                ("dof_pos_lo", f32 * 4), ("dof_pos_hi", f32 * 4), ("dof_vel_limits", f32 * 4),
                ("default_dof_pos", f32 * 4), ("dof_pos_noise_lo", f32 * 4), ("dof_pos_noise_span", f32 * 4), ("dof_vel_noise_lo", f32 * 4),
                ("dof_vel_noise_span", f32 * 4),
                ("base_init_state", f32 * 13), ("root_pos_noise_lo", f32 * 5), ("root_pos_noise_span", f32 * 5), ("root_vel_noise_lo", f32 * 6),
                ("root_vel_noise_span", f32 * 6), ("zero_action", f32 * 4),
                ("z_pos_scale", f32), ("lin_vel_scale", f32), ("ang_vel_scale", f32), ("dof_vel_scale", f32), ("clip_obs", f32), ("traj_scale", f32 * 2),
                ("noise_scale_vec", f32 * 14), ("seed_lo", u32), ("seed_hi", u32)]


_HOPPER_ENV_FIELDS = ["root_states", "dof_state", "contact_forces", "actions", "torques", "last_actions", "last_dof_vel", "last_root_vel", "base_lin_vel",
                      "base_ang_vel", "projected_gravity", "feet_air_time", "last_contacts", "episode_length_buf", "reset_buf", "time_out_buf", "rew_buf",
                      "episode_sums", "obs_buf", "trajectory", "gen_v", "prev_error", "time_until_next_push", "env_origins", "extras_out", "ws_sums",
                      "push_flag"]


class HopperEnvBuffersPOD(C.Structure):
    _fields_ = [(n, vp) for n in _HOPPER_ENV_FIELDS]


_HOPPER_FIELDS = ["actions", "dof_state", "contact_forces", "root_states", "base_ang_vel", "p_gain_random", "d_gain_random", "torque_limit_random",
                  "wheel_limit_random", "spring_stiffness", "spring_damping", "foot_pos_des", "torque_speed_bound_ratio_random", "torques",
                  "torques_clipped"]


class HopperTorqueBuffersPOD(C.Structure):
    _fields_ = [(n, vp) for n in _HOPPER_FIELDS]


_ROM_FIELDS = ["root_states", "trajectory", "v_trajectory", "v", "t", "k", "t_final", "weights", "sample_hold_input",
               "extreme_input", "ramp_v_start", "ramp_v_end", "ramp_t_start", "sin_mag", "sin_freq", "sin_off", "sin_mean",
               "stationary_inds", "rng_ctr", "env_trajectory", "obs", "center"]


class RomStatePOD(C.Structure):
    _fields_ = [(n, vp) for n in _ROM_FIELDS]


class MlpParamsPOD(C.Structure):
    _fields_ = [("batch", i32), ("num_layers", i32), ("in_dim", i32), ("in_stride", i32), ("out_dim", i32), ("pad", i32),
                ("dims", i32 * 7)]


class PpoLossParamsPOD(C.Structure):
    _fields_ = [("batch", i32), ("num_actions", i32), ("use_clipped_value_loss", i32), ("pad", i32),
                ("clip_param", f32), ("value_loss_coef", f32), ("entropy_coef", f32), ("inv_global_batch", f32)]


class GemmProblemPOD(C.Structure):
    _fields_ = [("a", vp), ("b", vp), ("out", vp), ("aux", vp), ("bias", vp), ("mode", i32), ("flags", i32),
                ("m", i32), ("n", i32), ("k", i32), ("lda", i32), ("ldb", i32), ("ldo", i32), ("ldaux", i32),
                ("m_real", i32), ("n_real", i32), ("splits", i32), ("scale", f32)]


GEMM_FWD, GEMM_DGRAD, GEMM_WGRAD = 0, 1, 2
PACK_MAX = 16


CHAIN_MAX_LAYERS = 6


class ChainNetPOD(C.Structure):
    _fields_ = [("x", vp), ("w16", vp), ("flat_param", vp), ("h", vp * CHAIN_MAX_LAYERS), ("dz", vp * CHAIN_MAX_LAYERS), ("out", vp),
                ("w_off", C.c_int64 * CHAIN_MAX_LAYERS), ("b_off", C.c_int64 * CHAIN_MAX_LAYERS),
                ("kp", i32 * CHAIN_MAX_LAYERS), ("np", i32 * CHAIN_MAX_LAYERS), ("n_real", i32 * CHAIN_MAX_LAYERS),
                ("num_layers", i32), ("ldx", i32), ("x32", vp), ("flat_grad", vp), ("w32_off", C.c_int64 * CHAIN_MAX_LAYERS),
                ("k_real", i32 * CHAIN_MAX_LAYERS), ("ldx32", i32), ("pad", i32)]   # pad = w_layout (0 row-major, 1 chunk-major)


class OptParamsPOD(C.Structure):
    _fields_ = [("n", C.c_int64), ("count", C.c_double), ("adaptive", i32), ("pad", i32), ("desired_kl", f32), ("max_grad_norm", f32),
                ("beta1", f32), ("beta2", f32), ("eps", f32), ("pad2", f32)]


class PackEntryPOD(C.Structure):
    _fields_ = [("src_off", C.c_int64), ("dst_off", C.c_int64), ("elem_end", C.c_int64), ("rows", i32), ("cols", i32), ("ld", i32),
                ("layout", i32)]


class PackTablePOD(C.Structure):
    _fields_ = [("e", PackEntryPOD * PACK_MAX), ("total", C.c_int64), ("n", i32), ("pad", i32)]


class PeerPtrsPOD(C.Structure):
    _fields_ = [("ptr", vp * 16)]


class PeerBasesPOD(C.Structure):
    _fields_ = [("base", vp * 16)]


OPT_MAX_CTAS = 160
PEER_WS_BYTES = 32 + 8 * OPT_MAX_CTAS


_lib = None


def lib():
    """Loads the shared library once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `make` or `python -c 'import __graft_entry__ as g; "
                           f"g.build()'` — the b200gym path has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.b200gym_version.restype = C.c_int
    L.b200gym_last_error.restype = C.c_char_p
    L.b200gym_sizeof.restype = C.c_int
    L.b200gym_sizeof.argtypes = [C.c_char_p]
    pp = C.POINTER(LeggedParamsPOD)
    L.b200gym_pd_torques.argtypes = [pp, vp, vp, vp, vp, vp, vp]
    L.b200gym_set_actuator_net.argtypes = [vp] * 10 + [f32, f32, f32]
    L.b200gym_lstm_torques.argtypes = [pp, vp, vp, vp, vp, vp, vp, vp]
    L.b200gym_post_physics.argtypes = [pp, C.POINTER(LeggedBuffersPOD), C.c_uint64, C.c_int64, vp]
    L.b200gym_legged_reset_idx.argtypes = [pp, C.POINTER(LeggedBuffersPOD), vp, C.c_uint64, C.c_int64, vp]
    L.b200gym_legged_reset_idx.restype = C.c_int
    L.b200gym_hopper_post_physics.argtypes = [C.POINTER(HopperEnvParamsPOD), C.POINTER(HopperEnvBuffersPOD), C.c_uint64, C.c_int64, vp]
    L.b200gym_hopper_post_physics.restype = C.c_int
    L.b200gym_hopper_reset_idx.argtypes = [C.POINTER(HopperEnvParamsPOD), C.POINTER(HopperEnvBuffersPOD), vp, C.c_uint64, C.c_int64, vp]
    L.b200gym_hopper_reset_idx.restype = C.c_int
    for name in ("b200gym_pd_torques", "b200gym_set_actuator_net", "b200gym_lstm_torques", "b200gym_post_physics"):
        getattr(L, name).restype = C.c_int
    rp, rs = C.POINTER(RomParamsPOD), C.POINTER(RomStatePOD)
    L.b200gym_rom_init.argtypes = [rp, rs, C.c_int64, vp]
    L.b200gym_rom_step.argtypes = [rp, rs, vp, vp, C.c_int64, vp]
    L.b200gym_rom_reset.argtypes = [rp, rs, vp, C.c_int64, vp]
    L.b200gym_rom_reset_from_root.argtypes = [rp, rs, vp, vp, C.c_int32, vp, C.c_int64, vp]
    L.b200gym_rom_tracking_policy.argtypes = [rp, vp, vp, vp]
    L.b200gym_raibert_policy.argtypes = [vp, C.c_int32, C.c_int64, f32, f32, f32, f32, f32, f32, vp, vp]
    L.b200gym_raibert_policy.restype = C.c_int
    L.b200gym_rom_rollout.argtypes = [rp, rs, vp, C.c_int32, vp, vp, vp, vp, vp, C.c_int64, vp]
    for name in ("b200gym_rom_init", "b200gym_rom_step", "b200gym_rom_reset", "b200gym_rom_reset_from_root", "b200gym_rom_tracking_policy",
                 "b200gym_rom_rollout"):
        getattr(L, name).restype = C.c_int
    L.b200gym_hopper_torques.argtypes = [C.POINTER(HopperTorqueParamsPOD), C.POINTER(HopperTorqueBuffersPOD), vp]
    L.b200gym_hopper_torques.restype = C.c_int
    L.b200gym_hopper_observations.argtypes = [C.POINTER(HopperObsParamsPOD), vp, vp, vp, vp, vp, vp, vp, C.c_uint64, C.c_int64, vp]
    L.b200gym_hopper_observations.restype = C.c_int
    L.b200gym_hopper_reward_terms.argtypes = [C.c_int32, f32, vp, vp, vp, vp, vp, vp]
    L.b200gym_hopper_reward_terms.restype = C.c_int
    fp = C.POINTER(RomFamilyParamsPOD)
    L.b200gym_romfam_f.argtypes = [C.c_int32, f32, vp, vp, vp, C.c_int64, vp]
    L.b200gym_romfam_des_pose_vel.argtypes = [C.c_int32, vp, vp, vp, vp, C.c_int64, vp]
    L.b200gym_romfam_input_bounds.argtypes = [fp, vp, vp, vp, vp, vp, C.c_int64, vp]
    L.b200gym_romfam_proj_z.argtypes = [C.c_int32, vp, vp, C.c_int64, vp]
    L.b200gym_romfam_gen_init.argtypes = [fp, rs, C.c_int64, vp]
    L.b200gym_romfam_gen_reset.argtypes = [fp, rs, vp, vp, C.c_int64, vp]
    L.b200gym_romfam_gen_step.argtypes = [fp, rs, vp, C.c_int64, vp]
    L.b200gym_romfam_gen_input.argtypes = [fp, rs, vp, vp, vp, C.c_int64, vp]
    for name in ("b200gym_romfam_f", "b200gym_romfam_des_pose_vel", "b200gym_romfam_input_bounds", "b200gym_romfam_proj_z",
                 "b200gym_romfam_gen_init", "b200gym_romfam_gen_reset", "b200gym_romfam_gen_step", "b200gym_romfam_gen_input"):
        getattr(L, name).restype = C.c_int
    f64 = C.c_double
    L.b200gym_gae_returns.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int32, C.c_int32, f32, f32, vp]
    L.b200gym_adv_normalize.argtypes = [vp, vp, C.c_int64, vp]
    L.b200gym_gather_rows.argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.c_int32, vp, C.c_int64, vp]
    L.b200gym_ppo_loss.argtypes = [C.POINTER(PpoLossParamsPOD)] + [vp] * 15
    L.b200gym_grad_sumsq.argtypes = [vp, C.c_int64, f32, vp, vp]
    L.b200gym_clip_adam.argtypes = [vp, vp, vp, vp, C.c_int64, f32, vp, f32, vp, f32, f32, f32, C.c_int32, vp]
    L.b200gym_adaptive_lr.argtypes = [vp, f64, f32, vp, vp]
    L.b200gym_grad_reduce_peers.argtypes = [C.POINTER(PeerPtrsPOD), C.c_int32, vp, C.c_int64, C.c_int64, vp, vp]
    L.b200gym_grad_reduce_peers.restype = C.c_int
    L.b200gym_adam_prepare.argtypes = [vp, vp, vp]
    L.b200gym_clip_adam_dev.argtypes = [vp, vp, vp, vp, C.c_int64, f32, vp, f32, vp, f32, f32, f32, vp, vp]
    L.b200gym_debug_mlp_trace.argtypes = [vp]
    L.b200gym_debug_set_lstm_variant.argtypes = [C.c_int]
    L.b200gym_debug_set_lstm_variant.restype = C.c_int
    for name in ("b200gym_gae_returns", "b200gym_adv_normalize", "b200gym_gather_rows", "b200gym_ppo_loss", "b200gym_grad_sumsq",
                 "b200gym_clip_adam", "b200gym_adaptive_lr", "b200gym_adam_prepare", "b200gym_clip_adam_dev", "b200gym_debug_mlp_trace", "b200gym_debug_chain_trace"):
        getattr(L, name).restype = C.c_int
    L.b200gym_mlp_forward.argtypes = [C.POINTER(MlpParamsPOD), vp, vp, vp, vp, vp]
    L.b200gym_tube_error.argtypes = [vp, vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, vp]
    L.b200gym_sliding_window.argtypes = [vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
    L.b200gym_tube_error.restype = L.b200gym_sliding_window.restype = C.c_int
    L.b200gym_mlp_forward.restype = C.c_int
    L.b200gym_mlp_forward_pair.argtypes = [C.POINTER(MlpParamsPOD), vp, vp, vp, vp, C.POINTER(MlpParamsPOD), vp, vp, vp, vp, vp]
    L.b200gym_mlp_forward_pair.restype = C.c_int
    L.b200gym_gemm_f16.argtypes = [C.POINTER(GemmProblemPOD), C.c_int32, vp]
    L.b200gym_gemm_f16.restype = C.c_int
    L.b200gym_rows_to_f16.argtypes = [vp, C.c_int64, C.c_int32, vp, vp, C.c_int32, C.c_int64, vp]
    L.b200gym_ppo_loss_gathered.argtypes = [C.POINTER(PpoLossParamsPOD), vp, vp, C.c_int32, vp, C.c_int32] + [vp] * 13
    L.b200gym_pack_params_f16.argtypes = [vp, C.POINTER(PackTablePOD), vp, vp]
    L.b200gym_ppo_chain.argtypes = [C.POINTER(ChainNetPOD), C.POINTER(ChainNetPOD), C.POINTER(PpoLossParamsPOD)] + [vp] * 12
    L.b200gym_ppo_chain.restype = C.c_int
    L.b200gym_ppo_optimizer_step.argtypes = [C.POINTER(OptParamsPOD)] + [vp] * 9 + [C.POINTER(PackTablePOD), vp, vp]
    L.b200gym_ppo_optimizer_step.restype = C.c_int
    L.b200gym_ppo_optimizer_step_peers.argtypes = ([C.POINTER(OptParamsPOD), C.POINTER(PeerBasesPOD), C.c_int32, C.c_int32, C.c_int64] + [vp] * 10
                                                   + [C.POINTER(PackTablePOD), vp, C.c_int32, vp])
    L.b200gym_ppo_optimizer_step_peers.restype = C.c_int
    u64 = C.c_uint64
    L.b200gym_ppo_act_store.argtypes = [i32, i32, i32, i32, vp, i32, vp, i32, vp, vp, C.c_int64, vp, C.c_int64, u64, u64, vp, u64] + [vp] * 8
    L.b200gym_ppo_store_step.argtypes = [i32] + [vp] * 8
    L.b200gym_ppo_act_store.restype = L.b200gym_ppo_store_step.restype = C.c_int
    L.b200gym_rows_to_f16.restype = L.b200gym_ppo_loss_gathered.restype = L.b200gym_pack_params_f16.restype = C.c_int
    for name, cls in (("B200MlpParams", MlpParamsPOD), ("B200PpoLossParams", PpoLossParamsPOD), ("B200LeggedParams", LeggedParamsPOD), ("B200LeggedBuffers", LeggedBuffersPOD),
                      ("B200RomParams", RomParamsPOD), ("B200RomState", RomStatePOD), ("B200PeerPtrs", PeerPtrsPOD), ("B200PeerBases", PeerBasesPOD), ("B200GemmProblem", GemmProblemPOD), ("B200PackTable", PackTablePOD), ("B200ChainNet", ChainNetPOD), ("B200OptParams", OptParamsPOD),
                      ("B200RomFamilyParams", RomFamilyParamsPOD), ("B200HopperTorqueParams", HopperTorqueParamsPOD),
                      ("B200HopperTorqueBuffers", HopperTorqueBuffersPOD), ("B200HopperObsParams", HopperObsParamsPOD),
                      ("B200HopperEnvParams", HopperEnvParamsPOD), ("B200HopperEnvBuffers", HopperEnvBuffersPOD)):
        n = L.b200gym_sizeof(name.encode())
        if n != C.sizeof(cls):
            raise RuntimeError(f"ABI mismatch: sizeof({name}) is {n} in the library, {C.sizeof(cls)} in the binding")
    _register_optional(L)
    _lib = L
    return L


_OPTIONAL = []


def _register_optional(L):
    for fn in _OPTIONAL:
        fn(L)


def check(rc, what=""):
    if rc != 0:
        msg = lib().b200gym_last_error().decode(errors="replace")
        raise RuntimeError(f"b200gym {what} failed ({rc}): {msg}")


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_bound_device = None


def require_current_device(device):
    """One process drives one GPU: the library keeps per-process state (constant-memory actuator net, kernel attributes, SM
    count) for the device it first ran on.  Raises when an object is created for a device that is not the current CUDA device, or for a
    second device in the same process, instead of computing with another device's constants."""
    global _bound_device
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"b200gym runs on CUDA devices only (got {dev})")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx != torch.cuda.current_device():
        raise RuntimeError(f"b200gym: device cuda:{idx} is not the current CUDA device (cuda:{torch.cuda.current_device()}): call "
                           "torch.cuda.set_device first — one process per GPU")
    if _bound_device is None:
        _bound_device = idx
    elif _bound_device != idx:
        raise RuntimeError(f"b200gym: this process already runs on cuda:{_bound_device}; one process per GPU (per-device kernel state is process-wide)")


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"b200gym: `{name}` lives on {t.device}; the fused path runs on CUDA only (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError(f"b200gym: `{name}` must be contiguous")


def fill_params(p, num_sum_rows, sum_row, zero_lstm_on_reset=False) -> LeggedParamsPOD:
    """LeggedParams (legged_gym_dev_b200.params) -> POD struct.  Differences `hi - lo` are taken in double
    here, exactly where the reference takes them in Python floats (helpers.py:129-130)."""
    s = LeggedParamsPOD()
    s.num_envs, s.num_obs, s.num_bodies, s.num_heights = p.num_envs, p.num_obs, p.num_bodies, p.num_height_points
    if len(p.feet_indices) != NUM_FEET or len(p.penalised_indices) != NUM_PEN or p.num_dof != NUM_DOF:
        raise ValueError("the fused step pipeline is specialised for 12 DOF / 4 feet / 8 penalised bodies (ANYmal-class)")
    if not (1 <= len(p.termination_indices) <= MAX_TERM):
        raise ValueError("1..4 termination bodies supported")
    s.feet_idx[:] = p.feet_indices
    s.pen_idx[:] = p.penalised_indices
    for i, v in enumerate(p.termination_indices):
        s.term_idx[i] = v
    s.num_term = len(p.termination_indices)
    s.control_type = p.control_type
    s.action_scale, s.sim_dt, s.dt = p.action_scale, p.sim_dt, p.dt
    s.clip_actions, s.clip_obs = p.clip_actions, p.clip_observations
    s.p_gains[:], s.d_gains[:], s.default_dof_pos[:] = p.p_gains, p.d_gains, p.default_dof_pos
    s.torque_limits[:], s.dof_vel_limits[:] = p.torque_limits, p.dof_vel_limits
    s.dof_pos_lo[:] = [a for a, _ in p.dof_pos_limits]
    s.dof_pos_hi[:] = [b for _, b in p.dof_pos_limits]
    s.obs_lin_vel, s.obs_ang_vel, s.obs_dof_pos = p.obs_lin_vel, p.obs_ang_vel, p.obs_dof_pos
    s.obs_dof_vel, s.obs_height = p.obs_dof_vel, p.obs_height
    s.add_noise = int(p.add_noise)
    s.noise_lin_vel, s.noise_ang_vel, s.noise_gravity = p.noise_lin_vel, p.noise_ang_vel, p.noise_gravity
    s.noise_dof_pos, s.noise_dof_vel, s.noise_height = p.noise_dof_pos, p.noise_dof_vel, p.noise_height
    s.heading_command, s.resample_steps = int(p.heading_command), p.resample_steps
    for i, r in enumerate((p.cmd_lin_vel_x, p.cmd_lin_vel_y, p.cmd_ang_vel_yaw, p.cmd_heading)):
        s.cmd_lo[i], s.cmd_span[i] = r[0], r[1] - r[0]
        s.cmd_lo_reset[i], s.cmd_span_reset[i] = r[0], r[1] - r[0]
    s.max_command_x = p.cmd_lin_vel_x[1]
    s.push_robots, s.push_time = int(p.push_robots), int(p.push_time)
    s.push_lo, s.push_span = -p.max_push_vel, p.max_push_vel - (-p.max_push_vel)
    s.max_episode_length, s.max_episode_length_s = p.max_episode_length, p.max_episode_length_s
    s.reward_scale[:] = p.reward_scales
    s.sum_row[:] = sum_row
    s.num_sum_rows, s.only_positive = num_sum_rows, int(p.only_positive_rewards)
    s.tracking_sigma, s.soft_dof_vel_limit, s.soft_torque_limit = p.tracking_sigma, p.soft_dof_vel_limit, p.soft_torque_limit
    s.base_height_target, s.max_contact_force = p.base_height_target, p.max_contact_force
    s.mesh_plane = int(p.mesh_type == "plane")
    s.terrain_curriculum = int(p.terrain_curriculum)
    s.terrain_rows, s.terrain_cols = p.terrain_rows, p.terrain_cols
    s.max_terrain_level, s.terrain_num_cols = p.max_terrain_level, p.terrain_num_cols
    s.border_size, s.horizontal_scale, s.vertical_scale = p.border_size, p.horizontal_scale, p.vertical_scale
    s.half_env_length = p.terrain_length / 2
    if p.measure_heights:
        if len(p.measured_points_x) > MAX_POINTS or len(p.measured_points_y) > MAX_POINTS:
            raise ValueError("at most 32 x 32 height points")
        s.n_px, s.n_py = len(p.measured_points_x), len(p.measured_points_y)
        for i, v in enumerate(p.measured_points_x):
            s.points_x[i] = v
        for i, v in enumerate(p.measured_points_y):
            s.points_y[i] = v
    s.custom_origins, s.zero_lstm_on_reset = int(p.custom_origins), int(zero_lstm_on_reset)
    s.base_init_state[:] = p.base_init_state
    s.seed_lo, s.seed_hi = p.seed & 0xFFFFFFFF, (p.seed >> 32) & 0xFFFFFFFF
    s.traj_mode, s.traj_n, s.traj_horizon = int(p.traj_mode), p.traj_n, p.traj_horizon
    s.traj_scale[:], s.traj_weight[:] = p.traj_scale, p.traj_weight
    s.diff_neg_slope, s.diff_pos_slope = p.diff_neg_slope, p.diff_pos_slope
    s.push_t_lo, s.push_t_span = p.time_between_pushes[0], p.time_between_pushes[1] - p.time_between_pushes[0]
    return s
