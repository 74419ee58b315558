/* b200gym — C ABI of the B200-native per-step tensor pipeline for legged_gym_dev.
 *
 * Drop-in boundary.  The reference has no native code: every entry point below
 * replaces a group of Python/torch methods, cited as file:line of /root/reference.  All functions
 *   - take raw device pointers + sizes + a cudaStream_t (as void*), never torch types,
 *   - only ENQUEUE work on the stream (no host sync, CUDA-graph capturable),
 *   - return 0 on success, <0 on error with a message in b200gym_last_error() (thread-local),
 *   - never allocate or free tensor memory (PyTorch / PhysX own it).
 * There is no CPU fallback.
 */
#ifndef B200GYM_H
#define B200GYM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GYM_VERSION 100
#define B200GYM_NUM_DOF 12
#define B200GYM_NUM_FEET 4
#define B200GYM_NUM_PEN 8
#define B200GYM_MAX_TERM 4
#define B200GYM_NUM_REWARD_TERMS 21 /* alphabetical _reward_* names of LeggedRobot + LeggedRobotTrajectory, `termination` last
                                       (legged_robot.py:605-629, legged_robot_trajectory.py:663-687) */
#define B200GYM_TRAJ_WIDTH 20 /* trajectory observation block: N = 10 knots of a 2-state rom (legged_robot_trajectory.py:277-283) */
#define B200GYM_MAX_POINTS 32

int b200gym_version(void);
const char* b200gym_last_error(void);
/* sizeof() of the POD structs below, so a binding can verify its mirror of the layout */
int b200gym_sizeof(const char* struct_name);

/* ------------------------------------------------------------------------------------------------
 * Group R — LeggedRobot / Anymal step pipeline
 * ---------------------------------------------------------------------------------------------- */
typedef struct B200LeggedParams {
    int32_t num_envs, num_obs, num_bodies, num_heights; /* num_heights = 0 unless cfg.terrain.measure_heights */
    int32_t feet_idx[B200GYM_NUM_FEET], pen_idx[B200GYM_NUM_PEN], term_idx[B200GYM_MAX_TERM], num_term;
    int32_t control_type; /* 0 P, 1 V, 2 T (legged_robot.py:402-412) */
    float action_scale, sim_dt, dt, clip_actions, clip_obs;
    float p_gains[B200GYM_NUM_DOF], d_gains[B200GYM_NUM_DOF], default_dof_pos[B200GYM_NUM_DOF];
    float torque_limits[B200GYM_NUM_DOF], dof_pos_lo[B200GYM_NUM_DOF], dof_pos_hi[B200GYM_NUM_DOF];
    float dof_vel_limits[B200GYM_NUM_DOF];
    float obs_lin_vel, obs_ang_vel, obs_dof_pos, obs_dof_vel, obs_height;
    int32_t add_noise;
    float noise_lin_vel, noise_ang_vel, noise_gravity, noise_dof_pos, noise_dof_vel, noise_height;
    int32_t heading_command, resample_steps;
    float cmd_lo[4], cmd_span[4]; /* lin_vel_x, lin_vel_y, ang_vel_yaw, heading: lower and (upper - lower) */
    /* the ranges the RESET resample of this launch draws from: equal to cmd_lo / cmd_span except on the step on which the
     * command curriculum advances (legged_robot.py:360-363 runs between the periodic resample and reset_idx); max_command_x =
     * extras["episode"]["max_command_x"] (:183-184), written to extras_out[num_sum_rows + 2] when an env reset */
    float cmd_lo_reset[4], cmd_span_reset[4], max_command_x;
    int32_t push_robots, push_time;
    float push_lo, push_span;
    float max_episode_length, max_episode_length_s;
    float reward_scale[B200GYM_NUM_REWARD_TERMS]; /* cfg scale * dt; 0 = term inactive */
    int32_t sum_row[B200GYM_NUM_REWARD_TERMS];    /* row of episode_sums[K,N] or -1 */
    int32_t num_sum_rows, only_positive;
    float tracking_sigma, soft_dof_vel_limit, soft_torque_limit, base_height_target, max_contact_force;
    int32_t mesh_plane, terrain_curriculum, terrain_rows, terrain_cols, max_terrain_level, terrain_num_cols;
    float border_size, horizontal_scale, vertical_scale, half_env_length;
    int32_t n_px, n_py;
    float points_x[B200GYM_MAX_POINTS], points_y[B200GYM_MAX_POINTS];
    int32_t custom_origins, zero_lstm_on_reset;
    float base_init_state[13];
    uint32_t seed_lo, seed_hi;
    /* LeggedRobotTrajectory (legged_gym/envs/base/legged_robot_trajectory.py): traj_mode != 0 selects its step semantics —
       no command resampling (:405-417), per-env push timers (:169-178), rewards tracking_rom (:1060-1069) and
       differential_error (:1100-1110), prev_error on reset (:233), trajectory observation block (:274-287). */
    int32_t traj_mode, traj_n, traj_horizon;
    float traj_scale[4];  /* normalization.obs_scales.trajectory, per rom state */
    float traj_weight[4]; /* rom.get_weighting_vector(cfg.rewards.reward_weighting) (rom_dynamics.py:209-211) */
    float diff_neg_slope, diff_pos_slope; /* cfg.rewards.differential_error */
    float push_t_lo, push_t_span;         /* cfg.domain_rand.time_between_pushes: lower, upper - lower */
} B200LeggedParams;

typedef struct B200LeggedBuffers {
    /* PhysX-owned, AoS as given (legged_robot.py:545-551) */
    float* root_states;          /* [N,13]  pos3 quat4(xyzw) lin3 ang3; written on push / reset */
    float* dof_state;            /* [N,12,2] (pos,vel); written on reset */
    const float* contact_forces; /* [N,B,3] */
    const float* actions;        /* [N,12] clipped */
    const float* torques;        /* [N,12] */
    /* env-owned state (legged_robot.py:561-581, base_task.py:70-79) */
    float* last_actions;         /* [N,12] */
    float* last_dof_vel;         /* [N,12] */
    float* last_root_vel;        /* [N,6] */
    float* commands;             /* [N,4] */
    float* feet_air_time;        /* [N,4] */
    uint8_t* last_contacts;      /* [N,4] bool */
    int64_t* episode_length_buf; /* [N] */
    uint8_t* reset_buf;          /* [N] bool (out) */
    uint8_t* time_out_buf;       /* [N] bool (out) */
    float* rew_buf;              /* [N] (out) */
    float* episode_sums;         /* [K,N] rows = active reward terms, alphabetical */
    float* obs_buf;              /* [N,O] (out, already clipped to +-clip_obs) */
    float* base_lin_vel;         /* [N,3] (out) */
    float* base_ang_vel;         /* [N,3] (out) */
    float* projected_gravity;    /* [N,3] (out) */
    float* measured_heights;     /* [N,H] (out) or NULL */
    const int16_t* height_samples; /* [rows,cols] or NULL */
    float* env_origins;          /* [N,3] */
    int64_t* terrain_levels;     /* [N] or NULL */
    const int64_t* terrain_types;  /* [N] or NULL */
    const float* terrain_origins;  /* [max_level, num_cols, 3] or NULL */
    float* lstm_h;               /* [2,N*12,8] or NULL: zeroed for reset envs (anymal.py:59-60) */
    float* lstm_c;
    /* extras["episode"] (legged_robot.py:175-182): [K] means / max_episode_length_s, [K] terrain_level mean,
       [K+1] number of envs reset this step.  Entries are left untouched on steps without a reset (:156-157). */
    float* extras_out;           /* [K+2] */
    double* ws_sums;             /* [K+2] workspace, zero-initialised by the caller once */
    uint32_t* ws_counter;        /* [1]   workspace, zero-initialised by the caller once */
    /* optional device-resident step counter: when non-NULL the kernels read the step (RNG event, push schedule) from it
       instead of the by-value argument and advance it after every step, which makes the launch sequence of whole env
       steps CUDA-graph replayable.  It must hold common_step_counter of the NEXT post-physics pass. */
    uint64_t* step_counter;
    /* traj_mode only */
    const float* trajectory;     /* [N, horizon, n] self.trajectory = traj_gen.get_trajectory() of this step (:410) */
    float* prev_error;           /* [N, n] */
    float* time_until_next_push; /* [N] */
    /* optional [num_sum_rows + 2] doubles: the RAW statistics behind extras_out of this step — per-term sums over the envs that reset,
     * sum of terrain levels over all envs, number of resets (0 when none) — so that env shards can all-reduce (sum, count) pairs and
     * log the same means a single process over all envs would (legged_robot.py:175-182) */
    double* extras_raw;
} B200LeggedBuffers;

/* LeggedRobot._compute_torques (legged_robot.py:389-413), optionally fused with the action clip of
 * LeggedRobot.step (:86-87): if actions_clipped != NULL the clipped actions are also written there.
 * last_dof_vel is only read for control_type V. */
int b200gym_pd_torques(const B200LeggedParams* p, const float* actions, float* actions_clipped, const float* dof_state,
                       const float* last_dof_vel, float* torques, void* stream);

/* Anymal._compute_torques, actuator-network branch (anymal.py:71-78) incl. the TorchScript LSTMsea forward.
 * h, c: [2, N*12, 8], updated in place. */
int b200gym_set_actuator_net(const float* w_ih0, const float* w_hh0, const float* b_ih0, const float* b_hh0,
                             const float* w_ih1, const float* w_hh1, const float* b_ih1, const float* b_hh1,
                             const float* w_lin, const float* b_lin, float in_scale0, float in_scale1, float out_scale);
int b200gym_lstm_torques(const B200LeggedParams* p, const float* actions, float* actions_clipped, const float* dof_state,
                         float* h, float* c, float* torques, void* stream);
/* A/B aid: 0 = scalar-fma kernel (default), 3 = the same kernel with packed FFMA2 gate mat-vecs (bit-identical, 4-8 % slower),
 * 1 = the gate mat-vecs on the tcgen05 tensor cores with the 3xTF32 operand split
 * ([A_hi | A_lo | A_hi] x [W_hi | W_hi | W_lo], fp32 accumulation in TMEM; meets the 1e-5 torque contract), -1 = re-read the
 * environment variable B200GYM_LSTM_VARIANT. */
int b200gym_debug_set_lstm_variant(int variant);

/* LeggedRobot.post_physics_step (legged_robot.py:106-134) + the obs clip of step() (:100-101), fused:
 * counters, body-frame velocities, command resampling / heading, height scan, pushes, termination,
 * all reward terms, in-place reset, observations + noise, last_* copies, extras["episode"] statistics.
 * `step` is common_step_counter AFTER its increment (:115); env_id_offset = global id of local env 0. */
int b200gym_post_physics(const B200LeggedParams* p, const B200LeggedBuffers* b, uint64_t step, int64_t env_id_offset,
                         void* stream);

/* LeggedRobot.reset_idx(env_ids) as an EXTERNAL call (legged_robot.py:147-187, anymal.py:56-60; BaseTask.reset, base_task.py:111-119)
 * for the envs flagged in reset_mask (uint8 [N]): terrain curriculum, dof / root state redraw, command resample, last_actions /
 * last_dof_vel / feet_air_time / episode_length_buf cleared, reset_buf set, episode_sums folded into extras_out (means over the
 * reset envs / max_episode_length_s, terrain-level mean over all envs) and cleared, actuator LSTM state zeroed.  time_out_buf,
 * obs_buf and rew_buf are left alone, as in the reference.  `event` keys the Philox draws (sites as in b200gym_post_physics); pass
 * a value no env step uses (the Python mirror: (number of external resets << 40) | common_step_counter).  traj_mode
 * (LeggedRobotTrajectory.reset_idx, legged_robot_trajectory.py:204-246): no command resample; prev_error is rewritten from the
 * trajectory buffer and the new root (:233); the caller then resets the generators (b200gym_rom_reset_from_root with the same mask). */
int b200gym_legged_reset_idx(const B200LeggedParams* p, const B200LeggedBuffers* b, const uint8_t* reset_mask, uint64_t event,
                             int64_t env_id_offset, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Group M — reduced-order-model rollout (trajopt/rom_dynamics.py, deep_tube_learning/custom_sim.py,
 * deep_tube_learning/controllers.py, deep_tube_learning/data_collection_trajectory.py)
 * ---------------------------------------------------------------------------------------------- */
#define B200GYM_ROM_SINGLE_INT_2D 0 /* rom_dynamics.py:182-211  n=2 m=2 */
#define B200GYM_ROM_DOUBLE_INT_2D 1 /* rom_dynamics.py:214-260  n=4 m=2 */
#define B200GYM_ROM_MAX_WINDOW 32
/* trajectory generator classes (rom_dynamics.py): the random TrajectoryGenerator :441-615 and its deterministic subclasses
 * ZeroTrajectoryGenerator :618-624, SquareTrajectoryGenerator :627-675, CircleTrajectoryGenerator :678-698 (SingleInt2D rom) */
#define B200GYM_GEN_RANDOM 0
#define B200GYM_GEN_ZERO 1
#define B200GYM_GEN_SQUARE 2
#define B200GYM_GEN_CIRCLE 3

typedef struct B200RomParams {
    int32_t num_envs, model_type, rom_type, window, dN, horizon; /* window = N*dN, horizon = N (rom_dynamics.py:485-486) */
    float model_dt, rom_dt, dt_loop;
    float model_z_min[4], model_z_max[4], model_v_min[2], model_v_max[2];
    float rom_z_min[4], rom_z_max[4], rom_v_min[2], rom_v_max[2];
    float t_low, t_span, freq_low, freq_high, prob_stationary; /* UniformSampleHoldDT (utils.py:27-43), :544,:520 */
    int32_t weight_zero_col; /* -1 UniformWeightSampler, 1 ...NoRamp, 2 ...NoExtreme (utils.py:46-79) */
    int32_t randomize_rom_distance;
    float max_rom_distance[4], zero_rom_dist_llh, noise_lower[4], noise_upper[4]; /* custom_sim.py:32-35,80-91 */
    float Kp, Kd; /* DoubleSingleTracking (controllers.py:80-92) */
    uint32_t seed_lo, seed_hi;
    int32_t gen_kind;  /* B200GYM_GEN_* */
    float gen_c[4];    /* Square: leg boundaries c1..c4 (rom_dynamics.py:633-636), computed by the host in fp32 like the reference */
    float gen_v[4];    /* Square: v_max[1]/2, v_max[0], v_min[1]/2, v_min[1] (:637-640, incl. its v_min[1] quirk); Circle: [0] = speed (:692) */
} B200RomParams;

/* Generator + sim state; every tensor row-major [N, ...] fp32 with the reference's shapes (rom_dynamics.py:487-508,
 * custom_sim.py:29-31) so the Python attributes traj_gen.{k,t,v,trajectory,v_trajectory,...} alias them directly. */
typedef struct B200RomState {
    float* root_states;  /* [N, model_n]                 CustomSim.root_states (may be NULL for generator-only use) */
    float* trajectory;   /* [N, window+1, rom_n]         traj_gen.trajectory */
    float* v_trajectory; /* [N, window, 2]               traj_gen.v_trajectory */
    float* v;            /* [N, 2] */
    float *t, *k, *t_final; /* [N] fp32 (rom_dynamics.py:488-490) */
    float* weights;      /* [N, 4] */
    float *sample_hold_input, *extreme_input, *ramp_v_start, *ramp_v_end; /* [N, 2] */
    float* ramp_t_start; /* [N] */
    float *sin_mag, *sin_freq, *sin_off, *sin_mean; /* [N, 2] */
    uint8_t* stationary_inds; /* [N] bool */
    int32_t* rng_ctr;    /* [N] per-env draw-event counter of the counter-based RNG */
    float* env_trajectory; /* [N, horizon, rom_n]        CustomSim.trajectory (interpolated, custom_sim.py:74) or NULL */
    float* obs;          /* [N, model_n + rom_n + 2]     CustomSim.get_observations (custom_sim.py:95-100) or NULL */
    float* center;       /* [N, 2]                       CircleTrajectoryGenerator.center or NULL */
} B200RomState;

/* TrajectoryGenerator.__init__ draw of ramp_v_end (rom_dynamics.py:495); all other state must be zero-filled. */
int b200gym_rom_init(const B200RomParams* p, const B200RomState* s, int64_t env_id_offset, void* stream);
/* CustomSim.step (custom_sim.py:71-75) = model.f + TrajectoryGenerator.step (rom_dynamics.py:568-590) + get_trajectory
 * (:607-612) + get_observations.  action == NULL: generator only (TrajectoryGenerator.step / step_idx).
 * step_mask (uint8 [N]) == NULL: all envs (step_idx(arange)). */
int b200gym_rom_step(const B200RomParams* p, const B200RomState* s, const float* action, const uint8_t* step_mask,
                     int64_t env_id_offset, void* stream);
/* CustomSim.reset_idx (custom_sim.py:80-93) incl. reset_traj and TrajectoryGenerator.reset_idx (rom_dynamics.py:595-605)
 * and the trailing zero-action step of ALL envs.  reset_mask == NULL: CustomSim.reset(). */
int b200gym_rom_reset(const B200RomParams* p, const B200RomState* s, const uint8_t* reset_mask, int64_t env_id_offset,
                      void* stream);
/* LeggedRobotTrajectory.reset_traj (legged_robot_trajectory.py:248-253) + TrajectoryGenerator.reset_idx (rom_dynamics.py:595-605)
 * for the envs flagged in reset_mask: p_zx = proj_z(root) read from root [N, root_stride] (SingleInt2D: first two columns),
 * optional random offset, generator reset and N*dN warm-up steps.  any_reset (device float, e.g. extras_out[K+1] of
 * b200gym_post_physics, or NULL): when it reads 0 the call does nothing, as reset_idx returns early on an empty id list (:217-218).
 * The interpolated env_trajectory is NOT refreshed (the reference keeps the pre-reset clone until the next step, :410). */
int b200gym_rom_reset_from_root(const B200RomParams* p, const B200RomState* s, const uint8_t* reset_mask, const float* root,
                                int32_t root_stride, const float* any_reset, int64_t env_id_offset, void* stream);
/* DoubleSingleTracking.__call__ (controllers.py:87-92) with DoubleInt2D.clip_v_z (rom_dynamics.py:234-250). */
int b200gym_rom_tracking_policy(const B200RomParams* p, const float* obs, float* action, void* stream);
/* RaibertHeuristic.raibert_policy (deep_tube_learning/controllers.py:38-73, the hopper's tracking controller — SURVEY 8f row 3):
 * obs [n, obs_stride >= 10] = (pos err x, y, vel err x, y, desired vel x, y, quat xyzw) -> action [n, 4] = desired orientation
 * quaternion (w, x, y, z) = omega_to_quat(clamped pitch, clamped roll, current yaw). */
int b200gym_raibert_policy(const float* obs, int32_t obs_stride, int64_t n, float Kp, float Kv, float Kff, float clip_pos, float clip_vel,
                           float clip_ang, float* action, void* stream);
/* One epoch of data_collection_trajectory.py:104-149 as ONE persistent launch: reset all envs, then T ROM steps of
 * {policy -> CustomSim.step} with generator state in registers; logs x [N,T+1,model_n] (may be NULL), z, pz_x
 * [N,T+1,rom_n], v [N,T,2], done [N,T] (bool).  obs_io [N,8]: in = observation the first action is computed from
 * (the reference reuses the previous epoch's, :94,:111), out = last observation. */
int b200gym_rom_rollout(const B200RomParams* p, const B200RomState* s, float* obs_io, int32_t T, float* x, float* z, float* pz_x,
                        float* v, uint8_t* done, int64_t env_id_offset, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Group M, SURVEY 8f row 4 — the whole RomDynamics family (trajopt/rom_dynamics.py:182-438) behind one generic generator.
 * The kernels above are the register-resident fast path of the shipped data-generation configs (SingleInt2D / DoubleInt2D,
 * 2 inputs); the entry points below cover every rom class — state dimension up to 6, up to 3 inputs — with the same
 * state tensors (B200RomState, [N, m] / [N, window+1, n] with the class's own n and m), the same draw events and, for
 * rom types 0 / 1, bit-identical results to the kernels above.
 * ---------------------------------------------------------------------------------------------- */
#define B200GYM_ROM_UNICYCLE 2                  /* rom_dynamics.py:263-305  n=3 m=2  [x, y, theta] / [v, omega] */
#define B200GYM_ROM_LATERAL_UNICYCLE 3          /* rom_dynamics.py:307-333  n=3 m=3  [x, y, theta] / [v, v_perp, omega] */
#define B200GYM_ROM_EXTENDED_UNICYCLE 4         /* rom_dynamics.py:336-394  n=5 m=2  [x, y, theta, v, omega] / [a, alpha] */
#define B200GYM_ROM_EXTENDED_LATERAL_UNICYCLE 5 /* rom_dynamics.py:397-438  n=6 m=3  [x, y, theta, v, v_perp, omega] / [a, a_perp, alpha] */
#define B200GYM_ROM_NUM_TYPES 6

typedef struct B200RomFamilyParams {
    int32_t num_envs, rom_type, window, dN; /* window = N*dN (rom_dynamics.py:485-486) */
    float rom_dt, dt_loop;
    float z_min[8], z_max[8], v_min[4], v_max[4]; /* RomDynamics.__init__ bounds (:16-33), first n / m entries used */
    float t_low, t_span, freq_low, freq_high, prob_stationary;
    int32_t weight_zero_col; /* as in B200RomParams */
    uint32_t seed_lo, seed_hi;
} B200RomFamilyParams;

/* RomDynamics.f (:192,:224,:273-278,:311-316,:345-352,:406-414): z [n_rows, n], v [n_rows, m] -> z_next [n_rows, n]. */
int b200gym_romfam_f(int32_t rom_type, float dt, const float* z, const float* v, float* z_next, int64_t n_rows, void* stream);
/* RomDynamics.des_pose_vel (:198,:230,:286-290,:318-322,:354-358,:416-420; incl. LateralUnicycle's om = v[:, 1]):
 * pose [n_rows, 3] = (x, y, yaw), vel [n_rows, 3] = (xdot, ydot, yawdot). */
int b200gym_romfam_des_pose_vel(int32_t rom_type, const float* z, const float* v, float* pose, float* vel, int64_t n_rows, void* stream);
/* RomDynamics.compute_state_dependent_input_bounds (:106-107,:234-246,:367-379) and, when v != NULL, clip_v_z (:201,:248-250,
 * :292,:381-383): v_lo / v_hi [n_rows, m] (either may be NULL), v_clipped [n_rows, m].  Uses p->rom_type, rom_dt, z_*, v_*. */
int b200gym_romfam_input_bounds(const B200RomFamilyParams* p, const float* z, const float* v, float* v_lo, float* v_hi, float* v_clipped,
                                int64_t n_rows, void* stream);
/* RomDynamics.proj_z (:195,:227,:280-284,:360-365,:422-427): x [n_rows, 13] = (pos 3, quat xyzw 4, lin vel 3, ang vel 3) ->
 * z [n_rows, n]; yaw = last angle of scipy's Rotation.as_euler('xyz'), evaluated on the device in fp32. */
int b200gym_romfam_proj_z(int32_t rom_type, const float* x, float* z, int64_t n_rows, void* stream);
/* TrajectoryGenerator.__init__ draw of ramp_v_end (:495) for m inputs. */
int b200gym_romfam_gen_init(const B200RomFamilyParams* p, const B200RomState* s, int64_t env_id_offset, void* stream);
/* TrajectoryGenerator.reset_idx(idx, z) (:595-605): z [N, n]; reset_mask uint8 [N] or NULL = reset(z).  Envs outside the mask
 * take part in the warm-up's input evaluations (resamples that are due, self.v), as in step_rom_idx (:577-580). */
int b200gym_romfam_gen_reset(const B200RomFamilyParams* p, const B200RomState* s, const float* z, const uint8_t* reset_mask,
                             int64_t env_id_offset, void* stream);
/* TrajectoryGenerator.step_idx(idx) (:571-590); step_mask uint8 [N] or NULL = step(). */
int b200gym_romfam_gen_step(const B200RomFamilyParams* p, const B200RomState* s, const uint8_t* step_mask, int64_t env_id_offset,
                            void* stream);
/* TrajectoryGenerator.get_input_t(t, z) (:560-566) with caller-supplied clocks t [N] and states z [N, n] (the open-loop use of
 * trajopt/trajectory_gen.py:35-41): resamples the envs with t > t_final, returns the weighted clipped input in v_out [N, m]. */
int b200gym_romfam_gen_input(const B200RomFamilyParams* p, const B200RomState* s, const float* t, const float* z, float* v_out,
                             int64_t env_id_offset, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SURVEY 8f row 3 (part) — the Hopper's torque law, Hopper._compute_torques (legged_gym/envs/hopper/hopper.py:168-237), for the control
 * types the reference method can run: "orientation" and "orientation_spindown" (hopper_config.py:62-63).  4 DOF: foot slide (dof 0) +
 * three reaction wheels (dofs 1-3); actions = desired base orientation, a raw (w, x, y, z) quaternion.
 * ---------------------------------------------------------------------------------------------- */
typedef struct B200HopperTorqueParams {
    int32_t num_envs, num_bodies, foot_body, spindown; /* contact = contact_forces[:, foot_body, 2] > 0.1 (:184); spindown: :204-206 */
    float action_scale, torque_speed_bound_ratio;      /* cfg.control.action_scale, cfg.asset.torque_speed_bound_ratio (:68) */
    float p_gains[4], d_gains[4];                      /* LeggedRobot._init_buffers' per-dof gains (hopper_config.py:35-48) */
    float kd_spindown[3], wheel_speed_limits[3];       /* hopper.py:390-403 */
    float torque_limits[4];                            /* asset effort limits */
    float rot_actuator[9];                             /* cfg.asset.rot_actuator, row-major; tau = local_tau @ R (pytorch3d Rotate, :67,:221) */
    /* 1: base_ang_vel = quat_rotate_inverse(root quaternion, root_states[:, 10:13]) is formed in the kernel from the root state the
     * previous sub-step left (the refresh of hopper_trajectory.py:124-126) instead of being read from the buffer; the env's first
     * sub-step of a step reads the buffer (it holds the pre-reset value, as in the reference) */
    int32_t ang_vel_from_root, pad;
} B200HopperTorqueParams;

typedef struct B200HopperTorqueBuffers {
    const float* actions;        /* [N, 4] */
    const float* dof_state;      /* [N, 4, 2] (pos, vel) */
    const float* contact_forces; /* [N, num_bodies, 3] */
    const float* root_states;    /* [N, 13]; quaternion xyzw in columns 3-6 */
    const float* base_ang_vel;   /* [N, 3] */
    const float *p_gain_random, *d_gain_random, *torque_limit_random; /* [N, 4] multipliers (:362-378) */
    const float* wheel_limit_random;                                  /* [N, 3] (:380-382) */
    const float *spring_stiffness, *spring_damping, *foot_pos_des, *torque_speed_bound_ratio_random; /* [N] ([N, 1] in the reference) */
    float* torques;              /* [N, 4] self.torques: after the torque-speed clip, before the torque-limit clip (:236) */
    float* torques_clipped;      /* [N, 4] the return value (:237) */
} B200HopperTorqueBuffers;

int b200gym_hopper_torques(const B200HopperTorqueParams* p, const B200HopperTorqueBuffers* b, void* stream);

/* Hopper.compute_observations (hopper.py:239-258, cfg.terrain.measure_heights = False as shipped, hopper_config.py:13) + the observation
 * clip of Hopper.step (:116-117): obs [N, 21] = (z * z_pos, base_quat xyzw, base_lin_vel * lin_vel, base_ang_vel * ang_vel,
 * wheel dof_vel * dof_vel, commands[:3] * commands_scale, actions / |actions| with qw >= 0) + (2u - 1) * noise_scale_vec, u = the
 * Philox uniforms of (seed, global env id, event, site OBS_NOISE, column) — event = common_step_counter, as in b200gym_post_physics. */
#define B200GYM_HOPPER_NUM_OBS 21
typedef struct B200HopperObsParams {
    int32_t num_envs, add_noise;
    float z_pos_scale, lin_vel_scale, ang_vel_scale, dof_vel_scale, clip_observations; /* cfg.normalization (hopper_config.py:92-100) */
    float commands_scale[3];                     /* LeggedRobot._init_buffers: (lin_vel, lin_vel, ang_vel) */
    float noise_scale_vec[B200GYM_HOPPER_NUM_OBS]; /* Hopper._get_noise_scale_vec (:407-430) */
    uint32_t seed_lo, seed_hi;
} B200HopperObsParams;
int b200gym_hopper_observations(const B200HopperObsParams* p, const float* root_states, const float* base_lin_vel, const float* base_ang_vel,
                                const float* dof_state, const float* commands, const float* actions, float* obs, uint64_t event,
                                int64_t env_id_offset, void* stream);
/* The Hopper's own reward terms (hopper.py:448-458), raw (before scale * dt): out [N, 3] = (_reward_torque_limits = sum |wheel torques|,
 * _reward_dof_acc = sum ((last_dof_vel - dof_vel) / dt)^2 over the wheels, _reward_unit_quat = (1 - |actions|)^2). */
int b200gym_hopper_reward_terms(int32_t num_envs, float dt, const float* torques, const float* dof_state, const float* last_dof_vel,
                                const float* actions, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SURVEY 8f row 3 — HopperTrajectory.post_physics_step (legged_gym/envs/hopper/hopper_trajectory.py:135-182) with the inherited
 * check_termination / compute_reward / reset_idx of legged_gym/envs/base/legged_robot_trajectory.py (:194-272), the Hopper's
 * _reset_dofs / _reset_root_states / _push_robots (:298-370), compute_observations (:255-282) + the observation clip of step
 * (:128-129).  The caller runs the trajectory generator around it as for the ANYmal trajectory env: b200gym_rom_step before (the
 * callback, legged_robot_trajectory.py:409-410), b200gym_rom_reset_from_root with reset_buf after (reset_traj, :248-253).
 * Specification pinned to the unmodified class: oracle/port_hopper_env.py.
 * ---------------------------------------------------------------------------------------------- */
#define B200GYM_HOPPER_NUM_TERMS 20     /* alphabetical: action_rate, ang_vel_xy, base_height, collision, differential_error, dof_acc,
                                           dof_pos_limits, dof_vel, dof_vel_limits, feet_air_time, feet_contact_forces, lin_vel_z, orientation,
                                           raibert, stumble, torque_limits, torques, tracking_rom, unit_quat; termination last */
#define B200GYM_HOPPER_TRAJ_NUM_OBS (14 + B200GYM_TRAJ_WIDTH + 4)
typedef struct B200HopperEnvParams {
    int32_t num_envs, num_bodies, foot_body, num_term, num_pen, num_sum_rows;
    int32_t term_idx[8], pen_idx[8];
    int32_t push_robots, only_positive, add_noise, randomize_yaw;
    float dt, push_dt, max_episode_length, max_episode_length_s; /* push_dt = decimation * sim_dt (:148) */
    float push_t_lo, push_t_span, max_push_vel[6];               /* time_between_pushes (:156-159), cfg.domain_rand.max_push_vel (:99) */
    float reward_scale[B200GYM_HOPPER_NUM_TERMS];                 /* cfg scale * dt; 0 = inactive */
    int32_t sum_row[B200GYM_HOPPER_NUM_TERMS];                    /* row of episode_sums [K, N] (alphabetical over the active names) or -1 */
    float tracking_sigma, soft_dof_vel_limit, base_height_target, max_contact_force, traj_weight[2], diff_neg_slope, diff_pos_slope;
    float raibert[6];                                             /* Kp, Kv, Kff, clip_pos, clip_vel, clip_ang (cfg.rewards.raibert) */
    float dof_pos_lo[4], dof_pos_hi[4], dof_vel_limits[4];
    float default_dof_pos[4], dof_pos_noise_lo[4], dof_pos_noise_span[4], dof_vel_noise_lo[4], dof_vel_noise_span[4]; /* span = fp32(upper) - fp32(lower) */
    float base_init_state[13], root_pos_noise_lo[5], root_pos_noise_span[5], root_vel_noise_lo[6], root_vel_noise_span[6]; /* columns 2..6 / 7..12 */
    float zero_action[4];
    float z_pos_scale, lin_vel_scale, ang_vel_scale, dof_vel_scale, clip_obs, traj_scale[2], noise_scale_vec[14];
    uint32_t seed_lo, seed_hi;
} B200HopperEnvParams;
typedef struct B200HopperEnvBuffers {
    float* root_states;           /* [N, 13] in/out (pushes, resets) */
    float* dof_state;             /* [N, 4, 2] in/out (resets) */
    const float* contact_forces;  /* [N, num_bodies, 3] */
    float* actions;               /* [N, 4] the clipped actions of this step; zero_action written for envs that reset (:314) */
    const float* torques;         /* [N, 4] the limit-clipped torques of the last sub-step (self.torques after :112) */
    float *last_actions, *last_dof_vel, *last_root_vel; /* [N, 4], [N, 4], [N, 6] */
    float *base_lin_vel, *base_ang_vel, *projected_gravity; /* [N, 3] out */
    float* feet_air_time;         /* [N] */
    uint8_t* last_contacts;       /* [N] */
    int64_t* episode_length_buf;  /* [N] */
    uint8_t *reset_buf, *time_out_buf; /* [N] out */
    float* rew_buf;               /* [N] out */
    float* episode_sums;          /* [K, N] */
    float* obs_buf;               /* [N, B200GYM_HOPPER_TRAJ_NUM_OBS] out */
    const float* trajectory;      /* [N, 10, 2] traj_gen.get_trajectory() after this step's generator step */
    const float* gen_v;           /* [N, 2] traj_gen.v */
    float* prev_error;            /* [N, 2] */
    float* time_until_next_push;  /* [N] */
    const float* env_origins;     /* [N, 3] */
    float* extras_out;            /* [K + 2]: means of episode_sums over the envs that reset / max_episode_length_s; [K + 1] = reset count */
    double* ws_sums;              /* [K + 2] zero-initialised workspace */
    uint32_t* push_flag;          /* 1 word, zero-initialised */
} B200HopperEnvBuffers;
int b200gym_hopper_post_physics(const B200HopperEnvParams* p, const B200HopperEnvBuffers* b, uint64_t step, int64_t env_id_offset, void* stream);
/* LeggedRobotTrajectory.reset_idx(env_ids) as an external call for the Hopper (legged_robot_trajectory.py:204-246 with the Hopper's _reset_* methods;
 * HopperTrajectory.reset, hopper_trajectory.py:286-296) for the envs flagged in reset_mask (uint8 [N]); draws keyed by `event` (the reference's own
 * call sequence keys them with common_step_counter).  The caller then resets the generators: b200gym_rom_reset_from_root with the same mask. */
int b200gym_hopper_reset_idx(const B200HopperEnvParams* p, const B200HopperEnvBuffers* b, const uint8_t* reset_mask, uint64_t event,
                             int64_t env_id_offset, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Group G — rsl_rl rollout storage + PPO update (rsl_rl v1.0.2, a fork of which the reference imports at
 * legged_gym/utils/task_registry.py:37-38; source NOT in /root/reference: arithmetic restated)
 * ---------------------------------------------------------------------------------------------- */
/* RolloutStorage.compute_returns (rsl_rl/storage/rollout_storage.py) as a per-env reverse scan over [T,N] tensors:
 *   nt = 1 - done_t; delta = r_t + nt*gamma*V_{t+1} - V_t; A = delta + nt*gamma*lam*A; returns_t = A + V_t;
 *   advantages_t = returns_t - V_t  (un-normalised).  If time_outs != NULL the time-out bootstrap of
 *   PPO.process_env_step (rewards += gamma * values * time_outs) is applied first, in place.
 * stats[0..2] (double, zeroed by the caller) receive sum(adv), sum(adv^2), count — all-reduce them across ranks
 * before b200gym_adv_normalize when envs are sharded. */
int b200gym_gae_returns(float* rewards, const float* values, const uint8_t* dones, const uint8_t* time_outs,
                        const float* last_values, float* returns, float* advantages, double* stats, int32_t T, int32_t N,
                        float gamma, float lam, void* stream);
/* advantages = (advantages - mean) / (std_unbiased + 1e-8) from stats (rollout_storage.py, end of compute_returns). */
int b200gym_adv_normalize(float* advantages, const double* stats, int64_t count, void* stream);
/* RolloutStorage.mini_batch_generator gather: dst[k][i, :] = src[k][idx[i], :] for k < n_tensors; row_bytes[k] bytes per row. */
int b200gym_gather_rows(void* const* dst, const void* const* src, const int32_t* row_bytes, int32_t n_tensors, const int64_t* idx,
                        int64_t n_rows, void* stream);
/* PPO.update loss (rsl_rl/algorithms/ppo.py) for one minibatch, forward AND gradient w.r.t. the network outputs:
 * diagonal-Gaussian log-prob / entropy with the learned std[A], KL(old || new), clipped surrogate, clipped value
 * loss, entropy bonus.  Writes d_mu [B,A], d_value [B] (already scaled by 1/B) and accumulates d_std [A] and
 * scalars[0..3] = {sum kl, sum surrogate, sum value_loss, sum entropy} (double, zeroed by the caller). */
typedef struct B200PpoLossParams {
    int32_t batch, num_actions, use_clipped_value_loss, pad;
    float clip_param, value_loss_coef, entropy_coef, inv_global_batch; /* 1 / (batch summed over ranks) */
} B200PpoLossParams;
int b200gym_ppo_loss(const B200PpoLossParams* p, const float* mu, const float* std, const float* value, const float* actions,
                     const float* old_log_prob, const float* advantages, const float* returns, const float* old_values,
                     const float* old_mu, const float* old_sigma, float* d_mu, float* d_value, float* d_std, double* scalars,
                     void* stream);
/* clip_grad_norm_(max_norm) + Adam.step over ONE flat fp32 parameter buffer: sumsq[0] (double, zeroed by the caller)
 * is filled by b200gym_grad_sumsq (all-reduced gradients), then b200gym_clip_adam applies
 * g *= min(1, max_norm / (sqrt(sumsq) + 1e-6)) and the bias-corrected Adam update with learning rate *lr (device). */
int b200gym_grad_sumsq(const float* grad, int64_t n, float grad_scale, double* sumsq, void* stream);
int b200gym_clip_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float grad_scale,
                      const double* sumsq, float max_norm, const float* lr, float beta1, float beta2, float eps, int32_t step,
                      void* stream);
/* The same optimiser step with the Adam step count in DEVICE memory, so that a whole minibatch update (gather, forward,
 * loss, backward, all-reduce, clip, Adam) can be captured once in a CUDA graph and replayed: b200gym_adam_prepare advances
 * *step_dev and zeroes *sumsq (launch it before b200gym_grad_sumsq), b200gym_clip_adam_dev derives the bias corrections
 * 1 - beta^step from *step_dev. */
int b200gym_adam_prepare(int32_t* step_dev, double* sumsq, void* stream);
int b200gym_clip_adam_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float grad_scale,
                          const double* sumsq, float max_norm, const float* lr, float beta1, float beta2, float eps,
                          const int32_t* step_dev, void* stream);
/* Data-parallel PPO.update: out[i] = sum over ranks r (in rank order) of peers->ptr[r][i], i < n_total, where
 * ptr[r] is rank r's flat gradient buffer mapped into this process (peer / symmetric memory over NVLink; ptr[own rank] is the
 * local buffer); sumsq (double, zeroed by the caller) += sum_{i < n_params} out[i]^2, the clip_grad_norm_ input.  Replaces the
 * NCCL all-reduce + b200gym_grad_sumsq pair with one graph-capturable launch; every rank computes bit-identical sums.  The caller
 * synchronises the ranks before (all gradients written) and after (no buffer overwritten while peers read) the call. */
#define B200GYM_MAX_PEERS 16
typedef struct B200PeerPtrs {
    const float* ptr[B200GYM_MAX_PEERS];
} B200PeerPtrs;
int b200gym_grad_reduce_peers(const B200PeerPtrs* peers, int32_t world, float* out, int64_t n_total, int64_t n_params, double* sumsq,
                              void* stream);

/* PPO.update adaptive schedule (ppo.py): kl_mean = *kl_sum / count; lr /= 1.5 if kl_mean > 2*desired_kl (floor 1e-5),
 * lr *= 1.5 if 0 < kl_mean < desired_kl/2 (cap 1e-2).  lr lives on the device, so no host sync per minibatch. */
int b200gym_adaptive_lr(const double* kl_sum, double count, float desired_kl, float* lr, void* stream);

/* ActorCritic MLP forward (rsl_rl/modules/actor_critic.py: Linear + ELU stack, last layer linear) on the tcgen05 tensor
 * cores (fp32 accumulation in TMEM), one launch for the whole stack.  dims[0..num_layers] are the PADDED layer widths
 * (K %% 16 == 0, N %% 16 == 0, N <= 256).  wpacked holds two sections back to back: (1) fp32, layer after layer, W_l [N_l x K_l]
 * as [K_l/4][N_l][4] (zero padded) — operands of the kind::tf32 kernels; (2) fp16, layer after layer, W_l as [K_l/8][N_l][8]
 * — operands of the kind::f16 four-slot kernel (used when every N_l <= 128 and in_dim %% 8 == 0).  bias holds the padded fp32
 * biases back to back.  x: [batch, in_dim] with row stride in_stride (floats); out: [batch, out_dim] contiguous.  The padded
 * weights must fit in shared memory (flat nets do). */
#define B200GYM_MLP_MAX_LAYERS 6
typedef struct B200MlpParams {
    int32_t batch, num_layers, in_dim, in_stride, out_dim, pad;
    int32_t dims[B200GYM_MLP_MAX_LAYERS + 1];
} B200MlpParams;
int b200gym_mlp_forward(const B200MlpParams* p, const float* x, const float* wpacked, const float* bias, float* out, void* stream);
/* Two nets in ONE launch — rsl_rl PPO.act's actor_critic.act(obs) + actor_critic.evaluate(critic_obs) (rsl_rl algorithms/ppo.py,
 * called from legged_gym/utils/task_registry.py's runner; not vendored under the reference tree).  CTAs [0, split) of the grid run
 * net a, the rest net b; results are those of two b200gym_mlp_forward calls.  Both nets must be ones the fp16 four-slot kernel takes
 * (every N_l <= 128, K_l %% 16 == 0, in_dim %% 8 == 0); B200GYM_EINVAL otherwise. */
int b200gym_mlp_forward_pair(const B200MlpParams* pa, const float* xa, const float* wa, const float* ba, float* outa,
                             const B200MlpParams* pb, const float* xb, const float* wb, const float* bb, float* outb, void* stream);
/* Debug aid (tools/trace_mlp.py): buf = device buffer of 5*4096*2 uint64 that CTA 0 of the pipelined forward kernel fills
 * with (event, clock64) pairs; NULL switches tracing off again (the default). */
int b200gym_debug_mlp_trace(void* buf);

/* ------------------------------------------------------------------------------------------------
 * PPO.update training contractions (rsl_rl ppo.py: `loss.backward()` through ActorCritic.actor / .critic)
 * on the tcgen05 tensor cores: ONE grouped fp16 GEMM kernel (fp32 accumulation in TMEM) in three modes.  All matrices are
 * row-major fp16 with leading dimensions that are multiples of 8; m/n/k below are the PADDED extents (multiples of 16 except
 * the row count of FWD/DGRAD and k of WGRAD), *_real the extents that exist in the fp32 parameter / gradient buffers.
 *   FWD    out[m,n]  = act(a[m,k] . b[n,k]^T + bias)      a = activations, b = packed weights W_l [n,k]; bias fp32 [n_real];
 *                                                           flags & 1: ELU, flags & 2: fp32 output (else fp16), ldo in elements
 *   DGRAD  out[m,n]  = (a[m,k] . b[k,n]) * elu'(aux[m,n])  a = dZ_l, b = the SAME packed W_l [k,n] (read MN-major), aux = H_{l-1}
 *                                                           (fp16, elu' = h > 0 ? 1 : h + 1); aux NULL: no derivative factor
 *   WGRAD  out[m,n] += scale * a[k,m]^T . b[k,n]           a = dZ_l [rows,m], b = H_{l-1} [rows,n], k = rows (split over `splits`
 *                                                           CTAs); out fp32 [m_real, n_real] with ldo, accumulated with red.add;
 *                                                           bias (fp32 [m_real], may be NULL) += scale * column sums of a
 * Up to B200GYM_GEMM_MAX_PROBLEMS problems per launch (the actor and the critic layer of the same depth, or every
 * weight-gradient GEMM of an update, share one launch). */
#define B200GYM_GEMM_FWD 0
#define B200GYM_GEMM_DGRAD 1
#define B200GYM_GEMM_WGRAD 2
#define B200GYM_GEMM_MAX_PROBLEMS 8
typedef struct B200GemmProblem {
    const void* a;
    const void* b;
    void* out;
    const void* aux;
    float* bias;
    int32_t mode, flags;
    int32_t m, n, k;
    int32_t lda, ldb, ldo, ldaux;
    int32_t m_real, n_real;
    int32_t splits;
    float scale;
} B200GemmProblem;
int b200gym_gemm_f16(const B200GemmProblem* problems, int32_t n_problems, void* stream);

/* Forward + loss + input-gradient chain of one PPO minibatch in ONE launch, for nets whose 128-row activation tiles fit in
 * shared memory (every layer output <= 128 wide: the 48-128-64-32 flat nets).  Per net (actor, critic) and layer l:
 * kp[l] / np[l] = padded input / output width (multiples of 16, kp[l+1] == np[l], np[last] == 16), n_real[l] = real output
 * width, w_off[l] = offset (halves) of the row-major fp16 copy [np[l], kp[l]] in w16, b_off[l] = offset (floats) of the bias
 * in flat_param.  x: fp16 [batch, ldx] input rows (b200gym_rows_to_f16).  Written: h[l] fp16 [batch, np[l]] (l < last),
 * dz[l] fp16 [batch, np[l]] (all l; unscaled, as in b200gym_ppo_loss_gathered), out (optional fp32 [batch, 16]), d_std and
 * scalars as in b200gym_ppo_loss.  Either follow with WGRAD problems of b200gym_gemm_f16 over (dz[l], h[l-1] | x), or pass
 * flat_grad (below) and the whole of rsl_rl's loss.backward() for the minibatch is this launch. */
#define B200GYM_CHAIN_MAX_LAYERS 6
typedef struct B200ChainNet {
    const void* x;
    const void* w16;
    const float* flat_param;
    void* h[B200GYM_CHAIN_MAX_LAYERS];
    void* dz[B200GYM_CHAIN_MAX_LAYERS];
    float* out;
    int64_t w_off[B200GYM_CHAIN_MAX_LAYERS], b_off[B200GYM_CHAIN_MAX_LAYERS];
    int32_t kp[B200GYM_CHAIN_MAX_LAYERS], np[B200GYM_CHAIN_MAX_LAYERS], n_real[B200GYM_CHAIN_MAX_LAYERS];
    int32_t num_layers, ldx;
    /* optional: x32 != NULL gathers the input rows in the kernel — fp32 [*, ldx32] rows taken through idx, k_real[0] columns — instead
     * of reading the fp16 copy x.  flat_grad != NULL also computes the weight and bias gradients in the kernel: per 128-row tile
     * dW_l = dZ_l^T . H_{l-1} / batch is added (fp32 red) to flat_grad + w32_off[l] ([n_real[l], k_real[l]] row-major, k_real = real
     * input width) and sum_rows dZ_l / batch to flat_grad + b_off[l]; h / dz may then be NULL (nothing but gradients leaves the SM). */
    const float* x32;
    float* flat_grad;
    int64_t w32_off[B200GYM_CHAIN_MAX_LAYERS];
    int32_t k_real[B200GYM_CHAIN_MAX_LAYERS];
    int32_t ldx32;
    int32_t w_layout; /* layout of the fp16 weights in w16 (B200PackEntry.layout): 0 row-major [np, kp] (cp.async pieces), 1 chunk-major
                       * [kp/8][np][8] (TMA bulk copies: a forward stage is 8 pieces of np*16 B, a backward stage kp/8 pieces of 1 KB) */
} B200ChainNet;
int b200gym_ppo_chain(const B200ChainNet* actor, const B200ChainNet* critic, const B200PpoLossParams* lp, const int64_t* idx,
                      const float* std, const float* actions, const float* old_log_prob, const float* advantages,
                      const float* returns, const float* old_values, const float* old_mu, const float* old_sigma, float* d_std,
                      double* scalars, void* stream);
/* Debug aid (tools/trace_chain.py): buf = device buffer of 6*256*2 uint64 that the first and last CTA of the chain kernel fill with
 * (event, clock64) pairs; NULL switches tracing off again (the default). */
int b200gym_debug_chain_trace(void* buf);
/* dst[i, 0:dst_ld] (fp16) = src[idx ? idx[i] : i, 0:cols] (fp32, row stride src_ld), zero padded to dst_ld (a multiple of 8):
 * the observation gather of RolloutStorage.mini_batch_generator fused with the operand conversion of the first layer. */
int b200gym_rows_to_f16(const float* src, int64_t src_ld, int32_t cols, const int64_t* idx, void* dst, int32_t dst_ld, int64_t n_rows,
                        void* stream);
/* b200gym_ppo_loss with (a) the per-sample storage columns read through the minibatch indices idx (NULL = identity) straight
 * from the [T*N, .] storage tensors, (b) mu / value read from the fp32 outputs of the last FWD GEMMs (row strides ld_mu /
 * ld_value, value in column 0), (c) the gradients written as the fp16 [batch, 16] dZ operands of the backward GEMMs, WITHOUT
 * the 1/batch factor (pass it as `scale` of the WGRAD problems); d_std and scalars as in b200gym_ppo_loss. */
int b200gym_ppo_loss_gathered(const B200PpoLossParams* p, const int64_t* idx, const float* mu_out, int32_t ld_mu, const float* value_out,
                              int32_t ld_value, const float* std, const float* actions, const float* old_log_prob, const float* advantages,
                              const float* returns, const float* old_values, const float* old_mu, const float* old_sigma, void* dz_actor,
                              void* dz_critic, float* d_std, double* scalars, void* stream);
/* fp32 master parameters (one flat buffer) -> fp16 operand copies, every layer in one launch.  Entry e copies the [rows, cols]
 * matrix at flat[src_off] to dst[dst_off] (in halves) as layout 0: row-major with leading dimension ld, or layout 1: the
 * chunk-major [cols/8][ld][8] form of b200gym_mlp_forward's fp16 section.  elem_end = running total of rows*cols. */
#define B200GYM_PACK_MAX 16
typedef struct B200PackEntry {
    int64_t src_off, dst_off, elem_end;
    int32_t rows, cols, ld, layout;
} B200PackEntry;
typedef struct B200PackTable {
    B200PackEntry e[B200GYM_PACK_MAX];
    int64_t total;
    int32_t n, pad;
} B200PackTable;
int b200gym_pack_params_f16(const float* flat, const B200PackTable* table, void* dst, void* stream);
/* Everything of a PPO minibatch step after the backward pass in ONE launch (rsl_rl PPO.update: adaptive-KL schedule,
 * clip_grad_norm_, Adam.step) + the fp16 operand copies of the updated weights (b200gym_pack_params_f16's table) + the
 * clearing of grad[0 .. n+8) and of the per-minibatch sums for the next minibatch.  mb_scalars[0..3] = this minibatch's
 * {sum kl, sum surrogate, sum value loss, sum entropy} (b200gym_ppo_chain / _ppo_loss_gathered), added to totals[0..3] and
 * zeroed; totals[4] = squared gradient norm of this step.  count = minibatch samples (KL mean denominator).  lr and the step
 * count live on the device.  workspace: 16 zero-initialised bytes, owned by the call sequence (reset on exit). */
typedef struct B200OptParams {
    int64_t n;
    double count;
    int32_t adaptive, pad;
    float desired_kl, max_grad_norm, beta1, beta2, eps, pad2;
} B200OptParams;
int b200gym_ppo_optimizer_step(const B200OptParams* p, float* param, float* grad, float* exp_avg, float* exp_avg_sq, float* lr,
                               int32_t* step_dev, double* mb_scalars, double* totals, void* workspace, const B200PackTable* table, void* w16,
                               void* stream);

/* The data-parallel form of b200gym_ppo_optimizer_step (replaces rsl_rl-style DDP: NCCL all-reduce of the
 * gradients + clip_grad_norm_ + Adam.step): ONE launch per rank exchanges the flat gradient buffers over NVLink peer memory
 * (push into every rank's symmetric buffer, flag signalling, no host / torch barrier), sums them in rank order (bit-identical on
 * every rank), and performs the step.  peers->base[r] = rank r's mapping of its symmetric buffer of 2 * world * n_pad floats
 * (slots [parity][source rank][n_pad]) followed by world uint32 flags, all zero-initialised before the first call;
 * n_pad = multiple of 4 in [n + 2, n + 8]; grad: this rank's LOCAL flat gradients (n + 8 floats, cleared on exit);
 * grad_sum: n_pad floats of scratch that receive the summed gradients (+ {sum kl, count} at [n], [n + 1]); the KL mean of the
 * adaptive schedule is the GLOBAL one.  workspace: B200GYM_PEER_WS_BYTES zero-initialised bytes owned by the call sequence
 * (holds the exchange number: every rank must make the same sequence of calls).  ctas: 0 = pick from n.  A rank that waits
 * more than 4 s for a peer sets workspace word 4 (error) instead of hanging. */
#define B200GYM_OPT_MAX_CTAS 160
#define B200GYM_PEER_WS_BYTES (32 + 8 * B200GYM_OPT_MAX_CTAS)
typedef struct B200PeerBases {
    float* base[B200GYM_MAX_PEERS];
} B200PeerBases;
int b200gym_ppo_optimizer_step_peers(const B200OptParams* p, const B200PeerBases* peers, int32_t world, int32_t rank, int64_t n_pad, float* param,
                                     float* grad, float* grad_sum, float* exp_avg, float* exp_avg_sq, float* lr, int32_t* step_dev,
                                     double* mb_scalars, double* totals, void* workspace, const B200PackTable* table, void* w16, int32_t ctas,
                                     void* stream);

/* Everything of rsl_rl PPO.act after the actor / critic forward, in ONE launch (rsl_rl/algorithms/ppo.py act(): Normal(mu, std)
 * .sample(), get_actions_log_prob, and the transition fields; rsl_rl/storage/rollout_storage.py add_transitions' copies): the
 * action sample (Philox4x32-10, site POLICY_SAMPLE, counter = (env_id_offset + env, event); Box-Muller, specification
 * oracle/port_ppo.py sample_actions), its summed log-prob, and observations / critic observations / actions / values / log-prob /
 * mu / sigma written to the storage ROW pointers st_* (row `step` of the [T, N, .] tensors).  mu_out [n_envs, ld_mu] and
 * value_out [n_envs, ld_value] are the MLP outputs; st_critic_obs may be NULL (no privileged observations).  event_dev (optional):
 * the act counter in device memory, read instead of `event` and advanced by b200gym_ppo_store_step — a whole rollout can then be
 * captured in a CUDA graph and replayed with fresh draws. */
int b200gym_ppo_act_store(int32_t n_envs, int32_t num_actions, int32_t num_obs, int32_t num_critic_obs, const float* mu_out, int32_t ld_mu,
                          const float* value_out, int32_t ld_value, const float* std, const float* obs, int64_t ld_obs, const float* critic_obs,
                          int64_t ld_critic_obs, uint64_t seed, uint64_t event, const uint64_t* event_dev, uint64_t env_id_offset, float* st_obs,
                          float* st_critic_obs, float* st_actions, float* st_values, float* st_log_prob, float* st_mu, float* st_sigma, void* stream);
/* PPO.process_env_step -> add_transitions: rewards / dones / time-out flags (byte tensors; time_outs may be NULL = none) of one env
 * step into the storage row pointers. */
int b200gym_ppo_store_step(int32_t n_envs, const float* rewards, const uint8_t* dones, const uint8_t* time_outs, float* st_rewards,
                           uint8_t* st_dones, uint8_t* st_time_outs, uint64_t* event_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Tube-dataset construction from the rollout logs (deep_tube_learning/datasets.py:60-71,
 * deep_tube_learning/evaluation/evaluate_tube_simple.py:28-46) — SURVEY.md §8f row 2
 * ---------------------------------------------------------------------------------------------- */
/* w[b,t] = || pz_x[b,t,:] - z[b,t,:] ||_2 for t < T; z, pz_x: [B, T1, n] logs of b200gym_rom_rollout (T1 = T + 1 there). */
int b200gym_tube_error(const float* z, const float* pz_x, float* w, int64_t B, int32_t T, int32_t T1, int32_t n, void* stream);
/* sliding_window(data, N, dN, m) (datasets.py:69-71): data [B,T,D] -> out [B,T,N*D], slice i = get_slice(data, i, dN, m). */
int b200gym_sliding_window(const float* data, float* out, int64_t B, int32_t T, int32_t D, int32_t N, int32_t dN, int32_t m, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200GYM_H */
