// the HopperTrajectory env step around the physics call (legged_gym/envs/hopper/hopper_trajectory.py:135-182 with what
// the class inherits from legged_gym/envs/base/legged_robot_trajectory.py: check_termination :194-202, reset_idx :204-246, compute_reward
// :255-272, the shared _reward_* terms :1000-1110).  The torque law of the sub-step loop is hopper.cu's; the trajectory generator is rom.cu's
// (b200gym_rom_step before, b200gym_rom_reset_from_root after, as for the ANYmal trajectory env).  Three launches per env step:
//
//   hopper_prologue_kernel      body-frame velocities and projected gravity from the root state the last sub-step left (:145-146, :126), the
//                               per-env push timers and the pushes themselves (:149-165, _push_robots :362-367).  `time_until_next_push` is
//                               [N, 1] in the reference and HopperTrajectory does not flatten its mask (:149,152): `nonzero().flatten()`
//                               interleaves the env ids with zeros, so env 0 is pushed whenever ANY env of the process is — reproduced
//                               through a device flag that env 0's thread of the next kernel consumes
//   hopper_post_physics_kernel  one thread per env: termination, the reward terms in the reference's (alphabetical) order with their episode
//                               sums, in-place reset (_reset_dofs / _reset_root_states :298-358: the HOP_* Philox sites, yaw randomisation
//                               through pytorch3d's euler -> matrix -> quaternion -> multiply chain), observations + noise + clip
//                               (:255-282, :128-129), last_* / prev_error epilogue (:175-178), extras["episode"] partial sums
//   hopper_extras_finalize      the means of legged_robot_trajectory.py:235-239
//
// Quirks of the reference that the parity depends on are listed in oracle/port_hopper_env.py (the specification, pinned to the unmodified
// class): base_quat is a view (post-reset quaternion in the observation) while base_lin_vel / base_ang_vel are pre-reset buffers;
// _reset_dofs overwrites actions with zero_action; prev_error is recomputed for all envs from the callback's trajectory clone; the reward
// terms see the limit-clipped torques.
//
// Algorithmic bytes per env (yaml reward table, K = 11 sum rows): read root 52 + dof 32 + contacts 60 + actions 16 + torques 16 + last_actions 16
// + last_dof_vel 16 + trajectory 80 + gen.v 8 + prev_error 8 + timer 4 + ep 8 + sums 4K + bav 12; written obs 152 + rew 4 + flags 2 + ep 8 +
// last_* 56 + prev_error 8 + blv/pg/bav 36 + timer 4 + sums 4K  ->  ~ 682 B.
#include <math.h>

#include "common.cuh"
#include "philox.cuh"
#include "../../include/b200gym.h"

namespace {

enum HopTerm {
    H_ACTION_RATE = 0, H_ANG_VEL_XY, H_BASE_HEIGHT, H_COLLISION, H_DIFF_ERROR, H_DOF_ACC, H_DOF_POS_LIMITS, H_DOF_VEL, H_DOF_VEL_LIMITS,
    H_FEET_AIR_TIME, H_FEET_CONTACT_FORCES, H_LIN_VEL_Z, H_ORIENTATION, H_RAIBERT, H_STUMBLE, H_TORQUE_LIMITS, H_TORQUES, H_TRACKING_ROM,
    H_UNIT_QUAT, H_TERMINATION
};
static_assert(H_TERMINATION + 1 == B200GYM_HOPPER_NUM_TERMS, "term table out of sync with the header");

__device__ __forceinline__ void rotate_inverse(float qx, float qy, float qz, float qw, float vx, float vy, float vz, float& ox, float& oy, float& oz) {
    // isaacgym.torch_utils.quat_rotate_inverse: a = v (2 w^2 - 1), b = cross(q_vec, v) * 2 w, c = q_vec (q_vec . v) 2; a - b + c
    const float k = sub_rn(mul_rn(mul_rn(2.0f, qw), qw), 1.0f);
    const float ax = mul_rn(vx, k), ay = mul_rn(vy, k), az = mul_rn(vz, k);
    const float w2 = mul_rn(qw, 2.0f);
    const float bx = mul_rn(sub_rn(mul_rn(qy, vz), mul_rn(qz, vy)), w2), by = mul_rn(sub_rn(mul_rn(qz, vx), mul_rn(qx, vz)), w2);
    const float bz = mul_rn(sub_rn(mul_rn(qx, vy), mul_rn(qy, vx)), w2);
    const float d = add_rn(add_rn(mul_rn(qx, vx), mul_rn(qy, vy)), mul_rn(qz, vz));
    const float cx = mul_rn(mul_rn(qx, d), 2.0f), cy = mul_rn(mul_rn(qy, d), 2.0f), cz = mul_rn(mul_rn(qz, d), 2.0f);
    ox = add_rn(sub_rn(ax, bx), cx), oy = add_rn(sub_rn(ay, by), cy), oz = add_rn(sub_rn(az, bz), cz);
}

__device__ __forceinline__ void push_draw(const B200HopperEnvParams& p, const philox::Stream& rng, float* root13) {
    const uint4 w0 = rng.words(philox::HOP_PUSH, 0), w1 = rng.words(philox::HOP_PUSH, 1);
    const float u[6] = {philox::u01(w0.x), philox::u01(w0.y), philox::u01(w0.z), philox::u01(w0.w), philox::u01(w1.x), philox::u01(w1.y)};
#pragma unroll
    for (int k = 0; k < 6; ++k) root13[7 + k] = affine_rn(sub_rn(p.max_push_vel[k], -p.max_push_vel[k]), u[k], -p.max_push_vel[k]);
}

__global__ void __launch_bounds__(256) hopper_prologue_kernel(const __grid_constant__ B200HopperEnvParams p, const __grid_constant__ B200HopperEnvBuffers b,
                                                              unsigned long long step, long long env_off) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.num_envs) return;
    const size_t i = static_cast<size_t>(e);
    float* r = b.root_states + i * 13;
    const float qx = r[3], qy = r[4], qz = r[5], qw = r[6];
    float x, y, z;
    rotate_inverse(qx, qy, qz, qw, r[7], r[8], r[9], x, y, z);
    b.base_lin_vel[i * 3] = x, b.base_lin_vel[i * 3 + 1] = y, b.base_lin_vel[i * 3 + 2] = z;
    rotate_inverse(qx, qy, qz, qw, r[10], r[11], r[12], x, y, z);
    b.base_ang_vel[i * 3] = x, b.base_ang_vel[i * 3 + 1] = y, b.base_ang_vel[i * 3 + 2] = z;
    rotate_inverse(qx, qy, qz, qw, 0.0f, 0.0f, -1.0f, x, y, z);
    b.projected_gravity[i * 3] = x, b.projected_gravity[i * 3 + 1] = y, b.projected_gravity[i * 3 + 2] = z;
    const float t = sub_rn(b.time_until_next_push[i], p.push_dt);
    float t_out = t;
    if (p.push_robots && t <= 0.0f) {
        const philox::Stream rng(p.seed_lo, p.seed_hi, static_cast<unsigned long long>(env_off) + i, step);
        push_draw(p, rng, r);
        t_out = affine_rn(p.push_t_span, philox::u01(rng.words(philox::PUSH_TIMER, 0).x), p.push_t_lo);
        atomicOr(b.push_flag, 1u);
    }
    b.time_until_next_push[i] = t_out;
}

// HopperTrajectory._reset_dofs + _reset_root_states for env i (hopper_trajectory.py:298-358): dof / action / root redraw at the HOP_* Philox
// sites, yaw randomisation through pytorch3d's euler -> matrix -> quaternion -> multiply chain; writes dof_state, actions, root_states and
// leaves the new values in q / qd / act / R.  Shared by the in-step reset and the external reset_idx launch.
__device__ __forceinline__ void hopper_reset_env(const B200HopperEnvParams& p, const B200HopperEnvBuffers& b, const philox::Stream& rng, size_t i,
                                                 float (&q)[4], float (&qd)[4], float (&act)[4], float (&R)[13]) {
    const uint4 wp = rng.words(philox::HOP_DOF_POS, 0), wv = rng.words(philox::HOP_DOF_VEL, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        q[k] = add_rn(p.default_dof_pos[k], affine_rn(p.dof_pos_noise_span[k], philox::u01(philox::word(wp, k)), p.dof_pos_noise_lo[k]));
        qd[k] = affine_rn(p.dof_vel_noise_span[k], philox::u01(philox::word(wv, k)), p.dof_vel_noise_lo[k]);
        act[k] = p.zero_action[k];                                                      // :314
    }
    *reinterpret_cast<float4*>(b.dof_state + i * 8) = make_float4(q[0], qd[0], q[1], qd[1]);
    *reinterpret_cast<float4*>(b.dof_state + i * 8 + 4) = make_float4(q[2], qd[2], q[3], qd[3]);
    *reinterpret_cast<float4*>(b.actions + i * 4) = make_float4(act[0], act[1], act[2], act[3]);
#pragma unroll
    for (int k = 0; k < 13; ++k) R[k] = p.base_init_state[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) R[k] = add_rn(R[k], b.env_origins[i * 3 + k]);
    const uint4 w0 = rng.words(philox::HOP_ROOT_POS, 0), w1 = rng.words(philox::HOP_ROOT_POS, 1);
    const float ur[5] = {philox::u01(w0.x), philox::u01(w0.y), philox::u01(w0.z), philox::u01(w0.w), philox::u01(w1.x)};
#pragma unroll
    for (int k = 0; k < 5; ++k) R[2 + k] = add_rn(R[2 + k], affine_rn(p.root_pos_noise_span[k], ur[k], p.root_pos_noise_lo[k]));   // :338-341
    const float nq = sqrtf(add_rn(add_rn(add_rn(mul_rn(R[3], R[3]), mul_rn(R[4], R[4])), mul_rn(R[5], R[5])), mul_rn(R[6], R[6])));
#pragma unroll
    for (int k = 3; k < 7; ++k) R[k] = div_rn(R[k], nq);
    if (p.randomize_yaw) {   // :343-348, pytorch3d: euler_angles_to_matrix([0, 0, yaw], "XYZ") = Rz(yaw) -> matrix_to_quaternion -> multiply
        const float pi = 3.14159265358979323846f;
        const float yaw = affine_rn(sub_rn(pi, -pi), philox::u01(rng.words(philox::HOP_YAW, 0).x), -pi);
        const float c = cosf(yaw), s = sinf(yaw);
        const float a0 = add_rn(add_rn(add_rn(1.0f, c), c), 1.0f), a3 = add_rn(sub_rn(sub_rn(1.0f, c), c), 1.0f);
        const float a1 = sub_rn(sub_rn(add_rn(1.0f, c), c), 1.0f), a2 = sub_rn(add_rn(sub_rn(1.0f, c), c), 1.0f);
        const float q0 = a0 > 0.f ? sqrtf(a0) : 0.f, q1 = a1 > 0.f ? sqrtf(a1) : 0.f, q2 = a2 > 0.f ? sqrtf(a2) : 0.f, q3 = a3 > 0.f ? sqrtf(a3) : 0.f;
        const float two_s = sub_rn(s, -s);   // m10 - m01
        float yw, yx = 0.0f, yy = 0.0f, yz;
        int best = 0;
        float bq = q0;
        if (q1 > bq) best = 1, bq = q1;
        if (q2 > bq) best = 2, bq = q2;
        if (q3 > bq) best = 3, bq = q3;
        const float den = mul_rn(2.0f, fmaxf(bq, 0.1f));
        if (best == 0) yw = div_rn(mul_rn(q0, q0), den), yz = div_rn(two_s, den);
        else if (best == 3) yw = div_rn(two_s, den), yz = div_rn(mul_rn(q3, q3), den);
        else if (best == 1) yw = 0.0f, yx = div_rn(mul_rn(q1, q1), den), yy = div_rn(add_rn(s, -s), den), yz = 0.0f;
        else yw = 0.0f, yx = div_rn(add_rn(s, -s), den), yy = div_rn(mul_rn(q2, q2), den), yz = 0.0f;
        if (yw < 0.0f) yw = -yw, yx = -yx, yy = -yy, yz = -yz;      // standardize_quaternion
        const float aw = R[6], ax = R[3], ay = R[4], az = R[5];      // wxyz_quat_inds
        float ow = sub_rn(sub_rn(sub_rn(mul_rn(aw, yw), mul_rn(ax, yx)), mul_rn(ay, yy)), mul_rn(az, yz));
        float ox = sub_rn(add_rn(add_rn(mul_rn(aw, yx), mul_rn(ax, yw)), mul_rn(ay, yz)), mul_rn(az, yy));
        float oy = add_rn(add_rn(sub_rn(mul_rn(aw, yy), mul_rn(ax, yz)), mul_rn(ay, yw)), mul_rn(az, yx));
        float oz = add_rn(sub_rn(add_rn(mul_rn(aw, yz), mul_rn(ax, yy)), mul_rn(ay, yx)), mul_rn(az, yw));
        if (ow < 0.0f) ow = -ow, ox = -ox, oy = -oy, oz = -oz;
        R[3] = ox, R[4] = oy, R[5] = oz, R[6] = ow;
    }
    const uint4 v0 = rng.words(philox::HOP_ROOT_VEL, 0), v1 = rng.words(philox::HOP_ROOT_VEL, 1);
    const float uv[6] = {philox::u01(v0.x), philox::u01(v0.y), philox::u01(v0.z), philox::u01(v0.w), philox::u01(v1.x), philox::u01(v1.y)};
#pragma unroll
    for (int k = 0; k < 6; ++k) R[7 + k] = affine_rn(p.root_vel_noise_span[k], uv[k], p.root_vel_noise_lo[k]);
#pragma unroll
    for (int k = 0; k < 13; ++k) b.root_states[i * 13 + k] = R[k];
}

// The row-strided tensors of a CTA's 128 envs (root [., 13], contacts [., B, 3], trajectory [., 20] in; observations [., 38] out) are contiguous
// byte ranges: they move between HBM and shared memory with coalesced cooperative copies (odd row strides in shared memory: conflict-free
// per-env access), the [., 4] tensors are read as one float4 per thread.  (First version: every thread walked its own rows in global memory —
// 13 + 15 + 20 strided loads and 38 strided stores per env, 40 % of HBM.)
#ifndef HOPPER_MINBLOCKS
#define HOPPER_MINBLOCKS 5   // resident CTAs per SM the register allocation aims at: 90 registers, no spills; 45.6 KB of shared memory still fit 5 (A/B: profiles/r2_hopper_occupancy.txt)
#endif
constexpr int HT_TILE = 128, HS_TRAJ = B200GYM_TRAJ_WIDTH + 1, HS_OBS = B200GYM_HOPPER_TRAJ_NUM_OBS + 1;

__global__ void __launch_bounds__(HT_TILE, HOPPER_MINBLOCKS) hopper_post_physics_kernel(const __grid_constant__ B200HopperEnvParams p,
                                                                        const __grid_constant__ B200HopperEnvBuffers b, unsigned long long step,
                                                                        long long env_off) {
    const int N = p.num_envs, K = p.num_sum_rows, B = p.num_bodies;
    const int tile0 = blockIdx.x * HT_TILE, t = threadIdx.x;
    const int e = tile0 + t, nvalid = min(HT_TILE, N - tile0);
    extern __shared__ __align__(16) float hsm[];
    float* s_root = hsm;                                   // [128][13]
    float* s_contact = s_root + HT_TILE * 13;              // [128][3B]
    float* s_traj = s_contact + HT_TILE * 3 * B;           // [128][21]
    float* s_obs = s_traj + HT_TILE * HS_TRAJ;             // [128][39]
    __shared__ double s_acc[B200GYM_HOPPER_NUM_TERMS + 2];
    if (t < K + 2) s_acc[t] = 0.0;
    {
        const float* g = b.root_states + static_cast<size_t>(tile0) * 13;
        for (int q = t; q < nvalid * 13; q += HT_TILE) s_root[q] = g[q];
        g = b.contact_forces + static_cast<size_t>(tile0) * 3 * B;
        for (int q = t; q < nvalid * 3 * B; q += HT_TILE) s_contact[q] = g[q];
        g = b.trajectory + static_cast<size_t>(tile0) * B200GYM_TRAJ_WIDTH;
        for (int q = t; q < nvalid * B200GYM_TRAJ_WIDTH; q += HT_TILE) s_traj[(q / B200GYM_TRAJ_WIDTH) * HS_TRAJ + q % B200GYM_TRAJ_WIDTH] = g[q];
    }
    __syncthreads();
    if (e < N) {
        const size_t i = static_cast<size_t>(e);
        const philox::Stream rng(p.seed_lo, p.seed_hi, static_cast<unsigned long long>(env_off) + i, step);
        float R[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) R[k] = s_root[t * 13 + k];
        if (e == 0 && p.push_robots && *b.push_flag != 0u) {   // the [N, 1]-mask quirk: env 0 rides along with every push of the process
            push_draw(p, rng, R);
#pragma unroll
            for (int k = 7; k < 13; ++k) b.root_states[k] = R[k];
            *b.push_flag = 0u;
        }
        const float4 d0 = *reinterpret_cast<const float4*>(b.dof_state + i * 8), d1 = *reinterpret_cast<const float4*>(b.dof_state + i * 8 + 4);
        float q[4] = {d0.x, d0.z, d1.x, d1.z}, qd[4] = {d0.y, d0.w, d1.y, d1.w};
        const float4 a4 = *reinterpret_cast<const float4*>(b.actions + i * 4);
        float act[4] = {a4.x, a4.y, a4.z, a4.w};
        const float4 t4 = *reinterpret_cast<const float4*>(b.torques + i * 4);
        const float tq[4] = {t4.x, t4.y, t4.z, t4.w};
        const float4 la4 = *reinterpret_cast<const float4*>(b.last_actions + i * 4), lv4 = *reinterpret_cast<const float4*>(b.last_dof_vel + i * 4);
        const float la[4] = {la4.x, la4.y, la4.z, la4.w}, lv[4] = {lv4.x, lv4.y, lv4.z, lv4.w};
        const float blv[3] = {b.base_lin_vel[i * 3], b.base_lin_vel[i * 3 + 1], b.base_lin_vel[i * 3 + 2]};
        const float bav[3] = {b.base_ang_vel[i * 3], b.base_ang_vel[i * 3 + 1], b.base_ang_vel[i * 3 + 2]};
        const float pg[3] = {b.projected_gravity[i * 3], b.projected_gravity[i * 3 + 1], b.projected_gravity[i * 3 + 2]};
        const float* traj = s_traj + t * HS_TRAJ;
        const float tr0x = traj[0], tr0y = traj[1];
        const float* cf = s_contact + t * B * 3;
        long long ep = reinterpret_cast<long long*>(b.episode_length_buf)[i] + 1;

        // termination (legged_robot_trajectory.py:194-202)
        bool term = false;
        for (int k = 0; k < p.num_term; ++k) {
            const float* f = cf + p.term_idx[k] * 3;
            term |= sqrtf(add_rn(add_rn(mul_rn(f[0], f[0]), mul_rn(f[1], f[1])), mul_rn(f[2], f[2]))) > 1.0f;
        }
        const bool time_out = static_cast<float>(ep) > p.max_episode_length;
        const bool reset = term | time_out;

        // rewards (legged_robot_trajectory.py:255-272), alphabetical order of the active terms
        // the per-term values stay in registers; the episode_sums rows are read-modify-written once, after every load of the step has
        // been issued (a global store per term in between would order the remaining loads behind it: 72 % long-scoreboard stalls)
        const float* rs = p.reward_scale;
        float rew = 0.0f;
        float tv[B200GYM_HOPPER_NUM_TERMS];
#pragma unroll
        for (int k = 0; k < B200GYM_HOPPER_NUM_TERMS; ++k) tv[k] = 0.0f;
        auto add_term = [&](int k, float val) {
            const float rr = mul_rn(val, rs[k]);
            rew = add_rn(rew, rr);
            tv[k] = rr;
        };
        const float ex = sub_rn(R[0], tr0x), ey = sub_rn(R[1], tr0y);
        const float te0 = mul_rn(ex, ex), te1 = mul_rn(ey, ey);   // square(proj_z(root) - trajectory[:, 0])
        if (rs[H_ACTION_RATE] != 0.f) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) s = add_rn(s, mul_rn(sub_rn(la[k], act[k]), sub_rn(la[k], act[k])));
            add_term(H_ACTION_RATE, s);
        }
        if (rs[H_ANG_VEL_XY] != 0.f) add_term(H_ANG_VEL_XY, add_rn(mul_rn(bav[0], bav[0]), mul_rn(bav[1], bav[1])));
        if (rs[H_BASE_HEIGHT] != 0.f) {
            const float d = sub_rn(R[2], p.base_height_target);
            add_term(H_BASE_HEIGHT, mul_rn(d, d));
        }
        if (rs[H_COLLISION] != 0.f) {
            float c = 0.f;
            for (int k = 0; k < p.num_pen; ++k) {
                const float* f = cf + p.pen_idx[k] * 3;
                c += sqrtf(add_rn(add_rn(mul_rn(f[0], f[0]), mul_rn(f[1], f[1])), mul_rn(f[2], f[2]))) > 0.1f ? 1.0f : 0.0f;
            }
            add_term(H_COLLISION, c);
        }
        if (rs[H_DIFF_ERROR] != 0.f) {
            const float err = sqrtf(add_rn(mul_rn(te0, te0), mul_rn(te1, te1)));
            const float pe0 = b.prev_error[i * 2], pe1 = b.prev_error[i * 2 + 1];
            const float de = sub_rn(err, sqrtf(add_rn(mul_rn(pe0, pe0), mul_rn(pe1, pe1))));
            add_term(H_DIFF_ERROR, mul_rn(de < 0.0f ? p.diff_neg_slope : p.diff_pos_slope, de));
        }
        if (rs[H_DOF_ACC] != 0.f) {   // wheels only (hopper_trajectory.py:474-476)
            float s = 0.f;
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                const float a = div_rn(sub_rn(lv[k], qd[k]), p.dt);
                s = add_rn(s, mul_rn(a, a));
            }
            add_term(H_DOF_ACC, s);
        }
        if (rs[H_DOF_POS_LIMITS] != 0.f) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) s = add_rn(s, add_rn(-fminf(sub_rn(q[k], p.dof_pos_lo[k]), 0.0f), fmaxf(sub_rn(q[k], p.dof_pos_hi[k]), 0.0f)));
            add_term(H_DOF_POS_LIMITS, s);
        }
        if (rs[H_DOF_VEL] != 0.f) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) s = add_rn(s, mul_rn(qd[k], qd[k]));
            add_term(H_DOF_VEL, s);
        }
        if (rs[H_DOF_VEL_LIMITS] != 0.f) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) s = add_rn(s, fminf(fmaxf(sub_rn(fabsf(qd[k]), mul_rn(p.dof_vel_limits[k], p.soft_dof_vel_limit)), 0.0f), 1.0f));
            add_term(H_DOF_VEL_LIMITS, s);
        }
        const float* ff = cf + p.foot_body * 3;
        float fat = b.feet_air_time[i];
        if (rs[H_FEET_AIR_TIME] != 0.f) {   // legged_robot_trajectory.py:1071-1083, one foot, no command gate
            const bool contact = ff[2] > 1.0f;
            const bool filt = contact | (b.last_contacts[i] != 0);
            b.last_contacts[i] = contact ? 1 : 0;
            const bool first = (fat > 0.0f) && filt;
            fat = add_rn(fat, p.dt);
            add_term(H_FEET_AIR_TIME, first ? sub_rn(fat, 0.5f) : 0.0f);
            if (filt) fat = 0.0f;
        }
        if (rs[H_FEET_CONTACT_FORCES] != 0.f)
            add_term(H_FEET_CONTACT_FORCES,
                     fmaxf(sub_rn(sqrtf(add_rn(add_rn(mul_rn(ff[0], ff[0]), mul_rn(ff[1], ff[1])), mul_rn(ff[2], ff[2]))), p.max_contact_force), 0.0f));
        if (rs[H_LIN_VEL_Z] != 0.f) add_term(H_LIN_VEL_Z, mul_rn(blv[2], blv[2]));
        if (rs[H_ORIENTATION] != 0.f) add_term(H_ORIENTATION, add_rn(mul_rn(pg[0], pg[0]), mul_rn(pg[1], pg[1])));
        if (rs[H_RAIBERT] != 0.f) {   // hopper_trajectory.py:482-505 + RaibertHeuristic.raibert_policy (controllers.py:38-73)
            float cvx, cvy, cvz;
            rotate_inverse(R[3], R[4], R[5], R[6], R[7], R[8], R[9], cvx, cvy, cvz);
            const float Kp = p.raibert[0], Kv = p.raibert[1], Kff = p.raibert[2], cpos = p.raibert[3], cvel = p.raibert[4], cang = p.raibert[5];
            const float pex = sub_rn(tr0x, R[0]), pey = -sub_rn(tr0y, R[1]), evx = -cvx, evy = cvy;
            const float dvx = b.gen_v[i * 2], dvy = -b.gen_v[i * 2 + 1];
            const float pitch_pos = clampf(mul_rn(-Kp, pex), -cpos, cpos), roll_pos = clampf(mul_rn(-Kp, pey), -cpos, cpos);
            const float vx = clampf(add_rn(mul_rn(-Kv, evx), mul_rn(Kff, dvx)), -cvel, cvel);
            const float vy = clampf(add_rn(mul_rn(-Kv, evy), mul_rn(Kff, dvy)), -cvel, cvel);
            const float pitch = clampf(add_rn(pitch_pos, vx), -cang, cang), roll = clampf(add_rn(roll_pos, vy), -cang, cang);
            const float yaw = atan2f(mul_rn(2.0f, add_rn(mul_rn(R[6], R[5]), mul_rn(R[3], R[4]))),
                                     sub_rn(1.0f, mul_rn(2.0f, add_rn(mul_rn(R[4], R[4]), mul_rn(R[5], R[5])))));
            const float cy = cosf(yaw * 0.5f), sy = sinf(yaw * 0.5f), cp = cosf(pitch * 0.5f), sp = sinf(pitch * 0.5f);
            const float cr = cosf(roll * 0.5f), sr = sinf(roll * 0.5f);
            const float rh[4] = {add_rn(mul_rn(mul_rn(cr, cp), cy), mul_rn(mul_rn(sr, sp), sy)), sub_rn(mul_rn(mul_rn(sr, cp), cy), mul_rn(mul_rn(cr, sp), sy)),
                                 add_rn(mul_rn(mul_rn(cr, sp), cy), mul_rn(mul_rn(sr, cp), sy)), sub_rn(mul_rn(mul_rn(cr, cp), sy), mul_rn(mul_rn(sr, sp), cy))};
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) s = add_rn(s, mul_rn(sub_rn(act[k], rh[k]), sub_rn(act[k], rh[k])));
            add_term(H_RAIBERT, s);
        }
        if (rs[H_STUMBLE] != 0.f)
            add_term(H_STUMBLE, sqrtf(add_rn(mul_rn(ff[0], ff[0]), mul_rn(ff[1], ff[1]))) > mul_rn(5.0f, fabsf(ff[2])) ? 1.0f : 0.0f);
        if (rs[H_TORQUE_LIMITS] != 0.f) add_term(H_TORQUE_LIMITS, add_rn(add_rn(fabsf(tq[1]), fabsf(tq[2])), fabsf(tq[3])));   // :470-472
        if (rs[H_TORQUES] != 0.f) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) s = add_rn(s, mul_rn(tq[k], tq[k]));
            add_term(H_TORQUES, s);
        }
        if (rs[H_TRACKING_ROM] != 0.f)
            add_term(H_TRACKING_ROM, expf(-div_rn(add_rn(mul_rn(te0, p.traj_weight[0]), mul_rn(te1, p.traj_weight[1])), p.tracking_sigma)));
        if (rs[H_UNIT_QUAT] != 0.f) {
            const float nrm = sqrtf(add_rn(add_rn(add_rn(mul_rn(act[0], act[0]), mul_rn(act[1], act[1])), mul_rn(act[2], act[2])), mul_rn(act[3], act[3])));
            add_term(H_UNIT_QUAT, mul_rn(sub_rn(1.0f, nrm), sub_rn(1.0f, nrm)));
        }
        if (p.only_positive) rew = fmaxf(rew, 0.0f);
        if (rs[H_TERMINATION] != 0.f) add_term(H_TERMINATION, (reset && !time_out) ? 1.0f : 0.0f);

        // in-place reset (legged_robot_trajectory.py:204-246 with HopperTrajectory._reset_dofs / _reset_root_states :298-358)
        if (reset) {
            hopper_reset_env(p, b, rng, i, q, qd, act, R);
            fat = 0.0f;
            ep = 0;
            atomicAdd(&s_acc[K + 1], 1.0);
        }
        // episode_sums += term (compute_reward) and, for an env that reset, extras["episode"] partial sums + clear (:235-239)
#pragma unroll
        for (int k = 0; k < B200GYM_HOPPER_NUM_TERMS; ++k) {
            const int row = p.sum_row[k];
            if (row >= 0) {
                float* sp = b.episode_sums + static_cast<size_t>(row) * N + i;
                float v = add_rn(*sp, tv[k]);
                if (reset) {
                    atomicAdd(&s_acc[row], static_cast<double>(v));
                    v = 0.0f;
                }
                *sp = v;
            }
        }

        // observations (hopper_trajectory.py:255-282) + clip (:128-129)
        float* o = s_obs + t * HS_OBS;
        float head[14] = {mul_rn(R[2], p.z_pos_scale), R[3], R[4], R[5], R[6], mul_rn(blv[0], p.lin_vel_scale), mul_rn(blv[1], p.lin_vel_scale),
                          mul_rn(blv[2], p.lin_vel_scale), mul_rn(bav[0], p.ang_vel_scale), mul_rn(bav[1], p.ang_vel_scale), mul_rn(bav[2], p.ang_vel_scale),
                          mul_rn(qd[1], p.dof_vel_scale), mul_rn(qd[2], p.dof_vel_scale), mul_rn(qd[3], p.dof_vel_scale)};
        if (p.add_noise) {
#pragma unroll
            for (int blk = 0; blk < 4; ++blk) {
                const uint4 w = rng.words(philox::OBS_NOISE, blk);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = 4 * blk + k;
                    if (c < 14) head[c] = add_rn(head[c], mul_rn(sub_rn(mul_rn(2.0f, philox::u01(philox::word(w, k))), 1.0f), p.noise_scale_vec[c]));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 14; ++c) o[c] = clampf(head[c], -p.clip_obs, p.clip_obs);
#pragma unroll
        for (int k = 0; k < B200GYM_TRAJ_WIDTH; k += 2) {
            o[14 + k] = clampf(mul_rn(sub_rn(traj[k], R[0]), p.traj_scale[0]), -p.clip_obs, p.clip_obs);
            o[15 + k] = clampf(mul_rn(sub_rn(traj[k + 1], R[1]), p.traj_scale[1]), -p.clip_obs, p.clip_obs);
        }
        {
            const float nrm = sqrtf(add_rn(add_rn(add_rn(mul_rn(act[0], act[0]), mul_rn(act[1], act[1])), mul_rn(act[2], act[2])), mul_rn(act[3], act[3])));
            const float sgn = div_rn(act[0], nrm) < 0.0f ? -1.0f : 1.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) o[14 + B200GYM_TRAJ_WIDTH + k] = clampf(mul_rn(div_rn(act[k], nrm), sgn), -p.clip_obs, p.clip_obs);
        }

        // epilogue (:175-178) + per-env outputs
        *reinterpret_cast<float4*>(b.last_actions + i * 4) = make_float4(act[0], act[1], act[2], act[3]);
        *reinterpret_cast<float4*>(b.last_dof_vel + i * 4) = make_float4(qd[0], qd[1], qd[2], qd[3]);
#pragma unroll
        for (int k = 0; k < 6; ++k) b.last_root_vel[i * 6 + k] = R[7 + k];
        const float fx = sub_rn(tr0x, R[0]), fy = sub_rn(tr0y, R[1]);
        b.prev_error[i * 2] = mul_rn(fx, fx), b.prev_error[i * 2 + 1] = mul_rn(fy, fy);
        b.feet_air_time[i] = fat;
        reinterpret_cast<long long*>(b.episode_length_buf)[i] = ep;
        b.rew_buf[i] = rew;
        b.reset_buf[i] = reset ? 1 : 0;
        b.time_out_buf[i] = time_out ? 1 : 0;
    }
    __syncthreads();
    {
        float* g = b.obs_buf + static_cast<size_t>(tile0) * B200GYM_HOPPER_TRAJ_NUM_OBS;
        for (int q = t; q < nvalid * B200GYM_HOPPER_TRAJ_NUM_OBS; q += HT_TILE)
            g[q] = s_obs[(q / B200GYM_HOPPER_TRAJ_NUM_OBS) * HS_OBS + q % B200GYM_HOPPER_TRAJ_NUM_OBS];
    }
    if (t < K + 2 && s_acc[t] != 0.0) atomicAdd(&b.ws_sums[t], s_acc[t]);
}


// LeggedRobotTrajectory.reset_idx(env_ids) called from OUTSIDE step() for the Hopper (legged_robot_trajectory.py:204-246 with the Hopper's
// _reset_dofs / _reset_root_states; HopperTrajectory.reset, hopper_trajectory.py:286-296): the same redraw as the in-step reset for the envs
// flagged in `mask`, buffers cleared, reset_buf set, prev_error from the (stale) trajectory buffer and the new root (:233), episode_sums folded
// into the extras statistics; time_out_buf, obs_buf, rew_buf untouched.  The caller then resets the generators (b200gym_rom_reset_from_root).
__global__ void __launch_bounds__(128) hopper_reset_idx_kernel(const __grid_constant__ B200HopperEnvParams p, const __grid_constant__ B200HopperEnvBuffers b,
                                                               const uint8_t* __restrict__ mask, unsigned long long event, long long env_off) {
    const int N = p.num_envs, K = p.num_sum_rows;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ double s_acc[B200GYM_HOPPER_NUM_TERMS + 2];
    if (threadIdx.x < K + 2) s_acc[threadIdx.x] = 0.0;
    __syncthreads();
    if (e < N && mask[e] != 0) {
        const size_t i = static_cast<size_t>(e);
        const philox::Stream rng(p.seed_lo, p.seed_hi, static_cast<unsigned long long>(env_off) + i, event);
        float q[4], qd[4], act[4], R[13];
        hopper_reset_env(p, b, rng, i, q, qd, act, R);
        *reinterpret_cast<float4*>(b.last_actions + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(b.last_dof_vel + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        b.feet_air_time[i] = 0.0f;
        reinterpret_cast<long long*>(b.episode_length_buf)[i] = 0;
        b.reset_buf[i] = 1;
        const float d0 = sub_rn(b.trajectory[i * B200GYM_TRAJ_WIDTH], R[0]), d1 = sub_rn(b.trajectory[i * B200GYM_TRAJ_WIDTH + 1], R[1]);
        b.prev_error[i * 2] = mul_rn(d0, d0), b.prev_error[i * 2 + 1] = mul_rn(d1, d1);
        for (int k = 0; k < K; ++k) {
            float* sp = b.episode_sums + static_cast<size_t>(k) * N + i;
            atomicAdd(&s_acc[k], static_cast<double>(*sp));
            *sp = 0.0f;
        }
        atomicAdd(&s_acc[K + 1], 1.0);
    }
    __syncthreads();
    if (threadIdx.x < K + 2 && s_acc[threadIdx.x] != 0.0) atomicAdd(&b.ws_sums[threadIdx.x], s_acc[threadIdx.x]);
}

__global__ void hopper_extras_finalize_kernel(const __grid_constant__ B200HopperEnvParams p, const __grid_constant__ B200HopperEnvBuffers b) {
    const int K = p.num_sum_rows, tid = threadIdx.x;
    const double cnt = b.ws_sums[K + 1];
    const double mine = tid < K + 2 ? b.ws_sums[tid] : 0.0;
    __syncthreads();
    if (tid < K && cnt > 0.0) b.extras_out[tid] = static_cast<float>(mine / cnt) / p.max_episode_length_s;
    if (tid == K + 1) b.extras_out[K + 1] = static_cast<float>(cnt);
    if (tid < K + 2) b.ws_sums[tid] = 0.0;
}

}  // namespace

extern "C" int b200gym_hopper_post_physics(const B200HopperEnvParams* p, const B200HopperEnvBuffers* b, uint64_t step, int64_t env_id_offset,
                                           void* stream) {
    B200_REQUIRE(p && b, B200GYM_EINVAL, "hopper_post_physics: null argument");
    B200_REQUIRE(p->num_envs > 0 && p->num_bodies > 0 && p->num_bodies <= 16 && p->foot_body >= 0 && p->foot_body < p->num_bodies, B200GYM_EINVAL,
                 "hopper_post_physics: bad sizes");
    B200_REQUIRE(p->num_term >= 0 && p->num_term <= 8 && p->num_pen >= 0 && p->num_pen <= 8 && p->num_sum_rows >= 0 &&
                     p->num_sum_rows <= B200GYM_HOPPER_NUM_TERMS,
                 B200GYM_EINVAL, "hopper_post_physics: bad index tables");
    for (int k = 0; k < B200GYM_HOPPER_NUM_TERMS; ++k)
        B200_REQUIRE(p->reward_scale[k] == 0.0f || (p->sum_row[k] >= 0 && p->sum_row[k] < p->num_sum_rows), B200GYM_EINVAL,
                     "hopper_post_physics: active term %d has no episode_sums row", k);
    const void* must[] = {b->root_states, b->dof_state, b->contact_forces, b->actions, b->torques, b->last_actions, b->last_dof_vel, b->last_root_vel,
                          b->base_lin_vel, b->base_ang_vel, b->projected_gravity, b->feet_air_time, b->last_contacts, b->episode_length_buf, b->reset_buf,
                          b->time_out_buf, b->rew_buf, b->obs_buf, b->trajectory, b->gen_v, b->prev_error, b->time_until_next_push, b->env_origins,
                          b->extras_out, b->ws_sums, b->push_flag};
    for (const void* q : must) B200_REQUIRE(q != nullptr, B200GYM_EINVAL, "hopper_post_physics: null buffer");
    B200_REQUIRE(p->num_sum_rows == 0 || b->episode_sums, B200GYM_EINVAL, "hopper_post_physics: episode_sums missing");
    const void* vec[] = {b->dof_state, b->actions, b->torques, b->last_actions, b->last_dof_vel};
    for (const void* q : vec) B200_REQUIRE(b200_aligned16(q), B200GYM_EALIGN, "hopper_post_physics: [N, 4] tensors must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    hopper_prologue_kernel<<<(p->num_envs + 255) / 256, 256, 0, st>>>(*p, *b, step, env_id_offset);
    B200_LAUNCH_CHECK("hopper_prologue");
    const size_t smem = static_cast<size_t>(HT_TILE) * (13 + 3 * p->num_bodies + HS_TRAJ + HS_OBS) * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(hopper_post_physics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "hopper_post_physics: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
        configured = smem;
    }
    hopper_post_physics_kernel<<<(p->num_envs + HT_TILE - 1) / HT_TILE, HT_TILE, smem, st>>>(*p, *b, step, env_id_offset);
    B200_LAUNCH_CHECK("hopper_post_physics");
    hopper_extras_finalize_kernel<<<1, 32, 0, st>>>(*p, *b);
    B200_LAUNCH_CHECK("hopper_extras_finalize");
    return B200GYM_OK;
}

extern "C" int b200gym_hopper_reset_idx(const B200HopperEnvParams* p, const B200HopperEnvBuffers* b, const uint8_t* reset_mask, uint64_t event,
                                        int64_t env_id_offset, void* stream) {
    B200_REQUIRE(p && b && reset_mask, B200GYM_EINVAL, "hopper_reset_idx: null argument");
    B200_REQUIRE(p->num_envs > 0 && p->num_sum_rows >= 0 && p->num_sum_rows <= B200GYM_HOPPER_NUM_TERMS, B200GYM_EINVAL, "hopper_reset_idx: bad sizes");
    const void* must[] = {b->root_states, b->dof_state, b->actions, b->last_actions, b->last_dof_vel, b->feet_air_time, b->episode_length_buf, b->reset_buf,
                          b->trajectory, b->prev_error, b->env_origins, b->extras_out, b->ws_sums};
    for (const void* q : must) B200_REQUIRE(q != nullptr, B200GYM_EINVAL, "hopper_reset_idx: null buffer");
    B200_REQUIRE(p->num_sum_rows == 0 || b->episode_sums, B200GYM_EINVAL, "hopper_reset_idx: episode_sums missing");
    const void* vec[] = {b->dof_state, b->actions, b->last_actions, b->last_dof_vel};
    for (const void* q : vec) B200_REQUIRE(b200_aligned16(q), B200GYM_EALIGN, "hopper_reset_idx: [N, 4] tensors must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    hopper_reset_idx_kernel<<<(p->num_envs + 127) / 128, 128, 0, st>>>(*p, *b, reset_mask, event, env_id_offset);
    B200_LAUNCH_CHECK("hopper_reset_idx");
    hopper_extras_finalize_kernel<<<1, 32, 0, st>>>(*p, *b);
    B200_LAUNCH_CHECK("hopper_reset_idx finalize");
    return B200GYM_OK;
}
