// the Hopper's torque law, Hopper._compute_torques (legged_gym/envs/hopper/hopper.py:168-237).
//
// One thread per env, one pass: the foot's contact-switched spring / PD torque, the reaction wheels' orientation control
// (quaternion error -> rotation matrix -> so3 log map -> PD in the body frame -> actuator rotation) or spin-down damping in contact,
// the torque-speed envelope and the torque limits, with the per-env randomised gain / limit multipliers.  The quaternion / so3 steps
// follow pytorch3d.transforms (quaternion_invert, quaternion_multiply, quaternion_to_matrix, so3_log_map with its linear acos
// continuation, Rotate.transform_points) operation by operation in fp32 round-to-nearest — the log map is ill-conditioned near
// 0 and pi, where the reference's own result depends on that order.
//
// Algorithmic bytes per env: actions 16 + dof_state 32 + foot contact z 4 + quaternion 16 + base_ang_vel 12 + multipliers 76 +
// two torque tensors 32 = 188 B; the contact read touches one 32 B sector of the [N, bodies, 3] tensor.
#include <math.h>

#include "common.cuh"
#include "philox.cuh"
#include "../../include/b200gym.h"

namespace {

struct AcosExtrap {   // acos_linear_extrapolation(x, (-1 + 1e-4, 1 - 1e-4)), pytorch3d/transforms/math.py, constants formed in double on the host
    float lo, hi, slope_lo, slope_hi, acos_lo, acos_hi;
};

__global__ void __launch_bounds__(256) hopper_torques_kernel(const __grid_constant__ B200HopperTorqueParams p,
                                                             const __grid_constant__ B200HopperTorqueBuffers b,
                                                             const __grid_constant__ AcosExtrap ax) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (e >= p.num_envs) return;
    const size_t i = static_cast<size_t>(e);
    const float4 act = *reinterpret_cast<const float4*>(b.actions + i * 4);
    const float4 ds0 = *reinterpret_cast<const float4*>(b.dof_state + i * 8), ds1 = *reinterpret_cast<const float4*>(b.dof_state + i * 8 + 4);
    const float foot_pos = ds0.x, foot_vel = ds0.y, wv[3] = {ds0.w, ds1.y, ds1.w};
    const bool contact = b.contact_forces[(i * p.num_bodies + p.foot_body) * 3 + 2] > 0.1f;                      // :184
    const float4 pr = *reinterpret_cast<const float4*>(b.p_gain_random + i * 4), dr = *reinterpret_cast<const float4*>(b.d_gain_random + i * 4);
    const float pg[4] = {mul_rn(p.p_gains[0], pr.x), mul_rn(p.p_gains[1], pr.y), mul_rn(p.p_gains[2], pr.z), mul_rn(p.p_gains[3], pr.w)};   // :191
    const float drv[4] = {dr.x, dr.y, dr.z, dr.w};
    const float dg[4] = {mul_rn(p.d_gains[0], dr.x), mul_rn(p.d_gains[1], dr.y), mul_rn(p.d_gains[2], dr.z), mul_rn(p.d_gains[3], dr.w)};   // :192
    float tq[4];

    // foot: zero + spring in contact, PD towards foot_pos_des in flight (:198-201)
    if (contact) {
        tq[0] = add_rn(0.0f, sub_rn(mul_rn(-b.spring_stiffness[i], foot_pos), mul_rn(b.spring_damping[i], foot_vel)));
    } else {
        tq[0] = sub_rn(mul_rn(-pg[0], sub_rn(foot_pos, b.foot_pos_des[i])), mul_rn(dg[0], foot_vel));
    }

    if (p.spindown && contact) {                                                                              // :204-205
#pragma unroll
        for (int j = 0; j < 3; ++j) tq[1 + j] = mul_rn(-mul_rn(p.kd_spindown[j], drv[1 + j]), wv[j]);          // kd_spindown * d_gain_random (:193)
    } else {                                                                                                  // :212-222
        const float a[4] = {mul_rn(act.x, p.action_scale), mul_rn(act.y, p.action_scale), mul_rn(act.z, p.action_scale), mul_rn(act.w, p.action_scale)};
        const float nrm = sqrtf(add_rn(add_rn(add_rn(mul_rn(a[0], a[0]), mul_rn(a[1], a[1])), mul_rn(a[2], a[2])), mul_rn(a[3], a[3])));
        // quaternion_invert(quat_des): (w, -x, -y, -z)
        const float aw = div_rn(a[0], nrm), ax_ = -div_rn(a[1], nrm), ay = -div_rn(a[2], nrm), az = -div_rn(a[3], nrm);
        const float* rs = b.root_states + i * 13;
        const float bw = rs[6], bx = rs[3], by = rs[4], bz = rs[5];                                            // wxyz_quat_inds = [6, 3, 4, 5] (:62)
        // quaternion_raw_multiply + standardize_quaternion
        float r = sub_rn(sub_rn(sub_rn(mul_rn(aw, bw), mul_rn(ax_, bx)), mul_rn(ay, by)), mul_rn(az, bz));
        float qi = sub_rn(add_rn(add_rn(mul_rn(aw, bx), mul_rn(ax_, bw)), mul_rn(ay, bz)), mul_rn(az, by));
        float qj = add_rn(add_rn(sub_rn(mul_rn(aw, by), mul_rn(ax_, bz)), mul_rn(ay, bw)), mul_rn(az, bx));
        float qk = add_rn(sub_rn(add_rn(mul_rn(aw, bz), mul_rn(ax_, by)), mul_rn(ay, bx)), mul_rn(az, bw));
        if (r < 0.0f) r = -r, qi = -qi, qj = -qj, qk = -qk;
        // quaternion_to_matrix
        const float two_s = div_rn(2.0f, add_rn(add_rn(add_rn(mul_rn(r, r), mul_rn(qi, qi)), mul_rn(qj, qj)), mul_rn(qk, qk)));
        const float R00 = sub_rn(1.0f, mul_rn(two_s, add_rn(mul_rn(qj, qj), mul_rn(qk, qk))));
        const float R01 = mul_rn(two_s, sub_rn(mul_rn(qi, qj), mul_rn(qk, r)));
        const float R02 = mul_rn(two_s, add_rn(mul_rn(qi, qk), mul_rn(qj, r)));
        const float R10 = mul_rn(two_s, add_rn(mul_rn(qi, qj), mul_rn(qk, r)));
        const float R11 = sub_rn(1.0f, mul_rn(two_s, add_rn(mul_rn(qi, qi), mul_rn(qk, qk))));
        const float R12 = mul_rn(two_s, sub_rn(mul_rn(qj, qk), mul_rn(qi, r)));
        const float R20 = mul_rn(two_s, sub_rn(mul_rn(qi, qk), mul_rn(qj, r)));
        const float R21 = mul_rn(two_s, add_rn(mul_rn(qj, qk), mul_rn(qi, r)));
        const float R22 = sub_rn(1.0f, mul_rn(two_s, add_rn(mul_rn(qi, qi), mul_rn(qj, qj))));
        // so3_rotation_angle + so3_log_map + hat_inv
        const float phi_cos = mul_rn(sub_rn(add_rn(add_rn(R00, R11), R22), 1.0f), 0.5f);
        float phi;
        if (phi_cos >= ax.hi) phi = add_rn(mul_rn(sub_rn(phi_cos, ax.hi), ax.slope_hi), ax.acos_hi);
        else if (phi_cos <= ax.lo) phi = add_rn(mul_rn(sub_rn(phi_cos, ax.lo), ax.slope_lo), ax.acos_lo);
        else phi = acosf(phi_cos);
        const float phi_sin = sinf(phi);
        const float factor = fabsf(phi_sin) > 5e-5f ? div_rn(phi, mul_rn(2.0f, phi_sin)) : add_rn(0.5f, mul_rn(mul_rn(phi, phi), 1.0f / 12));
        const float lg[3] = {mul_rn(factor, sub_rn(R21, R12)), mul_rn(factor, sub_rn(R02, R20)), mul_rn(factor, sub_rn(R10, R01))};
        float bav[3];
        if (p.ang_vel_from_root) {   // the refresh of the previous sub-step (hopper_trajectory.py:124-126): quat_rotate_inverse(base_quat, root[:, 10:13])
            const float vx = rs[10], vy = rs[11], vz = rs[12];
            const float k = sub_rn(mul_rn(mul_rn(2.0f, bw), bw), 1.0f), w2 = mul_rn(bw, 2.0f);
            const float d = add_rn(add_rn(mul_rn(bx, vx), mul_rn(by, vy)), mul_rn(bz, vz));
            bav[0] = add_rn(sub_rn(mul_rn(vx, k), mul_rn(sub_rn(mul_rn(by, vz), mul_rn(bz, vy)), w2)), mul_rn(mul_rn(bx, d), 2.0f));
            bav[1] = add_rn(sub_rn(mul_rn(vy, k), mul_rn(sub_rn(mul_rn(bz, vx), mul_rn(bx, vz)), w2)), mul_rn(mul_rn(by, d), 2.0f));
            bav[2] = add_rn(sub_rn(mul_rn(vz, k), mul_rn(sub_rn(mul_rn(bx, vy), mul_rn(by, vx)), w2)), mul_rn(mul_rn(bz, d), 2.0f));
        } else {
            bav[0] = b.base_ang_vel[i * 3], bav[1] = b.base_ang_vel[i * 3 + 1], bav[2] = b.base_ang_vel[i * 3 + 2];
        }
        float lt[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) lt[j] = sub_rn(mul_rn(-pg[1 + j], lg[j]), mul_rn(dg[1 + j], bav[j]));   // :219
        // Rotate(rot_actuator).transform_points: row vector times matrix
#pragma unroll
        for (int j = 0; j < 3; ++j) tq[1 + j] = fmaf(lt[2], p.rot_actuator[6 + j], fmaf(lt[1], p.rot_actuator[3 + j], mul_rn(lt[0], p.rot_actuator[j])));
    }

    // torque-speed envelope (:231-236), then the torque limits (:237)
    const float ts = mul_rn(p.torque_speed_bound_ratio, b.torque_speed_bound_ratio_random[i]);
    const float4 tl = *reinterpret_cast<const float4*>(b.torque_limit_random + i * 4);
    const float tb[4] = {mul_rn(p.torque_limits[0], tl.x), mul_rn(p.torque_limits[1], tl.y), mul_rn(p.torque_limits[2], tl.z), mul_rn(p.torque_limits[3], tl.w)};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float wb = mul_rn(p.wheel_speed_limits[j], b.wheel_limit_random[i * 3 + j]);
        const float k = div_rn(mul_rn(-ts, tb[1 + j]), wb);
        const float upper = mul_rn(k, sub_rn(wv[j], wb)), lower = mul_rn(k, add_rn(wv[j], wb));
        tq[1 + j] = fminf(fmaxf(tq[1 + j], lower), upper);
    }
    *reinterpret_cast<float4*>(b.torques + i * 4) = make_float4(tq[0], tq[1], tq[2], tq[3]);
    *reinterpret_cast<float4*>(b.torques_clipped + i * 4) = make_float4(fminf(fmaxf(tq[0], -tb[0]), tb[0]), fminf(fmaxf(tq[1], -tb[1]), tb[1]),
                                                                        fminf(fmaxf(tq[2], -tb[2]), tb[2]), fminf(fmaxf(tq[3], -tb[3]), tb[3]));
}

// Hopper.compute_observations (hopper.py:239-258) + the clip of Hopper.step (:116-117).  One thread per env computes the 21 columns; the
// [128, 21] tile of a CTA is one contiguous 10.75 KB block of the output, staged in shared memory (row stride 21: conflict-free) and
// written out with coalesced 128-bit stores.  96 B read + 84 B written per env.
constexpr int HOBS = B200GYM_HOPPER_NUM_OBS, HT = 128;

__global__ void __launch_bounds__(HT) hopper_obs_kernel(const __grid_constant__ B200HopperObsParams p, const float* __restrict__ root,
                                                        const float* __restrict__ lin_vel, const float* __restrict__ ang_vel,
                                                        const float* __restrict__ dof_state, const float* __restrict__ commands,
                                                        const float* __restrict__ actions, float* __restrict__ obs, unsigned long long event,
                                                        long long env_off) {
    __shared__ __align__(16) float tile[HT * HOBS];
    const int env0 = blockIdx.x * HT, e = threadIdx.x;
    const int nenv = min(HT, p.num_envs - env0);
    pdl_launch_dependents();
    pdl_wait();
    if (e < nenv) {
        const size_t i = static_cast<size_t>(env0 + e);
        float o[HOBS];
        const float* r = root + i * 13;
        o[0] = mul_rn(r[2], p.z_pos_scale);
        o[1] = r[3], o[2] = r[4], o[3] = r[5], o[4] = r[6];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            o[5 + c] = mul_rn(lin_vel[i * 3 + c], p.lin_vel_scale);
            o[8 + c] = mul_rn(ang_vel[i * 3 + c], p.ang_vel_scale);
            o[11 + c] = mul_rn(dof_state[i * 8 + 2 * (1 + c) + 1], p.dof_vel_scale);
            o[14 + c] = mul_rn(commands[i * 4 + c], p.commands_scale[c]);
        }
        const float4 a = *reinterpret_cast<const float4*>(actions + i * 4);   // normalised action quaternion, qw >= 0 (:242-244)
        const float nrm = sqrtf(add_rn(add_rn(add_rn(mul_rn(a.x, a.x), mul_rn(a.y, a.y)), mul_rn(a.z, a.z)), mul_rn(a.w, a.w)));
        const float sgn = div_rn(a.x, nrm) < 0.0f ? -1.0f : 1.0f;
        o[17] = mul_rn(div_rn(a.x, nrm), sgn), o[18] = mul_rn(div_rn(a.y, nrm), sgn), o[19] = mul_rn(div_rn(a.z, nrm), sgn), o[20] = mul_rn(div_rn(a.w, nrm), sgn);
        if (p.add_noise) {   // obs += (2 * rand_like(obs) - 1) * noise_scale_vec (:257-258); columns 14-20 carry zero scales: no draw needed
            const philox::Stream rng(p.seed_lo, p.seed_hi, static_cast<uint64_t>(env_off) + i, event);
#pragma unroll
            for (int blk = 0; blk < 4; ++blk) {
                const uint4 w = rng.words(philox::OBS_NOISE, blk);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = 4 * blk + k;
                    o[c] = add_rn(o[c], mul_rn(sub_rn(mul_rn(2.0f, philox::u01(philox::word(w, k))), 1.0f), p.noise_scale_vec[c]));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < HOBS; ++c) tile[e * HOBS + c] = fminf(fmaxf(o[c], -p.clip_observations), p.clip_observations);
    }
    __syncthreads();
    float* dst = obs + static_cast<size_t>(env0) * HOBS;   // 128 * 21 * 4 B per full tile: 16-byte aligned whenever obs is
    const int total = nenv * HOBS;
    if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        for (int q = e; q < total / 4; q += HT) reinterpret_cast<float4*>(dst)[q] = reinterpret_cast<const float4*>(tile)[q];
    } else {
        for (int q = e; q < total; q += HT) dst[q] = tile[q];
    }
}

__global__ void __launch_bounds__(256) hopper_reward_terms_kernel(int n, float dt, const float* __restrict__ torques, const float* __restrict__ dof_state,
                                                                  const float* __restrict__ last_dof_vel, const float* __restrict__ actions,
                                                                  float* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (e >= n) return;
    const size_t i = static_cast<size_t>(e);
    const float4 tq = *reinterpret_cast<const float4*>(torques + i * 4), lv = *reinterpret_cast<const float4*>(last_dof_vel + i * 4);
    const float4 a = *reinterpret_cast<const float4*>(actions + i * 4);
    const float4 d0 = *reinterpret_cast<const float4*>(dof_state + i * 8), d1 = *reinterpret_cast<const float4*>(dof_state + i * 8 + 4);
    const float tl = add_rn(add_rn(fabsf(tq.y), fabsf(tq.z)), fabsf(tq.w));                                   // :448-450
    const float q1 = div_rn(sub_rn(lv.y, d0.w), dt), q2 = div_rn(sub_rn(lv.z, d1.y), dt), q3 = div_rn(sub_rn(lv.w, d1.w), dt);
    const float acc = add_rn(add_rn(mul_rn(q1, q1), mul_rn(q2, q2)), mul_rn(q3, q3));                         // :452-454
    const float nrm = sqrtf(add_rn(add_rn(add_rn(mul_rn(a.x, a.x), mul_rn(a.y, a.y)), mul_rn(a.z, a.z)), mul_rn(a.w, a.w)));
    const float uq = sub_rn(1.0f, nrm);                                                                       // :456-458
    out[i * 3 + 0] = tl, out[i * 3 + 1] = acc, out[i * 3 + 2] = mul_rn(uq, uq);
}

}  // namespace

extern "C" int b200gym_hopper_observations(const B200HopperObsParams* p, const float* root_states, const float* base_lin_vel, const float* base_ang_vel,
                                           const float* dof_state, const float* commands, const float* actions, float* obs, uint64_t event,
                                           int64_t env_id_offset, void* stream) {
    B200_REQUIRE(p && root_states && base_lin_vel && base_ang_vel && dof_state && commands && actions && obs, B200GYM_EINVAL, "hopper_observations: null argument");
    B200_REQUIRE(p->num_envs > 0, B200GYM_EINVAL, "hopper_observations: num_envs must be positive (got %d)", p->num_envs);
    B200_REQUIRE(b200_aligned16(actions), B200GYM_EALIGN, "hopper_observations: actions must be 16-byte aligned");
    b200_launch_pdl(p->num_envs, hopper_obs_kernel, dim3((p->num_envs + HT - 1) / HT), dim3(HT), 0, static_cast<cudaStream_t>(stream), *p, root_states,
                    base_lin_vel, base_ang_vel, dof_state, commands, actions, obs, static_cast<unsigned long long>(event), static_cast<long long>(env_id_offset));
    B200_LAUNCH_CHECK("hopper_observations");
    return B200GYM_OK;
}

extern "C" int b200gym_hopper_reward_terms(int32_t num_envs, float dt, const float* torques, const float* dof_state, const float* last_dof_vel,
                                           const float* actions, float* out, void* stream) {
    B200_REQUIRE(torques && dof_state && last_dof_vel && actions && out, B200GYM_EINVAL, "hopper_reward_terms: null argument");
    B200_REQUIRE(num_envs > 0 && dt > 0.0f, B200GYM_EINVAL, "hopper_reward_terms: num_envs and dt must be positive");
    const void* vec[] = {torques, dof_state, last_dof_vel, actions};
    for (const void* q : vec) B200_REQUIRE(b200_aligned16(q), B200GYM_EALIGN, "hopper_reward_terms: [N, 4] tensors must be 16-byte aligned");
    b200_launch_pdl(num_envs, hopper_reward_terms_kernel, dim3((num_envs + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<int>(num_envs), dt,
                    torques, dof_state, last_dof_vel, actions, out);
    B200_LAUNCH_CHECK("hopper_reward_terms");
    return B200GYM_OK;
}

extern "C" int b200gym_hopper_torques(const B200HopperTorqueParams* p, const B200HopperTorqueBuffers* b, void* stream) {
    B200_REQUIRE(p && b, B200GYM_EINVAL, "hopper_torques: null argument");
    B200_REQUIRE(p->num_envs > 0, B200GYM_EINVAL, "hopper_torques: num_envs must be positive (got %d)", p->num_envs);
    B200_REQUIRE(p->num_bodies > 0 && p->foot_body >= 0 && p->foot_body < p->num_bodies, B200GYM_EINVAL, "hopper_torques: foot_body %d outside [0, %d)",
                 p->foot_body, p->num_bodies);
    const void* must[] = {b->actions, b->dof_state, b->contact_forces, b->root_states, b->base_ang_vel, b->p_gain_random, b->d_gain_random,
                          b->torque_limit_random, b->wheel_limit_random, b->spring_stiffness, b->spring_damping, b->foot_pos_des,
                          b->torque_speed_bound_ratio_random, b->torques, b->torques_clipped};
    for (const void* q : must) B200_REQUIRE(q != nullptr, B200GYM_EINVAL, "hopper_torques: null buffer");
    const void* vec[] = {b->actions, b->dof_state, b->p_gain_random, b->d_gain_random, b->torque_limit_random, b->torques, b->torques_clipped};
    for (const void* q : vec) B200_REQUIRE(b200_aligned16(q), B200GYM_EALIGN, "hopper_torques: [N, 4] tensors must be 16-byte aligned");
    const double lo = -1.0 + 1e-4, hi = 1.0 - 1e-4;   // so3_log_map's default cos_bound
    AcosExtrap ax;
    ax.lo = static_cast<float>(lo), ax.hi = static_cast<float>(hi);
    ax.slope_lo = static_cast<float>(-1.0 / sqrt(1.0 - lo * lo)), ax.slope_hi = static_cast<float>(-1.0 / sqrt(1.0 - hi * hi));
    ax.acos_lo = static_cast<float>(acos(lo)), ax.acos_hi = static_cast<float>(acos(hi));
    b200_launch_pdl(p->num_envs, hopper_torques_kernel, dim3((p->num_envs + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream), *p, *b, ax);
    B200_LAUNCH_CHECK("hopper_torques");
    return B200GYM_OK;
}
