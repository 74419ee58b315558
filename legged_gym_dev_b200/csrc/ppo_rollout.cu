// Rollout-side kernels of rsl_rl's PPO (SURVEY.md §8a G1/G3/G4 callers: PPO.act / PPO.process_env_step /
// RolloutStorage.add_transitions of rsl_rl v1.0.2, as legged_gym/utils/task_registry.py:148 drives them):
//
//   ppo_act_store_kernel   everything of PPO.act after the two MLP forwards, in ONE launch: Normal(mu, std).sample(),
//                          its log-prob, and the transition written straight into row `step` of the rollout storage
//                          (observations, critic observations, actions, values, log-prob, mu, sigma) — replaces
//                          torch.distributions.Normal + sample + log_prob + 7 copy_ launches (~20 eager launches per env step)
//   ppo_store_step_kernel  PPO.process_env_step's part of add_transitions: rewards, dones and the time-out flags of the env
//                          step into the same storage row (the time-out bootstrap itself is fused into gae_returns_kernel)
//
// The sample is drawn from Philox4x32-10 keyed like every other draw of the library (philox.cuh): counter = (global env id,
// act event, site POLICY_SAMPLE), Box-Muller on 24-bit uniforms.  rsl_rl draws from torch's global generator; as with the env
// kernels the random numbers are inputs of the parity contract (oracle/port_ppo.py sample_actions is the specification).
#include "common.cuh"
#include "philox.cuh"
#include "../../include/b200gym.h"

namespace {

constexpr int MAXA = 16;
__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__global__ void __launch_bounds__(256) ppo_act_store_kernel(int n_envs, int num_actions, int num_obs, int num_critic_obs, const float* __restrict__ mu_out,
                                                            int ld_mu, const float* __restrict__ value_out, int ld_value,
                                                            const float* __restrict__ stdv, const float* __restrict__ obs, long long ld_obs,
                                                            const float* __restrict__ critic_obs, long long ld_cobs, uint32_t seed_lo,
                                                            uint32_t seed_hi, unsigned long long event, const unsigned long long* __restrict__ event_dev,
                                                            unsigned long long env_id_offset, float* __restrict__ st_obs, float* __restrict__ st_cobs, float* __restrict__ st_actions,
                                                            float* __restrict__ st_values, float* __restrict__ st_logp, float* __restrict__ st_mu,
                                                            float* __restrict__ st_sigma) {
    const int A = num_actions;
    pdl_launch_dependents();   // programmatic dependent launch (common.cuh): scheduled while the policy forward drains
    pdl_wait();
    if (event_dev) event = *event_dev;   // graph-replayable mode: the act counter lives in device memory (advanced by ppo_store_step_kernel)
    const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
    // four lanes per env: lane blk draws the Philox block of actions 4 blk .. 4 blk + 3 (one thread per env left 16 CTAs doing ~3000
    // serial instructions each at 4096 envs); the log-probability is summed over the four lanes
    {
        const long long e = tid >> 2;
        const int blk = static_cast<int>(tid & 3);
        const bool on = e < n_envs;
        float logp = 0.0f;
        if (on && blk * 4 < A) {
            const philox::Stream rng(seed_lo, seed_hi, env_id_offset + static_cast<unsigned long long>(e), event);
            const float LOG_SQRT_2PI = 0.91893853320467274178f, TWO_PI = 6.283185307179586f;
            const uint4 w = rng.words(philox::POLICY_SAMPLE, blk);
            // two Box-Muller pairs per block: (w.x, w.y) -> columns 4 blk + {0, 1}, (w.z, w.w) -> columns 4 blk + {2, 3}
            const float u0 = (static_cast<float>(w.x >> 8) + 0.5f) * 5.9604644775390625e-08f, u1 = philox::u01(w.y);
            const float u2 = (static_cast<float>(w.z >> 8) + 0.5f) * 5.9604644775390625e-08f, u3 = philox::u01(w.w);
            const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
            float s0, c0, s1, c1;
            sincosf(TWO_PI * u1, &s0, &c0);
            sincosf(TWO_PI * u3, &s1, &c1);
            const float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int a = blk * 4 + j;
                if (a < A) {
                    const float sg = stdv[a], m = mu_out[e * ld_mu + a];
                    const float x = __fadd_rn(m, __fmul_rn(sg, z[j]));          // torch.normal: mean + std * eps
                    const float d = x - m;
                    logp += -(d * d) / (2.0f * sg * sg) - logf(sg) - LOG_SQRT_2PI;   // Normal.log_prob, summed over the action dim
                    st_actions[e * A + a] = x;
                    st_mu[e * A + a] = m;
                    st_sigma[e * A + a] = sg;
                }
            }
        }
        logp += __shfl_xor_sync(0xffffffffu, logp, 1);
        logp += __shfl_xor_sync(0xffffffffu, logp, 2);
        if (on && blk == 0) {
            st_logp[e] = logp;
            st_values[e] = value_out[e * ld_value];
        }
    }
    // the observations the action was computed from (the env rewrites its obs_buf in place on the next step)
    const long long n_o = static_cast<long long>(n_envs) * num_obs;
    if (ld_obs == num_obs && (num_obs & 3) == 0 && al16(obs) && al16(st_obs)) {
        for (long long i = tid; i < (n_o >> 2); i += nthreads) reinterpret_cast<float4*>(st_obs)[i] = __ldg(reinterpret_cast<const float4*>(obs) + i);
    } else {
        for (long long i = tid; i < n_o; i += nthreads) st_obs[i] = obs[(i / num_obs) * ld_obs + i % num_obs];
    }
    if (st_cobs != nullptr) {
        const long long n_c = static_cast<long long>(n_envs) * num_critic_obs;
        if (ld_cobs == num_critic_obs && (num_critic_obs & 3) == 0 && al16(critic_obs) && al16(st_cobs)) {
            for (long long i = tid; i < (n_c >> 2); i += nthreads)
                reinterpret_cast<float4*>(st_cobs)[i] = __ldg(reinterpret_cast<const float4*>(critic_obs) + i);
        } else {
            for (long long i = tid; i < n_c; i += nthreads) st_cobs[i] = critic_obs[(i / num_critic_obs) * ld_cobs + i % num_critic_obs];
        }
    }
}

__global__ void __launch_bounds__(256) ppo_store_step_kernel(int n_envs, const float* __restrict__ rewards, const uint8_t* __restrict__ dones,
                                                             const uint8_t* __restrict__ time_outs, float* __restrict__ st_rewards,
                                                             uint8_t* __restrict__ st_dones, uint8_t* __restrict__ st_time_outs,
                                                             unsigned long long* __restrict__ event_dev) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (e == 0 && event_dev) *event_dev += 1;   // the next PPO.act draws from the next event
    if (e >= n_envs) return;
    st_rewards[e] = rewards[e];
    st_dones[e] = dones[e] ? 1 : 0;
    st_time_outs[e] = time_outs ? (time_outs[e] ? 1 : 0) : 0;
}

}  // namespace

extern "C" {

int b200gym_ppo_act_store(int32_t n_envs, int32_t num_actions, int32_t num_obs, int32_t num_critic_obs, const float* mu_out, int32_t ld_mu,
                          const float* value_out, int32_t ld_value, const float* std, const float* obs, int64_t ld_obs, const float* critic_obs,
                          int64_t ld_critic_obs, uint64_t seed, uint64_t event, const uint64_t* event_dev, uint64_t env_id_offset, float* st_obs,
                          float* st_critic_obs, float* st_actions, float* st_values, float* st_log_prob, float* st_mu, float* st_sigma, void* stream) {
    B200_REQUIRE(mu_out && value_out && std && obs && st_obs && st_actions && st_values && st_log_prob && st_mu && st_sigma, B200GYM_EINVAL,
                 "ppo_act_store: null argument");
    B200_REQUIRE(n_envs > 0 && num_actions > 0 && num_actions <= MAXA && num_obs > 0 && ld_mu >= num_actions && ld_value >= 1 && ld_obs >= num_obs,
                 B200GYM_EINVAL, "ppo_act_store: n_envs > 0, 1..%d actions, leading dimensions >= widths", MAXA);
    B200_REQUIRE((st_critic_obs == nullptr) || (critic_obs != nullptr && num_critic_obs > 0 && ld_critic_obs >= num_critic_obs), B200GYM_EINVAL,
                 "ppo_act_store: critic observations need a source, a width and a leading dimension");
    const long long work = static_cast<long long>(n_envs) * ((num_obs + (st_critic_obs ? num_critic_obs : 0) + 3) / 4);
    long long blocks = ((work > n_envs ? work : n_envs) + 255) / 256;
    const long long min_blocks = (4LL * n_envs + 255) / 256;   // four lanes per env
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < min_blocks) blocks = min_blocks;
    b200_launch_pdl(0, ppo_act_store_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream),
        n_envs, num_actions, num_obs, num_critic_obs, mu_out, ld_mu, value_out, ld_value, std, obs, ld_obs, critic_obs, ld_critic_obs,
        static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), event, reinterpret_cast<const unsigned long long*>(event_dev), env_id_offset,
        st_obs, st_critic_obs, st_actions, st_values,
        st_log_prob, st_mu, st_sigma);
    B200_LAUNCH_CHECK("ppo_act_store");
    return B200GYM_OK;
}

int b200gym_ppo_store_step(int32_t n_envs, const float* rewards, const uint8_t* dones, const uint8_t* time_outs, float* st_rewards,
                           uint8_t* st_dones, uint8_t* st_time_outs, uint64_t* event_dev, void* stream) {
    B200_REQUIRE(n_envs > 0 && rewards && dones && st_rewards && st_dones && st_time_outs, B200GYM_EINVAL, "ppo_store_step: bad argument");
    b200_launch_pdl(0, ppo_store_step_kernel, dim3((n_envs + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream), n_envs, rewards, dones,
                    time_outs, st_rewards, st_dones, st_time_outs, reinterpret_cast<unsigned long long*>(event_dev));
    B200_LAUNCH_CHECK("ppo_store_step");
    return B200GYM_OK;
}

}  // extern "C"
