// Group G kernels: rsl_rl rollout storage + PPO update arithmetic (rsl_rl v1.0.2 restated —
// its source is not under /root/reference, see oracle/port_ppo.py for the CPU restatement these are tested against).
//
//   gae_returns_kernel    PPO.process_env_step bootstrap + RolloutStorage.compute_returns   (reverse scan, thread per env)
//   adv_normalize_kernel  advantage normalisation from (sum, sum^2, n) — all-reducible across env shards
//   gather_rows_kernel    RolloutStorage.mini_batch_generator gather of one minibatch, all tensors in one launch
//   ppo_loss_kernel       PPO.update loss forward + gradient w.r.t. network outputs (one warp-shuffle reduction per CTA)
//   grad_sumsq / clip_adam  clip_grad_norm_ + Adam.step over one flat parameter buffer
//   adaptive_lr_kernel    the KL-adaptive learning-rate schedule, kept on the device (no host sync per minibatch)
// All are HBM-streaming kernels; [T,N] storage tensors are read with consecutive threads on consecutive envs.
#include "common.cuh"
#include "../../include/b200gym.h"

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-level reduction of up to NV doubles per thread, one atomicAdd per value per CTA
template <int NV>
__device__ __forceinline__ void block_accumulate(double (&v)[NV], double* out) {
    __shared__ double sh[NV][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        v[k] = warp_sum(v[k]);
        if (lane == 0) sh[k][warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = lane < nw ? sh[k][lane] : 0.0;
            x = warp_sum(x);
            if (lane == 0 && x != 0.0) atomicAdd(out + k, x);
        }
    }
}

__global__ void __launch_bounds__(256) gae_returns_kernel(float* __restrict__ rewards, const float* __restrict__ values,
                                                          const uint8_t* __restrict__ dones, const uint8_t* __restrict__ time_outs,
                                                          const float* __restrict__ last_values, float* __restrict__ returns,
                                                          float* __restrict__ advantages, double* __restrict__ stats, int T, int N,
                                                          float gamma, float lam) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[3] = {0.0, 0.0, 0.0};
    if (n < N) {
        float adv = 0.0f, next_v = last_values[n];
        const float gl = mul_rn(gamma, lam);
        // The recurrence is sequential in t, the memory traffic is not: the inputs of CH time steps are fetched together
        // (CH independent loads in flight per thread) before the CH dependent updates run — at 4096 envs the kernel is a
        // chain of DRAM round trips, not a bandwidth problem (profiles/r1_kernels_ncu.md).
        constexpr int CH = 8;
        for (int t1 = T; t1 > 0; t1 -= CH) {
            const int cnt = t1 < CH ? t1 : CH;
            float v[CH], r[CH];
            unsigned char dn[CH], to[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                if (k < cnt) {
                    const size_t i = static_cast<size_t>(t1 - 1 - k) * N + n;
                    v[k] = values[i];
                    r[k] = rewards[i];
                    dn[k] = dones[i];
                    to[k] = time_outs ? time_outs[i] : static_cast<unsigned char>(0);
                }
            }
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                if (k < cnt) {
                    const size_t i = static_cast<size_t>(t1 - 1 - k) * N + n;
                    float rr = r[k];
                    if (to[k]) {   // PPO.process_env_step: rewards += gamma * values * time_outs
                        rr = add_rn(rr, mul_rn(gamma, v[k]));
                        rewards[i] = rr;
                    }
                    const float nt = dn[k] ? 0.0f : 1.0f;
                    const float delta = sub_rn(add_rn(rr, mul_rn(mul_rn(nt, gamma), next_v)), v[k]);
                    adv = add_rn(delta, mul_rn(mul_rn(nt, gl), adv));
                    const float ret = add_rn(adv, v[k]);
                    returns[i] = ret;
                    const float a = sub_rn(ret, v[k]);
                    advantages[i] = a;
                    acc[0] += a;
                    acc[1] += static_cast<double>(a) * a;
                    next_v = v[k];
                }
            }
        }
        acc[2] = T;
    }
    block_accumulate<3>(acc, stats);
}

__global__ void adv_normalize_kernel(float* __restrict__ adv, const double* __restrict__ stats, long long count) {
    const double n = stats[2], mean = stats[0] / n;
    const double var = fmax((stats[1] - n * mean * mean) / (n - 1.0), 0.0);   // torch.std: unbiased
    const float fm = static_cast<float>(mean), inv = 1.0f / (static_cast<float>(sqrt(var)) + 1e-8f);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        adv[i] = (adv[i] - fm) * inv;
}

constexpr int MAX_GATHER = 12;
struct GatherArgs {
    void* dst[MAX_GATHER];
    const void* src[MAX_GATHER];
    int row_bytes[MAX_GATHER];
    int n;
};

// one warp per output row; every tensor's row is copied with the widest aligned vector the row size allows
__global__ void __launch_bounds__(256) gather_rows_kernel(const __grid_constant__ GatherArgs a, const long long* __restrict__ idx,
                                                          long long n_rows) {
    const int lane = threadIdx.x & 31;
    const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const long long srow = idx[row];
    for (int k = 0; k < a.n; ++k) {
        const int rb = a.row_bytes[k];
        const char* s = static_cast<const char*>(a.src[k]) + srow * rb;
        char* d = static_cast<char*>(a.dst[k]) + row * rb;
        if ((rb & 15) == 0) {
            for (int o = lane * 16; o < rb; o += 512) *reinterpret_cast<float4*>(d + o) = *reinterpret_cast<const float4*>(s + o);
        } else if ((rb & 3) == 0) {
            for (int o = lane * 4; o < rb; o += 128) *reinterpret_cast<float*>(d + o) = *reinterpret_cast<const float*>(s + o);
        } else {
            for (int o = lane; o < rb; o += 32) d[o] = s[o];
        }
    }
}

constexpr int MAXA = 16;

__global__ void __launch_bounds__(256) ppo_loss_kernel(const __grid_constant__ B200PpoLossParams p, const float* __restrict__ mu,
                                                       const float* __restrict__ stdv, const float* __restrict__ value,
                                                       const float* __restrict__ actions, const float* __restrict__ old_logp,
                                                       const float* __restrict__ adv, const float* __restrict__ ret,
                                                       const float* __restrict__ old_v, const float* __restrict__ old_mu,
                                                       const float* __restrict__ old_sigma, float* __restrict__ d_mu,
                                                       float* __restrict__ d_value, float* __restrict__ d_std,
                                                       double* __restrict__ scalars) {
    const int A = p.num_actions;
    __shared__ float s_dstd[MAXA];
    if (threadIdx.x < MAXA) s_dstd[threadIdx.x] = 0.0f;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    float dstd[MAXA];
#pragma unroll
    for (int a = 0; a < MAXA; ++a) dstd[a] = 0.0f;
    if (i < p.batch) {
        const float LOG_SQRT_2PI = 0.91893853320467274178f;
        float logp = 0.0f, ent = 0.0f, kl = 0.0f;
        float diff[MAXA];
#pragma unroll
        for (int a = 0; a < MAXA; ++a) {
            if (a < A) {
                const float sg = stdv[a], m = mu[static_cast<size_t>(i) * A + a], x = actions[static_cast<size_t>(i) * A + a];
                const float ls = logf(sg);
                diff[a] = x - m;
                logp += -(diff[a] * diff[a]) / (2.0f * sg * sg) - ls - LOG_SQRT_2PI;
                ent += 0.5f + LOG_SQRT_2PI + ls;
                const float os = old_sigma[static_cast<size_t>(i) * A + a], dm = old_mu[static_cast<size_t>(i) * A + a] - m;
                kl += logf(sg / os + 1.e-5f) + (os * os + dm * dm) / (2.0f * sg * sg) - 0.5f;
            }
        }
        const float Ai = adv[i], ratio = expf(logp - old_logp[i]);
        const float lo = 1.0f - p.clip_param, hi = 1.0f + p.clip_param;
        const float s1 = -Ai * ratio, s2 = -Ai * fminf(fmaxf(ratio, lo), hi);
        const bool inside = ratio >= lo && ratio <= hi;
        const float g_ratio = s1 > s2 ? -Ai : (s1 == s2 ? (inside ? -Ai : -0.5f * Ai) : 0.0f);   // torch.max splits ties
        const float dlogp = g_ratio * ratio * p.inv_global_batch;
        const float v = value[i], R = ret[i];
        float vloss, dv;
        if (p.use_clipped_value_loss) {
            const float ov = old_v[i], dvo = v - ov;
            const float vc = ov + fminf(fmaxf(dvo, -p.clip_param), p.clip_param);
            const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
            const bool pass = dvo >= -p.clip_param && dvo <= p.clip_param;
            vloss = fmaxf(l1, l2);
            const float g1 = 2.0f * (v - R), g2 = pass ? 2.0f * (vc - R) : 0.0f;
            dv = l1 > l2 ? g1 : (l1 == l2 ? 0.5f * (g1 + g2) : g2);
        } else {
            vloss = (R - v) * (R - v);
            dv = 2.0f * (v - R);
        }
        d_value[i] = p.value_loss_coef * dv * p.inv_global_batch;
#pragma unroll
        for (int a = 0; a < MAXA; ++a) {
            if (a < A) {
                const float sg = stdv[a], inv2 = 1.0f / (sg * sg);
                d_mu[static_cast<size_t>(i) * A + a] = dlogp * diff[a] * inv2;
                dstd[a] = dlogp * (diff[a] * diff[a] * inv2 / sg - 1.0f / sg) - p.entropy_coef * p.inv_global_batch / sg;
            }
        }
        acc[0] = kl, acc[1] = fmaxf(s1, s2), acc[2] = vloss, acc[3] = ent;
    }
    // d_std: warp shuffle, then shared-memory atomics, then one global atomic per action per CTA
#pragma unroll
    for (int a = 0; a < MAXA; ++a) {
        if (a < A) {
            const float r = warp_sum(dstd[a]);
            if ((threadIdx.x & 31) == 0) atomicAdd(&s_dstd[a], r);
        }
    }
    block_accumulate<4>(acc, scalars);   // contains a __syncthreads()
    if (threadIdx.x < A) atomicAdd(d_std + threadIdx.x, s_dstd[threadIdx.x]);
}

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, long long n, float scale, double* __restrict__ out) {
    double acc[1] = {0.0};
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const double x = static_cast<double>(g[i]) * scale;
        acc[0] += x * x;
    }
    block_accumulate<1>(acc, out);
}

// Data-parallel PPO: the ranks' flat gradient buffers live in peer-mapped (symmetric) memory; every rank sums them
// ITSELF with P2P loads over NVLink, in rank order, so all ranks obtain bit-identical sums without a ring; the squared norm that
// clip_grad_norm_ needs is accumulated in the same pass.  One launch replaces NCCL all-reduce + grad_sumsq (the buffers are
// 135 KB - 2.3 MB: a latency problem, not a bandwidth one) and — unlike a process-group collective — it is a plain kernel that a
// CUDA graph captures.  The caller brackets it with a cross-rank barrier (gradients complete / nobody overwrites while peers read).
__global__ void __launch_bounds__(256) grad_reduce_peers_kernel(const B200PeerPtrs peers, int world, float* __restrict__ out, long long n_total,
                                                                long long n_params, double* __restrict__ sumsq) {
    double acc[1] = {0.0};
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float g = 0.0f;
        for (int r = 0; r < world; ++r) g = add_rn(g, __ldcv(peers.ptr[r] + i));   // volatile-class load: peers wrote it
        out[i] = g;
        if (i < n_params) acc[0] += static_cast<double>(g) * g;
    }
    block_accumulate<1>(acc, sumsq);
}

__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, float scale, const double* __restrict__ sumsq,
                                                        float max_norm, const float* __restrict__ lr, float b1, float b2, float eps,
                                                        float bc1, float bc2_sqrt) {
    const float total = static_cast<float>(sqrt(*sumsq));
    const float coef = fminf(max_norm / (total + 1e-6f), 1.0f) * scale;   // clip_grad_norm_
    const float step_size = *lr / bc1;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gi = g[i] * coef;
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi, v[i] = vi;
        p[i] -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

// graph-replayable optimiser step: the step count lives in device memory; this 1-thread kernel advances it and zeroes the
// gradient-norm accumulator in front of grad_sumsq, clip_adam_dev_kernel derives the bias corrections from it.
__global__ void adam_prepare_kernel(int* __restrict__ step, double* __restrict__ sumsq) {
    *step += 1;
    *sumsq = 0.0;
}

__global__ void __launch_bounds__(256) clip_adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                            float* __restrict__ v, long long n, float scale,
                                                            const double* __restrict__ sumsq, float max_norm,
                                                            const float* __restrict__ lr, float b1, float b2, float eps,
                                                            const int* __restrict__ step) {
    const float st = static_cast<float>(*step);
    const float bc1 = 1.0f - powf(b1, st), bc2_sqrt = sqrtf(1.0f - powf(b2, st));
    const float total = static_cast<float>(sqrt(*sumsq));
    const float coef = fminf(max_norm / (total + 1e-6f), 1.0f) * scale;   // clip_grad_norm_
    const float step_size = *lr / bc1;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gi = g[i] * coef;
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi, v[i] = vi;
        p[i] -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

__global__ void adaptive_lr_kernel(const double* __restrict__ kl_sum, double count, float desired_kl, float* __restrict__ lr) {
    const float kl_mean = static_cast<float>(*kl_sum / count);   // ppo.py: adaptive schedule
    float l = *lr;
    if (kl_mean > desired_kl * 2.0f) l = fmaxf(1e-5f, l / 1.5f);
    else if (kl_mean < desired_kl / 2.0f && kl_mean > 0.0f) l = fminf(1e-2f, l * 1.5f);
    *lr = l;
}

}  // namespace

extern "C" {

int b200gym_gae_returns(float* rewards, const float* values, const uint8_t* dones, const uint8_t* time_outs,
                        const float* last_values, float* returns, float* advantages, double* stats, int32_t T, int32_t N,
                        float gamma, float lam, void* stream) {
    B200_REQUIRE(rewards && values && dones && last_values && returns && advantages && stats, B200GYM_EINVAL, "gae_returns: null argument");
    B200_REQUIRE(T > 0 && N > 0, B200GYM_EINVAL, "gae_returns: T and N must be positive (got %d, %d)", T, N);
    const int block = N <= 148 * 256 ? 64 : 256;   // small rollouts: spread the envs over as many SMs as possible
    gae_returns_kernel<<<(N + block - 1) / block, block, 0, static_cast<cudaStream_t>(stream)>>>(rewards, values, dones, time_outs,
                                                                                               last_values, returns, advantages, stats, T, N,
                                                                                               gamma, lam);
    B200_LAUNCH_CHECK("gae_returns");
    return B200GYM_OK;
}

int b200gym_adv_normalize(float* advantages, const double* stats, int64_t count, void* stream) {
    B200_REQUIRE(advantages && stats && count > 0, B200GYM_EINVAL, "adv_normalize: bad argument");
    const int grid = static_cast<int>(count / 1024 < 1 ? 1 : (count / 1024 > 148 * 16 ? 148 * 16 : count / 1024));
    adv_normalize_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(advantages, stats, count);
    B200_LAUNCH_CHECK("adv_normalize");
    return B200GYM_OK;
}

int b200gym_gather_rows(void* const* dst, const void* const* src, const int32_t* row_bytes, int32_t n_tensors, const int64_t* idx,
                        int64_t n_rows, void* stream) {
    B200_REQUIRE(dst && src && row_bytes && idx, B200GYM_EINVAL, "gather_rows: null argument");
    B200_REQUIRE(n_tensors > 0 && n_tensors <= MAX_GATHER && n_rows > 0, B200GYM_EINVAL, "gather_rows: 1..%d tensors, n_rows > 0", MAX_GATHER);
    GatherArgs a;
    a.n = n_tensors;
    for (int k = 0; k < n_tensors; ++k) {
        B200_REQUIRE(dst[k] && src[k] && row_bytes[k] > 0, B200GYM_EINVAL, "gather_rows: tensor %d is null / empty", k);
        B200_REQUIRE((row_bytes[k] & 15) != 0 || (b200_aligned16(dst[k]) && b200_aligned16(src[k])), B200GYM_EALIGN,
                     "gather_rows: tensor %d must be 16-byte aligned", k);
        a.dst[k] = dst[k], a.src[k] = src[k], a.row_bytes[k] = row_bytes[k];
    }
    const long long threads = n_rows * 32;
    gather_rows_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        a, reinterpret_cast<const long long*>(idx), n_rows);
    B200_LAUNCH_CHECK("gather_rows");
    return B200GYM_OK;
}

int b200gym_ppo_loss(const B200PpoLossParams* p, const float* mu, const float* std, const float* value, const float* actions,
                     const float* old_log_prob, const float* advantages, const float* returns, const float* old_values,
                     const float* old_mu, const float* old_sigma, float* d_mu, float* d_value, float* d_std, double* scalars,
                     void* stream) {
    B200_REQUIRE(p && mu && std && value && actions && old_log_prob && advantages && returns && old_values && old_mu && old_sigma &&
                     d_mu && d_value && d_std && scalars,
                 B200GYM_EINVAL, "ppo_loss: null argument");
    B200_REQUIRE(p->batch > 0 && p->num_actions > 0 && p->num_actions <= MAXA, B200GYM_EINVAL, "ppo_loss: batch > 0, 1..%d actions", MAXA);
    ppo_loss_kernel<<<(p->batch + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        *p, mu, std, value, actions, old_log_prob, advantages, returns, old_values, old_mu, old_sigma, d_mu, d_value, d_std, scalars);
    B200_LAUNCH_CHECK("ppo_loss");
    return B200GYM_OK;
}

int b200gym_grad_sumsq(const float* grad, int64_t n, float grad_scale, double* sumsq, void* stream) {
    B200_REQUIRE(grad && sumsq && n > 0, B200GYM_EINVAL, "grad_sumsq: bad argument");
    const int grid = static_cast<int>((n + 2047) / 2048 > 148 * 8 ? 148 * 8 : (n + 2047) / 2048);
    grad_sumsq_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(grad, n, grad_scale, sumsq);
    B200_LAUNCH_CHECK("grad_sumsq");
    return B200GYM_OK;
}

int b200gym_grad_reduce_peers(const B200PeerPtrs* peers, int32_t world, float* out, int64_t n_total, int64_t n_params, double* sumsq,
                              void* stream) {
    B200_REQUIRE(peers && out && sumsq && n_total > 0 && n_params >= 0 && n_params <= n_total, B200GYM_EINVAL, "grad_reduce_peers: bad argument");
    B200_REQUIRE(world >= 1 && world <= B200GYM_MAX_PEERS, B200GYM_EINVAL, "grad_reduce_peers: 1..%d ranks (got %d)", B200GYM_MAX_PEERS, world);
    for (int r = 0; r < world; ++r) B200_REQUIRE(peers->ptr[r] != nullptr, B200GYM_EINVAL, "grad_reduce_peers: peer %d has no buffer", r);
    const int grid = static_cast<int>((n_total + 255) / 256 > 148 * 4 ? 148 * 4 : (n_total + 255) / 256);
    grad_reduce_peers_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*peers, world, out, n_total, n_params, sumsq);
    B200_LAUNCH_CHECK("grad_reduce_peers");
    return B200GYM_OK;
}

int b200gym_clip_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float grad_scale,
                      const double* sumsq, float max_norm, const float* lr, float beta1, float beta2, float eps, int32_t step,
                      void* stream) {
    B200_REQUIRE(param && grad && exp_avg && exp_avg_sq && sumsq && lr && n > 0 && step > 0, B200GYM_EINVAL, "clip_adam: bad argument");
    const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
    const float bc2 = sqrtf(1.0f - powf(beta2, static_cast<float>(step)));
    const int grid = static_cast<int>((n + 1023) / 1024 > 148 * 8 ? 148 * 8 : (n + 1023) / 1024);
    clip_adam_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, grad_scale, sumsq, max_norm, lr,
                                                                        beta1, beta2, eps, bc1, bc2);
    B200_LAUNCH_CHECK("clip_adam");
    return B200GYM_OK;
}

int b200gym_adam_prepare(int32_t* step_dev, double* sumsq, void* stream) {
    B200_REQUIRE(step_dev && sumsq, B200GYM_EINVAL, "adam_prepare: null argument");
    adam_prepare_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(step_dev, sumsq);
    B200_LAUNCH_CHECK("adam_prepare");
    return B200GYM_OK;
}

int b200gym_clip_adam_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float grad_scale,
                          const double* sumsq, float max_norm, const float* lr, float beta1, float beta2, float eps,
                          const int32_t* step_dev, void* stream) {
    B200_REQUIRE(param && grad && exp_avg && exp_avg_sq && sumsq && lr && step_dev && n > 0, B200GYM_EINVAL, "clip_adam_dev: bad argument");
    const int grid = static_cast<int>((n + 1023) / 1024 > 148 * 8 ? 148 * 8 : (n + 1023) / 1024);
    clip_adam_dev_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, grad_scale, sumsq, max_norm,
                                                                            lr, beta1, beta2, eps, step_dev);
    B200_LAUNCH_CHECK("clip_adam_dev");
    return B200GYM_OK;
}

int b200gym_adaptive_lr(const double* kl_sum, double count, float desired_kl, float* lr, void* stream) {
    B200_REQUIRE(kl_sum && lr && count > 0, B200GYM_EINVAL, "adaptive_lr: bad argument");
    adaptive_lr_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(kl_sum, count, desired_kl, lr);
    B200_LAUNCH_CHECK("adaptive_lr");
    return B200GYM_OK;
}

}  // extern "C"
