// The non-GEMM kernels of the tensor-core PPO minibatch step (rsl_rl PPO.update; the contractions are
// csrc/gemm.cu):
//   rows_to_f16_kernel        mini_batch_generator's observation gather, fused with the fp32 -> fp16 conversion and the zero
//                             padding of the first layer's A operand (also used without indices for the rollout forward)
//   ppo_loss_gathered_kernel  the PPO loss forward + its gradient w.r.t. (mu, value); the per-sample storage columns (actions,
//                             old log-prob, advantage, return, old value, old mu / sigma) are fetched through the minibatch
//                             indices here instead of being gathered into temporaries first.  Gradients leave UNSCALED
//                             (no 1/batch factor) as the fp16 dZ operands of the backward GEMMs — the 1/batch goes into the
//                             weight-gradient epilogue, so fp16 never sees 1e-5-sized values
//   pack_params_f16_kernel    fp32 master weights (one flat buffer) -> the fp16 operand copies, all layers in one launch
#include <cuda_fp16.h>
#include "common.cuh"
#include "../../include/b200gym.h"

namespace {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// one thread per 16-byte output piece (8 columns)
__global__ void __launch_bounds__(256) rows_to_f16_kernel(const float* __restrict__ src, long long src_ld, int cols, const long long* __restrict__ idx,
                                                          __half* __restrict__ dst, int dst_ld, long long n_rows) {
    const int pieces = dst_ld >> 3;
    const long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (q >= n_rows * pieces) return;
    const long long row = q / pieces;
    const int c0 = static_cast<int>(q - row * pieces) * 8;
    const float* s = src + (idx ? idx[row] : row) * src_ld + c0;
    float v[8];
    if (c0 + 8 <= cols && (src_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s) + 1);
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = c0 + j < cols ? __ldg(s + j) : 0.0f;
    }
    *reinterpret_cast<uint4*>(dst + row * dst_ld + c0) = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
}

constexpr int MAXA = 16;

__global__ void __launch_bounds__(256) ppo_loss_gathered_kernel(const __grid_constant__ B200PpoLossParams p, const long long* __restrict__ idx,
                                                                const float* __restrict__ mu_out, int ld_mu, const float* __restrict__ v_out, int ld_v,
                                                                const float* __restrict__ stdv, const float* __restrict__ actions,
                                                                const float* __restrict__ old_logp, const float* __restrict__ adv,
                                                                const float* __restrict__ ret, const float* __restrict__ old_v,
                                                                const float* __restrict__ old_mu, const float* __restrict__ old_sigma,
                                                                __half* __restrict__ dz_actor, __half* __restrict__ dz_critic, float* __restrict__ d_std,
                                                                double* __restrict__ scalars) {
    const int A = p.num_actions;
    __shared__ float s_dstd[MAXA];
    __shared__ double s_acc[4][8];
    if (threadIdx.x < MAXA) s_dstd[threadIdx.x] = 0.0f;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    float dstd[MAXA];
#pragma unroll
    for (int a = 0; a < MAXA; ++a) dstd[a] = 0.0f;
    if (i < p.batch) {
        const long long s = idx ? idx[i] : i;
        const float LOG_SQRT_2PI = 0.91893853320467274178f;
        float logp = 0.0f, ent = 0.0f, kl = 0.0f;
        float diff[MAXA];
#pragma unroll
        for (int a = 0; a < MAXA; ++a) {
            diff[a] = 0.0f;
            if (a < A) {
                const float sg = stdv[a], m = mu_out[static_cast<size_t>(i) * ld_mu + a], x = actions[s * A + a];
                const float ls = logf(sg);
                diff[a] = x - m;
                logp += -(diff[a] * diff[a]) / (2.0f * sg * sg) - ls - LOG_SQRT_2PI;
                ent += 0.5f + LOG_SQRT_2PI + ls;
                const float os = old_sigma[s * A + a], dm = old_mu[s * A + a] - m;
                kl += logf(sg / os + 1.e-5f) + (os * os + dm * dm) / (2.0f * sg * sg) - 0.5f;
            }
        }
        const float Ai = adv[s], ratio = expf(logp - old_logp[s]);
        const float lo = 1.0f - p.clip_param, hi = 1.0f + p.clip_param;
        const float s1 = -Ai * ratio, s2 = -Ai * fminf(fmaxf(ratio, lo), hi);
        const bool inside = ratio >= lo && ratio <= hi;
        const float g_ratio = s1 > s2 ? -Ai : (s1 == s2 ? (inside ? -Ai : -0.5f * Ai) : 0.0f);   // torch.max splits ties
        const float dlogp = g_ratio * ratio;   // unscaled: the 1/batch factor is applied by the weight-gradient epilogue
        const float v = v_out[static_cast<size_t>(i) * ld_v], R = ret[s];
        float vloss, dv;
        if (p.use_clipped_value_loss) {
            const float ov = old_v[s], dvo = v - ov;
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
        float dmu[MAXA];
#pragma unroll
        for (int a = 0; a < MAXA; ++a) {
            dmu[a] = 0.0f;
            if (a < A) {
                const float sg = stdv[a], inv2 = 1.0f / (sg * sg);
                dmu[a] = dlogp * diff[a] * inv2;
                dstd[a] = (dlogp * (diff[a] * diff[a] * inv2 / sg - 1.0f / sg) - p.entropy_coef / sg) * p.inv_global_batch;
            }
        }
        uint4* za = reinterpret_cast<uint4*>(dz_actor + static_cast<size_t>(i) * MAXA);
        za[0] = make_uint4(pack_h2(dmu[0], dmu[1]), pack_h2(dmu[2], dmu[3]), pack_h2(dmu[4], dmu[5]), pack_h2(dmu[6], dmu[7]));
        za[1] = make_uint4(pack_h2(dmu[8], dmu[9]), pack_h2(dmu[10], dmu[11]), pack_h2(dmu[12], dmu[13]), pack_h2(dmu[14], dmu[15]));
        uint4* zc = reinterpret_cast<uint4*>(dz_critic + static_cast<size_t>(i) * MAXA);
        zc[0] = make_uint4(pack_h2(p.value_loss_coef * dv, 0.0f), 0u, 0u, 0u);
        zc[1] = make_uint4(0u, 0u, 0u, 0u);
        acc[0] = kl, acc[1] = fmaxf(s1, s2), acc[2] = vloss, acc[3] = ent;
    }
#pragma unroll
    for (int a = 0; a < MAXA; ++a) {
        if (a < A) {
            const float r = warp_sum_f(dstd[a]);
            if ((threadIdx.x & 31) == 0) atomicAdd(&s_dstd[a], r);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double r = warp_sum_d(acc[k]);
        if (lane == 0) s_acc[k][warp] = r;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_acc[threadIdx.x][w];
        if (t != 0.0) atomicAdd(scalars + threadIdx.x, t);
    }
    if (threadIdx.x < A) atomicAdd(d_std + threadIdx.x, s_dstd[threadIdx.x]);
}

__global__ void __launch_bounds__(256) pack_params_f16_kernel(const float* __restrict__ flat, const __grid_constant__ B200PackTable tab,
                                                              __half* __restrict__ dst) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= tab.total) return;
    int e = 0;
    while (e < tab.n - 1 && i >= tab.e[e].elem_end) ++e;
    const B200PackEntry& t = tab.e[e];
    const long long local = i - (e ? tab.e[e - 1].elem_end : 0);
    const int n = static_cast<int>(local / t.cols), k = static_cast<int>(local - static_cast<long long>(n) * t.cols);
    const float w = flat[t.src_off + local];
    long long o;
    if (t.layout == 0) o = t.dst_off + static_cast<long long>(n) * t.ld + k;                                   // row-major [rows_pad, ld]
    else o = t.dst_off + (static_cast<long long>(k >> 3) * t.ld + n) * 8 + (k & 7);                            // chunk-major [cols/8][ld rows][8]
    dst[o] = __float2half_rn(w);
}


// ------------------------------------------------------------------------------------------------------------------
// ppo_optimizer_step_kernel — everything of a PPO minibatch step that follows the backward pass, in ONE launch:
//   clip_grad_norm_ (squared norm over the flat gradient buffer) -> KL-adaptive learning rate -> Adam -> fp16 operand
//   copies of the updated weights -> gradient buffer and per-minibatch loss sums cleared for the next minibatch.
// (rsl_rl PPO.update: the `if self.desired_kl ...` schedule, nn.utils.clip_grad_norm_, optimizer.step(); replaces
// adam_prepare + grad_sumsq + adaptive_lr + clip_adam_dev + pack_params_f16 + four torch fill / copy launches.)
// The grid is at most one CTA per SM, so all CTAs are co-resident and a device-wide barrier between the norm and the
// update is safe: ws[0] = squared norm, counters (arrive, depart) behind it; the last CTA to leave resets them.
// ------------------------------------------------------------------------------------------------------------------
struct OptWs {
    double sumsq;
    unsigned int arrive, depart;
};

__global__ void __launch_bounds__(512) ppo_optimizer_step_kernel(const __grid_constant__ B200OptParams p, float* __restrict__ param,
                                                                 float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                                                                 float* __restrict__ lr, int* __restrict__ step, double* __restrict__ mb,
                                                                 double* __restrict__ totals, OptWs* __restrict__ ws,
                                                                 const __grid_constant__ B200PackTable tab, __half* __restrict__ w16) {
    __shared__ double sh[16];
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // ---- phase 1: squared gradient norm ----
    double acc = 0.0;
    for (long long i = i0; i < p.n; i += stride) {
        const double g = grad[i];
        acc += g * g;
    }
    acc = warp_sum_d(acc);
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += sh[w];
        atomicAdd(&ws->sumsq, t);
        __threadfence();
        atomicAdd(&ws->arrive, 1u);
        while (atomicAdd(&ws->arrive, 0u) < gridDim.x) __nanosleep(32);   // all CTAs are resident (grid <= SM count)
        __threadfence();
    }
    __syncthreads();
    // ---- phase 2: schedule, clip, Adam, fp16 copies ----
    const double sumsq = *reinterpret_cast<volatile double*>(&ws->sumsq);
    float l = *lr;
    if (p.adaptive) {   // ppo.py: adaptive schedule on the KL of THIS minibatch, before the optimiser step
        const float kl_mean = static_cast<float>(mb[0] / p.count);
        if (kl_mean > p.desired_kl * 2.0f) l = fmaxf(1e-5f, l / 1.5f);
        else if (kl_mean < p.desired_kl / 2.0f && kl_mean > 0.0f) l = fminf(1e-2f, l * 1.5f);
    }
    const float st = static_cast<float>(*step + 1);
    const float bc1 = 1.0f - powf(p.beta1, st), bc2_sqrt = sqrtf(1.0f - powf(p.beta2, st));
    const float total = static_cast<float>(sqrt(sumsq));
    const float coef = fminf(p.max_grad_norm / (total + 1e-6f), 1.0f);   // clip_grad_norm_
    const float step_size = l / bc1;
    for (long long i = i0; i < p.n; i += stride) {
        const float gi = grad[i] * coef;
        const float mi = p.beta1 * m[i] + (1.0f - p.beta1) * gi;
        const float vi = p.beta2 * v[i] + (1.0f - p.beta2) * gi * gi;
        m[i] = mi, v[i] = vi;
        const float w = param[i] - step_size * mi / (sqrtf(vi) / bc2_sqrt + p.eps);
        param[i] = w;
        grad[i] = 0.0f;
        for (int e = 0; e < tab.n; ++e) {
            const B200PackEntry& t = tab.e[e];
            const long long local = i - t.src_off;
            if (local >= 0 && local < static_cast<long long>(t.rows) * t.cols) {
                const unsigned r = static_cast<unsigned>(local) / static_cast<unsigned>(t.cols), k = static_cast<unsigned>(local) - r * static_cast<unsigned>(t.cols);
                const long long o = t.layout == 0 ? t.dst_off + static_cast<long long>(r) * t.ld + k
                                                  : t.dst_off + (static_cast<long long>(k >> 3) * t.ld + r) * 8 + (k & 7);
                w16[o] = __float2half_rn(w);
                break;
            }
        }
    }
    for (long long i = p.n + i0; i < p.n + 8; i += stride) grad[i] = 0.0f;   // spare tail of the flat buffer
    // ---- leave: the last CTA publishes the scalars and resets the workspace ----
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&ws->depart, 1u) == gridDim.x - 1) {
            *lr = l;
            *step += 1;
            for (int k = 0; k < 4; ++k) {
                totals[k] += mb[k];
                mb[k] = 0.0;
            }
            totals[4] = sumsq;   // last gradient norm^2 (FlatAdam.grad_norm)
            ws->sumsq = 0.0;
            ws->arrive = 0u;
            ws->depart = 0u;
        }
    }
}

}  // namespace

extern "C" {

int b200gym_rows_to_f16(const float* src, int64_t src_ld, int32_t cols, const int64_t* idx, void* dst, int32_t dst_ld, int64_t n_rows,
                        void* stream) {
    B200_REQUIRE(src && dst && n_rows > 0 && cols > 0, B200GYM_EINVAL, "rows_to_f16: bad argument");
    B200_REQUIRE(dst_ld % 8 == 0 && dst_ld >= cols && src_ld >= cols && b200_aligned16(dst), B200GYM_EALIGN,
                 "rows_to_f16: dst must be 16-byte aligned with dst_ld %% 8 == 0, dst_ld >= cols");
    const long long total = n_rows * (dst_ld / 8);
    rows_to_f16_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, src_ld, cols, reinterpret_cast<const long long*>(idx), static_cast<__half*>(dst), dst_ld, n_rows);
    B200_LAUNCH_CHECK("rows_to_f16");
    return B200GYM_OK;
}

int b200gym_ppo_loss_gathered(const B200PpoLossParams* p, const int64_t* idx, const float* mu_out, int32_t ld_mu, const float* value_out,
                              int32_t ld_value, const float* std, const float* actions, const float* old_log_prob, const float* advantages,
                              const float* returns, const float* old_values, const float* old_mu, const float* old_sigma, void* dz_actor,
                              void* dz_critic, float* d_std, double* scalars, void* stream) {
    B200_REQUIRE(p && mu_out && value_out && std && actions && old_log_prob && advantages && returns && old_values && old_mu && old_sigma &&
                     dz_actor && dz_critic && d_std && scalars,
                 B200GYM_EINVAL, "ppo_loss_gathered: null argument");
    B200_REQUIRE(p->batch > 0 && p->num_actions > 0 && p->num_actions <= MAXA && ld_mu >= p->num_actions && ld_value >= 1, B200GYM_EINVAL,
                 "ppo_loss_gathered: batch > 0, 1..%d actions", MAXA);
    B200_REQUIRE(b200_aligned16(dz_actor) && b200_aligned16(dz_critic), B200GYM_EALIGN, "ppo_loss_gathered: dZ buffers must be 16-byte aligned");
    ppo_loss_gathered_kernel<<<(p->batch + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        *p, reinterpret_cast<const long long*>(idx), mu_out, ld_mu, value_out, ld_value, std, actions, old_log_prob, advantages, returns,
        old_values, old_mu, old_sigma, static_cast<__half*>(dz_actor), static_cast<__half*>(dz_critic), d_std, scalars);
    B200_LAUNCH_CHECK("ppo_loss_gathered");
    return B200GYM_OK;
}

int b200gym_pack_params_f16(const float* flat, const B200PackTable* table, void* dst, void* stream) {
    B200_REQUIRE(flat && table && dst && table->n >= 1 && table->n <= B200GYM_PACK_MAX && table->total > 0, B200GYM_EINVAL,
                 "pack_params_f16: bad argument");
    pack_params_f16_kernel<<<static_cast<unsigned>((table->total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        flat, *table, static_cast<__half*>(dst));
    B200_LAUNCH_CHECK("pack_params_f16");
    return B200GYM_OK;
}

int b200gym_ppo_optimizer_step(const B200OptParams* p, float* param, float* grad, float* exp_avg, float* exp_avg_sq, float* lr,
                               int32_t* step_dev, double* mb_scalars, double* totals, void* workspace, const B200PackTable* table, void* w16,
                               void* stream) {
    B200_REQUIRE(p && param && grad && exp_avg && exp_avg_sq && lr && step_dev && mb_scalars && totals && workspace && table && w16,
                 B200GYM_EINVAL, "ppo_optimizer_step: null argument");
    B200_REQUIRE(p->n > 0 && p->count > 0 && table->n >= 0 && table->n <= B200GYM_PACK_MAX, B200GYM_EINVAL, "ppo_optimizer_step: bad argument");
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // small nets (the 34 K-parameter flat nets): ONE CTA, no device-wide barrier on the critical path; otherwise ~8 parameters per thread
    long long grid = p->n <= 65536 ? 1 : (p->n + 512 * 8 - 1) / (512 * 8);
    grid = grid < 1 ? 1 : (grid > sms ? sms : grid);     // never more CTAs than SMs: the device-wide barrier needs co-residency
    ppo_optimizer_step_kernel<<<static_cast<unsigned>(grid), 512, 0, static_cast<cudaStream_t>(stream)>>>(
        *p, param, grad, exp_avg, exp_avg_sq, lr, step_dev, mb_scalars, totals, static_cast<OptWs*>(workspace), *table, static_cast<__half*>(w16));
    B200_LAUNCH_CHECK("ppo_optimizer_step");
    return B200GYM_OK;
}

}  // extern "C"
