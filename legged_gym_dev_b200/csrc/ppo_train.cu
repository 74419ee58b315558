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
#include <stdlib.h>
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
    pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
    pdl_wait();
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
    pdl_launch_dependents();
    pdl_wait();
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


// clip + Adam + fp16 operand copies over the flat buffers, elements i0, i0 + stride, ...; the loads of four elements are issued
// together (one CTA of the 34 K-parameter nets walks 66 elements per thread: a serial load -> store chain would cost ~1 us each).
// g_src: summed gradients (read through L2: another CTA may have written them); g_clear: buffer to zero behind the read, or null.
__device__ __forceinline__ void adam_pack_pass(const B200OptParams& p, const B200PackTable& tab, long long i0, long long stride,
                                               const float* __restrict__ g_src, float* __restrict__ g_clear, float coef, float step_size,
                                               float bc2_sqrt, float* __restrict__ param, float* __restrict__ m, float* __restrict__ v,
                                               __half* __restrict__ w16) {
    constexpr int U = 4;
    for (long long base = i0; base < p.n; base += U * stride) {
        float g[U], mm[U], vv[U], w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = base + u * stride;
            const bool ok = i < p.n;
            g[u] = ok ? __ldcg(g_src + i) : 0.0f;
            mm[u] = ok ? m[i] : 0.0f;
            vv[u] = ok ? v[i] : 0.0f;
            w[u] = ok ? param[i] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = base + u * stride;
            if (i >= p.n) break;
            const float gi = g[u] * coef;
            const float mi = p.beta1 * mm[u] + (1.0f - p.beta1) * gi;
            const float vi = p.beta2 * vv[u] + (1.0f - p.beta2) * gi * gi;
            m[i] = mi, v[i] = vi;
            const float wn = w[u] - step_size * mi / (sqrtf(vi) / bc2_sqrt + p.eps);
            param[i] = wn;
            if (g_clear) g_clear[i] = 0.0f;
            int e = 0;
            while (e < tab.n && !(i >= tab.e[e].src_off && i - tab.e[e].src_off < static_cast<long long>(tab.e[e].rows) * tab.e[e].cols)) ++e;
            if (e < tab.n) {
                const B200PackEntry& t = tab.e[e];
                const unsigned local = static_cast<unsigned>(i - t.src_off);
                const unsigned r = local / static_cast<unsigned>(t.cols), k = local - r * static_cast<unsigned>(t.cols);
                const long long o = t.layout == 0 ? t.dst_off + static_cast<long long>(r) * t.ld + k
                                                  : t.dst_off + (static_cast<long long>(k >> 3) * t.ld + r) * 8 + (k & 7);
                w16[o] = __float2half_rn(wn);
            }
        }
    }
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
    // programmatic dependent launch: this grid is scheduled while the minibatch kernel drains and lets the next minibatch kernel be
    // scheduled behind it; everything it reads was written by the launches before it, so wait first
    pdl_launch_dependents();
    pdl_wait();
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // ---- phase 1: squared gradient norm ----
    double acc = 0.0;
#pragma unroll 4
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
    adam_pack_pass(p, tab, i0, stride, grad, grad, coef, step_size, bc2_sqrt, param, m, v, w16);
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


// ------------------------------------------------------------------------------------------------------------------
// ppo_optimizer_step_peers_kernel — the data-parallel form of the launch above ("NCCL gradient allreduce"):
// gradient exchange over NVLink peer memory + rank-ordered sum + clip + KL schedule + Adam + fp16 copies in ONE launch per
// rank, with no host-side or torch barrier.
//   A  push   every rank stores its local gradients (+ {sum kl, count} in the tail) into slot [parity][rank] of EVERY rank's
//             symmetric buffer (posted NVLink writes), clears its local gradient buffer, then publishes the exchange number in
//             flags[rank] of every rank with a system-scope release store
//   B  wait   acquire-poll the LOCAL flags until every rank has published this exchange (bounded: a dead peer sets ws->error)
//   C  sum    slots [parity][0..world) summed in rank order -> bit-identical totals on all ranks; per-CTA partial squared
//             norms combined in CTA order after a device-wide barrier (deterministic: identical clip factor everywhere)
//   D  step   as the single-GPU kernel
// Slots are double-buffered by the parity of the exchange number: a rank can only overwrite parity q two exchanges later, i.e.
// after it has seen the flag of the exchange in between, which its owner publishes after finishing its reads of q.
// ------------------------------------------------------------------------------------------------------------------
struct PeerWs {
    unsigned int arrive, arrive2, depart, epoch, error, pad[3];
    double partial[B200GYM_OPT_MAX_CTAS];
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(512) ppo_optimizer_step_peers_kernel(const __grid_constant__ B200OptParams p, const __grid_constant__ B200PeerBases peers,
                                                                       int world, int rank, long long n_pad, float* __restrict__ param,
                                                                       float* __restrict__ grad, float* __restrict__ gsum, float* __restrict__ m,
                                                                       float* __restrict__ v, float* __restrict__ lr, int* __restrict__ step,
                                                                       double* __restrict__ mb, double* __restrict__ totals, PeerWs* __restrict__ ws,
                                                                       const __grid_constant__ B200PackTable tab, __half* __restrict__ w16) {
    __shared__ double sh[17];
    pdl_launch_dependents();   // as in ppo_optimizer_step_kernel
    pdl_wait();
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(&ws->epoch) + 1u;   // rewritten only after every CTA has passed barrier C
    const long long slot0 = static_cast<long long>(epoch & 1u) * world * n_pad;
    const long long n4 = n_pad >> 2;
    // ---- A: push ----
    for (long long i = i0; i < n4; i += stride) {
        float4 g = reinterpret_cast<const float4*>(grad)[i];
        const long long e = i << 2;
        if (e + 3 >= p.n) {   // the float4 that holds the piggy-backed tail: [n] = sum kl of this minibatch, [n + 1] = its sample count
            float t[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (e + j == p.n) t[j] = static_cast<float>(mb[0]);
                else if (e + j == p.n + 1) t[j] = static_cast<float>(p.count);
                else if (e + j > p.n + 1) t[j] = 0.0f;
            }
            g = make_float4(t[0], t[1], t[2], t[3]);
        }
        for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(peers.base[r] + slot0 + static_cast<long long>(rank) * n_pad)[i] = g;
        reinterpret_cast<float4*>(grad)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    for (long long i = n_pad + i0; i < p.n + 8; i += stride) grad[i] = 0.0f;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(&ws->arrive, 1u);
        if (blockIdx.x == 0) {
            while (atomicAdd(&ws->arrive, 0u) < gridDim.x) __nanosleep(32);   // all CTAs are resident (grid <= SM count)
            __threadfence_system();
        }
    }
    if (blockIdx.x == 0) {
        __syncthreads();
        if (threadIdx.x < world)
            st_release_sys(reinterpret_cast<unsigned int*>(peers.base[threadIdx.x] + 2ll * world * n_pad) + rank, epoch);
    }
    // ---- B: wait for every rank's publication of this exchange ----
    if (threadIdx.x < world) {
        const unsigned int* flag = reinterpret_cast<const unsigned int*>(peers.base[rank] + 2ll * world * n_pad) + threadIdx.x;
        const unsigned long long t0 = globaltimer_ns();
        while (static_cast<int>(ld_acquire_sys(flag) - epoch) < 0) {
            __nanosleep(64);
            if (globaltimer_ns() - t0 > 4000000000ull) {   // 4 s: a peer never arrived; flag it instead of hanging the device
                ws->error = 1u;
                break;
            }
        }
    }
    __syncthreads();
    // ---- C: rank-ordered sum + squared norm ----
    const float* mine = peers.base[rank] + slot0;
    double acc = 0.0;
    for (long long i = i0; i < n4; i += stride) {
        float4 s = __ldcg(reinterpret_cast<const float4*>(mine) + i);
        for (int r = 1; r < world; ++r) {
            const float4 a = __ldcg(reinterpret_cast<const float4*>(mine + static_cast<long long>(r) * n_pad) + i);
            s.x = __fadd_rn(s.x, a.x), s.y = __fadd_rn(s.y, a.y), s.z = __fadd_rn(s.z, a.z), s.w = __fadd_rn(s.w, a.w);
        }
        reinterpret_cast<float4*>(gsum)[i] = s;
        const long long e = i << 2;
        const float t[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (e + j < p.n) acc += static_cast<double>(t[j]) * t[j];
    }
    acc = warp_sum_d(acc);
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += sh[w];
        ws->partial[blockIdx.x] = t;
        __threadfence();
        atomicAdd(&ws->arrive2, 1u);
        while (atomicAdd(&ws->arrive2, 0u) < gridDim.x) __nanosleep(32);
        __threadfence();
        double tot = 0.0;
        for (unsigned int c = 0; c < gridDim.x; ++c) tot += __ldcg(&ws->partial[c]);   // CTA order: the same total on every rank
        sh[16] = tot;
    }
    __syncthreads();
    // ---- D: schedule, clip, Adam, fp16 copies ----
    const double sumsq = sh[16];
    float l = *lr;
    if (p.adaptive) {   // KL mean over the GLOBAL minibatch (all ranks), from the summed tail
        const float kl_mean = __ldcg(gsum + p.n) / __ldcg(gsum + p.n + 1);
        if (kl_mean > p.desired_kl * 2.0f) l = fmaxf(1e-5f, l / 1.5f);
        else if (kl_mean < p.desired_kl / 2.0f && kl_mean > 0.0f) l = fminf(1e-2f, l * 1.5f);
    }
    const float st = static_cast<float>(*step + 1);
    const float bc1 = 1.0f - powf(p.beta1, st), bc2_sqrt = sqrtf(1.0f - powf(p.beta2, st));
    const float total = static_cast<float>(sqrt(sumsq));
    const float coef = fminf(p.max_grad_norm / (total + 1e-6f), 1.0f);
    const float step_size = l / bc1;
    adam_pack_pass(p, tab, i0, stride, gsum, nullptr, coef, step_size, bc2_sqrt, param, m, v, w16);
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
            totals[4] = sumsq;
            ws->arrive = 0u;
            ws->arrive2 = 0u;
            ws->depart = 0u;
            ws->epoch = epoch;
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
    b200_launch_pdl(0, rows_to_f16_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), src, src_ld,
                    cols, reinterpret_cast<const long long*>(idx), static_cast<__half*>(dst), dst_ld, n_rows);
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
    b200_launch_pdl(0, ppo_loss_gathered_kernel, dim3((p->batch + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream), *p,
                    reinterpret_cast<const long long*>(idx), mu_out, ld_mu, value_out, ld_value, std, actions, old_log_prob, advantages, returns,
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
    // ~4 parameters per thread (one batch of adam_pack_pass); B200GYM_OPT_CTAS overrides (A/B)
    static const int forced = getenv("B200GYM_OPT_CTAS") ? atoi(getenv("B200GYM_OPT_CTAS")) : 0;
    long long grid = forced > 0 ? forced : (p->n + 512 * 4 - 1) / (512 * 4);
    grid = grid < 1 ? 1 : (grid > sms ? sms : grid);     // never more CTAs than SMs: the device-wide barrier needs co-residency
    b200_launch_pdl(0, ppo_optimizer_step_kernel, dim3(static_cast<unsigned>(grid)), dim3(512), 0, static_cast<cudaStream_t>(stream), *p, param, grad,
                    exp_avg, exp_avg_sq, lr, step_dev, mb_scalars, totals, static_cast<OptWs*>(workspace), *table, static_cast<__half*>(w16));
    B200_LAUNCH_CHECK("ppo_optimizer_step");
    return B200GYM_OK;
}

int b200gym_ppo_optimizer_step_peers(const B200OptParams* p, const B200PeerBases* peers, int32_t world, int32_t rank, int64_t n_pad, float* param,
                                     float* grad, float* grad_sum, float* exp_avg, float* exp_avg_sq, float* lr, int32_t* step_dev,
                                     double* mb_scalars, double* totals, void* workspace, const B200PackTable* table, void* w16, int32_t ctas,
                                     void* stream) {
    B200_REQUIRE(p && peers && param && grad && grad_sum && exp_avg && exp_avg_sq && lr && step_dev && mb_scalars && totals && workspace && table && w16,
                 B200GYM_EINVAL, "ppo_optimizer_step_peers: null argument");
    B200_REQUIRE(p->n > 0 && p->count > 0 && table->n >= 0 && table->n <= B200GYM_PACK_MAX, B200GYM_EINVAL, "ppo_optimizer_step_peers: bad argument");
    B200_REQUIRE(world >= 1 && world <= B200GYM_MAX_PEERS && rank >= 0 && rank < world, B200GYM_EINVAL,
                 "ppo_optimizer_step_peers: 1..%d ranks, 0 <= rank < world (got %d of %d)", B200GYM_MAX_PEERS, rank, world);
    B200_REQUIRE(n_pad % 4 == 0 && n_pad >= p->n + 2 && n_pad <= p->n + 8, B200GYM_EINVAL,
                 "ppo_optimizer_step_peers: n_pad must be a multiple of 4 in [n + 2, n + 8]");
    for (int r = 0; r < world; ++r)
        B200_REQUIRE(peers->base[r] != nullptr && b200_aligned16(peers->base[r]), B200GYM_EALIGN, "ppo_optimizer_step_peers: peer %d has no 16-byte aligned buffer", r);
    B200_REQUIRE(b200_aligned16(grad) && b200_aligned16(grad_sum), B200GYM_EALIGN, "ppo_optimizer_step_peers: grad / grad_sum must be 16-byte aligned");
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = ctas > 0 ? ctas : (p->n + 512 * 4 - 1) / (512 * 4);
    const long long cap = sms < B200GYM_OPT_MAX_CTAS ? sms : B200GYM_OPT_MAX_CTAS;
    grid = grid < 1 ? 1 : (grid > cap ? cap : grid);   // co-residency of the device-wide barriers
    b200_launch_pdl(0, ppo_optimizer_step_peers_kernel, dim3(static_cast<unsigned>(grid)), dim3(512), 0, static_cast<cudaStream_t>(stream), *p, *peers,
                    world, rank, n_pad, param, grad, grad_sum, exp_avg, exp_avg_sq, lr, step_dev, mb_scalars, totals,
                    static_cast<PeerWs*>(workspace), *table, static_cast<__half*>(w16));
    B200_LAUNCH_CHECK("ppo_optimizer_step_peers");
    return B200GYM_OK;
}

}  // extern "C"
