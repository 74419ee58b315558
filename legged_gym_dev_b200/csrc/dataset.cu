// Tube-dataset construction on the device: the consumers of the ROM rollout logs
// (deep_tube_learning/datasets.py:60-71 `get_slice` / `sliding_window`, evaluation/evaluate_tube_simple.py:28-46) build, per
// robot and time step, the tube error w = ||Pz(x) - z|| and an N-deep window of past samples.  With the epoch logs already in
// HBM (b200gym_rom_rollout) the round trip through pickle + numpy disappears; both kernels are pure data movement (bit-exact).
#include "common.cuh"
#include "../../include/b200gym.h"

namespace {

// w[b, t] = || pz_x[b, t, :] - z[b, t, :] ||_2 for t < T (the logs carry T1 >= T samples per robot); numpy's fp32 norm for a
// short last axis is sqrt of the left-to-right sum of squares, reproduced with separately rounded operations.
__global__ void __launch_bounds__(256) tube_error_kernel(const float* __restrict__ z, const float* __restrict__ pz_x, float* __restrict__ w,
                                                         long long B, int T, int T1, int n) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= B * T) return;
    const long long b = i / T;
    const int t = static_cast<int>(i - b * T);
    const float* zp = z + (b * T1 + t) * n;
    const float* pp = pz_x + (b * T1 + t) * n;
    float acc = 0.0f;
    for (int k = 0; k < n; ++k) {
        const float d = sub_rn(pp[k], zp[k]);
        acc = add_rn(acc, mul_rn(d, d));
    }
    w[i] = sqrtf(acc);
}

// out[b, t, i*D + d] = get_slice(data, i, dN, m)[b, t, d]   (datasets.py:60-71), i < N:
//   L_i = number of samples T-1-i*dN, T-1-i*dN-dN, ... >= 0;  rows t < T - L_i repeat data[b, 0, :] with its last m columns
//   zeroed (`start[:, :, -m:] = 0`: m == 0 zeroes the whole row, as numpy's [-0:] does), the remaining rows are those samples
//   in increasing time order.
// One CTA per robot: the robot's T x D source rows (1.6 KB for T = 200, D = 2) are staged in shared memory with coalesced loads,
// then the CTA streams the robot's T x N x D output block (16 KB) as 128-bit stores.  A thread decodes (t, slice, d) of the FIRST
// float of its float4 with two 32-bit divisions and walks the other three with carries; no 64-bit index arithmetic in the loop.
// (First version: one thread per (robot, time, slice) with 64-bit div/mod and 4-byte stores — 37 % of HBM, issue-bound.)
// DT = compile-time D (0: run time), DN1 = (dN == 1): the common shapes keep the per-element index math to a few instructions
template <int DT, bool DN1>
__global__ void __launch_bounds__(256) sliding_window_kernel(const float* __restrict__ data, float* __restrict__ out, long long B, int T, int D_,
                                                             int N, int dN, int m) {
    extern __shared__ float srow[];                               // [T * D]
    const int D = DT ? DT : D_;
    const int row_in = T * D, row_out = T * N * D;
    const unsigned ND_ = static_cast<unsigned>(N * D);
    const int dz = m == 0 ? 0 : D - m;                            // columns >= dz of a padding row are zero
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        const float* src = data + b * row_in;
        __syncthreads();                                          // previous robot's readers are done
        for (int k = threadIdx.x; k < row_in; k += blockDim.x) srow[k] = __ldg(src + k);
        __syncthreads();
        float* dst = out + b * row_out;
        auto value = [&](int t, int i, int d) -> float {
            if (DN1) return t < i ? (d >= dz ? 0.0f : srow[d]) : srow[(t - i) * D + d];   // L_i = T - i, pad = i, source row t - i
            const int last = T - i * dN - 1;                      // newest sample of slice i
            const int L = last >= 0 ? last / dN + 1 : 0;
            const int pad = T - L;
            if (t < pad) return d >= dz ? 0.0f : srow[d];
            return srow[(last - (L - 1 - (t - pad)) * dN) * D + d];
        };
        if ((row_out & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            for (unsigned q = threadIdx.x; q < static_cast<unsigned>(row_out) / 4; q += blockDim.x) {
                const unsigned e0 = 4 * q;
                int t = static_cast<int>(e0 / ND_);
                const unsigned r = e0 - static_cast<unsigned>(t) * ND_;
                int i = static_cast<int>(r / static_cast<unsigned>(D)), d = static_cast<int>(r) - i * D;
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    v[k] = value(t, i, d);
                    if (++d == D) {
                        d = 0;
                        if (++i == N) i = 0, ++t;
                    }
                }
                stg_stream4(reinterpret_cast<float4*>(dst) + q, make_float4(v[0], v[1], v[2], v[3]));
            }
        } else {
            for (unsigned e = threadIdx.x; e < static_cast<unsigned>(row_out); e += blockDim.x) {
                const int t = static_cast<int>(e / ND_), r = static_cast<int>(e) - t * static_cast<int>(ND_), i = r / D;
                dst[e] = value(t, i, r - i * D);
            }
        }
    }
}

template <int DT, bool DN1>
cudaError_t launch_sliding_window(unsigned grid, size_t smem, cudaStream_t st, const float* data, float* out, long long B, int T, int D, int N,
                                  int dN, int m) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sliding_window_kernel<DT, DN1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    sliding_window_kernel<DT, DN1><<<grid, 256, smem, st>>>(data, out, B, T, D, N, dN, m);
    return cudaSuccess;
}

}  // namespace

extern "C" {

int b200gym_tube_error(const float* z, const float* pz_x, float* w, int64_t B, int32_t T, int32_t T1, int32_t n, void* stream) {
    B200_REQUIRE(z && pz_x && w, B200GYM_EINVAL, "tube_error: null argument");
    B200_REQUIRE(B > 0 && T > 0 && T1 >= T && n > 0, B200GYM_EINVAL, "tube_error: need B > 0, 0 < T <= T1, n > 0");
    const long long total = static_cast<long long>(B) * T;
    tube_error_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(z, pz_x, w, B, T, T1, n);
    B200_LAUNCH_CHECK("tube_error");
    return B200GYM_OK;
}

int b200gym_sliding_window(const float* data, float* out, int64_t B, int32_t T, int32_t D, int32_t N, int32_t dN, int32_t m, void* stream) {
    B200_REQUIRE(data && out, B200GYM_EINVAL, "sliding_window: null argument");
    B200_REQUIRE(B > 0 && T > 0 && D > 0 && N > 0 && dN > 0 && m >= 0 && m <= D, B200GYM_EINVAL,
                 "sliding_window: need B, T, D, N, dN > 0 and 0 <= m <= D");
    const size_t smem = static_cast<size_t>(T) * D * sizeof(float);
    B200_REQUIRE(smem <= 200 * 1024 && static_cast<long long>(T) * N * D < (1LL << 30), B200GYM_EINVAL,
                 "sliding_window: one robot's rows must fit shared memory (T * D * 4 <= 200 KB)");
    const unsigned grid = static_cast<unsigned>(B < 148LL * 32 ? B : 148LL * 32);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (dN == 1 && D == 2) e = launch_sliding_window<2, true>(grid, smem, st, data, out, B, T, D, N, dN, m);
    else if (dN == 1 && D == 3) e = launch_sliding_window<3, true>(grid, smem, st, data, out, B, T, D, N, dN, m);
    else if (dN == 1 && D == 4) e = launch_sliding_window<4, true>(grid, smem, st, data, out, B, T, D, N, dN, m);
    else if (dN == 1) e = launch_sliding_window<0, true>(grid, smem, st, data, out, B, T, D, N, dN, m);
    else e = launch_sliding_window<0, false>(grid, smem, st, data, out, B, T, D, N, dN, m);
    B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "sliding_window: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    B200_LAUNCH_CHECK("sliding_window");
    return B200GYM_OK;
}

}  // extern "C"
