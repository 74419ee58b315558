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
__global__ void __launch_bounds__(256) sliding_window_kernel(const float* __restrict__ data, float* __restrict__ out, long long B, int T, int D,
                                                             int N, int dN, int m) {
    // one thread per (robot, time, slice): it copies the D contiguous floats of its source row; a warp writes 32*D contiguous floats
    const long long items = B * T * static_cast<long long>(N);
    for (long long it = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; it < items;
         it += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long bt = it / N;
        const int i = static_cast<int>(it - bt * N);
        const long long b = bt / T;
        const int t = static_cast<int>(bt - b * T);
        const int last = T - i * dN - 1;                         // newest sample of slice i
        const int L = last >= 0 ? last / dN + 1 : 0;
        const int pad = T - L;
        float* o = out + it * D;
        if (t < pad) {
            const float* s0 = data + (b * T) * D;
            for (int d = 0; d < D; ++d) o[d] = (m == 0 || d >= D - m) ? 0.0f : s0[d];
        } else {
            const float* src = data + (b * T + (last - (L - 1 - (t - pad)) * dN)) * D;
            for (int d = 0; d < D; ++d) o[d] = src[d];
        }
    }
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
    const long long total = static_cast<long long>(B) * T * N;
    const long long blocks = (total + 255) / 256;
    const unsigned grid = static_cast<unsigned>(blocks < 148LL * 64 ? blocks : 148LL * 64);
    sliding_window_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(data, out, B, T, D, N, dN, m);
    B200_LAUNCH_CHECK("sliding_window");
    return B200GYM_OK;
}

}  // extern "C"
