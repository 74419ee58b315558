// mlp_forward_kernel — the policy / critic MLP of rsl_rl's ActorCritic (Linear+ELU stack) as ONE
// launch on the 5th-generation tensor cores: tcgen05.mma (kind::tf32, M=128 rows per CTA tile, fp32 accumulators in
// TMEM), operands in shared memory in the canonical K-major no-swizzle UMMA layout, bias + ELU epilogue straight out
// of TMEM (tcgen05.ld), activations of layer l written back to shared memory as the A operand of layer l+1.  The
// whole weight set of the flat nets (48-128-64-32-12: 67 KB in fp32) stays resident in shared memory; CTAs are
// persistent over 128-row tiles.  Not under the 1e-5 contract (SURVEY.md §8d cfg 5: TF32 allowed for G4).
//
// Shared-memory operand layout (both A [rows x K] and B = W_l [N x K], K-major): 16-byte chunks of 4 consecutive k,
// chunk-major:  byte(r, k) = (k/4) * R*16 + r*16 + (k%4)*4,  R = rows of the operand.  In UMMA terms the 8x16B core
// matrices are contiguous (128 B), SBO (next 8 rows) = 128 B, LBO (next k-chunk) = R*16 B.
#include <stdlib.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "../../include/b200gym.h"

namespace {

constexpr int TM = 128;   // rows (samples) per tile = UMMA M = TMEM lanes

__device__ __forceinline__ uint64_t umma_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout NONE [61,64)
    return (static_cast<uint64_t>(smem_u32(smem) >> 4) & 0x3FFFull) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
           (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
    // cute::UMMA::InstrDescriptor: c_format F32=1 [4,6), a/b_format TF32=2 [7,10)/[10,13), K-major both, N>>3 [17,23), M>>4 [24,29)
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(TM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(TM, 1) mlp_forward_serial_kernel(const __grid_constant__ B200MlpParams p, const float* __restrict__ x,
                                                            const float* __restrict__ wpacked, const float* __restrict__ bias,
                                                            float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int L = p.num_layers, tid = threadIdx.x, warp = tid >> 5;
    int kmax = 0, wtot = 0, btot = 0, nmax = 0;
    for (int l = 0; l < L; ++l) {
        kmax = max(kmax, p.dims[l]);
        nmax = max(nmax, p.dims[l + 1]);
        wtot += p.dims[l] * p.dims[l + 1];
        btot += p.dims[l + 1];
    }
    float* sA = reinterpret_cast<float*>(smem);                 // [kmax/4][TM][4]
    float* sW = sA + static_cast<size_t>(TM) * kmax;            // per layer [K/4][N][4]
    float* sB = sW + wtot;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + ((btot + 3) & ~3));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    for (int i = tid * 4; i < wtot; i += TM * 4) *reinterpret_cast<float4*>(sW + i) = *reinterpret_cast<const float4*>(wpacked + i);
    for (int i = tid; i < btot; i += TM) sB[i] = bias[i];
    uint32_t ncols = 32;
    while (ncols < static_cast<uint32_t>(nmax)) ncols <<= 1;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t my_taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);   // this warp's 32 TMEM lanes
    uint32_t phase = 0;

    const int ntiles = (p.batch + TM - 1) / TM;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row = tile * TM + tid;
        const bool live = row < p.batch;
        // stage the input tile as the A operand (fp32 == tf32 container), zero-padded in rows and in k
        const int K0 = p.dims[0];
        for (int c = 0; c < K0 / 4; ++c) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                const float* src = x + static_cast<size_t>(row) * p.in_stride + 4 * c;
                if (4 * c + 3 < p.in_dim && (p.in_stride & 3) == 0) v = *reinterpret_cast<const float4*>(src);
                else {
                    if (4 * c + 0 < p.in_dim) v.x = src[0];
                    if (4 * c + 1 < p.in_dim) v.y = src[1];
                    if (4 * c + 2 < p.in_dim) v.z = src[2];
                    if (4 * c + 3 < p.in_dim) v.w = src[3];
                }
            }
            *reinterpret_cast<float4*>(sA + (static_cast<size_t>(c) * TM + tid) * 4) = v;
        }
        int woff = 0, boff = 0;
        for (int l = 0; l < L; ++l) {
            const int K = p.dims[l], N = p.dims[l + 1];
            fence_proxy_async();   // generic-proxy writes of the A operand -> visible to the tensor-core (async) proxy
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t idesc = umma_idesc_tf32(N);
                for (int ks = 0; ks < K / 8; ++ks) {   // UMMA K = 8 tf32 = two 16-byte chunks
                    const uint64_t da = umma_desc(sA + static_cast<size_t>(2 * ks) * TM * 4, TM * 16, 128);
                    const uint64_t db = umma_desc(sW + woff + static_cast<size_t>(2 * ks) * N * 4, N * 16, 128);
                    umma_tf32(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
                }
                umma_commit(bar);   // arrives on the mbarrier when every MMA above has completed (implies fence::before)
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            tc_fence_after();
            const bool last = (l == L - 1);
            for (int n0 = 0; n0 < N; n0 += 16) {
                float v[16];
                tmem_ld16(my_taddr + n0, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float a = v[j] + sB[boff + n0 + j];
                    if (!last) a = a > 0.0f ? a : expm1f(a);   // nn.ELU(alpha=1)
                    v[j] = a;
                }
                if (!last) {   // next layer's A operand: k = n, chunk-major
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(sA + (static_cast<size_t>(n0 / 4 + q) * TM + tid) * 4) =
                            make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                } else if (live) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + j < p.out_dim) out[static_cast<size_t>(row) * p.out_dim + n0 + j] = v[j];
                }
            }
            woff += K * N;
            boff += N;
        }
        tc_fence_before();   // order this tile's tcgen05.ld before the next tile's MMAs overwrite the accumulator
        __syncthreads();
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols) : "memory");
}


// ------------------------------------------------------------------------------------------------------------------
// mlp_forward_pipe_kernel — the same contraction as a warp-specialised, two-slot software pipeline (one persistent
// 512-thread CTA per SM).  The serial kernel above spends its time in a 4-warp epilogue (expm1f, dependent chains,
// nothing else resident on the SM: 25 us per 128-row tile, 146 GB/s); here
//   * two tile slots alternate: while slot 0's warpgroup runs its bias+ELU epilogue on the FMA/MUFU pipes, slot 1's
//     tcgen05.mma chain occupies the tensor pipe (and vice versa); each slot owns 128 TMEM columns and one operand buffer,
//   * two loader warpgroups prefetch the NEXT input tile of their slot into registers (12 x 128-bit loads in flight per
//     thread) and drop it into the slot's X region as soon as the layer that last reads that region has retired,
//   * ELU = max(t, ex2(min(t,0)*log2e) - 1): 1 MUFU + 5 FP32 instructions per activation, no branch,
//   * TMEM -> register loads are double-buffered against the epilogue arithmetic.
// Shared memory: weights resident (67 KB for 48-128-64-32-16) + 2 operand buffers [chunks][128 rows][4].
// Operand-buffer plan: layer l reads its A operand from chunks [0, K_l/4) (layer 0: the X region at chunk `x_off`) and
// writes ELU(h) to chunks [0, N_l/4).  x_off is chosen past every hidden activation written after layer 1, so the next
// tile's input can land while layers 2.. of the current tile are still running.
// ------------------------------------------------------------------------------------------------------------------
// optional event trace (tools/trace_mlp.py): clock64 stamps of CTA 0, written only when a buffer has been registered
__device__ unsigned long long* g_mlp_trace = nullptr;
__device__ __forceinline__ void trace_ev(unsigned long long*& cur, int code) {
    if (cur) {
        cur[0] = static_cast<unsigned long long>(code);
        cur[1] = clock64();
        cur += 2;
    }
}
constexpr int TRACE_EVENTS = 4096;   // per (slot, role) region, in (code, clock) pairs

constexpr int PIPE_THREADS = 800;   // warps 0-7 / 8-15: epilogue groups of slot 0 / 1 (two warps per TMEM lane quadrant, each takes half of the
                                    // columns); warps 16-19 / 20-23: loaders of slot 0 / 1; warp 24: MMA issue
constexpr int EPI_THREADS = 256;
constexpr int CHUNK_FLOATS = TM * 4 + 4;    // one k-chunk = 128 rows x 16 B, padded by 16 B: a row's chunks fall into different banks
constexpr int CHUNK_BYTES = CHUNK_FLOATS * 4;

struct PipePlan {
    int x_off;        // first chunk of the X region
    int free_layer;   // the X region is free again once this layer's MMAs have completed
    int buf_chunks;   // chunks per operand buffer
    int wtot, btot;
};

__device__ __forceinline__ void bar_sync_named(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// the registers are tied to the wait so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// nn.ELU(alpha=1): exp(t) - 1 >= t for every t, so elu(t) = max(t, exp(min(t, 0)) - 1)
__device__ __forceinline__ float elu_fast(float t) { return fmaxf(t, ex2_approx(fminf(t, 0.0f) * 1.4426950408889634f) - 1.0f); }

template <bool LAST>
__device__ __forceinline__ void pipe_epilogue_chunk(const uint32_t (&r)[16], const float* __restrict__ sBl, int n0, float* __restrict__ hrow,
                                                    float* __restrict__ orow, int out_dim, bool live) {
    float v[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b4 = *reinterpret_cast<const float4*>(sBl + n0 + 4 * q);   // same address in every lane: broadcast
        v[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + b4.x;
        v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + b4.y;
        v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + b4.z;
        v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + b4.w;
    }
    if (!LAST) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = elu_fast(v[j]);
#pragma unroll
        for (int q = 0; q < 4; ++q)   // next layer's A operand, k = n: chunk (n0/4 + q), this thread's row
            *reinterpret_cast<float4*>(hrow + static_cast<size_t>(n0 / 4 + q) * CHUNK_FLOATS) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else if (live) {
        if ((out_dim & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (n0 + 4 * q < out_dim)
                    *reinterpret_cast<float4*>(orow + n0 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (n0 + j < out_dim) orow[n0 + j] = v[j];
        }
    }
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// long waits (a loader waits a whole tile time for its X region): poll with a back-off instead of hammering the barrier
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    while (true) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(128);
    }
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {   // non-blocking phase test
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

struct LayerDesc {        // per layer, built once per CTA: everything the MMA warp needs
    uint64_t da0, db0;    // descriptors of the first K step (slot 0); slot 1 adds buf_bytes >> 4 to da0
    uint32_t idesc, ksteps, inc_b, pad;
};

__global__ void __launch_bounds__(PIPE_THREADS, 1) mlp_forward_pipe_kernel(const __grid_constant__ B200MlpParams p, const PipePlan plan,
                                                                            const float* __restrict__ x, const float* __restrict__ wpacked,
                                                                            const float* __restrict__ bias, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int L = p.num_layers, tid = threadIdx.x, warp = tid >> 5;
    const size_t buf_floats = static_cast<size_t>(plan.buf_chunks) * CHUNK_FLOATS;
    float* sH0 = reinterpret_cast<float*>(smem);
    float* sW = sH0 + 2 * buf_floats;
    float* sB = sW + plan.wtot;
    // barriers: [0,1] full (loader -> MMA), [2,3] empty (commit -> loader), [4,5] mma done (commit -> epilogue), [6,7] ready (epilogue -> MMA)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + ((plan.btot + 3) & ~3));
    LayerDesc* sL = reinterpret_cast<LayerDesc*>(bars + 8);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sL + B200GYM_MLP_MAX_LAYERS);

    for (int i = tid * 4; i < plan.wtot; i += PIPE_THREADS * 4)
        *reinterpret_cast<float4*>(sW + i) = *reinterpret_cast<const float4*>(wpacked + i);
    for (int i = tid; i < plan.btot; i += PIPE_THREADS) sB[i] = bias[i];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(bars + 0, TM), mbar_init(bars + 1, TM);   // full: every loader thread of the slot arrives
        mbar_init(bars + 2, 1), mbar_init(bars + 3, 1);     // empty: tcgen05.commit after the layer that last reads the X region
        mbar_init(bars + 4, 1), mbar_init(bars + 5, 1);     // MMA completion (tcgen05.commit)
        mbar_init(bars + 6, EPI_THREADS), mbar_init(bars + 7, EPI_THREADS);   // ready: every epilogue thread of the slot arrives
        fence_mbar_init();
    }
    if (tid >= 64 && tid < 64 + L) {
        const int l = tid - 64;
        int woff = 0;
        for (int j = 0; j < l; ++j) woff += p.dims[j] * p.dims[j + 1];
        const int N = p.dims[l + 1];
        LayerDesc d;
        d.da0 = umma_desc(sH0 + (l == 0 ? static_cast<size_t>(plan.x_off) * CHUNK_FLOATS : 0), CHUNK_BYTES, 128);
        d.db0 = umma_desc(sW + woff, N * 16, 128);
        d.idesc = umma_idesc_tf32(N);
        d.ksteps = p.dims[l] / 8;
        d.inc_b = (2 * N * 16) >> 4;   // two 16-byte k-chunks per UMMA K step, in descriptor (16-byte) units
        d.pad = 0;
        sL[l] = d;
    }
    fence_proxy_async();   // the weight tile was written through the generic proxy, the tensor core reads it through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int ntiles = (p.batch + TM - 1) / TM;
    const int slot = warp < 16 ? (warp >> 3) : ((warp >> 2) & 1);   // epilogue warps 0-7 | 8-15, loader warps 16-19 | 20-23
    const int t = tid & (TM - 1);                   // row within the tile (epilogue: TMEM lane = 32 * (warp % 4) + lane)
    float* sH = sH0 + slot * buf_floats;
    uint64_t *full = bars + slot, *empty = bars + 2 + slot, *mma_done = bars + 4 + slot, *ready = bars + 6 + slot;

    unsigned long long* tr = nullptr;
    if (g_mlp_trace && blockIdx.x == 0 && (tid & 31) == 0 && (warp == 0 || warp == 8 || warp == 16 || warp == 20 || warp == 24))
        tr = g_mlp_trace + static_cast<size_t>(warp == 0 ? 0 : warp == 8 ? 1 : warp == 16 ? 2 : warp == 20 ? 3 : 4) * TRACE_EVENTS * 2;
    if (warp == 24) {
        // ------------------------------ MMA warp: serves both slots, layer by layer ------------------------------
        // Each slot's chain is MMA(l) -> epilogue(l) -> MMA(l+1) ...; the warp serves whichever slot has its operands ready.
        // Slot 1 is held back until slot 0 has finished its first (longest) epilogue, so that the two slots run in
        // anti-phase: one slot's epilogue (MUFU/FMA pipes) overlaps the other slot's MMAs instead of its epilogue.
        const uint32_t buf_desc = static_cast<uint32_t>((buf_floats * 4) >> 4);
        const uint32_t inc_a = (2 * CHUNK_BYTES) >> 4;
        const int G = gridDim.x;
        int lay[2] = {0, 0}, tile_of[2] = {static_cast<int>(blockIdx.x), static_cast<int>(blockIdx.x) + G};
        uint32_t nstep[2] = {0, 0}, nfull[2] = {0, 0};
        const uint32_t gate = L < 2 ? 1u : 2u;
        while (tile_of[0] < ntiles || tile_of[1] < ntiles) {
            bool progress = false;
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                if (tile_of[sl] >= ntiles) continue;
                if (sl == 1 && tile_of[0] < ntiles && nstep[0] < gate) continue;
                // operands ready: the previous epilogue of this slot has written h and drained the accumulator ...
                if (nstep[sl] > 0 && !mbar_test(bars + 6 + sl, (nstep[sl] - 1) & 1)) continue;
                // ... and, for the first layer, the loader has delivered the tile
                if (lay[sl] == 0 && !mbar_test(bars + sl, nfull[sl] & 1)) continue;
                const int l = lay[sl];
                const LayerDesc d = sL[l];
                tc_fence_after();
                trace_ev(tr, 10 * l + sl);
                if (elect_one()) {
                    uint64_t da = d.da0 + (sl ? buf_desc : 0u), db = d.db0;
                    const uint32_t tacc = tmem + static_cast<uint32_t>(sl * 128);
                    for (uint32_t ks = 0; ks < d.ksteps; ++ks) {
                        umma_tf32(tacc, da, db, d.idesc, ks > 0 ? 1u : 0u);
                        da += inc_a;
                        db += d.inc_b;
                    }
                    umma_commit(bars + 4 + sl);
                    if (l == plan.free_layer) umma_commit(bars + 2 + sl);   // X region free once these MMAs have retired
                }
                __syncwarp();
                trace_ev(tr, 10 * l + sl + 5);
                progress = true;
                ++nstep[sl];
                if (++lay[sl] == L) {
                    lay[sl] = 0;
                    ++nfull[sl];
                    tile_of[sl] += 2 * G;
                }
            }
            if (!progress) __nanosleep(20);
        }
    } else if (warp >= 16) {
        // ------------------------------ loader warpgroup of this slot ------------------------------
        // 16-byte piece q of the tile (row-major, in_c pieces per row) is loaded by thread q % 128: a warp reads 512 contiguous
        // bytes per instruction; piece (r, c) lands at chunk c, row r of the X region (the padded chunk stride keeps the
        // scattered 128-bit stores of a quarter-warp in different banks)
        const int K0c = p.dims[0] / 4, in_c = p.in_dim / 4;
        const int pieces = TM * in_c;
        constexpr int PRE = 12;   // 16-byte pieces per thread prefetched into registers (48 floats: the flat observation width)
        float* xreg = sH + static_cast<size_t>(plan.x_off) * CHUNK_FLOATS;
        const int r_first = t / in_c, c_first = t - r_first * in_c, dr = TM / in_c, dc = TM - dr * in_c;
        uint32_t ph = 1;   // the first wait on a fresh "empty" barrier falls through
        for (int tile = blockIdx.x + slot * gridDim.x; tile < ntiles; tile += 2 * gridDim.x) {
            const int row0 = tile * TM;
            float4 v[PRE];
            int r = r_first, c = c_first;   // (row, piece) of q = t + TM * j, advanced without divisions
#pragma unroll
            for (int j = 0; j < PRE; ++j) {
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t + TM * j < pieces && row0 + r < p.batch)
                    v[j] = ldg_stream4(reinterpret_cast<const float4*>(x + static_cast<size_t>(row0 + r) * p.in_stride) + c);
                r += dr, c += dc;
                if (c >= in_c) c -= in_c, ++r;
            }
            trace_ev(tr, 100);
            mbar_wait_backoff(empty, ph);
            ph ^= 1;
            trace_ev(tr, 101);
            r = r_first, c = c_first;
#pragma unroll
            for (int j = 0; j < PRE; ++j) {
                if (t + TM * j < pieces) *reinterpret_cast<float4*>(xreg + static_cast<size_t>(c) * CHUNK_FLOATS + r * 4) = v[j];
                r += dr, c += dc;
                if (c >= in_c) c -= in_c, ++r;
            }
            for (int q = t + TM * PRE; q < pieces; q += TM) {   // inputs wider than 48 floats: the rest goes straight through
                const int r = q / in_c, c = q - r * in_c;
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row0 + r < p.batch) w = ldg_stream4(reinterpret_cast<const float4*>(x + static_cast<size_t>(row0 + r) * p.in_stride) + c);
                *reinterpret_cast<float4*>(xreg + static_cast<size_t>(c) * CHUNK_FLOATS + r * 4) = w;
            }
            for (int c = in_c; c < K0c; ++c)   // zero padding of the k dimension
                *reinterpret_cast<float4*>(xreg + static_cast<size_t>(c) * CHUNK_FLOATS + t * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            fence_proxy_async();
            mbar_arrive(full);
            trace_ev(tr, 102);
        }
    } else {
        // ------------------------------ epilogue warpgroup of this slot ------------------------------
        const uint32_t my_taddr = tmem + static_cast<uint32_t>(slot * 128) + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        float* hrow = sH + t * 4;
        uint32_t ph_mma = 0;
        for (int tile = blockIdx.x + slot * gridDim.x; tile < ntiles; tile += 2 * gridDim.x) {
            const int row = tile * TM + t;
            const bool live = row < p.batch;
            float* orow = out + static_cast<size_t>(row) * p.out_dim;
            int boff = 0;
            for (int l = 0; l < L; ++l) {
                const int N = p.dims[l + 1];
                trace_ev(tr, 10 * l + 0);
                mbar_wait(mma_done, ph_mma);
                ph_mma ^= 1;
                tc_fence_after();
                trace_ev(tr, 10 * l + 1);
                const float* sBl = sB + boff;
                // two warps share a TMEM lane quadrant: warp-group half h takes columns [h*N/2, (h+1)*N/2) (all 16 when N == 16)
                const int half = (warp >> 2) & 1;
                const int c_lo = N >= 32 ? half * (N / 2) : 0, c_hi = N >= 32 ? c_lo + N / 2 : (half == 0 ? N : 0);
                uint32_t ra[16], rb[16];
                if (c_lo < c_hi) tmem_ld16_issue(my_taddr + c_lo, ra);
                if (l == L - 1) {
                    for (int n0 = c_lo; n0 < c_hi; n0 += 16) {
                        tmem_ld16_wait(ra);
                        pipe_epilogue_chunk<true>(ra, sBl, n0, hrow, orow, p.out_dim, live);
                        if (n0 + 16 < c_hi) tmem_ld16_issue(my_taddr + n0 + 16, ra);
                    }
                } else {
                    for (int n0 = c_lo; n0 < c_hi; n0 += 32) {
                        tmem_ld16_wait(ra);
                        if (n0 + 16 < c_hi) tmem_ld16_issue(my_taddr + n0 + 16, rb);
                        pipe_epilogue_chunk<false>(ra, sBl, n0, hrow, orow, p.out_dim, live);
                        if (n0 + 16 < c_hi) {
                            tmem_ld16_wait(rb);
                            if (n0 + 32 < c_hi) tmem_ld16_issue(my_taddr + n0 + 32, ra);
                            pipe_epilogue_chunk<false>(rb, sBl, n0 + 16, hrow, orow, p.out_dim, live);
                        }
                    }
                }
                boff += N;
                // this thread's part of the next A operand is written and its accumulator reads are done
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(ready);
                trace_ev(tr, 10 * l + 2);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// mlp_forward_h4_kernel — FOUR tile slots in flight with fp16 operands (tcgen05.mma kind::f16, K = 16 per instruction,
// fp32 accumulation in TMEM).  The TF32 two-slot kernel above is bound by the per-layer hand-off latency of its two serial
// chains (profiles/r1_mlp_forward_ncu.txt); fp16 operands halve the operand buffers (4 slots x 33 KB + 34 KB of weights)
// and the number of dependent MMAs per layer, so four independent chains keep the tensor pipe, the MUFU pipe and the
// loaders busy at the same time.  fp16 has the same 10-bit mantissa TF32 keeps, so the result quality is unchanged
// (|x| <= 100 after the observation clip, ELU outputs >= -1: no range problem); bias + ELU stay in fp32.
//   warps  0-15: epilogue warpgroup of slot warp/4 (TMEM -> bias + ELU -> fp16 A operand of the next layer)
//   warps 16-19 / 20-23: loader groups; group g feeds slots g and g+2 alternately (fp32 rows -> fp16 chunks, register prefetch)
//   warp  24: MMA issue for whichever slot has its operands ready
// Operand layout (A and B, K-major, no swizzle): 16-byte chunks of 8 consecutive k; chunk c of row r at c*CH16 + r*16.
// ------------------------------------------------------------------------------------------------------------------
constexpr int NSLOT = 4;
constexpr int H4_THREADS = 800;
constexpr int CH16_BYTES = TM * 16 + 16;   // padded chunk stride (bank spreading for the loaders' scattered 16-byte stores)

__device__ __forceinline__ uint32_t umma_idesc_f16(int n) {
    // c_format F32 = 1 [4,6); a_format / b_format F16 = 0 [7,10) / [10,13); K-major both; N >> 3 [17,23); M >> 4 [24,29)
    return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(TM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

struct H4Plan {
    int x_off;        // first chunk (of 8 k) of the X region
    int free_layer;   // the X region is free again once this layer's MMAs have completed
    int buf_chunks;   // chunks per slot buffer
    int wtot, btot;
};

template <bool LAST>
__device__ __forceinline__ void h4_epilogue_chunk(const uint32_t (&r)[16], const float* __restrict__ sBl, int n0, unsigned char* __restrict__ hrow,
                                                  float* __restrict__ orow, int out_dim, bool live) {
    // bias add, the log2(e) scale and the "- 1" of the ELU run two activations per instruction (add/mul.rn.f32x2)
    float2 v[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b4 = *reinterpret_cast<const float4*>(sBl + n0 + 4 * q);   // same address in every lane: broadcast
        v[2 * q] = __fadd2_rn(make_float2(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1])), make_float2(b4.x, b4.y));
        v[2 * q + 1] = __fadd2_rn(make_float2(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])), make_float2(b4.z, b4.w));
    }
    if (!LAST) {
        const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f), mone = make_float2(-1.0f, -1.0f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // elu(t) = max(t, exp(min(t, 0)) - 1)
            const float2 a = __fmul2_rn(make_float2(fminf(v[j].x, 0.0f), fminf(v[j].y, 0.0f)), l2e);
            const float2 e = __fadd2_rn(make_float2(ex2_approx(a.x), ex2_approx(a.y)), mone);
            v[j] = make_float2(fmaxf(v[j].x, e.x), fmaxf(v[j].y, e.y));
        }
#pragma unroll
        for (int q = 0; q < 2; ++q)   // next layer's A operand, k = n: chunk (n0/8 + q), this thread's row
            *reinterpret_cast<uint4*>(hrow + static_cast<size_t>(n0 / 8 + q) * CH16_BYTES) =
                make_uint4(pack_h2(v[4 * q].x, v[4 * q].y), pack_h2(v[4 * q + 1].x, v[4 * q + 1].y), pack_h2(v[4 * q + 2].x, v[4 * q + 2].y),
                           pack_h2(v[4 * q + 3].x, v[4 * q + 3].y));
    } else if (live) {
        if ((out_dim & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (n0 + 4 * q < out_dim)
                    *reinterpret_cast<float4*>(orow + n0 + 4 * q) = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (n0 + 2 * j < out_dim) orow[n0 + 2 * j] = v[j].x;
                if (n0 + 2 * j + 1 < out_dim) orow[n0 + 2 * j + 1] = v[j].y;
            }
        }
    }
}

// One CTA's share of one net's forward: CTA `cta` of the `G` CTAs assigned to this net (the pair kernel below gives each of
// two nets its own contiguous range of the grid).
__device__ __forceinline__ void mlp_forward_h4_body(const B200MlpParams& p, const H4Plan& plan, const float* __restrict__ x,
                                                    const __half* __restrict__ wpacked16, const float* __restrict__ bias,
                                                    float* __restrict__ out, const int cta, const int G) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int L = p.num_layers, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t buf_bytes = static_cast<size_t>(plan.buf_chunks) * CH16_BYTES;
    unsigned char* sH0 = smem;
    __half* sW = reinterpret_cast<__half*>(sH0 + NSLOT * buf_bytes);
    float* sB = reinterpret_cast<float*>(sW + plan.wtot);
    // barriers: [0,4) full (loader -> MMA), [4,8) empty (commit -> loader), [8,12) mma done (commit -> epilogue), [12,16) ready
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + ((plan.btot + 3) & ~3));
    LayerDesc* sL = reinterpret_cast<LayerDesc*>(bars + 4 * NSLOT);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sL + B200GYM_MLP_MAX_LAYERS);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int sl = 0; sl < NSLOT; ++sl) {
            mbar_init(bars + sl, TM);                  // full: every loader thread of the group arrives
            mbar_init(bars + NSLOT + sl, 1);           // empty: tcgen05.commit after the layer that last reads the X region
            mbar_init(bars + 2 * NSLOT + sl, 1);       // MMA completion (tcgen05.commit)
            // ready: every epilogue thread of the slot arrives (single-tile CTAs: all four warpgroups work on slot 0)
            mbar_init(bars + 3 * NSLOT + sl, ((p.batch + TM - 1) / TM <= G && sl == 0) ? NSLOT * TM : TM);
        }
        fence_mbar_init();
    }
    // programmatic dependent launch: TMEM allocation and barrier set-up above overlap the previous kernel's tail; weights, biases
    // and observations are other kernels' output, so nothing in global memory is touched before this point
    pdl_launch_dependents();
    pdl_wait();
    for (int i = tid * 8; i < plan.wtot; i += H4_THREADS * 8)
        *reinterpret_cast<uint4*>(sW + i) = *reinterpret_cast<const uint4*>(wpacked16 + i);
    for (int i = tid; i < plan.btot; i += H4_THREADS) sB[i] = bias[i];
    if (tid >= 64 && tid < 64 + L) {
        const int l = tid - 64;
        int woff = 0;
        for (int j = 0; j < l; ++j) woff += p.dims[j] * p.dims[j + 1];
        const int N = p.dims[l + 1];
        LayerDesc d;
        d.da0 = umma_desc(sH0 + (l == 0 ? static_cast<size_t>(plan.x_off) * CH16_BYTES : 0), CH16_BYTES, 128);
        d.db0 = umma_desc(sW + woff, N * 16, 128);
        d.idesc = umma_idesc_f16(N);
        d.ksteps = p.dims[l] / 16;
        d.inc_b = (2 * N * 16) >> 4;   // two 16-byte k-chunks per UMMA K step, in descriptor (16-byte) units
        d.pad = 0;
        sL[l] = d;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int ntiles = (p.batch + TM - 1) / TM;
    const int t = tid & (TM - 1);   // row within the tile (epilogue: TMEM lane = 32 * (warp % 4) + lane)

    if (warp == 24) {
        // ------------------------------ MMA warp: serves whichever slot has its operands ready ------------------------------
        const uint32_t buf_desc = static_cast<uint32_t>(buf_bytes >> 4);
        const uint32_t inc_a = (2 * CH16_BYTES) >> 4;
        int lay[NSLOT], tile_of[NSLOT];
        uint32_t nstep[NSLOT], nfull[NSLOT];
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl) lay[sl] = 0, tile_of[sl] = cta + sl * G, nstep[sl] = 0, nfull[sl] = 0;
        while (true) {
            bool any = false, progress = false;
#pragma unroll
            for (int sl = 0; sl < NSLOT; ++sl) {
                if (tile_of[sl] >= ntiles) continue;
                any = true;
                if (nstep[sl] > 0 && !mbar_test(bars + 3 * NSLOT + sl, (nstep[sl] - 1) & 1)) continue;   // previous epilogue of the slot done
                if (lay[sl] == 0 && !mbar_test(bars + sl, nfull[sl] & 1)) continue;                      // input tile delivered
                const int l = lay[sl];
                const LayerDesc d = sL[l];
                tc_fence_after();
                if (elect_one()) {
                    uint64_t da = d.da0 + sl * buf_desc, db = d.db0;
                    const uint32_t tacc = tmem + static_cast<uint32_t>(sl * 128);
                    for (uint32_t ks = 0; ks < d.ksteps; ++ks) {
                        umma_f16(tacc, da, db, d.idesc, ks > 0 ? 1u : 0u);
                        da += inc_a;
                        db += d.inc_b;
                    }
                    umma_commit(bars + 2 * NSLOT + sl);
                    if (l == plan.free_layer) umma_commit(bars + NSLOT + sl);   // X region free once these MMAs have retired
                }
                __syncwarp();
                progress = true;
                ++nstep[sl];
                if (++lay[sl] == L) {
                    lay[sl] = 0;
                    ++nfull[sl];
                    tile_of[sl] += NSLOT * G;
                }
            }
            if (!any) break;
            if (!progress) __nanosleep(20);
        }
    } else if (warp >= 16) {
        // ------------------------------ loader group g: slots g and g + 2 alternately ------------------------------
        // 32-byte piece q of the tile (8 consecutive floats of a row) is handled by thread q % 128: a warp reads 1 KB
        // contiguous per step and stores 16-byte fp16 chunks; piece (r, c) lands at chunk c, row r of the X region.
        const int g = (warp - 16) >> 2;
        const int K0c = p.dims[0] / 8, in_c = p.in_dim / 8;
        const int pieces = TM * in_c;
        constexpr int PRE = 6;   // pieces per thread prefetched into registers (48 floats: the flat observation width)
        const int r_first = t / in_c, c_first = t - r_first * in_c, dr = TM / in_c, dc = TM - dr * in_c;
        uint32_t ph[2] = {1, 1};   // the first wait on a fresh "empty" barrier falls through
        for (int it = 0;; ++it) {
            const int sl = g + 2 * (it & 1);
            const int tile = cta + sl * G + (it >> 1) * NSLOT * G;
            // tiles of the two slots of this group interleave in increasing order; stop when both are exhausted
            if (tile >= ntiles) {
                const int other = cta + (g + 2 * ((it + 1) & 1)) * G + ((it + 1) >> 1) * NSLOT * G;
                if (other >= ntiles) break;
                continue;
            }
            const int row0 = tile * TM;
            unsigned char* xreg = sH0 + sl * buf_bytes + static_cast<size_t>(plan.x_off) * CH16_BYTES;
            float4 v[2 * PRE];
            int r = r_first, c = c_first;
#pragma unroll
            for (int j = 0; j < PRE; ++j) {
                v[2 * j] = v[2 * j + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t + TM * j < pieces && row0 + r < p.batch) {
                    const float4* src = reinterpret_cast<const float4*>(x + static_cast<size_t>(row0 + r) * p.in_stride) + 2 * c;
                    v[2 * j] = ldg_stream4(src);
                    v[2 * j + 1] = ldg_stream4(src + 1);
                }
                r += dr, c += dc;
                if (c >= in_c) c -= in_c, ++r;
            }
            mbar_wait_backoff(bars + NSLOT + sl, ph[it & 1]);
            ph[it & 1] ^= 1;
            r = r_first, c = c_first;
#pragma unroll
            for (int j = 0; j < PRE; ++j) {
                if (t + TM * j < pieces)
                    *reinterpret_cast<uint4*>(xreg + static_cast<size_t>(c) * CH16_BYTES + r * 16) =
                        make_uint4(pack_h2(v[2 * j].x, v[2 * j].y), pack_h2(v[2 * j].z, v[2 * j].w), pack_h2(v[2 * j + 1].x, v[2 * j + 1].y),
                                   pack_h2(v[2 * j + 1].z, v[2 * j + 1].w));
                r += dr, c += dc;
                if (c >= in_c) c -= in_c, ++r;
            }
            for (int q = t + TM * PRE; q < pieces; q += TM) {   // inputs wider than 48 floats: the rest goes straight through
                const int rr = q / in_c, cc = q - rr * in_c;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                if (row0 + rr < p.batch) {
                    const float4* src = reinterpret_cast<const float4*>(x + static_cast<size_t>(row0 + rr) * p.in_stride) + 2 * cc;
                    a = ldg_stream4(src), b = ldg_stream4(src + 1);
                }
                *reinterpret_cast<uint4*>(xreg + static_cast<size_t>(cc) * CH16_BYTES + rr * 16) =
                    make_uint4(pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(b.x, b.y), pack_h2(b.z, b.w));
            }
            for (int cz = in_c; cz < K0c; ++cz)   // zero padding of the k dimension
                *reinterpret_cast<uint4*>(xreg + static_cast<size_t>(cz) * CH16_BYTES + t * 16) = make_uint4(0u, 0u, 0u, 0u);
            fence_proxy_async();
            mbar_arrive(bars + sl);
        }
    } else {
        // ------------------------------ epilogue warpgroup of slot warp / 4 ------------------------------
        if (ntiles <= G) {
            // Small batches (the rollout's 4096 envs = 32 tiles): a CTA holds at most ONE tile and three of its four epilogue warpgroups
            // would idle while the layers run back to back, each waiting for a 128-thread epilogue.  Here the four warpgroups share the
            // tile instead: warpgroup w takes the 16-column blocks w, w + 4, ... of every layer's accumulator (all four cover the 128
            // TMEM lanes), so the per-layer epilogue is a quarter as long.  Same arithmetic per element: results bit-identical.
            const int wg = warp >> 2;
            const uint32_t my_taddr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
            unsigned char* hrow = sH0 + t * 16;
            uint64_t *mma_done = bars + 2 * NSLOT, *ready = bars + 3 * NSLOT;
            if (cta < ntiles) {
                const int row = cta * TM + t;
                const bool live = row < p.batch;
                float* orow = out + static_cast<size_t>(row) * p.out_dim;
                int boff = 0;
                for (int l = 0; l < L; ++l) {
                    const int N = p.dims[l + 1];
                    mbar_wait(mma_done, l & 1);
                    tc_fence_after();
                    const float* sBl = sB + boff;
                    for (int n0 = 16 * wg; n0 < N; n0 += 16 * NSLOT) {
                        uint32_t ra[16];
                        tmem_ld16_issue(my_taddr + n0, ra);
                        tmem_ld16_wait(ra);
                        if (l == L - 1) h4_epilogue_chunk<true>(ra, sBl, n0, hrow, orow, p.out_dim, live);
                        else h4_epilogue_chunk<false>(ra, sBl, n0, hrow, orow, p.out_dim, live);
                    }
                    boff += N;
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive(ready);
                }
            }
        } else {
        const int slot = warp >> 2;
        const uint32_t my_taddr = tmem + static_cast<uint32_t>(slot * 128) + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        unsigned char* hrow = sH0 + slot * buf_bytes + t * 16;
        uint64_t *mma_done = bars + 2 * NSLOT + slot, *ready = bars + 3 * NSLOT + slot;
        uint32_t ph_mma = 0;
        for (int tile = cta + slot * G; tile < ntiles; tile += NSLOT * G) {
            const int row = tile * TM + t;
            const bool live = row < p.batch;
            float* orow = out + static_cast<size_t>(row) * p.out_dim;
            int boff = 0;
            for (int l = 0; l < L; ++l) {
                const int N = p.dims[l + 1];
                mbar_wait(mma_done, ph_mma);
                ph_mma ^= 1;
                tc_fence_after();
                const float* sBl = sB + boff;
                uint32_t ra[16], rb[16];
                tmem_ld16_issue(my_taddr, ra);
                if (l == L - 1) {
                    for (int n0 = 0; n0 < N; n0 += 16) {
                        tmem_ld16_wait(ra);
                        h4_epilogue_chunk<true>(ra, sBl, n0, hrow, orow, p.out_dim, live);
                        if (n0 + 16 < N) tmem_ld16_issue(my_taddr + n0 + 16, ra);
                    }
                } else {
                    for (int n0 = 0; n0 < N; n0 += 32) {
                        tmem_ld16_wait(ra);
                        if (n0 + 16 < N) tmem_ld16_issue(my_taddr + n0 + 16, rb);
                        h4_epilogue_chunk<false>(ra, sBl, n0, hrow, orow, p.out_dim, live);
                        if (n0 + 16 < N) {
                            tmem_ld16_wait(rb);
                            if (n0 + 32 < N) tmem_ld16_issue(my_taddr + n0 + 32, ra);
                            h4_epilogue_chunk<false>(rb, sBl, n0 + 16, hrow, orow, p.out_dim, live);
                        }
                    }
                }
                boff += N;
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(ready);
            }
        }
        }
    }
    (void)lane;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

__global__ void __launch_bounds__(H4_THREADS, 1) mlp_forward_h4_kernel(const __grid_constant__ B200MlpParams p, const H4Plan plan,
                                                                        const float* __restrict__ x, const __half* __restrict__ wpacked16,
                                                                        const float* __restrict__ bias, float* __restrict__ out) {
    mlp_forward_h4_body(p, plan, x, wpacked16, bias, out, static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x));
}

// Two independent nets (PPO.act's actor and critic) in ONE launch: CTAs [0, split) run net 0, the rest net 1.  The choice is
// CTA-uniform, so each CTA is exactly the single-net kernel on its own range of the grid.
struct H4Net {
    B200MlpParams p;
    H4Plan plan;
    const float* x;
    const __half* w16;
    const float* bias;
    float* out;
};
struct H4Pair {
    H4Net net[2];
    int split;
};

__global__ void __launch_bounds__(H4_THREADS, 1) mlp_forward_h4_pair_kernel(const __grid_constant__ H4Pair pr) {
    const int b = static_cast<int>(blockIdx.x);
    const bool second = b >= pr.split;
    const H4Net& n = pr.net[second ? 1 : 0];
    mlp_forward_h4_body(n.p, n.plan, n.x, n.w16, n.bias, n.out, second ? b - pr.split : b,
                        second ? static_cast<int>(gridDim.x) - pr.split : pr.split);
}

// Argument checks shared by the forward entries; fills the totals the plan needs.
int mlp_check_params(const B200MlpParams* p, const char* who, size_t* wtot_out, size_t* btot_out) {
    B200_REQUIRE(p->batch > 0 && p->num_layers >= 1 && p->num_layers <= B200GYM_MLP_MAX_LAYERS, B200GYM_EINVAL,
                 "%s: batch > 0 and 1..%d layers", who, B200GYM_MLP_MAX_LAYERS);
    size_t wtot = 0, btot = 0;
    for (int l = 0; l < p->num_layers; ++l) {
        const int K = p->dims[l], N = p->dims[l + 1];
        B200_REQUIRE(K >= 8 && K % 8 == 0 && N >= 16 && N % 16 == 0 && N <= 256, B200GYM_EINVAL,
                     "%s: layer %d is %d -> %d; need K %% 8 == 0, N %% 16 == 0, N <= 256 (pad with zeros)", who, l, K, N);
        wtot += static_cast<size_t>(K) * N;
        btot += N;
    }
    B200_REQUIRE(p->in_dim > 0 && p->in_dim <= p->dims[0] && p->in_stride >= p->in_dim && p->out_dim > 0 &&
                     p->out_dim <= p->dims[p->num_layers],
                 B200GYM_EINVAL, "%s: inconsistent in_dim / out_dim", who);
    *wtot_out = wtot, *btot_out = btot;
    return B200GYM_OK;
}

// Shared-memory plan of the fp16 four-slot kernel; false when the net is outside what that kernel takes.
bool h4_make_plan(const B200MlpParams* p, size_t wtot, size_t btot, H4Plan* plan, size_t* smem) {
    const int L = p->num_layers;
    int nmax = 0, hidden_late = 0, hidden_all = 0;
    bool k16 = true;
    for (int l = 0; l < L; ++l) {
        nmax = p->dims[l + 1] > nmax ? p->dims[l + 1] : nmax;
        k16 = k16 && (p->dims[l] % 16 == 0);
    }
    for (int j = 1; j <= L - 1; ++j) hidden_all = p->dims[j] > hidden_all ? p->dims[j] : hidden_all;
    for (int j = 2; j <= L - 1; ++j) hidden_late = p->dims[j] > hidden_late ? p->dims[j] : hidden_late;
    plan->x_off = hidden_late / 8;
    plan->free_layer = L - 1 < 1 ? L - 1 : 1;
    plan->buf_chunks = hidden_all / 8 > plan->x_off + p->dims[0] / 8 ? hidden_all / 8 : plan->x_off + p->dims[0] / 8;
    plan->wtot = static_cast<int>(wtot), plan->btot = static_cast<int>(btot);
    *smem = NSLOT * static_cast<size_t>(plan->buf_chunks) * CH16_BYTES + wtot * 2 + ((btot + 3) & ~size_t(3)) * 4 + 8 * 4 * NSLOT +
            sizeof(LayerDesc) * B200GYM_MLP_MAX_LAYERS + 64;
    return k16 && nmax <= 128 && (p->in_dim & 7) == 0 && (p->in_stride & 3) == 0 && (wtot & 7) == 0 && *smem <= 227 * 1024;
}

int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

}  // namespace

/* debug: registers (or clears, with NULL) a device buffer of 5 * 4096 * 2 uint64 that CTA 0 of the pipelined kernel fills
 * with (event code, clock64) pairs — tools/trace_mlp.py turns it into a per-phase cycle budget. */
extern "C" int b200gym_debug_mlp_trace(void* buf) {
    unsigned long long* p = static_cast<unsigned long long*>(buf);
    cudaError_t e = cudaMemcpyToSymbol(g_mlp_trace, &p, sizeof(p));
    B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "debug_mlp_trace: %s", cudaGetErrorString(e));
    return B200GYM_OK;
}

extern "C" int b200gym_mlp_forward(const B200MlpParams* p, const float* x, const float* wpacked, const float* bias, float* out,
                                   void* stream) {
    B200_REQUIRE(p && x && wpacked && bias && out, B200GYM_EINVAL, "mlp_forward: null argument");
    B200_REQUIRE(b200_aligned16(x) && b200_aligned16(wpacked), B200GYM_EALIGN, "mlp_forward: x / weights must be 16-byte aligned");
    size_t kmax = 0, wtot = 0, btot = 0;
    if (int rc = mlp_check_params(p, "mlp_forward", &wtot, &btot)) return rc;
    for (int l = 0; l < p->num_layers; ++l) kmax = p->dims[l] > (int)kmax ? p->dims[l] : kmax;
    const int sms = sm_count();
    const int ntiles = (p->batch + TM - 1) / TM;
    const int L = p->num_layers;
    static int variant = -1;   // 0: best available, 1: force the TF32 two-slot kernel, 2: force the serial kernel (A/B runs)
    if (variant < 0) {
        const char* e = getenv("B200GYM_MLP_VARIANT");
        variant = e ? atoi(e) : 0;
    }
    // ---- fp16 four-slot kernel (weights: the fp16 section that follows the fp32 section of `wpacked`) --------------------
    {
        H4Plan plan;
        size_t smem_h4 = 0;
        if (h4_make_plan(p, wtot, btot, &plan, &smem_h4) && variant == 0) {
            static size_t configured_h4 = 0;
            if (smem_h4 > configured_h4) {
                cudaError_t e = cudaFuncSetAttribute(mlp_forward_h4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_h4));
                B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "mlp_forward: cannot reserve %zu B of shared memory: %s", smem_h4, cudaGetErrorString(e));
                configured_h4 = smem_h4;
            }
            b200_launch_pdl(0, mlp_forward_h4_kernel, dim3(ntiles < sms ? ntiles : sms), dim3(H4_THREADS), smem_h4, static_cast<cudaStream_t>(stream),
                            *p, plan, x, reinterpret_cast<const __half*>(wpacked + wtot), bias, out);
            B200_LAUNCH_CHECK("mlp_forward (fp16, 4 slots)");
            return B200GYM_OK;
        }
    }
    // ---- TF32 two-slot kernel --------------------------------------------------------------------------------------------
    {
        PipePlan plan;
        int nmax = 0, hidden_late = 0, hidden_all = 0;
        for (int l = 0; l < L; ++l) nmax = p->dims[l + 1] > nmax ? p->dims[l + 1] : nmax;
        for (int j = 1; j <= L - 1; ++j) hidden_all = p->dims[j] > hidden_all ? p->dims[j] : hidden_all;
        for (int j = 2; j <= L - 1; ++j) hidden_late = p->dims[j] > hidden_late ? p->dims[j] : hidden_late;
        plan.x_off = hidden_late / 4;
        plan.free_layer = L - 1 < 1 ? L - 1 : 1;
        plan.buf_chunks = hidden_all / 4 > plan.x_off + p->dims[0] / 4 ? hidden_all / 4 : plan.x_off + p->dims[0] / 4;
        plan.wtot = static_cast<int>(wtot), plan.btot = static_cast<int>(btot);
        const size_t smem_pipe = (2 * static_cast<size_t>(plan.buf_chunks) * CHUNK_FLOATS + wtot + ((btot + 3) & ~size_t(3))) * 4 + 64 +
                                 sizeof(LayerDesc) * B200GYM_MLP_MAX_LAYERS + 16;
        static int force_serial = -1;
        if (force_serial < 0) {
            const char* e = getenv("B200GYM_MLP_SERIAL");
            force_serial = (e && e[0] == '1') ? 1 : 0;
        }
        if (!force_serial && variant != 2 && nmax <= 128 && (p->in_dim & 3) == 0 && (p->in_stride & 3) == 0 && smem_pipe <= 227 * 1024) {
            static size_t configured_pipe = 0;
            if (smem_pipe > configured_pipe) {
                cudaError_t e = cudaFuncSetAttribute(mlp_forward_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_pipe));
                B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "mlp_forward: cannot reserve %zu B of shared memory: %s", smem_pipe, cudaGetErrorString(e));
                configured_pipe = smem_pipe;
            }
            mlp_forward_pipe_kernel<<<ntiles < sms ? ntiles : sms, PIPE_THREADS, smem_pipe, static_cast<cudaStream_t>(stream)>>>(*p, plan, x, wpacked,
                                                                                                                            bias, out);
            B200_LAUNCH_CHECK("mlp_forward (pipelined)");
            return B200GYM_OK;
        }
    }
    // ---- generic serial kernel (odd input widths, N up to 256) ---------------------------------------------------
    const size_t smem = (TM * kmax + wtot + ((btot + 3) & ~size_t(3))) * 4 + 16;
    B200_REQUIRE(smem <= 227 * 1024, B200GYM_EINVAL,
                 "mlp_forward: %zu B of shared memory needed (weights must stay resident); this net is too large for the fused kernel", smem);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(mlp_forward_serial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "mlp_forward: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
        configured = smem;
    }
    mlp_forward_serial_kernel<<<ntiles < sms ? ntiles : sms, TM, smem, static_cast<cudaStream_t>(stream)>>>(*p, x, wpacked, bias, out);
    B200_LAUNCH_CHECK("mlp_forward");
    return B200GYM_OK;
}

/* Two nets in one launch (PPO.act: actor on obs, critic on critic obs).  Both must be nets the fp16 four-slot kernel takes
 * (b200gym_mlp_forward's first variant); otherwise B200GYM_EINVAL and the caller issues two b200gym_mlp_forward calls. */
extern "C" int b200gym_mlp_forward_pair(const B200MlpParams* pa, const float* xa, const float* wa, const float* ba, float* outa,
                                        const B200MlpParams* pb, const float* xb, const float* wb, const float* bb, float* outb,
                                        void* stream) {
    B200_REQUIRE(pa && xa && wa && ba && outa && pb && xb && wb && bb && outb, B200GYM_EINVAL, "mlp_forward_pair: null argument");
    B200_REQUIRE(b200_aligned16(xa) && b200_aligned16(wa) && b200_aligned16(xb) && b200_aligned16(wb), B200GYM_EALIGN,
                 "mlp_forward_pair: x / weights must be 16-byte aligned");
    H4Pair pr;
    size_t smem = 0;
    const B200MlpParams* ps[2] = {pa, pb};
    const float* xs[2] = {xa, xb};
    const float* ws[2] = {wa, wb};
    const float* bs[2] = {ba, bb};
    float* os[2] = {outa, outb};
    int tiles[2];
    for (int i = 0; i < 2; ++i) {
        size_t wtot = 0, btot = 0, sm = 0;
        if (int rc = mlp_check_params(ps[i], "mlp_forward_pair", &wtot, &btot)) return rc;
        B200_REQUIRE(h4_make_plan(ps[i], wtot, btot, &pr.net[i].plan, &sm), B200GYM_EINVAL,
                     "mlp_forward_pair: net %d is outside the fp16 four-slot kernel (K %% 16, N <= 128, in_dim %% 8, weights resident)", i);
        pr.net[i].p = *ps[i];
        pr.net[i].x = xs[i], pr.net[i].w16 = reinterpret_cast<const __half*>(ws[i] + wtot), pr.net[i].bias = bs[i], pr.net[i].out = os[i];
        smem = sm > smem ? sm : smem;
        tiles[i] = (ps[i]->batch + TM - 1) / TM;
    }
    // one CTA per tile while the grid fits one wave; beyond that the SMs are split in proportion to the tile counts
    const int sms = sm_count();
    int ga = tiles[0], gb = tiles[1];
    if (ga + gb > sms) {
        ga = static_cast<int>(static_cast<long long>(sms) * tiles[0] / (tiles[0] + tiles[1]));
        ga = ga < 1 ? 1 : (ga > sms - 1 ? sms - 1 : ga);
        gb = sms - ga;
        ga = ga < tiles[0] ? ga : tiles[0], gb = gb < tiles[1] ? gb : tiles[1];
    }
    pr.split = ga;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(mlp_forward_h4_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "mlp_forward_pair: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
        configured = smem;
    }
    b200_launch_pdl(0, mlp_forward_h4_pair_kernel, dim3(ga + gb), dim3(H4_THREADS), smem, static_cast<cudaStream_t>(stream), pr);
    B200_LAUNCH_CHECK("mlp_forward_pair");
    return B200GYM_OK;
}
