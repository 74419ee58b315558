// gemm_f16_kernel — the contractions of the PPO update (rsl_rl PPO.update: forward of ActorCritic.actor / .critic,
// loss.backward() = input gradients + weight gradients) on the 5th-generation tensor cores.
// One grouped kernel, three modes (include/b200gym.h): FWD (activations x W^T, bias + ELU epilogue), DGRAD (dZ x W, ELU'
// epilogue) and WGRAD (dZ^T x H over the batch rows, split-K, fp32 red.add into the flat gradient buffer, bias gradient from
// a constant ones-column appended to the H tile).  tcgen05.mma kind::f16 (M = 128, N <= 144, K = 16), fp32 accumulators in
// TMEM, operands staged by cp.async into the chunk layout of tc.cuh and consumed K-major or MN-major by descriptor only —
// no transposed copies of activations, gradients or weights exist anywhere.
//
// CTA = 288 threads: warps 0-3 epilogue (TMEM lane = tile row), warps 4-7 loaders (cp.async, 3-stage ring, per-stage
// full/empty mbarriers), warp 8 issues the MMAs.  One output tile per CTA, two CTAs per SM (106 KB of shared memory, 256 TMEM
// columns each) so that one CTA's epilogue overlaps the other's main loop.
#include <stdlib.h>
#include "tc.cuh"
#include "../../include/b200gym.h"

namespace {

constexpr int TM = 128;               // UMMA M
constexpr int BN = 128;               // N tile
constexpr int KC = 64;                // K extent of one pipeline stage
constexpr int NSTAGE = 3;
constexpr int CH_ROWS128 = TM * 16 + 16;   // chunk stride of a 128-row tile (K-major operands)
constexpr int CH_ROWSKC = KC * 16 + 16;    // chunk stride of a KC-row tile (MN-major operands: rows = K)
constexpr int A_STAGE = 16 * CH_ROWSKC;            // max(8 * CH_ROWS128, 16 * CH_ROWSKC) = 16640
constexpr int B_STAGE = (BN / 8 + 2) * CH_ROWSKC;  // 18 chunks (BN + the 16-column ones/zero block of WGRAD) = 18720 >= 8 * CH_ROWS128
constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
constexpr int GEMM_THREADS = 288;
constexpr uint32_t TMEM_COLS = 256;
constexpr size_t GEMM_SMEM = static_cast<size_t>(NSTAGE) * STAGE_BYTES + 128 + BN * sizeof(float);   // stages, barriers, the tile's bias row
static_assert(8 * CH_ROWS128 <= A_STAGE && 8 * CH_ROWS128 <= B_STAGE, "stage regions too small");

struct GemmBatch {
    B200GemmProblem p[B200GYM_GEMM_MAX_PROBLEMS];
    int cta_end[B200GYM_GEMM_MAX_PROBLEMS];
    int n;
};

// cooperative cp.async of a [nrows x nchunks*8] tile of a row-major fp16 matrix into the chunk layout; out-of-range pieces are zero-filled
__device__ __forceinline__ void load_tile(unsigned char* dst, int chunk_stride, const __half* base, int ld, int row0, int col0, int nrows,
                                          int nchunks, int row_lim, int col_lim, int t) {
    const int pieces = nrows * nchunks;
    for (int q = t; q < pieces; q += 128) {
        // chunk counts are powers of two for every net in use (8 = one 64-column stage): shift / mask instead of an integer division
        const int r = (nchunks & (nchunks - 1)) == 0 ? q >> (31 - __clz(nchunks)) : q / nchunks, c = q - r * nchunks;
        const int gr = row0 + r, gc = col0 + 8 * c;
        const bool ok = gr < row_lim && gc < col_lim;
        const __half* src = ok ? base + static_cast<size_t>(gr) * ld + gc : base;
        tc::cp_async16(dst + c * chunk_stride + r * 16, src, ok ? 16u : 0u);
    }
}

__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_f16_kernel(const __grid_constant__ GemmBatch batch) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
    uint64_t* empty = full + NSTAGE;
    uint64_t* accum = empty + NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
    float* s_bias = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES + 128);

    // ---- which problem / tile / split is this CTA ----
    int pi = 0;
    while (pi < batch.n - 1 && static_cast<int>(blockIdx.x) >= batch.cta_end[pi]) ++pi;
    const B200GemmProblem& P = batch.p[pi];
    const int local = static_cast<int>(blockIdx.x) - (pi ? batch.cta_end[pi - 1] : 0);
    const int mode = P.mode;
    const int splits = mode == B200GYM_GEMM_WGRAD ? P.splits : 1;
    const int tiles_n = (P.n + BN - 1) / BN;
    const int split = local % splits, tile = local / splits;
    const int tile_m = tile / tiles_n, tile_n = tile - tile_m * tiles_n;
    const int bn_eff = min(BN, P.n - tile_n * BN);   // multiple of 16
    const bool ones_block = mode == B200GYM_GEMM_WGRAD && P.bias != nullptr && tile_n == 0;
    const int n_mma = bn_eff + (ones_block ? 16 : 0);

    int k_begin = 0, k_end = P.k;
    if (mode == B200GYM_GEMM_WGRAD) {
        const int per = (((P.k + splits - 1) / splits) + KC - 1) / KC * KC;
        k_begin = split * per;
        k_end = min(P.k, k_begin + per);
        if (k_begin >= k_end) return;   // uniform for the whole CTA
    }
    const int ktot = (k_end - k_begin + 15) & ~15;
    const int nstages = (ktot + KC - 1) / KC;

    if (warp == 0) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
    if (tid == 32) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(full + s, 128);   // every loader thread arrives once its share of the stage has landed
            mbar_init(empty + s, 1);    // tcgen05.commit after the MMAs that read the stage
        }
        mbar_init(accum, 1);
        fence_mbar_init();
    }
    // WGRAD: the dZ tile is [rows x 128 columns] whatever the layer width; columns beyond the layer's width are zero and never change,
    // so they are written once here instead of being re-issued as zero-fill copies in every stage
    const int a_chunks = mode == B200GYM_GEMM_WGRAD ? min(TM / 8, (P.m - tile_m * TM + 7) / 8) : 0;
    if (mode == B200GYM_GEMM_WGRAD && a_chunks < TM / 8 && warp >= 4 && warp < 8) {
        const int t = tid - 128, nz = TM / 8 - a_chunks;
        for (int q = t; q < NSTAGE * nz * KC; q += 128) {
            const int s = q / (nz * KC), rem = q - s * (nz * KC), c = a_chunks + rem / KC, r = rem % KC;
            *reinterpret_cast<uint4*>(smem + s * STAGE_BYTES + c * CH_ROWSKC + r * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
    }
    if (ones_block && warp >= 4 && warp < 8) {
        // the bias gradient rides the weight-gradient MMA: a constant [1,0,...,0] block of 16 columns behind the H tile
        const int t = tid - 128;
        for (int q = t; q < NSTAGE * KC; q += 128) {
            const int s = q / KC, r = q - s * KC;
            unsigned char* bt = smem + s * STAGE_BYTES + A_STAGE;
            *reinterpret_cast<uint4*>(bt + (bn_eff / 8) * CH_ROWSKC + r * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);   // fp16 1.0, then zeros
            *reinterpret_cast<uint4*>(bt + (bn_eff / 8 + 1) * CH_ROWSKC + r * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = *tmem_slot;
    // programmatic dependent launch: TMEM allocation and barrier set-up above overlap the previous launch's tail; operands and the
    // accumulation target are other kernels' output
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 8) {
        // ------------------------------ MMA issue ------------------------------
        const bool a_mn = mode == B200GYM_GEMM_WGRAD, b_mn = mode != B200GYM_GEMM_FWD;
        const uint32_t idesc = tc::idesc_f16(n_mma, a_mn, b_mn);
        constexpr uint32_t mn_lbo = 128, mn_sbo = CH_ROWSKC;   // MN-major: next 8 k = next 8 rows, next 8 m/n = next piece column
        const uint32_t a_step = a_mn ? (256u >> 4) : ((2u * CH_ROWS128) >> 4);
        const uint32_t b_step = b_mn ? (256u >> 4) : ((2u * CH_ROWS128) >> 4);
        for (int it = 0; it < nstages; ++it) {
            const int s = it % NSTAGE;
            tc::mbar_wait_spin(full + s, (it / NSTAGE) & 1);
            tc::fence_after();
            if (tc::elect_one()) {
                const unsigned char* at = smem + s * STAGE_BYTES;
                const unsigned char* bt = at + A_STAGE;
                uint64_t da = a_mn ? tc::smem_desc(at, mn_lbo, mn_sbo) : tc::smem_desc(at, CH_ROWS128, 128);
                uint64_t db = b_mn ? tc::smem_desc(bt, mn_lbo, mn_sbo) : tc::smem_desc(bt, CH_ROWS128, 128);
                const int nk = (min(KC, ktot - it * KC)) >> 4;
                for (int j = 0; j < nk; ++j) {
                    tc::mma_f16(tmem, da, db, idesc, (it | j) != 0 ? 1u : 0u);
                    da += a_step;
                    db += b_step;
                }
                tc::commit(empty + s);
                if (it == nstages - 1) tc::commit(accum);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ------------------------------ loaders ------------------------------
        const int t = tid - 128;
        const __half* A = static_cast<const __half*>(P.a);
        const __half* B = static_cast<const __half*>(P.b);
        for (int it = 0; it < nstages; ++it) {
            const int s = it % NSTAGE;
            tc::mbar_wait_sleep(empty + s, ((it / NSTAGE) & 1) ^ 1);   // a fresh barrier lets the first NSTAGE waits through
            unsigned char* at = smem + s * STAGE_BYTES;
            unsigned char* bt = at + A_STAGE;
            const int kc_eff = min(KC, ktot - it * KC);
            const int k0 = k_begin + it * KC;
            if (mode == B200GYM_GEMM_WGRAD) {
                load_tile(at, CH_ROWSKC, A, P.lda, k0, tile_m * TM, kc_eff, a_chunks, k_end, P.m, t);
                load_tile(bt, CH_ROWSKC, B, P.ldb, k0, tile_n * BN, kc_eff, bn_eff / 8, k_end, P.n, t);
            } else {
                load_tile(at, CH_ROWS128, A, P.lda, tile_m * TM, k0, TM, kc_eff / 8, P.m, P.k, t);
                if (mode == B200GYM_GEMM_FWD) load_tile(bt, CH_ROWS128, B, P.ldb, tile_n * BN, k0, bn_eff, kc_eff / 8, P.n, P.k, t);
                else load_tile(bt, CH_ROWSKC, B, P.ldb, k0, tile_n * BN, kc_eff, bn_eff / 8, P.k, P.n, t);
            }
            tc::cp_async_commit();
            if (it > 0) {   // the previous stage's copies of this thread have landed: publish them to the tensor-core proxy
                tc::cp_async_wait<1>();
                fence_proxy_async();
                tc::mbar_arrive(full + (it - 1) % NSTAGE);
            }
        }
        tc::cp_async_wait<0>();
        fence_proxy_async();
        tc::mbar_arrive(full + (nstages - 1) % NSTAGE);
    } else {
        // ------------------------------ epilogue: TMEM lane = tile row ------------------------------
        // (two co-resident CTAs of 106 KB leave no L1: a global load in this role is an L2 round trip on the critical path, so
        //  what the epilogue needs from global memory is fetched while the main loop runs)
        const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        const int grow = tile_m * TM + tid;
        const int gcol0 = tile_n * BN;
        uint4 aux_pre[8];   // DGRAD: the first 64 columns of this row's H (ELU derivative), in flight during the main loop
        if (mode == B200GYM_GEMM_FWD) {
            s_bias[tid] = (P.bias != nullptr && gcol0 + tid < P.n_real) ? __ldg(P.bias + gcol0 + tid) : 0.0f;
            asm volatile("bar.sync 1, 128;" ::: "memory");
        } else if (mode == B200GYM_GEMM_DGRAD) {
            const __half* aux = static_cast<const __half*>(P.aux);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                aux_pre[q] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);   // h = 1 -> derivative 1
                if (aux != nullptr && grow < P.m && 8 * q < bn_eff)
                    aux_pre[q] = __ldg(reinterpret_cast<const uint4*>(aux + static_cast<size_t>(grow) * P.ldaux + gcol0) + q);
            }
        }
        tc::mbar_wait_sleep(accum, 0);
        tc::fence_after();
        if (mode == B200GYM_GEMM_FWD) {
            const bool live = grow < P.m, elu = (P.flags & 1) != 0, f32out = (P.flags & 2) != 0;
            for (int n0 = 0; n0 < bn_eff; n0 += 16) {
                uint32_t r[16];
                tc::ld16_issue(taddr + n0, r);
                float bv[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 b4 = *reinterpret_cast<const float4*>(s_bias + n0 + 4 * q);
                    bv[4 * q] = b4.x, bv[4 * q + 1] = b4.y, bv[4 * q + 2] = b4.z, bv[4 * q + 3] = b4.w;
                }
                tc::ld16_wait(r);
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float x = __uint_as_float(r[j]) + bv[j];
                    v[j] = elu ? tc::elu_fast(x) : x;
                }
                if (live) {
                    if (f32out) {
                        float4* o = reinterpret_cast<float4*>(static_cast<float*>(P.out) + static_cast<size_t>(grow) * P.ldo + gcol0 + n0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    } else {
                        uint4* o = reinterpret_cast<uint4*>(static_cast<__half*>(P.out) + static_cast<size_t>(grow) * P.ldo + gcol0 + n0);
                        o[0] = make_uint4(tc::pack_h2(v[0], v[1]), tc::pack_h2(v[2], v[3]), tc::pack_h2(v[4], v[5]), tc::pack_h2(v[6], v[7]));
                        o[1] = make_uint4(tc::pack_h2(v[8], v[9]), tc::pack_h2(v[10], v[11]), tc::pack_h2(v[12], v[13]), tc::pack_h2(v[14], v[15]));
                    }
                }
            }
        } else if (mode == B200GYM_GEMM_DGRAD) {
            const bool live = grow < P.m;
            const __half* aux = static_cast<const __half*>(P.aux);
#pragma unroll
            for (int blk = 0; blk < BN / 16; ++blk) {
                const int n0 = 16 * blk;
                if (n0 >= bn_eff) break;
                uint32_t r[16];
                tc::ld16_issue(taddr + n0, r);
                if (blk == 4) {   // second half of the row: one batch of loads, issued together
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        aux_pre[q] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
                        if (aux != nullptr && live && 64 + 8 * q < bn_eff)
                            aux_pre[q] = __ldg(reinterpret_cast<const uint4*>(aux + static_cast<size_t>(grow) * P.ldaux + gcol0 + 64) + q);
                    }
                }
                const uint4 h0 = aux_pre[2 * (blk & 3)], h1 = aux_pre[2 * (blk & 3) + 1];
                tc::ld16_wait(r);
                const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                uint32_t o[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float2 h = tc::unpack_h2(hw[q]);
                    const float d0 = h.x > 0.0f ? 1.0f : h.x + 1.0f, d1 = h.y > 0.0f ? 1.0f : h.y + 1.0f;   // ELU'(z) from h = ELU(z)
                    o[q] = tc::pack_h2(__uint_as_float(r[2 * q]) * d0, __uint_as_float(r[2 * q + 1]) * d1);
                }
                if (live) {
                    uint4* op = reinterpret_cast<uint4*>(static_cast<__half*>(P.out) + static_cast<size_t>(grow) * P.ldo + gcol0 + n0);
                    op[0] = make_uint4(o[0], o[1], o[2], o[3]);
                    op[1] = make_uint4(o[4], o[5], o[6], o[7]);
                }
            }
        } else {
            const bool live = grow < P.m_real;
            float* out = static_cast<float*>(P.out);
            for (int n0 = 0; n0 < n_mma; n0 += 16) {
                uint32_t r[16];
                tc::ld16_issue(taddr + n0, r);
                tc::ld16_wait(r);
                if (live) {
                    if (n0 < bn_eff) {
                        float* o = out + static_cast<size_t>(grow) * P.ldo + gcol0 + n0;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (gcol0 + n0 + j < P.n_real) atomicAdd(o + j, __uint_as_float(r[j]) * P.scale);
                    } else {
                        atomicAdd(P.bias + grow, __uint_as_float(r[0]) * P.scale);   // column sums of dZ: the ones-column of the H tile
                    }
                }
                __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge before the next one
            }
        }
        tc::fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc::fence_after();
        tc::tmem_dealloc<TMEM_COLS>(tmem);
    }
}

}  // namespace

extern "C" int b200gym_gemm_f16(const B200GemmProblem* problems, int32_t n_problems, void* stream) {
    B200_REQUIRE(problems && n_problems >= 1 && n_problems <= B200GYM_GEMM_MAX_PROBLEMS, B200GYM_EINVAL, "gemm_f16: 1..%d problems",
                 B200GYM_GEMM_MAX_PROBLEMS);
    GemmBatch batch;
    int ctas = 0;
    for (int i = 0; i < n_problems; ++i) {
        const B200GemmProblem& p = problems[i];
        B200_REQUIRE(p.a && p.b && p.out, B200GYM_EINVAL, "gemm_f16: problem %d has a null operand", i);
        B200_REQUIRE(p.mode >= B200GYM_GEMM_FWD && p.mode <= B200GYM_GEMM_WGRAD, B200GYM_EINVAL, "gemm_f16: problem %d: unknown mode %d", i, p.mode);
        B200_REQUIRE(p.m > 0 && p.n > 0 && p.k > 0 && p.n % 16 == 0, B200GYM_EINVAL, "gemm_f16: problem %d: m, n, k > 0 and n %% 16 == 0 (got %d, %d, %d)",
                     i, p.m, p.n, p.k);
        B200_REQUIRE(p.lda % 8 == 0 && p.ldb % 8 == 0 && b200_aligned16(p.a) && b200_aligned16(p.b), B200GYM_EALIGN,
                     "gemm_f16: problem %d: operands must be 16-byte aligned with leading dimensions that are multiples of 8", i);
        int tiles_m;
        if (p.mode == B200GYM_GEMM_WGRAD) {
            B200_REQUIRE(p.m % 16 == 0 && p.splits >= 1 && p.m_real > 0 && p.m_real <= p.m && p.n_real > 0 && p.n_real <= p.n && p.ldo >= p.n_real,
                         B200GYM_EINVAL, "gemm_f16: problem %d (WGRAD): m %% 16 == 0, splits >= 1, 0 < m_real <= m, 0 < n_real <= n <= ldo", i);
            B200_REQUIRE(p.lda >= p.m && p.ldb >= p.n, B200GYM_EINVAL, "gemm_f16: problem %d (WGRAD): lda >= m and ldb >= n", i);
            tiles_m = (p.m + TM - 1) / TM;
        } else {
            B200_REQUIRE(p.k % 16 == 0 && p.lda >= p.k && p.ldo % 8 == 0 && p.ldo >= p.n && b200_aligned16(p.out), B200GYM_EINVAL,
                         "gemm_f16: problem %d: k %% 16 == 0, lda >= k, ldo %% 8 == 0, ldo >= n, 16-byte aligned output", i);
            B200_REQUIRE(p.mode == B200GYM_GEMM_FWD ? p.ldb >= p.k : p.ldb >= p.n, B200GYM_EINVAL, "gemm_f16: problem %d: ldb too small", i);
            B200_REQUIRE(p.aux == nullptr || (p.ldaux % 8 == 0 && p.ldaux >= p.n && b200_aligned16(p.aux)), B200GYM_EALIGN,
                         "gemm_f16: problem %d: aux must be 16-byte aligned with ldaux %% 8 == 0, ldaux >= n", i);
            tiles_m = (p.m + TM - 1) / TM;
        }
        const int tiles_n = (p.n + BN - 1) / BN;
        ctas += tiles_m * tiles_n * (p.mode == B200GYM_GEMM_WGRAD ? p.splits : 1);
        batch.p[i] = p;
        batch.cta_end[i] = ctas;
    }
    batch.n = n_problems;
    static bool configured = false;   // per process; the library runs one process per GPU
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(GEMM_SMEM));
        B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "gemm_f16: cannot reserve %zu B of shared memory: %s", GEMM_SMEM, cudaGetErrorString(e));
        configured = true;
    }
    b200_launch_pdl(0, gemm_f16_kernel, dim3(ctas), dim3(GEMM_THREADS), GEMM_SMEM, static_cast<cudaStream_t>(stream), batch);
    B200_LAUNCH_CHECK("gemm_f16");
    return B200GYM_OK;
}
