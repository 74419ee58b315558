// ppo_chain_kernel — forward, loss and input-gradient chain of one PPO minibatch for nets whose per-tile activations fit in
// shared memory (the 48-128-64-32 flat nets; rsl_rl PPO.update: ActorCritic forward, surrogate / value / entropy loss,
// loss.backward() down to the first layer's pre-activations).  One CTA owns (net, 128-row tile) and runs
//     FWD layer 0 .. L-1  ->  loss on the last layer's TMEM rows  ->  DGRAD layer L-1 .. 1
// back to back: every layer is a tcgen05.mma chain (kind::f16, fp32 accumulators in TMEM) whose A operand is the previous
// epilogue's output, kept in shared memory in the chunk layout of tc.cuh; weights stream through a 2-stage cp.async ring and are
// read K-major (forward) or MN-major (backward) from the same packed copy.  The loss separates per net — the actor CTA computes
// the clipped surrogate / KL / entropy terms and d(loss)/d(mu), the critic CTA the clipped value loss and d(loss)/d(value) — so
// no launch sits between forward and backward.
// Weight gradients (net.flat_grad != NULL): the operands of dW_l = dZ_l^T . H_{l-1} over the tile's 128 rows are the shared-memory
// regions the chain already holds, read MN-major, so each backward step issues that MMA chain next to its input-gradient chain
// (second accumulator, TMEM columns [128, 256)); the otherwise idle loader warps scale the tile's partial by 1/batch and add it to
// the flat gradient buffer with red.global.add.v4.f32 while the next step runs.  One last step holds dW_0 = dZ_0^T . X and the bias
// gradients of every layer (dZ_l^T against a constant [1,0,..] block).  The observation rows are gathered through the minibatch
// index and converted to fp16 in the kernel (net.x32), so a minibatch's forward + backward is this ONE launch and nothing but the
// gradient leaves the SM.  Without flat_grad the kernel is the round-2a form: H_l / dZ_l go to HBM (fp16) as the operands of the
// weight-gradient GEMM (gemm.cu, WGRAD).
// Two CTAs per SM (108 KB of shared memory, 256 TMEM columns each): one CTA's epilogue overlaps the other's MMAs.
#include "tc.cuh"
#include "../../include/b200gym.h"

namespace {

constexpr int TM = 128;
constexpr int KC = 64;
constexpr int WST = 2;                          // weight ring stages
constexpr int CH128 = TM * 16 + 16;             // chunk stride of 128-row tiles (activations, K-major weight tiles)
constexpr int CHKC = KC * 16 + 16;              // chunk stride of KC-row tiles (MN-major weight tiles)
constexpr int W_STAGE = 16 * CHKC;              // 16640 >= 8 * CH128
// EH = epilogue warps per TMEM lane quadrant (1 or 2): warps [0, 4 EH) epilogue, the next four loaders / gradient drain, then the MMA warp
__host__ __device__ constexpr int chain_threads(int eh) { return 128 * eh + 160; }
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t WG_COL = 128;              // first TMEM column of the weight-gradient accumulator
constexpr int MAXA = 16;
static_assert(8 * CH128 <= W_STAGE, "weight stage too small");

struct ChainArgs {
    B200ChainNet net[2];
    B200PpoLossParams lp;
    const long long* idx;
    const float *stdv, *actions, *old_logp, *adv, *ret, *old_v, *old_mu, *old_sigma;
    float* d_std;
    double* scalars;
    int tiles;
    int act_bytes[2];
};

__device__ __forceinline__ void load_tile(unsigned char* dst, int chunk_stride, const __half* base, long long ld, int row0, int col0, int nrows,
                                          int nchunks, int row_lim, int col_lim, int t) {
    const int pieces = nrows * nchunks;
    for (int q = t; q < pieces; q += 128) {
        // chunk counts are powers of two for every net in use (8 = one 64-column stage): shift / mask instead of an integer division
        const int r = (nchunks & (nchunks - 1)) == 0 ? q >> (31 - __clz(nchunks)) : q / nchunks, c = q - r * nchunks;
        const int gr = row0 + r, gc = col0 + 8 * c;
        const bool ok = gr < row_lim && gc < col_lim;
        const __half* src = ok ? base + static_cast<long long>(gr) * ld + gc : base;
        tc::cp_async16(dst + c * chunk_stride + r * 16, src, ok ? 16u : 0u);
    }
}

// debug trace (-DB200GYM_CHAIN_TRACE, tools/trace_chain.py): 6 regions of TRACE_EV (code, clock64) pairs — {epilogue thread 0, loader thread 128, MMA warp}
// of CTA 0 and of the last CTA of the grid (one that starts when a first-round CTA has left)
#ifdef B200GYM_CHAIN_TRACE
constexpr int TRACE_EV = 256;
__device__ unsigned long long* g_chain_trace = nullptr;
struct Trace {
    unsigned long long* p = nullptr;
    int n = 0;
    __device__ __forceinline__ void operator()(int code) {
        if (p != nullptr && n < TRACE_EV) {
            p[2 * n] = static_cast<unsigned long long>(code);
            p[2 * n + 1] = static_cast<unsigned long long>(clock64());
            ++n;
        }
    }
};
#else
struct Trace {   // the production build carries no trace code
    __device__ __forceinline__ void operator()(int) const {}
};
#endif

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int EH>
__global__ void __launch_bounds__(chain_threads(EH), 2) ppo_chain_kernel(const __grid_constant__ ChainArgs a) {
    constexpr int CHAIN_THREADS = chain_threads(EH), LOADER_WARP0 = 4 * EH, MMA_WARP = 4 * EH + 4, LOADER_TID0 = 128 * EH;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
#ifdef B200GYM_CHAIN_TRACE
    unsigned long long t_entry;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_entry));
#endif
    const int which = static_cast<int>(blockIdx.x) / a.tiles, tile = static_cast<int>(blockIdx.x) - which * a.tiles;
    const B200ChainNet& net = a.net[which];
    const int L = net.num_layers;
    const int batch = a.lp.batch;

    // shared memory: activation regions act[0..L-1] (act[l] = A operand of layer l), dz_last (2 chunks), weight ring, barriers
    __shared__ int act_off[B200GYM_CHAIN_MAX_LAYERS + 1];   // act_off[MAX] = end of the regions = dz_last
    int act_end = 0, nsum = 0;   // every thread sums the layer widths itself: one block-wide barrier in the whole set-up
    for (int l = 0; l < L; ++l) act_end += (net.kp[l] >> 3) * CH128, nsum += net.np[l];
    if (tid == 64) {
        int o = 0;
        for (int l = 0; l < B200GYM_CHAIN_MAX_LAYERS; ++l) {
            act_off[l] = o;
            if (l < L) o += (net.kp[l] >> 3) * CH128;
        }
        act_off[B200GYM_CHAIN_MAX_LAYERS] = o;
    }
    unsigned char* dz_last = smem + act_end;
    unsigned char* ring = dz_last + 2 * CH128;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + WST * W_STAGE);
    uint64_t* empty = full + WST;
    uint64_t* accum = empty + WST;
    uint64_t* aready = accum + 1;
    uint64_t* wdone = aready + 1;    // tcgen05.commit after a weight-gradient MMA chain
    uint64_t* wfree = wdone + 1;     // the loader warps have drained the weight-gradient accumulator
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfree + 1);
    float* s_red = reinterpret_cast<float*>(tmem_slot + 2);   // CTA sums: [0, 16) d_std partials, [16, 20) {kl, surrogate, value loss, entropy}
    // the biases of every layer, each zero padded to np[l], back to back; 192 bytes behind the (16-byte aligned) barrier block, by plain
    // pointer arithmetic so that the compiler keeps the shared address space (an integer round trip turns the reads into generic LD)
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(full) + 192);
    float* s_std = s_bias + nsum;   // per-action constants of the Gaussian policy (np[l] are multiples of 16: the block stays 16-byte aligned)

    if (warp == 0) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
    if (tid == 32) {
        for (int s = 0; s < WST; ++s) {
            mbar_init(full + s, net.w_layout == 1 ? 1 : 128);   // TMA bulk copies: one arrive.expect_tx + the bytes; cp.async: every loader thread
            mbar_init(empty + s, 1);
        }
        mbar_init(accum, 1);
        mbar_init(aready, 128 * EH);
        mbar_init(wdone, 1);
        mbar_init(wfree, 128);
        fence_mbar_init();
    }
    if (tid < 20) s_red[tid] = 0.0f;
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = *tmem_slot;
    // programmatic dependent launch: this grid is scheduled while the optimiser kernel of the previous minibatch runs.  Weights, biases
    // and the cleared gradient buffer are that kernel's output: every role executes pdl_wait() before it touches them.  The rollout
    // storage and the minibatch index were written before the update started, so the epilogue warps gather their observation rows (two
    // dependent L2 / HBM round trips) BEFORE the wait, under the optimiser kernel.
    pdl_launch_dependents();
    const bool fuse = net.flat_grad != nullptr;
    // The last step's dW_0 accumulator: when it fits (kp[0] <= 64 columns next to the L x 16 bias columns) it goes into the MAIN accumulator
    // region, which is idle by then, so that step does not wait for the drain of dW_1 (5 K cycles for the 64 x 128 block of the flat nets)
    const bool alt0 = net.kp[0] <= 64 && 16 * L <= 64;
    Trace tr;
#ifdef B200GYM_CHAIN_TRACE
    // CTA Gantt chart: (globaltimer at start, at end, SM id) of every CTA behind the six event regions
    unsigned long long* gantt = g_chain_trace != nullptr ? g_chain_trace + 6 * TRACE_EV * 2 + 4 * static_cast<size_t>(blockIdx.x) : nullptr;
    if (gantt != nullptr && tid == 0) {
        unsigned long long t;
        unsigned int smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        gantt[0] = t_entry, gantt[2] = smid, gantt[3] = t;
    }
    if (g_chain_trace != nullptr && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && (tid == 0 || tid == LOADER_TID0 || tid == 32 * MMA_WARP))
        tr.p = g_chain_trace + static_cast<size_t>((blockIdx.x == 0 ? 0 : 3) + (tid == 0 ? 0 : tid == LOADER_TID0 ? 1 : 2)) * TRACE_EV * 2;
#endif
    tr(0);
    const int nsteps = 2 * L - 1 + (fuse ? 1 : 0);   // L forward layers, DGRAD of layers L-1 .. 1, then (fused) dW_0 + the bias gradients

    if (warp == MMA_WARP) {
        // ------------------------------ MMA issue ------------------------------
        pdl_wait();
        int it = 0;
        for (int step = 0; step < nsteps; ++step) {
            const bool fwd = step < L, last = step == 2 * L - 1;
            const int l = fwd ? step : 2 * L - 1 - step;
            const int ktot = fwd ? net.kp[l] : net.np[l];
            const int n = fwd ? net.np[l] : net.kp[l];
            const unsigned char* dzl = smem + (l == L - 1 ? act_off[B200GYM_CHAIN_MAX_LAYERS] : act_off[l + 1]);   // dZ_l once the backward pass is there
            const unsigned char* abase = fwd ? smem + act_off[l] : dzl;
            const uint32_t idesc = tc::idesc_f16(n, false, !fwd);
            const bool wg = fuse && !fwd;
            tc::mbar_wait_sleep(aready, step & 1);   // long wait: leave the issue slots to the epilogue warps of this SM sub-partition
            tc::fence_after();
            tr(100 + step * 4);       // A operand ready
            const int ns = last ? 0 : (ktot + KC - 1) / KC;
            for (int j = 0; j < ns; ++j, ++it) {
                const int s = it % WST;
                tc::mbar_wait_spin(full + s, (it / WST) & 1);
                tc::fence_after();
                if (tc::elect_one()) {
                    const unsigned char* bt = ring + s * W_STAGE;
                    uint64_t da = tc::smem_desc(abase + (8 * j) * CH128, CH128, 128);
                    uint64_t db = fwd ? tc::smem_desc(bt, CH128, 128) : tc::smem_desc(bt, 128, CHKC);
                    const uint32_t b_step = fwd ? ((2u * CH128) >> 4) : (256u >> 4);
                    const int nk = min(KC, ktot - j * KC) >> 4;
                    for (int q = 0; q < nk; ++q) {
                        tc::mma_f16(tmem, da, db, idesc, (j | q) != 0 ? 1u : 0u);
                        da += (2u * CH128) >> 4;
                        db += b_step;
                    }
                    tc::commit(empty + s);
                    if (j == ns - 1 && !wg) tc::commit(accum);
                }
                __syncwarp();
            }
            tr(101 + step * 4);       // forward / input-gradient MMAs issued
            if (wg) {
                // dW_l = dZ_l^T . H_{l-1}: both operands are [128 rows x width] chunk regions, contracted over the rows (MN-major).
                // Accumulator = 128 TMEM columns: layer 0's input may be wider (kp[0] <= 256 + ...), so it goes in blocks of 128 columns.
                const int nblk = (net.kp[l] + 127) >> 7;
                int s = 0;
                for (int blk = 0; blk < nblk; ++blk) {
                    const int jw = step - L + blk;
                    const bool in_main = last && alt0;
                    if (jw > 0 && !in_main) {
                        tc::mbar_wait_sleep(wfree, (jw - 1) & 1);
                        tc::fence_after();
                    }
                    if (last && blk == 0) {
                        s = it % WST;
                        tc::mbar_wait_spin(full + s, (it / WST) & 1);   // the constant block of the bias-gradient MMAs
                        tc::fence_after();
                        ++it;
                    }
                    if (tc::elect_one()) {
                        const int ncols = min(128, net.kp[l] - 128 * blk);
                        const uint32_t idw = tc::idesc_f16(ncols, true, true);
                        uint64_t da = tc::smem_desc(dzl, 128, CH128), db = tc::smem_desc(smem + act_off[l] + 16 * blk * CH128, 128, CH128);
                        for (int q = 0; q < TM / 16; ++q) {
                            tc::mma_f16(tmem + (in_main ? 64u : WG_COL), da, db, idw, q != 0 ? 1u : 0u);
                            da += 256u >> 4;
                            db += 256u >> 4;
                        }
                        if (last && blk == 0) {
                            const uint32_t idb = tc::idesc_f16(16, true, true);
                            for (int m = 0; m < L; ++m) {
                                uint64_t dam = tc::smem_desc(smem + (m == L - 1 ? act_off[B200GYM_CHAIN_MAX_LAYERS] : act_off[m + 1]), 128, CH128);
                                uint64_t dbm = tc::smem_desc(ring + s * W_STAGE, 128, CH128);
                                for (int q = 0; q < TM / 16; ++q) {
                                    tc::mma_f16(tmem + 16 * m, dam, dbm, idb, q != 0 ? 1u : 0u);
                                    dam += 256u >> 4;
                                    dbm += 256u >> 4;
                                }
                            }
                        }
                        tc::commit(wdone);
                        // also orders the weight-gradient reads of H_{l-1} before the epilogue overwrites it with dZ_{l-1}
                        if (blk == 0) tc::commit(accum);
                    }
                    __syncwarp();
                }
                tr(102 + step * 4);   // weight-gradient MMAs issued
            }
        }
    } else if (warp >= LOADER_WARP0) {
        // ------------------------------ loaders: the input tile, the weight tiles of every step, the weight-gradient drain -------------
        const int t = tid - LOADER_TID0;
        pdl_wait();   // the fp16 weights are the previous optimiser kernel's output
        if (net.x32 == nullptr) {
            load_tile(smem + act_off[0], CH128, static_cast<const __half*>(net.x), net.ldx, tile * TM, 0, TM, net.kp[0] >> 3, batch, net.kp[0], t);
            tc::cp_async_commit();
            tc::cp_async_wait<0>();
            fence_proxy_async();
            tc::mbar_arrive(aready);   // the input tile = step 0's A operand
        }
        const __half* w16 = static_cast<const __half*>(net.w16);
        const bool wtma = net.w_layout == 1;
        const uint32_t taddr_q = tmem + (static_cast<uint32_t>((warp - LOADER_WARP0) * 32) << 16);
        // drains the jw-th weight-gradient accumulator (layer L-1-jw): TMEM lane = row of W_l, scaled by 1/batch, added to the flat gradient
        auto drain = [&](int jw) {
            const int l = jw < L - 1 ? L - 1 - jw : 0;
            const int col0 = jw < L - 1 ? 0 : 128 * (jw - (L - 1));   // layer 0 goes in blocks of 128 input columns
            const uint32_t taddr_w = taddr_q + ((alt0 && jw >= L - 1) ? 64u : WG_COL);
            tr(300 + jw * 4);
            tc::mbar_wait_sleep(wdone, jw & 1);
            tc::fence_after();
            tr(301 + jw * 4);
            const int kr = net.k_real[l];
            if ((warp - LOADER_WARP0) * 32 < net.n_real[l]) {   // warp-uniform: this warp owns live rows of W_l
                const bool rowlive = t < net.n_real[l];
                float* g = net.flat_grad + net.w32_off[l] + static_cast<size_t>(t) * kr;
                const bool vec = (kr & 3) == 0 && (net.w32_off[l] & 3) == 0;
                const float sc = a.lp.inv_global_batch;
                for (int n0 = 0; n0 < 128 && col0 + n0 < kr; n0 += 16) {
                    uint32_t r[16];
                    tc::ld16_issue(taddr_w + n0, r);
                    tc::ld16_wait(r);
                    if (rowlive) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int c = col0 + n0 + 4 * q;
                            if (vec && c + 4 <= kr) {
                                red_add_v4(g + c, __uint_as_float(r[4 * q]) * sc, __uint_as_float(r[4 * q + 1]) * sc, __uint_as_float(r[4 * q + 2]) * sc,
                                           __uint_as_float(r[4 * q + 3]) * sc);
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (c + e < kr) atomicAdd(g + c + e, __uint_as_float(r[4 * q + e]) * sc);
                            }
                        }
                    }
                    __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge before the next one
                }
            }
            tc::fence_before();
            tc::mbar_arrive(wfree);
            tr(302 + jw * 4);
        };
        int it = 0, pending = -1;   // pending: ring stage whose copies are committed but not yet published to the MMA warp
        for (int step = 0; step < nsteps; ++step) {
            const bool fwd = step < L, last = step == 2 * L - 1;
            const int l = fwd ? step : 2 * L - 1 - step;
            const int ktot = fwd ? net.kp[l] : net.np[l];
            const __half* W = w16 + net.w_off[l];   // row-major [np[l], kp[l]]
            const int ns = last ? 0 : (ktot + KC - 1) / KC;
            for (int j = 0; j < ns; ++j, ++it) {
                const int s = it % WST;
                unsigned char* bt = ring + s * W_STAGE;
                const int kc_eff = min(KC, ktot - j * KC);
                if (wtma) {
                    // chunk-major weights [kp/8][np][8]: the stage is a handful of contiguous pieces, moved by one thread as TMA bulk
                    // copies that complete on the stage's barrier (no per-thread 16-byte cp.async, no publish step)
                    if (t == 0) {
                        tc::mbar_wait_sleep(empty + s, ((it / WST) & 1) ^ 1);
                        const int np_l = net.np[l];
                        if (fwd) {   // [np rows x kc_eff cols], K-major: chunk c of the stage = np_l * 16 contiguous bytes
                            const int nch = kc_eff >> 3;
                            const uint32_t piece = static_cast<uint32_t>(np_l) * 16u;
                            mbar_expect_tx(full + s, piece * nch);
                            for (int c = 0; c < nch; ++c)
                                bulk_g2s(bt + c * CH128, W + (static_cast<size_t>(8 * j + c) * np_l) * 8, piece, full + s);
                        } else {     // [kc_eff rows (of np) x kp cols], MN-major: chunk c = rows j*KC .. of column chunk c, kc_eff * 16 bytes
                            const int nch = net.kp[l] >> 3;
                            const uint32_t piece = static_cast<uint32_t>(kc_eff) * 16u;
                            mbar_expect_tx(full + s, piece * nch);
                            for (int c = 0; c < nch; ++c)
                                bulk_g2s(bt + c * CHKC, W + (static_cast<size_t>(c) * np_l + j * KC) * 8, piece, full + s);
                        }
                    }
                    tr(200 + it);
                    continue;
                }
                if (pending >= 0 && !tc::mbar_try(empty + s, ((it / WST) & 1) ^ 1)) {
                    // about to block on the ring: hand the stage already in flight to the MMA warp first
                    tc::cp_async_wait<0>();
                    fence_proxy_async();
                    tc::mbar_arrive(full + pending);
                    pending = -1;
                }
                tc::mbar_wait_sleep(empty + s, ((it / WST) & 1) ^ 1);
                tr(250 + it);
                if (fwd) load_tile(bt, CH128, W, net.kp[l], 0, j * KC, net.np[l], kc_eff >> 3, net.np[l], net.kp[l], t);       // [n rows x k cols], K-major
                else load_tile(bt, CHKC, W, net.kp[l], j * KC, 0, kc_eff, net.kp[l] >> 3, net.np[l], net.kp[l], t);            // [n rows = K x k cols = N], MN-major
                tc::cp_async_commit();
                if (pending >= 0) {   // everything before this stage has landed: publish it to the tensor-core proxy
                    tc::cp_async_wait<1>();
                    fence_proxy_async();
                    tc::mbar_arrive(full + pending);
                }
                pending = s;
                tr(200 + it);         // stage `it` issued, stage `it - 1` published
            }
            if (fuse && !fwd) {
                if (pending >= 0) {
                    tc::cp_async_wait<0>();
                    fence_proxy_async();
                    tc::mbar_arrive(full + pending);
                    pending = -1;
                }
                if (last) {
                    // constant block of the bias-gradient MMAs: 16 columns x 128 rows, column 0 = 1 (fp16), the rest 0
                    const int s = it % WST;
                    tc::mbar_wait_sleep(empty + s, ((it / WST) & 1) ^ 1);
                    unsigned char* bt = ring + s * W_STAGE;
                    *reinterpret_cast<uint4*>(bt + t * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(bt + CH128 + t * 16) = make_uint4(0u, 0u, 0u, 0u);
                    fence_proxy_async();
                    if (wtma) {   // the stage barrier counts ONE arrival in this mode
                        asm volatile("bar.sync 2, 128;" ::: "memory");
                        if (t == 0) tc::mbar_arrive(full + s);
                    } else {
                        tc::mbar_arrive(full + s);
                    }
                    ++it;
                }
                if (step > L) drain(step - L - 1);   // the previous step's weight gradient, while this step's MMAs and epilogue run
            }
        }
        if (pending >= 0) {
            tc::cp_async_wait<0>();
            fence_proxy_async();
            tc::mbar_arrive(full + pending);
        }
        if (fuse && !alt0)   // (alt0: the epilogue warps drain dW_0 from the main accumulator)
            for (int blk = 0; blk < (net.kp[0] + 127) >> 7; ++blk) drain(L - 1 + blk);
    } else {
        // ------------------------------ epilogue warps: TMEM lane = tile row; two warps per lane quadrant, each takes half of the
        // accumulator's 16-column blocks (a lone warp per scheduler issues at ~0.25 IPC: the chain is bound by this role) --------------
        const int row = tid & 127, half = EH == 2 ? tid >> 7 : 0;
        const uint32_t taddr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const int grow = tile * TM + row;
        const bool live = grow < batch;
        const int A = a.lp.num_actions;
        // per-sample storage columns of this row (half 0 computes the loss), fetched through the minibatch index while the first layers run
        const long long srow = live ? (a.idx ? a.idx[grow] : grow) : 0;
        // (no L1 to speak of next to 2 x 113 KB of shared memory: every global or local access is an L2 round trip, so the loss step
        // takes its 38 per-row values in one batch of loads when it gets there, from lines prefetched into L2 here)
        float s_adv = 0.f, s_logp = 0.f, s_ret = 0.f, s_oldv = 0.f;
        if (half == 0 && live) {
            if (which == 0) {
                const float* rows[3] = {a.actions + srow * A, a.old_mu + srow * A, a.old_sigma + srow * A};
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(rows[q]));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(rows[q] + A - 1));
                }
                s_adv = __ldg(a.adv + srow), s_logp = __ldg(a.old_logp + srow);
            } else {
                s_ret = __ldg(a.ret + srow), s_oldv = __ldg(a.old_v + srow);
            }
        }
        if (net.x32 != nullptr) {
            // this thread's half of its observation row through the minibatch index, fp32 -> fp16, straight into the A operand of layer 0
            // (the loader warps are fetching the first weight tiles meanwhile)
            const int chunks0 = net.kp[0] >> 3, kr0 = net.k_real[0];
            const int c_lo = half ? (chunks0 + 1) >> 1 : 0, c_hi = (EH == 1 || half) ? chunks0 : (chunks0 + 1) >> 1;
            const float* src = net.x32 + static_cast<size_t>(srow) * net.ldx32;
            unsigned char* dst = smem + act_off[0] + row * 16;
            if ((kr0 & 7) == 0 && (net.ldx32 & 3) == 0) {
                for (int c0 = c_lo; c0 < c_hi; c0 += 4) {
                    float4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        v[u] = (live && c0 + (u >> 1) < c_hi && 8 * c0 + 4 * u < kr0) ? __ldg(reinterpret_cast<const float4*>(src + 8 * c0) + u)
                                                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (c0 + u < c_hi)
                            *reinterpret_cast<uint4*>(dst + (c0 + u) * CH128) =
                                make_uint4(tc::pack_h2(v[2 * u].x, v[2 * u].y), tc::pack_h2(v[2 * u].z, v[2 * u].w), tc::pack_h2(v[2 * u + 1].x, v[2 * u + 1].y),
                                           tc::pack_h2(v[2 * u + 1].z, v[2 * u + 1].w));
                }
            } else {
                for (int c = c_lo; c < c_hi; ++c) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = (live && 8 * c + e < kr0) ? __ldg(src + 8 * c + e) : 0.0f;
                    *reinterpret_cast<uint4*>(dst + c * CH128) =
                        make_uint4(tc::pack_h2(v[0], v[1]), tc::pack_h2(v[2], v[3]), tc::pack_h2(v[4], v[5]), tc::pack_h2(v[6], v[7]));
                }
            }
            fence_proxy_async();
            tc::mbar_arrive(aready);   // the input tile = step 0's A operand
            tr(399);
        } else if (EH == 2 && half == 1) {
            tc::mbar_arrive(aready);   // fp16 input copied by the 128 loader threads: make up the barrier's 256 arrivals
        }
        pdl_wait();   // from here on: parameters of the previous optimiser step, and (at the end) the gradient buffer it cleared
        // biases and policy constants into shared memory while step 0's MMAs run (only these warps read them): sigma, log sigma,
        // 1/(2 sigma^2), 1/sigma^2, 1/sigma are computed once per CTA instead of once per row
        for (int l = 0, o = 0; l < L; o += net.np[l], ++l)
            for (int c = tid; c < net.np[l]; c += 128 * EH) s_bias[o + c] = c < net.n_real[l] ? __ldg(net.flat_param + net.b_off[l] + c) : 0.0f;
        if (tid < MAXA) {
            const float sg = tid < a.lp.num_actions ? __ldg(a.stdv + tid) : 1.0f;
            s_std[tid] = sg, s_std[MAXA + tid] = logf(sg), s_std[2 * MAXA + tid] = 1.0f / (2.0f * sg * sg), s_std[3 * MAXA + tid] = 1.0f / (sg * sg);
            s_std[4 * MAXA + tid] = 1.0f / sg;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(128 * EH) : "memory");
        int sboff = 0;   // offset of layer l's bias in s_bias (forward steps visit the layers in order)
        for (int step = 0; step < nsteps; ++step) {
            const bool fwd = step < L;
            const int l = fwd ? step : 2 * L - 1 - step;
            tr(400 + step * 4);
            tc::mbar_wait_sleep(accum, step & 1);
            tc::fence_after();
            tr(401 + step * 4);       // accumulator ready
            if (fwd && l < L - 1) {
                // bias + ELU -> fp16: A operand of layer l+1 (shared memory) and, in the unfused form, H_l (HBM, operand of the weight-gradient GEMM)
                const int n = net.np[l], nb = n >> 4;
                const int b_lo = half ? (nb + 1) >> 1 : 0, b_hi = (EH == 1 || half) ? nb : (nb + 1) >> 1;
                const float* sb = s_bias + sboff;
                unsigned char* dst = smem + act_off[l + 1] + row * 16;
                __half* hg = static_cast<__half*>(net.h[l]) + static_cast<size_t>(grow) * n;
                const bool keep_h = live && net.h[l] != nullptr;
                for (int n0 = 16 * b_lo; n0 < 16 * b_hi; n0 += 16) {
                    uint32_t r[16];
                    if (step == 0) tr(500 + n0);
                    tc::ld16_issue(taddr + n0, r);
                    float bv[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 b4 = *reinterpret_cast<const float4*>(sb + n0 + 4 * q);
                        bv[4 * q] = b4.x, bv[4 * q + 1] = b4.y, bv[4 * q + 2] = b4.z, bv[4 * q + 3] = b4.w;
                    }
                    tc::ld16_wait(r);
                    if (step == 0) tr(501 + n0);
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = tc::elu_fast(__uint_as_float(r[j]) + bv[j]);
                    const uint4 p0 = make_uint4(tc::pack_h2(v[0], v[1]), tc::pack_h2(v[2], v[3]), tc::pack_h2(v[4], v[5]), tc::pack_h2(v[6], v[7]));
                    const uint4 p1 = make_uint4(tc::pack_h2(v[8], v[9]), tc::pack_h2(v[10], v[11]), tc::pack_h2(v[12], v[13]), tc::pack_h2(v[14], v[15]));
                    *reinterpret_cast<uint4*>(dst + (n0 >> 3) * CH128) = p0;
                    *reinterpret_cast<uint4*>(dst + ((n0 >> 3) + 1) * CH128) = p1;
                    if (step == 0) tr(502 + n0);
                    if (keep_h) {
                        reinterpret_cast<uint4*>(hg + n0)[0] = p0;
                        reinterpret_cast<uint4*>(hg + n0)[1] = p1;
                    }
                }
                sboff += n;
            } else if (fwd && half == 0) {
                // last layer: 16 output columns of this row -> loss terms and d(loss)/d(output), unscaled (1/batch lives in the gradient drain)
                uint32_t r[16];
                tc::ld16_issue(taddr, r);
                tc::ld16_wait(r);
                tr(601);
                const float* sb = s_bias + sboff;
                float o[16], dz[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    o[j] = __uint_as_float(r[j]) + sb[j];
                    dz[j] = 0.0f;
                }
                if (net.out != nullptr && live) {
                    float4* op = reinterpret_cast<float4*>(net.out + static_cast<size_t>(grow) * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) op[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                }
                tr(602);
                float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                float dstd[MAXA];
#pragma unroll
                for (int j = 0; j < MAXA; ++j) dstd[j] = 0.0f;
                if (which == 0 && live) {
                    const float LOG_SQRT_2PI = 0.91893853320467274178f;
                    float logp = 0.0f, ent = 0.0f, kl = 0.0f, diff[MAXA];
#pragma unroll
                    for (int j = 0; j < MAXA; ++j) diff[j] = 0.0f;
                    auto term = [&](int j, float x, float om, float os) {
                        const float sg = s_std[j], ls = s_std[MAXA + j], iden = s_std[2 * MAXA + j], m = o[j];
                        diff[j] = x - m;
                        logp += -(diff[j] * diff[j]) * iden - ls - LOG_SQRT_2PI;
                        ent += 0.5f + LOG_SQRT_2PI + ls;
                        const float dm = om - m;
                        kl += logf(__fdividef(sg, os) + 1.e-5f) + (os * os + dm * dm) * iden - 0.5f;
                    };
                    if (A == 12) {
                        float4 x4[3], m4[3], g4[3];
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            x4[q] = __ldg(reinterpret_cast<const float4*>(a.actions + srow * 12) + q);
                            m4[q] = __ldg(reinterpret_cast<const float4*>(a.old_mu + srow * 12) + q);
                            g4[q] = __ldg(reinterpret_cast<const float4*>(a.old_sigma + srow * 12) + q);
                        }
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            term(4 * q, x4[q].x, m4[q].x, g4[q].x);
                            term(4 * q + 1, x4[q].y, m4[q].y, g4[q].y);
                            term(4 * q + 2, x4[q].z, m4[q].z, g4[q].z);
                            term(4 * q + 3, x4[q].w, m4[q].w, g4[q].w);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < MAXA; ++j)
                            if (j < A) term(j, __ldg(a.actions + srow * A + j), __ldg(a.old_mu + srow * A + j), __ldg(a.old_sigma + srow * A + j));
                    }
                    const float ratio = expf(logp - s_logp);
                    const float lo = 1.0f - a.lp.clip_param, hi = 1.0f + a.lp.clip_param;
                    const float s1 = -s_adv * ratio, s2 = -s_adv * fminf(fmaxf(ratio, lo), hi);
                    const bool inside = ratio >= lo && ratio <= hi;
                    const float g_ratio = s1 > s2 ? -s_adv : (s1 == s2 ? (inside ? -s_adv : -0.5f * s_adv) : 0.0f);   // torch.max splits ties
                    const float dlogp = g_ratio * ratio;
#pragma unroll
                    for (int j = 0; j < MAXA; ++j) {
                        if (j < A) {
                            const float inv2 = s_std[3 * MAXA + j], isg = s_std[4 * MAXA + j];
                            dz[j] = dlogp * diff[j] * inv2;
                            dstd[j] = (dlogp * (diff[j] * diff[j] * inv2 * isg - isg) - a.lp.entropy_coef * isg) * a.lp.inv_global_batch;
                        }
                    }
                    acc[0] = kl, acc[1] = fmaxf(s1, s2), acc[3] = ent;
                } else if (which == 1 && live) {
                    const float v = o[0], R = s_ret;
                    float vloss, dv;
                    if (a.lp.use_clipped_value_loss) {
                        const float dvo = v - s_oldv;
                        const float vc = s_oldv + fminf(fmaxf(dvo, -a.lp.clip_param), a.lp.clip_param);
                        const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
                        const bool pass = dvo >= -a.lp.clip_param && dvo <= a.lp.clip_param;
                        vloss = fmaxf(l1, l2);
                        const float g1 = 2.0f * (v - R), g2 = pass ? 2.0f * (vc - R) : 0.0f;
                        dv = l1 > l2 ? g1 : (l1 == l2 ? 0.5f * (g1 + g2) : g2);
                    } else {
                        vloss = (R - v) * (R - v);
                        dv = 2.0f * (v - R);
                    }
                    dz[0] = a.lp.value_loss_coef * dv;
                    acc[2] = vloss;
                }
                tr(603 + (__float_as_uint(dz[0]) == 0x7fc12345u));
                const uint4 p0 = make_uint4(tc::pack_h2(dz[0], dz[1]), tc::pack_h2(dz[2], dz[3]), tc::pack_h2(dz[4], dz[5]), tc::pack_h2(dz[6], dz[7]));
                const uint4 p1 = make_uint4(tc::pack_h2(dz[8], dz[9]), tc::pack_h2(dz[10], dz[11]), tc::pack_h2(dz[12], dz[13]), tc::pack_h2(dz[14], dz[15]));
                *reinterpret_cast<uint4*>(dz_last + row * 16) = p0;
                *reinterpret_cast<uint4*>(dz_last + CH128 + row * 16) = p1;
                if (live && net.dz[l] != nullptr) {
                    uint4* zg = reinterpret_cast<uint4*>(static_cast<__half*>(net.dz[l]) + static_cast<size_t>(grow) * 16);
                    zg[0] = p0, zg[1] = p1;
                }
                tr(605);
                // CTA-level sums.  16 values per lane (d_std[0..12), the four loss sums) are reduced over the warp by exchanging half of the
                // remaining values per step (8 + 4 + 2 + 1 + 1 shuffles instead of 16 x 5); lane 2 k ends up with the warp sum of value
                // bitrev4(k), which it adds to the CTA's shared-memory sums.  More than 12 actions: a second pass for d_std[12..16).
                const int lane = tid & 31;
                auto warp_sum16 = [&](float (&v)[16]) -> float {
#pragma unroll
                    for (int w = 8, d = 16; w >= 1; w >>= 1, d >>= 1) {
                        const bool up = (lane & d) != 0;
#pragma unroll
                        for (int i = 0; i < w; ++i) {
                            const float send = up ? v[i] : v[i + w], keep = up ? v[i + w] : v[i];
                            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
                        }
                    }
                    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
                };
                const int slot = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);   // value index this lane ends up with
                {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 12; ++j) v[j] = dstd[j];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[12 + k] = acc[k];
                    const float sum = warp_sum16(v);
                    if ((lane & 1) == 0 && sum != 0.0f) atomicAdd(s_red + (slot < 12 ? slot : slot + 4), sum);
                }
                if (A > 12) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = j < 4 ? dstd[12 + j] : 0.0f;
                    const float sum = warp_sum16(v);
                    if ((lane & 1) == 0 && slot < 4 && sum != 0.0f) atomicAdd(s_red + 12 + slot, sum);
                }
                tr(606);
                // the CTA's sums go to global memory at the very end of the kernel: every CTA of a wave reaches this step at the same time,
                // and the proxy fence below would wait for ~300 same-address atomics to be acknowledged (measured: 12 us of a 40 us CTA)
            } else if (fwd) {
                // half 1 has no part in the 16-column loss step
            } else if (step == 2 * L - 1) {
                // bias gradients: column 16 m of the accumulator = sum over the tile's rows of dZ_m[:, lane]; layers alternate between the halves
                for (int m = half; m < L; m += EH) {
                    if ((warp & 3) * 32 < net.n_real[m]) {   // warp-uniform
                        uint32_t r[16];
                        tc::ld16_issue(taddr + 16 * m, r);
                        tc::ld16_wait(r);
                        if (row < net.n_real[m]) atomicAdd(net.flat_grad + net.b_off[m] + row, __uint_as_float(r[0]) * a.lp.inv_global_batch);
                        __syncwarp();
                    }
                }
                if (alt0 && (warp & 3) * 32 < net.n_real[0]) {
                    // dW_0 sits in columns [64, 64 + kp[0]) of this accumulator: the epilogue warps drain it (row = TMEM lane, the halves
                    // alternate over the 16-column blocks) while the loader warps are still draining dW_1
                    const int kr = net.k_real[0];
                    float* g = net.flat_grad + net.w32_off[0] + static_cast<size_t>(row) * kr;
                    const bool vec = (kr & 3) == 0 && (net.w32_off[0] & 3) == 0;
                    const float sc = a.lp.inv_global_batch;
                    for (int n0 = 16 * half; n0 < kr; n0 += 16 * EH) {
                        uint32_t r[16];
                        tc::ld16_issue(taddr + 64 + n0, r);
                        tc::ld16_wait(r);
                        if (row < net.n_real[0]) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int c = n0 + 4 * q;
                                if (vec && c + 4 <= kr) {
                                    red_add_v4(g + c, __uint_as_float(r[4 * q]) * sc, __uint_as_float(r[4 * q + 1]) * sc, __uint_as_float(r[4 * q + 2]) * sc,
                                               __uint_as_float(r[4 * q + 3]) * sc);
                                } else {
#pragma unroll
                                    for (int e = 0; e < 4; ++e)
                                        if (c + e < kr) atomicAdd(g + c + e, __uint_as_float(r[4 * q + e]) * sc);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            } else {
                // DGRAD of layer l: dZ_{l-1} = (dZ_l . W_l) * ELU'(H_{l-1}); H_{l-1} sits in act[l] and is overwritten in place by dZ_{l-1}
                const int n = net.kp[l], nb = n >> 4;
                const int b_lo = half ? (nb + 1) >> 1 : 0, b_hi = (EH == 1 || half) ? nb : (nb + 1) >> 1;
                unsigned char* hs = smem + act_off[l] + row * 16;
                __half* zg = static_cast<__half*>(net.dz[l - 1]) + static_cast<size_t>(grow) * n;
                const bool keep_dz = live && net.dz[l - 1] != nullptr;
                for (int n0 = 16 * b_lo; n0 < 16 * b_hi; n0 += 16) {
                    uint32_t r[16];
                    tc::ld16_issue(taddr + n0, r);
                    const uint4 h0 = *reinterpret_cast<const uint4*>(hs + (n0 >> 3) * CH128), h1 = *reinterpret_cast<const uint4*>(hs + ((n0 >> 3) + 1) * CH128);
                    tc::ld16_wait(r);
                    const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                    uint32_t o[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float2 h = tc::unpack_h2(hw[q]);
                        const float d0 = h.x > 0.0f ? 1.0f : h.x + 1.0f, d1 = h.y > 0.0f ? 1.0f : h.y + 1.0f;
                        o[q] = tc::pack_h2(__uint_as_float(r[2 * q]) * d0, __uint_as_float(r[2 * q + 1]) * d1);
                    }
                    const uint4 p0 = make_uint4(o[0], o[1], o[2], o[3]), p1 = make_uint4(o[4], o[5], o[6], o[7]);
                    *reinterpret_cast<uint4*>(hs + (n0 >> 3) * CH128) = p0;
                    *reinterpret_cast<uint4*>(hs + ((n0 >> 3) + 1) * CH128) = p1;
                    if (keep_dz) {
                        reinterpret_cast<uint4*>(zg + n0)[0] = p0;
                        reinterpret_cast<uint4*>(zg + n0)[1] = p1;
                    }
                }
            }
            // this thread's part of the next A operand is written and its accumulator reads are done
            if (step == 0 || step == L - 1) tr(598);
            fence_proxy_async();
            if (step == L - 1) tr(599);
            tc::fence_before();
            if (step + 1 < nsteps) tc::mbar_arrive(aready);
            tr(402 + step * 4);       // epilogue done
        }
    }
    tr(999);
    tc::fence_before();
    __syncthreads();
    if (tid < 4 && s_red[16 + tid] != 0.0f) atomicAdd(a.scalars + tid, static_cast<double>(s_red[16 + tid]));
    if (which == 0 && tid >= 32 && tid < 32 + a.lp.num_actions) atomicAdd(a.d_std + tid - 32, s_red[tid - 32]);
#ifdef B200GYM_CHAIN_TRACE
    if (gantt != nullptr && tid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        gantt[1] = t;
    }
#endif
    if (warp == 0) {
        tc::fence_after();
        tc::tmem_dealloc<TMEM_COLS>(tmem);
    }
}

size_t chain_smem(const B200ChainNet& n) {
    size_t chunks = 2;
    for (int l = 0; l < n.num_layers; ++l) chunks += static_cast<size_t>(n.kp[l] >> 3);
    size_t nsum = 0;
    for (int l = 0; l < n.num_layers; ++l) nsum += static_cast<size_t>(n.np[l]);
    return chunks * CH128 + static_cast<size_t>(WST) * W_STAGE + 256 + nsum * 4 + 5 * MAXA * 4;
}

}  // namespace

extern "C" int b200gym_ppo_chain(const B200ChainNet* actor, const B200ChainNet* critic, const B200PpoLossParams* lp, const int64_t* idx,
                                 const float* std, const float* actions, const float* old_log_prob, const float* advantages,
                                 const float* returns, const float* old_values, const float* old_mu, const float* old_sigma, float* d_std,
                                 double* scalars, void* stream) {
    B200_REQUIRE(actor && critic && lp && std && actions && old_log_prob && advantages && returns && old_values && old_mu && old_sigma &&
                     d_std && scalars,
                 B200GYM_EINVAL, "ppo_chain: null argument");
    B200_REQUIRE(lp->batch > 0 && lp->num_actions > 0 && lp->num_actions <= MAXA, B200GYM_EINVAL, "ppo_chain: batch > 0, 1..%d actions", MAXA);
    ChainArgs a;
    size_t smem = 0;
    const B200ChainNet* nets[2] = {actor, critic};
    for (int w = 0; w < 2; ++w) {
        const B200ChainNet& n = *nets[w];
        B200_REQUIRE(n.num_layers >= 2 && n.num_layers <= B200GYM_CHAIN_MAX_LAYERS, B200GYM_EINVAL, "ppo_chain: 2..%d layers", B200GYM_CHAIN_MAX_LAYERS);
        B200_REQUIRE(n.w16 && n.flat_param && b200_aligned16(n.w16), B200GYM_EALIGN, "ppo_chain: w16 must be 16-byte aligned");
        if (n.x32 != nullptr)
            B200_REQUIRE(b200_aligned16(n.x32) && n.ldx32 >= n.k_real[0] && n.k_real[0] > 0 && n.k_real[0] <= n.kp[0], B200GYM_EALIGN,
                         "ppo_chain: x32 must be 16-byte aligned with 0 < k_real[0] <= ldx32, kp[0]");
        else
            B200_REQUIRE(n.x && b200_aligned16(n.x) && n.ldx % 8 == 0 && n.ldx >= n.kp[0], B200GYM_EALIGN,
                         "ppo_chain: x must be 16-byte aligned, ldx %% 8 == 0");
        B200_REQUIRE(n.flat_grad == nullptr || b200_aligned16(n.flat_grad), B200GYM_EALIGN, "ppo_chain: flat_grad must be 16-byte aligned");
        for (int l = 0; l < n.num_layers; ++l) {
            B200_REQUIRE(n.kp[l] % 16 == 0 && n.np[l] % 16 == 0 && n.np[l] >= 16 && n.np[l] <= 128 && n.kp[l] >= 16 && n.n_real[l] <= n.np[l],
                         B200GYM_EINVAL, "ppo_chain: layer %d is %d -> %d; widths must be multiples of 16, outputs <= 128", l, n.kp[l], n.np[l]);
            B200_REQUIRE(l == 0 || (n.kp[l] == n.np[l - 1] && n.kp[l] <= 128), B200GYM_EINVAL, "ppo_chain: layer %d input %d != previous output %d", l, n.kp[l],
                         n.np[l - 1]);
            B200_REQUIRE(n.w_off[l] % 8 == 0 && b200_aligned16(n.dz[l]) && b200_aligned16(n.h[l]), B200GYM_EALIGN,
                         "ppo_chain: layer %d buffers misaligned", l);
            if (n.flat_grad == nullptr)
                B200_REQUIRE(n.dz[l] && (l == n.num_layers - 1 || n.h[l]), B200GYM_EINVAL,
                             "ppo_chain: layer %d: without flat_grad the h / dz operand buffers of the weight-gradient GEMM are required", l);
            else
                B200_REQUIRE(n.k_real[l] > 0 && n.k_real[l] <= n.kp[l] && n.w32_off[l] >= 0, B200GYM_EINVAL,
                             "ppo_chain: layer %d: 0 < k_real <= kp and w32_off >= 0 are required with flat_grad", l);
        }
        B200_REQUIRE(n.np[n.num_layers - 1] == 16, B200GYM_EINVAL, "ppo_chain: the last layer must be padded to 16 outputs");
        a.net[w] = n;
        const size_t s = chain_smem(n);
        smem = s > smem ? s : smem;
    }
    B200_REQUIRE(smem <= 227 * 1024, B200GYM_EINVAL, "ppo_chain: %zu B of shared memory needed per CTA: use the layered GEMM path for this net", smem);
    a.lp = *lp;
    a.idx = reinterpret_cast<const long long*>(idx);
    a.stdv = std, a.actions = actions, a.old_logp = old_log_prob, a.adv = advantages, a.ret = returns, a.old_v = old_values;
    a.old_mu = old_mu, a.old_sigma = old_sigma, a.d_std = d_std, a.scalars = scalars;
    a.tiles = (lp->batch + TM - 1) / TM;
    static int eh = 0;   // epilogue warps per lane quadrant: 2 by default, B200GYM_CHAIN_EPI=1 for the A/B
    if (!eh) {
        const char* e = getenv("B200GYM_CHAIN_EPI");
        eh = (e && e[0] == '1') ? 1 : 2;
    }
    auto kern = eh == 1 ? ppo_chain_kernel<1> : ppo_chain_kernel<2>;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "ppo_chain: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
        // two CTAs per SM need the whole unified L1 / shared memory as shared memory: say so up front
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "ppo_chain: carve-out preference: %s", cudaGetErrorString(e));
        configured = smem;
    }
    b200_launch_pdl(0, kern, dim3(2 * a.tiles), dim3(chain_threads(eh)), smem, static_cast<cudaStream_t>(stream), a);
    B200_LAUNCH_CHECK("ppo_chain");
    return B200GYM_OK;
}

/* debug: registers (or clears, with NULL) a device buffer of 6 * 256 * 2 uint64 that the first and the last CTA of ppo_chain_kernel
 * fill with (event code, clock64) pairs — tools/trace_chain.py prints the per-step budget. */
extern "C" int b200gym_debug_chain_trace(void* buf) {
#ifdef B200GYM_CHAIN_TRACE
    unsigned long long* p = static_cast<unsigned long long*>(buf);
    cudaError_t e = cudaMemcpyToSymbol(g_chain_trace, &p, sizeof(p));
    B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "debug_chain_trace: %s", cudaGetErrorString(e));
    return B200GYM_OK;
#else
    (void)buf;
    B200_REQUIRE(false, B200GYM_EINVAL, "debug_chain_trace: this library was built without -DB200GYM_CHAIN_TRACE (tools/trace_chain.py builds its own)");
    return B200GYM_EINVAL;
#endif
}
