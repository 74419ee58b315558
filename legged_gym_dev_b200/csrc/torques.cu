// Group R torque kernels.
//
//   pd_torques_kernel     LeggedRobot._compute_torques           legged_robot.py:389-413 (+ action clip :86-87)
//   lstm_torques_kernel   Anymal._compute_torques (LSTM branch)  anymal.py:71-78 + LSTMsea TorchScript
//   (post_physics_kernel lives in post_physics.cu)
//
// Both are elementwise over the flattened (env,dof) index and stream through HBM with 128-bit accesses.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "philox.cuh"
#include "../../include/b200gym.h"

#ifndef PD_MAX_CTAS
#define PD_MAX_CTAS 0      // 0 = one CTA per 256 float4 chunks (no cap); see profiles/ for the A/B of capped grids
#endif
#ifndef LSTM_MINBLOCKS
#define LSTM_MINBLOCKS 8
#endif

namespace {

constexpr int ND = B200GYM_NUM_DOF;

// ------------------------------------------------------------------------------------------------
// PD torques: purely elementwise over N*12; one thread = 4 consecutive (env,dof) entries.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pd_torques_kernel(const __grid_constant__ B200LeggedParams p,
                                                         const float4* __restrict__ actions,
                                                         float4* __restrict__ actions_clipped,
                                                         const float4* __restrict__ dof_state,
                                                         const float4* __restrict__ last_dof_vel,
                                                         float4* __restrict__ torques, int n4) {
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    // Per-dof constants as float4 rows for the three dof quads {0-3, 4-7, 8-11}: a thread's quad is i % 3, i.e. lane-dependent.
    // Indexing the __grid_constant__ arrays with it costs 16 constant-bank loads per thread, each replayed for the 3 distinct
    // addresses of a warp (the kernel was MIO-throttled at 79 % of HBM); four 128-bit shared-memory reads replace them.
    __shared__ float4 tab[4][3];
    if (threadIdx.x < 12) {
        const int k = threadIdx.x / 3, r = threadIdx.x % 3;
        const float* src = k == 0 ? p.p_gains : k == 1 ? p.d_gains : k == 2 ? p.default_dof_pos : p.torque_limits;
        tab[k][r] = make_float4(src[4 * r], src[4 * r + 1], src[4 * r + 2], src[4 * r + 3]);
    }
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();
    // grid-stride: large problems run on a capped grid (PD_MAX_CTAS) of long-lived CTAs instead of ~50 000 64-byte-per-thread CTAs
    for (int i = i0; i < n4; i += gridDim.x * blockDim.x) {
    float4 a4 = ldg_stream4(actions + i);
    const float4 s0 = ldg_stream4(dof_state + 2 * i), s1 = ldg_stream4(dof_state + 2 * i + 1);
    float a[4] = {a4.x, a4.y, a4.z, a4.w};
    const float q[4] = {s0.x, s0.z, s1.x, s1.z}, qd[4] = {s0.y, s0.w, s1.y, s1.w};
    float lv[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.control_type == 1) {
        const float4 l = ldg_stream4(last_dof_vel + i);
        lv[0] = l.x, lv[1] = l.y, lv[2] = l.z, lv[3] = l.w;
    }
    const int r = i % 3;   // dof quad: 12 % 4 == 0, so the 4 entries never straddle an env
    const float4 kp4 = tab[0][r], kd4 = tab[1][r], q04 = tab[2][r], tl4 = tab[3][r];
    const float kp[4] = {kp4.x, kp4.y, kp4.z, kp4.w}, kd[4] = {kd4.x, kd4.y, kd4.z, kd4.w};
    const float q0[4] = {q04.x, q04.y, q04.z, q04.w}, tl[4] = {tl4.x, tl4.y, tl4.z, tl4.w};
    float t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a[j] = clampf(a[j], -p.clip_actions, p.clip_actions);
        const float as = mul_rn(a[j], p.action_scale);
        float v;
        if (p.control_type == 0)
            v = sub_rn(mul_rn(kp[j], sub_rn(add_rn(as, q0[j]), q[j])), mul_rn(kd[j], qd[j]));
        else if (p.control_type == 1)
            v = sub_rn(mul_rn(kp[j], sub_rn(as, qd[j])), div_rn(mul_rn(kd[j], sub_rn(qd[j], lv[j])), p.sim_dt));
        else
            v = as;
        t[j] = clampf(v, -tl[j], tl[j]);
    }
    stg_stream4(torques + i, make_float4(t[0], t[1], t[2], t[3]));
    if (actions_clipped) actions_clipped[i] = make_float4(a[0], a[1], a[2], a[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// Actuator LSTM: one thread per actuator (env,dof).  969 weights live in constant memory, so every FFMA
// takes its weight operand straight from the constant bank (all lanes read the same address).
// ------------------------------------------------------------------------------------------------
// Gate rows are stored in PAIRS (2p, 2p+1) so that one packed fp32 FMA (fma.rn.f32x2, SASS FFMA2 — Blackwell issues two
// fp32 FMAs per lane per instruction) advances two gate pre-activations; the weight pair comes from the constant bank /
// uniform registers and the input is the broadcast scalar operand, so no packing moves are needed.
struct ActuatorNet {
    float2 w_ih0[16][2], w_hh0[16][8], b0[16], bh0[16];   // b0 = b_ih, bh0 = b_hh (summed separately to mirror ATen)
    float2 w_ih1[16][8], w_hh1[16][8], b1[16], bh1[16];
    float w_lin[8], b_lin, in0, in1, out_scale;
};
__constant__ ActuatorNet c_net;

// MUFU-based activations without the denormal-handling prologues of __expf/__fdividef: ex2.approx.ftz (rel. error 2^-22)
// and rcp.approx.ftz (1 ulp).  Saturation is exact: ex2 -> inf gives rcp -> 0, ex2 -> 0 (flushed) gives rcp(1) = 1.
// Absolute error ~1e-7 on sigma/tanh, inside the 1e-5 * S contract (ATen's own two LSTM paths differ by 1.1e-5).
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Gate non-linearities of one PAIR of hidden units with shared reciprocals.  With E_x = exp(-x) (E_g, E_c: exp(-2x)):
//   sigma(f) c + sigma(i) tanh(g) = [c (1+E_i)(1+E_g) + (1-E_g)(1+E_f)] / [(1+E_f)(1+E_i)(1+E_g)]        one rcp instead of three
//   sigma(o) tanh(c')             = (1-E_c) / [(1+E_o)(1+E_c)]                                           one rcp instead of two
// i.e. 7 MUFU ops per unit instead of 10 (the XU pipe was the busiest pipe of this kernel, profiles/r1_kernels_ncu.md);
// the surrounding arithmetic runs two units per instruction (mul/add/fma.rn.f32x2).  The exponent arguments are capped at
// 2^40 so that the triple product stays finite: sigma(x) for x < -27.7 becomes 9e-13 instead of a smaller number.
__device__ __forceinline__ float2 ex2_pair_capped(float2 x, float scale) {
    const float2 a = __fmul2_rn(x, make_float2(scale, scale));
    return make_float2(ex2_ftz(fminf(a.x, 40.0f)), ex2_ftz(fminf(a.y, 40.0f)));
}
__device__ __forceinline__ float2 rcp_pair(float2 x) { return make_float2(rcp_ftz(x.x), rcp_ftz(x.y)); }

// gate[p] = pre-activations of gate rows 2p, 2p+1: i (p 0-3), f (4-7), g (8-11), o (12-15) of hidden units 2(p%4), 2(p%4)+1
__device__ __forceinline__ void lstm_gates(const float2 (&gate)[16], float (&h)[8], float (&c)[8]) {
    const float L = 1.4426950408889634f;
    const float2 one = make_float2(1.0f, 1.0f), mone = make_float2(-1.0f, -1.0f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 Ei = ex2_pair_capped(gate[k], -L), Ef = ex2_pair_capped(gate[4 + k], -L);
        const float2 Eg = ex2_pair_capped(gate[8 + k], -2.0f * L), Eo = ex2_pair_capped(gate[12 + k], -L);
        const float2 D1 = __fadd2_rn(Ef, one), D2 = __fmul2_rn(__fadd2_rn(Ei, one), __fadd2_rn(Eg, one));
        const float2 N2 = __ffma2_rn(Eg, mone, one);
        const float2 cc = make_float2(c[2 * k], c[2 * k + 1]);
        const float2 num = __ffma2_rn(N2, D1, __fmul2_rn(cc, D2));
        const float2 cn = __fmul2_rn(num, rcp_pair(__fmul2_rn(D1, D2)));
        const float2 Ec = ex2_pair_capped(cn, -2.0f * L);
        const float2 hn = __fmul2_rn(__ffma2_rn(Ec, mone, one), rcp_pair(__fmul2_rn(__fadd2_rn(Eo, one), __fadd2_rn(Ec, one))));
        c[2 * k] = cn.x, c[2 * k + 1] = cn.y;
        h[2 * k] = hn.x, h[2 * k + 1] = hn.y;
    }
}

template <int NIN>
__device__ __forceinline__ void lstm_cell(const float2 (&w_ih)[16][NIN], const float2 (&w_hh)[16][8], const float2 (&bi)[16],
                                          const float2 (&bh)[16], const float (&x)[NIN], float (&h)[8], float (&c)[8]) {
    float2 gate[16];   // pair p = gate rows 2p, 2p+1: i (p 0-3), f (4-7), g (8-11), o (12-15) of hidden units 2(p%4), 2(p%4)+1
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        float2 a = bi[p], bsum = bh[p];
#pragma unroll
        for (int k = 0; k < NIN; ++k) a = __ffma2_rn(w_ih[p][k], make_float2(x[k], x[k]), a);
#pragma unroll
        for (int k = 0; k < 8; ++k) bsum = __ffma2_rn(w_hh[p][k], make_float2(h[k], h[k]), bsum);
        gate[p] = __fadd2_rn(a, bsum);
    }
    lstm_gates(gate, h, c);
}

// The same cell with scalar fma: each weight is a constant-bank OPERAND of its FFMA, where the packed form needs an LDC.64 into registers
// first (ncu at 4096 envs: 22 % issue utilisation, short-scoreboard waits on those loads); here ptxas fetches the weights four at a time.
// Twice the FMA-pipe issues, but measured faster at every size (profiles/r2_lstm_scalar_ab.txt), so it is the default; the packed
// form stays selectable (B200GYM_LSTM_VARIANT=3).  Same products, same rounding: bit-identical.
template <int NIN>
__device__ __forceinline__ void lstm_cell_scalar(const float2 (&w_ih)[16][NIN], const float2 (&w_hh)[16][8], const float2 (&bi)[16],
                                                 const float2 (&bh)[16], const float (&x)[NIN], float (&h)[8], float (&c)[8]) {
    float2 gate[16];
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        float ax = bi[p].x, ay = bi[p].y, bx = bh[p].x, by = bh[p].y;
#pragma unroll
        for (int k = 0; k < NIN; ++k) ax = __fmaf_rn(w_ih[p][k].x, x[k], ax), ay = __fmaf_rn(w_ih[p][k].y, x[k], ay);
#pragma unroll
        for (int k = 0; k < 8; ++k) bx = __fmaf_rn(w_hh[p][k].x, h[k], bx), by = __fmaf_rn(w_hh[p][k].y, h[k], by);
        gate[p] = make_float2(__fadd_rn(ax, bx), __fadd_rn(ay, by));
    }
    lstm_gates(gate, h, c);
}

template <bool SCALAR>
__global__ void __launch_bounds__(128, LSTM_MINBLOCKS) lstm_torques_kernel(const __grid_constant__ B200LeggedParams p,
                                                           const float* __restrict__ actions,
                                                           float* __restrict__ actions_clipped,
                                                           const float2* __restrict__ dof_state, float* __restrict__ hbuf,
                                                           float* __restrict__ cbuf, float* __restrict__ torques, int m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (i >= m) return;
    const int d = i % ND;
    const float a = clampf(actions[i], -p.clip_actions, p.clip_actions);
    const float2 s = dof_state[i];
    float x[2];
    x[0] = sub_rn(add_rn(mul_rn(a, p.action_scale), p.default_dof_pos[d]), s.x) * c_net.in0;
    x[1] = s.y * c_net.in1;
    float h0[8], c0[8], h1[8], c1[8];
    const size_t l1 = static_cast<size_t>(m) * 8;
    const float4* hp = reinterpret_cast<const float4*>(hbuf + static_cast<size_t>(i) * 8);
    const float4* cp = reinterpret_cast<const float4*>(cbuf + static_cast<size_t>(i) * 8);
    const float4* hp1 = reinterpret_cast<const float4*>(hbuf + l1 + static_cast<size_t>(i) * 8);
    const float4* cp1 = reinterpret_cast<const float4*>(cbuf + l1 + static_cast<size_t>(i) * 8);
    float4 v;
    // each thread owns 32 contiguous bytes per state row: two 128-bit loads that share their sectors, so they must be
    // allowed to allocate in L1 (a no-allocate streaming load would fetch every sector twice from L2)
#define LD8(dst, src)                                                  \
    v = __ldg(src);                                                    \
    dst[0] = v.x, dst[1] = v.y, dst[2] = v.z, dst[3] = v.w;            \
    v = __ldg(src + 1);                                                \
    dst[4] = v.x, dst[5] = v.y, dst[6] = v.z, dst[7] = v.w;
    LD8(h0, hp) LD8(c0, cp) LD8(h1, hp1) LD8(c1, cp1)
#undef LD8
    if (SCALAR) {
        lstm_cell_scalar<2>(c_net.w_ih0, c_net.w_hh0, c_net.b0, c_net.bh0, x, h0, c0);
        lstm_cell_scalar<8>(c_net.w_ih1, c_net.w_hh1, c_net.b1, c_net.bh1, h0, h1, c1);
    } else {
        lstm_cell<2>(c_net.w_ih0, c_net.w_hh0, c_net.b0, c_net.bh0, x, h0, c0);
        lstm_cell<8>(c_net.w_ih1, c_net.w_hh1, c_net.b1, c_net.bh1, h0, h1, c1);
    }
    float o = c_net.b_lin;
#pragma unroll
    for (int k = 0; k < 8; ++k) o = fmaf(c_net.w_lin[k], h1[k], o);
    torques[i] = c_net.out_scale * o;
    if (actions_clipped) actions_clipped[i] = a;
#define ST8(dstp, src)                                                                                 \
    stg_stream4(const_cast<float4*>(dstp), make_float4(src[0], src[1], src[2], src[3]));              \
    stg_stream4(const_cast<float4*>(dstp) + 1, make_float4(src[4], src[5], src[6], src[7]));
    ST8(hp, h0) ST8(cp, c0) ST8(hp1, h1) ST8(cp1, c1)
#undef ST8
}


// ------------------------------------------------------------------------------------------------
// lstm_torques_tc_kernel — the same actuator LSTM with the gate mat-vecs on the tcgen05 tensor cores.
// 128 actuators = one UMMA tile (M = 128 rows, N = 32 gate rows); per layer D[128 x 32] = A[128 x K] W^T accumulates in TMEM.
// The torque contract is 1e-5 (S = 80), so plain TF32 (10-bit mantissa) is not enough: every operand is split into
// hi = x with the low 13 mantissa bits cleared (exactly representable in TF32) and lo = x - hi (exact in fp32), and the
// product is evaluated as  A_hi W_hi + A_lo W_hi + A_hi W_lo  — ONE chain of K' = 3K TF32 MMAs with A' = [A_hi | A_lo | A_hi],
// W' = [W_hi | W_hi | W_lo] (the dropped lo*lo term is 2^-22 relative).  fp32 accumulation; biases, the shared-reciprocal
// gate non-linearities and the output layer stay in fp32 in the epilogue.  Per actuator this removes 444 FFMA2 from the FMA
// pipe (the kernel was issue / FMA / MUFU co-limited at 70 % of HBM, profiles/r1_kernels_ncu.md).
//   layer 1: K = 10 (x0, x1, h0[8])  -> K' = 30, padded to 32 (4 MMAs of K = 8)
//   layer 2: K = 16 (h0'[8], h1[8])  -> K' = 48            (6 MMAs)
// One 128-thread CTA = one tile at a time (thread = actuator = TMEM lane), several CTAs per SM hide each other's latencies.
// ------------------------------------------------------------------------------------------------
#ifndef LSTM_TC_BLOCKS
#define LSTM_TC_BLOCKS 4
#endif
constexpr int LSTM_K1 = 32, LSTM_K2 = 48, LSTM_NG = 32;
__device__ float g_lstm_wtc[(LSTM_K1 + LSTM_K2) * LSTM_NG];   // [K'/4][32][4] per layer, written by b200gym_set_actuator_net
__device__ float g_lstm_btc[2 * LSTM_NG];                      // b_ih + b_hh per layer

__device__ __forceinline__ uint64_t tc_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (static_cast<uint64_t>(smem_u32(smem) >> 4) & 0x3FFFull) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
           (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool tc_elect_one() {
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
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
        "%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// writes [hi(v) | lo(v) | hi(v)] (3K floats, zero padded to KP) as this thread's row of the K-major A operand
template <int K, int KP>
__device__ __forceinline__ void stage_split_row(float* __restrict__ sA, int row, const float (&v)[K]) {
    float a[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) a[j] = 0.0f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const float hi = tf32_hi(v[j]);
        a[j] = hi, a[K + j] = v[j] - hi, a[2 * K + j] = hi;
    }
#pragma unroll
    for (int cidx = 0; cidx < KP / 4; ++cidx)
        *reinterpret_cast<float4*>(sA + (static_cast<size_t>(cidx) * 128 + row) * 4) =
            make_float4(a[4 * cidx], a[4 * cidx + 1], a[4 * cidx + 2], a[4 * cidx + 3]);
}

__global__ void __launch_bounds__(128, LSTM_TC_BLOCKS) lstm_torques_tc_kernel(const __grid_constant__ B200LeggedParams p, const float* __restrict__ actions,
                                                                  float* __restrict__ actions_clipped, const float2* __restrict__ dof_state,
                                                                  float* __restrict__ hbuf, float* __restrict__ cbuf,
                                                                  float* __restrict__ torques, int m) {
    __shared__ __align__(128) float sA[LSTM_K2 * 128];                       // A operand of the current layer: [K'/4][128][4]
    __shared__ __align__(128) float sW[(LSTM_K1 + LSTM_K2) * LSTM_NG];       // [K'/4][32][4] per layer
    __shared__ float sBias[2 * LSTM_NG];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    pdl_launch_dependents();
    pdl_wait();
    for (int i = tid; i < (LSTM_K1 + LSTM_K2) * LSTM_NG; i += 128) sW[i] = g_lstm_wtc[i];
    if (tid < 2 * LSTM_NG) sBias[tid] = g_lstm_btc[tid];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    fence_proxy_async();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t my_taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    // instruction descriptor: c F32 [4,6)=1, a/b TF32 [7,10)/[10,13)=2, K-major, N>>3 [17,23), M>>4 [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(LSTM_NG >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    uint32_t phase = 0;
    const size_t l1 = static_cast<size_t>(m) * 8;
    const int ntiles = (m + 127) / 128;

    auto run_mma = [&](const float* sWl, int ksteps) {
        // every thread has staged its A row (generic proxy -> async proxy) and finished reading the accumulator
        fence_proxy_async();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tc_elect_one()) {
                for (int ks = 0; ks < ksteps; ++ks) {   // UMMA K = 8 tf32 = two 16-byte chunks
                    const uint64_t da = tc_desc(sA + static_cast<size_t>(2 * ks) * 128 * 4, 128 * 16, 128);
                    const uint64_t db = tc_desc(sWl + static_cast<size_t>(2 * ks) * LSTM_NG * 4, LSTM_NG * 16, 128);
                    tc_mma_tf32(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = tile * 128 + tid;
        const bool live = i < m;
        const int ii = live ? i : 0;
        const int d = ii % ND;
        const float a = clampf(live ? actions[ii] : 0.0f, -p.clip_actions, p.clip_actions);
        const float2 s = live ? dof_state[ii] : make_float2(0.f, 0.f);
        float h0[8], c0[8], h1[8], c1[8];
        float4* hp = reinterpret_cast<float4*>(hbuf + static_cast<size_t>(ii) * 8);
        float4* cp = reinterpret_cast<float4*>(cbuf + static_cast<size_t>(ii) * 8);
        float4* hp1 = reinterpret_cast<float4*>(hbuf + l1 + static_cast<size_t>(ii) * 8);
        float4* cp1 = reinterpret_cast<float4*>(cbuf + l1 + static_cast<size_t>(ii) * 8);
        float4 v;
#define LD8(dst, src)                                                  \
    v = __ldg(src);                                                    \
    dst[0] = v.x, dst[1] = v.y, dst[2] = v.z, dst[3] = v.w;            \
    v = __ldg(src + 1);                                                \
    dst[4] = v.x, dst[5] = v.y, dst[6] = v.z, dst[7] = v.w;
        LD8(h0, hp) LD8(c0, cp) LD8(h1, hp1) LD8(c1, cp1)
#undef LD8
        float in1[10];
        in1[0] = sub_rn(add_rn(mul_rn(a, p.action_scale), p.default_dof_pos[d]), s.x) * c_net.in0;
        in1[1] = s.y * c_net.in1;
#pragma unroll
        for (int k = 0; k < 8; ++k) in1[2 + k] = h0[k];
        stage_split_row<10, LSTM_K1>(sA, tid, in1);
        run_mma(sW, LSTM_K1 / 8);
        float g[32];
        float2 gate[16];
        tc_ld32(my_taddr, g);
#pragma unroll
        for (int k = 0; k < 16; ++k) gate[k] = make_float2(g[2 * k] + sBias[2 * k], g[2 * k + 1] + sBias[2 * k + 1]);
        lstm_gates(gate, h0, c0);
        float in2[16];
#pragma unroll
        for (int k = 0; k < 8; ++k) in2[k] = h0[k], in2[8 + k] = h1[k];
        stage_split_row<16, LSTM_K2>(sA, tid, in2);
        run_mma(sW + LSTM_K1 * LSTM_NG, LSTM_K2 / 8);
        tc_ld32(my_taddr, g);
#pragma unroll
        for (int k = 0; k < 16; ++k) gate[k] = make_float2(g[2 * k] + sBias[LSTM_NG + 2 * k], g[2 * k + 1] + sBias[LSTM_NG + 2 * k + 1]);
        lstm_gates(gate, h1, c1);
        if (live) {
            float o = c_net.b_lin;
#pragma unroll
            for (int k = 0; k < 8; ++k) o = fmaf(c_net.w_lin[k], h1[k], o);
            torques[i] = c_net.out_scale * o;
            if (actions_clipped) actions_clipped[i] = a;
#define ST8(dstp, src)                                                            \
    stg_stream4(dstp, make_float4(src[0], src[1], src[2], src[3]));              \
    stg_stream4(dstp + 1, make_float4(src[4], src[5], src[6], src[7]));
            ST8(hp, h0) ST8(cp, c0) ST8(hp1, h1) ST8(cp1, c1)
#undef ST8
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int b200gym_pd_torques(const B200LeggedParams* p, const float* actions, float* actions_clipped, const float* dof_state,
                       const float* last_dof_vel, float* torques, void* stream) {
    B200_REQUIRE(p && actions && dof_state && torques, B200GYM_EINVAL, "pd_torques: null argument");
    B200_REQUIRE(p->num_envs > 0, B200GYM_EINVAL, "pd_torques: num_envs must be positive (got %d)", p->num_envs);
    B200_REQUIRE(p->control_type >= 0 && p->control_type <= 2, B200GYM_EINVAL, "Unknown controller type: %d", p->control_type);
    B200_REQUIRE(p->control_type != 1 || last_dof_vel, B200GYM_EINVAL, "pd_torques: control_type V needs last_dof_vel");
    B200_REQUIRE(b200_aligned16(actions) && b200_aligned16(dof_state) && b200_aligned16(torques) &&
                     b200_aligned16(actions_clipped) && b200_aligned16(last_dof_vel),
                 B200GYM_EALIGN, "pd_torques: pointers must be 16-byte aligned");
    const int n4 = p->num_envs * ND / 4;
    static int max_ctas = -1;
    if (max_ctas < 0) {
        const char* e = getenv("B200GYM_PD_MAX_CTAS");
        max_ctas = e ? atoi(e) : PD_MAX_CTAS;
    }
    const int want = (n4 + 255) / 256, grid = (max_ctas > 0 && want > max_ctas) ? max_ctas : want;
    b200_launch_pdl(p->num_envs, pd_torques_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), *p,
                    reinterpret_cast<const float4*>(actions), reinterpret_cast<float4*>(actions_clipped),
                    reinterpret_cast<const float4*>(dof_state), reinterpret_cast<const float4*>(last_dof_vel),
                    reinterpret_cast<float4*>(torques), n4);
    B200_LAUNCH_CHECK("pd_torques");
    return B200GYM_OK;
}

int b200gym_set_actuator_net(const float* w_ih0, const float* w_hh0, const float* b_ih0, const float* b_hh0,
                             const float* w_ih1, const float* w_hh1, const float* b_ih1, const float* b_hh1,
                             const float* w_lin, const float* b_lin, float in_scale0, float in_scale1, float out_scale) {
    B200_REQUIRE(w_ih0 && w_hh0 && b_ih0 && b_hh0 && w_ih1 && w_hh1 && b_ih1 && b_hh1 && w_lin && b_lin, B200GYM_EINVAL,
                 "set_actuator_net: null argument");
    ActuatorNet n;
    for (int p = 0; p < 16; ++p) {   // rows 2p and 2p+1 side by side
        const int r0 = 2 * p, r1 = 2 * p + 1;
        for (int k = 0; k < 2; ++k) n.w_ih0[p][k] = make_float2(w_ih0[r0 * 2 + k], w_ih0[r1 * 2 + k]);
        for (int k = 0; k < 8; ++k) {
            n.w_hh0[p][k] = make_float2(w_hh0[r0 * 8 + k], w_hh0[r1 * 8 + k]);
            n.w_ih1[p][k] = make_float2(w_ih1[r0 * 8 + k], w_ih1[r1 * 8 + k]);
            n.w_hh1[p][k] = make_float2(w_hh1[r0 * 8 + k], w_hh1[r1 * 8 + k]);
        }
        n.b0[p] = make_float2(b_ih0[r0], b_ih0[r1]), n.bh0[p] = make_float2(b_hh0[r0], b_hh0[r1]);
        n.b1[p] = make_float2(b_ih1[r0], b_ih1[r1]), n.bh1[p] = make_float2(b_hh1[r0], b_hh1[r1]);
    }
    for (int k = 0; k < 8; ++k) n.w_lin[k] = w_lin[k];
    n.b_lin = b_lin[0], n.in0 = in_scale0, n.in1 = in_scale1, n.out_scale = out_scale;
    cudaError_t e = cudaMemcpyToSymbol(c_net, &n, sizeof(n));
    B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "set_actuator_net: %s", cudaGetErrorString(e));
    // tensor-core path: W' = [W_hi | W_hi | W_lo] per layer in the K-major chunk layout [K'/4][32][4]; hi = low 13 mantissa bits cleared
    static float wtc[(LSTM_K1 + LSTM_K2) * LSTM_NG], btc[2 * LSTM_NG];
    for (float& x : wtc) x = 0.0f;
    auto hi = [](float x) {
        uint32_t u;
        memcpy(&u, &x, 4);
        u &= 0xFFFFE000u;
        float r;
        memcpy(&r, &u, 4);
        return r;
    };
    auto put = [&](int layer_off, int kp, int row, float val) { wtc[layer_off + ((kp / 4) * LSTM_NG + row) * 4 + (kp % 4)] = val; };
    for (int r = 0; r < LSTM_NG; ++r) {
        float w1[10], w2[16];
        for (int k = 0; k < 2; ++k) w1[k] = w_ih0[r * 2 + k];
        for (int k = 0; k < 8; ++k) w1[2 + k] = w_hh0[r * 8 + k], w2[k] = w_ih1[r * 8 + k], w2[8 + k] = w_hh1[r * 8 + k];
        for (int k = 0; k < 10; ++k) {
            put(0, k, r, hi(w1[k])), put(0, 10 + k, r, hi(w1[k])), put(0, 20 + k, r, w1[k] - hi(w1[k]));
        }
        for (int k = 0; k < 16; ++k) {
            put(LSTM_K1 * LSTM_NG, k, r, hi(w2[k])), put(LSTM_K1 * LSTM_NG, 16 + k, r, hi(w2[k])), put(LSTM_K1 * LSTM_NG, 32 + k, r, w2[k] - hi(w2[k]));
        }
        btc[r] = b_ih0[r] + b_hh0[r];
        btc[LSTM_NG + r] = b_ih1[r] + b_hh1[r];
    }
    e = cudaMemcpyToSymbol(g_lstm_wtc, wtc, sizeof(wtc));
    B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "set_actuator_net: %s", cudaGetErrorString(e));
    e = cudaMemcpyToSymbol(g_lstm_btc, btc, sizeof(btc));
    B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "set_actuator_net: %s", cudaGetErrorString(e));
    return B200GYM_OK;
}

static int g_lstm_variant = -1;
/* debug / A-B aid: 0 = scalar-fma kernel (default), 1 = tcgen05 3xTF32 kernel, 3 = packed FFMA2 kernel, -1 = re-read B200GYM_LSTM_VARIANT */
int b200gym_debug_set_lstm_variant(int variant) {
    g_lstm_variant = variant;
    return B200GYM_OK;
}

int b200gym_lstm_torques(const B200LeggedParams* p, const float* actions, float* actions_clipped, const float* dof_state,
                         float* h, float* c, float* torques, void* stream) {
    B200_REQUIRE(p && actions && dof_state && h && c && torques, B200GYM_EINVAL, "lstm_torques: null argument");
    B200_REQUIRE(p->num_envs > 0, B200GYM_EINVAL, "lstm_torques: num_envs must be positive (got %d)", p->num_envs);
    B200_REQUIRE(b200_aligned16(h) && b200_aligned16(c) && b200_aligned16(dof_state), B200GYM_EALIGN,
                 "lstm_torques: state pointers must be 16-byte aligned");
    const int m = p->num_envs * ND;
    if (g_lstm_variant < 0) {   // 0 (default): scalar fma; 1: tcgen05 (3xTF32) gate mat-vecs — measured slower, see profiles/; 3: the packed
        const char* e = getenv("B200GYM_LSTM_VARIANT");   // FFMA2 form of the default scalar-fma kernel
        g_lstm_variant = e ? atoi(e) : 0;
    }
    if (g_lstm_variant == 1) {
        static int sms = 0;
        if (!sms) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
        const int ntiles = (m + 127) / 128, cap = sms * LSTM_TC_BLOCKS;
        b200_launch_pdl(p->num_envs, lstm_torques_tc_kernel, dim3(ntiles < cap ? ntiles : cap), dim3(128), 0, static_cast<cudaStream_t>(stream), *p, actions,
                        actions_clipped, reinterpret_cast<const float2*>(dof_state), h, c, torques, m);
        B200_LAUNCH_CHECK("lstm_torques (tcgen05)");
        return B200GYM_OK;
    }
    if (g_lstm_variant != 3)   // scalar fma: 4-8 % faster than the packed form from 16 384 to 1 M envs, equal at 4096 (profiles/r2_lstm_scalar_ab.txt)
        b200_launch_pdl(p->num_envs, lstm_torques_kernel<true>, dim3((m + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), *p, actions,
                        actions_clipped, reinterpret_cast<const float2*>(dof_state), h, c, torques, m);
    else
        b200_launch_pdl(p->num_envs, lstm_torques_kernel<false>, dim3((m + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), *p, actions,
                        actions_clipped, reinterpret_cast<const float2*>(dof_state), h, c, torques, m);
    B200_LAUNCH_CHECK("lstm_torques");
    return B200GYM_OK;
}

}  // extern "C"
