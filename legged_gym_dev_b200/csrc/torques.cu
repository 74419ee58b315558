// Group R torque kernels.
//
//   pd_torques_kernel     LeggedRobot._compute_torques           legged_robot.py:389-413 (+ action clip :86-87)
//   lstm_torques_kernel   Anymal._compute_torques (LSTM branch)  anymal.py:71-78 + LSTMsea TorchScript
//   (post_physics_kernel lives in post_physics.cu)
//
// Both are elementwise over the flattened (env,dof) index and stream through HBM with 128-bit accesses.
#include "common.cuh"
#include "philox.cuh"
#include "../../include/b200gym.h"

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
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a4 = ldg_stream4(actions + i);
    const float4 s0 = ldg_stream4(dof_state + 2 * i), s1 = ldg_stream4(dof_state + 2 * i + 1);
    float a[4] = {a4.x, a4.y, a4.z, a4.w};
    const float q[4] = {s0.x, s0.z, s1.x, s1.z}, qd[4] = {s0.y, s0.w, s1.y, s1.w};
    float lv[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.control_type == 1) {
        const float4 l = ldg_stream4(last_dof_vel + i);
        lv[0] = l.x, lv[1] = l.y, lv[2] = l.z, lv[3] = l.w;
    }
    const int d0 = (i * 4) % ND;   // 12 % 4 == 0, so the 4 entries never straddle an env
    float t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int d = d0 + j;
        a[j] = clampf(a[j], -p.clip_actions, p.clip_actions);
        const float as = mul_rn(a[j], p.action_scale);
        float v;
        if (p.control_type == 0)
            v = sub_rn(mul_rn(p.p_gains[d], sub_rn(add_rn(as, p.default_dof_pos[d]), q[j])), mul_rn(p.d_gains[d], qd[j]));
        else if (p.control_type == 1)
            v = sub_rn(mul_rn(p.p_gains[d], sub_rn(as, qd[j])), div_rn(mul_rn(p.d_gains[d], sub_rn(qd[j], lv[j])), p.sim_dt));
        else
            v = as;
        t[j] = clampf(v, -p.torque_limits[d], p.torque_limits[d]);
    }
    stg_stream4(torques + i, make_float4(t[0], t[1], t[2], t[3]));
    if (actions_clipped) actions_clipped[i] = make_float4(a[0], a[1], a[2], a[3]);
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

__global__ void __launch_bounds__(128, LSTM_MINBLOCKS) lstm_torques_kernel(const __grid_constant__ B200LeggedParams p,
                                                           const float* __restrict__ actions,
                                                           float* __restrict__ actions_clipped,
                                                           const float2* __restrict__ dof_state, float* __restrict__ hbuf,
                                                           float* __restrict__ cbuf, float* __restrict__ torques, int m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
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
    lstm_cell<2>(c_net.w_ih0, c_net.w_hh0, c_net.b0, c_net.bh0, x, h0, c0);
    lstm_cell<8>(c_net.w_ih1, c_net.w_hh1, c_net.b1, c_net.bh1, h0, h1, c1);
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
    pd_torques_kernel<<<(n4 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        *p, reinterpret_cast<const float4*>(actions), reinterpret_cast<float4*>(actions_clipped),
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
    return B200GYM_OK;
}

int b200gym_lstm_torques(const B200LeggedParams* p, const float* actions, float* actions_clipped, const float* dof_state,
                         float* h, float* c, float* torques, void* stream) {
    B200_REQUIRE(p && actions && dof_state && h && c && torques, B200GYM_EINVAL, "lstm_torques: null argument");
    B200_REQUIRE(p->num_envs > 0, B200GYM_EINVAL, "lstm_torques: num_envs must be positive (got %d)", p->num_envs);
    B200_REQUIRE(b200_aligned16(h) && b200_aligned16(c) && b200_aligned16(dof_state), B200GYM_EALIGN,
                 "lstm_torques: state pointers must be 16-byte aligned");
    const int m = p->num_envs * ND;
    lstm_torques_kernel<<<(m + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        *p, actions, actions_clipped, reinterpret_cast<const float2*>(dof_state), h, c, torques, m);
    B200_LAUNCH_CHECK("lstm_torques");
    return B200GYM_OK;
}

}  // extern "C"
