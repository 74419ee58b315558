// Group M kernels: batched reduced-order-model rollout + tracking controller.
//
//   rom_init_kernel      TrajectoryGenerator.__init__ draw                 trajopt/rom_dynamics.py:495
//   rom_step_kernel      CustomSim.step / TrajectoryGenerator.step         custom_sim.py:71-75, rom_dynamics.py:568-590,607-612
//   rom_reset_kernel     CustomSim.reset_idx + TrajectoryGenerator.reset_idx   custom_sim.py:80-93, rom_dynamics.py:595-605
//   rom_policy_kernel    DoubleSingleTracking.__call__                     deep_tube_learning/controllers.py:87-92
//   rom_rollout_kernel   one data-collection epoch, persistent            data_collection_trajectory.py:104-149
//
// Envs are independent: one thread owns one env.  The per-call kernels keep the reference's [N, ...] tensors as
// the state of record (so traj_gen.k / .t / .trajectory ... stay readable attributes) and exist to preserve the
// call-per-step API; the rollout kernel is the throughput path: it holds the whole generator state (incl. the
// 11x2 horizon window) in registers for T ROM steps and only materialises the logs, staged through shared
// memory so that each env's [t, t+8) rows leave the SM as contiguous segments.
//
// fp32 discipline: t/k/t_final arithmetic and everything feeding `t >= k*dt - 1e-5` and `t > t_final` uses
// explicit round-to-nearest single operations in the reference's order, so the step and
// resample masks are bit-exact; DoubleInt2D's A@x row uses the fused multiply-add the reference's sgemm uses.
#include "common.cuh"
#include "philox.cuh"
#include "../../include/b200gym.h"

namespace {

constexpr int WMAXR = B200GYM_ROM_MAX_WINDOW;

template <int RN, int W>   // RN = rom state dim (2|4), W = window capacity
struct Gen {
    float traj[(W + 1) * RN], vtraj[W * 2];
    float w[4], t_final, t, k, hold[2], ext[2], ramp_t0, rv0[2], rv1[2], smag[2], sfreq[2], soff[2], smean[2], v[2];
    float cen[2];   // CircleTrajectoryGenerator.center (rom_dynamics.py:680-683)
    bool stat;
    uint32_t ctr;
};

template <int RN>
__device__ __forceinline__ void rom_bounds(const B200RomParams& p, const float* z, float (&lo)[2], float (&hi)[2]) {
    // RomDynamics.compute_state_dependent_input_bounds: rom_dynamics.py:106-107 (SingleInt2D), :234-246 (DoubleInt2D)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (RN == 2) {
            lo[j] = p.rom_v_min[j], hi[j] = p.rom_v_max[j];
        } else {
            hi[j] = fminf(p.rom_v_max[j], div_rn(sub_rn(p.rom_z_max[2 + j], z[2 + j]), p.rom_dt));
            lo[j] = fmaxf(p.rom_v_min[j], div_rn(sub_rn(p.rom_z_min[2 + j], z[2 + j]), p.rom_dt));
        }
    }
}
template <int RN>
__device__ __forceinline__ float rom_clip(float v, float lo, float hi) {   // clip_v_z: :201-202 identity, :248-250
    return RN == 2 ? v : fmaxf(fminf(v, hi), lo);
}

template <int RN, int W>
__device__ __forceinline__ void resample(const B200RomParams& p, Gen<RN, W>& g, const float* z, uint64_t genv) {
    // TrajectoryGenerator.resample, rom_dynamics.py:510-545; draw order = SURVEY.md A.3
    const philox::Stream rng(p.seed_lo, p.seed_hi, genv, g.ctr);
    float lo[2], hi[2];
    rom_bounds<RN>(p, z, lo, hi);
    const uint4 wc = rng.words(philox::ROM_CONST, 0), wr = rng.words(philox::ROM_RAMP, 0), we = rng.words(philox::ROM_EXTREME, 0);
    const uint4 wm = rng.words(philox::ROM_SIN_MAG, 0), wn = rng.words(philox::ROM_SIN_MEAN, 0);
    const uint4 wf = rng.words(philox::ROM_SIN_FREQ, 0), wo = rng.words(philox::ROM_SIN_OFF, 0);
    const float pi = 3.14159265358979323846f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float span = sub_rn(hi[j], lo[j]);
        g.hold[j] = affine_rn(span, philox::u01(philox::word(wc, j)), lo[j]);
        g.rv0[j] = rom_clip<RN>(g.rv1[j], lo[j], hi[j]);
        g.rv1[j] = affine_rn(span, philox::u01(philox::word(wr, j)), lo[j]);
        const uint32_t c = philox::bounded(philox::word(we, j), 3u);
        g.ext[j] = c == 0 ? lo[j] : (c == 1 ? 0.0f : hi[j]);
        g.smag[j] = mul_rn(div_rn(span, 2.0f), philox::u01(philox::word(wm, j)));
        const float mlo = add_rn(lo[j], g.smag[j]), mhi = sub_rn(hi[j], g.smag[j]);
        g.smean[j] = affine_rn(sub_rn(mhi, mlo), philox::u01(philox::word(wn, j)), mlo);
        g.sfreq[j] = affine_rn(sub_rn(p.freq_high, p.freq_low), philox::u01(philox::word(wf, j)), p.freq_low);
        g.soff[j] = affine_rn(sub_rn(pi, -pi), philox::u01(philox::word(wo, j)), -pi);
    }
    g.ramp_t0 = g.t_final;
    g.t_final = add_rn(g.t_final, affine_rn(p.t_span, philox::u01(rng.words(philox::ROM_TFINAL, 0).x), p.t_low));
    const float4 uw = philox::u01(rng.words(philox::ROM_WEIGHTS, 0));
    const int zc = p.weight_zero_col;
    const float w0 = zc == 0 ? 0.0f : uw.x, w1 = zc == 1 ? 0.0f : uw.y, w2 = zc == 2 ? 0.0f : uw.z, w3 = zc == 3 ? 0.0f : uw.w;
    const float sum = add_rn(add_rn(add_rn(w0, w1), w2), w3);
    g.w[0] = div_rn(w0, sum), g.w[1] = div_rn(w1, sum), g.w[2] = div_rn(w2, sum), g.w[3] = div_rn(w3, sum);
    g.stat = philox::u01(rng.words(philox::ROM_STATIONARY, 0).x) < p.prob_stationary;
    g.ctr += 1;
}

template <int RN, int W>
__device__ __forceinline__ void get_input(const B200RomParams& p, Gen<RN, W>& g, int w, uint64_t genv) {
    // TrajectoryGenerator.get_input_t, rom_dynamics.py:550-566 (+ v[stationary] = 0 of :580)
    if (W != WMAXR) w = W;
    const float* z = g.traj + w * RN;
    if (p.gen_kind != B200GYM_GEN_RANDOM) {
        // Zero / Square / Circle generators override get_input_t (rom_dynamics.py:618-698): no resampling, no clip_v_z, no draws
        float v0 = 0.0f, v1 = 0.0f;
        if (p.gen_kind == B200GYM_GEN_SQUARE) {   // :630-640 (SingleInt2D): four legs selected by the env's own clock
            const float t = g.t;
            if (0.0f <= t && t < p.gen_c[0]) v1 = p.gen_v[0];
            if (p.gen_c[0] <= t && t < p.gen_c[1]) v0 = p.gen_v[1];
            if (p.gen_c[1] <= t && t < p.gen_c[2]) v1 = p.gen_v[2];
            if (p.gen_c[2] <= t && t < p.gen_c[3]) v0 = p.gen_v[3];
        } else if (p.gen_kind == B200GYM_GEN_CIRCLE) {   // :686-692 (SingleInt2D)
            const float e0 = sub_rn(z[0], g.cen[0]), e1 = sub_rn(z[1], g.cen[1]);
            float a0 = -e1, a1 = e0;
            const float n1 = sqrtf(add_rn(mul_rn(a0, a0), mul_rn(a1, a1)));
            a0 = add_rn(a0, -sub_rn(e0, div_rn(mul_rn(0.5f, e0), n1)));
            a1 = add_rn(a1, -sub_rn(e1, div_rn(mul_rn(0.5f, e1), n1)));
            const float n2 = sqrtf(add_rn(mul_rn(a0, a0), mul_rn(a1, a1)));
            v0 = mul_rn(div_rn(a0, n2), p.gen_v[0]);
            v1 = mul_rn(div_rn(a1, n2), p.gen_v[0]);
        }
        g.v[0] = g.stat ? 0.0f : v0;
        g.v[1] = g.stat ? 0.0f : v1;
        return;
    }
    if (g.t > g.t_final) resample(p, g, z, genv);
    float lo[2], hi[2];
    rom_bounds<RN>(p, z, lo, hi);
    const float frac = div_rn(sub_rn(g.t, g.ramp_t0), sub_rn(g.t_final, g.ramp_t0));
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float ramp = add_rn(g.rv0[j], mul_rn(sub_rn(g.rv1[j], g.rv0[j]), frac));
        const float sn = add_rn(mul_rn(g.smag[j], sinf(add_rn(mul_rn(g.sfreq[j], g.t), g.soff[j]))), g.smean[j]);
        float a = mul_rn(g.w[0], rom_clip<RN>(g.hold[j], lo[j], hi[j]));
        a = add_rn(a, mul_rn(g.w[1], rom_clip<RN>(ramp, lo[j], hi[j])));
        a = add_rn(a, mul_rn(g.w[2], rom_clip<RN>(g.ext[j], lo[j], hi[j])));
        a = add_rn(a, mul_rn(g.w[3], rom_clip<RN>(sn, lo[j], hi[j])));
        g.v[j] = g.stat ? 0.0f : a;
    }
}

template <int N_>
__device__ __forceinline__ void int2d_f(float dt, const float* x, const float* u, float* out) {
    // SingleInt2D.f (:192-193) / DoubleInt2D.f (:224-225): (A @ x.T).T + (B @ u.T).T
    if (N_ == 2) {
        out[0] = add_rn(x[0], mul_rn(dt, u[0]));
        out[1] = add_rn(x[1], mul_rn(dt, u[1]));
    } else {
        const float a0 = fmaf(dt, x[2], x[0]), a1 = fmaf(dt, x[3], x[1]);
        out[2] = add_rn(x[2], mul_rn(dt, u[0]));
        out[3] = add_rn(x[3], mul_rn(dt, u[1]));
        out[0] = a0, out[1] = a1;
    }
}

template <int RN, int W>
__device__ __forceinline__ void advance(const B200RomParams& p, Gen<RN, W>& g, int w, bool inc_rom_time) {
    // TrajectoryGenerator.step_rom_idx after the input evaluation, rom_dynamics.py:581-590
    if (W != WMAXR) w = W;
    float zn[RN];
    int2d_f<RN>(p.rom_dt, g.traj + w * RN, g.v, zn);
    if (RN == 4 && g.stat) zn[2] = 0.0f, zn[3] = 0.0f;
    if (W == WMAXR) {
        for (int i = 0; i < w * RN; ++i) g.traj[i] = g.traj[i + RN];
        for (int i = 0; i < (w - 1) * 2; ++i) g.vtraj[i] = g.vtraj[i + 2];
    } else {
#pragma unroll
        for (int i = 0; i < W * RN; ++i) g.traj[i] = g.traj[i + RN];
#pragma unroll
        for (int i = 0; i < (W - 1) * 2; ++i) g.vtraj[i] = g.vtraj[i + 2];
    }
#pragma unroll
    for (int c = 0; c < RN; ++c) g.traj[w * RN + c] = zn[c];
    g.vtraj[(w - 1) * 2] = g.v[0], g.vtraj[(w - 1) * 2 + 1] = g.v[1];
    g.k = add_rn(g.k, 1.0f);
    if (inc_rom_time) g.t = add_rn(g.t, p.rom_dt);
}

template <int RN, int W>
__device__ __forceinline__ void gen_step(const B200RomParams& p, Gen<RN, W>& g, int w, uint64_t genv, bool in_idx) {
    // TrajectoryGenerator.step_idx, rom_dynamics.py:571-575.  The input (and a possible resample) is evaluated for
    // every env on every call (A.4), only envs of `idx` whose ROM clock is due advance.
    const bool due = in_idx && (g.t >= sub_rn(mul_rn(g.k, p.rom_dt), 1e-5f));
    get_input(p, g, w, genv);
    if (due) advance(p, g, w, false);
    if (in_idx) g.t = add_rn(g.t, p.dt_loop);
}

template <int RN, int W>
__device__ __forceinline__ float interp_scale(const B200RomParams& p, const Gen<RN, W>& g) {
    return sub_rn(g.t, mul_rn(sub_rn(g.k, 1.0f), p.rom_dt));   // (t - (k-1)*dt), rom_dynamics.py:611
}
__device__ __forceinline__ float interp(float a, float b, float s, float dt) { return add_rn(a, div_rn(mul_rn(sub_rn(b, a), s), dt)); }

__device__ __forceinline__ void tracking_policy(const B200RomParams& p, const float* x, const float* zt, const float* vt, float* u) {
    // DoubleSingleTracking (controllers.py:87-92) + DoubleInt2D.clip_v_z of the MODEL (rom_dynamics.py:234-250)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float a = add_rn(mul_rn(p.Kp, sub_rn(zt[j], x[j])), mul_rn(p.Kd, sub_rn(vt[j], x[2 + j])));
        const float hi = fminf(p.model_v_max[j], div_rn(sub_rn(p.model_z_max[2 + j], x[2 + j]), p.model_dt));
        const float lo = fmaxf(p.model_v_min[j], div_rn(sub_rn(p.model_z_min[2 + j], x[2 + j]), p.model_dt));
        u[j] = fmaxf(fminf(a, hi), lo);
    }
}

// ---- state <-> registers ---------------------------------------------------------------------------
template <int RN, int W, bool WINDOWS = true>
__device__ __forceinline__ void load_gen(const B200RomState& s, size_t i, int w, Gen<RN, W>& g) {
    const float* tr = s.trajectory + i * (w + 1) * RN;
    const float* vt = s.v_trajectory + i * w * 2;
    if (!WINDOWS) {
        // generator parameters only (the caller neither reads nor advances the horizon windows)
    } else if (W == WMAXR) {
        for (int c = 0; c < (w + 1) * RN; ++c) g.traj[c] = tr[c];
        for (int c = 0; c < w * 2; ++c) g.vtraj[c] = vt[c];
    } else {
#pragma unroll
        for (int c = 0; c < (W + 1) * RN; c += 2) *reinterpret_cast<float2*>(g.traj + c) = *reinterpret_cast<const float2*>(tr + c);
#pragma unroll
        for (int c = 0; c < W * 2; c += 2) *reinterpret_cast<float2*>(g.vtraj + c) = *reinterpret_cast<const float2*>(vt + c);
    }
    const float4 w4 = *reinterpret_cast<const float4*>(s.weights + i * 4);
    g.w[0] = w4.x, g.w[1] = w4.y, g.w[2] = w4.z, g.w[3] = w4.w;
    g.t_final = s.t_final[i], g.t = s.t[i], g.k = s.k[i], g.ramp_t0 = s.ramp_t_start[i];
#define LD2(dst, src)                                                    \
    {                                                                    \
        const float2 q = *reinterpret_cast<const float2*>(src + i * 2); \
        dst[0] = q.x, dst[1] = q.y;                                      \
    }
    LD2(g.hold, s.sample_hold_input) LD2(g.ext, s.extreme_input) LD2(g.rv0, s.ramp_v_start) LD2(g.rv1, s.ramp_v_end)
    LD2(g.smag, s.sin_mag) LD2(g.sfreq, s.sin_freq) LD2(g.soff, s.sin_off) LD2(g.smean, s.sin_mean) LD2(g.v, s.v)
#undef LD2
    g.stat = s.stationary_inds[i] != 0;
    g.ctr = static_cast<uint32_t>(s.rng_ctr[i]);
    g.cen[0] = g.cen[1] = 0.0f;
    if (s.center) {
        const float2 c = *reinterpret_cast<const float2*>(s.center + i * 2);
        g.cen[0] = c.x, g.cen[1] = c.y;
    }
}

template <int RN, int W, bool WINDOWS = true>
__device__ __forceinline__ void store_gen(const B200RomState& s, size_t i, int w, const Gen<RN, W>& g) {
    float* tr = s.trajectory + i * (w + 1) * RN;
    float* vt = s.v_trajectory + i * w * 2;
    if (!WINDOWS) {
    } else if (W == WMAXR) {
        for (int c = 0; c < (w + 1) * RN; ++c) tr[c] = g.traj[c];
        for (int c = 0; c < w * 2; ++c) vt[c] = g.vtraj[c];
    } else {
#pragma unroll
        for (int c = 0; c < (W + 1) * RN; c += 2) *reinterpret_cast<float2*>(tr + c) = make_float2(g.traj[c], g.traj[c + 1]);
#pragma unroll
        for (int c = 0; c < W * 2; c += 2) *reinterpret_cast<float2*>(vt + c) = make_float2(g.vtraj[c], g.vtraj[c + 1]);
    }
    *reinterpret_cast<float4*>(s.weights + i * 4) = make_float4(g.w[0], g.w[1], g.w[2], g.w[3]);
    s.t_final[i] = g.t_final, s.t[i] = g.t, s.k[i] = g.k, s.ramp_t_start[i] = g.ramp_t0;
#define ST2(dst, src) *reinterpret_cast<float2*>(dst + i * 2) = make_float2(src[0], src[1]);
    ST2(s.sample_hold_input, g.hold) ST2(s.extreme_input, g.ext) ST2(s.ramp_v_start, g.rv0) ST2(s.ramp_v_end, g.rv1)
    ST2(s.sin_mag, g.smag) ST2(s.sin_freq, g.sfreq) ST2(s.sin_off, g.soff) ST2(s.sin_mean, g.smean) ST2(s.v, g.v)
#undef ST2
    s.stationary_inds[i] = g.stat ? 1 : 0;
    s.rng_ctr[i] = static_cast<int32_t>(g.ctr);
    if (s.center) *reinterpret_cast<float2*>(s.center + i * 2) = make_float2(g.cen[0], g.cen[1]);
}

// writes CustomSim.trajectory (interpolated window, custom_sim.py:74) and the observation (custom_sim.py:95-100)
template <int RN, int W>
__device__ __forceinline__ void write_views(const B200RomParams& p, const B200RomState& s, size_t i, const Gen<RN, W>& g,
                                            const float* root, int mn) {
    const float sc = interp_scale(p, g);
    if (s.env_trajectory) {
        float* et = s.env_trajectory + i * p.horizon * RN;
        if (W != WMAXR) {   // specialised instances are only dispatched for dN == 1: all indices compile-time
#pragma unroll
            for (int c = 0; c < W * RN; ++c) et[c] = interp(g.traj[c], g.traj[c + RN], sc, p.rom_dt);
        } else {
            for (int j = 0; j < p.horizon; ++j) {
                const int a = j * p.dN * RN;
#pragma unroll
                for (int c = 0; c < RN; ++c) et[j * RN + c] = interp(g.traj[a + c], g.traj[a + RN + c], sc, p.rom_dt);
            }
        }
    }
    if (s.obs && root) {
        float* o = s.obs + i * (mn + RN + 2);
        for (int c = 0; c < mn; ++c) o[c] = root[c];
#pragma unroll
        for (int c = 0; c < RN; ++c) o[mn + c] = interp(g.traj[c], g.traj[RN + c], sc, p.rom_dt);
        o[mn + RN] = g.vtraj[2], o[mn + RN + 1] = g.vtraj[3];
    }
}

__global__ void rom_init_kernel(const __grid_constant__ B200RomParams p, const __grid_constant__ B200RomState s, long long env_off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.num_envs) return;
    uint32_t ctr = static_cast<uint32_t>(s.rng_ctr[i]);
    const philox::Stream rng(p.seed_lo, p.seed_hi, static_cast<uint64_t>(env_off + i), ctr);
    const uint4 w = rng.words(philox::ROM_INIT, 0);
    s.ramp_v_end[i * 2 + 0] = affine_rn(sub_rn(p.rom_v_max[0], p.rom_v_min[0]), philox::u01(w.x), p.rom_v_min[0]);
    s.ramp_v_end[i * 2 + 1] = affine_rn(sub_rn(p.rom_v_max[1], p.rom_v_min[1]), philox::u01(w.y), p.rom_v_min[1]);
    s.rng_ctr[i] = static_cast<int32_t>(ctr + 1);
}

template <int RN, int W>
#ifndef ROM_STEP_MINBLOCKS
#define ROM_STEP_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(128, ROM_STEP_MINBLOCKS) rom_step_kernel(const __grid_constant__ B200RomParams p, const __grid_constant__ B200RomState s,
                                                       const float* __restrict__ action, const uint8_t* __restrict__ mask,
                                                       long long env_off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (i >= p.num_envs) return;
    const int w = p.window, mn = p.model_type ? 4 : 2;
    Gen<RN, W> g;
    load_gen(s, i, w, g);
    float root[4] = {0.f, 0.f, 0.f, 0.f};
    if (action) {   // CustomSim.step: root_states = model.f(root_states, action)
        float x[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < mn; ++c) x[c] = s.root_states[static_cast<size_t>(i) * mn + c];
        const float u[2] = {action[i * 2], action[i * 2 + 1]};
        if (mn == 4) int2d_f<4>(p.model_dt, x, u, root);
        else int2d_f<2>(p.model_dt, x, u, root);
        for (int c = 0; c < mn; ++c) s.root_states[static_cast<size_t>(i) * mn + c] = root[c];
    }
    const uint32_t ctr0 = g.ctr;
    const float k0 = g.k;
    gen_step(p, g, w, static_cast<uint64_t>(env_off + i), mask ? mask[i] != 0 : true);
    // The ROM clock is slower than the loop (rom.dt / dt_loop loop steps per knot) and resamples are rarer still: write back only
    // what this call changed — the horizon windows when a knot was appended, the parameters when a resample fired, else t and v.
    if (g.k != k0) {
        store_gen(s, i, w, g);
    } else if (g.ctr != ctr0) {
        store_gen<RN, W, false>(s, i, w, g);
    } else {
        s.t[i] = g.t;
        *reinterpret_cast<float2*>(s.v + static_cast<size_t>(i) * 2) = make_float2(g.v[0], g.v[1]);
    }
    write_views(p, s, i, g, action ? root : nullptr, mn);
}

// reset_traj + TrajectoryGenerator.reset_idx for one env held in registers (custom_sim.py:80-85 == legged_robot_trajectory.py:248-253,
// rom_dynamics.py:595-605); pz = proj_z of the (new) root state; `in_idx` = env is being reset
template <int RN, int W>
__device__ __forceinline__ void traj_reset(const B200RomParams& p, Gen<RN, W>& g, float* pz, int w, uint64_t genv, bool in_idx) {
    if (W != WMAXR) w = W;
    if (in_idx) {
        {
            const philox::Stream rng(p.seed_lo, p.seed_hi, genv, g.ctr);
            if (p.randomize_rom_distance && philox::u01(rng.words(philox::ROM_DIST_MASK, 0).x) > p.zero_rom_dist_llh) {
                const uint4 wd = rng.words(philox::ROM_DIST, 0);
#pragma unroll
                for (int c = 0; c < RN; ++c) {
                    const float d = p.max_rom_distance[c];
                    pz[c] = add_rn(pz[c], affine_rn(sub_rn(d, -d), philox::u01(philox::word(wd, c)), -d));
                }
            }
            g.ctr += 1;
        }
        // TrajectoryGenerator.reset_idx, rom_dynamics.py:595-602
        for (int c = 0; c < w * RN; ++c) g.traj[c] = 0.0f;
        for (int c = 0; c < w * 2; ++c) g.vtraj[c] = 0.0f;
#pragma unroll
        for (int c = 0; c < RN; ++c) g.traj[w * RN + c] = pz[c];
        g.k = -static_cast<float>(w);
        g.t = mul_rn(g.k, p.rom_dt);
        g.t_final = g.t;
        if (p.gen_kind == B200GYM_GEN_RANDOM) resample(p, g, pz, genv);
        else if (p.gen_kind == B200GYM_GEN_ZERO) g.stat = true;                    // ZeroTrajectoryGenerator.resample, :619-620
    }
    if (p.gen_kind == B200GYM_GEN_CIRCLE) {   // CircleTrajectoryGenerator.resample re-centres EVERY env from z (:680-683), reset or not
        g.cen[0] = sub_rn(pz[0], 0.5f);
        g.cen[1] = pz[1];
    }
    // warm-up, :604-605: the input is evaluated for every env, only reset envs advance (with the ROM clock)
    for (int it = 0; it < w; ++it) {
        get_input(p, g, w, genv);
        if (in_idx) advance(p, g, w, true);
    }
}

// CustomSim.reset_idx body for one env held in registers; `in_idx` = env is being reset
template <int RN, int W>
__device__ __forceinline__ void sim_reset(const B200RomParams& p, Gen<RN, W>& g, float* root, int mn, int w, uint64_t genv, bool in_idx) {
    float pz[RN];
#pragma unroll
    for (int c = 0; c < RN; ++c) pz[c] = 0.0f;
    if (in_idx) {
        {   // root_states[idx] = U(lower, upper), custom_sim.py:88-91
            const philox::Stream rng(p.seed_lo, p.seed_hi, genv, g.ctr);
            const float4 ur = philox::u01(rng.words(philox::ROM_ROOT, 0));
            root[0] = affine_rn(sub_rn(p.noise_upper[0], p.noise_lower[0]), ur.x, p.noise_lower[0]);
            root[1] = affine_rn(sub_rn(p.noise_upper[1], p.noise_lower[1]), ur.y, p.noise_lower[1]);
            if (mn == 4) {
                root[2] = affine_rn(sub_rn(p.noise_upper[2], p.noise_lower[2]), ur.z, p.noise_lower[2]);
                root[3] = affine_rn(sub_rn(p.noise_upper[3], p.noise_lower[3]), ur.w, p.noise_lower[3]);
            }
            g.ctr += 1;
        }
        // reset_traj, custom_sim.py:80-85 (proj_z of SingleInt2D = first two model states)
#pragma unroll
        for (int c = 0; c < RN; ++c) pz[c] = c < 2 ? root[c] : 0.0f;
    }
    traj_reset(p, g, pz, w, genv, in_idx);
}

// LeggedRobotTrajectory.reset_traj (legged_robot_trajectory.py:248-253): generator-only reset of the envs flagged in `mask`, p_zx read
// from the robot's root_states [N, stride] (SingleInt2D.proj_z = first two columns).  The reference returns from reset_idx before
// touching the generator when no env resets (:217-218), so nothing happens when *any_reset == 0; otherwise the warm-up evaluates the
// input (incl. due resamples) of EVERY env, as the reference's step_rom_idx does (rom_dynamics.py:577-580).
template <int RN, int W>
__global__ void __launch_bounds__(128) rom_reset_root_kernel(const __grid_constant__ B200RomParams p, const __grid_constant__ B200RomState s,
                                                             const uint8_t* __restrict__ mask, const float* __restrict__ root, int stride,
                                                             const float* __restrict__ any_reset, long long env_off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (i >= p.num_envs) return;
    if (any_reset && *any_reset == 0.0f) return;
    const int w = p.window;
    const uint64_t genv = static_cast<uint64_t>(env_off + i);
    Gen<RN, W> g;
    if (mask[i] != 0) {
        load_gen(s, i, w, g);
        float pz[RN];
#pragma unroll
        for (int c = 0; c < RN; ++c) pz[c] = root[static_cast<size_t>(i) * stride + c];
        traj_reset(p, g, pz, w, genv, true);
        store_gen(s, i, w, g);
    }
}

// The envs that do NOT reset (the other half of the call above, as its own kernel: without the horizon windows in registers it needs
// a quarter of the registers and runs at full occupancy).
template <int RN, int W>
__global__ void __launch_bounds__(256, 4) rom_reset_root_others_kernel(const __grid_constant__ B200RomParams p, const __grid_constant__ B200RomState s,
                                                                    const uint8_t* __restrict__ mask, const float* __restrict__ root, int stride,
                                                                    const float* __restrict__ any_reset, long long env_off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (i >= p.num_envs) return;
    if (any_reset && *any_reset == 0.0f) return;
    if (mask[i] != 0) return;
    const int w = p.window;
    const uint64_t genv = static_cast<uint64_t>(env_off + i);
    Gen<RN, W> g;
    // An env that does not reset only takes part in the warm-up's input evaluations (rom_dynamics.py:577-580): resamples that are
    // due, then v.  Its clock does not move, so once t <= t_final every further evaluation repeats the same v; the horizon windows
    // are neither read (SingleInt2D bounds are state-independent) nor written: ~110 B instead of ~600 B of traffic for such an env.
    static_assert(RN == 2, "state-independent input bounds assumed");
#pragma unroll
    for (int c = 0; c < (W + 1) * RN; ++c) g.traj[c] = 0.0f;
    load_gen<RN, W, false>(s, i, w, g);
    if (p.gen_kind == B200GYM_GEN_CIRCLE) {   // re-centred from this env's own root position (no offset: it does not reset), :680-683
#pragma unroll
        for (int c = 0; c < RN; ++c) g.traj[(W != WMAXR ? W : w) * RN + c] = s.trajectory[(static_cast<size_t>(i) * (w + 1) + w) * RN + c];
        g.cen[0] = sub_rn(root[static_cast<size_t>(i) * stride], 0.5f);
        g.cen[1] = root[static_cast<size_t>(i) * stride + 1];
        *reinterpret_cast<float2*>(s.center + static_cast<size_t>(i) * 2) = make_float2(g.cen[0], g.cen[1]);
    }
    const uint32_t ctr0 = g.ctr;
    int it = 0;
    while (it < w && g.t > g.t_final) {
        get_input(p, g, w, genv);
        ++it;
    }
    if (it < w) get_input(p, g, w, genv);
    if (g.ctr != ctr0) store_gen<RN, W, false>(s, i, w, g);
    else *reinterpret_cast<float2*>(s.v + static_cast<size_t>(i) * 2) = make_float2(g.v[0], g.v[1]);
}

template <int RN, int W>
__global__ void __launch_bounds__(128) rom_reset_kernel(const __grid_constant__ B200RomParams p, const __grid_constant__ B200RomState s,
                                                        const uint8_t* __restrict__ mask, long long env_off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.num_envs) return;
    const int w = p.window, mn = p.model_type ? 4 : 2;
    const uint64_t genv = static_cast<uint64_t>(env_off + i);
    Gen<RN, W> g;
    load_gen(s, i, w, g);
    float root[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < mn; ++c) root[c] = s.root_states[static_cast<size_t>(i) * mn + c];
    sim_reset(p, g, root, mn, w, genv, mask ? mask[i] != 0 : true);
    // self.step(zeros) for ALL envs, custom_sim.py:93
    float nx[4] = {0.f, 0.f, 0.f, 0.f};
    const float u0[2] = {0.0f, 0.0f};
    if (mn == 4) int2d_f<4>(p.model_dt, root, u0, nx);
    else int2d_f<2>(p.model_dt, root, u0, nx);
    for (int c = 0; c < mn; ++c) s.root_states[static_cast<size_t>(i) * mn + c] = nx[c];
    gen_step(p, g, w, genv, true);
    store_gen(s, i, w, g);
    write_views(p, s, i, g, nx, mn);
}

__global__ void rom_policy_kernel(const __grid_constant__ B200RomParams p, const float* __restrict__ obs, float* __restrict__ action) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (i >= p.num_envs) return;
    const float4 a = *reinterpret_cast<const float4*>(obs + static_cast<size_t>(i) * 8);
    const float4 b = *reinterpret_cast<const float4*>(obs + static_cast<size_t>(i) * 8 + 4);
    const float x[4] = {a.x, a.y, a.z, a.w}, zt[2] = {b.x, b.y}, vt[2] = {b.z, b.w};
    float u[2];
    tracking_policy(p, x, zt, vt, u);
    *reinterpret_cast<float2*>(action + static_cast<size_t>(i) * 2) = make_float2(u[0], u[1]);
}

// RaibertHeuristic.raibert_policy (deep_tube_learning/controllers.py:38-73): position / velocity errors -> clamped pitch and roll
// commands -> desired orientation quaternion (w, x, y, z) at the robot's current yaw.  One thread per env; pure elementwise.
__global__ void raibert_policy_kernel(const float* __restrict__ obs, int stride, long long n, float Kp, float Kv, float Kff, float clip_pos,
                                      float clip_vel, float clip_ang, float* __restrict__ action) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (i >= n) return;
    const float* o = obs + i * stride;
    const float pex = o[0], pey = -o[1], evx = -o[2], evy = o[3], dvx = o[4], dvy = -o[5];
    const float pitch_pos = clampf(mul_rn(-Kp, pex), -clip_pos, clip_pos), roll_pos = clampf(mul_rn(-Kp, pey), -clip_pos, clip_pos);
    const float vx = clampf(add_rn(mul_rn(-Kv, evx), mul_rn(Kff, dvx)), -clip_vel, clip_vel);
    const float vy = clampf(add_rn(mul_rn(-Kv, evy), mul_rn(Kff, dvy)), -clip_vel, clip_vel);
    const float pitch = clampf(add_rn(pitch_pos, vx), -clip_ang, clip_ang), roll = clampf(add_rn(roll_pos, vy), -clip_ang, clip_ang);
    const float qx = o[6], qy = o[7], qz = o[8], qw = o[9];                                    // quat_to_yaw, :75-81
    const float yaw = atan2f(mul_rn(2.0f, add_rn(mul_rn(qw, qz), mul_rn(qx, qy))),
                             sub_rn(1.0f, mul_rn(2.0f, add_rn(mul_rn(qy, qy), mul_rn(qz, qz)))));
    const float cy = cosf(yaw * 0.5f), sy = sinf(yaw * 0.5f), cp = cosf(pitch * 0.5f), sp = sinf(pitch * 0.5f);   // omega_to_quat, :23-36
    const float cr = cosf(roll * 0.5f), sr = sinf(roll * 0.5f);
    const float w = add_rn(mul_rn(mul_rn(cr, cp), cy), mul_rn(mul_rn(sr, sp), sy));
    const float x = sub_rn(mul_rn(mul_rn(sr, cp), cy), mul_rn(mul_rn(cr, sp), sy));
    const float y = add_rn(mul_rn(mul_rn(cr, sp), cy), mul_rn(mul_rn(sr, cp), sy));
    const float z = sub_rn(mul_rn(mul_rn(cr, cp), sy), mul_rn(mul_rn(sr, sp), cy));
    *reinterpret_cast<float4*>(action + i * 4) = make_float4(w, x, y, z);
}

// ---- persistent rollout: one data-collection epoch ---------------------------------------------------
#ifndef ROLLOUT_MINBLOCKS
#define ROLLOUT_MINBLOCKS 4   // resident CTAs per SM the register allocation must allow (A/B: profiles/r2_rom_rollout_occupancy.txt)
#endif
constexpr int RB = 128;   // envs (threads) per CTA
constexpr int TB = 8;     // timesteps staged per flush

template <int E>
__device__ __forceinline__ void flush_rows(float* __restrict__ dst, const float* stage, int env0, int nenv, int rows_total, int r0, int cnt) {
    // stage layout [env][TB*E + 1]; dst layout [N][rows_total][E]; each env contributes one contiguous segment
    const int seg = cnt * E;
    for (int i = threadIdx.x; i < nenv * seg; i += RB) {
        const int e = i / seg, r = i - e * seg;
        dst[(static_cast<size_t>(env0 + e) * rows_total + r0) * E + r] = stage[e * (TB * E + 1) + r];
    }
}

template <int W>
__global__ void __launch_bounds__(RB, ROLLOUT_MINBLOCKS) rom_rollout_kernel(const __grid_constant__ B200RomParams p, const __grid_constant__ B200RomState s,
                                                         float* __restrict__ obs_io, int T, float* __restrict__ lx, float* __restrict__ lz,
                                                         float* __restrict__ lpz, float* __restrict__ lv, uint8_t* __restrict__ ldone,
                                                         long long env_off) {
    constexpr int RN = 2, MN = 4;   // SingleInt2D rom tracked by a DoubleInt2D model (double_single_int.yaml)
    __shared__ float st_z[RB * (TB * 2 + 1)], st_pz[RB * (TB * 2 + 1)], st_v[RB * (TB * 2 + 1)], st_x[RB * (TB * 4 + 1)];
    const int env0 = blockIdx.x * RB;
    const int nenv = min(RB, p.num_envs - env0);
    const int e = threadIdx.x;
    const bool valid = e < nenv;
    const size_t i = static_cast<size_t>(env0 + (valid ? e : 0));
    const uint64_t genv = static_cast<uint64_t>(env_off) + i;
    const int w = p.window;
    Gen<RN, W> g;
    load_gen(s, i, w, g);
    float root[MN], obs[8];
#pragma unroll
    for (int c = 0; c < MN; ++c) root[c] = s.root_states[i * MN + c];
    {
        const float4 a = *reinterpret_cast<const float4*>(obs_io + i * 8), b = *reinterpret_cast<const float4*>(obs_io + i * 8 + 4);
        obs[0] = a.x, obs[1] = a.y, obs[2] = a.z, obs[3] = a.w, obs[4] = b.x, obs[5] = b.y, obs[6] = b.z, obs[7] = b.w;
    }
    auto sim_step = [&](const float* u) {   // CustomSim.step + get_observations
        float nx[MN];
        int2d_f<MN>(p.model_dt, root, u, nx);
#pragma unroll
        for (int c = 0; c < MN; ++c) root[c] = nx[c];
        gen_step(p, g, w, genv, true);
        const float sc = interp_scale(p, g);
#pragma unroll
        for (int c = 0; c < MN; ++c) obs[c] = root[c];
        obs[4] = interp(g.traj[0], g.traj[2], sc, p.rom_dt), obs[5] = interp(g.traj[1], g.traj[3], sc, p.rom_dt);
        obs[6] = g.vtraj[2], obs[7] = g.vtraj[3];
    };
    // env.reset(): reset_idx(all) incl. the zero-action step (data_collection_trajectory.py:111, custom_sim.py:77-93)
    sim_reset(p, g, root, MN, w, genv, true);
    {
        float sv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) sv[c] = obs[c];   // the caller's obs is NOT refreshed by reset (:94,:111)
        const float u0[2] = {0.0f, 0.0f};
        sim_step(u0);
#pragma unroll
        for (int c = 0; c < 8; ++c) obs[c] = sv[c];
    }
    // row 0: x = state, pz_x = proj_z(state), z = RAW oldest window entry (:112-116)
    auto stage_row = [&](int slot, float z0, float z1) {
        st_z[e * (TB * 2 + 1) + slot * 2] = z0, st_z[e * (TB * 2 + 1) + slot * 2 + 1] = z1;
        st_pz[e * (TB * 2 + 1) + slot * 2] = root[0], st_pz[e * (TB * 2 + 1) + slot * 2 + 1] = root[1];
        if (lx) {
#pragma unroll
            for (int c = 0; c < MN; ++c) st_x[e * (TB * 4 + 1) + slot * 4 + c] = root[c];
        }
    };
    stage_row(0, g.traj[0], g.traj[1]);
    int r0 = 0;   // first staged row of x/z/pz
    for (int t = 0; t < T; ++t) {
        const float k0 = g.k;
        int guard = 0;
        do {   // step the environment until the ROM steps (:121-138)
            float u[2];
            tracking_policy(p, obs, obs + 4, obs + 6, u);
            sim_step(u);
        } while (g.k == k0 && ++guard < 4096);
        const float sc = interp_scale(p, g);
        st_v[e * (TB * 2 + 1) + (t % TB) * 2] = g.v[0], st_v[e * (TB * 2 + 1) + (t % TB) * 2 + 1] = g.v[1];
        stage_row((t + 1) % TB, interp(g.traj[0], g.traj[2], sc, p.rom_dt), interp(g.traj[1], g.traj[3], sc, p.rom_dt));
        const bool flush_v = ((t + 1) % TB == 0) || (t == T - 1);
        const bool flush_r = ((t + 2) % TB == 0) || (t == T - 1);
        if (flush_v || flush_r) __syncthreads();
        if (flush_v) {
            const int v0 = (t / TB) * TB, cnt = t - v0 + 1;
            flush_rows<2>(lv, st_v, env0, nenv, T, v0, cnt);
            for (int q = threadIdx.x; q < nenv * cnt; q += RB) ldone[static_cast<size_t>(env0 + q / cnt) * T + v0 + q % cnt] = 0;
        }
        if (flush_r) {
            const int cnt = t + 1 - r0 + 1;
            flush_rows<2>(lz, st_z, env0, nenv, T + 1, r0, cnt);
            flush_rows<2>(lpz, st_pz, env0, nenv, T + 1, r0, cnt);
            if (lx) flush_rows<4>(lx, st_x, env0, nenv, T + 1, r0, cnt);
            r0 = t + 2;
        }
        if (flush_v || flush_r) __syncthreads();
    }
    if (T == 0) {
        __syncthreads();
        flush_rows<2>(lz, st_z, env0, nenv, 1, 0, 1);
        flush_rows<2>(lpz, st_pz, env0, nenv, 1, 0, 1);
        if (lx) flush_rows<4>(lx, st_x, env0, nenv, 1, 0, 1);
    }
    if (valid) {
        store_gen(s, i, w, g);
#pragma unroll
        for (int c = 0; c < MN; ++c) s.root_states[i * MN + c] = root[c];
        *reinterpret_cast<float4*>(obs_io + i * 8) = make_float4(obs[0], obs[1], obs[2], obs[3]);
        *reinterpret_cast<float4*>(obs_io + i * 8 + 4) = make_float4(obs[4], obs[5], obs[6], obs[7]);
        write_views(p, s, i, g, root, MN);
    }
}

int check_rom(const B200RomParams* p, const B200RomState* s, const char* what, bool need_root) {
    B200_REQUIRE(p && s, B200GYM_EINVAL, "%s: null argument", what);
    B200_REQUIRE(p->num_envs > 0, B200GYM_EINVAL, "%s: num_envs must be positive (got %d)", what, p->num_envs);
    B200_REQUIRE(p->rom_type == 0 || p->rom_type == 1, B200GYM_EINVAL, "%s: unknown rom_type %d", what, p->rom_type);
    B200_REQUIRE(p->model_type == 0 || p->model_type == 1, B200GYM_EINVAL, "%s: unknown model_type %d", what, p->model_type);
    B200_REQUIRE(p->window >= 2 && p->window <= WMAXR && p->window == p->horizon * p->dN, B200GYM_EINVAL,
                 "%s: window %d must equal N*dN and lie in [2,%d]", what, p->window, WMAXR);
    const void* must[] = {s->trajectory, s->v_trajectory, s->v, s->t, s->k, s->t_final, s->weights, s->sample_hold_input,
                          s->extreme_input, s->ramp_v_start, s->ramp_v_end, s->ramp_t_start, s->sin_mag, s->sin_freq, s->sin_off,
                          s->sin_mean, s->stationary_inds, s->rng_ctr};
    for (const void* q : must) {
        B200_REQUIRE(q != nullptr, B200GYM_EINVAL, "%s: null state tensor", what);
        B200_REQUIRE(b200_aligned16(q), B200GYM_EALIGN, "%s: state tensors must be 16-byte aligned", what);
    }
    B200_REQUIRE(!need_root || s->root_states, B200GYM_EINVAL, "%s: root_states missing", what);
    B200_REQUIRE(p->gen_kind >= B200GYM_GEN_RANDOM && p->gen_kind <= B200GYM_GEN_CIRCLE, B200GYM_EINVAL, "%s: unknown gen_kind %d", what, p->gen_kind);
    B200_REQUIRE(p->gen_kind == B200GYM_GEN_RANDOM || p->rom_type == 0, B200GYM_EINVAL,
                 "%s: the Zero / Square / Circle generators are implemented for a SingleInt2D rom", what);
    B200_REQUIRE(p->gen_kind != B200GYM_GEN_CIRCLE || (s->center && b200_aligned16(s->center)), B200GYM_EINVAL, "%s: circle centres missing", what);
    return B200GYM_OK;
}

}  // namespace

extern "C" {

int b200gym_rom_init(const B200RomParams* p, const B200RomState* s, int64_t env_id_offset, void* stream) {
    if (int rc = check_rom(p, s, "rom_init", false)) return rc;
    rom_init_kernel<<<(p->num_envs + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p, *s, env_id_offset);
    B200_LAUNCH_CHECK("rom_init");
    return B200GYM_OK;
}

int b200gym_rom_step(const B200RomParams* p, const B200RomState* s, const float* action, const uint8_t* step_mask,
                     int64_t env_id_offset, void* stream) {
    if (int rc = check_rom(p, s, "rom_step", action != nullptr)) return rc;
    B200_REQUIRE(!action || p->rom_type == 0, B200GYM_EINVAL, "rom_step: CustomSim stepping needs a SingleInt2D rom");
    const int grid = (p->num_envs + 127) / 128;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p->rom_type == 0) {
        if (p->window == 10 && p->dN == 1) b200_launch_pdl(p->num_envs, rom_step_kernel<2, 10>, dim3(grid), dim3(128), 0, st, *p, *s, action, step_mask, env_id_offset);
        else b200_launch_pdl(p->num_envs, rom_step_kernel<2, WMAXR>, dim3(grid), dim3(128), 0, st, *p, *s, action, step_mask, env_id_offset);
    } else {
        if (p->window == 10 && p->dN == 1) b200_launch_pdl(p->num_envs, rom_step_kernel<4, 10>, dim3(grid), dim3(128), 0, st, *p, *s, action, step_mask, env_id_offset);
        else b200_launch_pdl(p->num_envs, rom_step_kernel<4, WMAXR>, dim3(grid), dim3(128), 0, st, *p, *s, action, step_mask, env_id_offset);
    }
    B200_LAUNCH_CHECK("rom_step");
    return B200GYM_OK;
}

int b200gym_rom_reset(const B200RomParams* p, const B200RomState* s, const uint8_t* reset_mask, int64_t env_id_offset,
                      void* stream) {
    if (int rc = check_rom(p, s, "rom_reset", true)) return rc;
    B200_REQUIRE(p->gen_kind == B200GYM_GEN_RANDOM, B200GYM_EINVAL, "rom_reset: CustomSim only knows the random TrajectoryGenerator (custom_sim.py:1,53)");
    B200_REQUIRE(p->rom_type == 0, B200GYM_EINVAL, "rom_reset: CustomSim needs a SingleInt2D rom (proj_z of the model state)");
    const int grid = (p->num_envs + 127) / 128;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p->window == 10 && p->dN == 1) rom_reset_kernel<2, 10><<<grid, 128, 0, st>>>(*p, *s, reset_mask, env_id_offset);
    else rom_reset_kernel<2, WMAXR><<<grid, 128, 0, st>>>(*p, *s, reset_mask, env_id_offset);
    B200_LAUNCH_CHECK("rom_reset");
    return B200GYM_OK;
}

int b200gym_rom_reset_from_root(const B200RomParams* p, const B200RomState* s, const uint8_t* reset_mask, const float* root,
                                int32_t root_stride, const float* any_reset, int64_t env_id_offset, void* stream) {
    if (int rc = check_rom(p, s, "rom_reset_from_root", false)) return rc;
    B200_REQUIRE(reset_mask && root && root_stride >= 2, B200GYM_EINVAL, "rom_reset_from_root: reset_mask / root missing");
    B200_REQUIRE(p->rom_type == 0, B200GYM_EINVAL, "rom_reset_from_root: proj_z is implemented for SingleInt2D (first two root columns)");
    const int grid = (p->num_envs + 127) / 128;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid2 = (p->num_envs + 255) / 256;
    if (p->window == 10 && p->dN == 1) {
        b200_launch_pdl(p->num_envs, rom_reset_root_others_kernel<2, 10>, dim3(grid2), dim3(256), 0, st, *p, *s, reset_mask, root, root_stride, any_reset, env_id_offset);
        b200_launch_pdl(p->num_envs, rom_reset_root_kernel<2, 10>, dim3(grid), dim3(128), 0, st, *p, *s, reset_mask, root, root_stride, any_reset, env_id_offset);
    } else {
        b200_launch_pdl(p->num_envs, rom_reset_root_others_kernel<2, WMAXR>, dim3(grid2), dim3(256), 0, st, *p, *s, reset_mask, root, root_stride, any_reset, env_id_offset);
        b200_launch_pdl(p->num_envs, rom_reset_root_kernel<2, WMAXR>, dim3(grid), dim3(128), 0, st, *p, *s, reset_mask, root, root_stride, any_reset, env_id_offset);
    }
    B200_LAUNCH_CHECK("rom_reset_from_root");
    return B200GYM_OK;
}

int b200gym_rom_tracking_policy(const B200RomParams* p, const float* obs, float* action, void* stream) {
    B200_REQUIRE(p && obs && action, B200GYM_EINVAL, "rom_tracking_policy: null argument");
    B200_REQUIRE(p->num_envs > 0, B200GYM_EINVAL, "rom_tracking_policy: num_envs must be positive");
    B200_REQUIRE(p->model_type == 1, B200GYM_EINVAL, "rom_tracking_policy: DoubleSingleTracking needs a DoubleInt2D model");
    B200_REQUIRE(b200_aligned16(obs) && b200_aligned16(action), B200GYM_EALIGN, "rom_tracking_policy: pointers must be 16-byte aligned");
    b200_launch_pdl(p->num_envs, rom_policy_kernel, dim3((p->num_envs + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream), *p, obs, action);
    B200_LAUNCH_CHECK("rom_tracking_policy");
    return B200GYM_OK;
}

int b200gym_raibert_policy(const float* obs, int32_t obs_stride, int64_t n, float Kp, float Kv, float Kff, float clip_pos, float clip_vel,
                           float clip_ang, float* action, void* stream) {
    B200_REQUIRE(obs && action && n > 0 && obs_stride >= 10, B200GYM_EINVAL, "raibert_policy: need obs [n, >= 10] and an action buffer");
    B200_REQUIRE(b200_aligned16(action), B200GYM_EALIGN, "raibert_policy: action must be 16-byte aligned");
    b200_launch_pdl(static_cast<int>(n < (1 << 30) ? n : (1 << 30)), raibert_policy_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0,
                    static_cast<cudaStream_t>(stream), obs, obs_stride, n, Kp, Kv, Kff, clip_pos, clip_vel, clip_ang, action);
    B200_LAUNCH_CHECK("raibert_policy");
    return B200GYM_OK;
}

int b200gym_rom_rollout(const B200RomParams* p, const B200RomState* s, float* obs_io, int32_t T, float* x, float* z, float* pz_x,
                        float* v, uint8_t* done, int64_t env_id_offset, void* stream) {
    if (int rc = check_rom(p, s, "rom_rollout", true)) return rc;
    B200_REQUIRE(p->gen_kind == B200GYM_GEN_RANDOM, B200GYM_EINVAL, "rom_rollout: CustomSim only knows the random TrajectoryGenerator (custom_sim.py:1,53)");
    B200_REQUIRE(p->rom_type == 0 && p->model_type == 1, B200GYM_EINVAL,
                 "rom_rollout: built for the double_single_int configuration (DoubleInt2D model, SingleInt2D rom)");
    B200_REQUIRE(obs_io && z && pz_x && T >= 0 && (T == 0 || (v && done)), B200GYM_EINVAL, "rom_rollout: null log buffer or negative T");
    B200_REQUIRE(b200_aligned16(obs_io), B200GYM_EALIGN, "rom_rollout: obs must be 16-byte aligned");
    const int grid = (p->num_envs + RB - 1) / RB;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p->window == 10 && p->dN == 1) rom_rollout_kernel<10><<<grid, RB, 0, st>>>(*p, *s, obs_io, T, x, z, pz_x, v, done, env_id_offset);
    else rom_rollout_kernel<WMAXR><<<grid, RB, 0, st>>>(*p, *s, obs_io, T, x, z, pz_x, v, done, env_id_offset);
    B200_LAUNCH_CHECK("rom_rollout");
    return B200GYM_OK;
}

}  // extern "C"
