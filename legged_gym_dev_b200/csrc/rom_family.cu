// Group M, SURVEY.md §8f row 4: the whole RomDynamics family behind one generic trajectory generator.
//
//   fam_f_kernel        RomDynamics.f                                    trajopt/rom_dynamics.py:192,224,273-278,311-316,345-352,406-414
//   fam_des_kernel      RomDynamics.des_pose_vel                         :198,230,286-290,318-322,354-358,416-420
//   fam_bounds_kernel   compute_state_dependent_input_bounds / clip_v_z  :106-107,201,234-250,292,367-383
//   fam_proj_kernel     RomDynamics.proj_z                               :195,227,280-284,360-365,422-427
//   fam_init_kernel     TrajectoryGenerator.__init__ draw                :495
//   fam_reset_kernel    TrajectoryGenerator.reset_idx(idx, z)            :595-605
//   fam_step_kernel     TrajectoryGenerator.step_idx(idx)                :571-590
//   fam_input_kernel    TrajectoryGenerator.get_input_t(t, z)            :560-566
//
// rom.cu holds the register-resident fast path of the shipped configurations (2-input SingleInt2D / DoubleInt2D with the
// horizon window in registers).  Here the rom class is a template parameter <T> (state dim up to 6, up to 3 inputs) and the
// horizon windows stay in HBM: one thread owns one env, a reset writes its W+1 knots straight to their final rows (the
// reference's W shift-and-append passes produce exactly rows 0..W = the W+1 successive states), a step shifts the window in
// place only when the env's ROM clock is due.  Draw events, sites and the fp32 operation order are those of rom.cu, so rom
// types 0 / 1 reproduce its results bit for bit (tests/test_rom_family_gpu.py).
#include "common.cuh"
#include "philox.cuh"
#include "../../include/b200gym.h"

namespace {

template <int T>
struct Rom;   // n = state dim, m = input dim, vs = first velocity state (== n when the class has none; then n - vs == m otherwise)
template <> struct Rom<B200GYM_ROM_SINGLE_INT_2D> { static constexpr int n = 2, m = 2, vs = 2; };
template <> struct Rom<B200GYM_ROM_DOUBLE_INT_2D> { static constexpr int n = 4, m = 2, vs = 2; };
template <> struct Rom<B200GYM_ROM_UNICYCLE> { static constexpr int n = 3, m = 2, vs = 3; };
template <> struct Rom<B200GYM_ROM_LATERAL_UNICYCLE> { static constexpr int n = 3, m = 3, vs = 3; };
template <> struct Rom<B200GYM_ROM_EXTENDED_UNICYCLE> { static constexpr int n = 5, m = 2, vs = 3; };
template <> struct Rom<B200GYM_ROM_EXTENDED_LATERAL_UNICYCLE> { static constexpr int n = 6, m = 3, vs = 3; };

// ---- the rom classes' algebra --------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ void rom_f(float dt, const float* x, const float* u, float* out) {
    if constexpr (T == B200GYM_ROM_SINGLE_INT_2D) {          // (A @ x.T).T + (B @ u.T).T, :192-193
        out[0] = add_rn(x[0], mul_rn(dt, u[0]));
        out[1] = add_rn(x[1], mul_rn(dt, u[1]));
    } else if constexpr (T == B200GYM_ROM_DOUBLE_INT_2D) {   // :224-225 (the A @ x row is the fused multiply-add of the reference's sgemm)
        const float a0 = fmaf(dt, x[2], x[0]), a1 = fmaf(dt, x[3], x[1]);
        out[2] = add_rn(x[2], mul_rn(dt, u[0]));
        out[3] = add_rn(x[3], mul_rn(dt, u[1]));
        out[0] = a0, out[1] = a1;
    } else {                                       // x + dt * gu with gu assembled column by column
        const float c = cosf(x[2]), s = sinf(x[2]);
        float gu[Rom<T>::n];
        if constexpr (T == B200GYM_ROM_UNICYCLE) {                   // :273-278
            gu[0] = mul_rn(u[0], c), gu[1] = mul_rn(u[0], s), gu[2] = u[1];
        } else if constexpr (T == B200GYM_ROM_LATERAL_UNICYCLE) {    // :311-316
            gu[0] = sub_rn(mul_rn(u[0], c), mul_rn(u[1], s));
            gu[1] = add_rn(mul_rn(u[0], s), mul_rn(u[1], c));
            gu[2] = u[2];
        } else if constexpr (T == B200GYM_ROM_EXTENDED_UNICYCLE) {   // :345-352
            gu[0] = mul_rn(x[3], c), gu[1] = mul_rn(x[3], s), gu[2] = x[4];
            gu[3] = u[0], gu[4] = u[1];
        } else {                                           // :406-414
            gu[0] = sub_rn(mul_rn(x[3], c), mul_rn(x[4], s));
            gu[1] = add_rn(mul_rn(x[3], s), mul_rn(x[4], c));
            gu[2] = x[5];
            gu[3] = u[0], gu[4] = u[1], gu[5] = u[2];
        }
#pragma unroll
        for (int i = 0; i < Rom<T>::n; ++i) out[i] = add_rn(x[i], mul_rn(dt, gu[i]));
    }
}

template <int T>
__device__ __forceinline__ void rom_bounds(const B200RomFamilyParams& p, const float* z, float* lo, float* hi) {
    constexpr int M = Rom<T>::m, VS = Rom<T>::vs;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        if constexpr (VS == Rom<T>::n) {   // no velocity states: the static bounds, :106-107
            lo[j] = p.v_min[j], hi[j] = p.v_max[j];
        } else {                 // :243-245, :376-378
            hi[j] = fminf(p.v_max[j], div_rn(sub_rn(p.z_max[VS + j], z[VS + j]), p.rom_dt));
            lo[j] = fmaxf(p.v_min[j], div_rn(sub_rn(p.z_min[VS + j], z[VS + j]), p.rom_dt));
        }
    }
}
template <int T>
__device__ __forceinline__ float rom_clip(float v, float lo, float hi) {   // clip_v_z: identity without velocity states
    return Rom<T>::vs == Rom<T>::n ? v : fmaxf(fminf(v, hi), lo);
}

// ---- generator parameters of one env (everything but the horizon windows) -------------------------------
template <int M>
struct GenP {
    float w[4], t_final, t, k, ramp_t0;
    float hold[M], ext[M], rv0[M], rv1[M], smag[M], sfreq[M], soff[M], smean[M], v[M];
    bool stat;
    uint32_t ctr;
};

template <int M>
__device__ __forceinline__ void load_params(const B200RomState& s, size_t i, GenP<M>& g) {
    const float4 w4 = *reinterpret_cast<const float4*>(s.weights + i * 4);
    g.w[0] = w4.x, g.w[1] = w4.y, g.w[2] = w4.z, g.w[3] = w4.w;
    g.t_final = s.t_final[i], g.t = s.t[i], g.k = s.k[i], g.ramp_t0 = s.ramp_t_start[i];
#pragma unroll
    for (int j = 0; j < M; ++j) {
        g.hold[j] = s.sample_hold_input[i * M + j], g.ext[j] = s.extreme_input[i * M + j];
        g.rv0[j] = s.ramp_v_start[i * M + j], g.rv1[j] = s.ramp_v_end[i * M + j];
        g.smag[j] = s.sin_mag[i * M + j], g.sfreq[j] = s.sin_freq[i * M + j];
        g.soff[j] = s.sin_off[i * M + j], g.smean[j] = s.sin_mean[i * M + j];
        g.v[j] = s.v[i * M + j];
    }
    g.stat = s.stationary_inds[i] != 0;
    g.ctr = static_cast<uint32_t>(s.rng_ctr[i]);
}

template <int M>
__device__ __forceinline__ void store_params(const B200RomState& s, size_t i, const GenP<M>& g) {
    *reinterpret_cast<float4*>(s.weights + i * 4) = make_float4(g.w[0], g.w[1], g.w[2], g.w[3]);
    s.t_final[i] = g.t_final, s.ramp_t_start[i] = g.ramp_t0;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        s.sample_hold_input[i * M + j] = g.hold[j], s.extreme_input[i * M + j] = g.ext[j];
        s.ramp_v_start[i * M + j] = g.rv0[j], s.ramp_v_end[i * M + j] = g.rv1[j];
        s.sin_mag[i * M + j] = g.smag[j], s.sin_freq[i * M + j] = g.sfreq[j];
        s.sin_off[i * M + j] = g.soff[j], s.sin_mean[i * M + j] = g.smean[j];
    }
    s.stationary_inds[i] = g.stat ? 1 : 0;
    s.rng_ctr[i] = static_cast<int32_t>(g.ctr);
}

template <int T>
__device__ __forceinline__ void resample(const B200RomFamilyParams& p, GenP<Rom<T>::m>& g, const float* z, uint64_t genv) {
    // TrajectoryGenerator.resample, rom_dynamics.py:510-545; one draw event, one Philox block per site (column j = word j)
    constexpr int M = Rom<T>::m;
    const philox::Stream rng(p.seed_lo, p.seed_hi, genv, g.ctr);
    float lo[M], hi[M];
    rom_bounds<T>(p, z, lo, hi);
    const uint4 wc = rng.words(philox::ROM_CONST, 0), wr = rng.words(philox::ROM_RAMP, 0), we = rng.words(philox::ROM_EXTREME, 0);
    const uint4 wm = rng.words(philox::ROM_SIN_MAG, 0), wn = rng.words(philox::ROM_SIN_MEAN, 0);
    const uint4 wf = rng.words(philox::ROM_SIN_FREQ, 0), wo = rng.words(philox::ROM_SIN_OFF, 0);
    const float pi = 3.14159265358979323846f;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        const float span = sub_rn(hi[j], lo[j]);
        g.hold[j] = affine_rn(span, philox::u01(philox::word(wc, j)), lo[j]);                       // :528
        g.rv0[j] = rom_clip<T>(g.rv1[j], lo[j], hi[j]);                                             // :531
        g.rv1[j] = affine_rn(span, philox::u01(philox::word(wr, j)), lo[j]);                        // :532
        const uint32_t c = philox::bounded(philox::word(we, j), 3u);                                // :535-538
        g.ext[j] = c == 0 ? lo[j] : (c == 1 ? 0.0f : hi[j]);
        g.smag[j] = mul_rn(div_rn(span, 2.0f), philox::u01(philox::word(wm, j)));                   // :541
        const float mlo = add_rn(lo[j], g.smag[j]), mhi = sub_rn(hi[j], g.smag[j]);
        g.smean[j] = affine_rn(sub_rn(mhi, mlo), philox::u01(philox::word(wn, j)), mlo);            // :542
        g.sfreq[j] = affine_rn(sub_rn(p.freq_high, p.freq_low), philox::u01(philox::word(wf, j)), p.freq_low);
        g.soff[j] = affine_rn(sub_rn(pi, -pi), philox::u01(philox::word(wo, j)), -pi);
    }
    g.ramp_t0 = g.t_final;                                                                          // :533
    g.t_final = add_rn(g.t_final, affine_rn(p.t_span, philox::u01(rng.words(philox::ROM_TFINAL, 0).x), p.t_low));
    const float4 uw = philox::u01(rng.words(philox::ROM_WEIGHTS, 0));
    const int zc = p.weight_zero_col;
    const float w0 = zc == 0 ? 0.0f : uw.x, w1 = zc == 1 ? 0.0f : uw.y, w2 = zc == 2 ? 0.0f : uw.z, w3 = zc == 3 ? 0.0f : uw.w;
    const float sum = add_rn(add_rn(add_rn(w0, w1), w2), w3);
    g.w[0] = div_rn(w0, sum), g.w[1] = div_rn(w1, sum), g.w[2] = div_rn(w2, sum), g.w[3] = div_rn(w3, sum);
    g.stat = philox::u01(rng.words(philox::ROM_STATIONARY, 0).x) < p.prob_stationary;               // :520
    g.ctr += 1;
}

// TrajectoryGenerator.get_input_t(t, z), :560-566: a resample when t > t_final, then the weighted sum of the four clipped inputs
template <int T>
__device__ __forceinline__ void input_t(const B200RomFamilyParams& p, GenP<Rom<T>::m>& g, float t, const float* z, uint64_t genv, float* v) {
    constexpr int M = Rom<T>::m;
    if (t > g.t_final) resample<T>(p, g, z, genv);
    float lo[M], hi[M];
    rom_bounds<T>(p, z, lo, hi);
    const float frac = div_rn(sub_rn(t, g.ramp_t0), sub_rn(g.t_final, g.ramp_t0));
#pragma unroll
    for (int j = 0; j < M; ++j) {
        const float ramp = add_rn(g.rv0[j], mul_rn(sub_rn(g.rv1[j], g.rv0[j]), frac));
        const float sn = add_rn(mul_rn(g.smag[j], sinf(add_rn(mul_rn(g.sfreq[j], t), g.soff[j]))), g.smean[j]);
        float a = mul_rn(g.w[0], rom_clip<T>(g.hold[j], lo[j], hi[j]));
        a = add_rn(a, mul_rn(g.w[1], rom_clip<T>(ramp, lo[j], hi[j])));
        a = add_rn(a, mul_rn(g.w[2], rom_clip<T>(g.ext[j], lo[j], hi[j])));
        a = add_rn(a, mul_rn(g.w[3], rom_clip<T>(sn, lo[j], hi[j])));
        v[j] = a;
    }
}

// the head of step_rom_idx, :579-580: self.v = get_input_t(self.t, last knot); v[stationary] = 0
template <int T>
__device__ __forceinline__ void eval_v(const B200RomFamilyParams& p, GenP<Rom<T>::m>& g, const float* z, uint64_t genv) {
    input_t<T>(p, g, g.t, z, genv, g.v);
#pragma unroll
    for (int j = 0; j < Rom<T>::m; ++j) g.v[j] = g.stat ? 0.0f : g.v[j];
}

// z_next of step_rom_idx, :581-583 (velocity states of a stationary env are zeroed)
template <int T>
__device__ __forceinline__ void next_knot(const B200RomFamilyParams& p, const GenP<Rom<T>::m>& g, const float* z, float* zn) {
    rom_f<T>(p.rom_dt, z, g.v, zn);
    if (g.stat) {
#pragma unroll
        for (int c = Rom<T>::vs; c < Rom<T>::n; ++c) zn[c] = 0.0f;
    }
}

// ---- kernels --------------------------------------------------------------------------------------------
template <int T>
__global__ void fam_init_kernel(const __grid_constant__ B200RomFamilyParams p, const __grid_constant__ B200RomState s, long long env_off) {
    constexpr int M = Rom<T>::m;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.num_envs) return;
    const uint32_t ctr = static_cast<uint32_t>(s.rng_ctr[i]);
    const philox::Stream rng(p.seed_lo, p.seed_hi, static_cast<uint64_t>(env_off + i), ctr);
    const uint4 w = rng.words(philox::ROM_INIT, 0);
#pragma unroll
    for (int j = 0; j < M; ++j)
        s.ramp_v_end[static_cast<size_t>(i) * M + j] = affine_rn(sub_rn(p.v_max[j], p.v_min[j]), philox::u01(philox::word(w, j)), p.v_min[j]);
    s.rng_ctr[i] = static_cast<int32_t>(ctr + 1);
}

// TrajectoryGenerator.reset_idx for one env (rom_dynamics.py:595-605); tr / vt = this env's horizon windows (HBM or a shared-memory tile)
template <int T>
__device__ __forceinline__ void reset_env(const B200RomFamilyParams& p, const B200RomState& s, size_t i, uint64_t genv, const float* __restrict__ z_in,
                                          bool in_idx, float* tr, float* vt) {
    constexpr int RN = Rom<T>::n, M = Rom<T>::m;
    const int w = p.window;
    GenP<M> g;
    load_params(s, i, g);
    float z[RN];
    if (in_idx) {
        // reset_idx, :596-602: windows cleared, last knot = z, clocks rewound to -W knots, parameters resampled from z
#pragma unroll
        for (int c = 0; c < RN; ++c) z[c] = z_in[i * RN + c];
        g.k = -static_cast<float>(w);
        g.t = mul_rn(g.k, p.rom_dt);
        g.t_final = g.t;
        resample<T>(p, g, z, genv);
        // warm-up, :604-605: W ROM steps with the ROM clock; knot j lands in row j of the shifted window, input j in row j
#pragma unroll
        for (int c = 0; c < RN; ++c) tr[c] = z[c];
        for (int j = 0; j < w; ++j) {
            eval_v<T>(p, g, z, genv);
            float zn[RN];
            next_knot<T>(p, g, z, zn);
#pragma unroll
            for (int c = 0; c < RN; ++c) tr[(j + 1) * RN + c] = z[c] = zn[c];
#pragma unroll
            for (int c = 0; c < M; ++c) vt[j * M + c] = g.v[c];
            g.k = add_rn(g.k, 1.0f);
            g.t = add_rn(g.t, p.rom_dt);
        }
        s.t[i] = g.t, s.k[i] = g.k;
    } else {
        // not reset: the env still takes part in the W input evaluations of the warm-up (:579 runs over every env).  Its clock and
        // last knot do not move, so once t <= t_final every further evaluation repeats the same v.
#pragma unroll
        for (int c = 0; c < RN; ++c) z[c] = tr[w * RN + c];
        int it = 0;
        while (it < w && g.t > g.t_final) {
            eval_v<T>(p, g, z, genv);
            ++it;
        }
        if (it < w) eval_v<T>(p, g, z, genv);
    }
    store_params(s, i, g);
#pragma unroll
    for (int c = 0; c < M; ++c) s.v[i * M + c] = g.v[c];
}

template <int T>
__global__ void __launch_bounds__(128) fam_reset_kernel(const __grid_constant__ B200RomFamilyParams p, const __grid_constant__ B200RomState s,
                                                        const float* __restrict__ z_in, const uint8_t* __restrict__ mask, long long env_off) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.num_envs) return;
    const size_t i = static_cast<size_t>(e);
    reset_env<T>(p, s, i, static_cast<uint64_t>(env_off) + i, z_in, mask == nullptr || mask[i] != 0,
                 s.trajectory + i * (p.window + 1) * Rom<T>::n, s.v_trajectory + i * p.window * Rom<T>::m);
}

// LeggedRobotTrajectory's view of the generator (get_trajectory, rom_dynamics.py:607-612, written into env.trajectory [N, N_h, n]):
// knot j of the horizon interpolated towards the next one by (t - (k-1) rom.dt) / rom.dt — the arithmetic of rom.cu's write_views
template <int RN>
__device__ __forceinline__ void write_env_window(const B200RomFamilyParams& p, const float* tr, float t, float k, float* et) {
    const float sc = sub_rn(t, mul_rn(sub_rn(k, 1.0f), p.rom_dt));
    const int horizon = p.window / p.dN;
    for (int j = 0; j < horizon; ++j) {
        const int a = j * p.dN * RN;
#pragma unroll
        for (int c = 0; c < RN; ++c) et[j * RN + c] = add_rn(tr[a + c], div_rn(mul_rn(sub_rn(tr[a + RN + c], tr[a + c]), sc), p.rom_dt));
    }
}

template <int T>
__global__ void __launch_bounds__(128) fam_step_kernel(const __grid_constant__ B200RomFamilyParams p, const __grid_constant__ B200RomState s,
                                                       const uint8_t* __restrict__ mask, long long env_off) {
    constexpr int RN = Rom<T>::n, M = Rom<T>::m;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (e >= p.num_envs) return;
    const size_t i = static_cast<size_t>(e);
    const uint64_t genv = static_cast<uint64_t>(env_off) + i;
    const int w = p.window;
    float* tr = s.trajectory + i * (w + 1) * RN;
    float* vt = s.v_trajectory + i * w * M;
    GenP<M> g;
    load_params(s, i, g);
    const bool in_idx = mask == nullptr || mask[i] != 0;
    // step_idx, :571-575: only envs of idx whose ROM clock is due advance; the input (and a due resample) is evaluated for all
    const bool due = in_idx && (g.t >= sub_rn(mul_rn(g.k, p.rom_dt), 1e-5f));
    const uint32_t ctr0 = g.ctr;
    float z[RN];
#pragma unroll
    for (int c = 0; c < RN; ++c) z[c] = tr[w * RN + c];
    eval_v<T>(p, g, z, genv);
    if (due) {   // :581-588
        float zn[RN];
        next_knot<T>(p, g, z, zn);
        for (int c = 0; c < w * RN; ++c) tr[c] = tr[c + RN];
#pragma unroll
        for (int c = 0; c < RN; ++c) tr[w * RN + c] = zn[c];
        for (int c = 0; c < (w - 1) * M; ++c) vt[c] = vt[c + M];
#pragma unroll
        for (int c = 0; c < M; ++c) vt[(w - 1) * M + c] = g.v[c];
        g.k = add_rn(g.k, 1.0f);
        s.k[i] = g.k;
    }
    if (in_idx) {
        g.t = add_rn(g.t, p.dt_loop);
        s.t[i] = g.t;
    }
    if (g.ctr != ctr0) store_params(s, i, g);
#pragma unroll
    for (int c = 0; c < M; ++c) s.v[i * M + c] = g.v[c];
    if (s.env_trajectory) write_env_window<RN>(p, tr, g.t, g.k, s.env_trajectory + i * (w / p.dN) * RN);
}

// The same step with the CTA's horizon windows staged through shared memory.  The windows of 128 consecutive envs are ONE contiguous
// block of HBM ([N, W+1, n] row-major), so a CTA whose envs are due moves them with two 1-D bulk copies (TMA) each way instead of 128
// strided per-thread streams; each thread then shifts its own row in shared memory.  ROM clocks run in lock-step unless step_idx is
// used with partial masks, so `due` is CTA-uniform in practice: on the calls between knots (rom.dt / dt_loop - 1 of every rom.dt /
// dt_loop) no window is touched beyond the last knot.  A partial tail CTA whose block is not a multiple of 16 bytes takes the per-thread
// path of fam_step_kernel.
constexpr int FT = 128;   // envs (threads) per CTA

template <int T>
__global__ void __launch_bounds__(FT) fam_step_tile_kernel(const __grid_constant__ B200RomFamilyParams p, const __grid_constant__ B200RomState s,
                                                           const uint8_t* __restrict__ mask, long long env_off) {
    constexpr int RN = Rom<T>::n, M = Rom<T>::m;
    extern __shared__ __align__(128) float fam_smem[];
    __shared__ uint64_t bar;
    const int w = p.window;
    const int rl = (w + 1) * RN, vl = w * M, hl = (w / p.dN) * RN;   // floats per env: the two windows, the env's interpolated view
    float* s_tr = fam_smem;
    float* s_vt = fam_smem + FT * rl;
    float* s_et = s_vt + FT * vl;                                      // only carved / used with an attached env (s.env_trajectory)
    const int env0 = blockIdx.x * FT;
    const int nenv = min(FT, p.num_envs - env0);
    const int e = threadIdx.x;
    const bool valid = e < nenv;
    const size_t i = static_cast<size_t>(env0) + (valid ? e : 0);
    const uint64_t genv = static_cast<uint64_t>(env_off) + i;
    pdl_launch_dependents();
    pdl_wait();
    float* g_tr = s.trajectory + static_cast<size_t>(env0) * rl;
    float* g_vt = s.v_trajectory + static_cast<size_t>(env0) * vl;
    GenP<M> g;
    load_params(s, i, g);
    const bool in_idx = valid && (mask == nullptr || mask[i] != 0);
    const bool due = in_idx && (g.t >= sub_rn(mul_rn(g.k, p.rom_dt), 1e-5f));
    const uint32_t bytes_tr = static_cast<uint32_t>(nenv) * rl * 4u, bytes_vt = static_cast<uint32_t>(nenv) * vl * 4u;
    const uint32_t bytes_et = static_cast<uint32_t>(nenv) * hl * 4u;
    // an attached env reads the whole state window on EVERY call (its view is an interpolation over all knots), so that window is
    // staged on every call and the view leaves as one bulk store; the input window still moves only when a knot is appended
    float* g_et = s.env_trajectory ? s.env_trajectory + static_cast<size_t>(env0) * hl : nullptr;
    const bool tile_ok = ((bytes_tr | bytes_vt | (g_et ? bytes_et : 0u)) & 15u) == 0 &&
                         ((reinterpret_cast<uintptr_t>(g_tr) | reinterpret_cast<uintptr_t>(g_vt) | reinterpret_cast<uintptr_t>(g_et)) & 15u) == 0;
    const bool any_due = __syncthreads_or(due) != 0;
    const bool staged = any_due && tile_ok;              // both windows in shared memory, written back
    const bool staged_tr = tile_ok && (any_due || g_et != nullptr);
    if (staged_tr) {
        if (e == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (e == 0) {
            mbar_expect_tx(&bar, bytes_tr + (staged ? bytes_vt : 0u));
            bulk_g2s(s_tr, g_tr, bytes_tr, &bar);
            if (staged) bulk_g2s(s_vt, g_vt, bytes_vt, &bar);
        }
        mbar_wait(&bar, 0);
    }
    float* tr = staged_tr ? s_tr + e * rl : s.trajectory + i * rl;
    float* vt = staged ? s_vt + e * vl : s.v_trajectory + i * vl;
    const uint32_t ctr0 = g.ctr;
    if (valid) {
        float z[RN];
#pragma unroll
        for (int c = 0; c < RN; ++c) z[c] = tr[w * RN + c];
        eval_v<T>(p, g, z, genv);
        if (due) {   // :581-588
            float zn[RN];
            next_knot<T>(p, g, z, zn);
            for (int c = 0; c < w * RN; ++c) tr[c] = tr[c + RN];
#pragma unroll
            for (int c = 0; c < RN; ++c) tr[w * RN + c] = zn[c];
            for (int c = 0; c < (w - 1) * M; ++c) vt[c] = vt[c + M];
#pragma unroll
            for (int c = 0; c < M; ++c) vt[(w - 1) * M + c] = g.v[c];
            g.k = add_rn(g.k, 1.0f);
            s.k[i] = g.k;
        }
        if (in_idx) {
            g.t = add_rn(g.t, p.dt_loop);
            s.t[i] = g.t;
        }
        if (g.ctr != ctr0) store_params(s, i, g);
#pragma unroll
        for (int c = 0; c < M; ++c) s.v[i * M + c] = g.v[c];
        if (g_et) write_env_window<RN>(p, tr, g.t, g.k, staged_tr ? s_et + e * hl : s.env_trajectory + i * hl);
    }
    if (staged_tr) {
        fence_proxy_async();
        __syncthreads();
        if (e == 0) {
            if (staged) {
                bulk_s2g(g_tr, s_tr, bytes_tr);
                bulk_s2g(g_vt, s_vt, bytes_vt);
            }
            if (g_et) bulk_s2g(g_et, s_et, bytes_et);
            bulk_commit();
            bulk_wait0();
        }
    }
}

// reset_idx with the same staging.  Per-thread, every env streams its own W+1 rows to HBM (ncu at 1 M envs, n = 6: 26 sectors per store
// request, 603 MB written for 403 MB of windows, L2 at 66 %); staged, the CTA's block leaves as two bulk stores.  With a partial mask the
// block is fetched first so that the other envs' windows survive; reset() of every env overwrites all of it.
template <int T>
__global__ void __launch_bounds__(FT) fam_reset_tile_kernel(const __grid_constant__ B200RomFamilyParams p, const __grid_constant__ B200RomState s,
                                                            const float* __restrict__ z_in, const uint8_t* __restrict__ mask, long long env_off) {
    constexpr int RN = Rom<T>::n, M = Rom<T>::m;
    extern __shared__ __align__(128) float fam_smem[];
    __shared__ uint64_t bar;
    const int w = p.window;
    const int rl = (w + 1) * RN, vl = w * M;
    float* s_tr = fam_smem;
    float* s_vt = fam_smem + FT * rl;
    const int env0 = blockIdx.x * FT;
    const int nenv = min(FT, p.num_envs - env0);
    const int e = threadIdx.x;
    const bool valid = e < nenv;
    const size_t i = static_cast<size_t>(env0) + (valid ? e : 0);
    float* g_tr = s.trajectory + static_cast<size_t>(env0) * rl;
    float* g_vt = s.v_trajectory + static_cast<size_t>(env0) * vl;
    const uint32_t bytes_tr = static_cast<uint32_t>(nenv) * rl * 4u, bytes_vt = static_cast<uint32_t>(nenv) * vl * 4u;
    const bool staged = ((bytes_tr | bytes_vt) & 15u) == 0 && ((reinterpret_cast<uintptr_t>(g_tr) | reinterpret_cast<uintptr_t>(g_vt)) & 15u) == 0;
    pdl_launch_dependents();
    pdl_wait();
    if (staged && mask != nullptr) {
        if (e == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (e == 0) {
            mbar_expect_tx(&bar, bytes_tr + bytes_vt);
            bulk_g2s(s_tr, g_tr, bytes_tr, &bar);
            bulk_g2s(s_vt, g_vt, bytes_vt, &bar);
        }
        mbar_wait(&bar, 0);
    }
    if (valid)
        reset_env<T>(p, s, i, static_cast<uint64_t>(env_off) + i, z_in, mask == nullptr || mask[i] != 0, staged ? s_tr + e * rl : s.trajectory + i * rl,
                     staged ? s_vt + e * vl : s.v_trajectory + i * vl);
    if (staged) {
        fence_proxy_async();
        __syncthreads();
        if (e == 0) {
            bulk_s2g(g_tr, s_tr, bytes_tr);
            bulk_s2g(g_vt, s_vt, bytes_vt);
            bulk_commit();
            bulk_wait0();
        }
    }
}

template <int T>
__global__ void __launch_bounds__(128) fam_input_kernel(const __grid_constant__ B200RomFamilyParams p, const __grid_constant__ B200RomState s,
                                                        const float* __restrict__ t_in, const float* __restrict__ z_in, float* __restrict__ v_out,
                                                        long long env_off) {
    constexpr int RN = Rom<T>::n, M = Rom<T>::m;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.num_envs) return;
    const size_t i = static_cast<size_t>(e);
    GenP<M> g;
    load_params(s, i, g);
    const uint32_t ctr0 = g.ctr;
    float z[RN], v[M];
#pragma unroll
    for (int c = 0; c < RN; ++c) z[c] = z_in[i * RN + c];
    input_t<T>(p, g, t_in[i], z, static_cast<uint64_t>(env_off) + i, v);
    if (g.ctr != ctr0) store_params(s, i, g);
#pragma unroll
    for (int c = 0; c < M; ++c) v_out[i * M + c] = v[c];
}

template <int T>
__global__ void fam_f_kernel(float dt, const float* __restrict__ z, const float* __restrict__ v, float* __restrict__ out, long long n) {
    constexpr int RN = Rom<T>::n, M = Rom<T>::m;
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float x[RN], u[M], o[RN];
#pragma unroll
    for (int c = 0; c < RN; ++c) x[c] = z[i * RN + c];
#pragma unroll
    for (int c = 0; c < M; ++c) u[c] = v[i * M + c];
    rom_f<T>(dt, x, u, o);
#pragma unroll
    for (int c = 0; c < RN; ++c) out[i * RN + c] = o[c];
}

template <int T>
__global__ void fam_des_kernel(const float* __restrict__ z, const float* __restrict__ v, float* __restrict__ pose, float* __restrict__ vel,
                               long long n) {
    constexpr int RN = Rom<T>::n, M = Rom<T>::m;
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float x[RN], u[M], ps[3], vl[3];
#pragma unroll
    for (int c = 0; c < RN; ++c) x[c] = z[i * RN + c];
#pragma unroll
    for (int c = 0; c < M; ++c) u[c] = v[i * M + c];
    if constexpr (T == B200GYM_ROM_SINGLE_INT_2D) {          // :198-199
        ps[0] = x[0], ps[1] = x[1], ps[2] = atan2f(u[1], u[0]);
        vl[0] = u[0], vl[1] = u[1], vl[2] = 0.0f;
    } else if constexpr (T == B200GYM_ROM_DOUBLE_INT_2D) {   // :230-232
        ps[0] = x[0], ps[1] = x[1], ps[2] = atan2f(x[3], x[2]);
        vl[0] = x[2], vl[1] = x[3], vl[2] = 0.0f;
    } else {
        const float c = cosf(x[2]), s = sinf(x[2]);
        ps[0] = x[0], ps[1] = x[1], ps[2] = x[2];
        if constexpr (T == B200GYM_ROM_UNICYCLE) {                   // :286-290
            vl[0] = mul_rn(u[0], c), vl[1] = mul_rn(u[0], s), vl[2] = u[1];
        } else if constexpr (T == B200GYM_ROM_LATERAL_UNICYCLE) {    // :318-322 — om = v[:, 1], as the reference writes it
            vl[0] = sub_rn(mul_rn(u[0], c), mul_rn(u[1], s));
            vl[1] = add_rn(mul_rn(u[0], s), mul_rn(u[1], c));
            vl[2] = u[1];
        } else if constexpr (T == B200GYM_ROM_EXTENDED_UNICYCLE) {   // :354-358
            vl[0] = mul_rn(x[3], c), vl[1] = mul_rn(x[3], s), vl[2] = x[4];
        } else {                                           // :416-420
            vl[0] = sub_rn(mul_rn(x[3], c), mul_rn(x[4], s));
            vl[1] = add_rn(mul_rn(x[3], s), mul_rn(x[4], c));
            vl[2] = x[5];
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) pose[i * 3 + c] = ps[c], vel[i * 3 + c] = vl[c];
}

template <int T>
__global__ void fam_bounds_kernel(const __grid_constant__ B200RomFamilyParams p, const float* __restrict__ z, const float* __restrict__ v,
                                  float* __restrict__ v_lo, float* __restrict__ v_hi, float* __restrict__ v_clip, long long n) {
    constexpr int RN = Rom<T>::n, M = Rom<T>::m;
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float x[RN], lo[M], hi[M];
#pragma unroll
    for (int c = 0; c < RN; ++c) x[c] = z[i * RN + c];
    rom_bounds<T>(p, x, lo, hi);
#pragma unroll
    for (int c = 0; c < M; ++c) {
        if (v_lo) v_lo[i * M + c] = lo[c];
        if (v_hi) v_hi[i * M + c] = hi[c];
        if (v_clip) v_clip[i * M + c] = rom_clip<T>(v[i * M + c], lo[c], hi[c]);
    }
}

// yaw of scipy's Rotation.from_quat(q).as_euler('xyz')[..., -1] for q = (x, y, z, w): atan2(R10, R00) of the rotation matrix, written on
// the un-normalised quaternion (numerator and denominator both scale with |q|^2, scipy normalises first)
__device__ __forceinline__ float quat_yaw(float qx, float qy, float qz, float qw) {
    return atan2f(2.0f * (qx * qy + qw * qz), (qw * qw + qx * qx) - (qy * qy + qz * qz));
}

template <int T>
__global__ void fam_proj_kernel(const float* __restrict__ x, float* __restrict__ z, long long n) {
    constexpr int RN = Rom<T>::n;
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const float* r = x + i * 13;
    float o[RN];
    o[0] = r[0], o[1] = r[1];
    if constexpr (T == B200GYM_ROM_DOUBLE_INT_2D) {            // :227-228
        o[2] = r[7], o[3] = r[8];
    } else if constexpr (T != B200GYM_ROM_SINGLE_INT_2D) {     // :280-284, :360-365, :422-427
        const float yaw = quat_yaw(r[3], r[4], r[5], r[6]);
        o[2] = yaw;
        if constexpr (T == B200GYM_ROM_EXTENDED_UNICYCLE || T == B200GYM_ROM_EXTENDED_LATERAL_UNICYCLE) {
            const float cy = cosf(yaw), sy = sinf(yaw);   // yaw2rot(yaw) @ v_xy, deep_tube_learning/utils.py:88-96
            o[3] = cy * r[7] + sy * r[8];
            if constexpr (T == B200GYM_ROM_EXTENDED_LATERAL_UNICYCLE) o[4] = -sy * r[7] + cy * r[8];
            o[RN - 1] = r[12];                            // x[:, -1]: yaw rate
        }
    }
#pragma unroll
    for (int c = 0; c < RN; ++c) z[i * RN + c] = o[c];
}

int check_fam(const B200RomFamilyParams* p, const B200RomState* s, const char* what) {
    B200_REQUIRE(p && s, B200GYM_EINVAL, "%s: null argument", what);
    B200_REQUIRE(p->num_envs > 0, B200GYM_EINVAL, "%s: num_envs must be positive (got %d)", what, p->num_envs);
    B200_REQUIRE(p->rom_type >= 0 && p->rom_type < B200GYM_ROM_NUM_TYPES, B200GYM_EINVAL, "%s: unknown rom_type %d", what, p->rom_type);
    B200_REQUIRE(p->window >= 2 && p->dN >= 1 && p->window % p->dN == 0, B200GYM_EINVAL, "%s: window %d must be a positive multiple of dN %d (>= 2)",
                 what, p->window, p->dN);
    B200_REQUIRE(p->rom_dt > 0.0f, B200GYM_EINVAL, "%s: rom_dt must be positive", what);
    const void* must[] = {s->trajectory, s->v_trajectory, s->v, s->t, s->k, s->t_final, s->weights, s->sample_hold_input,
                          s->extreme_input, s->ramp_v_start, s->ramp_v_end, s->ramp_t_start, s->sin_mag, s->sin_freq, s->sin_off,
                          s->sin_mean, s->stationary_inds, s->rng_ctr};
    for (const void* q : must) B200_REQUIRE(q != nullptr, B200GYM_EINVAL, "%s: null state tensor", what);
    B200_REQUIRE(b200_aligned16(s->weights), B200GYM_EALIGN, "%s: weights must be 16-byte aligned", what);
    return B200GYM_OK;
}

#define FAM_DISPATCH(type, CALL)                                      \
    switch (type) {                                                   \
        case 0: { constexpr int T = 0; CALL; } break;                 \
        case 1: { constexpr int T = 1; CALL; } break;                 \
        case 2: { constexpr int T = 2; CALL; } break;                 \
        case 3: { constexpr int T = 3; CALL; } break;                 \
        case 4: { constexpr int T = 4; CALL; } break;                 \
        default: { constexpr int T = 5; CALL; } break;                \
    }

bool fam_tile_enabled() {
    static const bool on = []() { const char* e = getenv("B200GYM_ROMFAM_TILE"); return !(e && e[0] == '0'); }();
    return on;
}
size_t fam_tile_bytes(const B200RomFamilyParams* p, bool env_view = false) {   // both horizon windows of a CTA's 128 envs (+ the env's view)
    const int n_of[B200GYM_ROM_NUM_TYPES] = {2, 4, 3, 3, 5, 6}, m_of[B200GYM_ROM_NUM_TYPES] = {2, 2, 2, 3, 2, 3};
    const int hl = env_view ? (p->window / (p->dN > 0 ? p->dN : 1)) * n_of[p->rom_type] : 0;
    return static_cast<size_t>(128) * ((p->window + 1) * n_of[p->rom_type] + p->window * m_of[p->rom_type] + hl) * sizeof(float);
}

int check_rows(int32_t rom_type, int64_t n_rows, const char* what) {
    B200_REQUIRE(rom_type >= 0 && rom_type < B200GYM_ROM_NUM_TYPES, B200GYM_EINVAL, "%s: unknown rom_type %d", what, rom_type);
    B200_REQUIRE(n_rows > 0 && n_rows < (1ll << 31) * 256, B200GYM_EINVAL, "%s: n_rows out of range", what);
    return B200GYM_OK;
}

}  // namespace

extern "C" {

int b200gym_romfam_f(int32_t rom_type, float dt, const float* z, const float* v, float* z_next, int64_t n_rows, void* stream) {
    if (int rc = check_rows(rom_type, n_rows, "romfam_f")) return rc;
    B200_REQUIRE(z && v && z_next, B200GYM_EINVAL, "romfam_f: null argument");
    const unsigned grid = static_cast<unsigned>((n_rows + 255) / 256);
    FAM_DISPATCH(rom_type, (fam_f_kernel<T><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(dt, z, v, z_next, n_rows)));
    B200_LAUNCH_CHECK("romfam_f");
    return B200GYM_OK;
}

int b200gym_romfam_des_pose_vel(int32_t rom_type, const float* z, const float* v, float* pose, float* vel, int64_t n_rows, void* stream) {
    if (int rc = check_rows(rom_type, n_rows, "romfam_des_pose_vel")) return rc;
    B200_REQUIRE(z && v && pose && vel, B200GYM_EINVAL, "romfam_des_pose_vel: null argument");
    const unsigned grid = static_cast<unsigned>((n_rows + 255) / 256);
    FAM_DISPATCH(rom_type, (fam_des_kernel<T><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(z, v, pose, vel, n_rows)));
    B200_LAUNCH_CHECK("romfam_des_pose_vel");
    return B200GYM_OK;
}

int b200gym_romfam_input_bounds(const B200RomFamilyParams* p, const float* z, const float* v, float* v_lo, float* v_hi, float* v_clipped,
                                int64_t n_rows, void* stream) {
    B200_REQUIRE(p && z, B200GYM_EINVAL, "romfam_input_bounds: null argument");
    if (int rc = check_rows(p->rom_type, n_rows, "romfam_input_bounds")) return rc;
    B200_REQUIRE((v_clipped == nullptr) == (v == nullptr), B200GYM_EINVAL, "romfam_input_bounds: v and v_clipped go together");
    B200_REQUIRE(v_lo || v_hi || v_clipped, B200GYM_EINVAL, "romfam_input_bounds: no output requested");
    B200_REQUIRE(p->rom_dt > 0.0f, B200GYM_EINVAL, "romfam_input_bounds: rom_dt must be positive");
    const unsigned grid = static_cast<unsigned>((n_rows + 255) / 256);
    FAM_DISPATCH(p->rom_type, (fam_bounds_kernel<T><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*p, z, v, v_lo, v_hi, v_clipped, n_rows)));
    B200_LAUNCH_CHECK("romfam_input_bounds");
    return B200GYM_OK;
}

int b200gym_romfam_proj_z(int32_t rom_type, const float* x, float* z, int64_t n_rows, void* stream) {
    if (int rc = check_rows(rom_type, n_rows, "romfam_proj_z")) return rc;
    B200_REQUIRE(x && z, B200GYM_EINVAL, "romfam_proj_z: null argument");
    const unsigned grid = static_cast<unsigned>((n_rows + 255) / 256);
    FAM_DISPATCH(rom_type, (fam_proj_kernel<T><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, z, n_rows)));
    B200_LAUNCH_CHECK("romfam_proj_z");
    return B200GYM_OK;
}

int b200gym_romfam_gen_init(const B200RomFamilyParams* p, const B200RomState* s, int64_t env_id_offset, void* stream) {
    if (int rc = check_fam(p, s, "romfam_gen_init")) return rc;
    const int grid = (p->num_envs + 127) / 128;
    FAM_DISPATCH(p->rom_type, (fam_init_kernel<T><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p, *s, env_id_offset)));
    B200_LAUNCH_CHECK("romfam_gen_init");
    return B200GYM_OK;
}

int b200gym_romfam_gen_reset(const B200RomFamilyParams* p, const B200RomState* s, const float* z, const uint8_t* reset_mask,
                             int64_t env_id_offset, void* stream) {
    if (int rc = check_fam(p, s, "romfam_gen_reset")) return rc;
    B200_REQUIRE(z, B200GYM_EINVAL, "romfam_gen_reset: z missing");
    const int grid = (p->num_envs + 127) / 128;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = fam_tile_bytes(p);
    if (fam_tile_enabled() && smem <= 200 * 1024) {
        FAM_DISPATCH(p->rom_type, {
            static size_t granted = 0;
            if (smem + 1024 > 48 * 1024 && smem > granted) {
                cudaFuncSetAttribute(fam_reset_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
                granted = smem;
            }
            b200_launch_pdl(p->num_envs, fam_reset_tile_kernel<T>, dim3(grid), dim3(FT), smem, st, *p, *s, z, reset_mask, (long long)env_id_offset);
        });
    } else {
        FAM_DISPATCH(p->rom_type, (fam_reset_kernel<T><<<grid, 128, 0, st>>>(*p, *s, z, reset_mask, env_id_offset)));
    }
    B200_LAUNCH_CHECK("romfam_gen_reset");
    return B200GYM_OK;
}

int b200gym_romfam_gen_step(const B200RomFamilyParams* p, const B200RomState* s, const uint8_t* step_mask, int64_t env_id_offset,
                            void* stream) {
    if (int rc = check_fam(p, s, "romfam_gen_step")) return rc;
    const int grid = (p->num_envs + 127) / 128;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // windows staged through shared memory (TMA bulk tiles) unless they do not fit or B200GYM_ROMFAM_TILE=0 asks for the per-thread kernel
    B200_REQUIRE(s->env_trajectory == nullptr || (p->dN > 0 && p->window % p->dN == 0), B200GYM_EINVAL,
                 "romfam_gen_step: an attached env view needs window %% dN == 0");
    const size_t smem = fam_tile_bytes(p, s->env_trajectory != nullptr);
    if (fam_tile_enabled() && smem <= 200 * 1024) {
        FAM_DISPATCH(p->rom_type, {
            static size_t granted = 0;   // per rom class: raise the dynamic shared-memory limit once per size, not on every call
            if (smem + 1024 > 48 * 1024 && smem > granted) {
                cudaFuncSetAttribute(fam_step_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
                granted = smem;
            }
            b200_launch_pdl(p->num_envs, fam_step_tile_kernel<T>, dim3(grid), dim3(FT), smem, st, *p, *s, step_mask, (long long)env_id_offset);
        });
    } else {
        FAM_DISPATCH(p->rom_type, (b200_launch_pdl(p->num_envs, fam_step_kernel<T>, dim3(grid), dim3(128), 0, st, *p, *s, step_mask, (long long)env_id_offset)));
    }
    B200_LAUNCH_CHECK("romfam_gen_step");
    return B200GYM_OK;
}

int b200gym_romfam_gen_input(const B200RomFamilyParams* p, const B200RomState* s, const float* t, const float* z, float* v_out,
                             int64_t env_id_offset, void* stream) {
    if (int rc = check_fam(p, s, "romfam_gen_input")) return rc;
    B200_REQUIRE(t && z && v_out, B200GYM_EINVAL, "romfam_gen_input: null argument");
    const int grid = (p->num_envs + 127) / 128;
    FAM_DISPATCH(p->rom_type, (fam_input_kernel<T><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p, *s, t, z, v_out, env_id_offset)));
    B200_LAUNCH_CHECK("romfam_gen_input");
    return B200GYM_OK;
}

}  // extern "C"
