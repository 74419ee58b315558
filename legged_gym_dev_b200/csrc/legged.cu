// Group R kernels: LeggedRobot / Anymal step pipeline.
//
//   pd_torques_kernel     LeggedRobot._compute_torques           legged_robot.py:389-413 (+ action clip :86-87)
//   lstm_torques_kernel   Anymal._compute_torques (LSTM branch)  anymal.py:71-78 + LSTMsea TorchScript
//   post_physics_kernel   LeggedRobot.post_physics_step          legged_robot.py:106-134 and everything it calls
//
// Layout: every API-visible tensor keeps the reference's row-major [N, k] (AoS per env) layout, so the
// Python attributes are plain contiguous tensors and PhysX-owned buffers are consumed as given.  The
// post-physics kernel gets coalescing from the hardware instead: a CTA owns a tile of TILE consecutive
// envs, whose rows form ONE contiguous byte range per tensor; each range is moved HBM<->shared memory by
// a 1-D TMA bulk copy (cp.async.bulk / UBLKCP) issued by one thread and tracked by an mbarrier.  Four
// lanes ("quad") then work on one env out of shared memory: 3 DOFs, 1 foot, 2 penalised bodies, 1 of the
// body-frame vectors and 12 observation columns per lane, with 2-step shuffle reductions for per-env sums.
#include "common.cuh"
#include "philox.cuh"
#include "../../include/b200gym.h"

namespace {

constexpr int LPE = 4;          // lanes per env in the post-physics kernel
constexpr int ND = B200GYM_NUM_DOF;
constexpr int HPAD = 192;       // padded per-env stride of the raw height tile (>= 187)

enum Term {
    T_ACTION_RATE = 0, T_ANG_VEL_XY, T_BASE_HEIGHT, T_COLLISION, T_DOF_ACC, T_DOF_POS_LIMITS, T_DOF_VEL,
    T_DOF_VEL_LIMITS, T_FEET_AIR_TIME, T_FEET_CONTACT_FORCES, T_LIN_VEL_Z, T_ORIENTATION, T_STAND_STILL,
    T_STUMBLE, T_TORQUE_LIMITS, T_TORQUES, T_TRACKING_ANG_VEL, T_TRACKING_LIN_VEL, T_TERMINATION
};

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
struct ActuatorNet {
    float w_ih0[32][2], w_hh0[32][8], b0[32];   // b0 = b_ih0 (+ b_hh0 added separately to mirror ATen)
    float bh0[32];
    float w_ih1[32][8], w_hh1[32][8], b1[32], bh1[32];
    float w_lin[8], b_lin, in0, in1, out_scale;
};
__constant__ ActuatorNet c_net;

__device__ __forceinline__ float sigmoid_acc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_acc(float x) {
    // tanh(x) = 1 - 2/(exp(2x)+1); abs error ~1e-7, inside the 1e-5 contract
    const float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}

template <int NIN>
__device__ __forceinline__ void lstm_cell(const float (&w_ih)[32][NIN], const float (&w_hh)[32][8], const float (&bi)[32],
                                          const float (&bh)[32], const float (&x)[NIN], float (&h)[8], float (&c)[8]) {
    float gate[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        float a = bi[r], bsum = bh[r];
#pragma unroll
        for (int k = 0; k < NIN; ++k) a = fmaf(w_ih[r][k], x[k], a);
#pragma unroll
        for (int k = 0; k < 8; ++k) bsum = fmaf(w_hh[r][k], h[k], bsum);
        gate[r] = a + bsum;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const float ig = sigmoid_acc(gate[u]), fg = sigmoid_acc(gate[8 + u]);
        const float gg = tanh_acc(gate[16 + u]), og = sigmoid_acc(gate[24 + u]);
        c[u] = fg * c[u] + ig * gg;
        h[u] = og * tanh_acc(c[u]);
    }
}

__global__ void __launch_bounds__(128) lstm_torques_kernel(const __grid_constant__ B200LeggedParams p,
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
#define LD8(dst, src)                                                  \
    v = ldg_stream4(src);                                              \
    dst[0] = v.x, dst[1] = v.y, dst[2] = v.z, dst[3] = v.w;            \
    v = ldg_stream4(src + 1);                                          \
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

// ------------------------------------------------------------------------------------------------
// post-physics
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void quat_rotate_inverse(float qx, float qy, float qz, float qw, float vx, float vy, float vz,
                                                    float& rx, float& ry, float& rz) {
    // isaacgym.torch_utils.quat_rotate_inverse: a = v*(2w^2-1); b = cross(q,v)*w*2; c = q*dot(q,v)*2; a - b + c
    const float s = 2.0f * qw * qw - 1.0f;
    const float cx = qy * vz - qz * vy, cy = qz * vx - qx * vz, cz = qx * vy - qy * vx;
    const float dt2 = (qx * vx + qy * vy + qz * vz) * 2.0f;
    rx = vx * s - cx * qw * 2.0f + qx * dt2;
    ry = vy * s - cy * qw * 2.0f + qy * dt2;
    rz = vz * s - cz * qw * 2.0f + qz * dt2;
}

__device__ __forceinline__ float norm2_rn(float x, float y) { return sqrtf(add_rn(mul_rn(x, x), mul_rn(y, y))); }
__device__ __forceinline__ float norm3_rn(float x, float y, float z) {
    return sqrtf(add_rn(add_rn(mul_rn(x, x), mul_rn(y, y)), mul_rn(z, z)));
}

__device__ __forceinline__ float wrap_to_pi(float a) {
    // legged_gym/utils/math.py:45-48 on fp32 tensors: python-style remainder by fl32(2*pi), then -2*pi where > fl32(pi)
    const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
    float r = fmodf(a, two_pi);
    if (r != 0.0f && r < 0.0f) r = add_rn(r, two_pi);
    if (r > pi) r = sub_rn(r, two_pi);
    return r;
}

__device__ __forceinline__ void resample_commands(const B200LeggedParams& p, const philox::Stream& rng, uint32_t site, float& c0,
                                                  float& c1, float& c2, float& c3) {
    // legged_robot.py:365-387
    const uint4 w = rng.words(site, 0);
    c0 = affine_rn(p.cmd_span[0], philox::u01(w.x), p.cmd_lo[0]);
    c1 = affine_rn(p.cmd_span[1], philox::u01(w.y), p.cmd_lo[1]);
    if (p.heading_command)
        c3 = affine_rn(p.cmd_span[3], philox::u01(w.z), p.cmd_lo[3]);
    else
        c2 = affine_rn(p.cmd_span[2], philox::u01(w.z), p.cmd_lo[2]);
    const float m = norm2_rn(c0, c1) > 0.2f ? 1.0f : 0.0f;
    c0 = mul_rn(c0, m);
    c1 = mul_rn(c1, m);
}

__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

struct Carver {
    unsigned char* base;
    size_t off;
    template <typename T>
    __device__ __host__ T* take(size_t n) {
        T* r = reinterpret_cast<T*>(base + off);
        off += (n * sizeof(T) + 15) & ~static_cast<size_t>(15);
        return r;
    }
};

struct TileSmem {
    float *root, *dof, *contact, *act, *tq, *lact, *ldv, *cmd, *fat;
    uint8_t* lc;
    long long* ep;
    float *sums, *obs, *blv, *bav, *pg, *lrv, *rew;
    uint8_t *reset, *tout;
    int16_t* hraw;
    float *bh, *zpost, *stage;
    double* acc;
    uint64_t* bar;
    int* nreset;
    size_t bytes;
};

template <int TILE, bool ROUGH>
__device__ __host__ inline TileSmem carve_tile(unsigned char* base, int B, int K) {
    Carver c{base, 0};
    TileSmem s;
    s.bar = c.take<uint64_t>(2);
    s.root = c.take<float>(TILE * 13);
    s.dof = c.take<float>(TILE * 24);
    s.contact = c.take<float>(static_cast<size_t>(TILE) * B * 3);
    s.act = c.take<float>(TILE * ND);
    s.tq = c.take<float>(TILE * ND);
    s.lact = c.take<float>(TILE * ND);
    s.ldv = c.take<float>(TILE * ND);
    s.cmd = c.take<float>(TILE * 4);
    s.fat = c.take<float>(TILE * 4);
    s.lc = c.take<uint8_t>(TILE * 4);
    s.ep = c.take<long long>(TILE);
    s.sums = c.take<float>(static_cast<size_t>(K > 0 ? K : 1) * TILE);
    s.obs = c.take<float>(TILE * 48);
    s.blv = c.take<float>(TILE * 3);
    s.bav = c.take<float>(TILE * 3);
    s.pg = c.take<float>(TILE * 3);
    s.lrv = c.take<float>(TILE * 6);
    s.rew = c.take<float>(TILE);
    s.reset = c.take<uint8_t>(TILE);
    s.tout = c.take<uint8_t>(TILE);
    s.acc = c.take<double>(B200GYM_NUM_REWARD_TERMS + 2);
    s.nreset = c.take<int>(4);
    if (ROUGH) {
        s.hraw = c.take<int16_t>(TILE * HPAD);
        s.bh = c.take<float>(TILE);
        s.zpost = c.take<float>(TILE);
        s.stage = c.take<float>((TILE * LPE / 32) * HPAD);
    } else {
        s.hraw = nullptr;
        s.bh = s.zpost = s.stage = nullptr;
    }
    s.bytes = c.off;
    return s;
}

template <typename T>
__device__ __forceinline__ void coop_copy(T* dst, const T* src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

template <int TILE, bool ROUGH>
__global__ void __launch_bounds__(TILE* LPE) post_physics_kernel(const __grid_constant__ B200LeggedParams p,
                                                                 const __grid_constant__ B200LeggedBuffers b, uint64_t step,
                                                                 long long env_off, int do_push) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int B = p.num_bodies, K = p.num_sum_rows, N = p.num_envs, O = p.num_obs;
    const TileSmem s = carve_tile<TILE, ROUGH>(smem_raw, B, K);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile0 = blockIdx.x * TILE;
    const int nvalid = min(TILE, N - tile0);
    const bool full = (nvalid == TILE);

    if (tid == 0) {
        mbar_init(s.bar, 1);
        fence_mbar_init();
        *s.nreset = 0;
    }
    if (tid < B200GYM_NUM_REWARD_TERMS + 2) s.acc[tid] = 0.0;
    __syncthreads();

    // ---- stage the tile: one bulk copy per tensor -------------------------------------------------
    if (full) {
        if (tid == 0) {
            const uint32_t bytes = TILE * (13 + 24 + 3 * B + 4 * ND + 4 + 4) * 4 + TILE * 4 + TILE * 8 + K * TILE * 4;
            mbar_expect_tx(s.bar, bytes);
            bulk_g2s(s.root, b.root_states + static_cast<size_t>(tile0) * 13, TILE * 13 * 4, s.bar);
            bulk_g2s(s.dof, b.dof_state + static_cast<size_t>(tile0) * 24, TILE * 24 * 4, s.bar);
            bulk_g2s(s.contact, b.contact_forces + static_cast<size_t>(tile0) * B * 3, TILE * B * 3 * 4, s.bar);
            bulk_g2s(s.act, b.actions + static_cast<size_t>(tile0) * ND, TILE * ND * 4, s.bar);
            bulk_g2s(s.tq, b.torques + static_cast<size_t>(tile0) * ND, TILE * ND * 4, s.bar);
            bulk_g2s(s.lact, b.last_actions + static_cast<size_t>(tile0) * ND, TILE * ND * 4, s.bar);
            bulk_g2s(s.ldv, b.last_dof_vel + static_cast<size_t>(tile0) * ND, TILE * ND * 4, s.bar);
            bulk_g2s(s.cmd, b.commands + static_cast<size_t>(tile0) * 4, TILE * 4 * 4, s.bar);
            bulk_g2s(s.fat, b.feet_air_time + static_cast<size_t>(tile0) * 4, TILE * 4 * 4, s.bar);
            bulk_g2s(s.lc, b.last_contacts + static_cast<size_t>(tile0) * 4, TILE * 4, s.bar);
            bulk_g2s(s.ep, b.episode_length_buf + tile0, TILE * 8, s.bar);
            for (int k = 0; k < K; ++k)
                bulk_g2s(s.sums + k * TILE, b.episode_sums + static_cast<size_t>(k) * N + tile0, TILE * 4, s.bar);
        }
        mbar_wait(s.bar, 0);
    } else {
        coop_copy(s.root, b.root_states + static_cast<size_t>(tile0) * 13, nvalid * 13);
        coop_copy(s.dof, b.dof_state + static_cast<size_t>(tile0) * 24, nvalid * 24);
        coop_copy(s.contact, b.contact_forces + static_cast<size_t>(tile0) * B * 3, nvalid * B * 3);
        coop_copy(s.act, b.actions + static_cast<size_t>(tile0) * ND, nvalid * ND);
        coop_copy(s.tq, b.torques + static_cast<size_t>(tile0) * ND, nvalid * ND);
        coop_copy(s.lact, b.last_actions + static_cast<size_t>(tile0) * ND, nvalid * ND);
        coop_copy(s.ldv, b.last_dof_vel + static_cast<size_t>(tile0) * ND, nvalid * ND);
        coop_copy(s.cmd, b.commands + static_cast<size_t>(tile0) * 4, nvalid * 4);
        coop_copy(s.fat, b.feet_air_time + static_cast<size_t>(tile0) * 4, nvalid * 4);
        coop_copy(s.lc, b.last_contacts + static_cast<size_t>(tile0) * 4, nvalid * 4);
        coop_copy(s.ep, reinterpret_cast<const long long*>(b.episode_length_buf) + tile0, nvalid);
        for (int k = 0; k < K; ++k) coop_copy(s.sums + k * TILE, b.episode_sums + static_cast<size_t>(k) * N + tile0, nvalid);
        __syncthreads();
    }

    // ---- R6: height scan, one warp per env (legged_robot.py:877-915, math.py:38-42) ---------------
    if (ROUGH) {
        const int H = p.num_heights;
        const int rows = p.terrain_rows, cols = p.terrain_cols;
        for (int e = warp; e < nvalid; e += TILE * LPE / 32) {
            const float* R = s.root + e * 13;
            // quat_apply_yaw: zero x,y, renormalise, rotate (un-fused fp32 ops)
            const float nq = fmaxf(sqrtf(add_rn(mul_rn(R[5], R[5]), mul_rn(R[6], R[6]))), 1e-9f);
            const float qz = div_rn(R[5], nq), qw = div_rn(R[6], nq);
            float part = 0.0f;
            for (int pt = lane; pt < H; pt += 32) {
                int raw = 0;
                if (!p.mesh_plane) {
                    const float hx = p.points_x[pt / p.n_py], hy = p.points_y[pt % p.n_py];
                    const float tx = mul_rn(-mul_rn(qz, hy), 2.0f), ty = mul_rn(mul_rn(qz, hx), 2.0f);
                    float wx = add_rn(add_rn(hx, mul_rn(qw, tx)), -mul_rn(qz, ty));
                    float wy = add_rn(add_rn(hy, mul_rn(qw, ty)), mul_rn(qz, tx));
                    wx = div_rn(add_rn(add_rn(wx, R[0]), p.border_size), p.horizontal_scale);
                    wy = div_rn(add_rn(add_rn(wy, R[1]), p.border_size), p.horizontal_scale);
                    long long ix = static_cast<long long>(wx), iy = static_cast<long long>(wy);
                    ix = ix < 0 ? 0 : (ix > rows - 2 ? rows - 2 : ix);
                    iy = iy < 0 ? 0 : (iy > cols - 2 ? cols - 2 : iy);
                    const int16_t* hs = b.height_samples + ix * cols + iy;
                    const int h1 = __ldg(hs), h2 = __ldg(hs + cols), h3 = __ldg(hs + 1);
                    raw = min(min(h1, h2), h3);
                }
                s.hraw[e * HPAD + pt] = static_cast<int16_t>(raw);
                const float mh = mul_rn(static_cast<float>(raw), p.vertical_scale);
                b.measured_heights[static_cast<size_t>(tile0 + e) * H + pt] = mh;
                part += sub_rn(R[2], mh);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) s.bh[e] = div_rn(part, static_cast<float>(H));
        }
        __syncthreads();
    }

    // ---- quad phase: 4 lanes per env --------------------------------------------------------------
    {
        const int e = tid >> 2, g = tid & 3;
        const bool valid = e < nvalid;
        const uint64_t genv = static_cast<uint64_t>(env_off + tile0 + e);
        const philox::Stream rng(p.seed_lo, p.seed_hi, genv, step);
        float* R = s.root + e * 13;
        const float qx = R[3], qy = R[4], qz = R[5], qw = R[6];

        // R4: counters + body-frame vectors (legged_robot.py:114-121); lane g rotates vector g
        const long long ep = s.ep[e] + 1;
        float vx, vy, vz, rx, ry, rz;
        if (g == 0) vx = R[7], vy = R[8], vz = R[9];
        else if (g == 1) vx = R[10], vy = R[11], vz = R[12];
        else vx = 0.0f, vy = 0.0f, vz = -1.0f;
        quat_rotate_inverse(qx, qy, qz, qw, vx, vy, vz, rx, ry, rz);
        const int qb = lane & ~3;
        const float blx = __shfl_sync(0xffffffffu, rx, qb), bly = __shfl_sync(0xffffffffu, ry, qb), blz = __shfl_sync(0xffffffffu, rz, qb);
        const float bax = __shfl_sync(0xffffffffu, rx, qb + 1), bay = __shfl_sync(0xffffffffu, ry, qb + 1), baz = __shfl_sync(0xffffffffu, rz, qb + 1);
        const float pgx = __shfl_sync(0xffffffffu, rx, qb + 2), pgy = __shfl_sync(0xffffffffu, ry, qb + 2), pgz = __shfl_sync(0xffffffffu, rz, qb + 2);

        // R5: command resampling + heading (legged_robot.py:343-354)
        float c0 = s.cmd[e * 4 + 0], c1 = s.cmd[e * 4 + 1], c2 = s.cmd[e * 4 + 2], c3 = s.cmd[e * 4 + 3];
        if (static_cast<int>(ep) % p.resample_steps == 0) resample_commands(p, rng, philox::CMD_PERIODIC, c0, c1, c2, c3);
        if (p.heading_command) {
            // forward = quat_apply(q, [1,0,0]) (x,y only); heading = atan2(fy, fx)
            const float tyy = 2.0f * qz, tzz = -2.0f * qy;   // t = 2*cross(q_xyz, [1,0,0]) = (0, 2qz, -2qy)
            const float fx = 1.0f + (qy * tzz - qz * tyy);
            const float fy = qw * tyy + (-qx * tzz);
            const float heading = atan2f(fy, fx);
            c2 = clampf(mul_rn(0.5f, wrap_to_pi(sub_rn(c3, heading))), -1.0f, 1.0f);
        }

        // R7: pushes (legged_robot.py:456-461)
        float rvx = R[7], rvy = R[8];
        if (do_push) {
            const uint4 w = rng.words(philox::PUSH, 0);
            rvx = affine_rn(p.push_span, philox::u01(w.x), p.push_lo);
            rvy = affine_rn(p.push_span, philox::u01(w.y), p.push_lo);
        }

        // R8: termination (legged_robot.py:139-145)
        bool term = false;
        for (int t = 0; t < p.num_term; ++t) {
            const float* F = s.contact + (e * B + p.term_idx[t]) * 3;
            term |= norm3_rn(F[0], F[1], F[2]) > 1.0f;
        }
        const bool time_out = static_cast<float>(ep) > p.max_episode_length;
        const bool reset = term | time_out;

        // R9: rewards (legged_robot.py:189-206, 918-1015); lane g owns dofs 3g..3g+2, foot g, bodies g and g+4
        const float* rs = p.reward_scale;
        float pa_rate = 0.f, pd_acc = 0.f, pd_vel = 0.f, ptq = 0.f, ppos_lim = 0.f, pvel_lim = 0.f, ptq_lim = 0.f, pstand = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int d = 3 * g + j;
            const float a = s.act[e * ND + d], la = s.lact[e * ND + d], tq = s.tq[e * ND + d];
            const float q = s.dof[e * 24 + 2 * d], qd = s.dof[e * 24 + 2 * d + 1], ldv = s.ldv[e * ND + d];
            const float da = la - a;
            pa_rate += da * da;
            const float acc = div_rn(ldv - qd, p.dt);
            pd_acc += acc * acc;
            pd_vel += qd * qd;
            ptq += tq * tq;
            ppos_lim += -fminf(q - p.dof_pos_lo[d], 0.0f) + fmaxf(q - p.dof_pos_hi[d], 0.0f);
            pvel_lim += clampf(fabsf(qd) - p.dof_vel_limits[d] * p.soft_dof_vel_limit, 0.0f, 1.0f);
            ptq_lim += fmaxf(fabsf(tq) - p.torque_limits[d] * p.soft_torque_limit, 0.0f);
            pstand += fabsf(q - p.default_dof_pos[d]);
        }
        const float cmd_norm = norm2_rn(c0, c1);
        // feet
        const float* Ff = s.contact + (e * B + p.feet_idx[g]) * 3;
        float fat = s.fat[e * 4 + g];
        uint8_t lc = s.lc[e * 4 + g];
        float p_air = 0.f;
        if (rs[T_FEET_AIR_TIME] != 0.0f) {   // the reference only mutates this state when the term is active
            const bool contact = Ff[2] > 1.0f;
            const bool filt = contact || (lc != 0);
            lc = contact ? 1 : 0;
            const bool first = (fat > 0.0f) && filt;
            fat = add_rn(fat, p.dt);
            p_air = first ? sub_rn(fat, 0.5f) : 0.0f;
            if (filt) fat = 0.0f;
        }
        const float p_stumble = (norm2_rn(Ff[0], Ff[1]) > mul_rn(5.0f, fabsf(Ff[2]))) ? 1.0f : 0.0f;
        const float p_fcf = fmaxf(sub_rn(norm3_rn(Ff[0], Ff[1], Ff[2]), p.max_contact_force), 0.0f);
        float p_coll = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float* Fp = s.contact + (e * B + p.pen_idx[g + 4 * j]) * 3;
            p_coll += (norm3_rn(Fp[0], Fp[1], Fp[2]) > 0.1f) ? 1.0f : 0.0f;
        }

        float rew = 0.0f;
        float* sums = s.sums + e;
        auto add_term = [&](int k, float val) {
            const float r = mul_rn(val, rs[k]);
            rew = add_rn(rew, r);
            if (g == 0) sums[p.sum_row[k] * TILE] = add_rn(sums[p.sum_row[k] * TILE], r);
        };
        if (rs[T_ACTION_RATE] != 0.f) add_term(T_ACTION_RATE, quad_sum(pa_rate));
        if (rs[T_ANG_VEL_XY] != 0.f) add_term(T_ANG_VEL_XY, bax * bax + bay * bay);
        if (rs[T_BASE_HEIGHT] != 0.f) {
            const float bh = ROUGH ? s.bh[e] : R[2];
            const float dh = sub_rn(bh, p.base_height_target);
            add_term(T_BASE_HEIGHT, dh * dh);
        }
        if (rs[T_COLLISION] != 0.f) add_term(T_COLLISION, quad_sum(p_coll));
        if (rs[T_DOF_ACC] != 0.f) add_term(T_DOF_ACC, quad_sum(pd_acc));
        if (rs[T_DOF_POS_LIMITS] != 0.f) add_term(T_DOF_POS_LIMITS, quad_sum(ppos_lim));
        if (rs[T_DOF_VEL] != 0.f) add_term(T_DOF_VEL, quad_sum(pd_vel));
        if (rs[T_DOF_VEL_LIMITS] != 0.f) add_term(T_DOF_VEL_LIMITS, quad_sum(pvel_lim));
        if (rs[T_FEET_AIR_TIME] != 0.f) add_term(T_FEET_AIR_TIME, (cmd_norm > 0.1f) ? quad_sum(p_air) : mul_rn(quad_sum(p_air), 0.0f));
        if (rs[T_FEET_CONTACT_FORCES] != 0.f) add_term(T_FEET_CONTACT_FORCES, quad_sum(p_fcf));
        if (rs[T_LIN_VEL_Z] != 0.f) add_term(T_LIN_VEL_Z, blz * blz);
        if (rs[T_ORIENTATION] != 0.f) add_term(T_ORIENTATION, pgx * pgx + pgy * pgy);
        if (rs[T_STAND_STILL] != 0.f) add_term(T_STAND_STILL, (cmd_norm < 0.1f) ? quad_sum(pstand) : mul_rn(quad_sum(pstand), 0.0f));
        if (rs[T_STUMBLE] != 0.f) add_term(T_STUMBLE, quad_sum(p_stumble) > 0.0f ? 1.0f : 0.0f);
        if (rs[T_TORQUE_LIMITS] != 0.f) add_term(T_TORQUE_LIMITS, quad_sum(ptq_lim));
        if (rs[T_TORQUES] != 0.f) add_term(T_TORQUES, quad_sum(ptq));
        if (rs[T_TRACKING_ANG_VEL] != 0.f) {
            const float er = sub_rn(c2, baz);
            add_term(T_TRACKING_ANG_VEL, expf(div_rn(-(er * er), p.tracking_sigma)));
        }
        if (rs[T_TRACKING_LIN_VEL] != 0.f) {
            const float ex = sub_rn(c0, blx), ey = sub_rn(c1, bly);
            add_term(T_TRACKING_LIN_VEL, expf(div_rn(-(ex * ex + ey * ey), p.tracking_sigma)));
        }
        if (p.only_positive) rew = fmaxf(rew, 0.0f);
        if (rs[T_TERMINATION] != 0.f) add_term(T_TERMINATION, (reset && !time_out) ? 1.0f : 0.0f);

        // R10: in-place reset (legged_robot.py:147-187, anymal.py:56-60); no host compaction (H7)
        float q3[3], qd3[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) q3[j] = s.dof[e * 24 + 2 * (3 * g + j)], qd3[j] = s.dof[e * 24 + 2 * (3 * g + j) + 1];
        float zpost = R[2];
        float lrv[6] = {rvx, rvy, R[9], R[10], R[11], R[12]};
        long long ep_out = ep;
        long long level = 0;
        if (p.terrain_curriculum && valid) level = b.terrain_levels[tile0 + e];
        if (reset && valid) {
            const size_t ge = static_cast<size_t>(tile0 + e);
            float ox = b.env_origins[ge * 3 + 0], oy = b.env_origins[ge * 3 + 1], oz = b.env_origins[ge * 3 + 2];
            if (p.terrain_curriculum) {   // legged_robot.py:463-486
                const float dist = norm2_rn(sub_rn(R[0], ox), sub_rn(R[1], oy));
                const bool up = dist > p.half_env_length;
                const bool down = (dist < mul_rn(mul_rn(cmd_norm, p.max_episode_length_s), 0.5f)) && !up;
                level += (up ? 1 : 0) - (down ? 1 : 0);
                if (level >= p.max_terrain_level)
                    level = philox::bounded(rng.words(philox::TERRAIN, 0).x, static_cast<uint32_t>(p.max_terrain_level));
                else if (level < 0)
                    level = 0;
                const long long type = b.terrain_types[ge];
                const float* og = b.terrain_origins + (level * p.terrain_num_cols + type) * 3;
                ox = og[0], oy = og[1], oz = og[2];
                if (g == 0) {
                    b.terrain_levels[ge] = level;
                    b.env_origins[ge * 3 + 0] = ox, b.env_origins[ge * 3 + 1] = oy, b.env_origins[ge * 3 + 2] = oz;
                }
            }
            // dofs: q = q0 * U(0.5, 1.5), qd = 0
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int d = 3 * g + j;
                const uint4 w = rng.words(philox::RESET_DOF, d >> 2);
                q3[j] = mul_rn(p.default_dof_pos[d], affine_rn(1.0f, philox::u01(philox::word(w, d & 3)), 0.5f));
                qd3[j] = 0.0f;
                s.dof[e * 24 + 2 * d] = q3[j];
                s.dof[e * 24 + 2 * d + 1] = 0.0f;
                reinterpret_cast<float2*>(b.dof_state)[ge * ND + d] = make_float2(q3[j], 0.0f);
            }
            // root
            float nr[13];
#pragma unroll
            for (int k = 0; k < 13; ++k) nr[k] = p.base_init_state[k];
            nr[0] = add_rn(nr[0], ox), nr[1] = add_rn(nr[1], oy), nr[2] = add_rn(nr[2], oz);
            if (p.custom_origins) {
                const uint4 w = rng.words(philox::RESET_XY, 0);
                nr[0] = add_rn(nr[0], affine_rn(2.0f, philox::u01(w.x), -1.0f));
                nr[1] = add_rn(nr[1], affine_rn(2.0f, philox::u01(w.y), -1.0f));
            }
            {
                const uint4 w0 = rng.words(philox::RESET_VEL, 0), w1 = rng.words(philox::RESET_VEL, 1);
                nr[7] = affine_rn(1.0f, philox::u01(w0.x), -0.5f), nr[8] = affine_rn(1.0f, philox::u01(w0.y), -0.5f);
                nr[9] = affine_rn(1.0f, philox::u01(w0.z), -0.5f), nr[10] = affine_rn(1.0f, philox::u01(w0.w), -0.5f);
                nr[11] = affine_rn(1.0f, philox::u01(w1.x), -0.5f), nr[12] = affine_rn(1.0f, philox::u01(w1.y), -0.5f);
            }
            zpost = nr[2];
#pragma unroll
            for (int k = 0; k < 6; ++k) lrv[k] = nr[7 + k];
            if (g == 0) {
#pragma unroll
                for (int k = 0; k < 13; ++k) b.root_states[ge * 13 + k] = nr[k];
            }
            resample_commands(p, rng, philox::CMD_RESET, c0, c1, c2, c3);
            fat = 0.0f;
            ep_out = 0;
            // extras["episode"] statistics: per-CTA partial sums, then one double atomic per row per CTA
            if (g == 0) {
                for (int k = 0; k < K; ++k) {
                    atomicAdd(&s.acc[k], static_cast<double>(sums[k * TILE]));
                    sums[k * TILE] = 0.0f;
                }
                atomicAdd(s.nreset, 1);
            }
            if (p.zero_lstm_on_reset) {
                // h,c: [2, N*12, 8]; this env owns 2 x (12*8) floats per array; lane g clears a quarter
                const size_t M8 = static_cast<size_t>(N) * ND * 8;
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int l = 0; l < 2; ++l) {
                    float4* hp = reinterpret_cast<float4*>(b.lstm_h + l * M8 + ge * ND * 8) + g * 6;
                    float4* cp = reinterpret_cast<float4*>(b.lstm_c + l * M8 + ge * ND * 8) + g * 6;
#pragma unroll
                    for (int k = 0; k < 6; ++k) hp[k] = z4, cp[k] = z4;
                }
            }
        } else if (do_push && valid && g == 0) {
            b.root_states[static_cast<size_t>(tile0 + e) * 13 + 7] = rvx;
            b.root_states[static_cast<size_t>(tile0 + e) * 13 + 8] = rvy;
        }
        if (p.terrain_curriculum && valid && g == 0) atomicAdd(&s.acc[K], static_cast<double>(level));

        // R11: observations + noise (legged_robot.py:208-226), clipped as in step() (:100-101)
        float o[12];
        if (g == 0) o[0] = blx * p.obs_lin_vel, o[1] = bly * p.obs_lin_vel, o[2] = blz * p.obs_lin_vel;
        else if (g == 1) o[0] = bax * p.obs_ang_vel, o[1] = bay * p.obs_ang_vel, o[2] = baz * p.obs_ang_vel;
        else if (g == 2) o[0] = pgx, o[1] = pgy, o[2] = pgz;
        else o[0] = c0 * p.obs_lin_vel, o[1] = c1 * p.obs_lin_vel, o[2] = c2 * p.obs_ang_vel;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            o[3 + j] = mul_rn(sub_rn(q3[j], p.default_dof_pos[3 * g + j]), p.obs_dof_pos);
            o[6 + j] = mul_rn(qd3[j], p.obs_dof_vel);
            o[9 + j] = s.act[e * ND + 3 * g + j];
        }
        if (p.add_noise) {
            // uniforms for columns 0..35 = Philox blocks 0..8, spread over the quad through the obs tile
            float4* stage = reinterpret_cast<float4*>(s.obs + e * 48);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int blk = g + 4 * r;
                if (blk < 9) stage[blk] = philox::u01(rng.words(philox::OBS_NOISE, blk));
            }
            __syncwarp();
            const float n0 = g == 0 ? p.noise_lin_vel : g == 1 ? p.noise_ang_vel : g == 2 ? p.noise_gravity : 0.0f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float u0 = s.obs[e * 48 + 3 * g + j], u1 = s.obs[e * 48 + 12 + 3 * g + j], u2 = s.obs[e * 48 + 24 + 3 * g + j];
                if (g < 3) o[j] = add_rn(o[j], mul_rn(sub_rn(mul_rn(2.0f, u0), 1.0f), n0));
                o[3 + j] = add_rn(o[3 + j], mul_rn(sub_rn(mul_rn(2.0f, u1), 1.0f), p.noise_dof_pos));
                o[6 + j] = add_rn(o[6 + j], mul_rn(sub_rn(mul_rn(2.0f, u2), 1.0f), p.noise_dof_vel));
            }
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 3; ++j) s.obs[e * 48 + 12 * k + 3 * g + j] = clampf(o[3 * k + j], -p.clip_obs, p.clip_obs);

        // R12 + state write-back into the tile
#pragma unroll
        for (int j = 0; j < 3; ++j) s.ldv[e * ND + 3 * g + j] = qd3[j];
        s.fat[e * 4 + g] = fat;
        s.lc[e * 4 + g] = lc;
        if (g == 0) {
            s.blv[e * 3 + 0] = blx, s.blv[e * 3 + 1] = bly, s.blv[e * 3 + 2] = blz;
            s.cmd[e * 4 + 0] = c0, s.cmd[e * 4 + 1] = c1, s.cmd[e * 4 + 2] = c2, s.cmd[e * 4 + 3] = c3;
            s.ep[e] = ep_out;
            s.rew[e] = rew;
            s.reset[e] = reset ? 1 : 0;
            s.tout[e] = time_out ? 1 : 0;
        } else if (g == 1) {
            s.bav[e * 3 + 0] = bax, s.bav[e * 3 + 1] = bay, s.bav[e * 3 + 2] = baz;
#pragma unroll
            for (int k = 0; k < 6; ++k) s.lrv[e * 6 + k] = lrv[k];
        } else if (g == 2) {
            s.pg[e * 3 + 0] = pgx, s.pg[e * 3 + 1] = pgy, s.pg[e * 3 + 2] = pgz;
            if (ROUGH) s.zpost[e] = zpost;
        }
    }
    fence_proxy_async();   // make the generic-proxy smem writes visible to the bulk-store engine
    __syncthreads();

    // ---- write the tile back ----------------------------------------------------------------------
    const bool obs_bulk = full && (O == 48);
    if (full) {
        if (tid == 0) {
            bulk_s2g(b.last_actions + static_cast<size_t>(tile0) * ND, s.act, TILE * ND * 4);
            bulk_s2g(b.last_dof_vel + static_cast<size_t>(tile0) * ND, s.ldv, TILE * ND * 4);
            bulk_s2g(b.last_root_vel + static_cast<size_t>(tile0) * 6, s.lrv, TILE * 6 * 4);
            bulk_s2g(b.commands + static_cast<size_t>(tile0) * 4, s.cmd, TILE * 4 * 4);
            bulk_s2g(b.feet_air_time + static_cast<size_t>(tile0) * 4, s.fat, TILE * 4 * 4);
            bulk_s2g(b.last_contacts + static_cast<size_t>(tile0) * 4, s.lc, TILE * 4);
            bulk_s2g(b.episode_length_buf + tile0, s.ep, TILE * 8);
            bulk_s2g(b.reset_buf + tile0, s.reset, TILE);
            bulk_s2g(b.time_out_buf + tile0, s.tout, TILE);
            bulk_s2g(b.rew_buf + tile0, s.rew, TILE * 4);
            bulk_s2g(b.base_lin_vel + static_cast<size_t>(tile0) * 3, s.blv, TILE * 3 * 4);
            bulk_s2g(b.base_ang_vel + static_cast<size_t>(tile0) * 3, s.bav, TILE * 3 * 4);
            bulk_s2g(b.projected_gravity + static_cast<size_t>(tile0) * 3, s.pg, TILE * 3 * 4);
            for (int k = 0; k < K; ++k) bulk_s2g(b.episode_sums + static_cast<size_t>(k) * N + tile0, s.sums + k * TILE, TILE * 4);
            if (obs_bulk) bulk_s2g(b.obs_buf + static_cast<size_t>(tile0) * 48, s.obs, TILE * 48 * 4);
            bulk_commit();
        }
    } else {
        coop_copy(b.last_actions + static_cast<size_t>(tile0) * ND, s.act, nvalid * ND);
        coop_copy(b.last_dof_vel + static_cast<size_t>(tile0) * ND, s.ldv, nvalid * ND);
        coop_copy(b.last_root_vel + static_cast<size_t>(tile0) * 6, s.lrv, nvalid * 6);
        coop_copy(b.commands + static_cast<size_t>(tile0) * 4, s.cmd, nvalid * 4);
        coop_copy(b.feet_air_time + static_cast<size_t>(tile0) * 4, s.fat, nvalid * 4);
        coop_copy(b.last_contacts + static_cast<size_t>(tile0) * 4, s.lc, nvalid * 4);
        coop_copy(reinterpret_cast<long long*>(b.episode_length_buf) + tile0, s.ep, nvalid);
        coop_copy(b.reset_buf + tile0, s.reset, nvalid);
        coop_copy(b.time_out_buf + tile0, s.tout, nvalid);
        coop_copy(b.rew_buf + tile0, s.rew, nvalid);
        coop_copy(b.base_lin_vel + static_cast<size_t>(tile0) * 3, s.blv, nvalid * 3);
        coop_copy(b.base_ang_vel + static_cast<size_t>(tile0) * 3, s.bav, nvalid * 3);
        coop_copy(b.projected_gravity + static_cast<size_t>(tile0) * 3, s.pg, nvalid * 3);
        for (int k = 0; k < K; ++k) coop_copy(b.episode_sums + static_cast<size_t>(k) * N + tile0, s.sums + k * TILE, nvalid);
    }
    if (!obs_bulk) {
        for (int i = tid; i < nvalid * 48; i += TILE * LPE) b.obs_buf[static_cast<size_t>(tile0 + i / 48) * O + (i % 48)] = s.obs[i];
    }

    // ---- rough: height observations, one warp per env (legged_robot.py:220-226) -------------------
    if (ROUGH) {
        const int H = p.num_heights;
        float* stage = s.stage + warp * HPAD;
        for (int e = warp; e < nvalid; e += TILE * LPE / 32) {
            const float z = s.zpost[e];
            const philox::Stream rng(p.seed_lo, p.seed_hi, static_cast<uint64_t>(env_off + tile0 + e), step);
            for (int pb = lane; pb * 4 < H; pb += 32) {
                float4 u = make_float4(0.5f, 0.5f, 0.5f, 0.5f);
                if (p.add_noise) u = philox::u01(rng.words(philox::OBS_NOISE, 12 + pb));
                const float uu[4] = {u.x, u.y, u.z, u.w};
                float out[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int pt = 4 * pb + j;
                    const float mh = mul_rn(static_cast<float>(pt < H ? s.hraw[e * HPAD + pt] : 0), p.vertical_scale);
                    float v = mul_rn(clampf(sub_rn(sub_rn(z, 0.5f), mh), -1.0f, 1.0f), p.obs_height);
                    if (p.add_noise) v = add_rn(v, mul_rn(sub_rn(mul_rn(2.0f, uu[j]), 1.0f), p.noise_height));
                    out[j] = clampf(v, -p.clip_obs, p.clip_obs);
                }
                *reinterpret_cast<float4*>(stage + 4 * pb) = make_float4(out[0], out[1], out[2], out[3]);
            }
            __syncwarp();
            float* dst = b.obs_buf + static_cast<size_t>(tile0 + e) * O + 48;
            for (int pt = lane; pt < H; pt += 32) dst[pt] = stage[pt];
            __syncwarp();
        }
    }

    // ---- extras["episode"]: cross-CTA reduction, finalised by the last CTA to arrive ---------------
    __syncthreads();
    const int nreset = *s.nreset;
    if (nreset > 0 && tid < K) atomicAdd(&b.ws_sums[tid], s.acc[tid]);
    if (p.terrain_curriculum && tid == K) atomicAdd(&b.ws_sums[K], s.acc[K]);
    if (nreset > 0 && tid == K + 1) atomicAdd(&b.ws_sums[K + 1], static_cast<double>(nreset));
    __threadfence();
    __syncthreads();
    __shared__ unsigned int s_ticket;
    if (tid == 0) s_ticket = atomicAdd(b.ws_counter, 1u);
    __syncthreads();
    if (s_ticket == gridDim.x - 1) {
        __threadfence();
        volatile double* ws = b.ws_sums;
        const double cnt = ws[K + 1];
        if (tid < K && cnt > 0.0)
            b.extras_out[tid] = div_rn(static_cast<float>(ws[tid] / cnt), p.max_episode_length_s);
        if (tid == K && cnt > 0.0) b.extras_out[K] = static_cast<float>(ws[K] / static_cast<double>(N));
        if (tid == K + 1) b.extras_out[K + 1] = static_cast<float>(cnt);
        __syncthreads();
        if (tid < K + 2) b.ws_sums[tid] = 0.0;
        if (tid == 0) *b.ws_counter = 0u;
    }
    if (full && tid == 0) bulk_wait_read0();   // smem must stay alive until the bulk stores have read it
}

template <int TILE, bool ROUGH>
int launch_post_physics(const B200LeggedParams& p, const B200LeggedBuffers& b, uint64_t step, long long env_off, int do_push,
                        cudaStream_t stream) {
    const size_t smem = carve_tile<TILE, ROUGH>(nullptr, p.num_bodies, p.num_sum_rows).bytes;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(post_physics_kernel<TILE, ROUGH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "post_physics: cannot reserve %zu B of shared memory: %s", smem,
                     cudaGetErrorString(e));
        configured = smem;
    }
    const int grid = (p.num_envs + TILE - 1) / TILE;
    post_physics_kernel<TILE, ROUGH><<<grid, TILE * LPE, smem, stream>>>(p, b, step, env_off, do_push);
    B200_LAUNCH_CHECK("post_physics");
    return B200GYM_OK;
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
    for (int r = 0; r < 32; ++r) {
        for (int k = 0; k < 2; ++k) n.w_ih0[r][k] = w_ih0[r * 2 + k];
        for (int k = 0; k < 8; ++k) n.w_hh0[r][k] = w_hh0[r * 8 + k], n.w_ih1[r][k] = w_ih1[r * 8 + k], n.w_hh1[r][k] = w_hh1[r * 8 + k];
        n.b0[r] = b_ih0[r], n.bh0[r] = b_hh0[r], n.b1[r] = b_ih1[r], n.bh1[r] = b_hh1[r];
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

int b200gym_post_physics(const B200LeggedParams* p, const B200LeggedBuffers* b, uint64_t step, int64_t env_id_offset,
                         void* stream) {
    B200_REQUIRE(p && b, B200GYM_EINVAL, "post_physics: null argument");
    B200_REQUIRE(p->num_envs > 0, B200GYM_EINVAL, "post_physics: num_envs must be positive (got %d)", p->num_envs);
    B200_REQUIRE(p->num_bodies > 0 && p->num_bodies <= 64, B200GYM_EINVAL, "post_physics: num_bodies %d out of range", p->num_bodies);
    B200_REQUIRE(p->num_sum_rows >= 0 && p->num_sum_rows <= B200GYM_NUM_REWARD_TERMS, B200GYM_EINVAL, "post_physics: bad num_sum_rows");
    B200_REQUIRE(p->resample_steps > 0, B200GYM_EINVAL, "post_physics: resample_steps must be positive");
    const bool rough = p->num_heights > 0;
    B200_REQUIRE(p->num_obs == 48 + p->num_heights, B200GYM_EINVAL, "post_physics: num_obs %d != 48 + %d heights", p->num_obs,
                 p->num_heights);
    B200_REQUIRE(!rough || (p->num_heights <= HPAD && p->n_px * p->n_py == p->num_heights && p->n_px <= B200GYM_MAX_POINTS &&
                            p->n_py <= B200GYM_MAX_POINTS),
                 B200GYM_EINVAL, "post_physics: unsupported height grid %dx%d", p->n_px, p->n_py);
    B200_REQUIRE(!rough || b->measured_heights, B200GYM_EINVAL, "post_physics: measured_heights buffer missing");
    B200_REQUIRE(!rough || p->mesh_plane || (b->height_samples && p->terrain_rows >= 2 && p->terrain_cols >= 2), B200GYM_EINVAL,
                 "post_physics: height_samples missing");
    B200_REQUIRE(!p->terrain_curriculum || (b->terrain_levels && b->terrain_types && b->terrain_origins), B200GYM_EINVAL,
                 "post_physics: terrain curriculum buffers missing");
    B200_REQUIRE(!p->zero_lstm_on_reset || (b->lstm_h && b->lstm_c), B200GYM_EINVAL, "post_physics: LSTM state buffers missing");
    const void* must[] = {b->root_states, b->dof_state, b->contact_forces, b->actions, b->torques, b->last_actions, b->last_dof_vel,
                          b->last_root_vel, b->commands, b->feet_air_time, b->last_contacts, b->episode_length_buf, b->reset_buf,
                          b->time_out_buf, b->rew_buf, b->obs_buf, b->base_lin_vel, b->base_ang_vel, b->projected_gravity,
                          b->env_origins, b->extras_out, b->ws_sums, b->ws_counter};
    for (const void* q : must) {
        B200_REQUIRE(q != nullptr, B200GYM_EINVAL, "post_physics: null buffer");
        B200_REQUIRE(b200_aligned16(q), B200GYM_EALIGN, "post_physics: buffers must be 16-byte aligned");
    }
    B200_REQUIRE(p->num_sum_rows == 0 || (b->episode_sums && b200_aligned16(b->episode_sums) && p->num_envs % 4 == 0) ||
                     (b->episode_sums && p->num_envs < 64),
                 B200GYM_EALIGN, "post_physics: episode_sums rows must be 16-byte aligned (num_envs %% 4 == 0)");
    const int do_push = (p->push_robots && p->push_time > 0 && (step % static_cast<uint64_t>(p->push_time) == 0)) ? 1 : 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (rough) return launch_post_physics<64, true>(*p, *b, step, env_id_offset, do_push, st);
    return launch_post_physics<64, false>(*p, *b, step, env_id_offset, do_push, st);
}

}  // extern "C"
