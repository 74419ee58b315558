// post_physics_kernel — LeggedRobot.post_physics_step (legged_robot.py:106-134) and everything it calls,
// plus the observation clip of step() (:100-101), as ONE pass over per-env state.
//
// Data movement.  Every API-visible tensor keeps the reference's row-major [N, k] layout (plain contiguous
// torch tensors; PhysX-owned buffers consumed as given).  A CTA owns a tile of TILE consecutive envs, whose
// rows form ONE contiguous byte range per tensor; each range moves HBM<->shared memory as a 1-D TMA bulk
// copy (cp.async.bulk, SASS UBLKCP), issued by one thread and tracked by an mbarrier.  No thread ever
// computes a global address for the streaming data, and every DRAM sector is touched exactly once.
//
// Work mapping (chosen from the ncu profile of the first version, profiles/r1_post_physics_v1.md: the kernel
// was issue-bound because four lanes each repeated the per-env scalar work):
//   phase H  (rough only) one warp per env pair, one lane per sample point: 187-point height scan + height
//            observations (noise included)
//                                                                            legged_robot.py:877-915,220-226
//   phase W  4 lanes per env: per-DOF / per-foot / per-body terms, the DOF observation columns and all
//            Philox noise draws; 2-step shuffle reductions leave per-env partial sums in shared memory
//   phase S  1 thread per env: body-frame vectors, commands, termination, reward assembly, in-place reset,
//            first 12 observation columns — scalar work executed exactly once per env
//   phase H' (rough only) height observations redone for the (rare) envs that reset: they use the post-reset height
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "philox.cuh"
#include "../../include/b200gym.h"

#ifndef PP_SQNORM
#define PP_SQNORM 1        // threshold compares on squared norms with exact (host-computed) constants instead of sqrtf
#endif
#ifndef PP_ALIAS_OBS
#define PP_ALIAS_OBS 1     // the observation tile reuses the contact-force tile (dead after the contact pass of phase W)
#endif
#ifndef PP_MINBLOCKS_FLAT
#define PP_MINBLOCKS_FLAT 10   // flat, non-trajectory kernel: 48 registers (no spills) and 22.0 KB of shared memory -> 10 CTAs per SM
#endif
#ifndef PP_MINBLOCKS
#define PP_MINBLOCKS 9     // register cap so that 9 CTAs of 128 threads are resident per SM (A/B: profiles/r1_post_physics_ab.txt)
#endif
#ifndef PP_FINALIZE_KERNEL
#define PP_FINALIZE_KERNEL 1   // extras["episode"] finalised by a 1-warp follow-up kernel instead of a last-CTA ticket
#endif

namespace {

constexpr int LPE = 4;     // lanes per env in phase W

// sqrtf(x) > t  <=>  x > gt(t);  sqrtf(x) < t  <=>  x < lt(t)   (exact in fp32; computed on the host by sq_thresholds())
struct SqThr { float gt1, gt01, gt02, lt01; };
constexpr int ND = B200GYM_NUM_DOF;
constexpr int HPAD = 192;  // padded per-env stride of the raw height tile (>= 187)
// Terrain window of one env staged by ONE 2-D TMA box load (cp.async.bulk.tensor.2d, SASS UTMALDG.2D): the 17 x 11 sample grid lies within
// 0.95 m of the base, i.e. within +-10 cells at the shipped 0.1 m resolution; 22 rows (+1 for the min-of-3 neighbour) x 32 columns (the
// window's first column is rounded down to a multiple of 8: the box origin of the innermost dimension must be 16-byte aligned — an
// unaligned origin raises "illegal instruction", tools/probes/tma_probe_bisect.cu) = 1408 B = 11 x 128 B.
constexpr int HW_ROWS = 22, HW_COLS = 32, HW_ELEMS = HW_ROWS * HW_COLS, HW_HALF = 10;
// LeggedRobotTrajectory (legged_robot_trajectory.py:274-287): the 3 command columns become the N x rom.n trajectory block
constexpr int TRAJ_W = B200GYM_TRAJ_WIDTH;
template <bool TRAJ> struct ObsLayout {
    static constexpr int CW = TRAJ ? TRAJ_W : 3;       // command / trajectory block
    static constexpr int DOF = 9 + CW;                  // first dof_pos column
    static constexpr int OW = DOF + 3 * ND;             // columns before the height block (48 | 65)
    static constexpr int NB = (DOF + 2 * ND + 3) / 4;   // Philox blocks covering the noisy columns [0, DOF + 24)
    static constexpr int HB0 = OW / 4, HSH = OW % 4;    // height column j -> block HB0 + (j + HSH) / 4, word (j + HSH) % 4
};

enum Term {
    T_ACTION_RATE = 0, T_ANG_VEL_XY, T_BASE_HEIGHT, T_COLLISION, T_DIFFERENTIAL_ERROR, T_DOF_ACC, T_DOF_POS_LIMITS, T_DOF_VEL,
    T_DOF_VEL_LIMITS, T_FEET_AIR_TIME, T_FEET_CONTACT_FORCES, T_LIN_VEL_Z, T_ORIENTATION, T_STAND_STILL,
    T_STUMBLE, T_TORQUE_LIMITS, T_TORQUES, T_TRACKING_ANG_VEL, T_TRACKING_LIN_VEL, T_TRACKING_ROM, T_TERMINATION
};
// per-env partial sums handed from phase W to phase S
enum Part { P_ACTION_RATE = 0, P_DOF_ACC, P_DOF_VEL, P_TORQUES, P_POS_LIM, P_VEL_LIM, P_TQ_LIM, P_STAND, P_AIR, P_STUMBLE, P_FCF,
            P_COLL, P_TERM, NUM_PARTS };

__device__ __forceinline__ void quat_rotate_inverse(float qx, float qy, float qz, float qw, float vx, float vy, float vz,
                                                    float& rx, float& ry, float& rz) {
    // isaacgym.torch_utils.quat_rotate_inverse: a = v*(2w^2-1); b = cross(q,v)*w*2; c = q*dot(q,v)*2; a - b + c
    const float s = 2.0f * qw * qw - 1.0f;
    const float cx = qy * vz - qz * vy, cy = qz * vx - qx * vz, cz = qx * vy - qy * vx;
    const float dt2 = (qx * vx + qy * vy + qz * vz) * 2.0f, w2 = qw * 2.0f;
    rx = vx * s - cx * w2 + qx * dt2;
    ry = vy * s - cy * w2 + qy * dt2;
    rz = vz * s - cz * w2 + qz * dt2;
}

// norms that feed a threshold compare keep torch's un-fused rounding
__device__ __forceinline__ float norm2_rn(float x, float y) { return sqrtf(add_rn(mul_rn(x, x), mul_rn(y, y))); }
__device__ __forceinline__ float norm3_rn(float x, float y, float z) {
    return sqrtf(add_rn(add_rn(mul_rn(x, x), mul_rn(y, y)), mul_rn(z, z)));
}

__device__ __forceinline__ float sq2_rn(float x, float y) { return add_rn(mul_rn(x, x), mul_rn(y, y)); }
__device__ __forceinline__ float sq3_rn(float x, float y, float z) { return add_rn(add_rn(mul_rn(x, x), mul_rn(y, y)), mul_rn(z, z)); }

// Height-scan cell index of TWO sample points at once (legged_robot.py:896-907 with quat_apply_yaw, math.py:38-42), every
// operation a separately rounded fp32 mul/add exactly as torch evaluates it, issued as packed f32x2 instructions:
//   t = 2 * cross(q_yaw.xyz, p) = (-2 qz hy, 2 qz hx);  w = p + qw t + cross(q_yaw.xyz, t);  cell = trunc((w + base + border) / hs)
struct HeightConsts { float2 border, neg_hs, rcp_hs; };
template <bool PACKED>
__device__ __forceinline__ void height_cells2(float2 hx, float2 hy, float qz, float qw, float bx, float by, const HeightConsts& c,
                                              int rows, int cols, int& pk0, int& pk1) {
    const float2 qz2 = make_float2(qz, qz), nqz2 = make_float2(-qz, -qz), qw2 = make_float2(qw, qw);
    const float2 tx = __fmul2_rn(__fmul2_rn(qz2, hy), make_float2(-2.0f, -2.0f));   // == (-(qz*hy)) * 2
    const float2 ty = __fmul2_rn(__fmul2_rn(qz2, hx), make_float2(2.0f, 2.0f));
    float2 wx = __fadd2_rn(__fadd2_rn(hx, __fmul2_rn(qw2, tx)), __fmul2_rn(nqz2, ty));   // (hx + qw*tx) + (-(qz*ty))
    float2 wy = __fadd2_rn(__fadd2_rn(hy, __fmul2_rn(qw2, ty)), __fmul2_rn(qz2, tx));
    wx = __fadd2_rn(__fadd2_rn(wx, make_float2(bx, bx)), c.border);
    wy = __fadd2_rn(__fadd2_rn(wy, make_float2(by, by)), c.border);
    // correctly rounded x / hs for the loop-invariant cell size with rcp = RN(1/hs): q0 = RN(x*rcp); r = x - q0*hs (exact, FMA);
    // q = RN(q0 + r*rcp) (Markstein) — two lanes at a time; the parity tests compare the cells bit for bit
    const float2 qx0 = __fmul2_rn(wx, c.rcp_hs), qy0 = __fmul2_rn(wy, c.rcp_hs);
    wx = __ffma2_rn(__ffma2_rn(qx0, c.neg_hs, wx), c.rcp_hs, qx0);
    wy = __ffma2_rn(__ffma2_rn(qy0, c.neg_hs, wy), c.rcp_hs, qy0);
    // .long() truncation + clip (legged_robot.py:903-907); the saturating conversion keeps huge values clipped
    const int ix0 = min(max(__float2int_rz(wx.x), 0), rows - 2), iy0 = min(max(__float2int_rz(wy.x), 0), cols - 2);
    const int ix1 = min(max(__float2int_rz(wx.y), 0), rows - 2), iy1 = min(max(__float2int_rz(wy.y), 0), cols - 2);
    if (PACKED) {   // (row, column) of the cell, 16 bits each (the launcher checks rows, cols < 32768): the TMA path indexes a window
        pk0 = (ix0 << 16) | iy0;
        pk1 = (ix1 << 16) | iy1;
    } else {        // element offset into the field: the gather path
        pk0 = ix0 * cols + iy0;
        pk1 = ix1 * cols + iy1;
    }
}

__device__ __forceinline__ float wrap_to_pi(float a) {
    // legged_gym/utils/math.py:45-48 on fp32 tensors: python-style remainder by fl32(2*pi), then -2*pi where > fl32(pi)
    const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
    float r = fmodf(a, two_pi);
    if (r < 0.0f) r = add_rn(r, two_pi);
    if (r > pi) r = sub_rn(r, two_pi);
    return r;
}

__device__ __forceinline__ void resample_commands(const B200LeggedParams& p, const SqThr& thr, const philox::Stream& rng, uint32_t site,
                                                  float& c0, float& c1, float& c2, float& c3) {
    // legged_robot.py:365-387
    const uint4 w = rng.words(site, 0);
    const float* lo = site == philox::CMD_RESET ? p.cmd_lo_reset : p.cmd_lo;       // differ only while the command curriculum advances
    const float* span = site == philox::CMD_RESET ? p.cmd_span_reset : p.cmd_span;
    c0 = affine_rn(span[0], philox::u01(w.x), lo[0]);
    c1 = affine_rn(span[1], philox::u01(w.y), lo[1]);
    if (p.heading_command)
        c3 = affine_rn(span[3], philox::u01(w.z), lo[3]);
    else
        c2 = affine_rn(span[2], philox::u01(w.z), lo[2]);
#if PP_SQNORM
    const float m = sq2_rn(c0, c1) > thr.gt02 ? 1.0f : 0.0f;
#else
    const float m = norm2_rn(c0, c1) > 0.2f ? 1.0f : 0.0f;
#endif
    c0 *= m;
    c1 *= m;
}

__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// noise term of legged_robot.py:226: (2u - 1) * scale, added to the observation
__device__ __forceinline__ float add_noise(float v, float u, float scale) { return fmaf(fmaf(2.0f, u, -1.0f), scale, v); }

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
    uint64_t* bar;
    float *root, *dof, *contact, *act, *tq, *lact, *ldv, *cmd, *fat;
    uint8_t* lc;
    long long* ep;
    float *sums, *obs, *blv, *bav, *pg, *lrv, *rew, *part, *part_b, *part_c;   // part rows 0-1 / 2-7 / 8-12 (see carve_tile)
    uint8_t *reset, *tout;
    int16_t* hraw;
    float *bh, *zpost, *stage, *hsum, *unoise;
    float *traj, *perr, *tpush;
    float2 *pts, *yaw;
    int16_t* hwin;     // [warps][2][HW_ROWS][HW_COLS] terrain windows of the env pair a warp is scanning (TMA destination, 128-byte aligned)
    uint64_t* hbar;    // one mbarrier per warp
    double* acc;
    int* nreset;
    size_t bytes;
};

// row k (compile-time) of the per-env partial sums, env e
template <int TILE>
__device__ __forceinline__ float& part_at(const TileSmem& s, int k, int e) {
    return k < 2 ? s.part[k * TILE + e] : k < 8 ? s.part_b[(k - 2) * TILE + e] : s.part_c[(k - 8) * TILE + e];
}
#define PART(k) part_at<TILE>(s, (k), e)

template <int TILE, bool ROUGH, bool TRAJ>
__device__ __host__ inline TileSmem carve_tile(unsigned char* base, int B, int K, bool need_hpart, bool tma_heights = false) {
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
    // the obs tile reuses the contact tile when it fits (B*3 >= OW): contact forces are dead after phase W's contact pass
    constexpr int OW = ObsLayout<TRAJ>::OW;
    // (trajectory layout, 65 columns: the dof tile sits right in front of the contact tile and is dead at the same barrier, so the
    //  pair [dof | contact] = 24 + 3B floats per env hosts the observation tile)
    static_assert((TILE * 24 * sizeof(float)) % 16 == 0, "dof and contact tiles must be contiguous");
    s.obs = (PP_ALIAS_OBS && B * 3 >= OW) ? s.contact : (PP_ALIAS_OBS && TRAJ && 24 + B * 3 >= OW) ? s.dof : c.take<float>(TILE * OW);
    s.traj = TRAJ ? c.take<float>(TILE * TRAJ_W) : nullptr;
    s.perr = TRAJ ? c.take<float>(TILE * 2) : nullptr;
    s.tpush = TRAJ ? c.take<float>(TILE) : nullptr;
    // The torque and last-action tiles are only read in phase W; the per-env outputs of phase S (behind the barrier that ends phase W)
    // are staged on top of them: 64 B per env less, which is what lets a 10th CTA of the flat kernel fit on an SM (131 072 envs = 4096
    // tiles then take 3 rounds of 1480 resident CTAs instead of 4 rounds of 1332)
    static_assert((TILE * 3 * sizeof(float)) % 16 == 0 && (TILE * sizeof(float)) % 16 == 0, "aliased output tiles must stay 16-byte aligned");
    s.blv = s.tq;
    s.bav = s.tq + TILE * 3;
    s.pg = s.tq + TILE * 6;
    s.rew = s.tq + TILE * 9;
    s.lrv = s.lact;
    // Phase W's per-env partial sums (NUM_PARTS rows of TILE floats, written after the barrier that retires the contact tile, read in phase
    // S).  Trajectory layout: they fit in what is dead by then and not claimed by the phase-S outputs above — the last 2 rows of the torque
    // tile, the last 6 of the last-action tile and the 10 rows behind the 65-column observation tile in [dof | contact] — which brings
    // the trajectory kernel from 8 to 9 resident CTAs per SM.  Otherwise one block of their own.
    static_assert(NUM_PARTS == 13, "part rows are split 2 + 6 + 5");
    if (PP_ALIAS_OBS && TRAJ && B * 3 < OW && 24 + B * 3 >= OW + 5) {
        s.part = s.tq + TILE * 10;
        s.part_b = s.lact + TILE * 6;
        s.part_c = s.dof + TILE * OW;
    } else {
        s.part = c.take<float>(NUM_PARTS * TILE);
        s.part_b = s.part + TILE * 2;
        s.part_c = s.part + TILE * 8;
    }
    s.reset = c.take<uint8_t>(TILE);
    s.tout = c.take<uint8_t>(TILE);
    s.acc = c.take<double>(B200GYM_NUM_REWARD_TERMS + 2);
    s.nreset = c.take<int>(4);
    if (ROUGH) {
        s.hraw = nullptr;
        s.bh = nullptr;
        s.zpost = c.take<float>(TILE);
        s.stage = nullptr;
        s.yaw = c.take<float2>(TILE);
        s.hsum = need_hpart ? c.take<float>(TILE) : nullptr;
        s.unoise = c.take<float>(static_cast<size_t>(TILE * LPE / 32) * 2 * HPAD);   // per warp: noise uniforms of an env pair
        s.pts = c.take<float2>(HPAD);
        s.hwin = nullptr, s.hbar = nullptr;
        if (tma_heights) {
            s.hbar = c.take<uint64_t>(TILE * LPE / 32);
            c.off = (c.off + 127) & ~static_cast<size_t>(127);
            s.hwin = c.take<int16_t>(static_cast<size_t>(TILE * LPE / 32) * 2 * HW_ELEMS);
        }
    } else {
        s.hwin = nullptr, s.hbar = nullptr;
        s.hraw = nullptr;
        s.bh = s.zpost = s.stage = s.hsum = s.unoise = nullptr;
        s.pts = nullptr;
        s.yaw = nullptr;
    }
    s.bytes = c.off;
    return s;
}

template <typename T>
__device__ __forceinline__ void coop_copy(T* dst, const T* src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

template <int TILE, bool ROUGH, bool TRAJ, bool HTMA = false>
__global__ void __launch_bounds__(TILE* LPE, (TILE == 32 ? ((ROUGH || TRAJ) ? PP_MINBLOCKS : PP_MINBLOCKS_FLAT) : 1)) post_physics_kernel(const __grid_constant__ B200LeggedParams p,
                                                                               const __grid_constant__ B200LeggedBuffers b,
                                                                               uint64_t step, long long env_off, int do_push,
                                                                               const SqThr thr, const __grid_constant__ CUtensorMap hmap) {
    constexpr int use_hmap = (ROUGH && HTMA) ? 1 : 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int B = p.num_bodies, K = p.num_sum_rows, N = p.num_envs, O = p.num_obs;
    const TileSmem s = carve_tile<TILE, ROUGH, TRAJ>(smem_raw, B, K, p.reward_scale[T_BASE_HEIGHT] != 0.0f, use_hmap != 0);
    const int tid = threadIdx.x;
    const int tile0 = blockIdx.x * TILE;
    const int nvalid = min(TILE, N - tile0);
    const bool full = (nvalid == TILE);
    const float* rs = p.reward_scale;
    using OL = ObsLayout<TRAJ>;
    constexpr int OW = OL::OW;

    pdl_launch_dependents();
    pdl_wait();   // everything below reads what the torque kernels / the previous step wrote (incl. the device step counter)
    if (tid == 0) {
        mbar_init(s.bar, 1);
        if (ROUGH && use_hmap)
            for (int w = 0; w < TILE * LPE / 32; ++w) mbar_init(s.hbar + w, 1);
        fence_mbar_init();
        s.nreset[0] = 0;
        s.nreset[1] = do_push;
        if (b.step_counter) {   // graph-replayable mode: step and push schedule come from device memory
            const unsigned long long sc = *b.step_counter;
            s.nreset[1] = (p.push_robots && p.push_time > 0 && (sc % static_cast<unsigned long long>(p.push_time) == 0)) ? 1 : 0;
        }
    }
    if (b.step_counter) step = *b.step_counter;
    if (tid < B200GYM_NUM_REWARD_TERMS + 2) s.acc[tid] = 0.0;
    if (ROUGH) {   // base-frame sample points (legged_robot.py:861-875), tabulated once per CTA
        for (int pt = tid; pt < HPAD; pt += TILE * LPE)
            s.pts[pt] = pt < p.num_heights ? make_float2(p.points_x[pt / p.n_py], p.points_y[pt % p.n_py]) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    do_push = s.nreset[1];

    // ---- stage the tile: one bulk copy per tensor -------------------------------------------------
    if (full) {
        if (tid == 0) {
            const uint32_t bytes = TILE * (13 + 24 + 3 * B + 4 * ND + 4 + 4) * 4 + TILE * 4 + TILE * 8 + K * TILE * 4 +
                                   (TRAJ ? TILE * (TRAJ_W + 2 + 1) * 4 : 0);
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
            if (TRAJ) {
                bulk_g2s(s.traj, b.trajectory + static_cast<size_t>(tile0) * TRAJ_W, TILE * TRAJ_W * 4, s.bar);
                bulk_g2s(s.perr, b.prev_error + static_cast<size_t>(tile0) * 2, TILE * 2 * 4, s.bar);
                bulk_g2s(s.tpush, b.time_until_next_push + tile0, TILE * 4, s.bar);
            }
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
        if (TRAJ) {
            coop_copy(s.traj, b.trajectory + static_cast<size_t>(tile0) * TRAJ_W, nvalid * TRAJ_W);
            coop_copy(s.perr, b.prev_error + static_cast<size_t>(tile0) * 2, nvalid * 2);
            coop_copy(s.tpush, b.time_until_next_push + tile0, nvalid);
        }
        for (int k = 0; k < K; ++k) coop_copy(s.sums + k * TILE, b.episode_sums + static_cast<size_t>(k) * N + tile0, nvalid);
        __syncthreads();
    }

    // ---- phase H: height scan + height observations (legged_robot.py:877-915, math.py:38-42, :220-226) ----------------
    // A warp takes two envs at a time; a lane owns the sample points {lane, lane+32, ...} of the env, so one gather instruction
    // covers 32 CONSECUTIVE points (~3 columns of the 17 x 11 grid = few heightfield rows) and the measured_heights / height
    // observation stores are 128-B coalesced.  (The previous mapping, one lane per quad of points, spread every gather over
    // ~14 cache lines and was bound by the L1 tag rate: profiles/r1_post_physics_rough.md.)  The index chain keeps torch's
    // un-fused fp32 rounding (the cell index depends on it) but runs two points per instruction (f32x2).
    // Noise: Philox block q holds the draws of points 4q..4q+3, so lanes draw whole blocks and hand the uniforms over
    // through a per-warp scratch.  The observation needs the POST-reset base height; resets are rare, so it is produced
    // here with the current height and redone after phase S only for the envs that did reset (phase H').
    if (ROUGH) {
        if (tid < nvalid) {   // quat_apply_yaw: zero x,y, renormalise (un-fused fp32), once per env
            const float* R = s.root + tid * 13;
            const float nq = fmaxf(sqrtf(add_rn(mul_rn(R[5], R[5]), mul_rn(R[6], R[6]))), 1e-9f);
            s.yaw[tid] = make_float2(div_rn(R[5], nq), div_rn(R[6], nq));
        }
        __syncthreads();
        const int H = p.num_heights, Q = (H + OL::HSH + 3) >> 2;   // Philox blocks holding the height-noise draws
        const int rows = p.terrain_rows, cols = p.terrain_cols;
        const float inv_hs = div_rn(1.0f, p.horizontal_scale);
        const HeightConsts hc = {make_float2(p.border_size, p.border_size), make_float2(-p.horizontal_scale, -p.horizontal_scale),
                                 make_float2(inv_hs, inv_hs)};
        constexpr int NW = TILE * LPE / 32, RND = HPAD / 32;
        const int warp = tid >> 5, lane = tid & 31;
        float* un = s.unoise + warp * 2 * HPAD;
        constexpr bool tma_h = use_hmap != 0;   // the launcher selects HTMA only for a heightfield (never mesh_plane)
        int16_t* win = tma_h ? s.hwin + warp * 2 * HW_ELEMS : nullptr;
        uint32_t hphase = 0;
        for (int e0 = 2 * warp; e0 < nvalid; e0 += 2 * NW) {
            const int ne = min(2, nvalid - e0);
            int wx0[2] = {0, 0}, wy0[2] = {0, 0};
            if (tma_h) {
                // terrain windows of the pair: one 2-D box per env around the base cell, issued before the noise draws so that the
                // copy overlaps them; rows / columns outside the field are zero-filled by the TMA unit and never addressed (the cell
                // indices are clipped to the field, legged_robot.py:905-907)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (k < ne) {
                        const float* R = s.root + (e0 + k) * 13;
                        wx0[k] = __float2int_rd(mul_rn(add_rn(R[0], p.border_size), inv_hs)) - HW_HALF;
                        wy0[k] = (__float2int_rd(mul_rn(add_rn(R[1], p.border_size), inv_hs)) - HW_HALF) & ~7;
                    }
                }
                fence_proxy_async();   // the previous pair's generic-proxy reads of the windows precede the async-proxy writes
                __syncwarp();
                if (lane == 0) mbar_expect_tx(s.hbar + warp, static_cast<uint32_t>(ne) * HW_ELEMS * 2);
                __syncwarp();
                if (lane < ne) {
                    const int cy = lane == 0 ? wy0[0] : wy0[1], cx = lane == 0 ? wx0[0] : wx0[1];
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                                     smem_u32(win + lane * HW_ELEMS)),
                                 "l"(reinterpret_cast<uint64_t>(&hmap)), "r"(cy), "r"(cx), "r"(smem_u32(s.hbar + warp))
                                 : "memory");
                }
            }
            if (p.add_noise) {
                // the Q (= 47) Philox blocks of an env hold the noise of its 4Q points: lane L draws block L of both envs, then
                // the tails (blocks 32..Q-1) of the two envs share one more round (lanes 0-15 / 16-31); uniforms go through a
                // per-warp scratch so that a lane can pick up the draws of ITS points
                for (int k = 0; k < ne; ++k) {
                    const philox::Stream rng(p.seed_lo, p.seed_hi, static_cast<uint64_t>(env_off + tile0 + e0 + k), step);
                    if (lane < Q) *reinterpret_cast<float4*>(un + k * HPAD + 4 * lane) = philox::u01(rng.words(philox::OBS_NOISE, OL::HB0 + lane));
                }
                const int k = lane >> 4, blk = 32 + (lane & 15);
                if (k < ne && blk < Q) {
                    const philox::Stream rng(p.seed_lo, p.seed_hi, static_cast<uint64_t>(env_off + tile0 + e0 + k), step);
                    *reinterpret_cast<float4*>(un + k * HPAD + 4 * blk) = philox::u01(rng.words(philox::OBS_NOISE, OL::HB0 + blk));
                }
                __syncwarp();
            }
            for (int k = 0; k < ne; ++k) {
                const int e = e0 + k;
                const float* R = s.root + e * 13;
                const float2 yw = s.yaw[e];
                const float rz = R[2], z05 = sub_rn(rz, 0.5f);
                int raw[RND];
#pragma unroll
                for (int r = 0; r < RND; ++r) raw[r] = 0;
                if (!p.mesh_plane) {
                    int pk[RND];
#pragma unroll
                    for (int r = 0; r < RND; r += 2) {   // two points (rounds r, r+1) per packed instruction
                        const float2 pa = s.pts[32 * r + lane], pb = s.pts[32 * (r + 1) + lane];
                        height_cells2<tma_h>(make_float2(pa.x, pb.x), make_float2(pa.y, pb.y), yw.x, yw.y, R[0], R[1], hc, rows, cols, pk[r], pk[r + 1]);
                    }
                    if (tma_h) {
                        if (k == 0) {
                            mbar_wait(s.hbar + warp, hphase);
                            hphase ^= 1u;
                        }
                        const int16_t* w = win + k * HW_ELEMS;
                        const int ox = wx0[k], oy = wy0[k];
#pragma unroll
                        for (int r = 0; r < RND; ++r) {   // samples from the staged window; a cell outside it (robot off the field: its
                            const int lx = (pk[r] >> 16) - ox, ly = (pk[r] & 0xffff) - oy;   // clipped cells are far from the base) reads the field itself
                            if (static_cast<unsigned>(lx) <= static_cast<unsigned>(HW_ROWS - 2) && static_cast<unsigned>(ly) <= static_cast<unsigned>(HW_COLS - 2)) {
                                const int16_t* h = w + lx * HW_COLS + ly;
                                raw[r] = min(min(static_cast<int>(h[0]), static_cast<int>(h[HW_COLS])), static_cast<int>(h[1]));
                            } else {
                                const int16_t* h = b.height_samples + (pk[r] >> 16) * cols + (pk[r] & 0xffff);
                                raw[r] = min(min(static_cast<int>(__ldg(h)), static_cast<int>(__ldg(h + cols))), static_cast<int>(__ldg(h + 1)));
                            }
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < RND; ++r) {   // 18 gathers issued together; padding points read a valid (unused) cell
                            const int16_t* h = b.height_samples + pk[r];
                            raw[r] = min(min(static_cast<int>(__ldg(h)), static_cast<int>(__ldg(h + cols))), static_cast<int>(__ldg(h + 1)));
                        }
                    }
                }
                float* mh_out = b.measured_heights + static_cast<size_t>(tile0 + e) * H + lane;
                float* ob_out = b.obs_buf + static_cast<size_t>(tile0 + e) * O + OW + lane;
                const float* ue = un + k * HPAD + lane + OL::HSH;
                float part = 0.0f;
#pragma unroll
                for (int r = 0; r < RND; ++r) {
                    if (32 * r + lane < H) {
                        const float mh = mul_rn(static_cast<float>(raw[r]), p.vertical_scale);
                        mh_out[32 * r] = mh;
                        part += rz - mh;
                        float v = clampf(sub_rn(z05, mh), -1.0f, 1.0f) * p.obs_height;
                        if (p.add_noise) v = add_noise(v, ue[32 * r], p.noise_height);
                        ob_out[32 * r] = clampf(v, -p.clip_obs, p.clip_obs);
                    }
                }
                if (s.hsum) {
#pragma unroll
                    for (int m = 16; m > 0; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
                    if (lane == 0) s.hsum[e] = part;
                }
            }
            __syncwarp();   // the scratch is rewritten by the next env pair
        }
    }

    // ---- phase W: 4 lanes per env — per-DOF / per-foot / per-body work, DOF observations, all noise draws
    {
        const int e = tid >> 2, g = tid & 3;
        const philox::Stream rng(p.seed_lo, p.seed_hi, static_cast<uint64_t>(env_off + tile0 + e), step);
        float pa_rate = 0.f, pd_acc = 0.f, pd_vel = 0.f, ptq = 0.f, ppos_lim = 0.f, pvel_lim = 0.f, ptq_lim = 0.f, pstand = 0.f;
        float o_pos[3], o_vel[3], o_act[3];
        const float inv_dt = 1.0f / p.dt;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int d = 3 * g + j;
            const float a = s.act[e * ND + d], la = s.lact[e * ND + d], tq = s.tq[e * ND + d], ldv = s.ldv[e * ND + d];
            const float2 qv = *reinterpret_cast<const float2*>(s.dof + e * 24 + 2 * d);
            const float q = qv.x, qd = qv.y, q0 = p.default_dof_pos[d];
            const float da = la - a, acc = (ldv - qd) * inv_dt;
            pa_rate = fmaf(da, da, pa_rate);
            pd_acc = fmaf(acc, acc, pd_acc);
            pd_vel = fmaf(qd, qd, pd_vel);
            ptq = fmaf(tq, tq, ptq);
            if (rs[T_DOF_POS_LIMITS] != 0.f) ppos_lim += fmaxf(q - p.dof_pos_hi[d], 0.0f) - fminf(q - p.dof_pos_lo[d], 0.0f);
            if (rs[T_DOF_VEL_LIMITS] != 0.f) pvel_lim += clampf(fabsf(qd) - p.dof_vel_limits[d] * p.soft_dof_vel_limit, 0.0f, 1.0f);
            if (rs[T_TORQUE_LIMITS] != 0.f) ptq_lim += fmaxf(fabsf(tq) - p.torque_limits[d] * p.soft_torque_limit, 0.0f);
            pstand += fabsf(q - q0);
            o_pos[j] = (q - q0) * p.obs_dof_pos;
            o_vel[j] = qd * p.obs_dof_vel;
            o_act[j] = a;
            s.ldv[e * ND + d] = qd;   // R12: last_dof_vel <- dof_vel (a reset env is fixed up in phase S)
        }
        // feet: lane g owns foot g (legged_robot.py:988-1015)
        const float* Ff = s.contact + (e * B + p.feet_idx[g]) * 3;
        const float fz = Ff[2];
        float p_air = 0.f;
        if (rs[T_FEET_AIR_TIME] != 0.0f) {   // the reference only mutates this state when the term is active
            float fat = s.fat[e * 4 + g];
            const bool contact = fz > 1.0f;
            const bool filt = contact || (s.lc[e * 4 + g] != 0);
            const bool first = (fat > 0.0f) && filt;
            fat = add_rn(fat, p.dt);
            p_air = first ? fat - 0.5f : 0.0f;
            s.fat[e * 4 + g] = filt ? 0.0f : fat;
            s.lc[e * 4 + g] = contact ? 1 : 0;
        }
        float p_stumble = 0.f, p_fcf = 0.f, p_coll = 0.f;
        if (rs[T_STUMBLE] != 0.f) p_stumble = (norm2_rn(Ff[0], Ff[1]) > mul_rn(5.0f, fabsf(fz))) ? 1.0f : 0.0f;
        if (rs[T_FEET_CONTACT_FORCES] != 0.f) p_fcf = fmaxf(norm3_rn(Ff[0], Ff[1], fz) - p.max_contact_force, 0.0f);
        if (rs[T_COLLISION] != 0.f) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float* Fp = s.contact + (e * B + p.pen_idx[g + 4 * j]) * 3;
#if PP_SQNORM
                p_coll += (sq3_rn(Fp[0], Fp[1], Fp[2]) > thr.gt01) ? 1.0f : 0.0f;
#else
                p_coll += (norm3_rn(Fp[0], Fp[1], Fp[2]) > 0.1f) ? 1.0f : 0.0f;
#endif
            }
        }
        // R8 (contact part): lane t checks termination body t
        float p_term = 0.f;
        if (g < p.num_term) {
            const float* F = s.contact + (e * B + p.term_idx[g]) * 3;
#if PP_SQNORM
            p_term = sq3_rn(F[0], F[1], F[2]) > thr.gt1 ? 1.0f : 0.0f;
#else
            p_term = norm3_rn(F[0], F[1], F[2]) > 1.0f ? 1.0f : 0.0f;
#endif
        }
#if PP_ALIAS_OBS
        __syncthreads();   // every lane of the CTA is done with the contact tile: it becomes the obs tile
#endif
        // uniforms for the noisy observation columns (0..35 = Philox blocks 0..8; trajectory layout: 0..52 = blocks 0..13),
        // exchanged through the obs tile
        if (p.add_noise) {
            if (!TRAJ) {
                float4* stage = reinterpret_cast<float4*>(s.obs + e * OW);
                stage[g] = philox::u01(rng.words(philox::OBS_NOISE, g));
                stage[g + 4] = philox::u01(rng.words(philox::OBS_NOISE, g + 4));
                if (g == 0) stage[8] = philox::u01(rng.words(philox::OBS_NOISE, 8));
            } else {   // rows of 65 floats are not 16-byte aligned: scalar stores
#pragma unroll
                for (int r = 0; r < (OL::NB + 3) / 4; ++r) {
                    const int blk = g + 4 * r;
                    if (blk < OL::NB) {
                        const float4 u = philox::u01(rng.words(philox::OBS_NOISE, blk));
                        float* st = s.obs + e * OW + 4 * blk;
                        st[0] = u.x, st[1] = u.y, st[2] = u.z, st[3] = u.w;
                    }
                }
            }
        }
        // DOF observation columns 12..47 (+ noise), already clipped (legged_robot.py:100-101,208-226)
        if (p.add_noise) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float u1 = s.obs[e * OW + OL::DOF + 3 * g + j], u2 = s.obs[e * OW + OL::DOF + ND + 3 * g + j];
                o_pos[j] = add_noise(o_pos[j], u1, p.noise_dof_pos);
                o_vel[j] = add_noise(o_vel[j], u2, p.noise_dof_vel);
            }
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            s.obs[e * OW + OL::DOF + 3 * g + j] = clampf(o_pos[j], -p.clip_obs, p.clip_obs);
            s.obs[e * OW + OL::DOF + ND + 3 * g + j] = clampf(o_vel[j], -p.clip_obs, p.clip_obs);
            s.obs[e * OW + OL::DOF + 2 * ND + 3 * g + j] = clampf(o_act[j], -p.clip_obs, p.clip_obs);
        }
        // quad reductions -> per-env partial sums
        pa_rate = quad_sum(pa_rate), pd_acc = quad_sum(pd_acc), pd_vel = quad_sum(pd_vel), ptq = quad_sum(ptq);
        pstand = quad_sum(pstand), p_air = quad_sum(p_air), p_coll = quad_sum(p_coll), p_term = quad_sum(p_term);
        if (rs[T_DOF_POS_LIMITS] != 0.f) ppos_lim = quad_sum(ppos_lim);
        if (rs[T_DOF_VEL_LIMITS] != 0.f) pvel_lim = quad_sum(pvel_lim);
        if (rs[T_TORQUE_LIMITS] != 0.f) ptq_lim = quad_sum(ptq_lim);
        if (rs[T_STUMBLE] != 0.f) p_stumble = quad_sum(p_stumble);
        if (rs[T_FEET_CONTACT_FORCES] != 0.f) p_fcf = quad_sum(p_fcf);
        if (g == 0) {
            PART(P_ACTION_RATE) = pa_rate, PART(P_DOF_ACC) = pd_acc, PART(P_DOF_VEL) = pd_vel;
            PART(P_TORQUES) = ptq;
        } else if (g == 1) {
            PART(P_POS_LIM) = ppos_lim, PART(P_VEL_LIM) = pvel_lim, PART(P_TQ_LIM) = ptq_lim;
        } else if (g == 2) {
            PART(P_STAND) = pstand, PART(P_AIR) = p_air, PART(P_STUMBLE) = p_stumble;
        } else {
            PART(P_FCF) = p_fcf, PART(P_COLL) = p_coll, PART(P_TERM) = p_term;
        }
    }
    __syncthreads();

    // ---- phase S: one thread per env — everything that is scalar per env ---------------------------
    if (tid < TILE) {
        const int e = tid;
        const bool valid = e < nvalid;
        const size_t ge = static_cast<size_t>(tile0 + e);
        const philox::Stream rng(p.seed_lo, p.seed_hi, static_cast<uint64_t>(env_off + tile0 + e), step);
        const float* R = s.root + e * 13;
        const float qx = R[3], qy = R[4], qz = R[5], qw = R[6];

        // R4: counters + body-frame vectors (legged_robot.py:114-121)
        const long long ep = s.ep[e] + 1;
        float blx, bly, blz, bax, bay, baz, pgx, pgy, pgz;
        quat_rotate_inverse(qx, qy, qz, qw, R[7], R[8], R[9], blx, bly, blz);
        quat_rotate_inverse(qx, qy, qz, qw, R[10], R[11], R[12], bax, bay, baz);
        quat_rotate_inverse(qx, qy, qz, qw, 0.0f, 0.0f, -1.0f, pgx, pgy, pgz);

        // R5: command resampling + heading (legged_robot.py:343-354)
        const float4 cmd4 = *reinterpret_cast<const float4*>(s.cmd + e * 4);
        float c0 = cmd4.x, c1 = cmd4.y, c2 = cmd4.z, c3 = cmd4.w;
        // (LeggedRobotTrajectory has no command resampling: legged_robot_trajectory.py:405-417)
        if (!TRAJ && static_cast<int>(ep) % p.resample_steps == 0) resample_commands(p, thr, rng, philox::CMD_PERIODIC, c0, c1, c2, c3);
        if (!TRAJ && p.heading_command) {
            // forward = quat_apply(q, [1,0,0]) (x,y only): t = 2*cross(q_xyz, [1,0,0]) = (0, 2qz, -2qy)
            const float tyy = 2.0f * qz, tzz = -2.0f * qy;
            const float fx = 1.0f + (qy * tzz - qz * tyy);
            const float fy = qw * tyy + (-qx * tzz);
            c2 = clampf(0.5f * wrap_to_pi(sub_rn(c3, atan2f(fy, fx))), -1.0f, 1.0f);
        }

        // R7: pushes (legged_robot.py:456-461)
        float lrv[6] = {R[7], R[8], R[9], R[10], R[11], R[12]};
        bool pushed = do_push != 0;
        float tpush = 0.0f;
        if (TRAJ) {   // per-env push timers (legged_robot_trajectory.py:169-178)
            tpush = sub_rn(s.tpush[e], p.dt);
            pushed = tpush <= 0.0f;
        }
        if (pushed) {
            const uint4 w = rng.words(philox::PUSH, 0);
            lrv[0] = affine_rn(p.push_span, philox::u01(w.x), p.push_lo);
            lrv[1] = affine_rn(p.push_span, philox::u01(w.y), p.push_lo);
            if (TRAJ) tpush = affine_rn(p.push_t_span, philox::u01(rng.words(philox::PUSH_TIMER, 0).x), p.push_t_lo);
        }
        if (TRAJ) s.tpush[e] = tpush;

        // R8: termination (legged_robot.py:139-145)
        const bool time_out = static_cast<float>(ep) > p.max_episode_length;
        const bool reset = (PART(P_TERM) > 0.0f) | time_out;

        // R9: reward assembly in the reference's (alphabetical) order (legged_robot.py:189-206)
#if PP_SQNORM
        const float cmd_sq = sq2_rn(c0, c1);
        const bool cmd_gt01 = cmd_sq > thr.gt01, cmd_lt01 = cmd_sq < thr.lt01;
#else
        const float cmd_norm = norm2_rn(c0, c1);
        const bool cmd_gt01 = cmd_norm > 0.1f, cmd_lt01 = cmd_norm < 0.1f;
#endif
        float rew = 0.0f;
        float* sums = s.sums + e;
        auto add_term = [&](int k, float val) {
            const float r = val * rs[k];
            rew += r;
            sums[p.sum_row[k] * TILE] += r;
        };
        if (rs[T_ACTION_RATE] != 0.f) add_term(T_ACTION_RATE, PART(P_ACTION_RATE));
        if (rs[T_ANG_VEL_XY] != 0.f) add_term(T_ANG_VEL_XY, bax * bax + bay * bay);
        if (rs[T_BASE_HEIGHT] != 0.f) {
            float bh = R[2];
            if (ROUGH)   // mean(root_z - measured_heights) (legged_robot.py:938-939): phase H's warp-reduced sum (fixed order)
                bh = s.hsum[e] / static_cast<float>(p.num_heights);
            const float dh = bh - p.base_height_target;
            add_term(T_BASE_HEIGHT, dh * dh);
        }
        if (rs[T_COLLISION] != 0.f) add_term(T_COLLISION, PART(P_COLL));
        float te0 = 0.0f, te1 = 0.0f;   // square(proj_z(root) - trajectory[:, 0]) (legged_robot_trajectory.py:1064-1066, :1101-1103)
        if (TRAJ) {
            const float d0 = R[0] - s.traj[e * TRAJ_W], d1 = R[1] - s.traj[e * TRAJ_W + 1];
            te0 = d0 * d0, te1 = d1 * d1;
            if (rs[T_DIFFERENTIAL_ERROR] != 0.f) {   // :1100-1110
                const float pe0 = s.perr[e * 2], pe1 = s.perr[e * 2 + 1];
                const float de = sqrtf(te0 * te0 + te1 * te1) - sqrtf(pe0 * pe0 + pe1 * pe1);
                add_term(T_DIFFERENTIAL_ERROR, (de < 0.0f ? p.diff_neg_slope : p.diff_pos_slope) * de);
            }
        }
        if (rs[T_DOF_ACC] != 0.f) add_term(T_DOF_ACC, PART(P_DOF_ACC));
        if (rs[T_DOF_POS_LIMITS] != 0.f) add_term(T_DOF_POS_LIMITS, PART(P_POS_LIM));
        if (rs[T_DOF_VEL] != 0.f) add_term(T_DOF_VEL, PART(P_DOF_VEL));
        if (rs[T_DOF_VEL_LIMITS] != 0.f) add_term(T_DOF_VEL_LIMITS, PART(P_VEL_LIM));
        if (rs[T_FEET_AIR_TIME] != 0.f) add_term(T_FEET_AIR_TIME, TRAJ ? PART(P_AIR) : PART(P_AIR) * (cmd_gt01 ? 1.0f : 0.0f));
        if (rs[T_FEET_CONTACT_FORCES] != 0.f) add_term(T_FEET_CONTACT_FORCES, PART(P_FCF));
        if (rs[T_LIN_VEL_Z] != 0.f) add_term(T_LIN_VEL_Z, blz * blz);
        if (rs[T_ORIENTATION] != 0.f) add_term(T_ORIENTATION, pgx * pgx + pgy * pgy);
        if (rs[T_STAND_STILL] != 0.f) add_term(T_STAND_STILL, PART(P_STAND) * (cmd_lt01 ? 1.0f : 0.0f));
        if (rs[T_STUMBLE] != 0.f) add_term(T_STUMBLE, PART(P_STUMBLE) > 0.0f ? 1.0f : 0.0f);
        if (rs[T_TORQUE_LIMITS] != 0.f) add_term(T_TORQUE_LIMITS, PART(P_TQ_LIM));
        if (rs[T_TORQUES] != 0.f) add_term(T_TORQUES, PART(P_TORQUES));
        const float inv_sigma = 1.0f / p.tracking_sigma;
        if (rs[T_TRACKING_ANG_VEL] != 0.f) {
            const float er = c2 - baz;
            add_term(T_TRACKING_ANG_VEL, expf(-(er * er) * inv_sigma));
        }
        if (rs[T_TRACKING_LIN_VEL] != 0.f) {
            const float ex = c0 - blx, ey = c1 - bly;
            add_term(T_TRACKING_LIN_VEL, expf(-(ex * ex + ey * ey) * inv_sigma));
        }
        if (TRAJ && rs[T_TRACKING_ROM] != 0.f)   // exp(-inner(err^2, weighting) / sigma), legged_robot_trajectory.py:1060-1069
            add_term(T_TRACKING_ROM, expf(-(te0 * p.traj_weight[0] + te1 * p.traj_weight[1]) / p.tracking_sigma));
        if (p.only_positive) rew = fmaxf(rew, 0.0f);
        if (rs[T_TERMINATION] != 0.f) add_term(T_TERMINATION, (reset && !time_out) ? 1.0f : 0.0f);

        // R10: in-place reset (legged_robot.py:147-187, anymal.py:56-60) — a branch, not a host compaction (H7)
        float zpost = R[2], xpost = R[0], ypost = R[1];
        long long ep_out = ep, level = 0;
        if (p.terrain_curriculum && valid) level = b.terrain_levels[ge];
        if (reset && valid) {
            float ox = b.env_origins[ge * 3 + 0], oy = b.env_origins[ge * 3 + 1], oz = b.env_origins[ge * 3 + 2];
            if (p.terrain_curriculum) {   // legged_robot.py:463-486
                const float dist = norm2_rn(sub_rn(R[0], ox), sub_rn(R[1], oy));
                const bool up = dist > p.half_env_length;
                const bool down = (dist < mul_rn(mul_rn(norm2_rn(c0, c1), p.max_episode_length_s), 0.5f)) && !up;
                level += (up ? 1 : 0) - (down ? 1 : 0);
                if (level >= p.max_terrain_level)
                    level = philox::bounded(rng.words(philox::TERRAIN, 0).x, static_cast<uint32_t>(p.max_terrain_level));
                else if (level < 0)
                    level = 0;
                const float* og = b.terrain_origins + (level * p.terrain_num_cols + b.terrain_types[ge]) * 3;
                ox = og[0], oy = og[1], oz = og[2];
                b.terrain_levels[ge] = level;
                b.env_origins[ge * 3 + 0] = ox, b.env_origins[ge * 3 + 1] = oy, b.env_origins[ge * 3 + 2] = oz;
            }
            // dofs: q = q0 * U(0.5, 1.5), qd = 0 (legged_robot.py:423-425); fix up what phase W assumed
            if (!TRAJ) {
                for (int blk = 0; blk < 3; ++blk) {
                    const uint4 w = rng.words(philox::RESET_DOF, blk);
                    const uint4 n1 = rng.words(philox::OBS_NOISE, 3 + blk), n2 = rng.words(philox::OBS_NOISE, 6 + blk);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int d = 4 * blk + i;
                        const float q0 = p.default_dof_pos[d];
                        const float q = mul_rn(q0, affine_rn(1.0f, philox::u01(philox::word(w, i)), 0.5f));
                        reinterpret_cast<float2*>(b.dof_state)[ge * ND + d] = make_float2(q, 0.0f);
                        float op = (q - q0) * p.obs_dof_pos, ov = 0.0f;
                        if (p.add_noise) {
                            op = add_noise(op, philox::u01(philox::word(n1, i)), p.noise_dof_pos);
                            ov = add_noise(ov, philox::u01(philox::word(n2, i)), p.noise_dof_vel);
                        }
                        s.obs[e * OW + OL::DOF + d] = clampf(op, -p.clip_obs, p.clip_obs);
                        s.obs[e * OW + OL::DOF + ND + d] = clampf(ov, -p.clip_obs, p.clip_obs);
                        s.ldv[e * ND + d] = 0.0f;
                    }
                }
            } else {   // dof columns 29.. are not Philox-block aligned: (column >> 2, column & 3) per element (rare path)
                for (int blk = 0; blk < 3; ++blk) {
                    const uint4 w = rng.words(philox::RESET_DOF, blk);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int d = 4 * blk + i;
                        const float q0 = p.default_dof_pos[d];
                        const float q = mul_rn(q0, affine_rn(1.0f, philox::u01(philox::word(w, i)), 0.5f));
                        reinterpret_cast<float2*>(b.dof_state)[ge * ND + d] = make_float2(q, 0.0f);
                        float op = (q - q0) * p.obs_dof_pos, ov = 0.0f;
                        if (p.add_noise) {
                            const int c1 = OL::DOF + d, c2 = OL::DOF + ND + d;
                            const uint4 n1 = rng.words(philox::OBS_NOISE, c1 >> 2), n2 = rng.words(philox::OBS_NOISE, c2 >> 2);
                            const int k1 = c1 & 3, k2 = c2 & 3;
                            op = add_noise(op, philox::u01(k1 == 0 ? n1.x : k1 == 1 ? n1.y : k1 == 2 ? n1.z : n1.w), p.noise_dof_pos);
                            ov = add_noise(ov, philox::u01(k2 == 0 ? n2.x : k2 == 1 ? n2.y : k2 == 2 ? n2.z : n2.w), p.noise_dof_vel);
                        }
                        s.obs[e * OW + OL::DOF + d] = clampf(op, -p.clip_obs, p.clip_obs);
                        s.obs[e * OW + OL::DOF + ND + d] = clampf(ov, -p.clip_obs, p.clip_obs);
                        s.ldv[e * ND + d] = 0.0f;
                    }
                }
            }
            // root (legged_robot.py:441-449)
            float nr[13];
#pragma unroll
            for (int k = 0; k < 13; ++k) nr[k] = p.base_init_state[k];
            nr[0] = add_rn(nr[0], ox), nr[1] = add_rn(nr[1], oy), nr[2] = add_rn(nr[2], oz);
            if (p.custom_origins) {
                const uint4 w = rng.words(philox::RESET_XY, 0);
                nr[0] = add_rn(nr[0], affine_rn(2.0f, philox::u01(w.x), -1.0f));
                nr[1] = add_rn(nr[1], affine_rn(2.0f, philox::u01(w.y), -1.0f));
            }
            const uint4 w0 = rng.words(philox::RESET_VEL, 0), w1 = rng.words(philox::RESET_VEL, 1);
            nr[7] = affine_rn(1.0f, philox::u01(w0.x), -0.5f), nr[8] = affine_rn(1.0f, philox::u01(w0.y), -0.5f);
            nr[9] = affine_rn(1.0f, philox::u01(w0.z), -0.5f), nr[10] = affine_rn(1.0f, philox::u01(w0.w), -0.5f);
            nr[11] = affine_rn(1.0f, philox::u01(w1.x), -0.5f), nr[12] = affine_rn(1.0f, philox::u01(w1.y), -0.5f);
            zpost = nr[2], xpost = nr[0], ypost = nr[1];
#pragma unroll
            for (int k = 0; k < 6; ++k) lrv[k] = nr[7 + k];
#pragma unroll
            for (int k = 0; k < 13; ++k) b.root_states[ge * 13 + k] = nr[k];
            if (!TRAJ) resample_commands(p, thr, rng, philox::CMD_RESET, c0, c1, c2, c3);
            if (TRAJ) {   // prev_error from the (stale) trajectory and the NEW root position (legged_robot_trajectory.py:233)
                const float d0 = s.traj[e * TRAJ_W] - nr[0], d1 = s.traj[e * TRAJ_W + 1] - nr[1];
                s.perr[e * 2] = d0 * d0, s.perr[e * 2 + 1] = d1 * d1;
            }
            *reinterpret_cast<float4*>(s.fat + e * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            ep_out = 0;
            // extras["episode"] statistics (legged_robot.py:175-179): per-CTA partials in shared memory
            for (int k = 0; k < K; ++k) {
                atomicAdd(&s.acc[k], static_cast<double>(sums[k * TILE]));
                sums[k * TILE] = 0.0f;
            }
            atomicAdd(s.nreset, 1);
            if (p.zero_lstm_on_reset) {   // h,c: [2, N*12, 8]; this env owns 2 x 96 floats per array
                const size_t M8 = static_cast<size_t>(N) * ND * 8;
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int l = 0; l < 2; ++l) {
                    float4* hp = reinterpret_cast<float4*>(b.lstm_h + l * M8 + ge * ND * 8);
                    float4* cp = reinterpret_cast<float4*>(b.lstm_c + l * M8 + ge * ND * 8);
                    for (int k = 0; k < 24; ++k) hp[k] = z4, cp[k] = z4;
                }
            }
        } else if (pushed && valid) {
            b.root_states[ge * 13 + 7] = lrv[0];
            b.root_states[ge * 13 + 8] = lrv[1];
        }
        if (p.terrain_curriculum) {   // sum of terrain levels of the tile: warp shuffle first, one shared atomic per warp
            long long lv = valid ? level : 0;
            constexpr int SW = TILE < 32 ? TILE : 32;                              // phase S runs on the first TILE threads only
            constexpr unsigned SMASK = TILE < 32 ? ((1u << SW) - 1u) : 0xffffffffu;
#pragma unroll
            for (int o = SW / 2; o > 0; o >>= 1) lv += __shfl_xor_sync(SMASK, lv, o);
            if ((tid & 31) == 0) atomicAdd(&s.acc[K], static_cast<double>(lv));
        }

        // R11: first 12 observation columns (+ noise, clip); uniforms were staged by phase W
        float o[12] = {blx * p.obs_lin_vel, bly * p.obs_lin_vel, blz * p.obs_lin_vel, bax * p.obs_ang_vel, bay * p.obs_ang_vel,
                       baz * p.obs_ang_vel, pgx, pgy, pgz, c0 * p.obs_lin_vel, c1 * p.obs_lin_vel, c2 * p.obs_ang_vel};
        float4* o4 = reinterpret_cast<float4*>(s.obs + e * OW);
        float* orow = s.obs + e * OW;
        if (p.add_noise) {
            float4 u0, u1, u2;
            if (!TRAJ) u0 = o4[0], u1 = o4[1], u2 = o4[2];
            else u0 = make_float4(orow[0], orow[1], orow[2], orow[3]), u1 = make_float4(orow[4], orow[5], orow[6], orow[7]), u2.x = orow[8];
            o[0] = add_noise(o[0], u0.x, p.noise_lin_vel), o[1] = add_noise(o[1], u0.y, p.noise_lin_vel);
            o[2] = add_noise(o[2], u0.z, p.noise_lin_vel), o[3] = add_noise(o[3], u0.w, p.noise_ang_vel);
            o[4] = add_noise(o[4], u1.x, p.noise_ang_vel), o[5] = add_noise(o[5], u1.y, p.noise_ang_vel);
            o[6] = add_noise(o[6], u1.z, p.noise_gravity), o[7] = add_noise(o[7], u1.w, p.noise_gravity);
            o[8] = add_noise(o[8], u2.x, p.noise_gravity);
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) o[k] = clampf(o[k], -p.clip_obs, p.clip_obs);
        if (!TRAJ) {
            o4[0] = make_float4(o[0], o[1], o[2], o[3]);
            o4[1] = make_float4(o[4], o[5], o[6], o[7]);
            o4[2] = make_float4(o[8], o[9], o[10], o[11]);
        } else {   // (trajectory - proj_z(root)) * trajectory_scale, no noise (legged_robot_trajectory.py:277-283, :573)
#pragma unroll
            for (int k = 0; k < 9; ++k) orow[k] = o[k];
#pragma unroll
            for (int k = 0; k < TRAJ_W; k += 2) {
                orow[9 + k] = clampf((s.traj[e * TRAJ_W + k] - xpost) * p.traj_scale[0], -p.clip_obs, p.clip_obs);
                orow[10 + k] = clampf((s.traj[e * TRAJ_W + k + 1] - ypost) * p.traj_scale[1], -p.clip_obs, p.clip_obs);
            }
        }

        // R12 + per-env outputs
        s.blv[e * 3 + 0] = blx, s.blv[e * 3 + 1] = bly, s.blv[e * 3 + 2] = blz;
        s.bav[e * 3 + 0] = bax, s.bav[e * 3 + 1] = bay, s.bav[e * 3 + 2] = baz;
        s.pg[e * 3 + 0] = pgx, s.pg[e * 3 + 1] = pgy, s.pg[e * 3 + 2] = pgz;
#pragma unroll
        for (int k = 0; k < 6; ++k) s.lrv[e * 6 + k] = lrv[k];
        *reinterpret_cast<float4*>(s.cmd + e * 4) = make_float4(c0, c1, c2, c3);
        s.ep[e] = ep_out;
        s.rew[e] = rew;
        s.reset[e] = reset ? 1 : 0;
        s.tout[e] = time_out ? 1 : 0;
        if (ROUGH) s.zpost[e] = zpost;
    }
    fence_proxy_async();   // make the generic-proxy smem writes visible to the bulk-store engine
    __syncthreads();

    // ---- write the tile back ----------------------------------------------------------------------
    const bool obs_bulk = full && (O == OW);
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
            if (obs_bulk) bulk_s2g(b.obs_buf + static_cast<size_t>(tile0) * OW, s.obs, TILE * OW * 4);
            if (TRAJ) {
                bulk_s2g(b.prev_error + static_cast<size_t>(tile0) * 2, s.perr, TILE * 2 * 4);
                bulk_s2g(b.time_until_next_push + tile0, s.tpush, TILE * 4);
            }
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
        if (TRAJ) {
            coop_copy(b.prev_error + static_cast<size_t>(tile0) * 2, s.perr, nvalid * 2);
            coop_copy(b.time_until_next_push + tile0, s.tpush, nvalid);
        }
    }
    if (!obs_bulk) {
        for (int i = tid; i < nvalid * OW; i += TILE * LPE) b.obs_buf[static_cast<size_t>(tile0 + i / OW) * O + (i % OW)] = s.obs[i];
    }

    // ---- phase H': height observations of the envs that reset this step, redone with the post-reset base height ----
    if (ROUGH) {
        __syncthreads();   // measured_heights written in phase H by other threads of this CTA
        if (*s.nreset > 0) {
            const int H = p.num_heights, Q = (H + 3) >> 2;
            const int de = (TILE * LPE) / Q, dq = (TILE * LPE) - de * Q;
            int e = tid / Q, q = tid - e * Q;
            for (int i = tid; i < nvalid * Q; i += TILE * LPE) {
                if (s.reset[e]) {
                    const float z05 = sub_rn(s.zpost[e], 0.5f);
                    const int pt0 = 4 * q;
                    float un[4] = {0.5f, 0.5f, 0.5f, 0.5f};
                    if (p.add_noise) {   // height column j sits in Philox block HB0 + (j + HSH) / 4, word (j + HSH) % 4
                        const philox::Stream rng(p.seed_lo, p.seed_hi, static_cast<uint64_t>(env_off + tile0 + e), step);
                        const float4 ua = philox::u01(rng.words(philox::OBS_NOISE, OL::HB0 + q));
                        const float4 ub = OL::HSH ? philox::u01(rng.words(philox::OBS_NOISE, OL::HB0 + q + 1)) : ua;
                        const float u8[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) un[j] = u8[j + OL::HSH];
                    }
                    const float* mh_in = b.measured_heights + static_cast<size_t>(tile0 + e) * H + pt0;
                    float* ob_out = b.obs_buf + static_cast<size_t>(tile0 + e) * O + OW + pt0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (pt0 + j < H) {
                            float v = clampf(sub_rn(z05, mh_in[j]), -1.0f, 1.0f) * p.obs_height;
                            if (p.add_noise) v = add_noise(v, un[j], p.noise_height);
                            ob_out[j] = clampf(v, -p.clip_obs, p.clip_obs);
                        }
                    }
                }
                e += de, q += dq;
                if (q >= Q) q -= Q, ++e;
            }
        }
    }

    // ---- extras["episode"]: cross-CTA reduction, finalised by the last CTA to arrive ---------------
    __syncthreads();
    const int nreset = *s.nreset;
    if (nreset > 0 && tid < K) atomicAdd(&b.ws_sums[tid], s.acc[tid]);
    if (p.terrain_curriculum && tid == K) atomicAdd(&b.ws_sums[K], s.acc[K]);
    if (nreset > 0 && tid == K + 1) atomicAdd(&b.ws_sums[K + 1], static_cast<double>(nreset));
#if !PP_FINALIZE_KERNEL
    __threadfence();
    __syncthreads();
    __shared__ unsigned int s_ticket;
    if (tid == 0) s_ticket = atomicAdd(b.ws_counter, 1u);
    __syncthreads();
    if (s_ticket == gridDim.x - 1) {
        __threadfence();
        volatile double* ws = b.ws_sums;
        const double cnt = ws[K + 1];
        if (tid < K && cnt > 0.0) b.extras_out[tid] = static_cast<float>(ws[tid] / cnt) / p.max_episode_length_s;
        if (tid == K && cnt > 0.0) b.extras_out[K] = static_cast<float>(ws[K] / static_cast<double>(N));
        if (tid == K + 1) b.extras_out[K + 1] = static_cast<float>(cnt);
        __syncthreads();
        if (tid < K + 2) b.ws_sums[tid] = 0.0;
        if (tid == 0) *b.ws_counter = 0u;
    }
#endif
    if (full && tid == 0) bulk_wait_read0();   // shared memory must stay alive until the bulk stores have read it
}

#if PP_FINALIZE_KERNEL
// legged_robot.py:175-182: mean over the reset envs / max_episode_length_s; untouched when nothing reset (:156-157)
__global__ void extras_finalize_kernel(const __grid_constant__ B200LeggedParams p, const __grid_constant__ B200LeggedBuffers b) {
    const int K = p.num_sum_rows, tid = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    const double cnt = b.ws_sums[K + 1];
    const double mine = tid < K + 2 ? b.ws_sums[tid] : 0.0;
    __syncthreads();
    if (tid < K && cnt > 0.0) b.extras_out[tid] = static_cast<float>(mine / cnt) / p.max_episode_length_s;
    if (tid == K && cnt > 0.0) b.extras_out[K] = static_cast<float>(mine / static_cast<double>(p.num_envs));
    if (tid == K + 1) b.extras_out[K + 1] = static_cast<float>(cnt);
    if (tid == K + 2 && cnt > 0.0) b.extras_out[K + 2] = p.max_command_x;   // legged_robot.py:183-184
    if (b.extras_raw && tid < K + 2) b.extras_raw[tid] = mine;
    if (tid < K + 2) b.ws_sums[tid] = 0.0;
    if (tid == 0 && b.step_counter) *b.step_counter += 1;
}
#endif


// ------------------------------------------------------------------------------------------------------------------
// reset_idx_kernel — LeggedRobot.reset_idx(env_ids) called from OUTSIDE step() (legged_robot.py:147-187, anymal.py:56-60;
// BaseTask.reset, base_task.py:111-119): the same draws and buffer updates as the in-step reset branch of phase S, for the
// envs flagged in `mask`, with no termination / time-out flag touched and no observation recomputed (the reference leaves
// obs_buf alone until the next step).  One thread per env of the shard; every env contributes its terrain level to the
// extras["episode"]["terrain_level"] mean (:181).  The random draws are keyed by `event` (the host passes a number that no
// env step uses: the external-reset count in the high bits, common_step_counter below — oracle/port_legged.py reset_idx).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) reset_idx_kernel(const __grid_constant__ B200LeggedParams p, const __grid_constant__ B200LeggedBuffers b,
                                                        const uint8_t* __restrict__ mask, unsigned long long event, long long env_off,
                                                        const SqThr thr) {
    const int N = p.num_envs, K = p.num_sum_rows;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ double s_acc[B200GYM_NUM_REWARD_TERMS + 2];
    if (threadIdx.x < K + 2) s_acc[threadIdx.x] = 0.0;
    __syncthreads();
    const bool valid = e < N;
    const bool hit = valid && mask[e] != 0;
    long long level = (valid && p.terrain_curriculum) ? b.terrain_levels[e] : 0;
    if (hit) {
        const philox::Stream rng(p.seed_lo, p.seed_hi, static_cast<unsigned long long>(env_off + e), event);
        float c0 = b.commands[e * 4 + 0], c1 = b.commands[e * 4 + 1], c2 = b.commands[e * 4 + 2], c3 = b.commands[e * 4 + 3];
        float ox = b.env_origins[e * 3 + 0], oy = b.env_origins[e * 3 + 1], oz = b.env_origins[e * 3 + 2];
        if (p.terrain_curriculum) {   // legged_robot.py:463-486
            const float dist = norm2_rn(sub_rn(b.root_states[e * 13 + 0], ox), sub_rn(b.root_states[e * 13 + 1], oy));
            const bool up = dist > p.half_env_length;
            const bool down = (dist < mul_rn(mul_rn(norm2_rn(c0, c1), p.max_episode_length_s), 0.5f)) && !up;
            level += (up ? 1 : 0) - (down ? 1 : 0);
            if (level >= p.max_terrain_level)
                level = philox::bounded(rng.words(philox::TERRAIN, 0).x, static_cast<uint32_t>(p.max_terrain_level));
            else if (level < 0)
                level = 0;
            const float* og = b.terrain_origins + (level * p.terrain_num_cols + b.terrain_types[e]) * 3;
            ox = og[0], oy = og[1], oz = og[2];
            b.terrain_levels[e] = level;
            b.env_origins[e * 3 + 0] = ox, b.env_origins[e * 3 + 1] = oy, b.env_origins[e * 3 + 2] = oz;
        }
        for (int blk = 0; blk < 3; ++blk) {   // dofs: q = q0 * U(0.5, 1.5), qd = 0 (:423-425)
            const uint4 w = rng.words(philox::RESET_DOF, blk);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int d = 4 * blk + i;
                const float q = mul_rn(p.default_dof_pos[d], affine_rn(1.0f, philox::u01(philox::word(w, i)), 0.5f));
                reinterpret_cast<float2*>(b.dof_state)[e * ND + d] = make_float2(q, 0.0f);
                b.last_actions[e * ND + d] = 0.0f;
                b.last_dof_vel[e * ND + d] = 0.0f;
            }
        }
        float nr[13];   // root (:441-449)
#pragma unroll
        for (int k = 0; k < 13; ++k) nr[k] = p.base_init_state[k];
        nr[0] = add_rn(nr[0], ox), nr[1] = add_rn(nr[1], oy), nr[2] = add_rn(nr[2], oz);
        if (p.custom_origins) {
            const uint4 w = rng.words(philox::RESET_XY, 0);
            nr[0] = add_rn(nr[0], affine_rn(2.0f, philox::u01(w.x), -1.0f));
            nr[1] = add_rn(nr[1], affine_rn(2.0f, philox::u01(w.y), -1.0f));
        }
        const uint4 w0 = rng.words(philox::RESET_VEL, 0), w1 = rng.words(philox::RESET_VEL, 1);
        nr[7] = affine_rn(1.0f, philox::u01(w0.x), -0.5f), nr[8] = affine_rn(1.0f, philox::u01(w0.y), -0.5f);
        nr[9] = affine_rn(1.0f, philox::u01(w0.z), -0.5f), nr[10] = affine_rn(1.0f, philox::u01(w0.w), -0.5f);
        nr[11] = affine_rn(1.0f, philox::u01(w1.x), -0.5f), nr[12] = affine_rn(1.0f, philox::u01(w1.y), -0.5f);
#pragma unroll
        for (int k = 0; k < 13; ++k) b.root_states[e * 13 + k] = nr[k];
        if (!p.traj_mode) {
            resample_commands(p, thr, rng, philox::CMD_RESET, c0, c1, c2, c3);
            *reinterpret_cast<float4*>(b.commands + e * 4) = make_float4(c0, c1, c2, c3);
        } else {   // legged_robot_trajectory.py:233: prev_error from the (stale) trajectory clone and the NEW root; the caller resets the generator
            const float d0 = b.trajectory[static_cast<size_t>(e) * TRAJ_W] - nr[0], d1 = b.trajectory[static_cast<size_t>(e) * TRAJ_W + 1] - nr[1];
            b.prev_error[e * 2] = d0 * d0, b.prev_error[e * 2 + 1] = d1 * d1;
        }
        *reinterpret_cast<float4*>(b.feet_air_time + e * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<long long*>(b.episode_length_buf)[e] = 0;
        b.reset_buf[e] = 1;
        for (int k = 0; k < K; ++k) {   // extras["episode"] statistics (:175-179)
            atomicAdd(&s_acc[k], static_cast<double>(b.episode_sums[static_cast<size_t>(k) * N + e]));
            b.episode_sums[static_cast<size_t>(k) * N + e] = 0.0f;
        }
        atomicAdd(&s_acc[K + 1], 1.0);
        if (p.zero_lstm_on_reset) {
            const size_t M8 = static_cast<size_t>(N) * ND * 8;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int l = 0; l < 2; ++l) {
                float4* hp = reinterpret_cast<float4*>(b.lstm_h + l * M8 + static_cast<size_t>(e) * ND * 8);
                float4* cp = reinterpret_cast<float4*>(b.lstm_c + l * M8 + static_cast<size_t>(e) * ND * 8);
                for (int k = 0; k < 24; ++k) hp[k] = z4, cp[k] = z4;
            }
        }
    }
    if (p.terrain_curriculum) {
        long long lv = valid ? level : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lv += __shfl_xor_sync(0xffffffffu, lv, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[K], static_cast<double>(lv));
    }
    __syncthreads();
    if (threadIdx.x < K + 2 && s_acc[threadIdx.x] != 0.0) atomicAdd(&b.ws_sums[threadIdx.x], s_acc[threadIdx.x]);
}

// largest y with sqrtf(y) <= t  (so sqrtf(x) > t <=> x > y) and smallest y with sqrtf(y) >= t (sqrtf(x) < t <=> x < y)
static float sq_gt(float t) {
    float y = t * t;
    while (sqrtf(y) <= t) y = nextafterf(y, INFINITY);
    while (sqrtf(y) > t) y = nextafterf(y, -INFINITY);
    return y;
}
static float sq_lt(float t) {
    float y = t * t;
    while (sqrtf(y) >= t) y = nextafterf(y, -INFINITY);
    while (sqrtf(y) < t) y = nextafterf(y, INFINITY);
    return y;
}

// Tensor map of the int16 heightfield [rows, cols] (cuTensorMapEncodeTiled through the runtime's driver entry point: no link against
// libcuda), box = one env window; cached per (pointer, rows, cols).  Returns false when the field cannot be described (row pitch not a
// multiple of 16 bytes, misaligned base, no driver entry point): the kernel then gathers from the field directly.
static bool height_tensor_map(const int16_t* field, int rows, int cols, CUtensorMap* out) {
    struct Entry { const int16_t* f; int r, c; CUtensorMap m; bool ok; };
    static Entry cache[8];
    static int used = 0;
    for (int i = 0; i < used; ++i)
        if (cache[i].f == field && cache[i].r == rows && cache[i].c == cols) {
            *out = cache[i].m;
            return cache[i].ok;
        }
    Entry e{field, rows, cols, {}, false};
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if ((cols % 8) == 0 && b200_aligned16(field) && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn &&
        q == cudaDriverEntryPointSuccess) {
        const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
        const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(cols) * 2};
        const cuuint32_t box[2] = {HW_COLS, HW_ROWS};
        const cuuint32_t es[2] = {1, 1};
        e.ok = reinterpret_cast<EncodeFn>(fn)(&e.m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<int16_t*>(field), gdim, gstr, box, es,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    cache[used < 8 ? used++ : 7] = e;
    *out = e.m;
    return e.ok;
}

template <int TILE, bool ROUGH, bool TRAJ, bool HTMA>
int launch_post_physics_impl(const B200LeggedParams& p, const B200LeggedBuffers& b, uint64_t step, long long env_off, int do_push,
                             cudaStream_t stream, const CUtensorMap& hmap) {
    const size_t smem = carve_tile<TILE, ROUGH, TRAJ>(nullptr, p.num_bodies, p.num_sum_rows, p.reward_scale[T_BASE_HEIGHT] != 0.0f, HTMA).bytes;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(post_physics_kernel<TILE, ROUGH, TRAJ, HTMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        B200_REQUIRE(e == cudaSuccess, B200GYM_ECUDA, "post_physics: cannot reserve %zu B of shared memory: %s", smem,
                     cudaGetErrorString(e));
        configured = smem;
    }
    const int grid = (p.num_envs + TILE - 1) / TILE;
    static const SqThr thr = {sq_gt(1.0f), sq_gt(0.1f), sq_gt(0.2f), sq_lt(0.1f)};
    b200_launch_pdl(p.num_envs, post_physics_kernel<TILE, ROUGH, TRAJ, HTMA>, dim3(grid), dim3(TILE * LPE), smem, stream, p, b, step, env_off, do_push,
                    thr, hmap);
    B200_LAUNCH_CHECK("post_physics");
#if PP_FINALIZE_KERNEL
    b200_launch_pdl(p.num_envs, extras_finalize_kernel, dim3(1), dim3(32), 0, stream, p, b);
    B200_LAUNCH_CHECK("extras_finalize");
#endif
    return B200GYM_OK;
}

template <int TILE, bool ROUGH, bool TRAJ = false>
int launch_post_physics(const B200LeggedParams& p, const B200LeggedBuffers& b, uint64_t step, long long env_off, int do_push,
                        cudaStream_t stream) {
    // rough envs on a heightfield: the terrain window of every env is staged in shared memory by a 2-D TMA box load (B200GYM_HEIGHT_TMA=0:
    // per-point gathers from the field, the round-1 path — A/B in profiles/)
    static int want_tma = -1;
    if (want_tma < 0) {
        const char* t = getenv("B200GYM_HEIGHT_TMA");
        want_tma = t ? atoi(t) : 1;
    }
    CUtensorMap hmap;
    memset(&hmap, 0, sizeof(hmap));
    const bool within = p.horizontal_scale > 0.0f && 0.95f / p.horizontal_scale <= static_cast<float>(HW_HALF) - 0.5f;   // the scan fits the window
    const int use_hmap = (ROUGH && want_tma && !p.mesh_plane && within && p.terrain_rows < 32768 && p.terrain_cols < 32768 && height_tensor_map(b.height_samples, p.terrain_rows, p.terrain_cols, &hmap)) ? 1 : 0;
    if constexpr (ROUGH) {
        if (use_hmap) return launch_post_physics_impl<TILE, true, TRAJ, true>(p, b, step, env_off, do_push, stream, hmap);
    }
    return launch_post_physics_impl<TILE, ROUGH, TRAJ, false>(p, b, step, env_off, do_push, stream, hmap);
}

}  // namespace

extern "C" int b200gym_post_physics(const B200LeggedParams* p, const B200LeggedBuffers* b, uint64_t step, int64_t env_id_offset,
                                    void* stream) {
    B200_REQUIRE(p && b, B200GYM_EINVAL, "post_physics: null argument");
    B200_REQUIRE(p->num_envs > 0, B200GYM_EINVAL, "post_physics: num_envs must be positive (got %d)", p->num_envs);
    B200_REQUIRE(p->num_bodies > 0 && p->num_bodies <= 64, B200GYM_EINVAL, "post_physics: num_bodies %d out of range", p->num_bodies);
    B200_REQUIRE(p->num_sum_rows >= 0 && p->num_sum_rows <= B200GYM_NUM_REWARD_TERMS, B200GYM_EINVAL, "post_physics: bad num_sum_rows");
    B200_REQUIRE(p->resample_steps > 0, B200GYM_EINVAL, "post_physics: resample_steps must be positive");
    B200_REQUIRE(p->num_term >= 1 && p->num_term <= B200GYM_MAX_TERM, B200GYM_EINVAL, "post_physics: 1..4 termination bodies");
    const bool rough = p->num_heights > 0;
    const bool traj = p->traj_mode != 0;
    const int ow = traj ? ObsLayout<true>::OW : ObsLayout<false>::OW;
    B200_REQUIRE(p->num_obs == ow + p->num_heights, B200GYM_EINVAL, "post_physics: num_obs %d != %d + %d heights", p->num_obs, ow,
                 p->num_heights);
    B200_REQUIRE(!traj || (p->traj_n == 2 && p->traj_horizon * p->traj_n == TRAJ_W), B200GYM_EINVAL,
                 "post_physics: trajectory block must be %d columns of a 2-state rom (got N=%d, n=%d)", TRAJ_W, p->traj_horizon, p->traj_n);
    B200_REQUIRE(!traj || (b->trajectory && b->prev_error && b->time_until_next_push && b200_aligned16(b->trajectory) &&
                           b200_aligned16(b->prev_error) && b200_aligned16(b->time_until_next_push)),
                 B200GYM_EINVAL, "post_physics: trajectory / prev_error / time_until_next_push buffers missing or misaligned");
    B200_REQUIRE(!traj || !p->terrain_curriculum, B200GYM_EINVAL,
                 "post_physics: the trajectory env has no commands for the terrain curriculum (legged_robot_trajectory.py:508)");
    B200_REQUIRE(!traj || (p->reward_scale[T_STAND_STILL] == 0.f && p->reward_scale[T_TRACKING_ANG_VEL] == 0.f &&
                           p->reward_scale[T_TRACKING_LIN_VEL] == 0.f),
                 B200GYM_EINVAL, "post_physics: command-based reward terms do not exist in the trajectory env");
    B200_REQUIRE(traj || (p->reward_scale[T_TRACKING_ROM] == 0.f && p->reward_scale[T_DIFFERENTIAL_ERROR] == 0.f), B200GYM_EINVAL,
                 "post_physics: tracking_rom / differential_error need traj_mode");
    B200_REQUIRE(!rough || (p->num_heights <= HPAD && p->n_px * p->n_py == p->num_heights && p->n_px <= B200GYM_MAX_POINTS &&
                            p->n_py <= B200GYM_MAX_POINTS),
                 B200GYM_EINVAL, "post_physics: unsupported height grid %dx%d", p->n_px, p->n_py);
    B200_REQUIRE(!rough || b->measured_heights, B200GYM_EINVAL, "post_physics: measured_heights buffer missing");
    B200_REQUIRE(!rough || p->mesh_plane || (b->height_samples && p->terrain_rows >= 2 && p->terrain_cols >= 2), B200GYM_EINVAL,
                 "post_physics: height_samples missing");
    B200_REQUIRE(!rough || p->mesh_plane || (p->terrain_rows < 32768 && p->terrain_cols < 32768), B200GYM_EINVAL,
                 "post_physics: heightfields of up to 32767 x 32767 cells (cell coordinates are packed 16 + 16 bits)");
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
    const int do_push = (p->push_robots && p->push_time > 0 && (step % static_cast<uint64_t>(p->push_time) == 0)) ? 1 : 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static int tile = 0;   // tuning knob (profiles/): envs per CTA
    if (tile == 0) {
        const char* t = getenv("B200GYM_TILE");
        tile = t ? atoi(t) : 32;
    }
    // full tiles move the episode_sums rows [K, N] with 16-byte bulk copies: every row start k*N must be 16-byte aligned unless
    // the only CTA is a partial tile (cooperative copies) — checked against the tile size actually launched
    const int tile_used = traj ? 32 : (tile == 32 || tile == 16 ? tile : 64);
    B200_REQUIRE(p->num_sum_rows == 0 || (b->episode_sums && b200_aligned16(b->episode_sums) && (p->num_envs % 4 == 0 || p->num_envs < tile_used)),
                 B200GYM_EALIGN, "post_physics: episode_sums rows must be 16-byte aligned (num_envs %% 4 == 0 once num_envs >= %d)", tile_used);
    if (traj) {
        if (rough) return launch_post_physics<32, true, true>(*p, *b, step, env_id_offset, 0, st);
        return launch_post_physics<32, false, true>(*p, *b, step, env_id_offset, 0, st);
    }
    if (tile == 32) {
        if (rough) return launch_post_physics<32, true>(*p, *b, step, env_id_offset, do_push, st);
        return launch_post_physics<32, false>(*p, *b, step, env_id_offset, do_push, st);
    }
    if (tile == 16) {
        if (rough) return launch_post_physics<16, true>(*p, *b, step, env_id_offset, do_push, st);
        return launch_post_physics<16, false>(*p, *b, step, env_id_offset, do_push, st);
    }
    if (rough) return launch_post_physics<64, true>(*p, *b, step, env_id_offset, do_push, st);
    return launch_post_physics<64, false>(*p, *b, step, env_id_offset, do_push, st);
}


extern "C" int b200gym_legged_reset_idx(const B200LeggedParams* p, const B200LeggedBuffers* b, const uint8_t* reset_mask, uint64_t event,
                                        int64_t env_id_offset, void* stream) {
    B200_REQUIRE(p && b && reset_mask, B200GYM_EINVAL, "legged_reset_idx: null argument");
    B200_REQUIRE(p->num_envs > 0 && p->num_sum_rows >= 0 && p->num_sum_rows <= B200GYM_NUM_REWARD_TERMS, B200GYM_EINVAL, "legged_reset_idx: bad sizes");
    B200_REQUIRE(p->traj_mode == 0 || (b->trajectory && b->prev_error && !p->terrain_curriculum), B200GYM_EINVAL,
                 "legged_reset_idx: traj_mode needs the trajectory / prev_error buffers (and has no terrain curriculum)");
    B200_REQUIRE(!p->terrain_curriculum || (b->terrain_levels && b->terrain_types && b->terrain_origins), B200GYM_EINVAL,
                 "legged_reset_idx: terrain curriculum buffers missing");
    B200_REQUIRE(!p->zero_lstm_on_reset || (b->lstm_h && b->lstm_c), B200GYM_EINVAL, "legged_reset_idx: LSTM state buffers missing");
    const void* must[] = {b->root_states, b->dof_state, b->last_actions, b->last_dof_vel, b->commands, b->feet_air_time, b->episode_length_buf,
                          b->reset_buf, b->env_origins, b->extras_out, b->ws_sums};
    for (const void* q : must) {
        B200_REQUIRE(q != nullptr, B200GYM_EINVAL, "legged_reset_idx: null buffer");
        B200_REQUIRE(b200_aligned16(q), B200GYM_EALIGN, "legged_reset_idx: buffers must be 16-byte aligned");
    }
    B200_REQUIRE(p->num_sum_rows == 0 || b->episode_sums, B200GYM_EINVAL, "legged_reset_idx: episode_sums missing");
    static const SqThr thr = {sq_gt(1.0f), sq_gt(0.1f), sq_gt(0.2f), sq_lt(0.1f)};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    reset_idx_kernel<<<(p->num_envs + 127) / 128, 128, 0, st>>>(*p, *b, reset_mask, event, env_id_offset, thr);
    B200_LAUNCH_CHECK("legged_reset_idx");
#if PP_FINALIZE_KERNEL
    B200LeggedBuffers b2 = *b;
    b2.step_counter = nullptr;   // the means are finalised like a step's, but no env step has happened
    extras_finalize_kernel<<<1, 32, 0, st>>>(*p, b2);
    B200_LAUNCH_CHECK("legged_reset_idx finalize");
#else
#error "legged_reset_idx needs the stand-alone extras finaliser (PP_FINALIZE_KERNEL)"
#endif
    return B200GYM_OK;
}
