// Philox4x32-10 counter-based RNG.  Specification: oracle/philox.py (bit-for-bit).
//   key     = (seed_lo, seed_hi)
//   counter = (global_env_id, event_lo, event_hi, (site << 16) | block)
// Column j of a draw = word j%4 of block j/4; uniform = (x >> 8) * 2^-24; bounded int = mulhi(x, bound).
#pragma once
#include <stdint.h>

namespace philox {

enum Site : uint32_t {
    CMD_PERIODIC = 1, PUSH = 2, TERRAIN = 3, RESET_DOF = 4, RESET_XY = 5, RESET_VEL = 6, CMD_RESET = 7, OBS_NOISE = 8, PUSH_TIMER = 9,
    ROM_INIT = 16, ROM_ROOT = 17, ROM_DIST_MASK = 18, ROM_DIST = 19, ROM_CONST = 20, ROM_RAMP = 21, ROM_EXTREME = 22,
    ROM_SIN_MAG = 23, ROM_SIN_MEAN = 24, ROM_SIN_FREQ = 25, ROM_SIN_OFF = 26, ROM_TFINAL = 27, ROM_WEIGHTS = 28,
    ROM_STATIONARY = 29,
    // HopperTrajectory reset / push (oracle/philox.py; reserved, no kernel draws from them yet)
    HOP_DOF_POS = 32, HOP_DOF_VEL = 33, HOP_ROOT_POS = 34, HOP_YAW = 35, HOP_ROOT_VEL = 36, HOP_PUSH = 37,
    // PPO.act: Normal(mu, std).sample(), event = the runner's act counter (csrc/ppo_rollout.cu)
    POLICY_SAMPLE = 48
};

__device__ __forceinline__ uint4 block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = static_cast<uint64_t>(M0) * c0;
        const uint64_t p1 = static_cast<uint64_t>(M1) * c2;
        const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
        const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

struct Stream {
    uint32_t k0, k1, env, ev_lo, ev_hi;
    __device__ __forceinline__ Stream(uint32_t seed_lo, uint32_t seed_hi, uint64_t env_id, uint64_t event)
        : k0(seed_lo), k1(seed_hi), env(static_cast<uint32_t>(env_id)), ev_lo(static_cast<uint32_t>(event)),
          ev_hi(static_cast<uint32_t>(event >> 32)) {}
    __device__ __forceinline__ uint4 words(uint32_t site, uint32_t blk) const {
        return block(env, ev_lo, ev_hi, (site << 16) | blk, k0, k1);
    }
};

__device__ __forceinline__ float u01(uint32_t x) { return static_cast<float>(x >> 8) * 5.9604644775390625e-08f; }
__device__ __forceinline__ float4 u01(const uint4& w) { return make_float4(u01(w.x), u01(w.y), u01(w.z), u01(w.w)); }
__device__ __forceinline__ uint32_t bounded(uint32_t x, uint32_t bound) { return __umulhi(x, bound); }
__device__ __forceinline__ uint32_t word(const uint4& w, int i) { return i == 0 ? w.x : i == 1 ? w.y : i == 2 ? w.z : w.w; }

}  // namespace philox
