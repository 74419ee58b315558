// Shared device/host helpers for the b200gym kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "b200gym kernels are written for sm_100a (Blackwell B200) only"
#endif

// ------------------------------------------------------------------------------------------------
// error convention of the C ABI: 0 = enqueued, <0 = error, message via b200gym_last_error()
// ------------------------------------------------------------------------------------------------
#define B200GYM_OK 0
#define B200GYM_EINVAL (-1)
#define B200GYM_EALIGN (-2)
#define B200GYM_ECUDA (-3)
#define B200GYM_EARCH (-4)

void b200gym_set_error(const char* fmt, ...);

#define B200_REQUIRE(cond, code, ...)                \
    do {                                             \
        if (!(cond)) {                               \
            b200gym_set_error(__VA_ARGS__);          \
            return (code);                           \
        }                                            \
    } while (0)

#define B200_LAUNCH_CHECK(name)                                                            \
    do {                                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) {                                                          \
            b200gym_set_error("%s: launch failed: %s", (name), cudaGetErrorString(e__));   \
            return B200GYM_ECUDA;                                                          \
        }                                                                                  \
    } while (0)

static inline bool b200_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------------------------------------
// fp32 helpers whose rounding must match torch's un-fused elementwise ops
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
// torch_rand_float / TrajectoryGenerator.uniform: (hi - lo) * u + lo with two roundings
__device__ __forceinline__ float affine_rn(float span, float u, float lo) { return __fadd_rn(__fmul_rn(span, u), lo); }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// ------------------------------------------------------------------------------------------------
// TMA bulk copies (1-D cp.async.bulk, SASS UBLKCP) + mbarrier: whole AoS tiles of rows move between
// HBM and shared memory as single asynchronous transactions, issued by one thread.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global (bulk async group)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// streaming (read-once / write-once) vector accesses that bypass L1 allocation
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Programmatic Dependent Launch: the kernels of one env step form a dependent chain of short launches (at 4096 envs the
// step is launch latency, not bandwidth).  Every kernel of the chain starts with pdl_wait() — it blocks until the previous
// grid has completed and its writes are visible — and is launched with the programmatic-stream-serialization attribute,
// so its CTAs are scheduled and run their prologue while the previous kernel drains.  pdl_launch_dependents() at the top of
// a kernel lets the NEXT launch start being scheduled as soon as all of this kernel's CTAs are running.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#include <stdlib.h>
#include <utility>
// Measured (profiles/r1_pdl_ab.txt): PDL shortens the eagerly launched step by 13-15 % up to 262 144 envs, is neutral under
// CUDA-graph replay at those sizes and costs 8 % at 1 M envs (early-resident dependents that only wait) — so it is on by
// default for grids up to 262 144 envs; B200GYM_PDL=0 / 1 forces it off / on.
static inline bool b200_pdl_enabled(long long units) {
    static int mode = -2;
    if (mode == -2) {
        const char* e = getenv("B200GYM_PDL");
        mode = e ? (e[0] == '0' ? 0 : 1) : -1;
    }
    return mode >= 0 ? mode != 0 : units <= 262144;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t b200_launch_pdl(long long units, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                          Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = b200_pdl_enabled(units) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
