// tcgen05 / TMEM / cp.async helpers shared by the tensor-core kernels (sm_100a only).
//
// Shared-memory operand layout used everywhere in this library ("chunk layout", no swizzle): a [R rows x C cols] fp16 tile is
// stored as 16-byte pieces of 8 consecutive columns; piece (r, c8) lives at  c8 * chunk_stride + r * 16  with
// chunk_stride = R * 16 + 16 (the 16-byte pad spreads the loaders' scattered 16-byte stores over the banks).  Eight
// consecutive rows of one piece column are one contiguous 128-byte UMMA core matrix, so the SAME buffer is a legal operand
// in both majors (cute::UMMA::make_umma_desc, SWIZZLE_NONE / "interleave" canonical layouts):
//   * K-major  (rows = M or N index, cols = K):  LBO = chunk_stride (next 8 k),  SBO = 128 (next 8 rows);  16 k per MMA = 2 pieces
//   * MN-major (rows = K, cols = M or N index):  LBO = 128 (next 8 k = next 8 rows),  SBO = chunk_stride (next 8 m/n); 16 k = 256 B
// That is what lets one row-major activation matrix feed the forward GEMM (K-major A), the input-gradient GEMM (K-major A)
// and the weight-gradient GEMM (MN-major A and B, K = batch rows) without any transposed copy.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace tc {

__device__ __forceinline__ uint64_t smem_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout NONE [61,64)
    return (static_cast<uint64_t>(smem_u32(smem) >> 4) & 0x3FFFull) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
           (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor for kind::f16, fp16 operands, fp32 accumulator, M = 128:
// c_format F32 = 1 [4,6); a/b_format F16 = 0; a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major); N >> 3 [17,23); M >> 4 [24,29)
__device__ __forceinline__ uint32_t idesc_f16(int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) | (static_cast<uint32_t>(n >> 3) << 17) | (8u << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// the registers are tied to the wait so that no use of them can be scheduled above it
__device__ __forceinline__ void ld16_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) {
    }
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) __nanosleep(64);
}
// 16-byte asynchronous global -> shared copy; src_bytes = 0 zero-fills the destination (out-of-range pieces)
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) {
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// nn.ELU(alpha=1): exp(t) - 1 >= t for every t, so elu(t) = max(t, exp(min(t, 0)) - 1)
__device__ __forceinline__ float elu_fast(float t) { return fmaxf(t, ex2_approx(fminf(t, 0.0f) * 1.4426950408889634f) - 1.0f); }

}  // namespace tc
