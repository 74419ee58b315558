// Stand-alone probe of a 2-D TMA box load (cp.async.bulk.tensor.2d, SASS UTMALDG.2D) of an int16 heightfield-like array.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -lcuda ; ./tma_probe VARIANT [box_cols box_rows l2promo]
// VARIANT: 0 = descriptor as __grid_constant__ kernel parameter, `.tile` qualifier spelled out   1 = descriptor in global memory
//          2 = __grid_constant__ parameter, CUTLASS's instruction form (no `.tile`)              3 = as 2, one elected thread issues
// Each variant runs in its own process (tools/probes/run_tma_probe.sh): a faulting kernel poisons the context.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int VARIANT>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, int c0, int c1, int16_t* out, int box_elems, int nthreads) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    int16_t* dst = reinterpret_cast<int16_t*>(smem + 1024);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(box_elems * 2 * nthreads) : "memory");
    __syncthreads();
    if (threadIdx.x < nthreads) {
        const CUtensorMap* m = VARIANT == 1 ? gmap : &tmap;
        const uint32_t d = smem_u32(dst + threadIdx.x * box_elems), b = smem_u32(bar);
        const int x = c0 + (int)threadIdx.x, y = c1;
        if (VARIANT <= 1)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(d),
                         "l"(reinterpret_cast<uint64_t>(m)), "r"(x), "r"(y), "r"(b)
                         : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(d),
                         "l"(reinterpret_cast<uint64_t>(m)), "r"(b), "r"(x), "r"(y)
                         : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra.uni D;\nbra.uni W;\nD:\n}\n" ::"r"(smem_u32(bar)), "r"(0)
        : "memory");
    for (int i = threadIdx.x; i < box_elems * nthreads; i += blockDim.x) out[i] = dst[i];
}

int main(int argc, char** argv) {
    const int rows = 64, pitch = 128;
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int bc = argc > 2 ? atoi(argv[2]) : 32, br = argc > 3 ? atoi(argv[3]) : 16, promo = argc > 4 ? atoi(argv[4]) : 0;
    int drv = 0, rt = 0;
    cudaDriverGetVersion(&drv);
    cudaRuntimeGetVersion(&rt);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("variant %d box %d x %d promo %d | %s cc %d.%d driver %d runtime %d\n", variant, bc, br, promo, prop.name, prop.major, prop.minor, drv, rt);
    std::vector<int16_t> h(rows * pitch);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < pitch; ++c) h[r * pitch + c] = static_cast<int16_t>(r * 100 + c);
    int16_t *d, *o;
    cudaMalloc(&d, h.size() * 2);
    cudaMalloc(&o, bc * br * 2 * 32);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry %p q=%d (%s)\n", fn, (int)q, cudaGetErrorString(ge));
    alignas(64) CUtensorMap tm;
    const cuuint64_t gdim[2] = {pitch, rows};
    const cuuint64_t gstr[1] = {pitch * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(bc), static_cast<cuuint32_t>(br)};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeFn>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                               CU_TENSOR_MAP_SWIZZLE_NONE, static_cast<CUtensorMapL2promotion>(promo), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d, tensor map at %p (64-byte aligned: %d), global base %p\n", (int)r, (void*)&tm, (int)((reinterpret_cast<uintptr_t>(&tm) & 63) == 0), (void*)d);
    const unsigned long long* w = reinterpret_cast<const unsigned long long*>(&tm);
    printf("descriptor words: %016llx %016llx %016llx %016llx\n", w[0], w[1], w[2], w[3]);
    CUtensorMap* gm;
    cudaMalloc(&gm, sizeof(tm));
    cudaMemcpy(gm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    std::vector<int16_t> res(bc * br * 32);
    const int smem = 1024 + bc * br * 2 * 32;
    for (int nth : {1, 4, 32}) {
        if (variant == 3 && nth > 1) break;
        cudaError_t e;
#define RUN(V)                                                                                    \
    cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);              \
    probe<V><<<1, 128, smem>>>(tm, gm, 5, 3, o, bc * br, nth);
        if (variant == 0) { RUN(0) } else if (variant == 1) { RUN(1) } else { RUN(2) }
        e = cudaDeviceSynchronize();
        cudaMemcpy(res.data(), o, res.size() * 2, cudaMemcpyDeviceToHost);
        printf("nth %d: %s  first %d (want %d)  [t%d] row 1 col 1: %d (want %d)\n", nth, cudaGetErrorString(e), res[0], 3 * 100 + 5, nth - 1,
               res[(nth - 1) * bc * br + bc + 1], 4 * 100 + 5 + nth - 1 + 1);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
