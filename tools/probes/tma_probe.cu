// Stand-alone probe of a 2-D TMA box load (cp.async.bulk.tensor.2d) of an int16 field: nvcc -arch=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int MODE>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, int c0, int c1, int16_t* out, int box_elems, int nthreads) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    int16_t* dst = reinterpret_cast<int16_t*>(smem + 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(box_elems * 2 * nthreads) : "memory");
    if (threadIdx.x < nthreads) {
        const CUtensorMap* m = MODE == 0 ? &tmap : gmap;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         smem_u32(dst + threadIdx.x * box_elems)),
                     "l"(reinterpret_cast<uint64_t>(m)), "r"(c0 + (int)threadIdx.x), "r"(c1), "r"(smem_u32(bar))
                     : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra.uni D;\nbra.uni W;\nD:\n}\n" ::"r"(smem_u32(bar)), "r"(0)
        : "memory");
    for (int i = threadIdx.x; i < box_elems * nthreads; i += blockDim.x) out[i] = dst[i];
}

int main(int argc, char** argv) {
    const int rows = 64, pitch = 128;
    const int bc = argc > 1 ? atoi(argv[1]) : 32, br = argc > 2 ? atoi(argv[2]) : 22, promo = argc > 3 ? atoi(argv[3]) : 2;
    printf("box %d x %d promo %d\n", bc, br, promo);
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
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry %p q=%d\n", fn, (int)q);
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {pitch, rows};
    const cuuint64_t gstr[1] = {pitch * 2};
    const cuuint32_t box[2] = {bc, br};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeFn>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                               CU_TENSOR_MAP_SWIZZLE_NONE, static_cast<CUtensorMapL2promotion>(promo), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    CUtensorMap* gm;
    cudaMalloc(&gm, sizeof(tm));
    cudaMemcpy(gm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    std::vector<int16_t> res(bc * br * 32);
    for (int mode = 0; mode < 2; ++mode)
        for (int nth : {1, 4, 32}) {
            cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + bc * br * 2 * 32);
            cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + bc * br * 2 * 32);
            if (mode == 0) probe<0><<<1, 128, 128 + bc * br * 2 * 32>>>(tm, gm, 5, 3, o, bc * br, nth);
            else probe<1><<<1, 128, 128 + bc * br * 2 * 32>>>(tm, gm, 5, 3, o, bc * br, nth);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(res.data(), o, res.size() * 2, cudaMemcpyDeviceToHost);
            printf("mode %d nth %d: %s  first %d (want %d)  [t%d] %d (want %d)\n", mode, nth, cudaGetErrorString(e), res[0], 3 * 100 + 5, nth - 1,
                   res[(nth - 1) * bc * br + bc + 1], 4 * 100 + 5 + nth - 1 + 1);
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
