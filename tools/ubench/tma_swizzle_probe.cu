// Probe: where does TMA put a box whose inner dimension (64 B) is SHORTER than the swizzle span (SWIZZLE_128B)?
// Loads a {32 ch, 8 px, 4 rows} bf16 box (64 B inner rows) and dumps shared memory: prints, per 16 B chunk of smem, the
// source chunk it holds. Dense + address-based swizzle  <=>  chunk at byte o holds source chunk  (o ^ (((o >> 7) & 7) << 4)) / 16.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_swizzle_probe tma_swizzle_probe.cu -lcuda && ./tma_swizzle_probe
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(const __grid_constant__ CUtensorMap tm, uint32_t* out, int words) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    uint32_t* s = reinterpret_cast<uint32_t*>(smem + (base - smem_u32(smem)));
    for (int i = threadIdx.x; i < words; i += blockDim.x) s[i] = 0xFFFFFFFFu;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(32 * 8 * 4 * 2) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(base), "l"((uint64_t)&tm), "r"(smem_u32(&bar)), "r"(0), "r"(0), "r"(0) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < words; i += blockDim.x) out[i] = s[i];
}

int main() {
    const int C = 32, W = 8, H = 4;
    uint16_t h[C * W * H];
    for (int i = 0; i < C * W * H; ++i) h[i] = (uint16_t)(i / 8);   // value = index of the 16 B source chunk
    uint16_t* d; cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    CUtensorMap tm;
    cuuint64_t dims[3] = {C, W, H}; cuuint64_t strides[2] = {C * 2, C * W * 2}; cuuint32_t box[3] = {C, W, H}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((Enc)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    const int words = 8192 / 4;
    uint32_t* o; cudaMalloc(&o, words * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    k<<<1, 128, 8192 + 1024>>>(tm, o, words);
    printf("run: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    uint32_t ho[words]; cudaMemcpy(ho, o, words * 4, cudaMemcpyDeviceToHost);
    int dense_ok = 1;
    for (int c = 0; c < 8192 / 16; ++c) {
        const uint32_t v = ho[c * 4];
        const int src = (v == 0xFFFFFFFFu) ? -1 : (int)(v & 0xFFFF);
        if (c % 8 == 0) printf("\n%5d:", c * 16);
        printf(" %4d", src);
        const int o16 = c * 16, expect = (o16 < C * W * H * 2) ? ((o16 ^ (((o16 >> 7) & 7) << 4)) / 16) : -1;
        if (src != expect) dense_ok = 0;
    }
    printf("\ndense + address-bit swizzle: %s\n", dense_ok ? "YES" : "NO");
    return 0;
}
