// Micro-benchmark: tcgen05.mma (kind::f16, M=128, cta_group::1, SS operands, SWIZZLE_128B K-major) issue/execute rate vs N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(accum) : "memory");
}

// mode 0: every MMA reads a different A/B block (streaming, as in the conv main loop); mode 1: same A and B every time
template <int N, int MODE>
__global__ void k(long long* out, int slot) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tslot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // zero the operand area (bf16 zeros: no NaN slow paths)
    for (uint32_t i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem + (base - smem_u32(smem)))[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        constexpr int R = 144;  // MMAs per measurement
        long long t0 = 0, t1 = 0, t2 = 0;
        if (elect_one()) {
            t0 = clock64();
#pragma unroll 12
            for (int i = 0; i < R; ++i) {
                // A blocks of 16 KB (128 rows x 128 B), 4 K-steps of 32 B inside; B blocks of N*128 B
                const uint32_t blk = MODE ? 0u : (uint32_t)(i >> 2) % 4u;
                const uint32_t a_addr = base + blk * 16384u + 32u * (i & 3);
                const uint32_t b_addr = base + 65536u + blk * (uint32_t)(N * 128) + 32u * (i & 3);
                const uint64_t ad = ((uint64_t)desc_hi << 32) | (((a_addr >> 4) & 0x3FFFu) | (1u << 16));
                const uint64_t bd = ((uint64_t)desc_hi << 32) | (((b_addr >> 4) & 0x3FFFu) | (1u << 16));
                mma(tmem, ad, bd, idesc, i != 0);
            }
            t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        t2 = clock64();
        t0 = __shfl_sync(0xffffffffu, t0, 0);  // elected lane is lane 0 of a full warp
        t1 = __shfl_sync(0xffffffffu, t1, 0);
        if (threadIdx.x == 0) { out[2 * slot] = (t1 - t0) * 100 / R; out[2 * slot + 1] = (t2 - t0) * 100 / R; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, int MODE>
void run(long long* d, int slot, int ctas) {
    cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<N, MODE><<<ctas, 128, 200 * 1024>>>(d, slot);
    k<N, MODE><<<ctas, 128, 200 * 1024>>>(d, slot);
}

int main() {
    long long* d;
    cudaMalloc(&d, 64 * sizeof(long long));
    cudaMemset(d, 0, 64 * sizeof(long long));
    run<32, 0>(d, 0, 1); run<64, 0>(d, 1, 1); run<128, 0>(d, 2, 1); run<256, 0>(d, 3, 1);
    run<32, 1>(d, 4, 1); run<64, 1>(d, 5, 1); run<128, 1>(d, 6, 1); run<256, 1>(d, 7, 1);
    run<64, 0>(d, 8, 148); run<128, 0>(d, 9, 148); run<256, 0>(d, 10, 148);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[64];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("status %s\n", cudaGetErrorString(e));
    const char* names[] = {"N=32 stream", "N=64 stream", "N=128 stream", "N=256 stream", "N=32 same", "N=64 same", "N=128 same", "N=256 same",
                           "N=64 stream x148 CTAs", "N=128 stream x148", "N=256 stream x148"};
    for (int i = 0; i < 11; ++i)
        printf("%-24s issue %6.2f cyc/MMA   issue+drain %6.2f cyc/MMA   (floor N/2 = %d)\n", names[i], h[2 * i] / 100.0, h[2 * i + 1] / 100.0,
               (i < 8 ? (32 << (i & 3)) : (64 << (i - 8))) / 2);
    return 0;
}
