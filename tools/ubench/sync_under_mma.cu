// Micro-benchmark: cost of mbarrier.try_wait (on an already-complete barrier) and tcgen05.commit issued by the MMA thread
// right after a burst of tcgen05.mma, i.e. while the tensor pipe is draining.   ./sync_under_mma
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
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}

__global__ void k(long long* out, int nmma) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[8];
    __shared__ uint32_t tslot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t b0 = smem_u32(&bars[0]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b0 + 8 * i), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (uint32_t i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem + (base - smem_u32(smem)))[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        long long t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t acc = 0;
        for (int rep = 0; rep < 4; ++rep) {
            t[0] = clock64();
            if (elect_one()) {
                for (int i = 0; i < nmma; ++i) {
                    const uint32_t blk = (uint32_t)(i >> 2) % 4u;
                    const uint32_t a_addr = base + blk * 16384u + 32u * (i & 3);
                    const uint32_t b_addr = base + 65536u + blk * 8192u + 32u * (i & 3);
                    mma(tmem, ((uint64_t)desc_hi << 32) | (((a_addr >> 4) & 0x3FFFu) | (1u << 16)),
                        ((uint64_t)desc_hi << 32) | (((b_addr >> 4) & 0x3FFFu) | (1u << 16)), idesc, i != 0);
                }
                t[1] = clock64();
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b0 + 8) : "memory");
                t[2] = clock64();
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b0 + 16) : "memory");
                t[3] = clock64();
            }
            __syncwarp();
            t[4] = clock64();
            acc += mbar_try(b0, 1);   // complete barrier, all lanes
            t[5] = clock64();
            acc += mbar_try(b0 + 24, 1);
            t[6] = clock64();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            while (!mbar_try(b0 + 8, rep & 1)) {}
            while (!mbar_try(b0 + 16, rep & 1)) {}
            t[7] = clock64();
        }
        if (lane == 0) {
            out[0] = t[1] - t[0]; out[1] = t[2] - t[1]; out[2] = t[3] - t[2]; out[3] = t[4] - t[3];
            out[4] = t[5] - t[4]; out[5] = t[6] - t[5]; out[6] = t[7] - t[6]; out[7] = acc;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 64 * sizeof(long long));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int nmma : {0, 4, 12, 36}) {
        cudaMemset(d, 0, 64 * sizeof(long long));
        k<<<1, 128, 200 * 1024>>>(d, nmma);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[8];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("nmma %2d (%s): issue %lld | commit#1 %lld | commit#2 %lld | reconverge %lld | try_wait(done) %lld | try_wait(done) %lld | wait own commits %lld\n",
               nmma, cudaGetErrorString(e), h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    }
    return 0;
}
