// Micro-benchmark of the synchronisation primitives on the tcgen05 issue path (sm_100a): cycles per operation, one warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sync_cost sync_cost.cu && ./sync_cost
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__global__ void k(long long* out) {
    __shared__ __align__(8) unsigned long long bars[8];
    const uint32_t b0 = smem_u32(&bars[0]);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b0 + 8 * i), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int R = 64;
    long long t0, t1;
    uint32_t acc = 0;
    // (0) try_wait on a barrier whose waited phase (parity 1 = "previous phase") is already complete, all 32 lanes
    t0 = clock64();
    for (int i = 0; i < R; ++i) acc += mbar_try(b0, 1);
    t1 = clock64();
    if (lane == 0) out[0] = (t1 - t0) / R;
    // (1) same, one lane
    __syncwarp();
    t0 = clock64();
    if (lane == 0) for (int i = 0; i < R; ++i) acc += mbar_try(b0, 1);
    __syncwarp();
    t1 = clock64();
    if (lane == 0) out[1] = (t1 - t0) / R;
    // (2) test_wait, all lanes
    t0 = clock64();
    for (int i = 0; i < R; ++i) acc += mbar_test(b0, 1);
    t1 = clock64();
    if (lane == 0) out[2] = (t1 - t0) / R;
    // (3) elect.sync
    t0 = clock64();
    for (int i = 0; i < R; ++i) acc += elect_one();
    t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0) / R;
    // (4) tcgen05.commit (nothing outstanding) by an elected lane + try_wait for it to land (round trip)
    uint32_t ph = 0;
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b0 + 8) : "memory");
        __syncwarp();
        while (!mbar_try(b0 + 8, ph)) {}
        ph ^= 1;
    }
    t1 = clock64();
    if (lane == 0) out[4] = (t1 - t0) / R;
    // (5) tcgen05.commit issue cost only (arrivals accumulate on a count-64 barrier)
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b0 + 16), "r"(R));
    __syncwarp();
    t0 = clock64();
    if (elect_one()) for (int i = 0; i < R; ++i) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b0 + 16) : "memory");
    __syncwarp();
    t1 = clock64();
    if (lane == 0) out[5] = (t1 - t0) / R;
    // (6) plain mbarrier.arrive + try_wait round trip, one lane
    ph = 0;
    t0 = clock64();
    if (lane == 0) for (int i = 0; i < R; ++i) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b0 + 24) : "memory");
        while (!mbar_try(b0 + 24, ph)) {}
        ph ^= 1;
    }
    __syncwarp();
    t1 = clock64();
    if (lane == 0) out[6] = (t1 - t0) / R;
    // (7) tcgen05.fence::after_thread_sync
    t0 = clock64();
    for (int i = 0; i < R; ++i) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    t1 = clock64();
    if (lane == 0) out[7] = (t1 - t0) / R;
    // (8) __syncwarp
    t0 = clock64();
    for (int i = 0; i < R; ++i) __syncwarp();
    t1 = clock64();
    if (lane == 0) out[8] = (t1 - t0) / R;
    // (9) try_wait all lanes, each lane a different (complete) barrier
    t0 = clock64();
    for (int i = 0; i < R; ++i) acc += mbar_try(b0 + 8 * (lane & 3) + 32, 1);
    t1 = clock64();
    if (lane == 0) out[9] = (t1 - t0) / R;
    if (acc == 0xdeadbeef) out[15] = acc;
}

int main() {
    long long* d;
    cudaMalloc(&d, 16 * sizeof(long long));
    cudaMemset(d, 0, 16 * sizeof(long long));
    k<<<1, 32>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[] = {"try_wait (complete), 32 lanes", "try_wait (complete), 1 lane", "test_wait, 32 lanes", "elect.sync + selp",
                           "tcgen05.commit -> try_wait round trip", "tcgen05.commit issue", "mbarrier.arrive -> try_wait round trip (1 lane)",
                           "tcgen05.fence::after_thread_sync", "__syncwarp", "try_wait 32 lanes, 4 distinct barriers"};
    printf("status %s\n", cudaGetErrorString(e));
    for (int i = 0; i < 10; ++i) printf("%-52s %lld cycles\n", names[i], h[i]);
    return 0;
}
