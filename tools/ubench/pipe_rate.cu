// Micro-benchmark: per-SM throughput of the scalar / packed fp32, integer, conversion and MUFU instructions the row-wise stages of the
// fused SwinBlock kernel are made of (sm_100a), with 8 and 16 resident warps per SM (2 / 4 per scheduler) and 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rate pipe_rate.cu && ./pipe_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long p2;
__device__ __forceinline__ p2 fma2(p2 a, p2 b, p2 c) { p2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ p2 add2(p2 a, p2 b) { p2 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float tanh_a(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_a(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t lop(uint32_t a, uint32_t b) { uint32_t d; asm volatile("and.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t shl(uint32_t a) { uint32_t d; asm volatile("shl.b32 %0, %1, 16;" : "=r"(d) : "r"(a)); return d; }
__device__ __forceinline__ uint32_t cvtpk(float a, float b) { uint32_t d; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(b)); return d; }

__device__ __forceinline__ uint32_t tanh_h2(uint32_t x) { uint32_t y; asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t tanh_b2(uint32_t x) { uint32_t y; asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t cvth2(float a, float b) { uint32_t d; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(b)); return d; }

constexpr int ITERS = 512, CH = 8;

template <int OP>
__global__ void k(long long* out, float* sink, float seed) {
    float f[CH];
    p2 q[CH];
    uint32_t u[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { f[c] = seed + c + threadIdx.x; u[c] = threadIdx.x * 77u + c; q[c] = ((p2)__float_as_uint(f[c]) << 32) | __float_as_uint(f[c]); }
    const p2 qa = ((p2)__float_as_uint(1.0001f) << 32) | __float_as_uint(0.9999f), qb = ((p2)__float_as_uint(1e-3f) << 32) | __float_as_uint(1e-3f);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (OP == 0) f[c] = ffma(f[c], 1.0001f, 1e-3f);
            if (OP == 1) q[c] = fma2(q[c], qa, qb);
            if (OP == 2) q[c] = add2(q[c], qb);
            if (OP == 3) f[c] = tanh_a(f[c]);
            if (OP == 4) f[c] = ex2_a(f[c]);
            if (OP == 5) f[c] = rcp_a(f[c]);
            if (OP == 6) u[c] = lop(u[c], 0xffff0fffu + c);
            if (OP == 7) u[c] = shl(u[c]);
            if (OP == 8) u[c] = cvtpk(__uint_as_float(u[c]), f[c]);
            if (OP == 9) u[c] = tanh_h2(u[c]);
            if (OP == 10) u[c] = tanh_b2(u[c]);
            if (OP == 11) u[c] = ex2_b2(u[c]);
            if (OP == 12) u[c] = ex2_h2(u[c]);
            if (OP == 13) u[c] = cvth2(__uint_as_float(u[c]), f[c]);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += f[c] + __uint_as_float(u[c]) + __uint_as_float((uint32_t)q[c]) + __uint_as_float((uint32_t)(q[c] >> 32));
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>
void run(const char* name, long long* d_out, float* sink) {
    for (int threads : {128, 256, 512}) {
        k<OP><<<148, threads>>>(d_out, sink, 0.5f);
        cudaDeviceSynchronize();
        k<OP><<<148, threads>>>(d_out, sink, 0.5f);
        cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, d_out, sizeof(c), cudaMemcpyDeviceToHost);
        const double warp_instr = (double)ITERS * CH * (threads / 32);
        printf("%-28s %4d threads/SM: %7.2f cycles per warp instruction per SM  (%6.1f lanes/clk/SM)\n", name, threads, (double)c / warp_instr,
               warp_instr * 32.0 / (double)c);
    }
}

int main() {
    long long* d_out;
    float* sink;
    cudaMalloc(&d_out, 64);
    cudaMalloc(&sink, 148 * 512 * sizeof(float));
    run<0>("fma.rn.f32 (FFMA)", d_out, sink);
    run<1>("fma.rn.f32x2 (FFMA2)", d_out, sink);
    run<2>("add.rn.f32x2 (FADD2)", d_out, sink);
    run<3>("tanh.approx.f32 (MUFU.TANH)", d_out, sink);
    run<4>("ex2.approx.f32 (MUFU.EX2)", d_out, sink);
    run<5>("rcp.approx.f32 (MUFU.RCP)", d_out, sink);
    run<6>("and.b32 (LOP3)", d_out, sink);
    run<7>("shl.b32 (SHF)", d_out, sink);
    run<8>("cvt.rn.bf16x2.f32 (F2FP)", d_out, sink);
    run<9>("tanh.approx.f16x2", d_out, sink);
    run<10>("tanh.approx.bf16x2", d_out, sink);
    run<11>("ex2.approx.ftz.bf16x2", d_out, sink);
    run<12>("ex2.approx.f16x2", d_out, sink);
    run<13>("cvt.rn.f16x2.f32", d_out, sink);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
