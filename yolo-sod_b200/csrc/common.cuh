// Shared device/host helpers for the yolo-sod B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

// ---- 16-bit storage type of the build ---------------------------------------------------------------------------------
// libysod.so stores activations / weights as bf16 (north_star's low-precision mode). The SAME sources compiled with
// -DYSOD_HALF=1 give libysod_f16.so, the reference's own half mode (`model.half()`, nn/autobackend.py:154): IEEE fp16 storage
// and fp16 tensor-core inputs, fp32 accumulation and epilogue math unchanged. Dtype code 1 (YSOD_BF16) means "the build's
// 16-bit type" at the C ABI; ysod_storage_dtype() tells which one a library was built for.
#ifdef YSOD_HALF
#define __nv_bfloat16 __half
#define __nv_bfloat162 __half2
#define __float2bfloat16_rn __float2half_rn
#define __float2bfloat16 __float2half
#define __floats2bfloat162_rn __floats2half2_rn
#define __bfloat162float __half2float
#define __bfloat1622float2 __half22float2
#define YSOD_MMA_T "f16"                                    // mma.sync operand type
#define YSOD_TMAP_16 CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define YSOD_UMMA_AB_FORMAT 0u                              // tcgen05 kind::f16 instruction descriptor: A / B format 0 = f16
#else
#define YSOD_MMA_T "bf16"
#define YSOD_TMAP_16 CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define YSOD_UMMA_AB_FORMAT 1u                              // 1 = bf16
#endif

#define YSOD_OK 0
#define YSOD_ERR_INVALID 1
#define YSOD_ERR_CUDA 2
#define YSOD_ERR_UNSUPPORTED 3
#define YSOD_ERR_WORKSPACE 4

// thread-local last-error text, exposed through ysod_last_error()
void ysod_set_error(const char* fmt, ...);

#define YSOD_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            ysod_set_error(__VA_ARGS__);          \
            return YSOD_ERR_INVALID;              \
        }                                         \
    } while (0)

#define YSOD_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            ysod_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return YSOD_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define YSOD_LAUNCH_CHECK()                                                               \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            ysod_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return YSOD_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

static inline int ysod_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// dtype codes used across the C ABI
#define YSOD_F32 0
#define YSOD_BF16 1

// ysod_stem_mma src_fmt flag (ysod.h): the image argument is a device slot holding the image pointer
#define YSOD_STEM_INDIRECT 0x10

// activation codes for fused epilogues
#define YSOD_ACT_NONE 0
#define YSOD_ACT_SILU 1
#define YSOD_ACT_GELU 2   // exact erf GELU (torch.nn.GELU default)
#define YSOD_ACT_RELU 3
#define YSOD_ACT_SIGMOID 4
#define YSOD_ACT_HSIGMOID 5  // relu6(x+3)/6

#ifdef __CUDACC__
// Programmatic dependent launch (PDL): every kernel is launched with programmaticStreamSerializationAllowed, so its CTAs may be
// scheduled while the previous kernel in the stream is still draining; ysod_pdl_sync() at the top of the kernel blocks until
// the previous grid has completed and its memory is visible, then lets the next kernel start launching in turn. This hides
// launch latency and CTA scheduling between the ~130 back-to-back kernels of one forward (inside or outside a CUDA graph).
__device__ __forceinline__ void ysod_pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t ysod_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename T> __device__ __forceinline__ float ysod_ld(const T* p);
template <> __device__ __forceinline__ float ysod_ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ysod_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void ysod_st(T* p, float v);
template <> __device__ __forceinline__ void ysod_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void ysod_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// two packed 16-bit activations (one 32-bit word) -> floats
__device__ __forceinline__ float2 ysod_unpack2(uint32_t v) {
#ifdef YSOD_HALF
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
#else
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
#endif
}

// L2 prefetch of the 128 B line holding p (no register, no fault). Measured (round 2): pays in the tiled CBAM apply kernel, where the
// tile's lines travel while the 7x7 filter runs (59.8 -> 53.7 us); requesting lines two iterations ahead in the grid-stride streaming
// kernels (pool / scale / CoordAtt apply / CBAM statistics) changed nothing or cost 3-5 %, so those do not prefetch.
__device__ __forceinline__ void ysod_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ float ysod_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7): one rcp + one exp instead of erff's long polynomial path
__device__ __forceinline__ float ysod_erf_fast(float x) {
    const float ax = fabsf(x);
    const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = 1.0f - p * t * __expf(-ax * ax);
    return copysignf(e, x);
}

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) for the 16-bit tensor-core paths: erf(z) ~= tanh(a z + b z^3 + c z^5) (least-squares fit,
// max |err| 4.1e-5 on erf, 4.9e-5 on GELU), one MUFU (tanh.approx, the form the SiLU epilogues use) and 6 FP32 ops instead of
// two MUFUs and ~14 ops for the A&S erf. The error is far below the bf16 / fp16 rounding of the value it produces; the fp32
// parity mode keeps erff (ysod_act).
__device__ __forceinline__ float ysod_gelu_tanh(float x) {
    const float x2 = x * x;
    float q = fmaf(-3.2060743e-4f, x2, 3.6819429e-2f);      // c / 2^(5/2), b / 2^(3/2)
    q = fmaf(q, x2, 7.9770428e-1f);                          // a / sqrt 2
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * q));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

__device__ __forceinline__ float ysod_act(float x, int act) {
    switch (act) {
        case YSOD_ACT_SILU: return x / (1.0f + expf(-x));
        case YSOD_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
        case YSOD_ACT_RELU: return fmaxf(x, 0.0f);
        case YSOD_ACT_SIGMOID: return ysod_sigmoid(x);
        case YSOD_ACT_HSIGMOID: return fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
        default: return x;
    }
}

// 8 consecutive channels (16 B of bf16 / 32 B of fp32) <-> 8 floats
template <typename T> struct ysod_vec8;
template <> struct ysod_vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) {
        uint4 r = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* v) {
        uint4 r;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = r;
    }
};
template <> struct ysod_vec8<float> {
    static __device__ __forceinline__ void load(const float* p, float* v) {
        float4 a = *reinterpret_cast<const float4*>(p);
        float4 b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};

__device__ __forceinline__ float ysod_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float ysod_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif
