// CUDA-core convolutions for the shapes that do not belong on the tensor cores, and for the fp32 parity mode.
//
// Replaces (reference): ultralytics/nn/modules/conv.py:37-55 Conv / :102 DWConv for
//   * the Cin=3 stem (K = 27, HBM-bound; also performs the NCHW fp32 -> NHWC conversion of the input image),
//   * depthwise convs (SwinBlock.dw 3x3 blocks_transformer.py:137, AAttn.pe 5x5 block.py:1291, Detect DWConv head.py:51-52),
//   * grouped convs of yolov12 (yolov12.yaml:20,22),
//   * every conv when the model runs in fp32 mode (north_star: rtol 1e-4 mode cannot ride on bf16/tf32 MMA).
// Layouts: activations NHWC (channel-sliced views allowed), weights [Cout][kh][kw][Cin/groups] for dense/grouped,
// [kh][kw][C] for depthwise, [Cout][kh][kw][3] fp32 for the stem. Epilogue = +bias -> act -> (+residual).
#include "common.cuh"

namespace {

template <typename TI, typename TO>
__global__ void conv_direct_kernel(const TI* __restrict__ x, const TI* __restrict__ w, const float* __restrict__ bias,
                                   const TI* __restrict__ res, TO* __restrict__ out, int N, int H, int W, int Cin, int xcs,
                                   int Ho, int Wo, int Cout, int ocs, int rcs, int k, int s, int pad, int groups, int act) {
    ysod_pdl_sync();
    const int cog_n = (Cout + 3) >> 2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)N * Ho * Wo * cog_n;
    if (idx >= total) return;
    const int cog = (int)(idx % cog_n);
    const long long pix = idx / cog_n;
    const int ow = (int)(pix % Wo);
    const int oh = (int)((pix / Wo) % Ho);
    const int n = (int)(pix / ((long long)Wo * Ho));
    const int co0 = cog * 4;
    const int cpg = Cin / groups, opg = Cout / groups;
    const int g = co0 / opg;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int nco = (Cout - co0) < 4 ? (Cout - co0) : 4;
    for (int r = 0; r < k; ++r) {
        const int ih = oh * s + r - pad;
        if (ih < 0 || ih >= H) continue;
        for (int q = 0; q < k; ++q) {
            const int iw = ow * s + q - pad;
            if (iw < 0 || iw >= W) continue;
            const TI* xp = x + (((size_t)n * H + ih) * W + iw) * xcs + g * cpg;
            const TI* wp = w + ((size_t)(co0 * k + r) * k + q) * cpg;
            const size_t wstride = (size_t)k * k * cpg;
            if ((cpg & 7) == 0) {
                for (int c = 0; c < cpg; c += 8) {
                    float xv[8];
                    ysod_vec8<TI>::load(xp + c, xv);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < nco) {
                            float wv[8];
                            ysod_vec8<TI>::load(wp + j * wstride + c, wv);
#pragma unroll
                            for (int e = 0; e < 8; ++e) acc[j] = fmaf(xv[e], wv[e], acc[j]);
                        }
                    }
                }
            } else {
                for (int c = 0; c < cpg; ++c) {
                    const float xv = ysod_ld<TI>(xp + c);
                    for (int j = 0; j < nco; ++j) acc[j] = fmaf(xv, ysod_ld<TI>(wp + j * wstride + c), acc[j]);
                }
            }
        }
    }
    const size_t opix = (size_t)pix;
    for (int j = 0; j < nco; ++j) {
        float v = ysod_act(acc[j] + bias[co0 + j], act);
        if (res) v += ysod_ld<TI>(res + opix * rcs + co0 + j);
        ysod_st<TO>(out + opix * ocs + co0 + j, v);
    }
}

// depthwise k x k, stride s; thread = (pixel, 8 channels)
template <typename T>
__global__ void dwconv_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                              const T* __restrict__ res, T* __restrict__ out, int N, int H, int W, int C, int xcs, int Ho,
                              int Wo, int ocs, int rcs, int k, int s, int pad, int act) {
    ysod_pdl_sync();
    const int c8n = C >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)N * Ho * Wo * c8n;
    if (idx >= total) return;
    const int c0 = (int)(idx % c8n) * 8;
    const long long pix = idx / c8n;
    const int ow = (int)(pix % Wo);
    const int oh = (int)((pix / Wo) % Ho);
    const int n = (int)(pix / ((long long)Wo * Ho));
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bias[c0 + e];
    for (int r = 0; r < k; ++r) {
        const int ih = oh * s + r - pad;
        if (ih < 0 || ih >= H) continue;
        for (int q = 0; q < k; ++q) {
            const int iw = ow * s + q - pad;
            if (iw < 0 || iw >= W) continue;
            float xv[8], wv[8];
            ysod_vec8<T>::load(x + (((size_t)n * H + ih) * W + iw) * xcs + c0, xv);
            ysod_vec8<T>::load(w + (size_t)(r * k + q) * C + c0, wv);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(xv[e], wv[e], acc[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = ysod_act(acc[e], act);
    if (res) {
        float rv[8];
        ysod_vec8<T>::load(res + (size_t)pix * rcs + c0, rv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += rv[e];
    }
    ysod_vec8<T>::store(out + (size_t)pix * ocs + c0, acc);
}

// depthwise k x k, stride 1: a thread owns a horizontal strip of 4 output pixels x 8 channels and slides the filter window over
// the k x (k + 3) input patch it loads once (2x fewer 16 B loads than one pixel per thread for 3x3, 2.5x for 5x5).
template <typename T, int K>
__global__ void dwconv_strip_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                                    const T* __restrict__ res, T* __restrict__ out, int N, int H, int W, int C, int xcs, int ocs, int rcs,
                                    int act) {
    ysod_pdl_sync();
    constexpr int PAD = K / 2, SW = 4;
    const int c8n = C >> 3;
    const int wstrips = (W + SW - 1) / SW;
    const unsigned total = (unsigned)N * H * wstrips * c8n;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c0 = (int)(idx % c8n) * 8;
    unsigned r = idx / c8n;
    const int ws = (int)(r % wstrips); r /= wstrips;
    const int oh = (int)(r % H);
    const int n = (int)(r / H);
    const int ow0 = ws * SW;
    float acc[SW][8];
#pragma unroll
    for (int p = 0; p < SW; ++p)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[p][e] = bias[c0 + e];
#pragma unroll
    for (int kr = 0; kr < K; ++kr) {
        const int ih = oh + kr - PAD;
        if (ih < 0 || ih >= H) continue;
        float wv[K][8];
#pragma unroll
        for (int kq = 0; kq < K; ++kq) ysod_vec8<T>::load(w + (size_t)(kr * K + kq) * C + c0, wv[kq]);
        const T* xr = x + ((size_t)n * H + ih) * W * xcs + c0;
#pragma unroll
        for (int q = 0; q < SW + K - 1; ++q) {
            const int iw = ow0 + q - PAD;
            if (iw < 0 || iw >= W) continue;
            float xv[8];
            ysod_vec8<T>::load(xr + (size_t)iw * xcs, xv);
#pragma unroll
            for (int p = 0; p < SW; ++p) {
                const int kq = q - p;   // input column q feeds output p through tap kq
                if (kq >= 0 && kq < K) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[p][e] = fmaf(xv[e], wv[kq][e], acc[p][e]);
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < SW; ++p) {
        const int ow = ow0 + p;
        if (ow >= W) break;
        const size_t pix = ((size_t)n * H + oh) * W + ow;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[p][e] = ysod_act(acc[p][e], act);
        if (res) {
            float rv[8];
            ysod_vec8<T>::load(res + pix * rcs + c0, rv);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[p][e] += rv[e];
        }
        ysod_vec8<T>::store(out + pix * ocs + c0, acc[p]);
    }
}

// stem: NCHW fp32 image -> k x k stride-s conv (Cin = 3) -> NHWC; thread = (pixel, 8 output channels)
template <typename TO>
__global__ void stem_conv_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                                 TO* __restrict__ out, int N, int H, int W, int Ho, int Wo, int Cout, int ocs, int k, int s,
                                 int pad, int act) {
    ysod_pdl_sync();
    extern __shared__ float sw[];  // [Cout][k*k*3] + bias
    const int wn = Cout * k * k * 3;
    for (int i = threadIdx.x; i < wn; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[wn + i] = bias[i];
    __syncthreads();
    const int c8n = Cout >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)N * Ho * Wo * c8n;
    if (idx >= total) return;
    const int c0 = (int)(idx % c8n) * 8;
    const long long pix = idx / c8n;
    const int ow = (int)(pix % Wo);
    const int oh = (int)((pix / Wo) % Ho);
    const int n = (int)(pix / ((long long)Wo * Ho));
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = sw[wn + c0 + e];
    const int kk3 = k * k * 3;
    for (int r = 0; r < k; ++r) {
        const int ih = oh * s + r - pad;
        if (ih < 0 || ih >= H) continue;
        for (int q = 0; q < k; ++q) {
            const int iw = ow * s + q - pad;
            if (iw < 0 || iw >= W) continue;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float xv = __ldg(img + (((size_t)n * 3 + c) * H + ih) * W + iw);
                const float* wp = sw + (size_t)c0 * kk3 + (r * k + q) * 3 + c;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(xv, wp[e * kk3], acc[e]);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = ysod_act(acc[e], act);
    ysod_vec8<TO>::store(out + (size_t)pix * ocs + c0, acc);
}

// Stem fast path (3x3, stride 2, pad 1): a CTA computes a 32 x 8 tile of output pixels for CO output channels at a time.
// The 3 x 17 x 65 input patch is staged once in shared memory with coalesced row loads (this is also where NCHW becomes
// NHWC); each thread owns one pixel and CO accumulators, weights are read as float4 shared-memory broadcasts.
template <typename TO, int CO>
__global__ void __launch_bounds__(256)
stem_conv3x3s2_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                      TO* __restrict__ out, int H, int W, int Ho, int Wo, int Cout, int ocs, int act) {
    ysod_pdl_sync();
    constexpr int PH = 17, PW = 65, PWP = 66;
    __shared__ float patch[3 * PH * PWP];
    __shared__ __align__(16) float sw[27 * CO];
    __shared__ float sb[CO];
    const int n = blockIdx.z;
    const int oh0 = blockIdx.y * 8, ow0 = blockIdx.x * 32;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int ih0 = 2 * oh0 - 1, iw0 = 2 * ow0 - 1;
    for (int i = tid; i < 3 * PH * PW; i += 256) {
        const int c = i / (PH * PW);
        const int r = (i / PW) % PH;
        const int q = i % PW;
        const int ih = ih0 + r, iw = iw0 + q;
        float v = 0.f;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __ldg(img + (((size_t)n * 3 + c) * H + ih) * W + iw);
        patch[(c * PH + r) * PWP + q] = v;
    }
    const int oh = oh0 + threadIdx.y, ow = ow0 + threadIdx.x;
    const bool valid = oh < Ho && ow < Wo;
    for (int co0 = 0; co0 < Cout; co0 += CO) {
        __syncthreads();
        // weights arrive as [Cout][3][3][3] (co, r, s, c); stage transposed as [tap*3 + c][CO]
        for (int i = tid; i < 27 * CO; i += 256) {
            const int t = i / CO, co = i % CO;
            sw[i] = w[(size_t)(co0 + co) * 27 + t];
        }
        for (int i = tid; i < CO; i += 256) sb[i] = bias[co0 + i];
        __syncthreads();
        float acc[CO];
#pragma unroll
        for (int e = 0; e < CO; ++e) acc[e] = sb[e];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float xv = patch[(c * PH + 2 * threadIdx.y + r) * PWP + 2 * threadIdx.x + q];
                    const float4* wp = reinterpret_cast<const float4*>(&sw[((r * 3 + q) * 3 + c) * CO]);
#pragma unroll
                    for (int e = 0; e < CO / 4; ++e) {
                        const float4 w4 = wp[e];
                        acc[4 * e] = fmaf(xv, w4.x, acc[4 * e]);
                        acc[4 * e + 1] = fmaf(xv, w4.y, acc[4 * e + 1]);
                        acc[4 * e + 2] = fmaf(xv, w4.z, acc[4 * e + 2]);
                        acc[4 * e + 3] = fmaf(xv, w4.w, acc[4 * e + 3]);
                    }
                }
            }
        }
        if (valid) {
            TO* op = out + (((size_t)n * Ho + oh) * Wo + ow) * ocs + co0;
#pragma unroll
            for (int e = 0; e < CO; e += 8) {
                float o8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o8[j] = ysod_act(acc[e + j], act);
                ysod_vec8<TO>::store(op + e, o8);
            }
        }
    }
}

}  // namespace

extern "C" {

// Generic dense / grouped conv. dtype = element type of x, w, res; out_dtype may be fp32 for the raw head maps.
int ysod_conv_direct(const void* x, int dtype, int N, int H, int W, int Cin, int xcs, const void* w, const float* bias,
                     int Cout, int k, int s, int pad, int groups, void* out, int out_dtype, int ocs, const void* res, int rcs,
                     int act, cudaStream_t stream) {
    YSOD_CHECK_ARG(x && w && bias && out, "ysod_conv_direct: null pointer");
    YSOD_CHECK_ARG(groups >= 1 && Cin % groups == 0 && Cout % groups == 0, "ysod_conv_direct: bad groups %d", groups);
    YSOD_CHECK_ARG(groups == 1 || (Cout / groups) % 4 == 0, "ysod_conv_direct: Cout/groups must be a multiple of 4");
    const int Ho = (H + 2 * pad - k) / s + 1, Wo = (W + 2 * pad - k) / s + 1;
    const long long total = (long long)N * Ho * Wo * ((Cout + 3) / 4);
    const int blocks = ysod_cdiv(total, 256);
#define LAUNCH(TI, TO)                                                                                                    \
    ysod_launch(conv_direct_kernel<TI, TO>, blocks, 256, 0, stream, (const TI*)x, (const TI*)w, bias, (const TI*)res, (TO*)out, N, H, W, \
                                                           Cin, xcs, Ho, Wo, Cout, ocs, rcs, k, s, pad, groups, act)
    if (dtype == YSOD_F32 && out_dtype == YSOD_F32) LAUNCH(float, float);
    else if (dtype == YSOD_BF16 && out_dtype == YSOD_BF16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
    else if (dtype == YSOD_BF16 && out_dtype == YSOD_F32) LAUNCH(__nv_bfloat16, float);
    else {
        ysod_set_error("ysod_conv_direct: unsupported dtype combination %d -> %d", dtype, out_dtype);
        return YSOD_ERR_UNSUPPORTED;
    }
#undef LAUNCH
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_dwconv(const void* x, int dtype, int N, int H, int W, int C, int xcs, const void* w, const float* bias, int k, int s,
                int pad, void* out, int ocs, const void* res, int rcs, int act, cudaStream_t stream) {
    YSOD_CHECK_ARG(x && w && bias && out, "ysod_dwconv: null pointer");
    YSOD_CHECK_ARG(C % 8 == 0 && xcs % 8 == 0 && ocs % 8 == 0, "ysod_dwconv: channels must be a multiple of 8");
    const int Ho = (H + 2 * pad - k) / s + 1, Wo = (W + 2 * pad - k) / s + 1;
    const long long total = (long long)N * Ho * Wo * (C / 8);
    const int blocks = ysod_cdiv(total, 256);
    if (dtype == YSOD_BF16 && s == 1 && pad == k / 2 && (k == 3 || k == 5) && total < (1ll << 31)) {
        const long long strips = (long long)N * H * ((W + 3) / 4) * (C / 8);
        const int sb = ysod_cdiv(strips, 128);
        if (k == 3)
            ysod_launch(dwconv_strip_kernel<__nv_bfloat16, 3>, sb, 128, 0, stream, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias,
                        (const __nv_bfloat16*)res, (__nv_bfloat16*)out, N, H, W, C, xcs, ocs, rcs, act);
        else
            ysod_launch(dwconv_strip_kernel<__nv_bfloat16, 5>, sb, 128, 0, stream, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias,
                        (const __nv_bfloat16*)res, (__nv_bfloat16*)out, N, H, W, C, xcs, ocs, rcs, act);
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
    if (dtype == YSOD_F32)
        ysod_launch(dwconv_kernel<float>, blocks, 256, 0, stream, (const float*)x, (const float*)w, bias, (const float*)res, (float*)out, N,
                                                         H, W, C, xcs, Ho, Wo, ocs, rcs, k, s, pad, act);
    else
        ysod_launch(dwconv_kernel<__nv_bfloat16>, blocks, 256, 0, stream, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias,
                                                                 (const __nv_bfloat16*)res, (__nv_bfloat16*)out, N, H, W, C, xcs,
                                                                 Ho, Wo, ocs, rcs, k, s, pad, act);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

// img: (N,3,H,W) fp32 NCHW, exactly what the reference's DetectionModel.forward receives (tasks.py:129).
int ysod_stem_conv(const float* img, int N, int H, int W, const float* w, const float* bias, int Cout, int k, int s, int pad,
                   void* out, int out_dtype, int ocs, int act, cudaStream_t stream) {
    YSOD_CHECK_ARG(img && w && bias && out, "ysod_stem_conv: null pointer");
    YSOD_CHECK_ARG(Cout % 8 == 0 && ocs % 8 == 0, "ysod_stem_conv: Cout must be a multiple of 8");
    const int Ho = (H + 2 * pad - k) / s + 1, Wo = (W + 2 * pad - k) / s + 1;
    if (k == 3 && s == 2 && pad == 1 && Cout % 16 == 0) {
        dim3 grid(ysod_cdiv(Wo, 32), ysod_cdiv(Ho, 8), N), block(32, 8);
        if (Cout % 32 == 0) {
            if (out_dtype == YSOD_F32) ysod_launch(stem_conv3x3s2_kernel<float, 32>, grid, block, 0, stream, img, w, bias, (float*)out, H, W, Ho, Wo, Cout, ocs, act);
            else ysod_launch(stem_conv3x3s2_kernel<__nv_bfloat16, 32>, grid, block, 0, stream, img, w, bias, (__nv_bfloat16*)out, H, W, Ho, Wo, Cout, ocs, act);
        } else {
            if (out_dtype == YSOD_F32) ysod_launch(stem_conv3x3s2_kernel<float, 16>, grid, block, 0, stream, img, w, bias, (float*)out, H, W, Ho, Wo, Cout, ocs, act);
            else ysod_launch(stem_conv3x3s2_kernel<__nv_bfloat16, 16>, grid, block, 0, stream, img, w, bias, (__nv_bfloat16*)out, H, W, Ho, Wo, Cout, ocs, act);
        }
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
    const long long total = (long long)N * Ho * Wo * (Cout / 8);
    const int blocks = ysod_cdiv(total, 256);
    const size_t smem = (size_t)(Cout * k * k * 3 + Cout) * sizeof(float);
    YSOD_CHECK_ARG(smem <= 48 * 1024, "ysod_stem_conv: weights do not fit in shared memory");
    if (out_dtype == YSOD_F32)
        ysod_launch(stem_conv_kernel<float>, blocks, 256, smem, stream, img, w, bias, (float*)out, N, H, W, Ho, Wo, Cout, ocs, k, s, pad, act);
    else
        ysod_launch(stem_conv_kernel<__nv_bfloat16>, blocks, 256, smem, stream, img, w, bias, (__nv_bfloat16*)out, N, H, W, Ho, Wo, Cout, ocs,
                                                                       k, s, pad, act);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

}  // extern "C"
