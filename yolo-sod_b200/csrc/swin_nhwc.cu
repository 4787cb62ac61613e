// SwinBlock without token copies, for the levels the fully fused P2 kernel does not cover (the P4 block: 256 channels, 4 heads).
//
// Replaces (reference): ultralytics/nn/modules/blocks_transformer.py:133-171 SwinBlock.forward around its GEMMs --
//     dw 3x3 (:150) -> zero-pad to multiples of the window (:17-31) -> window_partition (:33-47) -> LayerNorm (:112) ... attention
//     (:116) ... window_reverse + crop (:49-79,125-129).
// LayerNorm, the in_proj / out_proj / MLP linears and the residual adds are per-token operations: they do not care in which order
// the tokens sit in memory. Only softmax(QK^T)V needs the window grouping. So the tokens stay in NHWC pixel order end to end:
//   * dwconv3_ln_kernel: depthwise 3x3 AND LayerNorm 1 in one pass (a warp owns a pixel's C channels), writing the raw tokens
//     (the residual) and the normalised tokens -- replaces ysod_dwconv + ysod_window_partition_ln;
//   * mha_window_nhwc_kernel: the window attention core reading q / k / v rows through the window -> pixel map and writing its output
//     rows back in pixel order -- no partition before, no ysod_window_reverse after. A window token that lies in the zero padding
//     (maps whose side is not a multiple of 7) was LayerNorm(0) = beta in the reference, i.e. its key / value are the constant vectors
//     in_proj_{k,v}(beta) + bias: they are precomputed by the host (kpad / vpad) and take part in the softmax unmasked, exactly as the
//     reference's padded tokens do; padded queries are never stored (the reference crops them).
// The linears then run on N*H*W real pixels instead of N*nW*49 padded tokens (-9 % at 40 x 40).
#include "common.cuh"

namespace {

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." YSOD_MMA_T "." YSOD_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ld32s(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t* r, const __nv_bfloat16* row) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t* r, const __nv_bfloat16* row) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ uint32_t pack2bf(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// Depthwise 3x3 (stride 1, pad 1, no bias) + LayerNorm over the channels. One warp owns a strip of 4 horizontally adjacent pixels,
// lane = V vectors of 8 channels (C = 256 * V): the 3 x 6 input pixels of a strip are loaded once and feed all four outputs (18 vector
// loads instead of 36), every tap's weights are read once per strip. The dw output is rounded to the storage type first and the
// LayerNorm runs on the rounded values, as when the two were separate kernels. w: [3][3][C] fp32.
template <int V>
__global__ void __launch_bounds__(256)
dwconv3_ln_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int xcs, const float* __restrict__ w, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y, int ycs, __nv_bfloat16* __restrict__ yn, int ncs) {
    ysod_pdl_sync();
    constexpr int C = 256 * V, SW = 4;
    extern __shared__ __align__(16) float sw[];   // [9][C] taps, then gamma[C], beta[C]
    for (int i = threadIdx.x; i < 9 * C; i += 256) sw[i] = w[i];
    for (int i = threadIdx.x; i < C; i += 256) { sw[9 * C + i] = gamma[i]; sw[10 * C + i] = beta[i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int spr = (W + SW - 1) / SW;                      // strips per row
    const long long nstrips = (long long)N * H * spr;
    for (long long st = (long long)blockIdx.x * 8 + warp; st < nstrips; st += (long long)gridDim.x * 8) {
        const int w0 = (int)(st % spr) * SW;
        const int ph = (int)((st / spr) % H);
        const long long n = st / ((long long)spr * H);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int c0 = (v * 32 + lane) * 8;
            // all 18 input pieces (3 rows x 6 pixels x 8 channels, 16 B each) are requested before the first one is used; the taps are then
            // accumulated in (r, s) order on fp32 pairs (fma.rn.f32x2: same rounding per lane, half the issue slots)
            uint4 xr[3][SW + 2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int ih = ph + r - 1;
#pragma unroll
                for (int q = 0; q < SW + 2; ++q) {
                    const int iw = w0 + q - 1;
                    xr[r][q] = make_uint4(0u, 0u, 0u, 0u);
                    if (ih >= 0 && ih < H && iw >= 0 && iw < W) xr[r][q] = *reinterpret_cast<const uint4*>(x + ((n * H + ih) * W + iw) * xcs + c0);
                }
            }
            unsigned long long acc2[SW][4];
#pragma unroll
            for (int j = 0; j < SW; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc2[j][e] = 0ull;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                if (ph + r - 1 < 0 || ph + r - 1 >= H) continue;   // (zero rows add nothing; skipping keeps the products identical)
                unsigned long long xin2[SW + 2][4];
#pragma unroll
                for (int q = 0; q < SW + 2; ++q) {
                    const uint32_t wv[4] = {xr[r][q].x, xr[r][q].y, xr[r][q].z, xr[r][q].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = ysod_unpack2(wv[e]);
                        asm("mov.b64 %0, {%1, %2};" : "=l"(xin2[q][e]) : "f"(f.x), "f"(f.y));
                    }
                }
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    unsigned long long wt2[4];
                    const float* wp = sw + (r * 3 + s) * C + c0;
                    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(wt2[0]), "=l"(wt2[1]) : "r"((uint32_t)__cvta_generic_to_shared(wp)) : "memory");
                    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(wt2[2]), "=l"(wt2[3]) : "r"((uint32_t)__cvta_generic_to_shared(wp + 4)) : "memory");
#pragma unroll
                    for (int j = 0; j < SW; ++j)
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[j][e]) : "l"(xin2[j + s][e]), "l"(wt2[e]));
                }
            }
            float acc[SW][8];
#pragma unroll
            for (int j = 0; j < SW; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[j][2 * e]), "=f"(acc[j][2 * e + 1]) : "l"(acc2[j][e]));
            // V == 1: the lane holds the pixel's whole share; V == 2 keeps the first half in registers until the second is done
            if (V == 1) {
                float su[SW], sq[SW], mean[SW], rstd[SW];
#pragma unroll
                for (int j = 0; j < SW; ++j) {
                    su[j] = 0.f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        acc[j][e] = __bfloat162float(__float2bfloat16_rn(acc[j][e]));   // the stored token value
                        su[j] += acc[j][e];
                    }
                }
#pragma unroll
                for (int j = 0; j < SW; ++j) mean[j] = ysod_warp_sum(su[j]) * (1.0f / (float)C);
#pragma unroll
                for (int j = 0; j < SW; ++j) {
                    sq[j] = 0.f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) { const float d = acc[j][e] - mean[j]; sq[j] = fmaf(d, d, sq[j]); }
                }
#pragma unroll
                for (int j = 0; j < SW; ++j) rstd[j] = rsqrtf(ysod_warp_sum(sq[j]) * (1.0f / (float)C) + eps);
                const float4 ga = *reinterpret_cast<const float4*>(sw + 9 * C + c0), gb = *reinterpret_cast<const float4*>(sw + 9 * C + c0 + 4);
                const float4 ba = *reinterpret_cast<const float4*>(sw + 10 * C + c0), bb = *reinterpret_cast<const float4*>(sw + 10 * C + c0 + 4);
                const float gm[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w}, bt[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int j = 0; j < SW; ++j) {
                    if (w0 + j >= W) continue;
                    const long long pix = (n * H + ph) * W + w0 + j;
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = (acc[j][e] - mean[j]) * rstd[j] * gm[e] + bt[e];
                    ysod_vec8<__nv_bfloat16>::store(y + pix * ycs + c0, acc[j]);
                    ysod_vec8<__nv_bfloat16>::store(yn + pix * ncs + c0, o);
                }
            } else {
                // C = 512: store the raw tokens per half; the LayerNorm runs below over both halves read back from y (L1 hits)
#pragma unroll
                for (int j = 0; j < SW; ++j)
                    if (w0 + j < W) ysod_vec8<__nv_bfloat16>::store(y + ((n * H + ph) * W + w0 + j) * ycs + c0, acc[j]);
            }
        }
        if (V > 1) {
            __syncwarp();
            for (int j = 0; j < SW && w0 + j < W; ++j) {
                const long long pix = (n * H + ph) * W + w0 + j;
                float tv[V][8];
                float su = 0.f;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    ysod_vec8<__nv_bfloat16>::load(y + pix * ycs + (v * 32 + lane) * 8, tv[v]);
#pragma unroll
                    for (int e = 0; e < 8; ++e) su += tv[v][e];
                }
                const float mean = ysod_warp_sum(su) * (1.0f / (float)C);
                float sq = 0.f;
#pragma unroll
                for (int v = 0; v < V; ++v)
#pragma unroll
                    for (int e = 0; e < 8; ++e) { const float d = tv[v][e] - mean; sq = fmaf(d, d, sq); }
                const float rstd = rsqrtf(ysod_warp_sum(sq) * (1.0f / (float)C) + eps);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int c0 = (v * 32 + lane) * 8;
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = (tv[v][e] - mean) * rstd * sw[9 * C + c0 + e] + sw[10 * C + c0 + e];
                    ysod_vec8<__nv_bfloat16>::store(yn + pix * ncs + c0, o);
                }
            }
        }
    }
}

// Window attention core on NHWC-ordered q / k / v rows (row = pixel, ld elements apart; head h = columns [h*D, (h+1)*D)).
// One CTA of 4 warps per (head, window); the window's 49 tokens (T = ws*ws <= 64) are MMA rows 0..48, rows T..63 are masked keys.
// Same arithmetic as attention.cu mha_win_kernel; only the row addressing differs.
template <int D>
__global__ void __launch_bounds__(128)
mha_window_nhwc_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v, int ld,
                       int H, int W, int ws, int nWh, int nWw, const __nv_bfloat16* __restrict__ kpad, const __nv_bfloat16* __restrict__ vpad,
                       float scale, __nv_bfloat16* __restrict__ out, int ldo, long long win0) {
    ysod_pdl_sync();
    constexpr int LDQ = D + 8;
    __shared__ __align__(16) __nv_bfloat16 Qs[64 * LDQ];
    __shared__ __align__(16) __nv_bfloat16 Ks[64 * LDQ];
    __shared__ __align__(16) __nv_bfloat16 Vs[64 * LDQ];
    __shared__ int rowpix[64];     // pixel row of token r (global pixel index), -1 = zero padding, -2 = MMA padding row
    const int h = blockIdx.x;
    const long long win = win0 + blockIdx.y;
    const int wj = (int)(win % nWw), wi = (int)((win / nWw) % nWh);
    const long long n = win / ((long long)nWw * nWh);
    const int T = ws * ws;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    if (tid < 64) {
        int rp = -2;
        if (tid < T) {
            const int ph = wi * ws + tid / ws, pw = wj * ws + tid % ws;
            rp = (ph < H && pw < W) ? (int)((n * H + ph) * W + pw) : -1;
        }
        rowpix[tid] = rp;
    }
    __syncthreads();
    {
        // every piece of the window's q / k / v rows is requested before the first shared-memory store (one latency, not NIT of them)
        constexpr int NIT = 64 * (D / 8) / 128;
        uint4 qv[NIT], kv[NIT], vv[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int i = tid + it * 128;
            const int r = i / (D / 8), pc = i - r * (D / 8);
            const int rp = rowpix[r];
            qv[it] = make_uint4(0, 0, 0, 0); kv[it] = qv[it]; vv[it] = qv[it];
            if (rp >= 0) {
                const size_t off = (size_t)rp * ld + h * D + pc * 8;
                qv[it] = *reinterpret_cast<const uint4*>(q + off);
                kv[it] = *reinterpret_cast<const uint4*>(k + off);
                vv[it] = *reinterpret_cast<const uint4*>(v + off);
            } else if (rp == -1) {   // a zero-padded window token: LayerNorm(0) = beta -> constant key / value rows
                kv[it] = *reinterpret_cast<const uint4*>(kpad + h * D + pc * 8);
                vv[it] = *reinterpret_cast<const uint4*>(vpad + h * D + pc * 8);
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int i = tid + it * 128;
            const int r = i / (D / 8), pc = i - r * (D / 8);
            *reinterpret_cast<uint4*>(&Qs[r * LDQ + pc * 8]) = qv[it];
            *reinterpret_cast<uint4*>(&Ks[r * LDQ + pc * 8]) = kv[it];
            *reinterpret_cast<uint4*>(&Vs[r * LDQ + pc * 8]) = vv[it];
        }
    }
    __syncthreads();
    const int row0 = warp * 16 + g;
    uint32_t qf[D / 16][4];
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
        qf[ks][0] = ld32s(&Qs[row0 * LDQ + ks * 16 + 2 * t]);
        qf[ks][1] = ld32s(&Qs[(row0 + 8) * LDQ + ks * 16 + 2 * t]);
        qf[ks][2] = ld32s(&Qs[row0 * LDQ + ks * 16 + 8 + 2 * t]);
        qf[ks][3] = ld32s(&Qs[(row0 + 8) * LDQ + ks * 16 + 8 + 2 * t]);
    }
    float s[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
#pragma unroll
        for (int k2 = 0; k2 < D / 32; ++k2) {
            uint32_t kb[4];
            ldsm_x4(kb, &Ks[(nb * 8 + (lane & 7)) * LDQ + k2 * 32 + (lane >> 3) * 8]);
            mma16816(s[nb], qf[2 * k2], kb[0], kb[1]);
            mma16816(s[nb], qf[2 * k2 + 1], kb[2], kb[3]);
        }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        const int key = nb * 8 + 2 * t;
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nb][e] *= scale;
        if (key >= T) { s[nb][0] = -INFINITY; s[nb][2] = -INFINITY; }
        if (key + 1 >= T) { s[nb][1] = -INFINITY; s[nb][3] = -INFINITY; }
        mx0 = fmaxf(mx0, fmaxf(s[nb][0], s[nb][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nb][2], s[nb][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        s[nb][0] = __expf(s[nb][0] - mx0); s[nb][1] = __expf(s[nb][1] - mx0);
        s[nb][2] = __expf(s[nb][2] - mx1); s[nb][3] = __expf(s[nb][3] - mx1);
        l0 += s[nb][0] + s[nb][1];
        l1 += s[nb][2] + s[nb][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    uint32_t pf[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pf[i][0] = pack2bf(s[2 * i][0], s[2 * i][1]);
        pf[i][1] = pack2bf(s[2 * i][2], s[2 * i][3]);
        pf[i][2] = pack2bf(s[2 * i + 1][0], s[2 * i + 1][1]);
        pf[i][3] = pack2bf(s[2 * i + 1][2], s[2 * i + 1][3]);
    }
    __syncthreads();   // every warp has read its Q rows: Qs becomes the output staging tile
#pragma unroll
    for (int nb = 0; nb < D / 8; ++nb) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i2 = 0; i2 < 2; ++i2) {
            uint32_t vb[4];
            ldsm_x4_trans(vb, &Vs[(i2 * 32 + lane) * LDQ + nb * 8]);
            mma16816(o, pf[2 * i2], vb[0], vb[1]);
            mma16816(o, pf[2 * i2 + 1], vb[2], vb[3]);
        }
        *reinterpret_cast<uint32_t*>(&Qs[row0 * LDQ + nb * 8 + 2 * t]) = pack2bf(o[0] * inv0, o[1] * inv0);
        *reinterpret_cast<uint32_t*>(&Qs[(row0 + 8) * LDQ + nb * 8 + 2 * t]) = pack2bf(o[2] * inv1, o[3] * inv1);
    }
    __syncthreads();
    for (int i = tid; i < T * (D / 8); i += 128) {
        const int r = i / (D / 8), pc = i - r * (D / 8);
        const int rp = rowpix[r];
        if (rp >= 0) *reinterpret_cast<uint4*>(out + (size_t)rp * ldo + h * D + pc * 8) = *reinterpret_cast<const uint4*>(&Qs[r * LDQ + pc * 8]);
    }
}

}  // namespace

extern "C" {

// y = dw3x3(x) (no bias), yn = LayerNorm(y) over the channels, both NHWC 16-bit views; w: [3][3][C] fp32. C in {256, 512}.
int ysod_dwconv3_ln(const void* x, int N, int H, int W, int C, int xcs, const float* w, const float* gamma, const float* beta, float eps,
                    void* y, int ycs, void* yn, int ncs, cudaStream_t stream) {
    YSOD_CHECK_ARG(x && w && gamma && beta && y && yn, "ysod_dwconv3_ln: null pointer");
    YSOD_CHECK_ARG(xcs % 8 == 0 && ycs % 8 == 0 && ncs % 8 == 0, "ysod_dwconv3_ln: pixel strides must be multiples of 8");
    if (C != 256 && C != 512) {
        ysod_set_error("ysod_dwconv3_ln: C %d unsupported (256, 512)", C);
        return YSOD_ERR_UNSUPPORTED;
    }
    const long long nstrips = (long long)N * H * ((W + 3) / 4);
    int sms = 148, dev = 0;
    YSOD_CUDA(cudaGetDevice(&dev));
    YSOD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long blocks = (nstrips + 7) / 8;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    const size_t smem = (size_t)11 * C * sizeof(float);
    if (C == 256) {
        ysod_launch(dwconv3_ln_kernel<1>, (unsigned)blocks, 256, smem, stream, (const __nv_bfloat16*)x, N, H, W, xcs, w, gamma, beta, eps,
                    (__nv_bfloat16*)y, ycs, (__nv_bfloat16*)yn, ncs);
    } else {
        ysod_launch(dwconv3_ln_kernel<2>, (unsigned)blocks, 256, smem, stream, (const __nv_bfloat16*)x, N, H, W, xcs, w, gamma, beta, eps,
                    (__nv_bfloat16*)y, ycs, (__nv_bfloat16*)yn, ncs);
    }
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

// softmax(q k^T * scale) v per (window, head) with q / k / v / out rows in NHWC pixel order (row stride ld / ldo elements, head h =
// columns [h*D, (h+1)*D)); windows of ws x ws tokens (ws*ws <= 64) tile the N x H x W map from its origin, nWh x nWw per image;
// kpad / vpad: [heads*D] key / value of a zero-padded token. D in {32, 64}.
int ysod_mha_window_nhwc(const void* q, const void* k, const void* v, int ld, int N, int H, int W, int ws, int heads, int D, const void* kpad,
                         const void* vpad, float scale, void* out, int ldo, cudaStream_t stream) {
    YSOD_CHECK_ARG(q && k && v && kpad && vpad && out, "ysod_mha_window_nhwc: null pointer");
    YSOD_CHECK_ARG(ws >= 1 && ws * ws <= 64 && ld % 8 == 0 && ldo % 8 == 0 && heads >= 1 && heads <= 65535, "ysod_mha_window_nhwc: bad geometry");
    YSOD_CHECK_ARG(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)out % 16) == 0 &&
                   ((uintptr_t)kpad % 16) == 0 && ((uintptr_t)vpad % 16) == 0, "ysod_mha_window_nhwc: 16 B alignment");
    YSOD_CHECK_ARG((long long)N * H * W < (1ll << 31), "ysod_mha_window_nhwc: map too large");
    if (D != 32 && D != 64) {
        ysod_set_error("ysod_mha_window_nhwc: head_dim %d unsupported (32, 64)", D);
        return YSOD_ERR_UNSUPPORTED;
    }
    const int nWh = ysod_cdiv(H, ws), nWw = ysod_cdiv(W, ws);
    const long long nwin = (long long)N * nWh * nWw;
    for (long long w0 = 0; w0 < nwin; w0 += 65535) {
        const unsigned nb = (unsigned)(nwin - w0 < 65535 ? nwin - w0 : 65535);
        dim3 grid(heads, nb);
        if (D == 32)
            ysod_launch(mha_window_nhwc_kernel<32>, grid, 128, 0, stream, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ld, H, W,
                        ws, nWh, nWw, (const __nv_bfloat16*)kpad, (const __nv_bfloat16*)vpad, scale, (__nv_bfloat16*)out, ldo, w0);
        else
            ysod_launch(mha_window_nhwc_kernel<64>, grid, 128, 0, stream, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ld, H, W,
                        ws, nWh, nWw, (const __nv_bfloat16*)kpad, (const __nv_bfloat16*)vpad, scale, (__nv_bfloat16*)out, ldo, w0);
        YSOD_LAUNCH_CHECK();
    }
    return YSOD_OK;
}

}  // extern "C"
