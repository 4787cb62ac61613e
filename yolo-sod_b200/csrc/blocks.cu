// Memory-bound blocks of the MAFN neck: SE, CBAM (channel + spatial), CoordAtt, SPPF pooling, nearest upsample /
// slice copy, LayerNorm, Swin window partition / reverse, A2 adaptive pooling + bilinear row upsample.
//
// Replaces (reference, ~5-12 ATen kernels each):
//   SE            ultralytics/nn/modules/smallobj_modules.py:57-92
//   CBAM_Block    ultralytics/nn/modules/cbam_block.py:8-55
//   CA_Block      ultralytics/nn/modules/ca_block.py:16-59
//   SPPF pooling  ultralytics/nn/modules/block.py:178-197  (three chained 5x5/s1 max-pools == 5x5, 9x9, 13x13 windows)
//   Upsample/Concat  nn.Upsample(nearest,2) + conv.py:323-334 Concat (producers write channel slices instead)
//   window_partition / window_reverse / LayerNorm   blocks_transformer.py:8-79,98-131
//   A2_Attn pooling + bilinear  a2_attn.py:44-60
// All are HBM-bound: NHWC, 8 channels (16 B of bf16) per thread access, fp32 math, warp-shuffle / shared-memory
// reductions, deterministic two-stage global reductions (no float atomics).
#include "common.cuh"
#include <stdlib.h>

namespace {

// ------------------------------------------------------------------------------------------------------------
// global average / max pool, stage 1: partial[n][s][c] over a pixel range
template <typename T>
__device__ __forceinline__ void gap_partial_body(const T* __restrict__ x, int HW, int C, int xcs, int S, float* __restrict__ psum,
                                                 float* __restrict__ pmax, float* sm) {
    const int n = blockIdx.y, s = blockIdx.x;
    const int c8n = C >> 3;
    const int PL = blockDim.x / c8n;
    const int cg = threadIdx.x % c8n, pl = threadIdx.x / c8n;
    const int chunk = (HW + S - 1) / S;
    const int p0 = s * chunk, p1 = min(HW, p0 + chunk);
    float su[8], mx[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { su[e] = 0.f; mx[e] = -INFINITY; }
    if (pl < PL) {
        const T* xb = x + (size_t)n * HW * xcs + cg * 8;
        int p = p0 + pl;
        for (; p + 7 * PL < p1; p += 8 * PL) {   // eight independent 16 B loads in flight per thread (a pure read stream needs the depth)
            float v[8][8];
#pragma unroll
            for (int u = 0; u < 8; ++u) ysod_vec8<T>::load(xb + (size_t)(p + u * PL) * xcs, v[u]);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                su[e] += ((v[0][e] + v[1][e]) + (v[2][e] + v[3][e])) + ((v[4][e] + v[5][e]) + (v[6][e] + v[7][e]));
                mx[e] = fmaxf(mx[e], fmaxf(fmaxf(fmaxf(v[0][e], v[1][e]), fmaxf(v[2][e], v[3][e])), fmaxf(fmaxf(v[4][e], v[5][e]), fmaxf(v[6][e], v[7][e]))));
            }
        }
        for (; p + 3 * PL < p1; p += 4 * PL) {   // four independent 16 B loads in flight per thread
            float v0[8], v1[8], v2[8], v3[8];
            ysod_vec8<T>::load(xb + (size_t)p * xcs, v0);
            ysod_vec8<T>::load(xb + (size_t)(p + PL) * xcs, v1);
            ysod_vec8<T>::load(xb + (size_t)(p + 2 * PL) * xcs, v2);
            ysod_vec8<T>::load(xb + (size_t)(p + 3 * PL) * xcs, v3);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                su[e] += (v0[e] + v1[e]) + (v2[e] + v3[e]);
                mx[e] = fmaxf(fmaxf(mx[e], fmaxf(v0[e], v1[e])), fmaxf(v2[e], v3[e]));
            }
        }
        for (; p < p1; p += PL) {
            float v[8];
            ysod_vec8<T>::load(xb + (size_t)p * xcs, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) { su[e] += v[e]; mx[e] = fmaxf(mx[e], v[e]); }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            sm[pl * C + cg * 8 + e] = su[e];
            sm[(PL + pl) * C + cg * 8 + e] = mx[e];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, m = -INFINITY;
        for (int q = 0; q < PL; ++q) { a += sm[q * C + c]; m = fmaxf(m, sm[(PL + q) * C + c]); }
        psum[((size_t)n * S + s) * C + c] = a;
        if (pmax) pmax[((size_t)n * S + s) * C + c] = m;
    }
}
template <typename T>
__global__ void gap_partial_kernel(const T* __restrict__ x, int HW, int C, int xcs, int S, float* __restrict__ psum,
                                   float* __restrict__ pmax) {
    ysod_pdl_sync();
    extern __shared__ float sm[];  // [PL][C] sums, then [PL][C] maxes
    gap_partial_body<T>(x, HW, C, xcs, S, psum, pmax, sm);
}

// SE gate: mean -> fc1(+bias) -> ReLU -> fc2(+bias) -> sigmoid for image n, by one CTA. fp32 as in the reference (:81-91).
// sm: C + hid floats. `ld` reads the partials through L2 (they may have been written by other SMs of the same launch).
__device__ __forceinline__ void se_gate_body(int n, const float* psum, int S, int HW, int C, const float* __restrict__ w1,
                                             const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                             int hid, float* __restrict__ gate, float* sm) {
    float* mean = sm;
    float* h = sm + C;
    // many partials (the per-tile sums a producing kernel wrote): blockDim / C threads per channel, each a strided share with four loads in flight
    {
        // parts = threads per channel (fixed by blockDim and C, so the summation order is fixed)
        const int parts = (S > 32 && C <= 128 && blockDim.x >= 256) ? 256 / C : 1;
        if (parts == 1) {
            for (int c = threadIdx.x; c < C; c += blockDim.x) {
                float a = 0.f;
                for (int s = 0; s < S; ++s) a += __ldcg(psum + ((size_t)n * S + s) * C + c);
                mean[c] = a / (float)HW;
            }
        } else {
            __shared__ float red[256];
            const int c = threadIdx.x % C, part = threadIdx.x / C;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            if (part < parts) {
                int s = part;
                for (; s + 3 * parts < S; s += 4 * parts) {
                    a0 += __ldcg(psum + ((size_t)n * S + s) * C + c);
                    a1 += __ldcg(psum + ((size_t)n * S + s + parts) * C + c);
                    a2 += __ldcg(psum + ((size_t)n * S + s + 2 * parts) * C + c);
                    a3 += __ldcg(psum + ((size_t)n * S + s + 3 * parts) * C + c);
                }
                for (; s < S; s += parts) a0 += __ldcg(psum + ((size_t)n * S + s) * C + c);
            }
            if ((int)threadIdx.x < parts * C) red[threadIdx.x] = (a0 + a1) + (a2 + a3);
            __syncthreads();
            if ((int)threadIdx.x < C) {
                float a = 0.f;
                for (int q = 0; q < parts; ++q) a += red[q * C + threadIdx.x];
                mean[threadIdx.x] = a / (float)HW;
            }
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = warp; j < hid; j += nw) {
        float a = 0.f;
        for (int c = lane; c < C; c += 32) a = fmaf(w1[(size_t)j * C + c], mean[c], a);
        a = ysod_warp_sum(a);
        if (lane == 0) h[j] = fmaxf(a + b1[j], 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = b2[c];
        for (int j = 0; j < hid; ++j) a = fmaf(w2[(size_t)c * hid + j], h[j], a);
        gate[(size_t)n * C + c] = ysod_sigmoid(a);
    }
}
__global__ void se_gate_kernel(const float* __restrict__ psum, int S, int HW, int C, const float* __restrict__ w1,
                               const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2, int hid,
                               float* __restrict__ gate) {
    ysod_pdl_sync();
    extern __shared__ float sm[];  // mean[C], h[hid]
    se_gate_body(blockIdx.x, psum, S, HW, C, w1, b1, w2, b2, hid, gate, sm);
}

// CBAM channel gate: sigmoid(fc(avg) + fc(max)), fc = conv1x1(no bias) -> ReLU -> conv1x1(no bias) (cbam_block.py:14-23), image n,
// one CTA. sm: 2C + 2 hid floats.
__device__ __forceinline__ void cbam_gate_body(int n, const float* psum, const float* pmax, int S, int HW, int C,
                                               const float* __restrict__ w1, const float* __restrict__ w2, int hid,
                                               float* __restrict__ gate, float* sm) {
    float* avg = sm;
    float* mx = sm + C;
    float* ha = sm + 2 * C;
    float* hm = ha + hid;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, m = -INFINITY;
        for (int s = 0; s < S; ++s) {
            a += __ldcg(psum + ((size_t)n * S + s) * C + c);
            m = fmaxf(m, __ldcg(pmax + ((size_t)n * S + s) * C + c));
        }
        avg[c] = a / (float)HW;
        mx[c] = m;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = warp; j < hid; j += nw) {
        float a = 0.f, m = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float w = w1[(size_t)j * C + c];
            a = fmaf(w, avg[c], a);
            m = fmaf(w, mx[c], m);
        }
        a = ysod_warp_sum(a);
        m = ysod_warp_sum(m);
        if (lane == 0) { ha[j] = fmaxf(a, 0.f); hm[j] = fmaxf(m, 0.f); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, m = 0.f;
        for (int j = 0; j < hid; ++j) {
            const float w = w2[(size_t)c * hid + j];
            a = fmaf(w, ha[j], a);
            m = fmaf(w, hm[j], m);
        }
        gate[(size_t)n * C + c] = ysod_sigmoid(a + m);
    }
}
__global__ void cbam_gate_kernel(const float* __restrict__ psum, const float* __restrict__ pmax, int S, int HW, int C,
                                 const float* __restrict__ w1, const float* __restrict__ w2, int hid, float* __restrict__ gate) {
    ysod_pdl_sync();
    extern __shared__ float sm[];  // avg[C], mx[C], ha[hid], hm[hid]
    cbam_gate_body(blockIdx.x, psum, pmax, S, HW, C, w1, w2, hid, gate, sm);
}

// Pool + gate in one launch: every CTA writes its partial sums / maxes, takes a ticket on a per-image counter, and the CTA that
// draws the last ticket of its image runs the gate MLP (kind 0 = SE, 1 = CBAM channel attention) and re-arms the counter. Saves the
// separate one-CTA-per-image gate launch (a ~5-8 us dependent launch on a path that is launch-bound at batch 1).
struct GateArgs {
    unsigned* counter;     // [N], zero before the first launch; left at zero by every launch
    int kind, hid;
    const float *w1, *b1, *w2, *b2;
    float* gate;
};
template <typename T>
__global__ void gap_gate_kernel(const T* __restrict__ x, int HW, int C, int xcs, int S, float* __restrict__ psum, float* __restrict__ pmax,
                                GateArgs g) {
    ysod_pdl_sync();
    extern __shared__ float sm[];
    __shared__ int is_last;
    gap_partial_body<T>(x, HW, C, xcs, S, psum, pmax, sm);
    __threadfence();                  // this thread's partials are visible device-wide before the ticket is taken
    __syncthreads();
    const int n = blockIdx.y;
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(g.counter + n, 1u);
        is_last = (t == (unsigned)S - 1u);
        if (is_last) g.counter[n] = 0u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (g.kind == 0) se_gate_body(n, psum, S, HW, C, g.w1, g.b1, g.w2, g.b2, g.hid, g.gate, sm);
    else cbam_gate_body(n, psum, pmax, S, HW, C, g.w1, g.w2, g.hid, g.gate, sm);
}

// Streaming-kernel indexing: a thread keeps one fixed 8-channel group (cg) and walks pixels pix0, pix0 + pstep, ... with 32-bit
// arithmetic (the grid size is a multiple of C/8 threads), so the loops contain no 64-bit divisions.
struct StreamIdx {
    unsigned cg, pix0, pstep;
    __device__ __forceinline__ StreamIdx(int c8n) {
        const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
        cg = gtid % (unsigned)c8n;
        pix0 = gtid / (unsigned)c8n;
        pstep = (gridDim.x * blockDim.x) / (unsigned)c8n;
    }
};

// out = x * gate[n][c]. Grid-stride, four independent 16 B loads in flight per thread.
template <typename T>
__global__ void scale_channels_kernel(const T* __restrict__ x, int HW, int C, int xcs, const float* __restrict__ gate,
                                      T* __restrict__ out, int ocs, unsigned npix) {
    ysod_pdl_sync();
    const StreamIdx si(C >> 3);
    for (unsigned p0 = si.pix0; p0 < npix; p0 += 4 * si.pstep) {
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned pix = p0 + u * si.pstep;
            if (pix < npix) ysod_vec8<T>::load(x + (size_t)pix * xcs + si.cg * 8, v[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned pix = p0 + u * si.pstep;
            if (pix < npix) {
                const unsigned n = pix / (unsigned)HW;
                const float4 g0 = *reinterpret_cast<const float4*>(gate + (size_t)n * C + si.cg * 8);
                const float4 g1 = *reinterpret_cast<const float4*>(gate + (size_t)n * C + si.cg * 8 + 4);
                v[u][0] *= g0.x; v[u][1] *= g0.y; v[u][2] *= g0.z; v[u][3] *= g0.w;
                v[u][4] *= g1.x; v[u][5] *= g1.y; v[u][6] *= g1.z; v[u][7] *= g1.w;
                ysod_vec8<T>::store(out + (size_t)pix * ocs + si.cg * 8, v[u]);
            }
        }
    }
}

// CBAM spatial statistics of x*gate: stats[pix] = (mean_c, max_c). A group of G lanes (G = min(C/8,32)) per pixel.
template <typename T>
__global__ void cbam_stats_kernel(const T* __restrict__ x, int HW, int C, int xcs, const float* __restrict__ gate,
                                  float2* __restrict__ stats, long long npix) {
    ysod_pdl_sync();
    const int c8n = C >> 3;
    const int G = c8n < 32 ? c8n : 32;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, gl = lane % G;
    const int ppw = 32 / G;  // pixels per warp
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long pix = warp_id * ppw + sub;
    float su = 0.f, mx = -INFINITY;
    if (pix < npix) {
        const int n = (int)(pix / HW);
        for (int cg = gl; cg < c8n; cg += G) {
            float v[8];
            ysod_vec8<T>::load(x + (size_t)pix * xcs + cg * 8, v);
            const float* g = gate + (size_t)n * C + cg * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float t = v[e] * g[e];
                su += t;
                mx = fmaxf(mx, t);
            }
        }
    }
    for (int o = G >> 1; o > 0; o >>= 1) {
        su += __shfl_xor_sync(0xffffffffu, su, o);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (pix < npix && gl == 0) stats[pix] = make_float2(su / (float)C, mx);
}

// CBAM apply: sa = sigmoid(conv7x7([mean,max])) ; out = x * gate * sa. Same lane grouping; taps split across the group.
template <typename T>
__global__ void cbam_apply_kernel(const T* __restrict__ x, int H, int W, int C, int xcs, const float* __restrict__ gate,
                                  const float2* __restrict__ stats, const float* __restrict__ wsp, int ks, T* __restrict__ out,
                                  int ocs, long long npix) {
    ysod_pdl_sync();
    extern __shared__ float swsp[];  // [2][ks][ks]
    for (int i = threadIdx.x; i < 2 * ks * ks; i += blockDim.x) swsp[i] = wsp[i];
    __syncthreads();
    const int c8n = C >> 3;
    const int G = c8n < 32 ? c8n : 32;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, gl = lane % G;
    const int ppw = 32 / G;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long pix = warp_id * ppw + sub;
    const int HW = H * W;
    float a = 0.f;
    int n = 0;
    if (pix < npix) {
        n = (int)(pix / HW);
        const int rem = (int)(pix % HW);
        const int h = rem / W, w = rem % W;
        const int pad = ks / 2;
        for (int t = gl; t < ks * ks; t += G) {
            const int r = t / ks, q = t % ks;
            const int ih = h + r - pad, iw = w + q - pad;
            if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
                const float2 st = stats[(size_t)n * HW + ih * W + iw];
                a = fmaf(swsp[t], st.x, a);
                a = fmaf(swsp[ks * ks + t], st.y, a);
            }
        }
    }
    for (int o = G >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (pix < npix) {
        const float sa = ysod_sigmoid(a);
        for (int cg = gl; cg < c8n; cg += G) {
            float v[8];
            ysod_vec8<T>::load(x + (size_t)pix * xcs + cg * 8, v);
            const float* g = gate + (size_t)n * C + cg * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = v[e] * g[e] * sa;
            ysod_vec8<T>::store(out + (size_t)pix * ocs + cg * 8, v);
        }
    }
}

// CBAM spatial statistics, streaming variant for C <= 256 (G = C/8 lanes own one pixel, one 16 B load each): grid-stride with
// four pixels in flight per lane group; the trip count is uniform across the grid so the shuffles stay warp-converged.
template <typename T>
__global__ void cbam_stats_stream_kernel(const T* __restrict__ x, int HW, int C, int xcs, const float* __restrict__ gate,
                                         float2* __restrict__ stats, long long npix, int iters) {
    ysod_pdl_sync();
    const int G = C >> 3;
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const long long ngroups = ((long long)gridDim.x * blockDim.x) / G;
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    for (int it = 0; it < iters; ++it) {
        const long long pix0 = gid + (long long)it * 4 * ngroups;
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long pix = pix0 + u * ngroups;
            if (pix < npix) ysod_vec8<T>::load(x + (size_t)pix * xcs + gl * 8, v[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long pix = pix0 + u * ngroups;
            float su = 0.f, mx = -INFINITY;
            if (pix < npix) {
                const float* g = gate + (size_t)(pix / HW) * C + gl * 8;
                const float4 g0 = *reinterpret_cast<const float4*>(g), g1 = *reinterpret_cast<const float4*>(g + 4);
                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float tv = v[u][e] * gg[e];
                    su += tv;
                    mx = fmaxf(mx, tv);
                }
            }
            for (int o = G >> 1; o > 0; o >>= 1) {
                su += __shfl_xor_sync(0xffffffffu, su, o);
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            if (pix < npix && gl == 0) stats[pix] = make_float2(su / (float)C, mx);
        }
    }
}

// CBAM spatial statistics, per-image variant for C <= 256: grid (pixel chunks, images). G = C/8 lanes own one pixel (one 16 B
// load each, four pixels in flight per lane group); the lane's eight gate values live in registers for the whole CTA, so the
// inner loop has no division and no gate loads. The trip count is uniform per CTA so the shuffles stay warp-converged.
template <typename T>
__global__ void __launch_bounds__(256)
cbam_stats_img_kernel(const T* __restrict__ x, int HW, int C, int xcs, const float* __restrict__ gate, float2* __restrict__ stats,
                      int chunk, int iters) {
    ysod_pdl_sync();
    const int G = C >> 3;
    const int lg = __ffs(G) - 1;
    const int n = blockIdx.y;
    const int gl = threadIdx.x & (G - 1), grp = threadIdx.x >> lg, ngr = blockDim.x >> lg;
    const int p0 = blockIdx.x * chunk, p1 = min(HW, p0 + chunk);
    float gg[8];
    {
        const float4 g0 = *reinterpret_cast<const float4*>(gate + (size_t)n * C + gl * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(gate + (size_t)n * C + gl * 8 + 4);
        gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w; gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
    }
    const T* xb = x + (size_t)n * HW * xcs + gl * 8;
    float2* sb = stats + (size_t)n * HW;
    const float inv_c = 1.0f / (float)C;
    for (int it = 0; it < iters; ++it) {
        const int pb = p0 + grp + it * 4 * ngr;
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int pix = pb + u * ngr;
            if (pix < p1) ysod_vec8<T>::load(xb + (size_t)pix * xcs, v[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int pix = pb + u * ngr;
            float su = 0.f, mx = -INFINITY;
            if (pix < p1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float tv = v[u][e] * gg[e];
                    su += tv;
                    mx = fmaxf(mx, tv);
                }
            }
            for (int o = G >> 1; o > 0; o >>= 1) {
                su += __shfl_xor_sync(0xffffffffu, su, o);
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            if (pix < p1 && gl == 0) sb[pix] = make_float2(su * inv_c, mx);
        }
    }
}

// Per-thread asynchronous prefetch ring (cp.async / LDGSTS): a thread requests its next DEPTH 16-byte pieces straight into its own
// shared-memory slots ([slot][thread]: conflict-free) without holding registers for them, and consumes the oldest. Nobody else reads a
// thread's slots, so the ring needs no barrier -- only cp.async.wait_group. This multiplies the bytes a streaming kernel keeps in
// flight (DEPTH x 16 B x threads) at constant register count.
__device__ __forceinline__ void ysod_cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ysod_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ysod_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// cbam_stats_img_kernel with the prefetch ring: identical arithmetic (same per-lane sums, same shuffle tree), DEPTH pixels in flight
// per lane group instead of 4 register-held ones.
template <typename T, int DEPTH>
__global__ void __launch_bounds__(256)
cbam_stats_ring_kernel(const T* __restrict__ x, int HW, int C, int xcs, const float* __restrict__ gate, float2* __restrict__ stats,
                       int chunk, int iters) {
    ysod_pdl_sync();
    extern __shared__ __align__(16) uint8_t ring_raw[];       // [DEPTH][256] x 16 B (bf16) -- T = float uses two pieces per slot
    constexpr int PIECES = sizeof(T) == 2 ? 1 : 2;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(ring_raw);
    const int G = C >> 3;
    const int lg = __ffs(G) - 1;
    const int n = blockIdx.y;
    const int tid = threadIdx.x;
    const int gl = tid & (G - 1), grp = tid >> lg, ngr = blockDim.x >> lg;
    const int p0 = blockIdx.x * chunk, p1 = min(HW, p0 + chunk);
    float gg[8];
    {
        const float4 g0 = *reinterpret_cast<const float4*>(gate + (size_t)n * C + gl * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(gate + (size_t)n * C + gl * 8 + 4);
        gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w; gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
    }
    const T* xb = x + (size_t)n * HW * xcs + gl * 8;
    float2* sb = stats + (size_t)n * HW;
    const float inv_c = 1.0f / (float)C;
    const int total = iters * 4;                 // pixel slots of this lane group: pixel j = p0 + grp + j * ngr (uniform trip count)
    auto request = [&](int j) {
        const int pix = p0 + grp + j * ngr;
        const bool ok = j < total && pix < p1;
        const uint32_t dst = ring + (uint32_t)(((j % DEPTH) * 256 + tid) * 16 * PIECES);
        const T* src = xb + (size_t)(ok ? pix : p0) * xcs;
#pragma unroll
        for (int q = 0; q < PIECES; ++q) ysod_cp_async16(dst + 16 * q, reinterpret_cast<const char*>(src) + 16 * q, ok);
        ysod_cp_async_commit();
    };
#pragma unroll
    for (int j = 0; j < DEPTH - 1; ++j) request(j);
    for (int j = 0; j < total; ++j) {
        request(j + DEPTH - 1);
        ysod_cp_async_wait<DEPTH - 1>();         // the group of slot j has landed
        const int pix = p0 + grp + j * ngr;
        float v[8];
        ysod_vec8<T>::load(reinterpret_cast<const T*>(ring_raw + (size_t)(((j % DEPTH) * 256 + tid) * 16 * PIECES)), v);
        float su = 0.f, mx = -INFINITY;
        if (pix < p1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float tv = v[e] * gg[e];
                su += tv;
                mx = fmaxf(mx, tv);
            }
        }
        for (int o = G >> 1; o > 0; o >>= 1) {
            su += __shfl_xor_sync(0xffffffffu, su, o);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (pix < p1 && gl == 0) sb[pix] = make_float2(su * inv_c, mx);
    }
}

// CBAM apply: sa = sigmoid(conv7x7([mean,max])) ; out = x * gate * sa. One CTA = a 16 x 16 pixel tile of one image: the
// (16+6)^2 statistics halo is staged in shared memory, each thread evaluates the 7x7x2 filter for one pixel, then the CTA
// streams the tile's channels with four 16 B loads in flight per thread.
template <typename T>
__global__ void __launch_bounds__(256)
cbam_apply_tile_kernel(const T* __restrict__ x, int H, int W, int C, int xcs, const float* __restrict__ gate,
                       const float2* __restrict__ stats, const float* __restrict__ wsp, T* __restrict__ out, int ocs) {
    ysod_pdl_sync();
    constexpr int TS = 16, HALO = 3, PS = TS + 2 * HALO;
    __shared__ float2 st[PS * PS];
    __shared__ float swsp[2 * 49];
    __shared__ float sa_s[TS * TS];
    const int n = blockIdx.z, h0 = blockIdx.y * TS, w0 = blockIdx.x * TS;
    const int tid = threadIdx.x;
    {   // request the tile's activations now: they travel to L2 while the statistics halo is staged and the 7x7 filter runs
        const int lpp = (C * (int)sizeof(T) + 127) / 128;           // 128 B lines per pixel
        for (int i = tid; i < TS * TS * lpp; i += 256) {
            const int pl = i / lpp, l = i - pl * lpp;
            const int h = h0 + (pl >> 4), w = w0 + (pl & 15);
            if (h < H && w < W) ysod_prefetch_l2(reinterpret_cast<const char*>(x + (((size_t)n * H + h) * W + w) * xcs) + l * 128);
        }
    }
    if (tid < 98) swsp[tid] = wsp[tid];
    for (int i = tid; i < PS * PS; i += 256) {
        const int ih = h0 - HALO + i / PS, iw = w0 - HALO + i % PS;
        st[i] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? stats[((size_t)n * H + ih) * W + iw] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    {
        const int ph = tid / TS, pw = tid % TS;
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                const float2 v = st[(ph + r) * PS + pw + q];
                a = fmaf(swsp[r * 7 + q], v.x, a);
                a = fmaf(swsp[49 + r * 7 + q], v.y, a);
            }
        sa_s[tid] = ysod_sigmoid(a);
    }
    __syncthreads();
    // C/8 is a power of two <= 256 (host check): a thread keeps one channel group for the whole tile, its eight gate values in
    // registers; pixel / channel-group indices are shifts
    const int c8n = C >> 3;
    const int lg = __ffs(c8n) - 1;
    const int total = TS * TS * c8n;
    const int cg = tid & (c8n - 1);
    float gg[8];
    {
        const float* gn = gate + (size_t)n * C + cg * 8;
        const float4 g0 = *reinterpret_cast<const float4*>(gn), g1 = *reinterpret_cast<const float4*>(gn + 4);
        gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w; gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
    }
    const T* xn = x + (size_t)n * H * W * xcs + cg * 8;
    T* on = out + (size_t)n * H * W * ocs + cg * 8;
    for (int base = tid; base < total; base += 4 * 256) {
        float v[4][8];
        bool ok[4];
        int pl[4], gp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + u * 256;
            pl[u] = i >> lg;
            const int h = h0 + (pl[u] >> 4), w = w0 + (pl[u] & 15);
            gp[u] = h * W + w;
            ok[u] = (i < total) && h < H && w < W;
            if (ok[u]) ysod_vec8<T>::load(xn + (size_t)gp[u] * xcs, v[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!ok[u]) continue;
            const float sa = sa_s[pl[u]];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[u][e] *= gg[e] * sa;
            ysod_vec8<T>::store(on + (size_t)gp[u] * ocs, v[u]);
        }
    }
}

// CBAM spatial attention in ONE pass over the map (cbam_block.py:25-55): a CTA owns a 32 x 16 pixel tile of one image. It computes the
// channel statistics (mean_c, max_c of x * gate) of the tile plus its 3-pixel halo straight from x (38 x 22 pixels; the halo re-reads
// are L2 hits: neighbouring tiles run at the same time), evaluates sigmoid(conv7x7([mean, max])) for the tile from shared memory and
// streams x * gate * sa out. With the pooling pass that feeds the channel gate this makes CBAM 2 reads + 1 write of the map (the floor:
// the gate needs the whole map first) instead of 3 + 1, and the statistics map never exists in HBM. The per-pixel arithmetic is that
// of cbam_stats_img_kernel + cbam_apply_tile_kernel, expression for expression, so the result is bit-identical to the two-pass path.
template <typename T>
__global__ void __launch_bounds__(256, 3)
cbam_spatial_fused_kernel(const T* __restrict__ x, int H, int W, int C, int xcs, const float* __restrict__ gate,
                          const float* __restrict__ wsp, T* __restrict__ out, int ocs) {
    ysod_pdl_sync();
    constexpr int TW_ = 32, TH_ = 16, HALO = 3, PW_ = TW_ + 2 * HALO, PH_ = TH_ + 2 * HALO, NP = PW_ * PH_;   // 38 x 22 = 836
    __shared__ float2 st[NP];
    __shared__ float swsp[2 * 49];
    __shared__ float sa_s[TW_ * TH_];
    const int n = blockIdx.z, h0 = blockIdx.y * TH_, w0 = blockIdx.x * TW_;
    const int tid = threadIdx.x;
    if (tid < 98) swsp[tid] = wsp[tid];
    const int c8n = C >> 3;                      // power of two <= 32 (host check)
    const int lg = __ffs(c8n) - 1;
    const int gl = tid & (c8n - 1), grp = tid >> lg, ngr = 256 >> lg;
    float gg[8];
    {
        const float* gn = gate + (size_t)n * C + gl * 8;
        const float4 g0 = *reinterpret_cast<const float4*>(gn), g1 = *reinterpret_cast<const float4*>(gn + 4);
        gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w; gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
    }
    const T* xn = x + (size_t)n * H * W * xcs + gl * 8;
    const float inv_c = 1.0f / (float)C;
    // ---- statistics of the tile + halo: a group of c8n lanes per pixel, four pixels in flight per group (uniform trip count: the
    //      shuffles stay converged); pixels outside the image are the zero padding of the 7x7 conv
    const int iters = (NP + 4 * ngr - 1) / (4 * ngr);
    for (int it = 0; it < iters; ++it) {
        const int pb = grp + it * 4 * ngr;
        float v[4][8];
        bool in[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = pb + u * ngr;
            const int ih = h0 - HALO + i / PW_, iw = w0 - HALO + i % PW_;
            in[u] = i < NP && ih >= 0 && ih < H && iw >= 0 && iw < W;
            if (in[u]) ysod_vec8<T>::load(xn + ((size_t)ih * W + iw) * xcs, v[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = pb + u * ngr;
            float su = 0.f, mx = -INFINITY;
            if (in[u]) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float tv = v[u][e] * gg[e];
                    su += tv;
                    mx = fmaxf(mx, tv);
                }
            }
            for (int o = c8n >> 1; o > 0; o >>= 1) {
                su += __shfl_xor_sync(0xffffffffu, su, o);
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            if (i < NP && gl == 0) st[i] = in[u] ? make_float2(su * inv_c, mx) : make_float2(0.f, 0.f);
        }
    }
    __syncthreads();
    // ---- sa = sigmoid(conv7x7([mean, max])): two pixels per thread
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
        const int pl = tid + k * 256;
        const int ph = pl >> 5, pw = pl & 31;
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                const float2 sv = st[(ph + r) * PW_ + pw + q];
                a = fmaf(swsp[r * 7 + q], sv.x, a);
                a = fmaf(swsp[49 + r * 7 + q], sv.y, a);
            }
        sa_s[pl] = ysod_sigmoid(a);
    }
    __syncthreads();
    // ---- out = x * gate * sa for the tile (x was read a moment ago by this CTA: L1 / L2 hits), four 16 B loads in flight
    T* on = out + (size_t)n * H * W * ocs + gl * 8;
    const int total = TW_ * TH_ * c8n;
    for (int base = tid; base < total; base += 4 * 256) {
        float v[4][8];
        bool ok[4];
        int pl[4], gp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + u * 256;
            pl[u] = i >> lg;
            const int h = h0 + (pl[u] >> 5), w = w0 + (pl[u] & 31);
            gp[u] = h * W + w;
            ok[u] = (i < total) && h < H && w < W;
            if (ok[u]) ysod_vec8<T>::load(xn + (size_t)gp[u] * xcs, v[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!ok[u]) continue;
            const float sa = sa_s[pl[u]];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[u][e] *= gg[e] * sa;
            ysod_vec8<T>::store(on + (size_t)gp[u] * ocs, v[u]);
        }
    }
}

// CoordAtt strip pools in ONE pass over the map (ca_block.py:42-45): a CTA owns 16 rows of one image; thread = (8-channel group,
// column lane). Every pixel is read once and added to its row sum (reduced over the column lanes through shared memory, fixed
// order) and to its column partial sum (kept in registers over the CTA's 16 rows, written as part[n][row block][w][c]); a second
// tiny launch adds the row blocks in fixed order. Deterministic (no float atomics). The former two-pass kernel ran 16 threads per
// CTA with one load in flight and read the map twice (25 % of HBM peak).
template <typename T, int WL>
__global__ void __launch_bounds__(256)
ca_pool_rows_kernel(const T* __restrict__ x, int H, int W, int C, int xcs, float* __restrict__ pooled, float* __restrict__ part, int RB) {
    ysod_pdl_sync();
    constexpr int ROWS = 16, MAXW = 8;           // a column lane owns columns wl, wl + WL, ... (at most MAXW of them: W <= WL * MAXW)
    extern __shared__ float sm[];                // [WL][C] row partials
    const int n = blockIdx.y, rb = blockIdx.x;
    const int c8n = C >> 3;                      // 256 / WL channel groups
    const int cg = threadIdx.x % c8n, wl = threadIdx.x / c8n;
    const int r0 = rb * ROWS, r1 = min(H, r0 + ROWS);
    float col[MAXW][8];
#pragma unroll
    for (int j = 0; j < MAXW; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) col[j][e] = 0.f;
    const T* xb = x + (size_t)n * H * W * xcs + cg * 8;
    const float inv_w = 1.0f / (float)W;
    for (int r = r0; r < r1; ++r) {
        float v[MAXW][8];
#pragma unroll
        for (int j = 0; j < MAXW; ++j) {          // all of a row's loads of this thread are in flight together
            const int w = wl + j * WL;
            if (w < W) ysod_vec8<T>::load(xb + ((size_t)r * W + w) * xcs, v[j]);
        }
        float rs[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) rs[e] = 0.f;
#pragma unroll
        for (int j = 0; j < MAXW; ++j) {
            const int w = wl + j * WL;
            if (w < W) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { rs[e] += v[j][e]; col[j][e] += v[j][e]; }
            }
        }
        __syncthreads();                          // the previous row's partials have been consumed
#pragma unroll
        for (int e = 0; e < 8; ++e) sm[wl * C + cg * 8 + e] = rs[e];
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += 256) {
            float a = 0.f;
            for (int q = 0; q < WL; ++q) a += sm[q * C + c];
            pooled[((size_t)n * (H + W) + r) * C + c] = a * inv_w;
        }
    }
#pragma unroll
    for (int j = 0; j < MAXW; ++j) {
        const int w = wl + j * WL;
        if (w < W) ysod_vec8<float>::store(part + (((size_t)n * RB + rb) * W + w) * C + cg * 8, col[j]);
    }
}
// pooled[n][H + w][c] = (sum over row blocks of part[n][rb][w][c]) / H
__global__ void ca_pool_cols_kernel(const float* __restrict__ part, int H, int W, int C, int RB, float* __restrict__ pooled, int total) {
    ysod_pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = i % C, w = (i / C) % W, n = i / (C * W);
    float a = 0.f;
    for (int rb = 0; rb < RB; ++rb) a += part[(((size_t)n * RB + rb) * W + w) * C + c];
    pooled[((size_t)n * (H + W) + H + w) * C + c] = a / (float)H;
}

// CoordAtt strip pools -> pooled[n][H + W][C] fp32: rows [0,H) = mean over w, rows [H,H+W) = mean over h (ca_block.py:42-45)
template <typename T>
__global__ void ca_pool_kernel(const T* __restrict__ x, int H, int W, int C, int xcs, float* __restrict__ pooled) {
    ysod_pdl_sync();
    const int n = blockIdx.y;
    const int row = blockIdx.x;  // 0..H+W-1
    const int c8n = C >> 3;
    for (int cg = threadIdx.x; cg < c8n; cg += blockDim.x) {
        float a[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = 0.f;
        if (row < H) {
            for (int w = 0; w < W; ++w) {
                float v[8];
                ysod_vec8<T>::load(x + (((size_t)n * H + row) * W + w) * xcs + cg * 8, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] += v[e];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] /= (float)W;
        } else {
            const int w = row - H;
            for (int h = 0; h < H; ++h) {
                float v[8];
                ysod_vec8<T>::load(x + (((size_t)n * H + h) * W + w) * xcs + cg * 8, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] += v[e];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] /= (float)H;
        }
        ysod_vec8<float>::store(pooled + ((size_t)n * (H + W) + row) * C + cg * 8, a);
    }
}

// CoordAtt gates: y = hsigmoid(conv1'(pooled)) (BN folded), a = sigmoid(conv_{h|w}(y)) (ca_block.py:47-57)
__global__ void ca_gate_kernel(const float* __restrict__ pooled, int H, int W, int C, int mip, const float* __restrict__ w1,
                               const float* __restrict__ b1, const float* __restrict__ wh, const float* __restrict__ bh,
                               const float* __restrict__ ww, const float* __restrict__ bw, float* __restrict__ att) {
    ysod_pdl_sync();
    extern __shared__ float sm[];  // y[mip]
    const int n = blockIdx.y, row = blockIdx.x;
    const float* p = pooled + ((size_t)n * (H + W) + row) * C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = warp; j < mip; j += nw) {
        float a = 0.f;
        for (int c = lane; c < C; c += 32) a = fmaf(w1[(size_t)j * C + c], p[c], a);
        a = ysod_warp_sum(a);
        if (lane == 0) sm[j] = ysod_act(a + b1[j], YSOD_ACT_HSIGMOID);
    }
    __syncthreads();
    const float* wo = row < H ? wh : ww;
    const float* bo = row < H ? bh : bw;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = bo[c];
        for (int j = 0; j < mip; ++j) a = fmaf(wo[(size_t)c * mip + j], sm[j], a);
        att[((size_t)n * (H + W) + row) * C + c] = ysod_sigmoid(a);
    }
}

// out = x * a_w[n][w][c] * a_h[n][h][c]  (ca_block.py:57). Grid-stride, four independent 16 B loads in flight per thread.
template <typename T>
__global__ void ca_apply_kernel(const T* __restrict__ x, int H, int W, int C, int xcs, const float* __restrict__ att,
                                T* __restrict__ out, int ocs, unsigned npix) {
    ysod_pdl_sync();
    const StreamIdx si(C >> 3);
    for (unsigned p0 = si.pix0; p0 < npix; p0 += 4 * si.pstep) {
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned pix = p0 + u * si.pstep;
            if (pix < npix) ysod_vec8<T>::load(x + (size_t)pix * xcs + si.cg * 8, v[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned pix = p0 + u * si.pstep;
            if (pix < npix) {
                const unsigned row = pix / (unsigned)W, w = pix - row * (unsigned)W;
                const unsigned n = row / (unsigned)H, h = row - n * (unsigned)H;
                float ah[8], aw[8];
                ysod_vec8<float>::load(att + ((size_t)n * (H + W) + h) * C + si.cg * 8, ah);
                ysod_vec8<float>::load(att + ((size_t)n * (H + W) + H + w) * C + si.cg * 8, aw);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[u][e] = v[u][e] * aw[e] * ah[e];
                ysod_vec8<T>::store(out + (size_t)pix * ocs + si.cg * 8, v[u]);
            }
        }
    }
}

// SPPF: o1 = max 5x5, o2 = max 9x9, o3 = max 13x13 of y0 (== three chained 5x5/s1/p2 max pools)
template <typename T>
__global__ void sppf_pool_kernel(const T* __restrict__ y0, int H, int W, int C, int xcs, int k, T* __restrict__ o1,
                                 T* __restrict__ o2, T* __restrict__ o3, int ocs, long long total) {
    ysod_pdl_sync();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c8n = C >> 3;
    const int cg = (int)(idx % c8n);
    const long long pix = idx / c8n;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    const int r1 = k / 2, r2 = 2 * r1, r3 = 3 * r1;
    float m1[8], m2[8], m3[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m1[e] = m2[e] = m3[e] = -INFINITY;
    for (int dh = -r3; dh <= r3; ++dh) {
        const int ih = h + dh;
        if (ih < 0 || ih >= H) continue;
        for (int dw = -r3; dw <= r3; ++dw) {
            const int iw = w + dw;
            if (iw < 0 || iw >= W) continue;
            float v[8];
            ysod_vec8<T>::load(y0 + (((size_t)n * H + ih) * W + iw) * xcs + cg * 8, v);
            const int ad = max(abs(dh), abs(dw));
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                m3[e] = fmaxf(m3[e], v[e]);
                if (ad <= r2) m2[e] = fmaxf(m2[e], v[e]);
                if (ad <= r1) m1[e] = fmaxf(m1[e], v[e]);
            }
        }
    }
    ysod_vec8<T>::store(o1 + (size_t)pix * ocs + cg * 8, m1);
    ysod_vec8<T>::store(o2 + (size_t)pix * ocs + cg * 8, m2);
    ysod_vec8<T>::store(o3 + (size_t)pix * ocs + cg * 8, m3);
}

// SPPF for small maps (H*W <= 1024, i.e. P5 up to 1024^2 inputs): one CTA per (image, 8-channel group) keeps the plane in
// shared memory and applies the separable 5x5 max three times (row pass, column pass), writing o1, o2, o3.
template <typename T>
__global__ void __launch_bounds__(256)
sppf_plane_kernel(const T* __restrict__ y0, int H, int W, int xcs, int k, T* __restrict__ o1, T* __restrict__ o2,
                  T* __restrict__ o3, int ocs) {
    ysod_pdl_sync();
    extern __shared__ float sppf_sm[];
    const int n = blockIdx.y, cg = blockIdx.x;
    const int HW = H * W, r = k / 2;
    float* a = sppf_sm;
    float* b = sppf_sm + HW * 8;
    for (int p = threadIdx.x; p < HW; p += blockDim.x) {
        float v[8];
        ysod_vec8<T>::load(y0 + ((size_t)n * HW + p) * xcs + cg * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[p * 8 + e] = v[e];
    }
    __syncthreads();
    T* outs[3] = {o1, o2, o3};
    for (int pass = 0; pass < 3; ++pass) {
        for (int i = threadIdx.x; i < HW * 8; i += blockDim.x) {  // row max: a -> b
            const int e = i & 7, p = i >> 3, h = p / W, w = p % W;
            float m = -INFINITY;
            for (int d = -r; d <= r; ++d) {
                const int ww = w + d;
                if (ww >= 0 && ww < W) m = fmaxf(m, a[(h * W + ww) * 8 + e]);
            }
            b[i] = m;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < HW * 8; i += blockDim.x) {  // column max: b -> a
            const int e = i & 7, p = i >> 3, h = p / W, w = p % W;
            float m = -INFINITY;
            for (int d = -r; d <= r; ++d) {
                const int hh = h + d;
                if (hh >= 0 && hh < H) m = fmaxf(m, b[(hh * W + w) * 8 + e]);
            }
            a[i] = m;
        }
        __syncthreads();
        for (int p = threadIdx.x; p < HW; p += blockDim.x) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = a[p * 8 + e];
            ysod_vec8<T>::store(outs[pass] + ((size_t)n * HW + p) * ocs + cg * 8, v);
        }
    }
}

// bf16 variant of the plane kernel: one 1024-thread CTA per (image, 16-channel group); every item is one pixel x 8 channels
// held as a uint4 of four bf16x2 (max is exact in bf16, so the result equals the fp32-staged kernel bit for bit). A thread
// owns ITEMS fixed items, so pixel coordinates are computed once; the three chained 5x5 pools are six separable passes of packed
// __hmax2 over shared memory. (Kept small on purpose: a fully unrolled 8-items-per-thread version stalled on instruction fetch.)
__device__ __forceinline__ uint4 sppf_max4(uint4 a, uint4 b) {
    uint4 r;
    const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
    __nv_bfloat162* z = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) z[i] = __hmax2(x[i], y[i]);
    return r;
}
template <int ITEMS>
__global__ void __launch_bounds__(1024)
sppf_plane_bf16_kernel(const __nv_bfloat16* __restrict__ y0, int H, int W, int xcs, int k, __nv_bfloat16* __restrict__ o1,
                       __nv_bfloat16* __restrict__ o2, __nv_bfloat16* __restrict__ o3, int ocs) {
    ysod_pdl_sync();
    extern __shared__ uint4 sppf_q[];
    const int n = blockIdx.y, c0 = blockIdx.x * 16;
    const int HW = H * W, r = k / 2, items = HW * 2;
    uint4* a = sppf_q;
    uint4* b = sppf_q + items;
    int ph[ITEMS], pw[ITEMS];   // item = pixel * 2 + half
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const int it = threadIdx.x + j * 1024;
        const int p = it >> 1;
        ph[j] = p / W;
        pw[j] = p - ph[j] * W;
        if (it < items) a[it] = *reinterpret_cast<const uint4*>(y0 + ((size_t)n * HW + p) * xcs + c0 + (it & 1) * 8);
    }
    __syncthreads();
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        __nv_bfloat16* outp = pass == 0 ? o1 : (pass == 1 ? o2 : o3);
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {   // row max: a -> b
            const int it = threadIdx.x + j * 1024;
            if (it < items) {
                uint4 m = a[it];
                for (int d = 1; d <= r; ++d) {
                    if (pw[j] - d >= 0) m = sppf_max4(m, a[it - 2 * d]);
                    if (pw[j] + d < W) m = sppf_max4(m, a[it + 2 * d]);
                }
                b[it] = m;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {   // column max: b -> a, and out
            const int it = threadIdx.x + j * 1024;
            if (it < items) {
                uint4 m = b[it];
                for (int d = 1; d <= r; ++d) {
                    if (ph[j] - d >= 0) m = sppf_max4(m, b[it - 2 * d * W]);
                    if (ph[j] + d < H) m = sppf_max4(m, b[it + 2 * d * W]);
                }
                a[it] = m;
                *reinterpret_cast<uint4*>(outp + ((size_t)n * HW + (it >> 1)) * ocs + c0 + (it & 1) * 8) = m;
            }
        }
        __syncthreads();
    }
}

// ---- MambaBlock (GLU fallback) plumbing: blocks_mamba.py:84-103, 167-236 -------------------------------------------------------
// F.avg_pool2d(y, r, r): out (N, H/r, W/r, C), fp32 accumulation, divide by r*r. Thread = 8 channels of one output pixel.
template <typename T>
__global__ void avgpool_kernel(const T* __restrict__ x, int H, int W, int C, int xcs, int r, T* __restrict__ out, int ocs, long long total) {
    ysod_pdl_sync();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c8n = C >> 3, Ho = H / r, Wo = W / r;
    const int cg = (int)(idx % c8n);
    const long long pix = idx / c8n;
    const int ow = (int)(pix % Wo), oh = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int dh = 0; dh < r; ++dh)
        for (int dw = 0; dw < r; ++dw) {
            float v[8];
            ysod_vec8<T>::load(x + (((size_t)n * H + oh * r + dh) * W + ow * r + dw) * xcs + cg * 8, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += v[e];
        }
    const float inv = 1.0f / (float)(r * r);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= inv;
    ysod_vec8<T>::store(out + (size_t)pix * ocs + cg * 8, acc);
}

// GLU gate: out[p][c] = sigmoid(x[p][hid + c]) * x[p][c], c < hid  (`a, g = pw1(x).chunk(2, 1); sigmoid(g) * a`)
template <typename T>
__global__ void glu_kernel(const T* __restrict__ x, int hid, int xcs, T* __restrict__ out, int ocs, long long total) {
    ysod_pdl_sync();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c8n = hid >> 3;
    const int cg = (int)(idx % c8n);
    const long long pix = idx / c8n;
    float a[8], g[8];
    ysod_vec8<T>::load(x + (size_t)pix * xcs + cg * 8, a);
    ysod_vec8<T>::load(x + (size_t)pix * xcs + hid + cg * 8, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] *= ysod_sigmoid(g[e]);
    ysod_vec8<T>::store(out + (size_t)pix * ocs + cg * 8, a);
}

// out = res + F.interpolate(y, size=(H, W), mode="nearest"): source index = min(floor(dst * in / out), in - 1) as ATen computes it
template <typename T>
__global__ void upsample_add_kernel(const T* __restrict__ y, int Hh, int Wh, int C, int ycs, const T* __restrict__ res, int rcs, int H, int W,
                                    T* __restrict__ out, int ocs, long long total) {
    ysod_pdl_sync();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c8n = C >> 3;
    const int cg = (int)(idx % c8n);
    const long long pix = idx / c8n;
    const int w = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((long long)W * H));
    const int sh = min((int)floorf((float)h * ((float)Hh / (float)H)), Hh - 1);
    const int sw = min((int)floorf((float)w * ((float)Wh / (float)W)), Wh - 1);
    float a[8], b[8];
    ysod_vec8<T>::load(y + (((size_t)n * Hh + sh) * Wh + sw) * ycs + cg * 8, a);
    ysod_vec8<T>::load(res + (size_t)pix * rcs + cg * 8, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += b[e];
    ysod_vec8<T>::store(out + (size_t)pix * ocs + cg * 8, a);
}

// nearest-neighbour upsample by `scale` (1 = plain slice copy) into a channel slice. Grid-stride, four loads in flight.
template <typename T>
__global__ void upsample_copy_kernel(const T* __restrict__ x, int H, int W, int C, int xcs, int scale, T* __restrict__ out,
                                     int ocs, unsigned npix) {
    ysod_pdl_sync();
    const StreamIdx si(C >> 3);
    const unsigned Wo = W * scale, Ho = H * scale;
    for (unsigned p0 = si.pix0; p0 < npix; p0 += 4 * si.pstep) {
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned pix = p0 + u * si.pstep;
            if (pix < npix) {
                const unsigned row = pix / Wo, ow = pix - row * Wo;
                const unsigned n = row / Ho, oh = row - n * Ho;
                ysod_vec8<T>::load(x + (((size_t)n * H + oh / scale) * W + ow / scale) * xcs + si.cg * 8, v[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned pix = p0 + u * si.pstep;
            if (pix < npix) ysod_vec8<T>::store(out + (size_t)pix * ocs + si.cg * 8, v[u]);
        }
    }
}

// LayerNorm over the last dim (eps inside sqrt, biased variance, as torch.nn.LayerNorm). A group of G lanes (power of two,
// G = min(32, C/8 rounded up)) owns one row, so narrow rows (C = 64 -> 8 lanes) still fill the warp: 32/G rows per warp.
template <typename T, bool GATHER>
__global__ void layernorm_kernel(const T* __restrict__ x, long long rows, int C, int ldx, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, T* __restrict__ out, int ldo, int G,
                                 // window-partition gather (GATHER): x is NHWC, rows are window tokens
                                 int H, int W, int wh, int ww, int nWh, int nWw, T* __restrict__ raw_out) {
    ysod_pdl_sync();
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long row = warp_id * (32 / G) + lane / G;
    const bool valid = row < rows;
    const int c8n = C >> 3;
    const T* src = nullptr;
    if (valid) {
        if (GATHER) {
            const int tpw = wh * ww;
            const int t = (int)(row % tpw);
            const long long win = row / tpw;
            const int wj = (int)(win % nWw);
            const int wi = (int)((win / nWw) % nWh);
            const int n = (int)(win / ((long long)nWw * nWh));
            const int h = wi * wh + t / ww, w = wj * ww + t % ww;
            if (h < H && w < W) src = x + (((size_t)n * H + h) * W + w) * ldx;  // else: zero-padded token (blocks_transformer.py:31-36)
        } else {
            src = x + (size_t)row * ldx;
        }
    }
    constexpr int MAXG = 8;  // C <= 8 * 8 * G
    float v[MAXG][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXG; ++i) {
        const int cg = gl + G * i;
        if (cg < c8n) {
            if (src) ysod_vec8<T>::load(src + cg * 8, v[i]);
            else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) s += v[i][e];
        }
    }
    for (int o = G >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXG; ++i) {
        const int cg = gl + G * i;
        if (cg < c8n) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float d = v[i][e] - mean; q += d * d; }
        }
    }
    for (int o = G >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)C + eps);
    if (!valid) return;
#pragma unroll
    for (int i = 0; i < MAXG; ++i) {
        const int cg = gl + G * i;
        if (cg < c8n) {
            if (GATHER && raw_out) ysod_vec8<T>::store(raw_out + (size_t)row * ldo + cg * 8, v[i]);
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * gamma[cg * 8 + e] + beta[cg * 8 + e];
            ysod_vec8<T>::store(out + (size_t)row * ldo + cg * 8, o);
        }
    }
}

// LayerNorm, streaming variant for C <= 256 (G = C/8 lanes own one row, one 16 B load each): grid-stride with four rows in
// flight per lane group and a grid-uniform trip count (the shuffles stay warp-converged). GATHER = window-partition gather.
template <typename T, bool GATHER>
__global__ void layernorm_stream_kernel(const T* __restrict__ x, long long rows, int C, int ldx, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, float eps, T* __restrict__ out, int ldo, int iters,
                                        int H, int W, int wh, int ww, int nWh, int nWw, T* __restrict__ raw_out) {
    ysod_pdl_sync();
    const int G = C >> 3;
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const long long ngroups = ((long long)gridDim.x * blockDim.x) / G;
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    float gm[8], bt[8];
    ysod_vec8<float>::load(gamma + gl * 8, gm);
    ysod_vec8<float>::load(beta + gl * 8, bt);
    const float invC = 1.0f / (float)C;
    for (int it = 0; it < iters; ++it) {
        const long long row0 = gid + (long long)it * 4 * ngroups;
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long row = row0 + u * ngroups;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[u][e] = 0.f;
            if (row < rows) {
                const T* src;
                if (GATHER) {
                    const int tpw = wh * ww;
                    const int t = (int)(row % tpw);
                    const long long win = row / tpw;
                    const int wj = (int)(win % nWw);
                    const int wi = (int)((win / nWw) % nWh);
                    const int n = (int)(win / ((long long)nWw * nWh));
                    const int h = wi * wh + t / ww, w = wj * ww + t % ww;
                    src = (h < H && w < W) ? x + (((size_t)n * H + h) * W + w) * ldx : nullptr;   // else zero-padded token
                } else {
                    src = x + (size_t)row * ldx;
                }
                if (src) ysod_vec8<T>::load(src + gl * 8, v[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long row = row0 + u * ngroups;
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) s += v[u][e];
            for (int o = G >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s * invC;
            float q = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float d = v[u][e] - mean; q += d * d; }
            for (int o = G >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            const float rstd = rsqrtf(q * invC + eps);
            if (row < rows) {
                if (GATHER && raw_out) ysod_vec8<T>::store(raw_out + (size_t)row * ldo + gl * 8, v[u]);
                float o8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o8[e] = (v[u][e] - mean) * rstd * gm[e] + bt[e];
                ysod_vec8<T>::store(out + (size_t)row * ldo + gl * 8, o8);
            }
        }
    }
}

// window tokens -> NHWC (crop the padding) (blocks_transformer.py:49-79,125-129). Grid-stride, four loads in flight.
template <typename T>
__global__ void window_reverse_kernel(const T* __restrict__ tok, int ldt, int H, int W, int C, int wh, int ww, int nWh, int nWw,
                                      T* __restrict__ out, int ocs, unsigned npix) {
    ysod_pdl_sync();
    const StreamIdx si(C >> 3);
    for (unsigned p0 = si.pix0; p0 < npix; p0 += 4 * si.pstep) {
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned pix = p0 + u * si.pstep;
            if (pix < npix) {
                const unsigned r = pix / (unsigned)W, w = pix - r * (unsigned)W;
                const unsigned n = r / (unsigned)H, h = r - n * (unsigned)H;
                const unsigned win = (n * nWh + h / wh) * nWw + w / ww;
                const size_t row = (size_t)win * (wh * ww) + (h % wh) * ww + (w % ww);
                ysod_vec8<T>::load(tok + row * ldt + si.cg * 8, v[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned pix = p0 + u * si.pstep;
            if (pix < npix) ysod_vec8<T>::store(out + (size_t)pix * ocs + si.cg * 8, v[u]);
        }
    }
}

// adaptive_avg_pool2d over H only: (H, W) -> (OH, W); bin i = [floor(i*H/OH), ceil((i+1)*H/OH))
template <typename T>
__global__ void adaptive_pool_rows_kernel(const T* __restrict__ x, int H, int W, int C, int xcs, int OH, T* __restrict__ out,
                                          int ocs, long long total) {
    ysod_pdl_sync();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c8n = C >> 3;
    const int cg = (int)(idx % c8n);
    const long long pix = idx / c8n;
    const int w = (int)(pix % W);
    const int oh = (int)((pix / W) % OH);
    const int n = (int)(pix / ((long long)W * OH));
    const int h0 = (oh * H) / OH, h1 = ((oh + 1) * H + OH - 1) / OH;
    float a[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = 0.f;
    for (int h = h0; h < h1; ++h) {
        float v[8];
        ysod_vec8<T>::load(x + (((size_t)n * H + h) * W + w) * xcs + cg * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += v[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] /= (float)(h1 - h0);
    ysod_vec8<T>::store(out + (size_t)pix * ocs + cg * 8, a);
}

// bilinear (align_corners=False) upsample over H only: (IH, W) -> (OH, W) (a2_attn.py:60; width scale is 1 = identity)
template <typename T>
__global__ void bilinear_rows_kernel(const T* __restrict__ x, int IH, int W, int C, int xcs, int OH, T* __restrict__ out, int ocs,
                                     long long total) {
    ysod_pdl_sync();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c8n = C >> 3;
    const int cg = (int)(idx % c8n);
    const long long pix = idx / c8n;
    const int w = (int)(pix % W);
    const int oh = (int)((pix / W) % OH);
    const int n = (int)(pix / ((long long)W * OH));
    const float scale = (float)IH / (float)OH;
    float src = scale * ((float)oh + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    const int i0 = (int)src;
    const int i1 = i0 + ((i0 < IH - 1) ? 1 : 0);
    const float l1 = src - (float)i0, l0 = 1.0f - l1;
    float a[8], b[8];
    ysod_vec8<T>::load(x + (((size_t)n * IH + i0) * W + w) * xcs + cg * 8, a);
    ysod_vec8<T>::load(x + (((size_t)n * IH + i1) * W + w) * xcs + cg * 8, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = l0 * a[e] + l1 * b[e];
    ysod_vec8<T>::store(out + (size_t)pix * ocs + cg * 8, a);
}

inline int blocks_for(long long total, int threads) { return ysod_cdiv(total, threads); }
// grid for a grid-stride streaming kernel with `unroll` items in flight per thread: enough CTAs to fill 148 SMs x 8, no more
inline int stream_blocks(long long total, int threads, int unroll) {
    const long long need = (total + (long long)threads * unroll - 1) / ((long long)threads * unroll);
    const long long cap = 148 * 16;
    return (int)(need < 1 ? 1 : (need > cap ? cap : need));
}
// block size for StreamIdx kernels: a multiple of C/8 so that every thread keeps a fixed channel group
inline int stream_threads(int C) { const int c8n = C / 8; return c8n >= 256 ? c8n : (256 / c8n) * c8n; }
inline int ln_group(int C) {
    int g = 1;
    while (g < 32 && g < C / 8) g <<= 1;
    return g;
}

}  // namespace

#define YSOD_DISPATCH(dtype, ...)                                   \
    do {                                                            \
        if ((dtype) == YSOD_F32) { using T = float; __VA_ARGS__; }  \
        else if ((dtype) == YSOD_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
        else { ysod_set_error("bad dtype %d", (int)(dtype)); return YSOD_ERR_INVALID; } \
    } while (0)

extern "C" {

int ysod_gap_partial(const void* x, int dtype, int N, int HW, int C, int xcs, int S, float* psum, float* pmax, cudaStream_t st) {
    YSOD_CHECK_ARG(x && psum && C % 8 == 0 && xcs % 8 == 0 && S >= 1, "ysod_gap_partial: bad args");
    const int c8n = C / 8;
    YSOD_CHECK_ARG(c8n <= 256, "ysod_gap_partial: C too large");
    const int PL = 256 / c8n;
    const size_t smem = (size_t)2 * PL * C * sizeof(float);
    YSOD_CHECK_ARG(smem <= 48 * 1024, "ysod_gap_partial: smem");
    dim3 grid(S, N);
    YSOD_DISPATCH(dtype, (ysod_launch(gap_partial_kernel<T>, grid, 256, smem, st, (const T*)x, HW, C, xcs, S, psum, pmax)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

// ysod_gap_partial + ysod_se_gate (kind 0) / ysod_cbam_gate (kind 1) in one launch. counter: N zero-initialised uint32, owned by the
// caller, left at zero by every call. SE: w1 [hid][C], b1 [hid], w2 [C][hid], b2 [C]; CBAM: w1, w2 only (bias-free), pmax required.
int ysod_gap_gate(const void* x, int dtype, int N, int HW, int C, int xcs, int S, float* psum, float* pmax, void* counter, int kind,
                  const float* w1, const float* b1, const float* w2, const float* b2, int hid, float* gate, cudaStream_t st) {
    YSOD_CHECK_ARG(x && psum && counter && w1 && w2 && gate && C % 8 == 0 && xcs % 8 == 0 && S >= 1 && hid >= 1, "ysod_gap_gate: bad args");
    YSOD_CHECK_ARG(kind == 0 ? (b1 && b2) : (kind == 1 && pmax), "ysod_gap_gate: kind %d needs %s", kind, kind == 0 ? "biases" : "pmax");
    const int c8n = C / 8;
    YSOD_CHECK_ARG(c8n <= 256, "ysod_gap_gate: C too large");
    const int PL = 256 / c8n;
    size_t smem = (size_t)2 * PL * C * sizeof(float);
    const size_t gate_smem = (size_t)(2 * C + 2 * hid) * sizeof(float);
    if (smem < gate_smem) smem = gate_smem;
    YSOD_CHECK_ARG(smem <= 48 * 1024, "ysod_gap_gate: smem");
    GateArgs g;
    g.counter = (unsigned*)counter; g.kind = kind; g.hid = hid; g.w1 = w1; g.b1 = b1; g.w2 = w2; g.b2 = b2; g.gate = gate;
    dim3 grid(S, N);
    YSOD_DISPATCH(dtype, (ysod_launch(gap_gate_kernel<T>, grid, 256, smem, st, (const T*)x, HW, C, xcs, S, psum, pmax, g)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_se_gate(const float* psum, int N, int S, int HW, int C, const float* w1, const float* b1, const float* w2,
                 const float* b2, int hid, float* gate, cudaStream_t st) {
    YSOD_CHECK_ARG(psum && w1 && b1 && w2 && b2 && gate, "ysod_se_gate: null pointer");
    ysod_launch(se_gate_kernel, N, 256, (C + hid) * sizeof(float), st, psum, S, HW, C, w1, b1, w2, b2, hid, gate);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_cbam_gate(const float* psum, const float* pmax, int N, int S, int HW, int C, const float* w1, const float* w2, int hid,
                   float* gate, cudaStream_t st) {
    YSOD_CHECK_ARG(psum && pmax && w1 && w2 && gate, "ysod_cbam_gate: null pointer");
    ysod_launch(cbam_gate_kernel, N, 256, (2 * C + 2 * hid) * sizeof(float), st, psum, pmax, S, HW, C, w1, w2, hid, gate);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_scale_channels(const void* x, int dtype, int N, int HW, int C, int xcs, const float* gate, void* out, int ocs,
                        cudaStream_t st) {
    YSOD_CHECK_ARG(x && gate && out && C % 8 == 0 && xcs % 8 == 0 && ocs % 8 == 0, "ysod_scale_channels: bad args");
    const long long total = (long long)N * HW * (C / 8);
    YSOD_DISPATCH(dtype, (ysod_launch(scale_channels_kernel<T>, stream_blocks(total, stream_threads(C), 4), stream_threads(C), 0, st, (const T*)x, HW, C, xcs, gate, (T*)out, ocs, (unsigned)(total / (C / 8)))));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_cbam_stats(const void* x, int dtype, int N, int HW, int C, int xcs, const float* gate, float* stats, cudaStream_t st) {
    YSOD_CHECK_ARG(x && gate && stats && C % 8 == 0 && xcs % 8 == 0, "ysod_cbam_stats: bad args");
    const int c8n = C / 8;
    YSOD_CHECK_ARG((c8n & (c8n - 1)) == 0, "ysod_cbam_stats: C/8 must be a power of two");
    const int G = c8n < 32 ? c8n : 32;
    const long long npix = (long long)N * HW;
    if (c8n <= 32 && N <= 65535) {   // per-image streaming variant: one 16 B load per lane and pixel, four pixels in flight, gate in registers
        int S = ysod_cdiv(1184, N);                      // ~8 CTAs per SM in total
        const int ngr = 256 / c8n;
        if (S > ysod_cdiv(HW, 4 * ngr)) S = ysod_cdiv(HW, 4 * ngr);
        const int chunk = ysod_cdiv(HW, S);
        S = ysod_cdiv(HW, chunk);
        const int iters = ysod_cdiv(chunk, 4 * ngr);
        dim3 grid(S, N);
        // cp.async prefetch ring, 8 pixels in flight per lane group (profiles/r02_ab_blocks.json: 49.2 -> 41.0 us at 64 x 160^2, bit-identical;
        // depth 16 loses occupancy to its 64 KB of shared memory); YSOD_RING_DEPTH=0 selects the register-held 4-deep kernel (A/B)
        static const int ring_depth = getenv("YSOD_RING_DEPTH") ? atoi(getenv("YSOD_RING_DEPTH")) : 8;
        if (ring_depth == 8 || ring_depth == 16) {
            const size_t es = dtype == YSOD_BF16 ? 1 : 2;
            if (ring_depth == 8) {
                // (fp32 parity mode: 64 KB of ring, above the 48 KB a kernel may use without opting in)
                if (8 * 256 * 16 * es > 48 * 1024) YSOD_DISPATCH(dtype, (cudaFuncSetAttribute(cbam_stats_ring_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * 256 * 16 * es))));
                YSOD_DISPATCH(dtype, (ysod_launch(cbam_stats_ring_kernel<T, 8>, grid, 256, 8 * 256 * 16 * es, st, (const T*)x, HW, C, xcs, gate, (float2*)stats, chunk, iters)));
            }
            else {
                YSOD_DISPATCH(dtype, (cudaFuncSetAttribute(cbam_stats_ring_kernel<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(16 * 256 * 16 * es))));
                YSOD_DISPATCH(dtype, (ysod_launch(cbam_stats_ring_kernel<T, 16>, grid, 256, 16 * 256 * 16 * es, st, (const T*)x, HW, C, xcs, gate, (float2*)stats, chunk, iters)));
            }
            YSOD_LAUNCH_CHECK();
            return YSOD_OK;
        }
        YSOD_DISPATCH(dtype, (ysod_launch(cbam_stats_img_kernel<T>, grid, 256, 0, st, (const T*)x, HW, C, xcs, gate, (float2*)stats, chunk, iters)));
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
    if (c8n <= 32) {   // streaming variant: one 16 B load per lane and pixel, four pixels in flight
        const int blocks = stream_blocks(npix * c8n, 256, 4);
        const long long ngroups = (long long)blocks * 256 / c8n;
        const int iters = (int)((npix + 4 * ngroups - 1) / (4 * ngroups));
        YSOD_DISPATCH(dtype, (ysod_launch(cbam_stats_stream_kernel<T>, blocks, 256, 0, st, (const T*)x, HW, C, xcs, gate, (float2*)stats, npix, iters)));
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
    const long long warps = (npix + (32 / G) - 1) / (32 / G);
    YSOD_DISPATCH(dtype, (ysod_launch(cbam_stats_kernel<T>, blocks_for(warps * 32, 256), 256, 0, st, (const T*)x, HW, C, xcs, gate, (float2*)stats, npix)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_cbam_apply(const void* x, int dtype, int N, int H, int W, int C, int xcs, const float* gate, const float* stats,
                    const float* wsp, int ks, void* out, int ocs, cudaStream_t st) {
    YSOD_CHECK_ARG(x && gate && stats && wsp && out && C % 8 == 0 && xcs % 8 == 0 && ocs % 8 == 0, "ysod_cbam_apply: bad args");
    const int c8n = C / 8;
    YSOD_CHECK_ARG((c8n & (c8n - 1)) == 0, "ysod_cbam_apply: C/8 must be a power of two");
    const int G = c8n < 32 ? c8n : 32;
    const long long npix = (long long)N * H * W;
    if (ks == 7 && N <= 65535 && c8n <= 256) {   // tiled variant (cbam_block.py:27: kernel_size 7)
        dim3 grid(ysod_cdiv(W, 16), ysod_cdiv(H, 16), N);
        YSOD_DISPATCH(dtype, (ysod_launch(cbam_apply_tile_kernel<T>, grid, 256, 0, st, (const T*)x, H, W, C, xcs, gate, (const float2*)stats, wsp, (T*)out, ocs)));
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
    const long long warps = (npix + (32 / G) - 1) / (32 / G);
    YSOD_DISPATCH(dtype, (ysod_launch(cbam_apply_kernel<T>, blocks_for(warps * 32, 256), 256, 2 * ks * ks * sizeof(float), st, 
                             (const T*)x, H, W, C, xcs, gate, (const float2*)stats, wsp, ks, (T*)out, ocs, npix)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

// CBAM spatial attention (statistics + 7x7 conv + apply) in one pass over the map: see cbam_spatial_fused_kernel. Covers kernel_size 7
// (cbam_block.py:27) and C/8 a power of two <= 32; returns YSOD_ERR_UNSUPPORTED otherwise (callers fall back to stats + apply).
int ysod_cbam_spatial(const void* x, int dtype, int N, int H, int W, int C, int xcs, const float* gate, const float* wsp, int ks, void* out,
                      int ocs, cudaStream_t st) {
    YSOD_CHECK_ARG(x && gate && wsp && out && C % 8 == 0 && xcs % 8 == 0 && ocs % 8 == 0, "ysod_cbam_spatial: bad args");
    const int c8n = C / 8;
    if (ks != 7 || (c8n & (c8n - 1)) != 0 || c8n > 32 || N > 65535) {
        ysod_set_error("ysod_cbam_spatial: unsupported (ks %d, C %d)", ks, C);
        return YSOD_ERR_UNSUPPORTED;
    }
    dim3 grid(ysod_cdiv(W, 32), ysod_cdiv(H, 16), N);
    YSOD_DISPATCH(dtype, (ysod_launch(cbam_spatial_fused_kernel<T>, grid, 256, 0, st, (const T*)x, H, W, C, xcs, gate, wsp, (T*)out, ocs)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

// workspace floats needed by ysod_ca_pool's single-pass plan (0: the plan does not apply and the two-pass kernel runs)
long long ysod_ca_pool_workspace_floats(int N, int H, int W, int C) {
    const int c8n = C / 8;
    if (C % 8 != 0 || c8n < 1 || 256 % c8n != 0) return 0;
    const int WL = 256 / c8n;
    if (!(WL == 8 || WL == 16 || WL == 32) || W > WL * 8) return 0;
    return (long long)N * ysod_cdiv(H, 16) * W * C;
}

int ysod_ca_pool(const void* x, int dtype, int N, int H, int W, int C, int xcs, float* pooled, float* workspace, cudaStream_t st) {
    YSOD_CHECK_ARG(x && pooled && C % 8 == 0 && xcs % 8 == 0, "ysod_ca_pool: bad args");
    if (workspace && ysod_ca_pool_workspace_floats(N, H, W, C) > 0 && N <= 65535) {
        const int WL = 256 / (C / 8), RB = ysod_cdiv(H, 16);
        dim3 g2(RB, N);
        const size_t smem = (size_t)WL * C * sizeof(float);
        if (WL == 8) YSOD_DISPATCH(dtype, (ysod_launch(ca_pool_rows_kernel<T, 8>, g2, 256, smem, st, (const T*)x, H, W, C, xcs, pooled, workspace, RB)));
        else if (WL == 16) YSOD_DISPATCH(dtype, (ysod_launch(ca_pool_rows_kernel<T, 16>, g2, 256, smem, st, (const T*)x, H, W, C, xcs, pooled, workspace, RB)));
        else YSOD_DISPATCH(dtype, (ysod_launch(ca_pool_rows_kernel<T, 32>, g2, 256, smem, st, (const T*)x, H, W, C, xcs, pooled, workspace, RB)));
        YSOD_LAUNCH_CHECK();
        const int total = N * W * C;
        ysod_launch(ca_pool_cols_kernel, ysod_cdiv(total, 256), 256, 0, st, (const float*)workspace, H, W, C, RB, pooled, total);
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
    dim3 grid(H + W, N);
    const int threads = (C / 8) < 32 ? 32 : ((C / 8 + 31) / 32) * 32;
    YSOD_DISPATCH(dtype, (ysod_launch(ca_pool_kernel<T>, grid, threads > 256 ? 256 : threads, 0, st, (const T*)x, H, W, C, xcs, pooled)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_ca_gate(const float* pooled, int N, int H, int W, int C, int mip, const float* w1, const float* b1, const float* wh,
                 const float* bh, const float* ww, const float* bw, float* att, cudaStream_t st) {
    YSOD_CHECK_ARG(pooled && w1 && b1 && wh && bh && ww && bw && att, "ysod_ca_gate: null pointer");
    dim3 grid(H + W, N);
    ysod_launch(ca_gate_kernel, grid, 128, mip * sizeof(float), st, pooled, H, W, C, mip, w1, b1, wh, bh, ww, bw, att);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_ca_apply(const void* x, int dtype, int N, int H, int W, int C, int xcs, const float* att, void* out, int ocs,
                  cudaStream_t st) {
    YSOD_CHECK_ARG(x && att && out && C % 8 == 0 && xcs % 8 == 0 && ocs % 8 == 0, "ysod_ca_apply: bad args");
    const long long total = (long long)N * H * W * (C / 8);
    YSOD_DISPATCH(dtype, (ysod_launch(ca_apply_kernel<T>, stream_blocks(total, stream_threads(C), 4), stream_threads(C), 0, st, (const T*)x, H, W, C, xcs, att, (T*)out, ocs, (unsigned)(total / (C / 8)))));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_sppf_pool(const void* y0, int dtype, int N, int H, int W, int C, int xcs, int k, void* o1, void* o2, void* o3, int ocs,
                   cudaStream_t st) {
    YSOD_CHECK_ARG(y0 && o1 && o2 && o3 && C % 8 == 0 && xcs % 8 == 0 && ocs % 8 == 0 && (k & 1), "ysod_sppf_pool: bad args");
    if (dtype == YSOD_BF16 && C % 16 == 0 && H * W <= 1024) {
        dim3 grid(C / 16, N);
        const size_t smem = (size_t)2 * H * W * 2 * sizeof(uint4);
        const __nv_bfloat16* yi = (const __nv_bfloat16*)y0;
        __nv_bfloat16 *p1 = (__nv_bfloat16*)o1, *p2 = (__nv_bfloat16*)o2, *p3 = (__nv_bfloat16*)o3;
        if (H * W <= 512) {
            ysod_launch(sppf_plane_bf16_kernel<1>, grid, 1024, smem, st, yi, H, W, xcs, k, p1, p2, p3, ocs);
        } else {
            YSOD_CUDA(cudaFuncSetAttribute(sppf_plane_bf16_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            ysod_launch(sppf_plane_bf16_kernel<2>, grid, 1024, smem, st, yi, H, W, xcs, k, p1, p2, p3, ocs);
        }
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
    if (H * W <= 1024) {
        dim3 grid(C / 8, N);
        const size_t smem = (size_t)2 * H * W * 8 * sizeof(float);
        if (smem > 48 * 1024) {
            YSOD_CUDA(cudaFuncSetAttribute(sppf_plane_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            YSOD_CUDA(cudaFuncSetAttribute(sppf_plane_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        }
        YSOD_DISPATCH(dtype, (ysod_launch(sppf_plane_kernel<T>, grid, 256, smem, st, (const T*)y0, H, W, xcs, k, (T*)o1, (T*)o2, (T*)o3, ocs)));
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
    const long long total = (long long)N * H * W * (C / 8);
    YSOD_DISPATCH(dtype, (ysod_launch(sppf_pool_kernel<T>, blocks_for(total, 128), 128, 0, st, (const T*)y0, H, W, C, xcs, k, (T*)o1, (T*)o2, (T*)o3, ocs, total)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_upsample_copy(const void* x, int dtype, int N, int H, int W, int C, int xcs, int scale, void* out, int ocs,
                       cudaStream_t st) {
    YSOD_CHECK_ARG(x && out && C % 8 == 0 && xcs % 8 == 0 && ocs % 8 == 0 && scale >= 1, "ysod_upsample_copy: bad args");
    const long long total = (long long)N * H * scale * W * scale * (C / 8);
    YSOD_DISPATCH(dtype, (ysod_launch(upsample_copy_kernel<T>, stream_blocks(total, stream_threads(C), 4), stream_threads(C), 0, st, (const T*)x, H, W, C, xcs, scale, (T*)out, ocs, (unsigned)(total / (C / 8)))));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_avgpool2d(const void* x, int dtype, int N, int H, int W, int C, int xcs, int r, void* out, int ocs, cudaStream_t st) {
    YSOD_CHECK_ARG(x && out && C % 8 == 0 && xcs % 8 == 0 && ocs % 8 == 0 && r >= 1 && H >= r && W >= r, "ysod_avgpool2d: bad args");
    const long long total = (long long)N * (H / r) * (W / r) * (C / 8);
    YSOD_DISPATCH(dtype, (ysod_launch(avgpool_kernel<T>, blocks_for(total, 256), 256, 0, st, (const T*)x, H, W, C, xcs, r, (T*)out, ocs, total)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_glu(const void* x, int dtype, long long npix, int hid, int xcs, void* out, int ocs, cudaStream_t st) {
    YSOD_CHECK_ARG(x && out && hid % 8 == 0 && xcs % 8 == 0 && ocs % 8 == 0 && xcs >= 2 * hid && npix > 0, "ysod_glu: bad args");
    const long long total = npix * (hid / 8);
    YSOD_DISPATCH(dtype, (ysod_launch(glu_kernel<T>, blocks_for(total, 256), 256, 0, st, (const T*)x, hid, xcs, (T*)out, ocs, total)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_upsample_add(const void* y, int dtype, int N, int Hh, int Wh, int C, int ycs, const void* res, int rcs, int H, int W, void* out,
                      int ocs, cudaStream_t st) {
    YSOD_CHECK_ARG(y && res && out && C % 8 == 0 && ycs % 8 == 0 && rcs % 8 == 0 && ocs % 8 == 0 && Hh > 0 && Wh > 0 && H >= Hh && W >= Wh,
                   "ysod_upsample_add: bad args");
    const long long total = (long long)N * H * W * (C / 8);
    YSOD_DISPATCH(dtype, (ysod_launch(upsample_add_kernel<T>, blocks_for(total, 256), 256, 0, st, (const T*)y, Hh, Wh, C, ycs, (const T*)res, rcs, H, W,
                                      (T*)out, ocs, total)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_layernorm(const void* x, int dtype, long long rows, int C, int ldx, const float* gamma, const float* beta, float eps,
                   void* out, int ldo, cudaStream_t st) {
    YSOD_CHECK_ARG(x && gamma && beta && out && C % 8 == 0 && C <= 2048 && ldx % 8 == 0 && ldo % 8 == 0, "ysod_layernorm: bad args");
    const int G = ln_group(C);
    if (C / 8 <= 32 && ((C / 8) & (C / 8 - 1)) == 0) {
        const int blocks = stream_blocks(rows * (C / 8), 256, 4);
        const long long ngroups = (long long)blocks * 256 / (C / 8);
        const int iters = (int)((rows + 4 * ngroups - 1) / (4 * ngroups));
        YSOD_DISPATCH(dtype, (ysod_launch(layernorm_stream_kernel<T, false>, blocks, 256, 0, st, (const T*)x, rows, C, ldx, gamma, beta, eps, (T*)out, ldo,
                                                                                      iters, 0, 0, 1, 1, 1, 1, nullptr)));
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
    YSOD_DISPATCH(dtype, (ysod_launch(layernorm_kernel<T, false>, blocks_for(rows * G, 256), 256, 0, st, 
                             (const T*)x, rows, C, ldx, gamma, beta, eps, (T*)out, ldo, G, 0, 0, 1, 1, 1, 1, nullptr)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

// x: NHWC view; raw_out / norm_out: [N*nWh*nWw*wh*ww][C] token matrices (row stride ldo)
int ysod_window_partition_ln(const void* x, int dtype, int N, int H, int W, int C, int xcs, int wh, int ww, int nWh, int nWw,
                             const float* gamma, const float* beta, float eps, void* raw_out, void* norm_out, int ldo,
                             cudaStream_t st) {
    YSOD_CHECK_ARG(x && gamma && beta && norm_out && C % 8 == 0 && C <= 2048 && xcs % 8 == 0 && ldo % 8 == 0, "ysod_window_partition_ln: bad args");
    const long long rows = (long long)N * nWh * nWw * wh * ww;
    const int G = ln_group(C);
    if (C / 8 <= 32 && ((C / 8) & (C / 8 - 1)) == 0) {
        const int blocks = stream_blocks(rows * (C / 8), 256, 4);
        const long long ngroups = (long long)blocks * 256 / (C / 8);
        const int iters = (int)((rows + 4 * ngroups - 1) / (4 * ngroups));
        YSOD_DISPATCH(dtype, (ysod_launch(layernorm_stream_kernel<T, true>, blocks, 256, 0, st, (const T*)x, rows, C, xcs, gamma, beta, eps, (T*)norm_out, ldo,
                                                                                     iters, H, W, wh, ww, nWh, nWw, (T*)raw_out)));
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
    YSOD_DISPATCH(dtype, (ysod_launch(layernorm_kernel<T, true>, blocks_for(rows * G, 256), 256, 0, st, 
                             (const T*)x, rows, C, xcs, gamma, beta, eps, (T*)norm_out, ldo, G, H, W, wh, ww, nWh, nWw, (T*)raw_out)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_window_reverse(const void* tok, int dtype, int ldt, int N, int H, int W, int C, int wh, int ww, int nWh, int nWw, void* out,
                        int ocs, cudaStream_t st) {
    YSOD_CHECK_ARG(tok && out && C % 8 == 0 && ldt % 8 == 0 && ocs % 8 == 0, "ysod_window_reverse: bad args");
    const long long total = (long long)N * H * W * (C / 8);
    YSOD_DISPATCH(dtype, (ysod_launch(window_reverse_kernel<T>, stream_blocks(total, stream_threads(C), 4), stream_threads(C), 0, st, (const T*)tok, ldt, H, W, C, wh, ww, nWh, nWw, (T*)out, ocs, (unsigned)(total / (C / 8)))));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_adaptive_pool_rows(const void* x, int dtype, int N, int H, int W, int C, int xcs, int OH, void* out, int ocs,
                            cudaStream_t st) {
    YSOD_CHECK_ARG(x && out && C % 8 == 0 && xcs % 8 == 0 && ocs % 8 == 0 && OH >= 1, "ysod_adaptive_pool_rows: bad args");
    const long long total = (long long)N * OH * W * (C / 8);
    YSOD_DISPATCH(dtype, (ysod_launch(adaptive_pool_rows_kernel<T>, blocks_for(total, 256), 256, 0, st, (const T*)x, H, W, C, xcs, OH, (T*)out, ocs, total)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_bilinear_rows(const void* x, int dtype, int N, int IH, int W, int C, int xcs, int OH, void* out, int ocs, cudaStream_t st) {
    YSOD_CHECK_ARG(x && out && C % 8 == 0 && xcs % 8 == 0 && ocs % 8 == 0 && OH >= 1, "ysod_bilinear_rows: bad args");
    const long long total = (long long)N * OH * W * (C / 8);
    YSOD_DISPATCH(dtype, (ysod_launch(bilinear_rows_kernel<T>, blocks_for(total, 256), 256, 0, st, (const T*)x, IH, W, C, xcs, OH, (T*)out, ocs, total)));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

}  // extern "C"
