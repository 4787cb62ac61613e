// Multi-head softmax attention core: out = softmax(q k^T * scale) v, per (batch, head), streaming over key chunks
// with an online softmax (fp32 math, bf16/fp32 storage).
//
// Replaces (reference): the attention inside nn.MultiheadAttention of WindowAttention (blocks_transformer.py:98,116;
// 49-token windows, padded tokens attend unmasked) and A2_Attn (a2_attn.py:29,53; L = 8*W), and the manual softmax path
// of AAttn (block.py:1348-1357; area attention, L = H*W/area) -- the oracle for AAttn is that manual path, not the fp16
// flash-attn branch.
// q/k/v are addressed as ptr + batch*bs + token*ld + head*D, so packed in_proj outputs ([L][3E]) and the
// AAttn qk / v tensors are consumed in place; heads are contiguous D-slices of the embedding, as in torch.
// One thread owns one query row (q and the output accumulator live in registers); K/V chunks of 64 keys are staged
// in shared memory with coalesced 16 B loads and read back as warp-wide broadcasts.
#include "common.cuh"

namespace {

template <typename T, int D>
__global__ void __launch_bounds__(64)
mha_core_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, int L, int ldq, int ldk, int ldv,
                long long bsq, long long bsk, long long bsv, float scale, T* __restrict__ out, int ldo, long long bso) {
    ysod_pdl_sync();
    constexpr int KC = 64;
    constexpr int LDS = D + 4;
    __shared__ float Ks[KC * LDS];
    __shared__ float Vs[KC * LDS];
    const int t = threadIdx.x;
    const int h = blockIdx.y;
    const int b = blockIdx.z;
    const int qi = blockIdx.x * 64 + t;
    const bool valid = qi < L;
    float qr[D], o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { qr[d] = 0.f; o[d] = 0.f; }
    if (valid) {
        const T* qp = q + (size_t)b * bsq + (size_t)qi * ldq + h * D;
#pragma unroll
        for (int d = 0; d < D; d += 8) {
            ysod_vec8<T>::load(qp + d, qr + d);
#pragma unroll
            for (int e = 0; e < 8; ++e) qr[d + e] *= scale;
        }
    }
    float m = -INFINITY, l = 0.f;
    const T* kb = k + (size_t)b * bsk + h * D;
    const T* vb = v + (size_t)b * bsv + h * D;
    for (int kc = 0; kc < L; kc += KC) {
        const int nk = min(KC, L - kc);
        __syncthreads();
        for (int p = t; p < nk * (D / 8); p += 64) {
            const int key = p / (D / 8), part = p % (D / 8);
            float tmp[8];
            ysod_vec8<T>::load(kb + (size_t)(kc + key) * ldk + part * 8, tmp);
            *reinterpret_cast<float4*>(&Ks[key * LDS + part * 8]) = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
            *reinterpret_cast<float4*>(&Ks[key * LDS + part * 8 + 4]) = make_float4(tmp[4], tmp[5], tmp[6], tmp[7]);
            ysod_vec8<T>::load(vb + (size_t)(kc + key) * ldv + part * 8, tmp);
            *reinterpret_cast<float4*>(&Vs[key * LDS + part * 8]) = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
            *reinterpret_cast<float4*>(&Vs[key * LDS + part * 8 + 4]) = make_float4(tmp[4], tmp[5], tmp[6], tmp[7]);
        }
        __syncthreads();
        for (int j = 0; j < nk; ++j) {
            const float* kr = &Ks[j * LDS];
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) s = fmaf(qr[d], kr[d], s);
            if (s > m) {
                const float corr = __expf(m - s);
                l *= corr;
#pragma unroll
                for (int d = 0; d < D; ++d) o[d] *= corr;
                m = s;
            }
            const float pj = __expf(s - m);
            l += pj;
            const float* vr = &Vs[j * LDS];
#pragma unroll
            for (int d = 0; d < D; ++d) o[d] = fmaf(pj, vr[d], o[d]);
        }
    }
    if (valid) {
        const float inv = 1.0f / l;
        T* op = out + (size_t)b * bso + (size_t)qi * ldo + h * D;
#pragma unroll
        for (int d = 0; d < D; d += 8) {
            float r[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) r[e] = o[d + e] * inv;
            ysod_vec8<T>::store(op + d, r);
        }
    }
}

// ---- windowed variant: L <= 64 tokens (Swin 7x7 windows, 49 tokens), bf16, on warp-level tensor cores --------------------------
// One CTA of 4 warps per (batch = window, head); warp w owns query rows 16w..16w+15. Q, K ([token][D]) and V^T ([D][token]) are
// staged in shared memory (padded rows: conflict-free 32-bit fragment loads), S = QK^T and O = PV are mma.sync m16n8k16 with
// fp32 accumulators, the softmax lives in registers (quad shuffles), P is re-packed from the S accumulators into A fragments.
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." YSOD_MMA_T "." YSOD_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ld32s(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }
// ldmatrix: lane i supplies the 16 B row (i & 7) of 8x8 bf16 matrix (i >> 3); per matrix a lane receives the element pair
// (row lane/4, columns 2*(lane%4), +1) -- or, with .trans, (rows 2*(lane%4), +1, column lane/4): the m16n8k16 B fragment of a
// [n][k] (plain) or [k][n] (.trans) operand.
__device__ __forceinline__ void ldsm_x4(uint32_t* r, const __nv_bfloat16* row) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2(uint32_t* r, const __nv_bfloat16* row) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t* r, const __nv_bfloat16* row) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ uint32_t pack2bf(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <int D>
__global__ void __launch_bounds__(128)
mha_win_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v, int L,
               int ldq, int ldk, int ldv, long long bsq, long long bsk, long long bsv, float scale, __nv_bfloat16* __restrict__ out,
               int ldo, long long bso) {
    ysod_pdl_sync();
    constexpr int LDQ = D + 8;      // row stride 2D + 16 bytes: 16 B aligned, conflict-free for 8-row ldmatrix phases (D = 16, 32, 64)
    __shared__ __align__(16) __nv_bfloat16 Qs[64 * LDQ];
    __shared__ __align__(16) __nv_bfloat16 Ks[64 * LDQ];
    __shared__ __align__(16) __nv_bfloat16 Vs[64 * LDQ];   // V stays row-major [key][d]; PV reads it through ldmatrix.trans
    const int h = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const __nv_bfloat16* qb = q + (size_t)b * bsq + h * D;
    const __nv_bfloat16* kb = k + (size_t)b * bsk + h * D;
    const __nv_bfloat16* vb = v + (size_t)b * bsv + h * D;
    for (int i = tid; i < 64 * (D / 8); i += 128) {
        const int r = i / (D / 8), pc = i - r * (D / 8);
        uint4 qv = make_uint4(0, 0, 0, 0), kv = qv, vv = qv;
        if (r < L) {
            qv = *reinterpret_cast<const uint4*>(qb + (size_t)r * ldq + pc * 8);
            kv = *reinterpret_cast<const uint4*>(kb + (size_t)r * ldk + pc * 8);
            vv = *reinterpret_cast<const uint4*>(vb + (size_t)r * ldv + pc * 8);
        }
        *reinterpret_cast<uint4*>(&Qs[r * LDQ + pc * 8]) = qv;
        *reinterpret_cast<uint4*>(&Ks[r * LDQ + pc * 8]) = kv;
        *reinterpret_cast<uint4*>(&Vs[r * LDQ + pc * 8]) = vv;
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
        // K rows nb*8 .. nb*8+7 as [n = key][k = d]: one ldmatrix.x4 = the (b0, b1) pairs of two K steps
#pragma unroll
        for (int k2 = 0; k2 < D / 32; ++k2) {
            uint32_t kb[4];
            ldsm_x4(kb, &Ks[(nb * 8 + (lane & 7)) * LDQ + k2 * 32 + (lane >> 3) * 8]);
            mma16816(s[nb], qf[2 * k2], kb[0], kb[1]);
            mma16816(s[nb], qf[2 * k2 + 1], kb[2], kb[3]);
        }
        if (D % 32) {   // D = 16: a single K step
            uint32_t kb[2];
            ldsm_x2(kb, &Ks[(nb * 8 + (lane & 7)) * LDQ + (D / 32) * 32 + ((lane >> 3) & 1) * 8]);
            mma16816(s[nb], qf[D / 16 - 1], kb[0], kb[1]);
        }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        const int key = nb * 8 + 2 * t;
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nb][e] *= scale;
        if (key >= L) { s[nb][0] = -INFINITY; s[nb][2] = -INFINITY; }
        if (key + 1 >= L) { s[nb][1] = -INFINITY; s[nb][3] = -INFINITY; }
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
        // V as [k = key][n = d]: ldmatrix.x4.trans over keys 32*i2 .. 32*i2+31 (lane = key row), d columns nb*8 .. +7
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
    __nv_bfloat16* ob = out + (size_t)b * bso + h * D;
    for (int i = tid; i < L * (D / 8); i += 128) {
        const int r = i / (D / 8), pc = i - r * (D / 8);
        *reinterpret_cast<uint4*>(ob + (size_t)r * ldo + pc * 8) = *reinterpret_cast<const uint4*>(&Qs[r * LDQ + pc * 8]);
    }
}

// ---- general variant: any L, bf16, flash-style on warp-level tensor cores -------------------------------------------------------
// CTA = 64 query rows of one (batch, head) (warp w: rows 16w..16w+15), looping over key blocks of 64 with an online softmax:
// m' = max(m, rowmax S), P = exp(S - m'), l = l*exp(m - m') + rowsum P, O = O*exp(m - m') + P V. Used by A2_Attn (L = 8*W) and the
// yolov12 area attention (L = H*W/area), whose reference semantics are the manual softmax path (block.py:1348-1357).
template <int D>
__global__ void __launch_bounds__(128)
mha_flash_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v, int L,
                 int ldq, int ldk, int ldv, long long bsq, long long bsk, long long bsv, float scale, __nv_bfloat16* __restrict__ out,
                 int ldo, long long bso) {
    ysod_pdl_sync();
    constexpr int LDQ = D + 8, LDV = 64 + 8;
    __shared__ __align__(16) __nv_bfloat16 Qs[64 * LDQ];
    __shared__ __align__(16) __nv_bfloat16 Ks[64 * LDQ];
    __shared__ __align__(16) __nv_bfloat16 Vt[D * LDV];
    const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const __nv_bfloat16* qb = q + (size_t)b * bsq + h * D;
    const __nv_bfloat16* kb = k + (size_t)b * bsk + h * D;
    const __nv_bfloat16* vb = v + (size_t)b * bsv + h * D;
    constexpr int NIT = 64 * (D / 8) / 128;   // 16 B pieces of a 64-row tile per thread
    {
        uint4 qv[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int i = tid + it * 128;
            const int r = i / (D / 8), pc = i - r * (D / 8);
            qv[it] = make_uint4(0, 0, 0, 0);
            if (q0 + r < L) qv[it] = *reinterpret_cast<const uint4*>(qb + (size_t)(q0 + r) * ldq + pc * 8);
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int i = tid + it * 128;
            const int r = i / (D / 8), pc = i - r * (D / 8);
            *reinterpret_cast<uint4*>(&Qs[r * LDQ + pc * 8]) = qv[it];
        }
    }
    // the K / V rows of a 64-key block are requested one block ahead: they travel under the previous block's MMAs and softmax
    uint4 kvr[NIT], vvr[NIT];
    auto request_kv = [&](int k0) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int i = tid + it * 128;
            const int r = i / (D / 8), pc = i - r * (D / 8);
            kvr[it] = make_uint4(0, 0, 0, 0); vvr[it] = kvr[it];
            if (k0 + r < L) {
                kvr[it] = *reinterpret_cast<const uint4*>(kb + (size_t)(k0 + r) * ldk + pc * 8);
                vvr[it] = *reinterpret_cast<const uint4*>(vb + (size_t)(k0 + r) * ldv + pc * 8);
            }
        }
    };
    request_kv(0);
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
    float o[D / 8][4];
#pragma unroll
    for (int nb = 0; nb < D / 8; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int k0 = 0; k0 < L; k0 += 64) {
        __syncthreads();   // previous block's K / V^T fully consumed
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int i = tid + it * 128;
            const int r = i / (D / 8), pc = i - r * (D / 8);
            *reinterpret_cast<uint4*>(&Ks[r * LDQ + pc * 8]) = kvr[it];
            const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&vvr[it]);
#pragma unroll
            for (int e = 0; e < 8; ++e) Vt[(pc * 8 + e) * LDV + r] = ve[e];
        }
        __syncthreads();
        if (k0 + 64 < L) request_kv(k0 + 64);
        float s[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
            const __nv_bfloat16* kr = &Ks[(nb * 8 + g) * LDQ + 2 * t];
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) mma16816(s[nb], qf[ks], ld32s(kr + ks * 16), ld32s(kr + ks * 16 + 8));
        }
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int key = k0 + nb * 8 + 2 * t;
#pragma unroll
            for (int e = 0; e < 4; ++e) s[nb][e] *= scale;
            if (key >= L) { s[nb][0] = -INFINITY; s[nb][2] = -INFINITY; }
            if (key + 1 >= L) { s[nb][1] = -INFINITY; s[nb][3] = -INFINITY; }
            mx0 = fmaxf(mx0, fmaxf(s[nb][0], s[nb][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nb][2], s[nb][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float c0 = __expf(m0 - mx0), c1 = __expf(m1 - mx1);   // exp(-inf) = 0 on the first block
        m0 = mx0; m1 = mx1;
        float r0 = 0.f, r1 = 0.f;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            s[nb][0] = __expf(s[nb][0] - m0); s[nb][1] = __expf(s[nb][1] - m0);
            s[nb][2] = __expf(s[nb][2] - m1); s[nb][3] = __expf(s[nb][3] - m1);
            r0 += s[nb][0] + s[nb][1];
            r1 += s[nb][2] + s[nb][3];
        }
        r0 += __shfl_xor_sync(0xffffffffu, r0, 1); r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
        r1 += __shfl_xor_sync(0xffffffffu, r1, 1); r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
        l0 = l0 * c0 + r0;
        l1 = l1 * c1 + r1;
        uint32_t pf[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            pf[i][0] = pack2bf(s[2 * i][0], s[2 * i][1]);
            pf[i][1] = pack2bf(s[2 * i][2], s[2 * i][3]);
            pf[i][2] = pack2bf(s[2 * i + 1][0], s[2 * i + 1][1]);
            pf[i][3] = pack2bf(s[2 * i + 1][2], s[2 * i + 1][3]);
        }
#pragma unroll
        for (int nb = 0; nb < D / 8; ++nb) {
            o[nb][0] *= c0; o[nb][1] *= c0; o[nb][2] *= c1; o[nb][3] *= c1;
            const __nv_bfloat16* vr = &Vt[(nb * 8 + g) * LDV + 2 * t];
#pragma unroll
            for (int i = 0; i < 4; ++i) mma16816(o[nb], pf[i], ld32s(vr + i * 16), ld32s(vr + i * 16 + 8));
        }
    }
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    // Qs rows of this warp were only read by this warp (fragments are in registers): reuse them as the output staging tile
#pragma unroll
    for (int nb = 0; nb < D / 8; ++nb) {
        *reinterpret_cast<uint32_t*>(&Qs[row0 * LDQ + nb * 8 + 2 * t]) = pack2bf(o[nb][0] * inv0, o[nb][1] * inv0);
        *reinterpret_cast<uint32_t*>(&Qs[(row0 + 8) * LDQ + nb * 8 + 2 * t]) = pack2bf(o[nb][2] * inv1, o[nb][3] * inv1);
    }
    __syncthreads();
    __nv_bfloat16* ob = out + (size_t)b * bso + h * D;
    for (int i = tid; i < 64 * (D / 8); i += 128) {
        const int r = i / (D / 8), pc = i - r * (D / 8);
        if (q0 + r < L) *reinterpret_cast<uint4*>(ob + (size_t)(q0 + r) * ldo + pc * 8) = *reinterpret_cast<const uint4*>(&Qs[r * LDQ + pc * 8]);
    }
}

}  // namespace

static int mha_core_chunk(const void* q, const void* k, const void* v, int dtype, int batch, int L, int heads, int D, int ldq,
                          int ldk, int ldv, long long bsq, long long bsk, long long bsv, float scale, void* out, int ldo,
                          long long bso, cudaStream_t st) {
    if (dtype == YSOD_BF16 && L <= 64 && (D == 16 || D == 32 || D == 64) && batch <= 65535 && ((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 &&
        ((uintptr_t)v % 16) == 0 && ((uintptr_t)out % 16) == 0 && bsq % 8 == 0 && bsk % 8 == 0 && bsv % 8 == 0 && bso % 8 == 0) {
        dim3 wgrid(heads, batch);
        if (D == 16)   // 4 heads over 64 channels: the P2 SwinBlock of yolov12-sod-fusion-v5-stable.yaml
            ysod_launch(mha_win_kernel<16>, wgrid, 128, 0, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, L, ldq, ldk, ldv,
                                                      bsq, bsk, bsv, scale, (__nv_bfloat16*)out, ldo, bso);
        else if (D == 32)
            ysod_launch(mha_win_kernel<32>, wgrid, 128, 0, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, L, ldq, ldk, ldv,
                                                      bsq, bsk, bsv, scale, (__nv_bfloat16*)out, ldo, bso);
        else
            ysod_launch(mha_win_kernel<64>, wgrid, 128, 0, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, L, ldq, ldk, ldv,
                                                      bsq, bsk, bsv, scale, (__nv_bfloat16*)out, ldo, bso);
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
    dim3 grid(ysod_cdiv(L, 64), heads, batch);
    YSOD_CHECK_ARG(batch <= 65535 && heads <= 65535, "ysod_mha_core: grid too large (batch %d)", batch);
    if (dtype == YSOD_BF16 && (D == 32 || D == 64) && ((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 &&
        ((uintptr_t)out % 16) == 0 && bsq % 8 == 0 && bsk % 8 == 0 && bsv % 8 == 0 && bso % 8 == 0) {
        if (D == 32)
            ysod_launch(mha_flash_kernel<32>, grid, 128, 0, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, L, ldq, ldk, ldv,
                                                       bsq, bsk, bsv, scale, (__nv_bfloat16*)out, ldo, bso);
        else
            ysod_launch(mha_flash_kernel<64>, grid, 128, 0, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, L, ldq, ldk, ldv,
                                                       bsq, bsk, bsv, scale, (__nv_bfloat16*)out, ldo, bso);
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
#define LAUNCH(T, DD) \
    ysod_launch(mha_core_kernel<T, DD>, grid, 64, 0, st, (const T*)q, (const T*)k, (const T*)v, L, ldq, ldk, ldv, bsq, bsk, bsv, scale, (T*)out, ldo, bso)
    if (dtype == YSOD_F32 && D == 16) LAUNCH(float, 16);
    else if (dtype == YSOD_BF16 && D == 16) LAUNCH(__nv_bfloat16, 16);
    else if (dtype == YSOD_F32 && D == 32) LAUNCH(float, 32);
    else if (dtype == YSOD_F32 && D == 64) LAUNCH(float, 64);
    else if (dtype == YSOD_BF16 && D == 32) LAUNCH(__nv_bfloat16, 32);
    else if (dtype == YSOD_BF16 && D == 64) LAUNCH(__nv_bfloat16, 64);
    else {
        ysod_set_error("ysod_mha_core: unsupported head_dim %d / dtype %d (supported: 16, 32, 64)", D, dtype);
        return YSOD_ERR_UNSUPPORTED;
    }
#undef LAUNCH
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

int ysod_mha_tc_launch(const void* q, const void* k, const void* v, int batch, int L, int heads, int D, int ldq, int ldk, int ldv,
                       long long bsq, long long bsk, long long bsv, float scale, void* out, int ldo, long long bso, cudaStream_t st);

// impl: 0 = auto (see below), 1 = this file's mma.sync / CUDA-core kernels only, 2 = the tcgen05 / TMEM kernel of attention_tc.cu
// (16-bit storage, head_dim 32 / 64; error if the shape is not covered).
extern "C" int ysod_mha_core_ex(const void* q, const void* k, const void* v, int dtype, int batch, int L, int heads, int D, int ldq,
                                int ldk, int ldv, long long bsq, long long bsk, long long bsv, float scale, void* out, int ldo,
                                long long bso, int impl, cudaStream_t st) {
    YSOD_CHECK_ARG(q && k && v && out, "ysod_mha_core: null pointer");
    YSOD_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "ysod_mha_core: row strides must be multiples of 8");
    YSOD_CHECK_ARG(impl >= 0 && impl <= 2, "ysod_mha_core_ex: impl %d", impl);
    // auto: the measured-faster kernel per shape (profiles/r02_ab_attention.json). The first tcgen05 version serialises load -> MMA ->
    // softmax -> MMA inside a CTA and is 0.4-0.7x the mma.sync kernels on every shape the models launch, so auto keeps those;
    // the tcgen05 kernel runs on request (impl 2: engine option attn_impl=2).
    if (impl == 2 && dtype == YSOD_BF16) {
        const int rc = ysod_mha_tc_launch(q, k, v, batch, L, heads, D, ldq, ldk, ldv, bsq, bsk, bsv, scale, out, ldo, bso, st);
        if (rc != YSOD_ERR_UNSUPPORTED) return rc;
    }
    YSOD_CHECK_ARG(impl != 2, "ysod_mha_core_ex: shape not covered by the tcgen05 kernel (head_dim %d, dtype %d)", D, dtype);
    // gridDim.y / .z are limited to 65535: large window batches (e.g. the unfused P2 SwinBlock at B >= 124) run as chunks of
    // <= 65535 (batch, ...) slices with offset base pointers
    const size_t es = dtype == YSOD_BF16 ? 2 : 4;
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
        const int rc = mha_core_chunk((const char*)q + (size_t)b0 * bsq * es, (const char*)k + (size_t)b0 * bsk * es,
                                      (const char*)v + (size_t)b0 * bsv * es, dtype, nb, L, heads, D, ldq, ldk, ldv, bsq, bsk, bsv, scale,
                                      (char*)out + (size_t)b0 * bso * es, ldo, bso, st);
        if (rc != YSOD_OK) return rc;
    }
    return YSOD_OK;
}

extern "C" int ysod_mha_core(const void* q, const void* k, const void* v, int dtype, int batch, int L, int heads, int D, int ldq,
                             int ldk, int ldv, long long bsq, long long bsk, long long bsv, float scale, void* out, int ldo,
                             long long bso, cudaStream_t st) {
    return ysod_mha_core_ex(q, k, v, dtype, batch, L, heads, D, ldq, ldk, ldv, bsq, bsk, bsv, scale, out, ldo, bso, 0, st);
}
