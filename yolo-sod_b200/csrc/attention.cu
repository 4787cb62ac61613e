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

}  // namespace

extern "C" int ysod_mha_core(const void* q, const void* k, const void* v, int dtype, int batch, int L, int heads, int D, int ldq,
                             int ldk, int ldv, long long bsq, long long bsk, long long bsv, float scale, void* out, int ldo,
                             long long bso, cudaStream_t st) {
    YSOD_CHECK_ARG(q && k && v && out, "ysod_mha_core: null pointer");
    YSOD_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "ysod_mha_core: row strides must be multiples of 8");
    YSOD_CHECK_ARG(batch <= 65535 || true, "ysod_mha_core: batch");
    dim3 grid(ysod_cdiv(L, 64), heads, batch);
    YSOD_CHECK_ARG(batch <= 65535 && heads <= 65535, "ysod_mha_core: grid too large (batch %d)", batch);
#define LAUNCH(T, DD) \
    mha_core_kernel<T, DD><<<grid, 64, 0, st>>>((const T*)q, (const T*)k, (const T*)v, L, ldq, ldk, ldv, bsq, bsk, bsv, scale, (T*)out, ldo, bso)
    if (dtype == YSOD_F32 && D == 32) LAUNCH(float, 32);
    else if (dtype == YSOD_F32 && D == 64) LAUNCH(float, 64);
    else if (dtype == YSOD_BF16 && D == 32) LAUNCH(__nv_bfloat16, 32);
    else if (dtype == YSOD_BF16 && D == 64) LAUNCH(__nv_bfloat16, 64);
    else {
        ysod_set_error("ysod_mha_core: unsupported head_dim %d / dtype %d (supported: 32, 64)", D, dtype);
        return YSOD_ERR_UNSUPPORTED;
    }
#undef LAUNCH
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}
