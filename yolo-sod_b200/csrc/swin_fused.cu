// Fully fused SwinBlock for the P2 level (C = 64, 2 heads of 32, 7x7 windows): one kernel instead of eleven.
//
// Replaces (reference): ultralytics/nn/modules/blocks_transformer.py:133-171 SwinBlock.forward
//     dw 3x3 (no bias) -> zero-pad to multiples of 7 -> window_partition (:8-47) -> WindowAttention (:81-131:
//     x + MHA(LN(x)), x + MLP(LN(x)) with Linear -> exact GELU -> Linear) -> window_reverse / crop (:49-79) ->
//     pw 1x1 (no bias) -> BN -> SiLU -> + identity
// for the shape it takes at P2 (layer 28 of yolov12-sod-fusion-v5-simple: 64 channels, 160x160 @640 -> 529 windows per image,
// 829 k tokens per 32-image batch). Unfused, the eleven launches move ~3.4 GB through HBM for 210 MB of real input+output.
//
// One group of 4 warps owns one window (49 tokens padded to 64 MMA rows, warp w = token rows 16w..16w+15), three groups per
// persistent CTA share the ~82 KB of weights kept in shared memory (12 warps per SM). Everything between the input patch load and the output
// store lives in registers / shared memory: every GEMM (QKV, QK^T, PV, out_proj, MLP, pw) is warp-level mma.sync m16n8k16
// (bf16 in, fp32 accumulate) with the accumulator -> A-fragment register re-packing trick, LayerNorm / softmax row
// reductions are quad shuffles.
// This is the mma.sync version (round 1), kept as the A/B baseline of swin_tc.cu: the tcgen05 / TMEM / TMA kernel there (two windows per
// M = 128 tile, row-wise stages on fp32 pairs) runs the same block in 322 us against 467 us for this one at B = 32; round 1's
// guess that K = 32..128 per GEMM and M = 49 tokens per window make tcgen05 the wrong shape did not survive the measurement.
// Zero-padded tokens (pixels beyond H/W) take part in attention unmasked, exactly as in the reference; only the MMA padding
// rows 49..63 are masked as keys.
#include "common.cuh"

namespace {

constexpr int C = 64, WS = 7, T = 49, HEADS = 2, HD = 32;
constexpr int LDW = C + 8;        // padded row stride (bf16) of 64-wide operand rows: conflict-free 32-bit fragment loads
constexpr int LDW2 = 128 + 8;     // rows of the 128-wide mlp.2 weight
constexpr int LDK = HD + 8;       // K rows  [key][d]
constexpr int LDV = 64 + 8;       // V^T rows [d][key]
constexpr int PATCH = 81;         // 9 x 9 input pixels around a 7 x 7 window
constexpr int NGROUP = 3;         // windows in flight per CTA
constexpr int THREADS = NGROUP * 128;

// bf16 weight blob offsets (elements) in global memory: dw[9][64] | wqkv[192][64] | wo[64][64] | w1[128][64] | w2[64][128] | wpw[64][64]
constexpr int G_DW = 0, G_QKV = 576, G_WO = G_QKV + 192 * 64, G_W1 = G_WO + 64 * 64, G_W2 = G_W1 + 128 * 64, G_PW = G_W2 + 64 * 128;
// fp32 blob: ln1_g | ln1_b | bqkv[192] | bo | ln2_g | ln2_b | b1[128] | b2 | bpw
constexpr int F_LN1G = 0, F_LN1B = 64, F_BQKV = 128, F_BO = 320, F_LN2G = 384, F_LN2B = 448, F_B1 = 512, F_B2 = 640, F_BPW = 704, F_TOTAL = 768;

struct Smem {
    __nv_bfloat16 wqkv[192 * LDW];
    __nv_bfloat16 wo[64 * LDW];
    __nv_bfloat16 w1[128 * LDW];
    __nv_bfloat16 w2[64 * LDW2];
    __nv_bfloat16 wpw[64 * LDW];
    float pf[F_TOTAL];
    struct Group {
        __nv_bfloat16 patch[PATCH * LDW];     // SwinBlock input around the window (identity residual = its centre)
        __nv_bfloat16 tok[64 * LDW];          // dw output tokens; reused as the output staging tile
        __nv_bfloat16 ks[HEADS * 64 * LDK];   // K   [head][key][d]
        __nv_bfloat16 vt[HEADS * HD * LDV];   // V^T [head][d][key]
    } g[NGROUP];
};

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." YSOD_MMA_T "." YSOD_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x * sigmoid(x) = h + h * tanh(h), h = x / 2: one MUFU op (same form as the tcgen05 conv epilogue)
__device__ __forceinline__ float silu_tanh(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
__device__ __forceinline__ uint32_t ld32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t v) { return ysod_unpack2(v); }
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return v;
}
__device__ __forceinline__ void group_barrier(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

// Four 8x8 bf16 matrices in one shared-memory instruction: lane i supplies the 16 B row (i & 7) of matrix (i >> 3); every lane
// receives, per matrix, the two elements (row lane/4, columns 2*(lane%4), +1) -- exactly the m16n8k16 B-fragment registers.
__device__ __forceinline__ void ldsm4(uint32_t* r, const __nv_bfloat16* row) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// B fragments of two consecutive K steps (32 columns starting at `w8`'s row base) for the 8 operand rows starting at row 0 of `W8`:
// matrices 0..3 = columns [0,8) [8,16) [16,24) [24,32)  ->  (b0, b1) of K step 0, (b0, b1) of K step 1
__device__ __forceinline__ void ldsm_b2(uint32_t* r, const __nv_bfloat16* W8, int ldw, int lane) {
    ldsm4(r, W8 + (size_t)(lane & 7) * ldw + (lane >> 3) * 8);
}

// acc[NB][4] += A[KS][4] (16 rows x 16*KS) * W^T, W rows nb*8 .. nb*8+7 (row stride ldw, 16 B aligned rows), KS even
template <int NB, int KS>
__device__ __forceinline__ void gemm_frag(float (*acc)[4], const uint32_t (*a)[4], const __nv_bfloat16* W, int ldw, int lane) {
    static_assert(KS % 2 == 0, "two K steps per ldmatrix.x4");
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
        for (int k2 = 0; k2 < KS / 2; ++k2) {
            uint32_t b[4];
            ldsm_b2(b, W + (size_t)(nb * 8) * ldw + k2 * 32, ldw, lane);
            mma16816(acc[nb], a[2 * k2], b[0], b[1]);
            mma16816(acc[nb], a[2 * k2 + 1], b[2], b[3]);
        }
    }
}
// accumulator layout (row g: [nb][0..1], row g+8: [nb][2..3], columns 8nb + 2t) -> A fragments of the next GEMM (K step i = n-blocks 2i, 2i+1)
template <int KS>
__device__ __forceinline__ void acc_to_frag(uint32_t (*a)[4], const float (*acc)[4]) {
#pragma unroll
    for (int i = 0; i < KS; ++i) {
        a[i][0] = pack2(acc[2 * i][0], acc[2 * i][1]);
        a[i][1] = pack2(acc[2 * i][2], acc[2 * i][3]);
        a[i][2] = pack2(acc[2 * i + 1][0], acc[2 * i + 1][1]);
        a[i][3] = pack2(acc[2 * i + 1][2], acc[2 * i + 1][3]);
    }
}
__device__ __forceinline__ void bias_init(float (*acc)[4], const float* b, int t) {
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        const float2 bv = *reinterpret_cast<const float2*>(b + nb * 8 + 2 * t);
        acc[nb][0] = bv.x; acc[nb][1] = bv.y; acc[nb][2] = bv.x; acc[nb][3] = bv.y;
    }
}
// LayerNorm over the 64 channels of rows g and g+8 (values in accumulator layout), result as A fragments. The affine part (gamma, beta)
// is folded into the weights / bias of the linear layer that consumes the result (W' = W diag(gamma), b' = b + W beta: the caller's
// job, see ysod_swin64_fused), so normalising is one FFMA per element: v * rstd - mean * rstd.
__device__ __forceinline__ void layernorm_frag(uint32_t (*a)[4], const float (*v)[4]) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) { s0 += v[nb][0] + v[nb][1]; s1 += v[nb][2] + v[nb][3]; }
    const float m0 = quad_sum(s0) * (1.0f / 64.0f), m1 = quad_sum(s1) * (1.0f / 64.0f);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        const float d0 = v[nb][0] - m0, d1 = v[nb][1] - m0, d2 = v[nb][2] - m1, d3 = v[nb][3] - m1;
        q0 = fmaf(d0, d0, fmaf(d1, d1, q0));
        q1 = fmaf(d2, d2, fmaf(d3, d3, q1));
    }
    const float r0 = rsqrtf(quad_sum(q0) * (1.0f / 64.0f) + 1e-5f), r1 = rsqrtf(quad_sum(q1) * (1.0f / 64.0f) + 1e-5f);
    const float c0 = -m0 * r0, c1 = -m1 * r1;
    float y[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        y[nb][0] = fmaf(v[nb][0], r0, c0);
        y[nb][1] = fmaf(v[nb][1], r0, c0);
        y[nb][2] = fmaf(v[nb][2], r1, c1);
        y[nb][3] = fmaf(v[nb][3], r1, c1);
    }
    acc_to_frag<4>(a, y);
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(THREADS, 1)
swin64_fused_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int xcs, const __nv_bfloat16* __restrict__ wb,
                    const float* __restrict__ pf, __nv_bfloat16* __restrict__ out, int ocs, int nWh, int nWw) {
    ysod_pdl_sync();
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x;

    // ---- weights -> shared memory (padded rows), once per CTA
    for (int i = tid; i < 192 * 8; i += THREADS) *reinterpret_cast<uint4*>(&sm.wqkv[(i >> 3) * LDW + (i & 7) * 8]) = *reinterpret_cast<const uint4*>(wb + G_QKV + i * 8);
    for (int i = tid; i < 64 * 8; i += THREADS) *reinterpret_cast<uint4*>(&sm.wo[(i >> 3) * LDW + (i & 7) * 8]) = *reinterpret_cast<const uint4*>(wb + G_WO + i * 8);
    for (int i = tid; i < 128 * 8; i += THREADS) *reinterpret_cast<uint4*>(&sm.w1[(i >> 3) * LDW + (i & 7) * 8]) = *reinterpret_cast<const uint4*>(wb + G_W1 + i * 8);
    for (int i = tid; i < 64 * 16; i += THREADS) *reinterpret_cast<uint4*>(&sm.w2[(i >> 4) * LDW2 + (i & 15) * 8]) = *reinterpret_cast<const uint4*>(wb + G_W2 + i * 8);
    for (int i = tid; i < 64 * 8; i += THREADS) *reinterpret_cast<uint4*>(&sm.wpw[(i >> 3) * LDW + (i & 7) * 8]) = *reinterpret_cast<const uint4*>(wb + G_PW + i * 8);
    for (int i = tid; i < F_TOTAL; i += THREADS) sm.pf[i] = pf[i];

    const int grp = tid >> 7, gt = tid & 127;          // window group and thread inside it
    const int warp = gt >> 5, lane = gt & 31;
    const int g = lane >> 2, t = lane & 3;
    Smem::Group& sg = sm.g[grp];
    // depthwise weights of this thread's channel pair (dw stage: thread = channel pair x token residue class)
    const int dwc = (gt & 31) * 2, dwr = gt >> 5;
    float2 wdw[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wdw[k] = unpack2(ld32(wb + G_DW + k * 64 + dwc));
    // MMA padding rows 49..63 of the token tile stay zero for the whole kernel
    for (int i = gt; i < (64 - T) * (C / 2); i += 128) {
        const int r = T + i / (C / 2), c2 = i % (C / 2);
        *reinterpret_cast<uint32_t*>(&sg.tok[r * LDW + 2 * c2]) = 0u;
    }
    __syncthreads();

    const long long nwin = (long long)N * nWh * nWw;
    for (long long win = (long long)blockIdx.x * NGROUP + grp; win < nwin; win += (long long)gridDim.x * NGROUP) {
        const int wj = (int)(win % nWw);
        const int wi = (int)((win / nWw) % nWh);
        const int n = (int)(win / ((long long)nWw * nWh));
        const int h0 = wi * WS, w0 = wj * WS;

        // ---- A. input patch (9 x 9 pixels x 64 ch, zero outside the image)
        {
            // all of a thread's loads are issued before the first store, so the ~6 global round trips overlap instead of serialising
            constexpr int NIT = (PATCH * 8 + 127) / 128;
            uint4 pv[NIT];
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int i = gt + it * 128;
                const int p = i >> 3, pc = i & 7;
                const int pr = p / 9;
                const int ih = h0 - 1 + pr, iw = w0 - 1 + (p - pr * 9);
                pv[it] = make_uint4(0, 0, 0, 0);
                if (i < PATCH * 8 && ih >= 0 && ih < H && iw >= 0 && iw < W)
                    pv[it] = *reinterpret_cast<const uint4*>(x + (((size_t)n * H + ih) * W + iw) * xcs + pc * 8);
            }
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int i = gt + it * 128;
                if (i < PATCH * 8) *reinterpret_cast<uint4*>(&sg.patch[(i >> 3) * LDW + (i & 7) * 8]) = pv[it];
            }
        }
        group_barrier(grp);

        // ---- B. depthwise 3x3 -> tokens (bf16). Tokens beyond the image are the zero padding of window_partition.
        //      A warp owns whole window columns (j = warp, warp + 4), a lane one channel pair: walking down the 9 patch rows of a
        //      column, every loaded pixel feeds the three output rows it belongs to (27 shared loads per column instead of 63);
        //      each output still accumulates its taps in (r, s) order, so the rounding is unchanged.
        for (int j = dwr; j < WS; j += 4) {
            float2 acc[WS];
#pragma unroll
            for (int i = 0; i < WS; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
            for (int ri = 0; ri < WS + 2; ++ri) {
                const __nv_bfloat16* pr = &sg.patch[(ri * 9 + j) * LDW + dwc];
                const float2 x0 = unpack2(ld32(pr)), x1 = unpack2(ld32(pr + LDW)), x2 = unpack2(ld32(pr + 2 * LDW));
#pragma unroll
                for (int r = 2; r >= 0; --r) {        // output row i = ri - r; for a fixed i the rows arrive in r = 0, 1, 2 order
                    const int i = ri - r;
                    if (i >= 0 && i < WS) {
                        acc[i].x = fmaf(x0.x, wdw[r * 3].x, acc[i].x);     acc[i].y = fmaf(x0.y, wdw[r * 3].y, acc[i].y);
                        acc[i].x = fmaf(x1.x, wdw[r * 3 + 1].x, acc[i].x); acc[i].y = fmaf(x1.y, wdw[r * 3 + 1].y, acc[i].y);
                        acc[i].x = fmaf(x2.x, wdw[r * 3 + 2].x, acc[i].x); acc[i].y = fmaf(x2.y, wdw[r * 3 + 2].y, acc[i].y);
                    }
                }
            }
            const bool col_in = w0 + j < W;
#pragma unroll
            for (int i = 0; i < WS; ++i) {
                const bool in = col_in && (h0 + i < H);
                *reinterpret_cast<uint32_t*>(&sg.tok[(i * WS + j) * LDW + dwc]) = in ? pack2(acc[i].x, acc[i].y) : 0u;
            }
        }
        group_barrier(grp);

        // ---- C. raw token fragments of this warp's 16 rows, LayerNorm 1
        const int row0 = warp * 16 + g;
        uint32_t raw[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            raw[ks][0] = ld32(&sg.tok[row0 * LDW + ks * 16 + 2 * t]);
            raw[ks][1] = ld32(&sg.tok[(row0 + 8) * LDW + ks * 16 + 2 * t]);
            raw[ks][2] = ld32(&sg.tok[row0 * LDW + ks * 16 + 8 + 2 * t]);
            raw[ks][3] = ld32(&sg.tok[(row0 + 8) * LDW + ks * 16 + 8 + 2 * t]);
        }
        float x1[8][4];   // running token value in accumulator layout: raw tokens now, x + attn after step F
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            float2 f;
            f = unpack2(raw[ks][0]); x1[2 * ks][0] = f.x; x1[2 * ks][1] = f.y;
            f = unpack2(raw[ks][1]); x1[2 * ks][2] = f.x; x1[2 * ks][3] = f.y;
            f = unpack2(raw[ks][2]); x1[2 * ks + 1][0] = f.x; x1[2 * ks + 1][1] = f.y;
            f = unpack2(raw[ks][3]); x1[2 * ks + 1][2] = f.x; x1[2 * ks + 1][3] = f.y;
        }
        uint32_t xn[4][4];
        layernorm_frag(xn, x1);

        // ---- D. packed in_proj: K and V go to shared memory (all rows of the window are needed by every warp), Q stays in registers
        uint32_t qf[4][4];
        {
            float acc[8][4];
            bias_init(acc, sm.pf + F_BQKV + 64, t);                         // K = rows [64,128) of in_proj_weight
            gemm_frag<8, 4>(acc, xn, sm.wqkv + 64 * LDW, LDW, lane);
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const int h = nb >> 2, d = (nb & 3) * 8 + 2 * t;
                *reinterpret_cast<uint32_t*>(&sg.ks[(h * 64 + row0) * LDK + d]) = pack2(acc[nb][0], acc[nb][1]);
                *reinterpret_cast<uint32_t*>(&sg.ks[(h * 64 + row0 + 8) * LDK + d]) = pack2(acc[nb][2], acc[nb][3]);
            }
            bias_init(acc, sm.pf + F_BQKV + 128, t);                        // V = rows [128,192)
            gemm_frag<8, 4>(acc, xn, sm.wqkv + 128 * LDW, LDW, lane);
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const int h = nb >> 2, d = (nb & 3) * 8 + 2 * t;
                sg.vt[(h * HD + d) * LDV + row0] = __float2bfloat16_rn(acc[nb][0]);
                sg.vt[(h * HD + d + 1) * LDV + row0] = __float2bfloat16_rn(acc[nb][1]);
                sg.vt[(h * HD + d) * LDV + row0 + 8] = __float2bfloat16_rn(acc[nb][2]);
                sg.vt[(h * HD + d + 1) * LDV + row0 + 8] = __float2bfloat16_rn(acc[nb][3]);
            }
            bias_init(acc, sm.pf + F_BQKV, t);                              // Q = rows [0,64), scaled by 1/sqrt(d)
            gemm_frag<8, 4>(acc, xn, sm.wqkv, LDW, lane);   // the caller folded log2(e) / sqrt(d) into these rows: S comes out in log2 units
            acc_to_frag<4>(qf, acc);   // K steps 0,1 = head 0 (d 0..31), 2,3 = head 1
        }
        group_barrier(grp);

        // ---- E. attention per head: S = q k^T (16 x 64 keys), softmax over the 49 window tokens, O = P V
        uint32_t af[4][4];   // attention output as A fragments of out_proj (K step 2h + i = channels 32h + 16i ..)
#pragma unroll
        for (int h = 0; h < HEADS; ++h) {
            float s[8][4];
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
                uint32_t kb[4];
                ldsm_b2(kb, &sg.ks[(h * 64 + nb * 8) * LDK], LDK, lane);
                mma16816(s[nb], qf[2 * h], kb[0], kb[1]);
                mma16816(s[nb], qf[2 * h + 1], kb[2], kb[3]);
            }
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const int key = nb * 8 + 2 * t;
                if (key >= T) { s[nb][0] = -INFINITY; s[nb][2] = -INFINITY; }
                if (key + 1 >= T) { s[nb][1] = -INFINITY; s[nb][3] = -INFINITY; }
                mx0 = fmaxf(mx0, fmaxf(s[nb][0], s[nb][1]));
                mx1 = fmaxf(mx1, fmaxf(s[nb][2], s[nb][3]));
            }
            mx0 = quad_max(mx0); mx1 = quad_max(mx1);
            float l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                s[nb][0] = ex2_approx(s[nb][0] - mx0); s[nb][1] = ex2_approx(s[nb][1] - mx0);
                s[nb][2] = ex2_approx(s[nb][2] - mx1); s[nb][3] = ex2_approx(s[nb][3] - mx1);
                l0 += s[nb][0] + s[nb][1];
                l1 += s[nb][2] + s[nb][3];
            }
            const float inv0 = 1.0f / quad_sum(l0), inv1 = 1.0f / quad_sum(l1);
            uint32_t pfr[4][4];
            acc_to_frag<4>(pfr, s);
            float o[4][4];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
#pragma unroll
                for (int i2 = 0; i2 < 2; ++i2) {
                    uint32_t vb[4];
                    ldsm_b2(vb, &sg.vt[(h * HD + nb * 8) * LDV + i2 * 32], LDV, lane);
                    mma16816(o[nb], pfr[2 * i2], vb[0], vb[1]);
                    mma16816(o[nb], pfr[2 * i2 + 1], vb[2], vb[3]);
                }
                o[nb][0] *= inv0; o[nb][1] *= inv0; o[nb][2] *= inv1; o[nb][3] *= inv1;
            }
            acc_to_frag<2>(&af[2 * h], o);
        }

        // ---- F. out_proj + residual (raw tokens)
        {
            float acc[8][4];
            bias_init(acc, sm.pf + F_BO, t);
            gemm_frag<8, 4>(acc, af, sm.wo, LDW, lane);
#pragma unroll
            for (int nb = 0; nb < 8; ++nb)
#pragma unroll
                for (int e = 0; e < 4; ++e) x1[nb][e] += acc[nb][e];
        }
        // ---- G/H. LayerNorm 2 -> Linear(64,128) -> exact GELU -> Linear(128,64) + residual
        {
            uint32_t xn2[4][4];
            layernorm_frag(xn2, x1);
            uint32_t hf[8][4];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float acc[8][4];
                bias_init(acc, sm.pf + F_B1 + half * 64, t);
                gemm_frag<8, 4>(acc, xn2, sm.w1 + half * 64 * LDW, LDW, lane);
#pragma unroll
                for (int nb = 0; nb < 8; ++nb)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[nb][e] = ysod_gelu_tanh(acc[nb][e]);
                acc_to_frag<4>(&hf[half * 4], acc);
            }
            float acc[8][4];
            bias_init(acc, sm.pf + F_B2, t);
            gemm_frag<8, 8>(acc, hf, sm.w2, LDW2, lane);
#pragma unroll
            for (int nb = 0; nb < 8; ++nb)
#pragma unroll
                for (int e = 0; e < 4; ++e) x1[nb][e] += acc[nb][e];
        }
        // ---- I. pw 1x1 (BN folded) -> SiLU -> + identity -> staging tile -> coalesced store of the in-image pixels
        {
            uint32_t xf[4][4];
            acc_to_frag<4>(xf, x1);
            float acc[8][4];
            bias_init(acc, sm.pf + F_BPW, t);
            gemm_frag<8, 4>(acc, xf, sm.wpw, LDW, lane);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int tk = row0 + rr * 8;
                if (tk < T) {
                    const int i = tk / WS, j = tk - i * WS;
                    const __nv_bfloat16* idp = &sg.patch[((i + 1) * 9 + j + 1) * LDW + 2 * t];
#pragma unroll
                    for (int nb = 0; nb < 8; ++nb) {
                        const float2 id = unpack2(ld32(idp + nb * 8));
                        float v0 = acc[nb][2 * rr], v1 = acc[nb][2 * rr + 1];
                        v0 = silu_tanh(v0) + id.x;
                        v1 = silu_tanh(v1) + id.y;
                        *reinterpret_cast<uint32_t*>(&sg.tok[tk * LDW + nb * 8 + 2 * t]) = pack2(v0, v1);
                    }
                }
            }
        }
        group_barrier(grp);
        for (int i = gt; i < T * 8; i += 128) {
            const int tk = i >> 3, pc = i & 7;
            const int ti = tk / WS, tj = tk - ti * WS;
            const int oh = h0 + ti, ow = w0 + tj;
            if (oh < H && ow < W)
                *reinterpret_cast<uint4*>(out + (((size_t)n * H + oh) * W + ow) * ocs + pc * 8) = *reinterpret_cast<const uint4*>(&sg.tok[tk * LDW + pc * 8]);
        }
        group_barrier(grp);   // tok / patch / K / V are rewritten by the next window
    }
}

}  // namespace

// x / out: NHWC bf16 views (pixel strides xcs / ocs, multiples of 8, 16 B aligned) with 64 channels.
// The caller pre-folds what is linear (engine.py does): norm1's gamma / beta into in_proj (W diag(gamma), b + W beta), norm2's into
// mlp.0, and log2(e) / sqrt(head_dim) into the Q rows of in_proj (weight and bias) -- the kernel normalises without an affine part
// and exponentiates with ex2. The norm slots of pf32 are ignored.
// wbf16: dw[3][3][64] | in_proj_weight[192][64] | out_proj.weight[64][64] | mlp.0.weight[128][64] | mlp.2.weight[64][128] |
//        pw.weight (BN folded)[64][64]                                                                   (37440 bf16)
// pf32:  norm1.weight | norm1.bias | in_proj_bias[192] | out_proj.bias | norm2.weight | norm2.bias | mlp.0.bias[128] |
//        mlp.2.bias | folded BN bias                                                                     (768 fp32)
extern "C" int ysod_swin64_fused(const void* x, int N, int H, int W, int xcs, const void* wbf16, const float* pf32, void* out, int ocs,
                                 int window, int heads, cudaStream_t stream) {
    YSOD_CHECK_ARG(x && wbf16 && pf32 && out, "ysod_swin64_fused: null pointer");
    YSOD_CHECK_ARG(window == WS && heads == HEADS, "ysod_swin64_fused: only 7x7 windows / 2 heads (C = 64) are fused");
    YSOD_CHECK_ARG(H > WS && W > WS, "ysod_swin64_fused: map smaller than a window");
    YSOD_CHECK_ARG(xcs % 8 == 0 && ocs % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)wbf16 % 16) == 0,
                   "ysod_swin64_fused: views must be 16 B aligned");
    const int nWh = ysod_cdiv(H, WS), nWw = ysod_cdiv(W, WS);
    int dev = 0, sms = 148;
    YSOD_CUDA(cudaGetDevice(&dev));
    YSOD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long nwin = (long long)N * nWh * nWw;
    long long grid = (nwin + NGROUP - 1) / NGROUP;
    if (grid > sms) grid = sms;
    const size_t smem = sizeof(Smem);
    YSOD_CUDA(cudaFuncSetAttribute(swin64_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ysod_launch(swin64_fused_kernel, (unsigned)grid, THREADS, smem, stream, (const __nv_bfloat16*)x, N, H, W, xcs, (const __nv_bfloat16*)wbf16, pf32,
                                                                   (__nv_bfloat16*)out, ocs, nWh, nWw);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}
