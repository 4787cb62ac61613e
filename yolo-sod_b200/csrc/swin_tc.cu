// Fully fused SwinBlock for the P2 level (C = 64, 2 heads of 32, 7x7 windows) with EVERY contraction on the 5th-gen tensor cores:
// tcgen05.mma from shared-memory operands into TMEM accumulators. Same arithmetic and rounding points as swin_fused.cu (the mma.sync
// version, kept as the A/B baseline: ysod_swin64_fused).
//
// Replaces (reference): ultralytics/nn/modules/blocks_transformer.py:133-171 SwinBlock.forward
//     dw 3x3 (no bias) -> zero-pad to multiples of 7 -> window_partition (:8-47) -> WindowAttention (:81-131:
//     x + MHA(LN(x)), x + MLP(LN(x)) with Linear -> GELU -> Linear) -> window_reverse / crop (:49-79) -> pw 1x1 -> BN -> SiLU -> + identity
//
// Tile = TWO windows = 128 MMA rows (rows 0..63 / 64..127, 49 real tokens each): thread r of a 128-thread group <-> token row r <->
// TMEM lane r, so LayerNorm / softmax / GELU are per-thread loops over the thread's own row (no shuffles), read from TMEM with
// tcgen05.ld and written back to shared memory as the next GEMM's A operand (K-major, 128 B swizzled rows).
//   per tile:  patch -> dw 3x3 + LN1 -> [QKV 128x192x64] -> Q,K,V tiles -> [S_h = Q_h K_h^T 128x128x32, both heads] -> softmax (own
//              window's 49 keys; the cross-window quadrants of P are zero) -> [O_h = P_h V_h 128x32x128, V as an MN-major operand]
//              -> [out_proj 128x64x64] + residual -> LN2 -> [MLP1 128x128x64] -> GELU -> [MLP2 128x64x128] + residual
//              -> [pw 128x64x64] -> SiLU + identity -> global.        [..] = tcgen05.mma groups, 28 MMAs per tile.
// One persistent CTA per SM holds the 72 KB of weights in the UMMA operand layout; TWO independent 128-thread groups work on
// different tiles with their own operand buffers (64 KB), mbarrier and 256 TMEM columns, so one group's TMEM reads / element-wise
// math overlap the other's MMAs. The accumulators never round-trip through registers between the GEMM and its consumer more than
// the algorithm requires (every GEMM here is followed by a row-wise non-linearity).
#include "umma.cuh"

namespace {
using namespace umma;

constexpr int WS = 7, T = 49;
// bf16 weight blob (global): dw[9][64] | wqkv[192][64] | wo[64][64] | w1[128][64] | w2[64][128] | wpw[64][64]   (as ysod_swin64_fused)
constexpr int G_DW = 0, G_QKV = 576, G_WO = G_QKV + 192 * 64, G_W1 = G_WO + 64 * 64, G_W2 = G_W1 + 128 * 64, G_PW = G_W2 + 64 * 128;
// fp32 blob (global): ln1_g | ln1_b | bqkv[192] | bo | ln2_g | ln2_b | b1[128] | b2 | bpw
constexpr int F_BQKV = 128, F_BO = 320, F_B1 = 512, F_B2 = 640, F_BPW = 704;

// shared memory (bytes from a 1 KB aligned base)
constexpr uint32_t S_WQKV = 0, S_WO = 24576, S_W1 = 32768, S_W2 = 49152, S_WPW = 65536, S_GRP = 73728;
constexpr uint32_t GRP_BYTES = 65536;                       // per group: A1 16 KB | QK 32 KB | V 16 KB
constexpr uint32_t G_A1 = 0, G_QK = 16384, G_V = 49152;
//   A1: the 128 x 64 A operand of QKV / out_proj / MLP1 / pw (LN1 output, attention output, LN2 output, x2 in turn)
//   QK: Q_h0 | Q_h1 | K_h0 | K_h1 (8 KB each, 64 B rows) -- before that the two 9x9 input patches, after S the P tile (2 x 16 KB),
//       after PV the MLP hidden tile (2 x 16 KB)
//   V : V_h0 | V_h1 (8 KB each, 64 B rows, row = key: an MN-major B operand)
constexpr uint32_t PATCH_BYTES = 81 * 128;
constexpr uint32_t S_PF = S_GRP + 2 * GRP_BYTES;
// fp32 parameters in shared memory: dw[9][64] | bqkv[192] | bo[64] | b1[128] | b2[64] | bpw[64]
constexpr int P_DW = 0, P_BQKV = 576, P_BO = 768, P_B1 = 832, P_B2 = 960, P_BPW = 1024, P_TOTAL = 1088;
constexpr uint32_t S_BAR = S_PF + P_TOTAL * 4;              // two mbarriers, TMEM slot
constexpr uint32_t SMEM_BYTES = S_BAR + 64 + 1024;

__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float silu_tanh(float x) {   // x * sigmoid(x) = h + h * tanh(h), h = x / 2
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
__device__ __forceinline__ void group_barrier(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

// 16 B chunk c of row r of a K-major tile with 128 B rows (SWIZZLE_128B) / 64 B rows (SWIZZLE_64B); tiles are 1 KB aligned
__device__ __forceinline__ uint32_t row128(uint32_t tile, int r, int c) { return tile + (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t row64(uint32_t tile, int r, int c) { return tile + (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

// rows x 64 bf16 weights (row stride ld elements) -> K-major SWIZZLE_128B tile
__device__ __forceinline__ void stage_weights(uint32_t dst, const __nv_bfloat16* __restrict__ src, int rows, int ld, int tid) {
    for (int i = tid; i < rows * 8; i += 256) {
        const int row = i >> 3, c = i & 7;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)row * ld + c * 8));
        st_shared_v4(row128(dst, row, c), v.x, v.y, v.z, v.w);
    }
}

// this thread's 64 fp32 values -> bf16 -> row r of a 128 B-row tile
__device__ __forceinline__ void store_row64(uint32_t tile, int r, const float* v) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
        st_shared_v4(row128(tile, r, c), pack2(v[8 * c], v[8 * c + 1]), pack2(v[8 * c + 2], v[8 * c + 3]), pack2(v[8 * c + 4], v[8 * c + 5]),
                     pack2(v[8 * c + 6], v[8 * c + 7]));
}
// LayerNorm over the thread's 64 values without the affine part (folded into the next linear layer by the caller) -> bf16 row
__device__ __forceinline__ void layernorm_store(uint32_t tile, int r, const float* v) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 64; ++k) s += v[k];
    const float mean = s * (1.0f / 64.0f);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 64; ++k) { const float d = v[k] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(q * (1.0f / 64.0f) + 1e-5f), c0 = -mean * rstd;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = fmaf(v[8 * c + k], rstd, c0);
        st_shared_v4(row128(tile, r, c), pack2(y[0], y[1]), pack2(y[2], y[3]), pack2(y[4], y[5]), pack2(y[6], y[7]));
    }
}
// 16 accumulator columns of this thread's TMEM lane + 16 fp32 biases from shared memory
__device__ __forceinline__ void load_acc16(uint32_t taddr, uint32_t bias, float* f) {
    uint32_t v[16];
    tmem_ld16(taddr, v);
    const float4 b0 = lds_f4(bias), b1 = lds_f4(bias + 16u), b2 = lds_f4(bias + 32u), b3 = lds_f4(bias + 48u);
    tmem_ld_wait(v);
    const float bb[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + bb[j];
}

__global__ void __launch_bounds__(256, 1)
swin64_tc_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int xcs, const __nv_bfloat16* __restrict__ wb,
                 const float* __restrict__ pf, __nv_bfloat16* __restrict__ out, int ocs, int nWh, int nWw) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = smem_u32(smem_raw);
    const uint32_t base = (smem0 + 1023u) & ~1023u;
    float* const prm = reinterpret_cast<float*>(smem_raw + (base - smem0) + S_PF);
    const uint32_t prm_s = base + S_PF;
    const int tid = threadIdx.x, warp = tid >> 5, grp = tid >> 7, r = tid & 127;
    const uint32_t gb = base + S_GRP + (uint32_t)grp * GRP_BYTES;
    const uint32_t a1 = gb + G_A1, qk = gb + G_QK, vv = gb + G_V;
    const uint32_t bar = base + S_BAR + 8u * (uint32_t)grp, tmem_slot = base + S_BAR + 16u;

    // ---- weights (static) -> UMMA operand tiles, parameters -> fp32, before the dependency wait
    stage_weights(base + S_WQKV, wb + G_QKV, 192, 64, tid);
    stage_weights(base + S_WO, wb + G_WO, 64, 64, tid);
    stage_weights(base + S_W1, wb + G_W1, 128, 64, tid);
    stage_weights(base + S_W2, wb + G_W2, 64, 128, tid);               // K block 0: input channels 0..63
    stage_weights(base + S_W2 + 8192u, wb + G_W2 + 64, 64, 128, tid);  // K block 1: 64..127
    stage_weights(base + S_WPW, wb + G_PW, 64, 64, tid);
    for (int i = tid; i < 576; i += 256) prm[P_DW + i] = __bfloat162float(wb[G_DW + i]);
    for (int i = tid; i < 192; i += 256) prm[P_BQKV + i] = pf[F_BQKV + i];
    for (int i = tid; i < 128; i += 256) prm[P_B1 + i] = pf[F_B1 + i];
    if (tid < 64) { prm[P_BO + tid] = pf[F_BO + tid]; prm[P_B2 + tid] = pf[F_B2 + tid]; prm[P_BPW + tid] = pf[F_BPW + tid]; }
    if (tid == 0) {
        mbar_init(base + S_BAR, 1);
        mbar_init(base + S_BAR + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");
    const uint32_t tcol = tmem + (uint32_t)grp * 256u;                               // this group's 256 accumulator columns
    const uint32_t trow = tcol + ((uint32_t)((warp & 3) * 32) << 16);                // ... seen from this warp's lane quarter
    const bool issuer_warp = (warp & 3) == 0;
    ysod_pdl_sync();   // the input map is written by the previous kernel

    constexpr uint32_t ID_QKV = idesc_f16(128, 192, 0, 0), ID_N128 = idesc_f16(128, 128, 0, 0), ID_N64 = idesc_f16(128, 64, 0, 0),
                       ID_PV = idesc_f16(128, 32, 0, 1);
    const int wsel = r >> 6, t = r & 63;                 // this row's window of the pair, token inside the window
    const int ti = t / WS, tj = t - ti * WS;
    const long long nwin = (long long)N * nWh * nWw, npairs = (nwin + 1) / 2;
    uint32_t phase = 0;

    for (long long pr = (long long)blockIdx.x * 2 + grp; pr < npairs; pr += (long long)gridDim.x * 2) {
        // the pair's two windows (the second may not exist): image, origin
        const long long wa = 2 * pr, wb2 = 2 * pr + 1;
        const bool ok_a = true, ok_b = wb2 < nwin;
        const long long wcb = ok_b ? wb2 : wa;
        const int w0_a = (int)(wa % nWw) * WS, h0_a = (int)((wa / nWw) % nWh) * WS, n_a = (int)(wa / ((long long)nWw * nWh));
        const int w0_b = (int)(wcb % nWw) * WS, h0_b = (int)((wcb / nWw) % nWh) * WS, n_b = (int)(wcb / ((long long)nWw * nWh));
        // ---- A. the two 9 x 9 input patches (64 ch, zero outside the image) -> the QK region; all loads before the first store
        {
            constexpr int NCH = 2 * 81 * 8, NIT = (NCH + 127) / 128;
            uint4 pv[NIT];
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int i = r + it * 128;
                const int s = i >= 648 ? 1 : 0, rem = i - s * 648;
                const int pp = rem >> 3, pc = rem & 7;
                const int prow = pp / 9, pcol = pp - prow * 9;
                const int ih = (s ? h0_b : h0_a) - 1 + prow, iw = (s ? w0_b : w0_a) - 1 + pcol;
                pv[it] = make_uint4(0, 0, 0, 0);
                if (i < NCH && (s ? ok_b : ok_a) && ih >= 0 && ih < H && iw >= 0 && iw < W)
                    pv[it] = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)(s ? n_b : n_a) * H + ih) * W + iw) * xcs + pc * 8));
            }
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int i = r + it * 128;
                if (i < NCH) {
                    const int s = i >= 648 ? 1 : 0, rem = i - s * 648;
                    const int pp = rem >> 3, pc = rem & 7;
                    st_shared_v4(row128(qk + (uint32_t)s * PATCH_BYTES, pp, pc), pv[it].x, pv[it].y, pv[it].z, pv[it].w);
                }
            }
        }
        group_barrier(grp);

        // ---- B. depthwise 3x3 (taps accumulated in (r, s) order, fp32) -> bf16 token; tokens beyond the image are window_partition's
        //         zero padding; MMA padding rows 49..63 are zero. Then LayerNorm 1 -> A1.
        const int oh = (wsel ? h0_b : h0_a) + ti, ow = (wsel ? w0_b : w0_a) + tj, on = wsel ? n_b : n_a;
        const bool in_img = t < T && (wsel ? ok_b : ok_a) && oh < H && ow < W;
        float x1[64];   // running token value: dw output, + attention, + MLP
#pragma unroll
        for (int k = 0; k < 64; ++k) x1[k] = 0.f;
        if (in_img) {
            const uint32_t pbase = qk + (uint32_t)wsel * PATCH_BYTES;
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
                const int dr = tap / 3, ds = tap - dr * 3;
                const int pp = (ti + dr) * 9 + tj + ds;
                const uint32_t wt = prm_s + 4u * (uint32_t)(P_DW + tap * 64);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 v = lds_v4(row128(pbase, pp, c));
                    const float4 w0 = lds_f4(wt + (uint32_t)(c * 32)), w1 = lds_f4(wt + (uint32_t)(c * 32 + 16));
                    const float2 p0 = ysod_unpack2(v.x), p1 = ysod_unpack2(v.y), p2 = ysod_unpack2(v.z), p3 = ysod_unpack2(v.w);
                    x1[8 * c + 0] = fmaf(p0.x, w0.x, x1[8 * c + 0]); x1[8 * c + 1] = fmaf(p0.y, w0.y, x1[8 * c + 1]);
                    x1[8 * c + 2] = fmaf(p1.x, w0.z, x1[8 * c + 2]); x1[8 * c + 3] = fmaf(p1.y, w0.w, x1[8 * c + 3]);
                    x1[8 * c + 4] = fmaf(p2.x, w1.x, x1[8 * c + 4]); x1[8 * c + 5] = fmaf(p2.y, w1.y, x1[8 * c + 5]);
                    x1[8 * c + 6] = fmaf(p3.x, w1.z, x1[8 * c + 6]); x1[8 * c + 7] = fmaf(p3.y, w1.w, x1[8 * c + 7]);
                }
            }
#pragma unroll
            for (int k = 0; k < 64; ++k) x1[k] = __bfloat162float(__float2bfloat16_rn(x1[k]));   // the token as the reference's 16-bit dw output
        }
        layernorm_store(a1, r, x1);
        tc_fence_before();
        fence_async_smem();
        group_barrier(grp);     // also: every thread is done with the patches (QK region)
        if (issuer_warp && elect_one()) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                tc_mma(tcol, smem_desc(a1 + ks * 32u, 16u, 1024u, 2u), smem_desc(base + S_WQKV + ks * 32u, 16u, 1024u, 2u), ID_QKV, (uint32_t)(ks > 0));
            tc_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();

        // ---- C. in_proj epilogue: + bias -> bf16 -> per-head Q / K / V tiles (64 B rows). (Q rows carry log2(e) / sqrt(d): caller.)
#pragma unroll
        for (int ch = 0; ch < 12; ++ch) {
            float f[16];
            load_acc16(trow + (uint32_t)(ch * 16), prm_s + 4u * (uint32_t)(P_BQKV + ch * 16), f);
            const int sel = ch >> 2, hh = (ch >> 1) & 1, cp = (ch & 1) * 2;
            const uint32_t tile = (sel == 0 ? qk : sel == 1 ? qk + 16384u : vv) + (uint32_t)hh * 8192u;
            st_shared_v4(row64(tile, r, cp), pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
            st_shared_v4(row64(tile, r, cp + 1), pack2(f[8], f[9]), pack2(f[10], f[11]), pack2(f[12], f[13]), pack2(f[14], f[15]));
        }
        tc_fence_before();
        fence_async_smem();
        group_barrier(grp);
        if (issuer_warp && elect_one()) {
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)   // S_h[128 q][128 keys] = Q_h K_h^T, head_dim 32 = two K steps
                    tc_mma(tcol + (uint32_t)h * 128u, smem_desc(qk + (uint32_t)h * 8192u + ks * 32u, 16u, 512u, 4u),
                           smem_desc(qk + 16384u + (uint32_t)h * 8192u + ks * 32u, 16u, 512u, 4u), ID_N128, (uint32_t)(ks > 0));
            tc_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();

        // ---- D. softmax over the 49 keys of the row's own window (S is in log2 units), P (bf16, unnormalised) -> the QK region as two
        //         K-major chunks of 64 keys; the other window's chunk of this row is zero. O_h = P V_h.
        float inv[2];
        uint32_t pk[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float s[64];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[16];
                tmem_ld16(trow + (uint32_t)(h * 128 + wsel * 64 + c * 16), v);
                tmem_ld_wait(v);
#pragma unroll
                for (int j = 0; j < 16; ++j) s[c * 16 + j] = __uint_as_float(v[j]);
            }
            float mx = s[0];
#pragma unroll
            for (int k = 1; k < T; ++k) mx = fmaxf(mx, s[k]);
            float l = 0.f;
#pragma unroll
            for (int k = 0; k < 64; ++k) {
                s[k] = k < T ? ex2_approx(s[k] - mx) : 0.f;
                l += s[k];
            }
            inv[h] = 1.0f / l;
#pragma unroll
            for (int k = 0; k < 32; ++k) pk[k] = pack2(s[2 * k], s[2 * k + 1]);
            if (h == 1) {   // P_0 must have been consumed (O_0 complete) before P_1 overwrites it
                mbar_wait(bar, phase); phase ^= 1u;
                tc_fence_after();
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                st_shared_v4(row128(qk + (uint32_t)wsel * 16384u, r, c), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                if (h == 0) st_shared_v4(row128(qk + (uint32_t)(1 - wsel) * 16384u, r, c), 0u, 0u, 0u, 0u);
            }
            tc_fence_before();
            fence_async_smem();
            group_barrier(grp);
            if (issuer_warp && elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)   // 16 keys per MMA: A advances 32 B inside a 64-key chunk, B (MN-major V_h) advances 16 key rows
                    tc_mma(tcol + (uint32_t)h * 32u, smem_desc(qk + (uint32_t)(ks >> 2) * 16384u + (uint32_t)(ks & 3) * 32u, 16u, 1024u, 2u),
                           smem_desc(vv + (uint32_t)h * 8192u + (uint32_t)ks * 1024u, 512u, 512u, 4u), ID_PV, (uint32_t)(ks > 0));
                tc_commit(bar);
            }
        }
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();

        // ---- E. attention output (normalised) -> A1 -> out_proj
        {
            float o[64];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[16];
                tmem_ld16(trow + (uint32_t)(c * 16), v);
                tmem_ld_wait(v);
#pragma unroll
                for (int j = 0; j < 16; ++j) o[c * 16 + j] = __uint_as_float(v[j]) * inv[c >> 1];
            }
            store_row64(a1, r, o);
        }
        tc_fence_before();
        fence_async_smem();
        group_barrier(grp);
        if (issuer_warp && elect_one()) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                tc_mma(tcol + 64u, smem_desc(a1 + ks * 32u, 16u, 1024u, 2u), smem_desc(base + S_WO + ks * 32u, 16u, 1024u, 2u), ID_N64, (uint32_t)(ks > 0));
            tc_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();

        // ---- F. + bias + residual -> LayerNorm 2 -> A1 -> MLP linear 1
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float f[16];
            load_acc16(trow + 64u + (uint32_t)(c * 16), prm_s + 4u * (uint32_t)(P_BO + c * 16), f);
#pragma unroll
            for (int j = 0; j < 16; ++j) x1[c * 16 + j] += f[j];
        }
        layernorm_store(a1, r, x1);
        tc_fence_before();
        fence_async_smem();
        group_barrier(grp);
        if (issuer_warp && elect_one()) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                tc_mma(tcol + 128u, smem_desc(a1 + ks * 32u, 16u, 1024u, 2u), smem_desc(base + S_W1 + ks * 32u, 16u, 1024u, 2u), ID_N128, (uint32_t)(ks > 0));
            tc_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();

        // ---- G. + bias -> GELU -> bf16 hidden tile (two K chunks of 64 in the QK region) -> MLP linear 2
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float f[16];
            load_acc16(trow + 128u + (uint32_t)(c * 16), prm_s + 4u * (uint32_t)(P_B1 + c * 16), f);
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = ysod_gelu_tanh(f[j]);
            const uint32_t tile = qk + (uint32_t)(c >> 2) * 16384u;
            const int c2 = (c & 3) * 2;
            st_shared_v4(row128(tile, r, c2), pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
            st_shared_v4(row128(tile, r, c2 + 1), pack2(f[8], f[9]), pack2(f[10], f[11]), pack2(f[12], f[13]), pack2(f[14], f[15]));
        }
        tc_fence_before();
        fence_async_smem();
        group_barrier(grp);
        if (issuer_warp && elect_one()) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
                tc_mma(tcol, smem_desc(qk + (uint32_t)(ks >> 2) * 16384u + (uint32_t)(ks & 3) * 32u, 16u, 1024u, 2u),
                       smem_desc(base + S_W2 + (uint32_t)(ks >> 2) * 8192u + (uint32_t)(ks & 3) * 32u, 16u, 1024u, 2u), ID_N64, (uint32_t)(ks > 0));
            tc_commit(bar);
        }
        // the identity row travels while the last two GEMMs run
        uint4 idv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) idv[c] = make_uint4(0, 0, 0, 0);
        if (in_img) {
            const __nv_bfloat16* ip = x + (((size_t)on * H + oh) * W + ow) * xcs;
#pragma unroll
            for (int c = 0; c < 8; ++c) idv[c] = __ldg(reinterpret_cast<const uint4*>(ip + c * 8));
        }
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();

        // ---- H. + bias + residual -> bf16 -> A1 -> pw 1x1 (BN folded)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float f[16];
            load_acc16(trow + (uint32_t)(c * 16), prm_s + 4u * (uint32_t)(P_B2 + c * 16), f);
#pragma unroll
            for (int j = 0; j < 16; ++j) x1[c * 16 + j] += f[j];
        }
        store_row64(a1, r, x1);
        tc_fence_before();
        fence_async_smem();
        group_barrier(grp);
        if (issuer_warp && elect_one()) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                tc_mma(tcol + 64u, smem_desc(a1 + ks * 32u, 16u, 1024u, 2u), smem_desc(base + S_WPW + ks * 32u, 16u, 1024u, 2u), ID_N64, (uint32_t)(ks > 0));
            tc_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();

        // ---- I. + bias -> SiLU -> + identity -> the pixel's 128 B row in global memory
        {
            __nv_bfloat16* op = out + (((size_t)on * H + oh) * W + ow) * ocs;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float f[16];
                load_acc16(trow + 64u + (uint32_t)(c * 16), prm_s + 4u * (uint32_t)(P_BPW + c * 16), f);
                if (in_img) {
                    const uint32_t iw[8] = {idv[2 * c].x, idv[2 * c].y, idv[2 * c].z, idv[2 * c].w, idv[2 * c + 1].x, idv[2 * c + 1].y, idv[2 * c + 1].z, idv[2 * c + 1].w};
                    uint32_t ov[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float2 id = ysod_unpack2(iw[j]);
                        ov[j] = pack2(silu_tanh(f[2 * j]) + id.x, silu_tanh(f[2 * j + 1]) + id.y);
                    }
                    *reinterpret_cast<uint4*>(op + c * 16) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
                    *reinterpret_cast<uint4*>(op + c * 16 + 8) = make_uint4(ov[4], ov[5], ov[6], ov[7]);
                }
            }
        }
        tc_fence_before();   // this tile's TMEM reads are ordered before the barriers of the next tile's first MMA group
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

}  // namespace

// Same contract as ysod_swin64_fused (swin_fused.cu): x / out NHWC 16-bit views with 64 channels, the caller pre-folds the LayerNorm
// affine parts into in_proj / mlp.0 and log2(e) / sqrt(head_dim) into the Q rows; wbf16 / pf32 are the same blobs.
extern "C" int ysod_swin64_tc(const void* x, int N, int H, int W, int xcs, const void* wbf16, const float* pf32, void* out, int ocs,
                              int window, int heads, cudaStream_t stream) {
    YSOD_CHECK_ARG(x && wbf16 && pf32 && out, "ysod_swin64_tc: null pointer");
    YSOD_CHECK_ARG(window == WS && heads == 2, "ysod_swin64_tc: only 7x7 windows / 2 heads (C = 64) are fused");
    YSOD_CHECK_ARG(H > WS && W > WS, "ysod_swin64_tc: map smaller than a window");
    YSOD_CHECK_ARG(xcs % 8 == 0 && ocs % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)wbf16 % 16) == 0,
                   "ysod_swin64_tc: views must be 16 B aligned");
    const int nWh = ysod_cdiv(H, WS), nWw = ysod_cdiv(W, WS);
    int dev = 0, sms = 148;
    YSOD_CUDA(cudaGetDevice(&dev));
    YSOD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long npairs = ((long long)N * nWh * nWw + 1) / 2;
    long long grid = (npairs + 1) / 2;
    if (grid > sms) grid = sms;
    YSOD_CUDA(cudaFuncSetAttribute(swin64_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    ysod_launch(swin64_tc_kernel, (unsigned)grid, 256, SMEM_BYTES, stream, (const __nv_bfloat16*)x, N, H, W, xcs, (const __nv_bfloat16*)wbf16, pf32,
                (__nv_bfloat16*)out, ocs, nWh, nWw);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}
