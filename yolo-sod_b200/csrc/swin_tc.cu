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
#include <cuda.h>

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
constexpr uint32_t PATCH_BYTES = 81 * 128, PATCH_STRIDE = 11 * 1024;   // the second patch starts on a swizzle-atom boundary
constexpr uint32_t S_PF = S_GRP + 2 * GRP_BYTES;
// fp32 parameters in shared memory: dw[9][64] | bqkv[192] | bo[64] | b1[128] | b2[64] | bpw[64]
constexpr int P_DW = 0, P_BQKV = 576, P_BO = 768, P_B1 = 832, P_B2 = 960, P_BPW = 1024, P_TOTAL = 1088;
constexpr uint32_t S_BAR = S_PF + P_TOTAL * 4;              // MMA mbarrier per group (2 x 8 B) | TMEM slot (@16) | patch mbarrier per group (@24, @32)
constexpr uint32_t SMEM_BYTES = S_BAR + 64 + 1024;

__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
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
__device__ __forceinline__ float tanh_approx(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
}
__device__ __forceinline__ void group_barrier(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// ---- packed fp32 pairs (sm_100 add / mul / fma .f32x2: FADD2 / FMUL2 / FFMA2, two lanes per issue slot). Every row-wise stage of this
// kernel is a long stream of independent fp32 operations issued by 2 warps per scheduler, so the instruction count is what bounds it.
typedef unsigned long long p2;
__device__ __forceinline__ p2 pk2(float lo, float hi) { p2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ p2 pk2u(uint32_t lo, uint32_t hi) { p2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ float2 up2(p2 v) { float2 f; asm("mov.b64 {%0, %1}, %2;" : "=f"(f.x), "=f"(f.y) : "l"(v)); return f; }
__device__ __forceinline__ p2 add2(p2 a, p2 b) { p2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ p2 mul2(p2 a, p2 b) { p2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ p2 fma2(p2 a, p2 b, p2 c) { p2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ p2 unpack_p2(uint32_t v) { const float2 f = ysod_unpack2(v); return pk2(f.x, f.y); }   // two 16-bit activations
__device__ __forceinline__ uint32_t pack_p2(p2 v) { const float2 f = up2(v); return pack2(f.x, f.y); }
// 4 consecutive fp32 pairs (32 B) from shared memory
__device__ __forceinline__ void lds_p2x4(uint32_t addr, p2* v) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v[0]), "=l"(v[1]) : "r"(addr) : "memory");
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v[2]), "=l"(v[3]) : "r"(addr + 16u) : "memory");
}

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

// 32 accumulator columns of this thread's TMEM lane (asynchronous: tmem_wait32 before the registers are read)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
}
// waits for every outstanding tcgen05.ld of the thread; the "+r" operands tie the loaded registers to the wait
__device__ __forceinline__ void tmem_wait32(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                   "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                   "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}
// Streams n32 chunks of 32 accumulator columns starting at taddr through `body(i, v)`: chunk i + 1 is in flight while chunk i is processed.
template <int N32, typename F>
__device__ __forceinline__ void for_acc32(uint32_t taddr, F&& body) {
    uint32_t va[32], vb[32];
    tmem_ld32(taddr, va);
#pragma unroll
    for (int i = 0; i < N32; ++i) {
        uint32_t* cur = (i & 1) ? vb : va;
        uint32_t* nxt = (i & 1) ? va : vb;
        tmem_wait32(cur);
        if (i + 1 < N32) tmem_ld32(taddr + (uint32_t)((i + 1) * 32), nxt);
        body(i, cur);
    }
}
// v[2j], v[2j+1] (accumulator columns) + bias pair j from shared memory, 16 pairs
__device__ __forceinline__ void acc_bias32(const uint32_t* v, uint32_t bias, p2* f) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        p2 b[4];
        lds_p2x4(bias + (uint32_t)(q * 32), b);
#pragma unroll
        for (int j = 0; j < 4; ++j) f[4 * q + j] = add2(pk2u(v[8 * q + 2 * j], v[8 * q + 2 * j + 1]), b[j]);
    }
}
// the thread's 32 pairs -> bf16 -> row r of a 128 B-row tile
__device__ __forceinline__ void store_row64(uint32_t tile, int r, const p2* v) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
        st_shared_v4(row128(tile, r, c), pack_p2(v[4 * c]), pack_p2(v[4 * c + 1]), pack_p2(v[4 * c + 2]), pack_p2(v[4 * c + 3]));
}
// LayerNorm over the thread's 64 values without the affine part (folded into the next linear layer by the caller) -> bf16 row
__device__ __forceinline__ void layernorm_store(uint32_t tile, int r, const p2* v) {
    p2 sa = v[0], sb = v[1];
#pragma unroll
    for (int k = 2; k < 32; k += 2) { sa = add2(sa, v[k]); sb = add2(sb, v[k + 1]); }
    const float2 s = up2(add2(sa, sb));
    const float mean = (s.x + s.y) * (1.0f / 64.0f);
    const p2 nm = pk2(-mean, -mean);
    p2 qa = pk2(0.f, 0.f), qb = qa;
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
        const p2 d0 = add2(v[k], nm), d1 = add2(v[k + 1], nm);
        qa = fma2(d0, d0, qa);
        qb = fma2(d1, d1, qb);
    }
    const float2 q = up2(add2(qa, qb));
    const float rstd = rsqrtf((q.x + q.y) * (1.0f / 64.0f) + 1e-5f), c0 = -mean * rstd;
    const p2 rs = pk2(rstd, rstd), cc = pk2(c0, c0);
#pragma unroll
    for (int c = 0; c < 8; ++c)
        st_shared_v4(row128(tile, r, c), pack_p2(fma2(v[4 * c], rs, cc)), pack_p2(fma2(v[4 * c + 1], rs, cc)), pack_p2(fma2(v[4 * c + 2], rs, cc)),
                     pack_p2(fma2(v[4 * c + 3], rs, cc)));
}

// profiling only (ysod_swin64_tc_trace): clock64 at the stage boundaries of the first tiles of CTA 0 / group 0 / thread 0
__device__ long long g_trace[8 * 24];
__device__ int g_trace_on = 0;
#define TR(k) do { if (tr_on && tile_i < 8) g_trace[tile_i * 24 + (k)] = clock64(); } while (0)

__global__ void __launch_bounds__(256, 1)
swin64_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX7, const __grid_constant__ CUtensorMap tmO7,
                 const __nv_bfloat16* __restrict__ x, int N, int H, int W, int xcs, const __nv_bfloat16* __restrict__ wb,
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
        mbar_init(base + S_BAR + 24u, 1);
        mbar_init(base + S_BAR + 32u, 1);
        mbar_init(base + S_BAR + 40u, 1);
        mbar_init(base + S_BAR + 48u, 1);
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
    const uint32_t nwin = (uint32_t)N * (uint32_t)nWh * (uint32_t)nWw, npairs = (nwin + 1u) / 2u, per_img = (uint32_t)nWh * (uint32_t)nWw;   // host: < 2^31
    uint32_t phase = 0;
    // The two 9 x 9 x 64 input patches of a window pair arrive by TMA (one box each, zero fill outside the image = the conv padding) in the
    // QK region, which is idle from the end of MLP linear 2 of the previous tile until this tile's in_proj epilogue.
    const uint32_t pbar = base + S_BAR + 24u + 8u * (uint32_t)grp;
    uint32_t pphase = 0;
    // The identity rows (the block's input at the window's own 7 x 7 pixels) arrive by TMA in the V region once PV is done; the output tile is
    // written over them row by row and leaves by TMA store (out-of-image tokens are clipped by the store, like window_reverse's crop).
    const uint32_t ibar = base + S_BAR + 40u + 8u * (uint32_t)grp;
    uint32_t iphase = 0;
    constexpr uint32_t ID_BYTES = 49u * 128u, ID_STRIDE = 8192u;
    auto request_patches = [&](uint32_t prn) {
        if (prn < npairs) {
            const uint32_t qa = 2u * prn, qb = 2u * prn + 1u;
            const bool okb = qb < nwin;
            const int na = (int)(qa / per_img), rra = (int)(qa - (uint32_t)na * per_img), ha = (rra / nWw) * WS, wa0 = (rra % nWw) * WS;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pbar), "r"(okb ? 2u * PATCH_BYTES : PATCH_BYTES) : "memory");
            tma_load_4d(qk, &tmX, pbar, 0, wa0 - 1, ha - 1, na);
            if (okb) {
                const int nb = (int)(qb / per_img), rrb = (int)(qb - (uint32_t)nb * per_img), hb = (rrb / nWw) * WS, wb0 = (rrb % nWw) * WS;
                tma_load_4d(qk + PATCH_STRIDE, &tmX, pbar, 0, wb0 - 1, hb - 1, nb);
            }
        }
    };
    if (issuer_warp && elect_one()) request_patches(blockIdx.x * 2u + (uint32_t)grp);
    const bool tr_on = g_trace_on && blockIdx.x == 0 && tid == 0;
    int tile_i = -1;

    for (uint32_t pr = blockIdx.x * 2u + (uint32_t)grp; pr < npairs; pr += gridDim.x * 2u) {
        ++tile_i;
        TR(0);
        // the pair's two windows (the second may not exist): image, origin
        const uint32_t wa = 2u * pr, wb2 = 2u * pr + 1u;
        const bool ok_a = true, ok_b = wb2 < nwin;
        const uint32_t wcb = ok_b ? wb2 : wa;
        const int n_a = (int)(wa / per_img), ra = (int)(wa - (uint32_t)n_a * per_img), h0_a = (ra / nWw) * WS, w0_a = (ra % nWw) * WS;
        const int n_b = (int)(wcb / per_img), rb = (int)(wcb - (uint32_t)n_b * per_img), h0_b = (rb / nWw) * WS, w0_b = (rb % nWw) * WS;
        // ---- A. the two 9 x 9 input patches were requested during the previous tile
        mbar_wait(pbar, pphase); pphase ^= 1u;
        TR(1);

        // ---- B. depthwise 3x3 as a sliding window: thread = (window, token column tj, 8-channel chunk) walks the 9 patch rows once, every
        //         loaded pixel feeds the (up to) three output rows it belongs to; the nine filter taps of the chunk stay in registers. Each
        //         output accumulates its taps in (r, s) order in fp32, like the reference's conv. Tokens beyond the image are
        //         window_partition's zero padding. The bf16 tokens go to A1 (row = token), then thread = token row applies LayerNorm 1.
        const int oh = (wsel ? h0_b : h0_a) + ti, ow = (wsel ? w0_b : w0_a) + tj, on = wsel ? n_b : n_a;
        const bool in_img = t < T && (wsel ? ok_b : ok_a) && oh < H && ow < W;
        if (t < 56) {
            const int dwc = t & 7, dwj = t >> 3;
            const uint32_t pbase = qk + (uint32_t)wsel * PATCH_STRIDE;
            p2 wv[9][4], acc[7][4];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) lds_p2x4(prm_s + 4u * (uint32_t)(P_DW + tap * 64 + dwc * 8), wv[tap]);
#pragma unroll
            for (int i = 0; i < 7; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[i][k] = pk2(0.f, 0.f);
#pragma unroll
            for (int prw = 0; prw < 9; ++prw) {
#pragma unroll
                for (int ds = 0; ds < 3; ++ds) {
                    const uint4 v = lds_v4(row128(pbase, prw * 9 + dwj + ds, dwc));
                    const p2 px[4] = {unpack_p2(v.x), unpack_p2(v.y), unpack_p2(v.z), unpack_p2(v.w)};
#pragma unroll
                    for (int dr = 0; dr < 3; ++dr) {
                        const int orow = prw - dr;
                        if (orow >= 0 && orow < 7) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) acc[orow][k] = fma2(px[k], wv[dr * 3 + ds][k], acc[orow][k]);
                        }
                    }
                }
            }
            const bool col_ok = (wsel ? ok_b : ok_a) && (wsel ? w0_b : w0_a) + dwj < W;
            const int hrow0 = wsel ? h0_b : h0_a;
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const bool live = col_ok && hrow0 + i < H;
                st_shared_v4(row128(a1, wsel * 64 + i * 7 + dwj, dwc), live ? pack_p2(acc[i][0]) : 0u, live ? pack_p2(acc[i][1]) : 0u,
                             live ? pack_p2(acc[i][2]) : 0u, live ? pack_p2(acc[i][3]) : 0u);
            }
        }
        group_barrier(grp);     // tokens complete; every thread is done with the patches (QK region)
        TR(2);
        p2 x1[32];   // running token value: dw output (as the reference's 16-bit conv output), + attention, + MLP
#pragma unroll
        for (int k = 0; k < 32; ++k) x1[k] = pk2(0.f, 0.f);
        if (t < T) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 v = lds_v4(row128(a1, r, c));
                x1[4 * c] = unpack_p2(v.x); x1[4 * c + 1] = unpack_p2(v.y); x1[4 * c + 2] = unpack_p2(v.z); x1[4 * c + 3] = unpack_p2(v.w);
            }
        }
        layernorm_store(a1, r, x1);   // in place: a thread reads and writes its own row only (MMA padding rows 49..63 become zero rows)
        if (r == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous tile's output store has read the V region (in_proj's epilogue rewrites it)
        TR(3);
        tc_fence_before();
        fence_async_smem();
        group_barrier(grp);
        if (issuer_warp && elect_one()) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                tc_mma(tcol, smem_desc(a1 + ks * 32u, 16u, 1024u, 2u), smem_desc(base + S_WQKV + ks * 32u, 16u, 1024u, 2u), ID_QKV, (uint32_t)(ks > 0));
            tc_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();
        TR(4);

        // ---- C. in_proj epilogue: + bias -> bf16 -> per-head Q / K / V tiles (64 B rows = one head's 32 dims). (Q rows carry
        //         log2(e) / sqrt(d): caller.)
        for_acc32<6>(trow, [&](int i, const uint32_t* v) {
            p2 f[16];
            if (i < 2) {
                acc_bias32(v, prm_s + 4u * (uint32_t)(P_BQKV + i * 32), f);
            } else {   // K / V carry no bias here: a key bias is softmax-invariant, the value bias is folded into out_proj's by the caller
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = pk2u(v[2 * j], v[2 * j + 1]);
            }
            const int sel = i >> 1, hh = i & 1;
            const uint32_t tile = (sel == 0 ? qk : sel == 1 ? qk + 16384u : vv) + (uint32_t)hh * 8192u;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                st_shared_v4(row64(tile, r, c), pack_p2(f[4 * c]), pack_p2(f[4 * c + 1]), pack_p2(f[4 * c + 2]), pack_p2(f[4 * c + 3]));
        });
        TR(5);
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
        TR(6);

        // ---- D. softmax over the 49 keys of the row's own window (S is in log2 units), P (bf16, unnormalised) -> the QK region as two
        //         K-major chunks of 64 keys; the other window's chunk of this row is zero. O_h = P V_h.
        float inv[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t sv[64];
            tmem_ld32(trow + (uint32_t)(h * 128 + wsel * 64), sv);
            tmem_ld32(trow + (uint32_t)(h * 128 + wsel * 64 + 32), sv + 32);
            tmem_wait32(sv);
            tmem_wait32(sv + 32);
            float mx = __uint_as_float(sv[0]);
#pragma unroll
            for (int k = 1; k < T; ++k) mx = fmaxf(mx, __uint_as_float(sv[k]));
            const p2 nmx = pk2(-mx, -mx);
            uint32_t pk[32];
            p2 la = pk2(0.f, 0.f), lb = la;
#pragma unroll
            for (int k = 0; k < 24; ++k) {
                const float2 d = up2(add2(pk2u(sv[2 * k], sv[2 * k + 1]), nmx));
                const float e0 = ex2_approx(d.x), e1 = ex2_approx(d.y);
                if (k & 1) lb = add2(lb, pk2(e0, e1)); else la = add2(la, pk2(e0, e1));
                pk[k] = pack2(e0, e1);
            }
            const float e48 = ex2_approx(__uint_as_float(sv[48]) - mx);
            pk[24] = pack2(e48, 0.f);
#pragma unroll
            for (int k = 25; k < 32; ++k) pk[k] = 0u;
            const float2 l2 = up2(add2(la, lb));
            inv[h] = 1.0f / (l2.x + l2.y + e48);
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
        TR(7);

        // ---- E. attention output (normalised) -> A1 -> out_proj. V is consumed: the identity rows of both windows take its place.
        if (r == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ibar), "r"(ok_b ? 2u * ID_BYTES : ID_BYTES) : "memory");
            tma_load_4d(vv, &tmX7, ibar, 0, w0_a, h0_a, n_a);
            if (ok_b) tma_load_4d(vv + ID_STRIDE, &tmX7, ibar, 0, w0_b, h0_b, n_b);
        }
        for_acc32<2>(trow, [&](int i, const uint32_t* v) {
            const p2 sc = pk2(inv[i], inv[i]);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                st_shared_v4(row128(a1, r, 4 * i + c), pack_p2(mul2(pk2u(v[8 * c], v[8 * c + 1]), sc)), pack_p2(mul2(pk2u(v[8 * c + 2], v[8 * c + 3]), sc)),
                             pack_p2(mul2(pk2u(v[8 * c + 4], v[8 * c + 5]), sc)), pack_p2(mul2(pk2u(v[8 * c + 6], v[8 * c + 7]), sc)));
        });
        TR(8);
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
        TR(9);

        // ---- F. + bias + residual -> LayerNorm 2 -> A1 -> MLP linear 1
        for_acc32<2>(trow + 64u, [&](int i, const uint32_t* v) {
            p2 f[16];
            acc_bias32(v, prm_s + 4u * (uint32_t)(P_BO + i * 32), f);
#pragma unroll
            for (int j = 0; j < 16; ++j) x1[16 * i + j] = add2(x1[16 * i + j], f[j]);
        });
        layernorm_store(a1, r, x1);
        TR(10);
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
        TR(11);

        // ---- G. + bias -> GELU (tanh form, common.cuh ysod_gelu_tanh, on pairs) -> bf16 hidden tile (two K chunks of 64 in the QK region)
        //         -> MLP linear 2
        {
            const p2 g3 = pk2(-3.2060743e-4f, -3.2060743e-4f), g2 = pk2(3.6819429e-2f, 3.6819429e-2f), g1 = pk2(7.9770428e-1f, 7.9770428e-1f),
                     half2 = pk2(0.5f, 0.5f);
            for_acc32<4>(trow + 128u, [&](int i, const uint32_t* v) {
                p2 f[16];
                acc_bias32(v, prm_s + 4u * (uint32_t)(P_B1 + i * 32), f);
                uint32_t o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const p2 xx = mul2(f[j], f[j]);
                    const p2 q = fma2(fma2(g3, xx, g2), xx, g1);
                    const float2 a = up2(mul2(f[j], q));
                    const p2 hx = mul2(f[j], half2);
                    o[j] = pack_p2(fma2(hx, pk2(tanh_approx(a.x), tanh_approx(a.y)), hx));
                }
                const uint32_t tile = qk + (uint32_t)(i >> 1) * 16384u;
#pragma unroll
                for (int c = 0; c < 4; ++c) st_shared_v4(row128(tile, r, (i & 1) * 4 + c), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
            });
        }
        TR(12);
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
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();
        if (issuer_warp && elect_one()) request_patches(pr + gridDim.x * 2u);   // the hidden tile is consumed: the QK region takes the next tile's patches
        TR(13);

        // ---- H. + bias + residual -> bf16 -> A1 -> pw 1x1 (BN folded)
        for_acc32<2>(trow, [&](int i, const uint32_t* v) {
            p2 f[16];
            acc_bias32(v, prm_s + 4u * (uint32_t)(P_B2 + i * 32), f);
#pragma unroll
            for (int j = 0; j < 16; ++j) x1[16 * i + j] = add2(x1[16 * i + j], f[j]);
        });
        store_row64(a1, r, x1);
        TR(14);
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
        TR(15);

        // ---- I. + bias -> SiLU (x * sigmoid(x) = h + h * tanh(h), h = x / 2) -> + identity (read from the thread's own row of the V region) -> bf16
        //         into the same row -> one TMA store per window (7 x 7 x 64 box, clipped at the image border)
        mbar_wait(ibar, iphase); iphase ^= 1u;
        {
            const uint32_t orow = vv + (uint32_t)wsel * ID_STRIDE;
            const p2 half2 = pk2(0.5f, 0.5f);
            for_acc32<2>(trow + 64u, [&](int i, const uint32_t* v) {
                p2 f[16];
                acc_bias32(v, prm_s + 4u * (uint32_t)(P_BPW + i * 32), f);
                if (t < T) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t addr = row128(orow, t, 4 * i + c);
                        const uint4 idw = lds_v4(addr);
                        const uint32_t iw[4] = {idw.x, idw.y, idw.z, idw.w};
                        uint32_t ov[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const p2 hx = mul2(f[4 * c + j], half2);
                            const float2 hf = up2(hx);
                            ov[j] = pack_p2(add2(fma2(hx, pk2(tanh_approx(hf.x), tanh_approx(hf.y)), hx), unpack_p2(iw[j])));
                        }
                        st_shared_v4(addr, ov[0], ov[1], ov[2], ov[3]);
                    }
                }
            });
        }
        tc_fence_before();   // this tile's TMEM reads are ordered before the barriers of the next tile's first MMA group
        fence_async_smem();
        group_barrier(grp);
        if (r == 0) {
            tma_store_4d(&tmO7, vv, 0, w0_a, h0_a, n_a);
            if (ok_b) tma_store_4d(&tmO7, vv + ID_STRIDE, 0, w0_b, h0_b, n_b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        TR(16);
        tc_fence_before();   // this tile's TMEM reads are ordered before the barriers of the next tile's first MMA group
    }

    if (r == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the last output stores are complete before shared memory goes away
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// profiling only: enable / read the stage trace of swin64_tc_kernel (8 tiles x 24 clock64 stamps of CTA 0, thread 0)
extern "C" int ysod_swin64_tc_trace(int enable, long long* host_out) {
    if (host_out) YSOD_CUDA(cudaMemcpyFromSymbol(host_out, g_trace, sizeof(long long) * 8 * 24));
    YSOD_CUDA(cudaMemcpyToSymbol(g_trace_on, &enable, sizeof(int)));
    return YSOD_OK;
}

// Same contract as ysod_swin64_fused (swin_fused.cu): x / out NHWC 16-bit views with 64 channels, the caller pre-folds the LayerNorm
// affine parts into in_proj / mlp.0 and log2(e) / sqrt(head_dim) into the Q rows; wbf16 / pf32 are the same blobs. In addition this kernel
// does not read the K and V thirds of the in_proj bias (pf32[192..320)): the caller zeroes them -- a bias added to every key shifts all scores
// of a query by the same amount (softmax-invariant), a bias added to every value passes through the attention average and belongs to
// out_proj's bias (bo += Wo bv).
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
    YSOD_CHECK_ARG((long long)N * nWh * nWw < (1ll << 31), "ysod_swin64_tc: too many windows");
    const long long npairs = ((long long)N * nWh * nWw + 1) / 2;
    long long grid = (npairs + 1) / 2;
    if (grid > sms) grid = sms;
    // the input map as a 4-D tensor (C = 64 of xcs, W, H, N); one TMA box = a window's 9 x 9 patch in 128 B swizzled pixel rows
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) enc = (EncodeTiledFn)fp;
    }
    YSOD_CHECK_ARG(enc != nullptr, "ysod_swin64_tc: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    CUtensorMap tmX, tmX7, tmO7;
    {
        const cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        struct { CUtensorMap* m; const void* ptr; int cs; cuuint32_t side; } maps[3] = {{&tmX, x, xcs, 9}, {&tmX7, x, xcs, 7}, {&tmO7, out, ocs, 7}};
        for (auto& mp : maps) {   // 9 x 9 input patches (conv halo), 7 x 7 identity windows, 7 x 7 output windows
            const cuuint64_t strides[3] = {(cuuint64_t)mp.cs * 2, (cuuint64_t)W * mp.cs * 2, (cuuint64_t)H * W * mp.cs * 2};
            const cuuint32_t box[4] = {64, mp.side, mp.side, 1};
            const CUresult cr = enc(mp.m, YSOD_TMAP_16, 4, const_cast<void*>(mp.ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) {
                ysod_set_error("ysod_swin64_tc: cuTensorMapEncodeTiled failed with %d (W %d H %d N %d pixel stride %d)", (int)cr, W, H, N, mp.cs);
                return YSOD_ERR_CUDA;
            }
        }
    }
    YSOD_CUDA(cudaFuncSetAttribute(swin64_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    ysod_launch(swin64_tc_kernel, (unsigned)grid, 256, SMEM_BYTES, stream, tmX, tmX7, tmO7, (const __nv_bfloat16*)x, N, H, W, xcs, (const __nv_bfloat16*)wbf16, pf32,
                (__nv_bfloat16*)out, ocs, nWh, nWw);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}
