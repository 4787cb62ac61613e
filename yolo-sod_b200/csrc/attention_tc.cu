// softmax(q k^T * scale) v on the 5th-gen tensor cores: both contractions are tcgen05.mma with the accumulators in TMEM.
//
// Replaces (reference): the attention core inside nn.MultiheadAttention of WindowAttention (blocks_transformer.py:98,116: the P4
// SwinBlock, 49-token windows, head_dim 64) and of A2_Attn (a2_attn.py:29,53), and the manual softmax path of AAttn
// (block.py:1348-1357: area attention of A2C2f, 400..1600 tokens per area, head_dim 32) -- the QK^T / PV contractions that
// north_star assigns to tcgen05 / TMEM. The mma.sync kernels in attention.cu remain as the A/B baseline (ysod_mha_core_ex impl 1)
// and for head_dim 16.
//
// One CTA (4 warps) owns a tile of 128 query rows of one head; thread t <-> query row t <-> TMEM lane t.
//   * L <= 64 (windows): two windows are packed into one tile (rows 0..63 / 64..127, likewise the 128 key rows); S = Q K^T is ONE
//     128 x 128 x D MMA group, the cross-window quadrants and the padding keys are masked in the softmax (2x the useful S FLOPs, which
//     are a few % of the block) -- this is what gives a 49-token window an M = 128 tensor-core tile.
//   * L > 64: 128 consecutive queries of one sequence, key tiles of 128 streamed with an online softmax.
// Operands are staged in shared memory in the canonical UMMA layouts by plain 16 B loads (rows are 64 / 128 B, strided in global
// memory: in_proj output is [token][3E]): Q and K as K-major SWIZZLE_128B (D = 64) / SWIZZLE_64B (D = 32) tiles, V row-major
// [key][d] = an MN-major B operand (no transpose anywhere), P (bf16) as two K-major SWIZZLE_128B chunks of 64 keys.
// Per key tile: S -> TMEM (D/16 MMAs) -> every thread reads its own row (tcgen05.ld): max, exp, sum need no shuffles -> P to shared
// memory -> O_tile = P V -> TMEM (8 MMAs, into the columns S just vacated) -> registers, rescaled and accumulated in fp32.
#include "umma.cuh"

namespace {
using namespace umma;

// row r of a 128-row tile -> (batch item, token, in range). packed: rows 0..63 = item b0, 64..127 = item b0 + 1 (tokens 0..63);
// otherwise item b0, tokens t0 + r.
struct RowMap {
    int packed, b0, t0, L, batch;
    __device__ __forceinline__ bool map(int r, int& bi, int& ti) const {
        if (packed) { bi = b0 + (r >> 6); ti = r & 63; return ti < L && bi < batch; }
        bi = b0; ti = t0 + r;
        return ti < L;
    }
};

// 128 rows x D bf16 (one head's slice of the token rows) -> shared memory, 16 B chunk c of row r at r*ROWB + ((c ^ swz(r)) << 4):
// the 128 B / 64 B swizzle TMA would produce (a function of the address bits [7..9] / [7..8]; the tile is 1 KB aligned). Rows
// out of range are zero-filled (a garbage V row would turn 0 * NaN into NaN). All of a thread's loads are issued before its stores.
template <int D>
struct TileRegs { uint4 v[D / 8]; };
template <int D>
__device__ __forceinline__ void fetch_tile(TileRegs<D>& t, const __nv_bfloat16* __restrict__ src, int ld, long long bs, int h,
                                           const RowMap& rm, int tid) {
    constexpr int CH = D / 8;
#pragma unroll
    for (int it = 0; it < CH; ++it) {
        const int i = tid + it * 128;
        const int r = i / CH, c = i - r * CH;
        int bi, ti;
        t.v[it] = make_uint4(0, 0, 0, 0);
        if (rm.map(r, bi, ti)) t.v[it] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)bi * bs + (size_t)ti * ld + h * D + c * 8));
    }
}
template <int D>
__device__ __forceinline__ void stash_tile(uint32_t dst, const TileRegs<D>& t, int tid) {
    constexpr int CH = D / 8, ROWB = D * 2;
#pragma unroll
    for (int it = 0; it < CH; ++it) {
        const int i = tid + it * 128;
        const int r = i / CH, c = i - r * CH;
        const int swz = (D == 64) ? (r & 7) : ((r >> 1) & 3);
        st_shared_v4(dst + (uint32_t)(r * ROWB + ((c ^ swz) << 4)), t.v[it].x, t.v[it].y, t.v[it].z, t.v[it].w);
    }
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <int D>
__global__ void __launch_bounds__(128, 3)
mha_tc_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v, int L, int ldq,
              int ldk, int ldv, long long bsq, long long bsk, long long bsv, float scale, __nv_bfloat16* __restrict__ out, int ldo,
              long long bso, int batch, int packed, int n_qt) {
    ysod_pdl_sync();
    constexpr int ROWB = D * 2;                 // bytes per operand row (one swizzle span)
    constexpr uint32_t TILE = 128u * ROWB;      // Q / K / V tile
    constexpr uint32_t LAYOUT = (D == 64) ? 2u : 4u;   // SWIZZLE_128B : SWIZZLE_64B
    constexpr uint32_t SBO = 8u * ROWB;         // 8-row group pitch
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_s = base, k_s = base + TILE, v_s = base + 2u * TILE, p_s = base + 3u * TILE;   // p_s: 2 chunks x 128 rows x 128 B
    const uint32_t bar = p_s + 32768u, tmem_slot = bar + 8u;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int h = blockIdx.y;

    RowMap qm;
    qm.packed = packed; qm.L = L; qm.batch = batch;
    if (packed) { qm.b0 = blockIdx.x * 2; qm.t0 = 0; }
    else { qm.b0 = blockIdx.x / n_qt; qm.t0 = (blockIdx.x - qm.b0 * n_qt) * 128; }

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // all three operand tiles of the first key tile are requested before any is stored: one global round trip, not three
    TileRegs<D> kr, vr;
    {
        TileRegs<D> qr;
        fetch_tile<D>(qr, q, ldq, bsq, h, qm, tid);
        fetch_tile<D>(kr, k, ldk, bsk, h, qm.packed ? qm : RowMap{0, qm.b0, 0, L, batch}, tid);
        fetch_tile<D>(vr, v, ldv, bsv, h, qm.packed ? qm : RowMap{0, qm.b0, 0, L, batch}, tid);
        stash_tile<D>(q_s, qr, tid);
    }
    if (packed) {
        // a row only attends to its own window: the off-diagonal quadrants of P (rows 0..63 x keys 64..127 and vice versa) stay zero
        // for the whole kernel -- chunk 1 of rows 0..63, chunk 0 of rows 64..127 (the swizzle permutes inside a row only)
        const uint32_t za = p_s + (uint32_t)(1 - (tid >> 6)) * 16384u + (uint32_t)tid * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j) st_shared_v4(za + (uint32_t)(j << 4), 0u, 0u, 0u, 0u);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's TMEM lane quarter

    constexpr uint32_t IDESC_S = idesc_f16(128, 128, 0, 0);       // S[128 q][128 keys] = Q (K-major) x K^T (K-major)
    constexpr uint32_t IDESC_PV = idesc_f16(128, D, 0, 1);        // O[128 q][D]      = P (K-major) x V (MN-major: [key][d])

    float o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint32_t phase = 0;
    const int r = tid;
    const int nkt = packed ? 1 : (L + 127) / 128;
    const int c0 = packed ? (r >> 6) * 4 : 0, c1 = packed ? c0 + 4 : 8;   // 16-column chunks of S this row attends to
    for (int kt = 0; kt < nkt; ++kt) {
        stash_tile<D>(k_s, kr, tid);
        stash_tile<D>(v_s, vr, tid);
        fence_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (kt + 1 < nkt) {   // the next key tile travels while this one is multiplied and soft-maxed
            RowMap km = qm;
            km.t0 = (kt + 1) * 128;
            fetch_tile<D>(kr, k, ldk, bsk, h, km, tid);
            fetch_tile<D>(vr, v, ldv, bsv, h, km, tid);
        }
        if (warp == 0 && elect_one()) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks)   // UMMA_K = 16 elements = 32 B along the swizzled row
                tc_mma(tmem, smem_desc(q_s + ks * 32u, 16u, SBO, LAYOUT), smem_desc(k_s + ks * 32u, 16u, SBO, LAYOUT), IDESC_S, (uint32_t)(ks > 0));
            tc_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();

        // ---- softmax of this thread's row. Pass 1: row max over the valid keys.
        const int kvalid = packed ? L : min(128, L - kt * 128);          // valid keys per 64-key half (packed) / in this tile
        const int klo = packed ? (r >> 6) * 64 : 0;                      // first key column this row may attend to
        float mx = -INFINITY;
#pragma unroll 2
        for (int c = c0; c < c1; ++c) {
            uint32_t sv[16];
            tmem_ld16(trow + (uint32_t)(c * 16), sv);
            tmem_ld_wait(sv);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int col = c * 16 + j - klo;
                if (col < kvalid) mx = fmaxf(mx, __uint_as_float(sv[j]) * scale);
            }
        }
        const float m_new = fmaxf(m_run, mx);
        const float corr = __expf(m_run - m_new);      // first tile: exp(-inf) = 0 (o and l are 0 anyway)
        float l_add = 0.f;
        // Pass 2: P = exp(s - m) as bf16 -> shared memory (K-major SWIZZLE_128B, 64 keys per chunk), row sum in fp32
#pragma unroll 2
        for (int c = c0; c < c1; ++c) {
            uint32_t sv[16];
            tmem_ld16(trow + (uint32_t)(c * 16), sv);
            tmem_ld_wait(sv);
            float p[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int col = c * 16 + j - klo;
                p[j] = (col < kvalid) ? __expf(__uint_as_float(sv[j]) * scale - m_new) : 0.f;
                l_add += p[j];
            }
            const uint32_t rowaddr = p_s + (uint32_t)(c >> 2) * 16384u + (uint32_t)r * 128u;
            const uint32_t pc = (uint32_t)(c & 3) * 2u, sw = (uint32_t)(r & 7);
            st_shared_v4(rowaddr + (((pc) ^ sw) << 4), pack2(p[0], p[1]), pack2(p[2], p[3]), pack2(p[4], p[5]), pack2(p[6], p[7]));
            st_shared_v4(rowaddr + (((pc + 1u) ^ sw) << 4), pack2(p[8], p[9]), pack2(p[10], p[11]), pack2(p[12], p[13]), pack2(p[14], p[15]));
        }
        l_run = l_run * corr + l_add;
        m_run = m_new;
        // every thread's reads of S are complete (tcgen05.wait::ld): the PV product may overwrite those TMEM columns
        tc_fence_before();
        fence_async_smem();
        __syncthreads();
        if (warp == 0 && elect_one()) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)   // 16 keys per MMA: A advances 32 B inside a 64-key chunk, B advances 16 V rows
                tc_mma(tmem, smem_desc(p_s + (uint32_t)(ks >> 2) * 16384u + (uint32_t)(ks & 3) * 32u, 16u, 1024u, 2u),
                       smem_desc(v_s + (uint32_t)ks * 16u * ROWB, SBO, SBO, LAYOUT), IDESC_PV, (uint32_t)(ks > 0));
            tc_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < D / 16; ++c) {
            uint32_t ov[16];
            tmem_ld16(trow + (uint32_t)(c * 16), ov);
            tmem_ld_wait(ov);
#pragma unroll
            for (int j = 0; j < 16; ++j) o[c * 16 + j] = fmaf(o[c * 16 + j], corr, __uint_as_float(ov[j]));
        }
        tc_fence_before();
        __syncthreads();   // K / V / P tiles and the TMEM columns are reused by the next key tile
    }

    int bi, ti;
    if (qm.map(r, bi, ti)) {
        const float inv = 1.0f / l_run;
        __nv_bfloat16* op = out + (size_t)bi * bso + (size_t)ti * ldo + h * D;
#pragma unroll
        for (int d = 0; d < D; d += 8) {
            uint4 w;
            w.x = pack2(o[d] * inv, o[d + 1] * inv);
            w.y = pack2(o[d + 2] * inv, o[d + 3] * inv);
            w.z = pack2(o[d + 4] * inv, o[d + 5] * inv);
            w.w = pack2(o[d + 6] * inv, o[d + 7] * inv);
            *reinterpret_cast<uint4*>(op + d) = w;
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
    }
}

}  // namespace

// tcgen05 attention core; same addressing as ysod_mha_core. Returns YSOD_ERR_UNSUPPORTED (without touching the error text of a
// successful fallback) when the shape is not covered: head_dim 32 / 64, 16 B aligned rows, bf16.
int ysod_mha_tc_launch(const void* q, const void* k, const void* v, int batch, int L, int heads, int D, int ldq, int ldk, int ldv,
                       long long bsq, long long bsk, long long bsv, float scale, void* out, int ldo, long long bso, cudaStream_t st) {
    if (!(D == 32 || D == 64) || L < 1 || batch < 1 || heads < 1 || heads > 65535) return YSOD_ERR_UNSUPPORTED;
    if (((uintptr_t)q % 16) || ((uintptr_t)k % 16) || ((uintptr_t)v % 16) || ((uintptr_t)out % 16) || (ldq % 8) || (ldk % 8) || (ldv % 8) ||
        (ldo % 8) || (bsq % 8) || (bsk % 8) || (bsv % 8) || (bso % 8))
        return YSOD_ERR_UNSUPPORTED;
    const int packed = L <= 64 ? 1 : 0;
    const int n_qt = packed ? 1 : (L + 127) / 128;
    const long long gx = packed ? (batch + 1) / 2 : (long long)batch * n_qt;
    if (gx > 2147483647ll) return YSOD_ERR_UNSUPPORTED;
    dim3 grid((unsigned)gx, (unsigned)heads);
    const size_t smem = 3u * 128u * (size_t)D * 2u + 32768u + 64u + 1024u;
    if (D == 64) {
        YSOD_CUDA(cudaFuncSetAttribute(mha_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ysod_launch(mha_tc_kernel<64>, grid, 128, smem, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, L, ldq, ldk, ldv, bsq,
                    bsk, bsv, scale, (__nv_bfloat16*)out, ldo, bso, batch, packed, n_qt);
    } else {
        YSOD_CUDA(cudaFuncSetAttribute(mha_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ysod_launch(mha_tc_kernel<32>, grid, 128, smem, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, L, ldq, ldk, ldv, bsq,
                    bsk, bsv, scale, (__nv_bfloat16*)out, ldo, bso, batch, packed, n_qt);
    }
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}
