// Implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (tcgen05 + TMEM), operands fed by TMA.
//
// Replaces (reference): every dense Conv2d(+BN)+act on the path -- ultralytics/nn/modules/conv.py:37-55 (Conv),
// block.py:233-356 (C2f/Bottleneck convs), head.py:43-57 (Detect cv2/cv3), and the nn.Linear / in_proj / out_proj
// GEMMs of nn.MultiheadAttention (blocks_transformer.py:98-106, a2_attn.py:29) -- which the reference hands to
// cuDNN/cuBLAS plus separate BN / SiLU / add / cat kernels.
//
// GEMM view: D[M = pixels, N = Cout] = sum over K = (r, s, cin) of A[pixel shifted by tap (r,s), cin] * Wt[cout, K].
//   * activations are NHWC bf16 (any channel-sliced view: pixel stride `xcs`), weights [Cout_pad][k*k*Cin] bf16
//     (K-major, BN already folded in), bias fp32.
//   * A CTA owns one TH x TW patch of output pixels of one image (TH*TW <= 128 = UMMA M) and BN output channels.
//   * For every K block (one filter tap x BK input channels) ONE 4-D tiled TMA box {BK, TW, TH, 1} lands the
//     shifted patch directly in the canonical K-major 128B/64B-swizzled UMMA layout (row = pixel, 128/64 B of
//     channels); the conv zero padding is TMA out-of-bounds fill; stride-2 convs use TMA element strides {1,2,2,1}.
//     No im2col buffer ever exists in HBM.
//   * ONE persistent CTA of 608 threads per SM: warps 0..15 = epilogue (two groups of 8 warps, one per tile parity: tcgen05.ld ->
//     +bias -> act (fp32 pairs, fma.rn.f32x2) -> (+residual) -> bf16/fp32 -> swizzled staging tile -> TMA tensor store into a channel
//     slice of the consumer's NHWC buffer: this is what removes Concat/chunk copies), warp 16 = TMA producer (one elected lane),
//     warps 17 / 18 = tcgen05.mma issuers (one elected lane each; the second one only for single-burst tiles), warp 17 also owns TMEM.
//   * smem rings of {A,B} tiles (generic mode) or separate halo-copy / weight-tap rings (3x3 stride-1 mode) with full/empty mbarriers;
//     four accumulators (128 lanes x BN fp32 columns each) in TMEM, so the epilogue of tile i overlaps the main loop of tile i+1.
//
// BN/BK/stage counts are runtime values (instruction + smem descriptors are built from them), so one kernel serves every layer shape;
// the plan (ysod_conv_tc_create) sizes the rings against the 227 KB of one SM.
#include "common.cuh"
#include <cuda.h>
#include <new>

namespace {

struct TcParams {
    int N, Ho, Wo, TH, TW, tiles_h, tiles_w;
    int Cin, ksize, stride, pad;
    int BN, BK, stages, tmem_cols, num_k, n_tiles;
    int step2_tw, step2_th, step2_img, step2_nt;   // the same for a stride of 2 * gridDim.x tiles (epilogue groups; tile pairs)
    int pair;        // 1: the CTA works on PAIRS of raster-adjacent output tiles (2c, 2c+1, then + 2*gridDim.x ...) that share every
                     // weight (B) fetch: each B tile / tap landed in shared memory is multiplied with both tiles' A operands into two
                     // TMEM accumulators. Deep layers are bound by the chip-wide L2 -> SM throughput (~12 TB/s), not by the tensor
                     // pipe; sharing B cuts the bytes per MMA by 25 % (generic stages) to 45 % (streamed 3x3 taps).
    int step_tw, step_th, step_img, step_nt;   // gridDim.x decomposed in the mixed radix (tiles_w, tiles_h, N): per-tile coordinate update without divisions
    int issuers;     // MMA issuer threads: 2 = tiles alternate between two issuers on separate sub-rings (short tiles), 1 = one issuer, whole ring
    int kgroup;      // generic mode: K blocks (slots) per pipeline stage = per mbarrier handshake / per elected issue burst
    int sgroup;      // halo mode: filter columns (A copies) per issue burst (3 when the weights are resident, else 1)
    int a_stages, b_stages, b_resident, cchunks;   // stages per sub-ring
    int a_slots, b_slots;                          // total smem slots / barrier pairs carved for A and B   // halo mode (3x3/s1): separate A (halo copies) and B (weight taps) rings
    int Cout;
    void* out;
    int out_f32;
    int ocs;
    const float* bias;
    const __nv_bfloat16* res;
    int rcs;
    int act;
    uint32_t idesc;
    uint32_t desc_hi;  // SBO | version | layout type (upper 32 bits of the smem descriptor)
    uint32_t a_bytes, b_bytes, a_tx;
    // epilogue staging: output rows of `row_bytes` (<= 128 B, one swizzle span) per store unit of `unit_cols` columns
    int unit_cols, n_units, swz_mask, cout_pad, stage_bufs;
    int stage_split;   // 1: the staging buffer holds ONE store unit; the epilogue makes n_units passes (frees smem for resident weight taps)
    uint32_t row_bytes;
    // fused Detect decode (ysod_conv_tc_set_decode): the layer is the level's final 1x1 head conv with output channels
    // [0,64) = DFL box logits (4 sides x 16 bins), [64,64+nc) = class logits; besides the raw map the epilogue writes
    // y[img][0..4+nc)[a_off + pixel] = (cx, cy, w, h) * stride, sigmoid(cls)   (head.py:100-131, block.py:64-83, tal.py:333-357)
    float* dec_y;
    int dec_A, dec_off, dec_nc;
    float dec_stride;
    int w_img_rows;   // > 0: per-image weights (an SE channel gate folded into the next conv): image i uses weight rows [i*w_img_rows, +Cout_pad)
    int no_store;   // the plan's own output is not written (decode-only head conv of the predict path)
    // Back-to-back GEMM (ysod_conv_tc_set_b2b): the staged bf16 output tile (128 pixels x 64 channels, 128 B swizzled rows) IS a K-major
    // SWIZZLE_128B UMMA A operand, so a following 1x1 conv (Detect cv2[i][2] / cv3[i][2], head.py:43-57) runs as a second tcgen05.mma
    // group straight from the staging buffer into its own TMEM columns: D2[128 x b2_n] = tile[128 x 64] * W2[b2_n x 64]. The second
    // epilogue adds the bias and runs the Detect decode (b2b 1: DFL box branch -> y[0..4); 2: class branch -> sigmoid -> y[4..4+nc)),
    // optionally also writing the fp32 raw map. The 3x3 conv's own output never goes to HBM.
    int b2b, b2_n, b2_raw_cs, b2_raw_off, b2_act;   // b2b 3: plain second layer (bias + b2_act -> bf16 -> the plan's own output map)
    int b2_k2;      // b2b 4 (ysod_conv_tc_set_b2b_cat): the second layer's input is cat(extra 64-channel tensor, this tile): b2_k2 = 64 extra K
                    // channels arrive per tile by TMA (tmA2) into their own two-slot ring; the second layer's output leaves through tmO2
    const __nv_bfloat16* b2_w;   // [b2_n][64] bf16, K-major
    const float* b2_bias;        // [b2_n]
    float* b2_raw;               // NHWC fp32 raw map (pixel stride b2_raw_cs), channels [b2_raw_off, ...), or nullptr
    float* b2_y;                 // y (B, 4+nc, A): geometry in dec_A / dec_off / dec_nc / dec_stride (dec_y stays null: no main-tile decode)
    int duo;    // pixel-duo plan of the halo kernel for 3x3 / stride-1 convs with 32 input and 32 output channels: one MMA row = TWO horizontally
                // adjacent output pixels (N = 64 = 2 x 32 channels), one K block = one 128 B shared-memory row = two adjacent input pixels.
                // Output pair j of a row reads input pixels 2j-1 .. 2j+2 = the two pairs at halo offsets j and j+1, so the 3 x 3 filter becomes
                // 3 rows x 2 pair-taps of K = 64 with a [64][6 * 64] weight matrix that holds each filter row twice, shifted by one pixel
                // (25 % structural zeros). Against the 32-channel plan (N = 32, 64 B rows): half the MMA instructions per pixel at the
                // N = 64 rate, and 128 B swizzled operand rows instead of 64 B ones (whose 8-row groups hit every bank twice: 83 cycles per
                // N = 32 MMA measured, against 48 in isolation).
    int up2;    // nn.Upsample(scale 2, nearest) fused into the store: every output pixel is written to its 2 x 2 block of the 2Ho x 2Wo destination
    int debug;  // profiling only (mode >> 8): 1 = epilogue drains without work, 2 = producer skips TMA, 4 = MMA issuer skips tcgen05.mma,
                // 8 = epilogue skips the TMA store, 16 = epilogue skips the activation
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: a waiting warp sleeps in hardware until the phase completes instead of spinning in the
// issue slots. (The SMSP arbiter favours higher warp ids, so a spinning epilogue warp starves a lower-numbered issuer warp.)
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    return ok;
}
// Bounded wait: a protocol bug (wrong tx count, bad descriptor) traps after ~2 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// fp32 pairs (sm_100 add / mul / fma .f32x2): two lanes per issue slot, each lane rounded like the scalar instruction (bit-identical results)
typedef unsigned long long p2;
__device__ __forceinline__ p2 pk2(float lo, float hi) { p2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ p2 pk2u(uint32_t lo, uint32_t hi) { p2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ float2 up2(p2 v) { float2 f; asm("mov.b64 {%0, %1}, %2;" : "=f"(f.x), "=f"(f.y) : "l"(v)); return f; }
__device__ __forceinline__ p2 add2(p2 a, p2 b) { p2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ p2 mul2(p2 a, p2 b) { p2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ p2 fma2(p2 a, p2 b, p2 c) { p2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
// the "+r" operands tie the loaded registers to the wait so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}

// Profiling aid (debug bit 32): CTA 0 logs clock64() at pipeline events of its tiles; read back with ysod_debug_trace().
// Record = (role << 56 | event << 48 | tile << 32 | index) , clock.  roles: 0 producer, 1 MMA issuer 0, 2 epilogue (warp 0
// lane 0), 3 MMA issuer 1. Each role appends to its own quarter of the buffer with a private counter (no atomics: a trace
// point is one clock read + one fire-and-forget store).
constexpr int TRACE_CAP = 8192;
__device__ unsigned long long g_trace[2 * TRACE_CAP];
__device__ __forceinline__ void trace(bool on, int role, int ev, int tile, int idx, unsigned int& cnt) {
    if (on && cnt < TRACE_CAP / 4) {
        const unsigned int i = role * (TRACE_CAP / 4) + cnt++;
        g_trace[2 * i] = ((unsigned long long)role << 56) | ((unsigned long long)ev << 48) | ((unsigned long long)tile << 32) | (unsigned int)idx;
        g_trace[2 * i + 1] = clock64();
    }
}

// One fat persistent CTA per SM: warps 0..15 epilogue, warp 16 TMA producer, warps 17 and 18 MMA issuers.
//  * 16 epilogue warps = 4 per SMSP: enough warps to hide TMEM-load / MUFU / shared-store latency. Warp w owns TMEM lane
//    quarter w % 4 (a hardware rule) and the 16-column chunks  w / 4, w / 4 + 4, ...  of the accumulator.
//  * two issuers ping-pong over the CTA's tiles (issuer i takes every second tile and always accumulator i): a
//    tcgen05.mma burst blocks its issuing thread at tensor-pipe rate (the hardware queue is shallow), so with a single
//    issuer the barrier handshakes between bursts (~1000+ cycles, measured) leave the tensor pipe idle.
//  * the issuer/producer warps carry the highest warp ids of their SMSPs, which the hi-wid-first arbiter favours.
constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int PRODUCER_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1;
constexpr int TC_THREADS = EPI_THREADS + 96;

// Programmatic dependent launch: the next kernel in the stream may start its prologue while this grid drains; everything that
// touches activations produced by the previous kernel happens after pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// One elected lane of a converged warp. Unlike `lane == 0`, the compiler knows the branch holds exactly one thread, so
// tcgen05.mma / tcgen05.commit / TMA (which take uniform-register operands) are emitted straight, without a per-lane
// ELECT + BRA.U.ANY serialisation loop around every instruction (that loop cost ~150 cycles per MMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t desc_hi, uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; }
// low word of the smem descriptor: start address >> 4 (14 bits) | LBO = 1 (ignored for swizzled K-major) @16
__device__ __forceinline__ uint32_t umma_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); }

template <int ACT>
__device__ __forceinline__ void act16(float* f) {
    if (ACT == YSOD_ACT_GELU) {   // ysod_gelu_tanh (common.cuh) on fp32 pairs: same operations, same rounding
        const p2 g3 = pk2(-3.2060743e-4f, -3.2060743e-4f), g2 = pk2(3.6819429e-2f, 3.6819429e-2f), g1 = pk2(7.9770428e-1f, 7.9770428e-1f), half2 = pk2(0.5f, 0.5f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const p2 x = pk2(f[2 * j], f[2 * j + 1]);
            const p2 xx = mul2(x, x);
            const float2 a = up2(mul2(x, fma2(fma2(g3, xx, g2), xx, g1)));
            const p2 hx = mul2(half2, x);
            const float2 o = up2(fma2(hx, pk2(tanh_approx(a.x), tanh_approx(a.y)), hx));
            f[2 * j] = o.x; f[2 * j + 1] = o.y;
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (ACT == YSOD_ACT_SILU) {  // x*sigmoid(x) = h + h*tanh(h), h = x/2: one MUFU op per element
            const float h = 0.5f * f[j];
            f[j] = fmaf(h, tanh_approx(h), h);
        }
        else if (ACT == YSOD_ACT_GELU) f[j] = ysod_gelu_tanh(f[j]);
        else if (ACT == YSOD_ACT_RELU) f[j] = fmaxf(f[j], 0.0f);
    }
}

// Output-tile coordinates of a persistent CTA, advanced by gridDim.x tiles per step with carries instead of div/mod: a
// single-thread role pays ~150-200 cycles per integer division (a dependent ~35-instruction chain), and four of them per
// tile used to cost more than the tile's MMAs.
struct TileIter {
    int tw, th, img, nt;
    __device__ __forceinline__ void init(int t, const TcParams& p) {
        tw = t % p.tiles_w; t /= p.tiles_w;
        th = t % p.tiles_h; t /= p.tiles_h;
        img = t % p.N;
        nt = t / p.N;
    }
    __device__ __forceinline__ void step(const TcParams& p) {
        tw += p.step_tw; if (tw >= p.tiles_w) { tw -= p.tiles_w; ++th; }
        th += p.step_th; if (th >= p.tiles_h) { th -= p.tiles_h; ++img; }
        img += p.step_img; if (img >= p.N) { img -= p.N; ++nt; }
        nt += p.step_nt;
    }
    __device__ __forceinline__ void step2(const TcParams& p) {   // 2 * gridDim.x tiles ahead
        tw += p.step2_tw; if (tw >= p.tiles_w) { tw -= p.tiles_w; ++th; }
        th += p.step2_th; if (th >= p.tiles_h) { th -= p.tiles_h; ++img; }
        img += p.step2_img; if (img >= p.N) { img -= p.N; ++nt; }
        nt += p.step2_nt;
    }
    __device__ __forceinline__ void step1(const TcParams& p) {   // the next tile in raster order
        if (++tw >= p.tiles_w) { tw = 0; if (++th >= p.tiles_h) { th = 0; if (++img >= p.N) { img = 0; ++nt; } } }
    }
    __device__ __forceinline__ bool valid(const TcParams& p) const { return nt < p.n_tiles; }
};

// Persistent, warp-specialised: each CTA loops over output tiles (tile = blockIdx.x + i*gridDim.x). The smem ring keeps
// streaming across tile boundaries and the accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of
// tile i overlaps the TMA/MMA main loop of tile i+1; barriers and TMEM are set up once per CTA.
//
// HALO = true is the 3x3 / stride-1 specialisation that removes the 9x re-fetch of the activation tile from L2: an output
// tile is 16 rows x 8 columns; for every 64-channel chunk ONE halo copy (18 rows x 10 px x 128 B = 22.5 KB) is landed by TMA
// as a flat array of 128 B pixel rows (SWIZZLE_128B, conv zero padding = TMA out-of-bounds fill). Filter tap (r, s) is then
// just a different UMMA A descriptor into the same copy: start address = copy + (r*10 + s) * 128 B, the 16 groups of 8
// output pixels (one output row each) are SBO = 10 pixels = 1280 B apart. The 128 B swizzle is a function of absolute
// shared-memory address bits (verified on B200: base_offset = 0 is correct for any 128 B-aligned start in a 1 KB-aligned slot). A traffic per tile drops from 9 x 16 KB to 22.5 KB; weight taps stay
// resident in shared memory for the whole CTA lifetime when they fit (64->64: 72 KB), else stream through their own ring.
template <bool HALO>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmO2,
               const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = smem_u32(smem_raw);
    const uint32_t base = (smem0 + 1023u) & ~1023u;
    const uint32_t a_base = base;
    const int a_slots = p.a_slots, b_slots = p.b_slots;   // total ring slots (both sub-rings)
    const uint32_t b_base = base + (uint32_t)a_slots * p.a_bytes;
    const uint32_t bar_base = b_base + (uint32_t)b_slots * p.b_bytes;  // 8-byte aligned (tiles are 1 KB multiples)
    // fullA[i], emptyA[i] (i < a_stages), fullB[i], emptyB[i] (i < b_stages), then tfull[2], tempty[2], tmem slot.
    // Generic mode uses the A barriers for the combined {A,B} stage.
    const uint32_t fullA = bar_base, emptyA = bar_base + 8u * a_slots;
    const uint32_t fullB = bar_base + 16u * a_slots, emptyB = fullB + 8u * b_slots;
    const uint32_t tfull_bar = fullB + 16u * b_slots;
    const uint32_t tempty_bar = tfull_bar + 32u;   // 4 accumulators: 2 per issuer / epilogue group (double-buffered per tile parity)
    const uint32_t tmem_slot = tempty_bar + 32u;
    // after the barriers: bias[cout_pad] fp32, then the (1 KB aligned) output staging: nbuf x n_units x 128 rows x row_bytes
    const uint32_t bias_smem = tmem_slot + 16u;
    const uint32_t stage_out = (bias_smem + 4u * (uint32_t)p.cout_pad + 1023u) & ~1023u;
    float* const bias_s = reinterpret_cast<float*>(smem_raw + (bias_smem - smem0));
    // back-to-back GEMM operands (b2b): W2 tile (1 KB aligned, b2_n rows of 128 B), its bias, one mbarrier per epilogue group
    const uint32_t b2_w_smem = stage_out + (uint32_t)p.stage_bufs * (uint32_t)(p.stage_split ? 1 : p.n_units) * 128u * p.row_bytes;
    const uint32_t b2_wb_smem = b2_w_smem + 128u * (uint32_t)p.b2_n;                     // b2b 4: W2 columns of the staged 32 channels (64 B rows)
    const uint32_t b2_bias_smem = b2_wb_smem + (p.b2b == 4 ? 64u * (uint32_t)p.b2_n : 0u);
    const uint32_t b2_bar = b2_bias_smem + 4u * (uint32_t)p.b2_n;                        // b2_bar[2], then (b2b 4) a2_full[2], a2_empty[2]
    const uint32_t a2_full = b2_bar + 16u, a2_empty = b2_bar + 32u;
    const uint32_t a2_smem = (b2_bar + 48u + 1023u) & ~1023u;                            // b2b 4: two 16 KB tiles of the extra operand
    const uint32_t stage2_smem = a2_smem + 2u * 16384u;                                  // b2b 4: two 16 KB staging tiles of the second layer's output
    float* const b2_bias_s = reinterpret_cast<float*>(smem_raw + (b2_bias_smem - smem0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool tr = (p.debug & 32) && blockIdx.x == 0;
    unsigned int tcnt = 0;
    trace(tr && threadIdx.x == 0, 2, 10, 0, 0, tcnt);   // kernel entry (epilogue warp 0 lane 0 logs CTA-level events)

    if (threadIdx.x == 32 && !(p.debug & 64)) {
        // the TMA unit fetches the 128 B descriptors on first use: request them now, while barriers / TMEM / bias are being set up
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmO) : "memory");
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < a_slots; ++s) {
            mbar_init(fullA + 8u * s, 1);
            mbar_init(emptyA + 8u * s, 1);
        }
        for (int s = 0; s < b_slots; ++s) {
            mbar_init(fullB + 8u * s, 1);
            mbar_init(emptyB + 8u * s, 1);
        }
        for (int a = 0; a < 4; ++a) {
            mbar_init(tfull_bar + 8u * a, 1);
            mbar_init(tempty_bar + 8u * a, EPI_WARPS / 2);  // one arrive per warp of the accumulator's epilogue group
        }
        if (p.b2b) {
            mbar_init(b2_bar, 1);
            mbar_init(b2_bar + 8u, 1);
            if (p.b2b == 4) {
                for (int g2 = 0; g2 < 2; ++g2) {
                    mbar_init(a2_full + 8u * g2, 1);
                    mbar_init(a2_empty + 8u * g2, 1);
                }
            }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp < EPI_WARPS) {
        // SiLU layers keep 0.5 * bias: the epilogue needs h = (acc + bias) / 2 = fma(acc, 0.5, 0.5 * bias), one instruction
        const float bscale = (p.act == YSOD_ACT_SILU && !(p.debug & 16)) ? 0.5f : 1.0f;
        for (int i = threadIdx.x; i < p.cout_pad; i += EPI_THREADS) bias_s[i] = bscale * __ldg(p.bias + (p.duo ? (i & 31) : i));   // duo: both pixels of a row
        if (p.b2b) {
            // W2 rows (64 bf16 = 128 B) in the K-major SWIZZLE_128B layout: 16 B piece c of row r at r * 128 + ((c ^ (r & 7)) << 4)
            const int w2ld = 64 + (p.b2b == 4 ? 32 : 0);   // b2b 4: W2 rows are [64 extra-operand channels | 32 staged channels]
            for (int i = threadIdx.x; i < p.b2_n * 8; i += EPI_THREADS) {
                const int r = i >> 3, c = i & 7;
                const uint4 w = __ldg(reinterpret_cast<const uint4*>(p.b2_w + (size_t)r * w2ld + c * 8));
                st_shared_v4(b2_w_smem + (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)), w.x, w.y, w.z, w.w);
            }
            if (p.b2b == 4) {   // the staged-channel columns as 64 B rows, SWIZZLE_64B: piece c of row r at r * 64 + ((c ^ ((r >> 1) & 3)) << 4)
                for (int i = threadIdx.x; i < p.b2_n * 4; i += EPI_THREADS) {
                    const int r = i >> 2, c = i & 3;
                    const uint4 w = __ldg(reinterpret_cast<const uint4*>(p.b2_w + (size_t)r * w2ld + 64 + c * 8));
                    st_shared_v4(b2_wb_smem + (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)), w.x, w.y, w.z, w.w);
                }
            }
            for (int i = threadIdx.x; i < p.b2_n; i += EPI_THREADS) b2_bias_s[i] = __ldg(p.b2_bias + i);
            fence_async_smem();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_acc;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot) : "memory");
    pdl_launch_dependents();   // the next kernel's CTAs may be scheduled as soon as SMs free up (they block in pdl_wait())
    trace(tr && threadIdx.x == 0, 2, 11, 0, 0, tcnt);   // prologue done (barriers, TMEM, bias)

    if (warp == PRODUCER_WARP) {
        // ===== TMA producer (one elected lane) =====
        if (elect_one()) {
            const int cchunks = p.cchunks;
            const bool two = p.issuers == 2;
            if (!HALO && p.pair) {
                // tile pairs: stage = {A of tile 0, A of tile 1, B}; A slots 2s / 2s + 1, B slot s
                const uint32_t tx = 2u * p.a_tx + p.b_bytes;
                const int nstages = p.stages, num_k = p.num_k, BK = p.BK, ksize = p.ksize, cstride = p.stride, pad = p.pad, Cin = p.Cin;
                int st = 0;
                uint32_t ph = 0;
                TileIter t0;
                t0.init(2 * blockIdx.x, p);
                pdl_wait();
                for (int tcount = 0; t0.valid(p); t0.step2(p), ++tcount) {
                    TileIter t1 = t0;
                    t1.step1(p);
                    const int y0 = t0.th * p.TH * cstride - pad, x0 = t0.tw * p.TW * cstride - pad;
                    const int y1 = t1.th * p.TH * cstride - pad, x1 = t1.tw * p.TW * cstride - pad;
                    const int n0 = t0.nt * p.BN;
                    int tap = 0, cc = 0, r = 0, s = 0;
                    for (int kb = 0; kb < num_k; ++kb) {
                        mbar_wait(emptyA + 8u * st, ph ^ 1u);
                        trace(tr, 0, 1, tcount, kb, tcnt);
                        const uint32_t full = fullA + 8u * st;
                        if (p.debug & 2) mbar_arrive(full);
                        else {
                            mbar_expect_tx(full, tx);
                            tma_load_4d(a_base + (uint32_t)(2 * st) * p.a_bytes, &tmA, full, cc * BK, x0 + s, y0 + r, t0.img);
                            tma_load_4d(a_base + (uint32_t)(2 * st + 1) * p.a_bytes, &tmA, full, cc * BK, x1 + s, y1 + r, t1.img);
                            tma_load_2d(b_base + (uint32_t)st * p.b_bytes, &tmB, full, tap * Cin + cc * BK, n0);
                        }
                        if (++cc == p.cchunks) {
                            cc = 0; ++tap;
                            if (++s == ksize) { s = 0; ++r; }
                        }
                        if (++st == nstages) { st = 0; ph ^= 1u; }
                    }
                }
            } else if (!HALO) {
                const uint32_t tx = p.a_tx + p.b_bytes;
                const int nstages = p.stages, G = p.kgroup, ngroups = p.num_k / p.kgroup, BK = p.BK, ksize = p.ksize, cstride = p.stride, pad = p.pad, Cin = p.Cin;
                // two sub-rings of `nstages` stages: even tiles (issuer 0) use stages [0, nstages), odd tiles [nstages, 2*nstages)
                int st_c = 0, st_o = 0, base_c = 0, base_o = two ? nstages : 0;
                uint32_t ph_c = 0, ph_o = 0;
                TileIter ti;
                ti.init(blockIdx.x, p);
                pdl_wait();   // the input activations are written by the previous kernel
                for (int tcount = 0; ti.valid(p); ti.step(p), ++tcount) {
                    const int img = ti.img;
                    const int oh0 = ti.th * p.TH, ow0 = ti.tw * p.TW;
                    const int n0 = ti.nt * p.BN;
                    int tap = 0, cc = 0, r = 0, s = 0;
                    for (int grp = 0; grp < ngroups; ++grp) {
                        const int stage = base_c + st_c;
                        mbar_wait(emptyA + 8u * stage, ph_c ^ 1u);
                        trace(tr, 0, 1, tcount, grp, tcnt);
                        const uint32_t full = fullA + 8u * stage;
                        if (p.debug & 2) mbar_arrive(full);
                        else mbar_expect_tx(full, tx * (uint32_t)G);
                        for (int g = 0; g < G; ++g) {
                            const uint32_t slot = (uint32_t)(stage * G + g);
                            if (!(p.debug & 2)) {
                                tma_load_4d(a_base + slot * p.a_bytes, &tmA, full, cc * BK, ow0 * cstride + s - pad, oh0 * cstride + r - pad, img);
                                tma_load_2d(b_base + slot * p.b_bytes, &tmB, full, tap * Cin + cc * BK, n0 + img * p.w_img_rows);
                            }
                            if (++cc == cchunks) {
                                cc = 0; ++tap;
                                if (++s == ksize) { s = 0; ++r; }
                            }
                        }
                        if (++st_c == nstages) { st_c = 0; ph_c ^= 1u; }
                    }
                    if (two) {
                        { int x = st_c; st_c = st_o; st_o = x; x = base_c; base_c = base_o; base_o = x; }
                        { uint32_t x = ph_c; ph_c = ph_o; ph_o = x; }
                    }
                }
            } else {
                // sub-rings: even tiles use A slots [0, ah) and streamed-B slots [0, bh); odd tiles the second halves
                const int ah = p.a_stages, bh = p.b_stages;
                const bool b_res = p.b_resident != 0;
                int sa_c = 0, sa_o = 0, sb_c = 0, sb_o = 0, abase_c = 0, abase_o = two ? ah : 0, bbase_c = 0, bbase_o = (b_res || !two) ? 0 : bh;
                uint32_t pa_c = 0, pa_o = 0, pb_c = 0, pb_o = 0;
                bool first = true;
                // filter columns per row of taps / K elements per tap in the weight matrix (duo: 2 pair-taps of 64)
                const int kw = p.duo ? 2 : 3, ntaps = 3 * kw, kcin = p.duo ? 64 : p.Cin;
                TileIter ti;
                ti.init(blockIdx.x, p);
                if (b_res && ti.valid(p) && p.n_tiles == 1) {
                    // resident weights are static: fetch them while the previous kernel is still draining
                    for (int cc = 0; cc < cchunks; ++cc)
                        for (int s = 0; s < kw; ++s)
                            for (int r = 0; r < 3; ++r) {
                                const int slot = cc * ntaps + s * 3 + r;
                                mbar_expect_tx(fullB + 8u * slot, p.b_bytes);
                                tma_load_2d(b_base + (uint32_t)slot * p.b_bytes, &tmB, fullB + 8u * slot, (r * kw + s) * kcin + cc * p.BK, 0);
                            }
                    first = false;
                }
                pdl_wait();   // the input activations are written by the previous kernel
                if (p.pair) {
                    // tile pairs with streamed taps: per 64-channel chunk the two tiles' halo copies, then the nine taps ONCE
                    TileIter t0;
                    t0.init(2 * blockIdx.x, p);
                    for (int tcount = 0; t0.valid(p); t0.step2(p), ++tcount) {
                        TileIter t1 = t0;
                        t1.step1(p);
                        const int n0 = t0.nt * p.BN;
                        for (int cc = 0; cc < cchunks; ++cc) {
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                const TileIter& tt = half ? t1 : t0;
                                mbar_wait(emptyA + 8u * sa_c, pa_c ^ 1u);
                                trace(tr, 0, 1, tcount, cc * 2 + half, tcnt);
                                if (p.debug & 2) mbar_arrive(fullA + 8u * sa_c);
                                else {
                                    mbar_expect_tx(fullA + 8u * sa_c, p.a_tx);
                                    tma_load_4d(a_base + (uint32_t)sa_c * p.a_bytes, &tmA, fullA + 8u * sa_c, cc * p.BK, tt.tw * 8 - 1, tt.th * 16 - 1, tt.img);
                                }
                                if (++sa_c == ah) { sa_c = 0; pa_c ^= 1u; }
                            }
                            for (int s = 0; s < 3; ++s) {
                                for (int r = 0; r < 3; ++r) {
                                    mbar_wait(emptyB + 8u * sb_c, pb_c ^ 1u);
                                    mbar_expect_tx(fullB + 8u * sb_c, p.b_bytes);
                                    tma_load_2d(b_base + (uint32_t)sb_c * p.b_bytes, &tmB, fullB + 8u * sb_c, (r * 3 + s) * p.Cin + cc * p.BK, n0);
                                    if (++sb_c == bh) { sb_c = 0; pb_c ^= 1u; }
                                }
                            }
                        }
                    }
                } else
                for (int tcount = 0; ti.valid(p); ti.step(p), ++tcount) {
                    const int img = ti.img;
                    const int oh0 = ti.th * 16, ow0 = ti.tw * 8;   // (duo: the A map's W axis counts pixel pairs, a tile is 16 rows x 8 pairs)
                    const int n0 = ti.nt * p.BN;
                    if (p.b2b == 4) {
                        // the tile's 16 x 8 pixels of the second layer's other input (64 channels, 128 B rows): slot = tile parity =
                        // the epilogue group that consumes it; freed by the commit of that group's second MMA burst
                        const int g2 = tcount & 1;
                        mbar_wait(a2_empty + 8u * g2, (uint32_t)(((tcount >> 1) & 1) ^ 1));
                        mbar_expect_tx(a2_full + 8u * g2, 16384u);
                        tma_load_4d(a2_smem + (uint32_t)g2 * 16384u, &tmA2, a2_full + 8u * g2, 0, ow0, oh0, img);
                    }
                    for (int cc = 0; cc < cchunks; ++cc) {
                        // ONE halo copy per 64-channel chunk: input rows oh0-1..oh0+16, columns ow0-1..ow0+8 (180 pixel rows of 128 B)
                        const int aslot = abase_c + sa_c;
                        mbar_wait(emptyA + 8u * aslot, pa_c ^ 1u);
                        trace(tr, 0, 1, tcount, cc, tcnt);             // A slot free
                        if (p.debug & 2) mbar_arrive(fullA + 8u * aslot);
                        else {
                            mbar_expect_tx(fullA + 8u * aslot, p.a_tx);
                            tma_load_4d(a_base + (uint32_t)aslot * p.a_bytes, &tmA, fullA + 8u * aslot, cc * p.BK, ow0 - 1, oh0 - 1, img);
                        }
                        if (++sa_c == ah) { sa_c = 0; pa_c ^= 1u; }
                        if (b_res && !first) continue;  // weights already in shared memory
                        for (int s = 0; s < kw; ++s) {
                            for (int r = 0; r < 3; ++r) {
                                const int slot = b_res ? (cc * ntaps + s * 3 + r) : (bbase_c + sb_c);
                                if (!b_res) mbar_wait(emptyB + 8u * slot, pb_c ^ 1u);
                                mbar_expect_tx(fullB + 8u * slot, p.b_bytes);
                                tma_load_2d(b_base + (uint32_t)slot * p.b_bytes, &tmB, fullB + 8u * slot, (r * kw + s) * kcin + cc * p.BK, n0);
                                if (!b_res && ++sb_c == bh) { sb_c = 0; pb_c ^= 1u; }
                            }
                        }
                    }
                    first = false;
                    if (two) {
                        { int x = sa_c; sa_c = sa_o; sa_o = x; x = sb_c; sb_c = sb_o; sb_o = x; x = abase_c; abase_c = abase_o; abase_o = x; x = bbase_c; bbase_c = bbase_o; bbase_o = x; }
                        { uint32_t x = pa_c; pa_c = pa_o; pa_o = x; x = pb_c; pb_c = pb_o; pb_o = x; }
                    }
                }
            }
        }
    } else if (warp == MMA_WARP || warp == MMA_WARP + 1) {
        // ===== MMA issuers: one elected thread per issuer warp runs the whole loop (tcgen05.mma / commit / mbarrier waits are
        //       per-thread operations; staying inside one elected branch avoids a warp reconvergence per burst).
        //       two issuers (p.issuers == 2): issuer `me` owns the CTA's tiles me, me + 2, ... and sub-ring `me`;
        //       one issuer: issuer 0 owns every tile and the whole ring. Accumulator of tile i = i % 4. =====
        const int me = warp - MMA_WARP;
        const bool two = p.issuers == 2;
        if ((me == 0 || two) && elect_one()) {
            const int trole = me ? 3 : 1;
            const uint32_t idesc = p.idesc, desc_hi = p.desc_hi;
            const uint32_t a_bytes = p.a_bytes, b_bytes = p.b_bytes;
            const bool no_mma = (p.debug & 4) != 0;
            const int BN = p.BN;
            const int astep = two ? 2 : 1;   // accumulator / tile stride of this issuer
            int aidx = me;                   // accumulator of the current tile (tile index % 4)
            uint32_t acc_phase = 0;
            TileIter ti;
            ti.init(blockIdx.x, p);
            if (me) ti.step(p);
            if (p.pair) {
                // ===== tile pairs (single issuer): accumulators 2*asel (tile 0) and 2*asel + 1 (tile 1) of the pair; every B operand
                //       landed in shared memory feeds both
                TileIter t0;
                t0.init(2 * blockIdx.x, p);
                const int ksteps = p.BK >> 4;
                int asel = 0;
                if (!HALO) {
                    const int nstages = p.stages, num_k = p.num_k;
                    int st = 0;
                    uint32_t phase = 0;
                    for (int tcount = 0; t0.valid(p); t0.step2(p), ++tcount) {
                        const uint32_t d0 = tmem_acc + (uint32_t)(2 * asel * BN), d1 = d0 + (uint32_t)BN;
                        mbar_wait(tempty_bar + 8u * (2 * asel), acc_phase ^ 1u);
                        mbar_wait(tempty_bar + 8u * (2 * asel + 1), acc_phase ^ 1u);
                        trace(tr, trole, 1, tcount, 0, tcnt);
                        tc_fence_after();
                        for (int kb = 0; kb < num_k; ++kb) {
                            mbar_wait(fullA + 8u * st, phase);
                            trace(tr, trole, 2, tcount, kb, tcnt);
                            tc_fence_after();
                            const uint32_t a0 = umma_lo(a_base + (uint32_t)(2 * st) * a_bytes), a1 = umma_lo(a_base + (uint32_t)(2 * st + 1) * a_bytes);
                            const uint32_t b_lo = umma_lo(b_base + (uint32_t)st * b_bytes);
                            if (!no_mma) {
#pragma unroll 4
                                for (int k = 0; k < ksteps; ++k)
                                    tc_mma_bf16(d0, umma_desc(desc_hi, a0 + 2u * k), umma_desc(desc_hi, b_lo + 2u * k), idesc, (uint32_t)((kb | k) != 0));
#pragma unroll 4
                                for (int k = 0; k < ksteps; ++k)
                                    tc_mma_bf16(d1, umma_desc(desc_hi, a1 + 2u * k), umma_desc(desc_hi, b_lo + 2u * k), idesc, (uint32_t)((kb | k) != 0));
                            }
                            tc_commit(emptyA + 8u * st);
                            if (kb == num_k - 1) {
                                tc_commit(tfull_bar + 8u * (2 * asel));
                                tc_commit(tfull_bar + 8u * (2 * asel + 1));
                            }
                            trace(tr, trole, 3, tcount, kb, tcnt);
                            if (++st == nstages) { st = 0; phase ^= 1u; }
                        }
                        asel ^= 1;
                        if (asel == 0) acc_phase ^= 1u;
                    }
                } else {
                    const int a_stages = p.a_stages, b_stages = p.b_stages, cchunks = p.cchunks;
                    const uint32_t row_b = 128u;   // one pixel = one 128 B row (32-channel pixels are padded, see the single-tile path)
                    const uint32_t halo_desc_hi = ((10u * row_b) >> 4) | (1u << 14) | (2u << 29);   // SBO = 10 pixels, SWIZZLE_128B
                    int sa = 0, sb = 0;
                    uint32_t pa = 0, pb = 0;
                    for (int tcount = 0; t0.valid(p); t0.step2(p), ++tcount) {
                        const uint32_t d0 = tmem_acc + (uint32_t)(2 * asel * BN), d1 = d0 + (uint32_t)BN;
                        mbar_wait(tempty_bar + 8u * (2 * asel), acc_phase ^ 1u);
                        mbar_wait(tempty_bar + 8u * (2 * asel + 1), acc_phase ^ 1u);
                        trace(tr, trole, 1, tcount, 0, tcnt);
                        tc_fence_after();
                        for (int cc = 0; cc < cchunks; ++cc) {
                            const int as0 = sa;
                            mbar_wait(fullA + 8u * as0, pa);
                            if (++sa == a_stages) { sa = 0; pa ^= 1u; }
                            const int as1 = sa;
                            mbar_wait(fullA + 8u * as1, pa);
                            if (++sa == a_stages) { sa = 0; pa ^= 1u; }
                            const uint32_t a0_addr = a_base + (uint32_t)as0 * a_bytes, a1_addr = a_base + (uint32_t)as1 * a_bytes;
                            for (int s = 0; s < 3; ++s) {
#pragma unroll
                                for (int r = 0; r < 3; ++r) mbar_wait(fullB + 8u * (sb + r), pb);
                                tc_fence_after();
                                trace(tr, trole, 2, tcount, cc * 3 + s, tcnt);
                                if (!no_mma) {
#pragma unroll
                                    for (int half = 0; half < 2; ++half) {
                                        const uint32_t dd = half ? d1 : d0;
                                        const uint32_t aa = half ? a1_addr : a0_addr;
#pragma unroll
                                        for (int r = 0; r < 3; ++r) {
                                            const uint32_t a_lo = umma_lo(aa + (uint32_t)(r * 10 + s) * row_b);
                                            const uint32_t b_lo = umma_lo(b_base + (uint32_t)(sb + r) * b_bytes);
#pragma unroll 4
                                            for (int k = 0; k < ksteps; ++k)
                                                tc_mma_bf16(dd, umma_desc(halo_desc_hi, a_lo + 2u * k), umma_desc(desc_hi, b_lo + 2u * k), idesc,
                                                            (uint32_t)((cc | s | r | k) != 0));
                                        }
                                    }
                                }
#pragma unroll
                                for (int r = 0; r < 3; ++r) tc_commit(emptyB + 8u * (sb + r));
                                sb += 3;
                                if (sb == b_stages) { sb = 0; pb ^= 1u; }
                                if (s == 2) {
                                    tc_commit(emptyA + 8u * as0);
                                    tc_commit(emptyA + 8u * as1);
                                    if (cc == cchunks - 1) {
                                        tc_commit(tfull_bar + 8u * (2 * asel));
                                        tc_commit(tfull_bar + 8u * (2 * asel + 1));
                                    }
                                }
                                trace(tr, trole, 3, tcount, cc * 3 + s, tcnt);
                            }
                        }
                        asel ^= 1;
                        if (asel == 0) acc_phase ^= 1u;
                    }
                }
            } else if (!HALO) {
                const int nstages = p.stages, G = p.kgroup, ngroups = p.num_k / p.kgroup;
                const bool k4 = (p.BK == 64);
                int st = 0;
                uint32_t phase = 0;
                const int sbase = me * nstages;   // this issuer's sub-ring (0 when there is a single issuer)
                const uint32_t a_step = a_bytes >> 4, b_step = b_bytes >> 4;
                for (int tcount = me; ti.valid(p); tcount += astep) {
                    const uint32_t d_tmem = tmem_acc + (uint32_t)(aidx * BN);
                    mbar_wait(tempty_bar + 8u * aidx, acc_phase ^ 1u);  // epilogue has drained this accumulator
                    trace(tr, trole, 1, tcount, 0, tcnt);
                    tc_fence_after();
                    for (int grp = 0; grp < ngroups; ++grp) {
                        const int stage = sbase + st;
                        mbar_wait(fullA + 8u * stage, phase);
                        trace(tr, trole, 2, tcount, grp, tcnt);
                        tc_fence_after();
                        // one burst = G K-blocks; descriptor start address advances by 32 B (>>4 = 2) per UMMA_K = 16
                        uint32_t a_lo = umma_lo(a_base + (uint32_t)(stage * G) * a_bytes);
                        uint32_t b_lo = umma_lo(b_base + (uint32_t)(stage * G) * b_bytes);
                        if (!no_mma) {
                            for (int g = 0; g < G; ++g) {
                                tc_mma_bf16(d_tmem, umma_desc(desc_hi, a_lo), umma_desc(desc_hi, b_lo), idesc, (uint32_t)((grp | g) != 0));
                                tc_mma_bf16(d_tmem, umma_desc(desc_hi, a_lo + 2u), umma_desc(desc_hi, b_lo + 2u), idesc, 1u);
                                if (k4) {
                                    tc_mma_bf16(d_tmem, umma_desc(desc_hi, a_lo + 4u), umma_desc(desc_hi, b_lo + 4u), idesc, 1u);
                                    tc_mma_bf16(d_tmem, umma_desc(desc_hi, a_lo + 6u), umma_desc(desc_hi, b_lo + 6u), idesc, 1u);
                                }
                                a_lo += a_step;
                                b_lo += b_step;
                            }
                        }
                        tc_commit(emptyA + 8u * stage);  // frees the smem slots when these MMAs retire
                        if (grp == ngroups - 1) tc_commit(tfull_bar + 8u * aidx);  // accumulator complete
                        trace(tr, trole, 3, tcount, grp, tcnt);
                        if (++st == nstages) { st = 0; phase ^= 1u; }
                    }
                    aidx += astep;
                    if (aidx >= 4) { aidx -= 4; acc_phase ^= 1u; }
                    ti.step(p);
                    if (two) ti.step(p);
                }
            } else {
                const int a_stages = p.a_stages, b_stages = p.b_stages, cchunks = p.cchunks, SG = p.sgroup;
                const bool b_res = p.b_resident != 0;
                // one pixel (duo: pixel pair) of the halo copy is ONE 128 B shared-memory row under the 128 B swizzle. A 32-channel pixel
                // fills half of its row: TMA pads an inner box dimension shorter than the swizzle span (probe: tools/ubench/
                // tma_swizzle_probe.cu), and the padded copy is the canonical conflict-free K-major layout, whereas packed 64 B rows
                // (SWIZZLE_64B) make every 8-row operand fetch hit each bank twice (83 cycles per N = 32 MMA measured, against 48).
                const uint32_t row_b = 128u;
                const uint32_t halo_desc_hi = ((10u * row_b) >> 4) | (1u << 14) | (2u << 29);   // SBO = 10 rows, SWIZZLE_128B
                const int ksteps = p.BK >> 4;
                const int kw = p.duo ? 2 : 3, ntaps = 3 * kw;   // filter columns (duo: pair-taps)
                // duo: the halo copy starts one pair (two pixels) left of the tile, the four pixels a row reads start at its second pixel
                const uint32_t a_skew = p.duo ? 64u : 0u;
                bool first = true;
                int sa = 0, sb = 0;          // positions inside this issuer's (sub-)rings
                uint32_t pa = 0, pb = 0;
                const int abase = me * a_stages, bbase = b_res ? 0 : me * b_stages;
                for (int tcount = me; ti.valid(p); tcount += astep) {
                    const uint32_t d_tmem = tmem_acc + (uint32_t)(aidx * BN);
                    mbar_wait(tempty_bar + 8u * aidx, acc_phase ^ 1u);
                    trace(tr, trole, 1, tcount, 0, tcnt);                     // accumulator free
                    tc_fence_after();
                    for (int cc = 0; cc < cchunks; ++cc) {
                        const int aslot = abase + sa;
                        mbar_wait(fullA + 8u * aslot, pa);
                        if (b_res && first) {  // resident weights land once, during the CTA's first tile
                            for (int j = 0; j < ntaps; ++j) mbar_wait(fullB + 8u * (cc * ntaps + j), 0u);
                        }
                        const uint32_t a_slot_addr = a_base + (uint32_t)aslot * a_bytes;
                        // resident weights: one burst of 36 MMAs per chunk; streamed weights: one burst of 12 per filter column
                        for (int s0 = 0; s0 < kw; s0 += SG) {
                            if (!b_res) {
#pragma unroll
                                for (int r = 0; r < 3; ++r) mbar_wait(fullB + 8u * (bbase + sb + r), pb);
                            }
                            tc_fence_after();
                            trace(tr, trole, 2, tcount, cc * 3 + s0, tcnt);   // operands landed
                            if (!no_mma) {
                                for (int s = s0; s < s0 + SG; ++s) {
                                    const uint32_t slot0 = b_res ? (uint32_t)(cc * ntaps + s * 3) : (uint32_t)(bbase + sb);
#pragma unroll
                                    for (int r = 0; r < 3; ++r) {
                                        // tap (r, s): the MMA's pixel rows start (r*10 + s) pixels into the 18 x 10 halo copy; the
                                        // 8-pixel row groups are 10 pixels (SBO = 1280 B) apart.
                                        const uint32_t a_lo = umma_lo(a_slot_addr + (uint32_t)(r * 10 + s) * row_b + a_skew);
                                        const uint32_t b_lo = umma_lo(b_base + (slot0 + r) * b_bytes);
#pragma unroll 4
                                        for (int k = 0; k < ksteps; ++k)
                                            tc_mma_bf16(d_tmem, umma_desc(halo_desc_hi, a_lo + 2u * k), umma_desc(desc_hi, b_lo + 2u * k), idesc,
                                                        (uint32_t)((cc | s | r | k) != 0));
                                    }
                                }
                            }
                            if (!b_res) {
#pragma unroll
                                for (int r = 0; r < 3; ++r) tc_commit(emptyB + 8u * (bbase + sb + r));
                                sb += 3;
                                if (sb == b_stages) { sb = 0; pb ^= 1u; }
                            }
                            if (s0 + SG == kw) {
                                tc_commit(emptyA + 8u * aslot);
                                if (cc == cchunks - 1) tc_commit(tfull_bar + 8u * aidx);
                            }
                            trace(tr, trole, 3, tcount, cc * 3 + s0, tcnt);           // burst issued + committed
                        }
                        if (++sa == a_stages) { sa = 0; pa ^= 1u; }
                    }
                    first = false;
                    aidx += astep;
                    if (aidx >= 4) { aidx -= 4; acc_phase ^= 1u; }
                    ti.step(p);
                    if (two) ti.step(p);
                }
            }
        }
    } else {
        // ===== epilogue (warps 0..15) = two groups of 8 warps; group g owns the CTA's tiles g, g + 2, ..., TMEM accumulator g
        //       and staging buffer g, so one group's per-tile bookkeeping / barriers / store issue overlap the other's math.
        //       Inside a group: TMEM lane quarter = warp % 4 (hardware rule), 16-column chunks  (warp%8)/4, +2, +4 ... =====
        // Per tile: TMEM -> registers -> (+bias, act, +residual) -> swizzled shared-memory rows of the whole BN-wide tile
        // (n_units sub-buffers of <= 128 B rows) -> ONE named barrier -> n_units TMA tensor stores. TMA clips ragged tiles
        // and the Cout padding, and writes whole lines to L2.
        const int grp = warp >> 3;
        const int q = warp & 3;
        const int cg = (warp & 7) >> 2;
        const int m = q * 32 + lane;
        const int TW = p.TW, TH = p.TH, BN = p.BN, unit_cols = p.unit_cols, n_units = p.n_units, act = (p.debug & 16) ? YSOD_ACT_NONE : p.act;
        const int th = m / TW, tw = m - th * TW;
        const bool out_f32 = p.out_f32 != 0;
        const uint32_t swz_mask = (uint32_t)p.swz_mask;
        const uint32_t unit_bytes = 128u * p.row_bytes;
        const int chunks_per_unit = unit_cols >> 4;
        const int nchunks = BN >> 4;
        const __nv_bfloat16* const res = p.res;
        const bool leader_warp = (warp & 7) == 0;
        const bool tre = tr && leader_warp && lane == 0 && grp == 0;
        const bool split = p.stage_split != 0;
        const int passes = split ? n_units : 1;
        const uint32_t sb = stage_out + (uint32_t)grp * ((uint32_t)(split ? 1 : n_units) * unit_bytes);
        const uint32_t trow0 = tmem_acc + ((uint32_t)(q * 32) << 16);
        int asel = 0;   // accumulator of tile i is i % 4 = 2 * asel + grp for this group's tiles
        const uint32_t row0 = sb + (uint32_t)m * p.row_bytes;
        const uint32_t swz_x = (row0 >> 7) & swz_mask;   // swizzle XOR term of this thread's staging row
        uint32_t acc_phase = 0, b2_phase = 0, a2_phase = 0;
        TileIter ti;
        ti.init(p.pair ? 2 * blockIdx.x + grp : blockIdx.x, p);   // pairs: group g owns tile g of every pair of this CTA
        if (!p.pair && grp) ti.step(p);
        if (res != nullptr) pdl_wait();   // the residual may be the previous kernel's output
        const int duo = p.duo;   // duo plan: row m = the pixel pair (th, 2 tw), (th, 2 tw + 1); columns [0,32) / [32,64) = their 32 channels
        for (int tcount = grp; ti.valid(p); ti.step2(p), tcount += 2) {
            const int img = ti.img;
            const int oh0 = ti.th * TH, ow0 = (ti.tw * TW) << duo;
            const int n0 = ti.nt * BN;
            // residual rows do not depend on the accumulator: fetch the first chunk's before waiting for the MMAs
            bool add_res = false;
            const __nv_bfloat16* rp = nullptr;
            uint4 ra = make_uint4(0, 0, 0, 0), rb = ra;
            if (res != nullptr) {
                const int oh = oh0 + th, ow = ow0 + (tw << duo);
                add_res = (m < TH * TW) && (oh < p.Ho) && (ow < p.Wo);
                rp = res + (((size_t)img * p.Ho + oh) * p.Wo + ow) * p.rcs + n0;
                // (duo: chunks 0, 1 are the first pixel's channels; chunk ch of the pair sits at (ch >> 1) * rcs + (ch & 1) * 16)
                if (add_res && cg < nchunks) { ra = *reinterpret_cast<const uint4*>(rp + cg * 16); rb = *reinterpret_cast<const uint4*>(rp + cg * 16 + 8); }
            }
            const int aidx = 2 * asel + grp;
            const uint32_t my_tfull = tfull_bar + 8u * aidx, my_tempty = tempty_bar + 8u * aidx;
            const uint32_t trow = trow0 + (uint32_t)(aidx * BN);
            mbar_wait(my_tfull, acc_phase);
            trace(tre, 2, 1, tcount, 0, tcnt);       // accumulator full seen
            tc_fence_after();
            if (p.debug & 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(my_tempty);
                asel ^= 1;
                if (asel == 0) acc_phase ^= 1u;
                continue;
            }
            // the bulk stores issued from this group's staging buffer (previous own tile) must have finished reading it
            const bool b2b = p.b2b != 0;
            const bool no_store = p.no_store != 0 && !b2b;   // decode-only plan: nothing is staged or stored, so no barriers either
            if (!no_store) {
                if (leader_warp) bulk_wait_read<0>();
                epi_barrier(grp);
            }
            trace(tre, 2, 2, tcount, 0, tcnt);       // staging buffer free
            float dd0 = 0.f, dd1 = 0.f;   // fused decode: DFL distances of this thread's two sides (cg = 0: left, right; cg = 1: top, bottom)
            const bool dec = p.dec_y != nullptr;
            const int d_oh = oh0 + th, d_ow = ow0 + tw;
            const bool d_ok = dec && (m < TH * TW) && d_oh < p.Ho && d_ow < p.Wo;
            float* const d_y = dec ? p.dec_y + (size_t)img * (4 + p.dec_nc) * p.dec_A + p.dec_off + d_oh * p.Wo + d_ow : nullptr;
            const int chunks_per_pass = split ? chunks_per_unit : nchunks;
            for (int ps = 0; ps < passes; ++ps) {
            if (ps > 0) {   // split staging: the previous unit's store must have read the buffer before it is overwritten
                if (leader_warp) bulk_wait_read<0>();
                epi_barrier(grp);
            }
            for (int ch = ps * chunks_per_pass + cg; ch < (ps + 1) * chunks_per_pass && ch < nchunks; ch += 2) {   // last unit may be partial (BN = 80)
                const int c0 = ch * 16;
                uint32_t v[16];
                tmem_ld16(trow + (uint32_t)c0, v);
                const float4* bs = reinterpret_cast<const float4*>(bias_s + n0 + c0);
                const float4 b0 = bs[0], b1 = bs[1], b2 = bs[2], b3 = bs[3];
                tmem_ld_wait(v);
                float f[16];
                const float bb[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
                if (act == YSOD_ACT_SILU) {
                    // x * sigmoid(x) = h + h * tanh(h) with h = x / 2: FFMA + MUFU + FFMA per element
                    const p2 half2 = pk2(0.5f, 0.5f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {   // on fp32 pairs: half the FMA issue slots of the epilogue
                        const p2 h2 = fma2(pk2u(v[2 * j], v[2 * j + 1]), half2, pk2(bb[2 * j], bb[2 * j + 1]));
                        const float2 h = up2(h2);
                        const float2 o = up2(fma2(h2, pk2(tanh_approx(h.x), tanh_approx(h.y)), h2));
                        f[2 * j] = o.x; f[2 * j + 1] = o.y;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + bb[j];
                    if (act == YSOD_ACT_GELU) act16<YSOD_ACT_GELU>(f);
                    else if (act == YSOD_ACT_RELU) act16<YSOD_ACT_RELU>(f);
                }
                if (add_res) {
                    const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float2 r2 = ysod_unpack2(rw[j]);
                        const float2 o = up2(add2(pk2(f[2 * j], f[2 * j + 1]), pk2(r2.x, r2.y)));
                        f[2 * j] = o.x; f[2 * j + 1] = o.y;
                    }
                    if (ch + 2 < nchunks) {
                        const __nv_bfloat16* rn = duo ? rp + p.rcs + (ch & 1) * 16 : rp + c0 + 32;   // duo: chunk ch + 2 = the same channels of the second pixel
                        ra = *reinterpret_cast<const uint4*>(rn); rb = *reinterpret_cast<const uint4*>(rn + 8);
                    }
                }
                if (dec) {
                    if (ch < 4) {
                        // one 16-column chunk = the 16 DFL bins of side `ch`: softmax expectation (same expression order as decode.cu)
                        float mx = f[0];
#pragma unroll
                        for (int j = 1; j < 16; ++j) mx = fmaxf(mx, f[j]);
                        float sum = 0.f, acc = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float e = expf(f[j] - mx);
                            sum += e;
                            acc += e * (float)j;
                        }
                        const float dist = acc / sum;
                        if (ch < 2) dd0 = dist; else dd1 = dist;
                    } else if (d_ok) {
                        const int nc = p.dec_nc;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if ((ch - 4) * 16 + j < nc) d_y[(size_t)(4 + (ch - 4) * 16 + j) * p.dec_A] = ysod_sigmoid(f[j]);
                    }
                }
                // swizzled store into unit u = ch / chunks_per_unit: 16-byte piece index ^= (address bits [7..]) & mask (== TMA swizzle).
                // The staging units are 1 KB aligned and a thread always writes row m, so the XOR term is a per-thread constant.
                const int u = ch / chunks_per_unit, cu = ch - u * chunks_per_unit;
                const uint32_t row_addr = row0 + (split ? 0u : (uint32_t)u * unit_bytes);
                if (no_store) continue;
                if (out_f32) {
                    const uint32_t p0 = (uint32_t)(cu * 4);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        st_shared_v4(row_addr + (((p0 + j) ^ swz_x) << 4), __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                                     __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
                } else {
                    const uint32_t p0 = (uint32_t)(cu * 2);
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        st_shared_v4(row_addr + (((p0 + j) ^ swz_x) << 4), pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                     pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
                }
            }
            if (ps == passes - 1) {
            if (d_ok) {
                // cg = 0 holds (left, right) -> cx, w ; cg = 1 holds (top, bottom) -> cy, h   (anchor = cell centre, tal.py:341-344)
                const float anc = (float)(cg == 0 ? d_ow : d_oh) + 0.5f;
                const float lo = anc - dd0, hi = anc + dd1;
                d_y[(size_t)cg * p.dec_A] = (lo + hi) * 0.5f * p.dec_stride;
                d_y[(size_t)(2 + cg) * p.dec_A] = (hi - lo) * p.dec_stride;
            }
            // all TMEM reads of this warp are complete (tcgen05.wait::ld above): hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(my_tempty);
            trace(tre, 2, 3, tcount, 0, tcnt);       // accumulator handed back
            }
            if (no_store) continue;
            fence_async_smem();  // generic-proxy smem writes -> visible to the TMA / the tensor core (async proxy)
            epi_barrier(grp);
            if (b2b) {
                // ===== back-to-back GEMM: the staged tile is the A operand of the 1x1 head conv =====
                const uint32_t d2col = (uint32_t)(4 * BN + grp * p.b2_n);
                if (leader_warp && elect_one()) {
                    const uint32_t idesc2 = (1u << 4) | (YSOD_UMMA_AB_FORMAT << 7) | (YSOD_UMMA_AB_FORMAT << 10) | ((uint32_t)(p.b2_n >> 3) << 17) | ((128u >> 4) << 24);
                    const uint32_t hi128 = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 8 rows x 128 B, SWIZZLE_128B
                    if (p.b2b == 4) {
                        // D2 = [extra operand (64 ch, TMA tile) | staged tile (32 ch, 64 B rows)] x W2^T
                        mbar_wait(a2_full + 8u * grp, a2_phase);
                        tc_fence_after();
                        const uint32_t hi64 = (512u >> 4) | (1u << 14) | (4u << 29);  // SBO = 8 rows x 64 B, SWIZZLE_64B
                        const uint32_t x_lo = umma_lo(a2_smem + (uint32_t)grp * 16384u), wa_lo = umma_lo(b2_w_smem);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_bf16(tmem_acc + d2col, umma_desc(hi128, x_lo + 2u * k), umma_desc(hi128, wa_lo + 2u * k), idesc2, (uint32_t)(k != 0));
                        const uint32_t s_lo = umma_lo(sb), wb_lo = umma_lo(b2_wb_smem);
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            tc_mma_bf16(tmem_acc + d2col, umma_desc(hi64, s_lo + 2u * k), umma_desc(hi64, wb_lo + 2u * k), idesc2, 1u);
                        tc_commit(a2_empty + 8u * grp);   // the extra-operand slot may be refilled when these MMAs retire
                    } else {
                        tc_fence_after();
                        const uint32_t a_lo = umma_lo(sb), b_lo = umma_lo(b2_w_smem);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_bf16(tmem_acc + d2col, umma_desc(hi128, a_lo + 2u * k), umma_desc(hi128, b_lo + 2u * k), idesc2, (uint32_t)(k != 0));
                    }
                    tc_commit(b2_bar + 8u * grp);
                }
                a2_phase ^= 1u;
                mbar_wait(b2_bar + 8u * grp, b2_phase);
                b2_phase ^= 1u;
                tc_fence_after();
                const int n2chunks = p.b2_n >> 4;
                if (p.b2b >= 3) {
                    // plain second layer (a 1x1 Conv + BN + act, conv.py:37-55): bias, activation, bf16, back into the SAME staging
                    // buffer (its first life, the A operand, ended when b2_bar fired) and out through the plan's own output map
                    for (int ch2 = cg; ch2 < n2chunks; ch2 += 2) {
                        uint32_t v2[16];
                        tmem_ld16(trow0 + d2col + (uint32_t)(ch2 * 16), v2);
                        tmem_ld_wait(v2);
                        float g[16];
                        if (p.b2_act == YSOD_ACT_SILU) {
                            const p2 half2 = pk2(0.5f, 0.5f);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const p2 h2 = mul2(half2, add2(pk2u(v2[2 * j], v2[2 * j + 1]), pk2(b2_bias_s[ch2 * 16 + 2 * j], b2_bias_s[ch2 * 16 + 2 * j + 1])));
                                const float2 hh = up2(h2);
                                const float2 o = up2(fma2(h2, pk2(tanh_approx(hh.x), tanh_approx(hh.y)), h2));
                                g[2 * j] = o.x; g[2 * j + 1] = o.y;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) g[j] = __uint_as_float(v2[j]) + b2_bias_s[ch2 * 16 + j];
                        }
                        // kind 3 re-uses the staging buffer (128 B rows); kind 4 has its own 128 B-row output staging tile
                        const uint32_t orow = p.b2b == 4 ? stage2_smem + (uint32_t)grp * 16384u + (uint32_t)m * 128u : row0;
                        const uint32_t oswz = p.b2b == 4 ? (uint32_t)(m & 7) : swz_x;
                        const uint32_t p0 = (uint32_t)(ch2 * 2);
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            st_shared_v4(orow + (((p0 + j) ^ oswz) << 4), pack_bf16(g[8 * j], g[8 * j + 1]), pack_bf16(g[8 * j + 2], g[8 * j + 3]),
                                         pack_bf16(g[8 * j + 4], g[8 * j + 5]), pack_bf16(g[8 * j + 6], g[8 * j + 7]));
                    }
                    tc_fence_before();
                    fence_async_smem();
                    epi_barrier(grp);
                    if (leader_warp && elect_one()) {
                        if (p.b2b == 4) tma_store_4d(&tmO2, stage2_smem + (uint32_t)grp * 16384u, 0, ow0, oh0, img);
                        else tma_store_4d(&tmO, sb, 0, ow0, oh0, img);
                        bulk_commit();
                    }
                    continue;
                }
                float* const rawp = (p.b2_raw != nullptr && (m < TH * TW) && d_oh < p.Ho && d_ow < p.Wo)
                                        ? p.b2_raw + (((size_t)img * p.Ho + d_oh) * p.Wo + d_ow) * p.b2_raw_cs + p.b2_raw_off : nullptr;
                const bool y_ok = (m < TH * TW) && d_oh < p.Ho && d_ow < p.Wo;
                float* const yb = p.b2_y + (size_t)img * (4 + p.dec_nc) * p.dec_A + p.dec_off + d_oh * p.Wo + d_ow;
                float e0 = 0.f, e1 = 0.f;   // box branch: DFL distances of this thread's two sides (cg 0: left, right; cg 1: top, bottom)
                for (int ch2 = cg; ch2 < n2chunks; ch2 += 2) {
                    uint32_t v2[16];
                    tmem_ld16(trow0 + d2col + (uint32_t)(ch2 * 16), v2);
                    tmem_ld_wait(v2);
                    float g[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) g[j] = __uint_as_float(v2[j]) + b2_bias_s[ch2 * 16 + j];
                    if (rawp) {
                        const int valid = p.b2b == 1 ? 16 : min(16, p.dec_nc - ch2 * 16);   // the class branch has nc (< 16 * chunks) real channels
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4)
                            if (j4 * 4 + 3 < valid) *reinterpret_cast<float4*>(rawp + ch2 * 16 + j4 * 4) = make_float4(g[4 * j4], g[4 * j4 + 1], g[4 * j4 + 2], g[4 * j4 + 3]);
                            else
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (j4 * 4 + j < valid) rawp[ch2 * 16 + j4 * 4 + j] = g[4 * j4 + j];
                    }
                    if (p.b2b == 1) {
                        // one 16-column chunk = the 16 DFL bins of side ch2 (l, t, r, b): softmax expectation, as in decode.cu
                        float mx = g[0];
#pragma unroll
                        for (int j = 1; j < 16; ++j) mx = fmaxf(mx, g[j]);
                        float sum = 0.f, acc = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float e = __expf(g[j] - mx);   // ex2.approx (2^-21 relative): this epilogue shares the SM with a tensor-bound main loop
                            sum += e;
                            acc = fmaf(e, (float)j, acc);
                        }
                        const float dist = __fdividef(acc, sum);
                        if (ch2 < 2) e0 = dist; else e1 = dist;
                    } else if (y_ok) {
                        const int nc = p.dec_nc;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (ch2 * 16 + j < nc) yb[(size_t)(4 + ch2 * 16 + j) * p.dec_A] = __fdividef(1.0f, 1.0f + __expf(-g[j]));
                    }
                }
                if (p.b2b == 1 && y_ok) {
                    const float anc = (float)(cg == 0 ? d_ow : d_oh) + 0.5f;
                    const float lo = anc - e0, hi = anc + e1;
                    yb[(size_t)cg * p.dec_A] = (lo + hi) * 0.5f * p.dec_stride;
                    yb[(size_t)(2 + cg) * p.dec_A] = (hi - lo) * p.dec_stride;
                }
                tc_fence_before();   // this warp's reads of D2 are complete before the group's next barrier lets the next MMA overwrite it
                continue;
            }
            if (leader_warp && elect_one()) {
                if (!(p.debug & 8)) {
                    if (split) {
                        if (p.up2) {
#pragma unroll
                            for (int d = 0; d < 4; ++d) tma_store_4d(&tmO, sb, n0 + ps * unit_cols, 2 * ow0 + (d & 1), 2 * oh0 + (d >> 1), img);
                        } else {
                            tma_store_4d(&tmO, sb, n0 + ps * unit_cols, ow0, oh0, img);
                        }
                    } else if (p.up2) {
                        // the output map walks the 2x-upsampled destination with element strides {1,2,2,1}: four stores of the
                        // same staged tile, one per (dy, dx) phase of the 2 x 2 replication
                        for (int u = 0; u < n_units; ++u)
#pragma unroll
                            for (int d = 0; d < 4; ++d)
                                tma_store_4d(&tmO, sb + (uint32_t)u * unit_bytes, n0 + u * unit_cols, 2 * ow0 + (d & 1), 2 * oh0 + (d >> 1), img);
                    } else if (duo) {
                        for (int u = 0; u < 2; ++u) tma_store_4d(&tmO, sb + (uint32_t)u * unit_bytes, 0, ow0 + u, oh0, img);   // unit = pixel parity
                    } else {
                        for (int u = 0; u < n_units; ++u) tma_store_4d(&tmO, sb + (uint32_t)u * unit_bytes, n0 + u * unit_cols, ow0, oh0, img);
                    }
                }
                bulk_commit();
            }
            }   // passes
            trace(tre, 2, 4, tcount, 0, tcnt);       // stores issued
            asel ^= 1;
            if (asel == 0) acc_phase ^= 1u;
        }
        if (leader_warp) bulk_wait_read<0>();  // smem must stay valid until the last bulk store has read it
    }

    tc_fence_before();
    __syncthreads();
    trace(tr && threadIdx.x == 0, 2, 12, 0, 0, tcnt);   // all roles finished (stores read out)
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// Weight matrix of the pixel-duo plan (TcParams::duo): w [32][3][3][32] (16-bit elements, K ordered (r, s, cin)) ->
// wd [64][3][2][2][32]: row n = output pixel (n >> 5) of the pair, channel n & 31; K = (filter row r, pair-tap t, pixel e of that pair, cin).
// Halo pixel h = 2t + e of the four a row reads feeds output pixel j through filter column s = h - j (zero when s is outside 0..2).
__global__ void duo_weights_kernel(const uint16_t* __restrict__ w, uint16_t* __restrict__ wd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * 384) return;
    const int n = i / 384, k = i - n * 384;
    const int r = k >> 7, t = (k >> 6) & 1, e = (k >> 5) & 1, ci = k & 31;
    const int s = 2 * t + e - (n >> 5);
    wd[i] = (s >= 0 && s < 3) ? w[(n & 31) * 288 + (r * 3 + s) * 32 + ci] : (uint16_t)0;
}

__global__ void scale_weights_kernel(const float* __restrict__ w, int rows, int K, int Cin, const float* __restrict__ gate,
                                     __nv_bfloat16* __restrict__ out, long long total) {
    ysod_pdl_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int k = (int)(i % K);
    const long long rk = (long long)rows * K;
    const int n = (int)(i / rk);
    out[i] = __float2bfloat16_rn(w[i - (long long)n * rk] * gate[(size_t)n * Cin + (k % Cin)]);
}

// ---- host side --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

struct ConvTc {
    bool halo;
    void* owned_w;   // duo plan: the repacked weight matrix (device memory owned by the plan)
    CUtensorMap tmA, tmB, tmO, tmA2, tmO2;
    TcParams p;
    dim3 grid;
    size_t smem;
};

}  // namespace

struct ysod_conv_tc { ConvTc c; };

extern "C" {

// Creates a launch plan (TMA descriptors + tiling) for one conv / linear layer with fixed device pointers.
//   x      : NHWC bf16 input view, N x H x W x Cin, pixel stride xcs elements (xcs % 8 == 0, 16 B aligned)
//   wgt    : [Cout_pad][ksize*ksize*Cin] bf16, K ordered (r, s, cin); Cout_pad % 16 == 0, rows >= Cout are zero
//   bias   : [Cout_pad] fp32 (BN folded / conv bias / zeros)
//   out    : NHWC view, pixel stride ocs elements; out_dtype YSOD_BF16 or YSOD_F32
//   res    : optional NHWC bf16 residual added AFTER the activation (Bottleneck / transformer skip), stride rcs
// ksize in {1,3}, stride in {1,2} (pad = ksize/2, conv.py:28 autopad), groups == 1, Cin % 32 == 0.
// 3x3 / stride-1 convs with 32 input and 32 output channels (bf16, even width) run the halo kernel's pixel-duo plan (TcParams::duo)
// unless mode carries 0x10000 (YSOD_CONV_NO_DUO).
// mode: 0 = auto, 1 = generic per-tap kernel, 2 = force the 3x3 halo-reuse kernel (error if the shape does not qualify);
//       | 0x40 = fuse nn.Upsample(scale_factor=2, mode='nearest') into the store: `out` is the N x 2Ho x 2Wo destination view.
//       | 0x80 = per-image weights: `wgt` is [N][Cout_pad][K] (ysod_scale_weights: an SE channel gate folded into this conv).
int ysod_conv_tc_create_ex(ysod_conv_tc** handle, const void* x, int N, int H, int W, int Cin, int xcs, const void* wgt,
                           const float* bias, int Cout, int Cout_pad, int ksize, int stride, void* out, int out_dtype, int ocs,
                           const void* res, int rcs, int act, int mode) {
    YSOD_CHECK_ARG(handle && x && wgt && bias && out, "ysod_conv_tc_create: null pointer");
    YSOD_CHECK_ARG(ksize == 1 || ksize == 3, "ysod_conv_tc_create: ksize %d unsupported", ksize);
    YSOD_CHECK_ARG(stride == 1 || stride == 2, "ysod_conv_tc_create: stride %d unsupported", stride);
    YSOD_CHECK_ARG(Cin % 32 == 0 && Cin >= 32, "ysod_conv_tc_create: Cin %d must be a multiple of 32", Cin);
    YSOD_CHECK_ARG(Cout_pad % 16 == 0 && Cout <= Cout_pad && Cout > 0, "ysod_conv_tc_create: bad Cout %d / pad %d", Cout, Cout_pad);
    YSOD_CHECK_ARG(xcs % 8 == 0 && ((uintptr_t)x % 16) == 0, "ysod_conv_tc_create: input view not 16 B aligned");
    YSOD_CHECK_ARG(out_dtype == YSOD_BF16 || out_dtype == YSOD_F32, "ysod_conv_tc_create: bad out dtype");
    const int oalign = out_dtype == YSOD_BF16 ? 8 : 4;
    YSOD_CHECK_ARG(ocs % oalign == 0 && ((uintptr_t)out % 16) == 0, "ysod_conv_tc_create: output view not 16 B aligned");
    YSOD_CHECK_ARG(!res || (rcs % 8 == 0 && ((uintptr_t)res % 16) == 0), "ysod_conv_tc_create: residual view not 16 B aligned");
    YSOD_CHECK_ARG(stride == 1 || (H % 2 == 0 && W % 2 == 0), "ysod_conv_tc_create: stride-2 needs even H, W");
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        ysod_set_error("ysod_conv_tc_create: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return YSOD_ERR_CUDA;
    }
    ConvTc c;
    memset(&c, 0, sizeof(c));
    TcParams& p = c.p;
    const int pad = ksize / 2;
    const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
    // pick the output patch TH x TW (<= 128 pixels) with the best row utilisation of the M = 128 MMA
    int bestTW = 1, bestTH = 1;
    double best = -1.0;
    const int maxTW = Wo < 128 ? Wo : 128;
    for (int tw = 1; tw <= maxTW; ++tw) {
        int th = 128 / tw;
        if (th > Ho) th = Ho;
        if (th > 128) th = 128;
        const long long tiles = (long long)ysod_cdiv(Ho, th) * ysod_cdiv(Wo, tw);
        const double util = (double)Ho * Wo / ((double)tiles * 128.0);
        if (util > best + 1e-9 || (util > best - 1e-9 && tw > bestTW && tw <= 32)) { best = util; bestTW = tw; bestTH = th; }
    }
    // 3x3 / stride-1 halo-reuse specialisation: fixed 16 x 8 output tile; worth it when that tiling wastes < 25 % of M
    bool halo = false;
    const int dbg = (mode >> 8) & 0xff;
    const bool no_duo = (mode & 0x10000) != 0;
    const bool up2 = (mode & 0x40) != 0;
    const bool img_w = (mode & 0x80) != 0;
    const int split_exp = mode & 0x10;
    const bool no_store = (mode & 0x20) != 0;
    const bool no_pair = (mode & 0x04) != 0;
    const bool force_pair = (mode & 0x08) != 0;
    mode &= 0x03;
    if (img_w) {
        YSOD_CHECK_ARG(mode != 2, "ysod_conv_tc_create_ex: per-image weights need the generic kernel (resident taps are shared by all images)");
        mode = 1;
    }
    if (ksize == 3 && stride == 1 && (Cin % 64 == 0 || Cin == 32)) {
        const double hutil = (double)Ho * Wo / ((double)ysod_cdiv(Ho, 16) * ysod_cdiv(Wo, 8) * 128.0);
        halo = (mode == 2) || (mode == 0 && hutil >= 0.75);
    }
    YSOD_CHECK_ARG(mode != 2 || halo, "ysod_conv_tc_create_ex: shape does not qualify for the halo kernel");
    // pixel-duo plan (TcParams::duo): 32 -> 32 channels, bf16 out, even width, 16 x 16 pixel tiles that waste < 25 % of the MMA rows
    bool duo = false;
    // (xcs == 32: the two pixels of a pair must be contiguous -- TMA cannot pack two strided 64 B pixels into one 128 B row)
    if (halo && !no_duo && Cin == 32 && xcs == 32 && Cout == 32 && Cout_pad == 32 && out_dtype == YSOD_BF16 && !up2 && !no_store && W % 2 == 0) {
        const double dutil = (double)Ho * Wo / ((double)ysod_cdiv(Ho, 16) * ysod_cdiv(Wo, 16) * 256.0);
        duo = dutil >= 0.75;
    }
    if (duo) Cout_pad = 64;   // the GEMM's N: two pixels x 32 channels (bias duplicated by the kernel prologue)
    p.duo = duo ? 1 : 0;
    p.debug = dbg;
    p.up2 = up2 ? 1 : 0;
    p.no_store = no_store ? 1 : 0;
    p.w_img_rows = img_w ? Cout_pad : 0;
    if (halo) { bestTH = 16; bestTW = 8; }
    p.N = N; p.Ho = Ho; p.Wo = Wo; p.TH = bestTH; p.TW = bestTW;
    p.tiles_h = ysod_cdiv(Ho, bestTH); p.tiles_w = ysod_cdiv(Wo, bestTW << (duo ? 1 : 0));   // duo: TW counts pixel pairs
    p.Cin = Cin; p.ksize = ksize; p.stride = stride; p.pad = pad;
    p.BK = (Cin % 64 == 0 || duo) ? 64 : 32;
    int BN = Cout_pad;
    if (BN > 128) {
        // <= 128 output channels per tile: the two epilogue groups each stage a whole tile (<= 32 KB), wide layers get more
        // (smaller) tiles to spread over the 148 SMs, and N = 128 MMAs already run at 93 % of the N = 256 rate
        BN = 128;
        while (Cout_pad % BN != 0) BN >>= 1;  // 128 / 64 / 32 / 16: with several N tiles every store unit must be full width
    }
    p.BN = BN;
    p.tmem_cols = 4 * BN <= 32 ? 32 : 4 * BN <= 64 ? 64 : 4 * BN <= 128 ? 128 : 4 * BN <= 256 ? 256 : 512;  // 2 issuers x double-buffered accumulator
    p.n_tiles = Cout_pad / BN;
    p.num_k = duo ? 6 : ksize * ksize * (Cin / p.BK);
    p.Cout = Cout;
    p.out = out; p.out_f32 = (out_dtype == YSOD_F32); p.ocs = ocs;
    p.bias = bias; p.res = (const __nv_bfloat16*)res; p.rcs = rcs; p.act = act;
    p.a_bytes = 128u * p.BK * 2u;
    p.b_bytes = (uint32_t)BN * p.BK * 2u;
    p.a_tx = (uint32_t)(bestTH * bestTW) * p.BK * 2u;
    // instruction descriptor (kind::f16): D=f32, A=B=bf16, both K-major, N>>3 @17, M>>4 @24
    p.idesc = (1u << 4) | (YSOD_UMMA_AB_FORMAT << 7) | (YSOD_UMMA_AB_FORMAT << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
    // smem descriptor high word: SBO (8 rows x swizzle span) >> 4 @ bits 32-45, version 1 @ 46, layout @ 61
    const uint32_t sbo = (p.BK == 64 ? 1024u : 512u) >> 4;
    const uint32_t layout = (p.BK == 64) ? 2u : 4u;  // SWIZZLE_128B : SWIZZLE_64B
    p.desc_hi = sbo | (1u << 14) | (layout << 29);
    p.cchunks = duo ? 1 : Cin / p.BK;
    {
        const uint32_t es = p.out_f32 ? 4u : 2u;
        uint32_t rb = 32;
        while (rb < 128 && rb < (uint32_t)BN * es) rb <<= 1;
        p.row_bytes = rb;
        p.unit_cols = (int)(rb / es);
        p.n_units = (int)(((uint32_t)BN * es + rb - 1) / rb);
        p.swz_mask = rb == 128 ? 7 : rb == 64 ? 3 : 1;
        p.cout_pad = Cout_pad;
        if (duo) {
            // the output view may be strided (a slice of a concat buffer), so a staged row cannot leave as one pixel pair: two store
            // units of 64 B rows, unit e = pixel e of every pair, each stored through a W axis walked with element stride 2
            p.row_bytes = 64; p.unit_cols = 32; p.n_units = 2; p.swz_mask = 3;
        }
    }
    // epilogue staging holds the whole BN-wide output tile (n_units sub-buffers of 128 rows); double-buffered up to 32 KB
    uint32_t tile_stage_bytes = (uint32_t)p.n_units * 128u * p.row_bytes;
    p.stage_bufs = 2;   // one per epilogue group
    p.stage_split = 0;
    if (halo && p.n_units > 1 && !duo) {
        // 3x3 halo plan whose weight taps only fit resident if the staging buffer shrinks to one store unit (64 -> 128 at P2:
        // 144 KB of taps): streamed taps are bound by the latency x depth of the small tap ring (measured 45 % tensor-pipe
        // activity), so trade a second barrier per tile for resident weights. Exact shared-memory accounting, not the 224 KB rule.
        const uint32_t a_halo = ((18u * 10u * 128u) + 1023u) & ~1023u;
        const uint32_t b_all = 9u * (uint32_t)p.cchunks * p.b_bytes;   // (never the duo plan: its N = 64 tile is one store unit)
        const uint32_t nbar_res = 2u * 2u + 2u * 9u * (uint32_t)p.cchunks;
        const uint32_t full_need = b_all + 2u * a_halo + 2u * tile_stage_bytes + 4u * (uint32_t)Cout_pad + 3u * 1024u;
        const uint32_t split_need = b_all + 2u * a_halo + 1024u + (8u * nbar_res + 128u) + 4u * (uint32_t)Cout_pad + 1024u + 2u * 128u * p.row_bytes;
        if (full_need > 224u * 1024u && split_need <= 227u * 1024u) {
            p.stage_split = 1;
            tile_stage_bytes = 128u * p.row_bytes;
        } else if (full_need > 224u * 1024u && !split_exp) {
            // taps cannot be resident (Cin >= 128 at BN = 128): they stream through their own ring, whose depth x L2 latency bounds
            // the layer. Split staging frees 32 KB = a third filter column of taps in flight (9 slots instead of 6).
            p.stage_split = 2;   // 2 = split for the streamed-tap ring
            tile_stage_bytes = 128u * p.row_bytes;
        }
    }
    if (!halo && p.n_units > 1 && p.num_k >= 4 && !(split_exp)) {
        // deep-K generic layers (20^2 / 40^2 maps, 1x1 with K >= 256) are bound by the latency x depth of the operand ring (L2 -> SM,
        // measured ~66 GB/s per SM with 4 slots in flight): a one-unit staging buffer frees 32 KB = a fifth {A,B} slot, +6..10 % on
        // every such layer (256->256 3x3 @20^2: 34.8 -> 32.8 us, 256->512 3x3/s2: 59.4 -> 53.2, 512->256 1x1 @40^2: 38.9 -> 34.8).
        // mode bit 0x10 disables it (A/B measurements).
        p.stage_split = 1;
        tile_stage_bytes = 128u * p.row_bytes;
    }
    const uint32_t staging = (uint32_t)p.stage_bufs * tile_stage_bytes;
    const uint32_t fixed = staging + 4u * (uint32_t)Cout_pad + 3u * 1024u;   // staging + bias + slack
    size_t ring_bytes = 0;
    int nbar = 0;
    // p.stages / p.a_stages / p.b_stages count stages PER (SUB-)RING: with two issuers the ring is split into one sub-ring per
    // issuer (tile parity), with one issuer there is a single ring. p.a_slots / p.b_slots are the smem slots carved in total.
    if (!halo) {
        // Pipeline stage = G K-blocks ("slots") behind one mbarrier pair: with narrow N an MMA retires in ~N/2 cycles, so a
        // stage must carry enough MMAs (>= ~512 tensor-pipe cycles) to amortise the issuer's per-handshake cost.
        // Tile pairs (see TcParams::pair): deep-K layers whose operand stream, not the tensor pipe, bounds them. Needs an even number
        // of spatial tiles per N tile (a pair never straddles two weight tiles), shared (not per-image) weights, >= 2 tiles per CTA.
        const long long spatial = (long long)N * p.tiles_h * p.tiles_w;
        // Measured (profiles/r02_conv_pair_ab.txt): the generic kernel's steady state gains 24 % per tile (512->256 1x1 @40^2: 4870 ->
        // 3700 cycles), but these launches are 25-35 us long with ~3 pairs per CTA, so the coarser tail (pairs quantise the last wave
        // twice as hard) and the larger first stage cancel it: off unless forced (YSOD_CONV_FORCE_PAIR). The streamed-tap 3x3 plan
        // below, where the shared operand is 85 % of the bytes, keeps it on.
        p.pair = (force_pair && !no_pair && !img_w && p.num_k >= 4 && BN >= 64 && spatial % 2 == 0 && spatial * p.n_tiles >= 4) ? 1 : 0;
        if (p.pair) {
            const uint32_t stage_bytes = 2u * p.a_bytes + p.b_bytes;
            int stages = (int)((224u * 1024u - fixed) / stage_bytes);
            if (stages > 6) stages = 6;
            if (stages < 2) p.pair = 0;
            else {
                p.issuers = 1;
                p.kgroup = 1;
                p.sgroup = 1;
                p.stages = p.a_stages = p.b_stages = stages;
                p.a_slots = 2 * stages;
                p.b_slots = stages;
                ring_bytes = (size_t)stages * stage_bytes;
                nbar = 2 * p.a_slots + 2 * p.b_slots;
            }
        }
        const uint32_t slot_bytes = p.a_bytes + p.b_bytes;
        const uint32_t budget = 224u * 1024u - fixed;   // one fat CTA per SM
        int slots = (int)(budget / slot_bytes);
        if (slots > 16) slots = 16;
        const int kb_cycles = (p.BK / 16) * (BN / 2);
        // two issuers only pay off for tiles with a single short burst (1x1 convs with K <= 64..128)
        if (!p.pair) p.issuers = (p.num_k * kb_cycles <= 512 && slots >= 2 * p.num_k) ? 2 : 1;
        const int rings = p.issuers;
        int gmax = 512 / kb_cycles;
        if (gmax < 1) gmax = 1;
        if (gmax > slots / (2 * rings)) gmax = slots / (2 * rings);   // keep >= 2 stages per (sub-)ring
        if (gmax < 1) gmax = 1;
        int G = 1;
        for (int g = 1; g <= gmax; ++g)
            if (p.num_k % g == 0) G = g;
        if (p.stage_split && slots % (rings * G) != 0) G = 1;   // use the slot the split staging freed: 5 single-block stages, not 2 x 2
        int stages = slots / (rings * G);
        if (stages > 8) stages = 8;
        if (stages < 1) stages = 1;
        if (!p.pair) {
            p.kgroup = G;
            p.sgroup = 1;
            p.stages = p.a_stages = p.b_stages = stages;
            p.a_slots = p.b_slots = rings * stages * G;
            ring_bytes = (size_t)p.a_slots * slot_bytes;
            nbar = 4 * p.a_slots;
        }
    } else {
        p.a_tx = 18u * 10u * 2u * (uint32_t)p.BK;     // bytes landed per halo copy: 18 rows x 10 px x BK ch bf16 (22.5 KB / 11.25 KB)
        p.a_bytes = (18u * 10u * 128u + 1023u) & ~1023u;   // slot stride: one 128 B row per pixel (32-channel pixels fill half a row)
        const uint32_t ntaps = duo ? 6u : 9u;         // duo: 3 rows x 2 pair-taps; its halo copy is 18 rows x 10 pixel pairs of 128 B
        const uint32_t b_all = ntaps * (uint32_t)p.cchunks * p.b_bytes;
        // resident split plan: sized exactly above; streamed split plan: exact accounting against the 227 KB limit (<= 13 barrier
        // pairs, bias, two 1 KB alignment pads, staging) instead of the conservative 224 KB rule -- the ninth tap slot needs it
        const uint32_t avail = p.stage_split == 1 ? b_all + 2u * p.a_bytes
                             : p.stage_split == 2 ? 227u * 1024u - (2048u + 8u * 26u + 128u + 4u * (uint32_t)Cout_pad + staging)
                                                  : 224u * 1024u - fixed;
        int a_total;
        if (b_all + 2u * p.a_bytes <= avail) {   // all weight taps stay resident in shared memory for the CTA lifetime
            p.b_resident = 1;
            p.issuers = 2;            // one 36-MMA burst per tile and chunk: ping-pong two issuers over the tiles
            p.b_stages = (int)ntaps * p.cchunks;
            p.b_slots = p.b_stages;
            a_total = (int)((avail - b_all) / p.a_bytes);
            if (a_total > 8) a_total = 8;
            a_total &= ~1;
            p.a_stages = a_total / 2;
            p.sgroup = duo ? 2 : 3;
        } else {
            YSOD_CHECK_ARG(!duo, "ysod_conv_tc_create: the duo plan keeps its weights resident");
            p.b_resident = 0;
            p.issuers = 1;            // many bursts per tile: one issuer, one deep ring
            p.b_stages = (9u * p.b_bytes + 2u * p.a_bytes <= avail) ? 9 : 6;   // three (else two) filter columns of three taps in flight
            {
                // tile pairs need the two tiles' halo copies of a chunk resident at once; with room for only two copies the next
                // chunk's halos could not be prefetched, so a pair plan trades the third tap column for four halo slots
                const long long spatial = (long long)N * p.tiles_h * p.tiles_w;
                const bool pair_ok = !no_pair && spatial % 2 == 0 && spatial * p.n_tiles >= 4;
                if (pair_ok && p.b_stages == 9 && (avail - 9u * p.b_bytes) / p.a_bytes < 4u && 6u * p.b_bytes + 4u * p.a_bytes <= avail) p.b_stages = 6;
            }
            p.b_slots = p.b_stages;
            YSOD_CHECK_ARG((uint32_t)p.b_stages * p.b_bytes + 2u * p.a_bytes <= avail, "ysod_conv_tc_create: halo plan does not fit in shared memory (BN %d)", BN);
            a_total = (int)((avail - (uint32_t)p.b_stages * p.b_bytes) / p.a_bytes);
            if (a_total > 4) a_total = 4;
            p.a_stages = a_total;
            p.sgroup = 1;
            // streamed taps are the dominant operand stream (9 x 16 KB per chunk against one 22.5 KB halo copy): share them
            // between the two tiles of a pair (TcParams::pair)
            const long long spatial = (long long)N * p.tiles_h * p.tiles_w;
            p.pair = (!no_pair && a_total >= 2 && spatial % 2 == 0 && spatial * p.n_tiles >= 4) ? 1 : 0;
        }
        p.a_slots = a_total;
        ring_bytes = (size_t)p.a_slots * p.a_bytes + (size_t)p.b_slots * p.b_bytes;
        nbar = 2 * p.a_slots + 2 * p.b_slots;
        p.kgroup = 1;
        p.stages = p.a_stages;
    }
    // ring + 1 KB alignment slack + barriers/slot + bias + alignment + output staging
    c.smem = ring_bytes + 1024 + (8 * (size_t)nbar + 128) + 4 * (size_t)Cout_pad + 1024 + staging;
    c.halo = halo;
    {
        int dev = 0, sms = 148;
        YSOD_CUDA(cudaGetDevice(&dev));
        YSOD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const long long total = (long long)N * p.tiles_h * p.tiles_w * p.n_tiles;
        long long g = (long long)sms;   // one persistent CTA per SM
        const long long units = p.pair ? total / 2 : total;   // work items: tiles, or tile pairs
        if (g > units) g = units;
        c.grid = dim3((unsigned)g, 1, 1);
        long long r = g;
        p.step_tw = (int)(r % p.tiles_w); r /= p.tiles_w;
        p.step_th = (int)(r % p.tiles_h); r /= p.tiles_h;
        p.step_img = (int)(r % N);
        p.step_nt = (int)(r / N);
        r = 2 * g;
        p.step2_tw = (int)(r % p.tiles_w); r /= p.tiles_w;
        p.step2_th = (int)(r % p.tiles_h); r /= p.tiles_h;
        p.step2_img = (int)(r % N);
        p.step2_nt = (int)(r / N);
    }

    const CUtensorMapSwizzle swz = p.BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)xcs * 2, (cuuint64_t)W * xcs * 2, (cuuint64_t)H * W * xcs * 2};
        if (duo) { dims[0] = 64; dims[1] = (cuuint64_t)(W / 2); strides[0] = 128; }   // W axis in pixel pairs (contiguous: xcs == 32)
        cuuint32_t box[4] = {(cuuint32_t)p.BK, (cuuint32_t)(bestTW * stride), (cuuint32_t)(bestTH * stride), 1};
        if (halo) { box[1] = 10; box[2] = 18; }   // the halo copy: rows oh0-1 .. oh0+16, columns ow0-1 .. ow0+8
        cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        // the halo copy is always 128 B-swizzled rows of 128 B: a 32-channel pixel is padded to a row by TMA, a duo row is a pixel pair
        CUresult r = enc(&c.tmA, YSOD_TMAP_16, 4, const_cast<void*>(x), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, halo ? CU_TENSOR_MAP_SWIZZLE_128B : swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            ysod_set_error("ysod_conv_tc_create: cuTensorMapEncodeTiled(A) failed with %d (Cin %d W %d H %d N %d xcs %d box %d,%d,%d)",
                           (int)r, Cin, W, H, N, xcs, p.BK, bestTW * stride, bestTH * stride);
            return YSOD_ERR_CUDA;
        }
    }
    struct OwnedGuard {   // frees the plan-owned weight copy on every error return below
        void*& ptr; bool keep;
        ~OwnedGuard() { if (!keep && ptr) { cudaFree(ptr); ptr = nullptr; } }
    } owned_guard{c.owned_w, false};
    if (duo) {
        YSOD_CUDA(cudaMalloc(&c.owned_w, 64 * 384 * 2));
        duo_weights_kernel<<<(64 * 384 + 255) / 256, 256>>>((const uint16_t*)wgt, (uint16_t*)c.owned_w);
        cudaError_t e = cudaDeviceSynchronize();   // plan creation is not on the hot path; `wgt` need not outlive it
        if (e != cudaSuccess) {
            ysod_set_error("ysod_conv_tc_create: duo weight repack failed: %s", cudaGetErrorString(e));
            return YSOD_ERR_CUDA;
        }
        wgt = c.owned_w;
    }
    {
        const cuuint64_t K = duo ? 384 : (cuuint64_t)ksize * ksize * Cin;
        cuuint64_t dims[2] = {K, (cuuint64_t)Cout_pad * (cuuint64_t)(img_w ? N : 1)};
        cuuint64_t strides[1] = {K * 2};
        cuuint32_t box[2] = {(cuuint32_t)p.BK, (cuuint32_t)BN};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&c.tmB, YSOD_TMAP_16, 2, const_cast<void*>(wgt), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            ysod_set_error("ysod_conv_tc_create: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
            return YSOD_ERR_CUDA;
        }
    }
    {
        // output map: channel slice [Cout] x Wo x Ho x N, box = {unit_cols, TW, TH, 1}; stores clip ragged tiles / Cout padding
        const cuuint64_t es = p.out_f32 ? 4 : 2;
        const cuuint64_t us = up2 ? 2 : 1;   // fused nearest upsample: the destination is 2Ho x 2Wo, traversed with element stride 2
        cuuint64_t dims[4] = {(cuuint64_t)Cout, us * Wo, us * Ho, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)ocs * es, us * Wo * ocs * es, us * Ho * us * Wo * ocs * es};
        cuuint32_t box[4] = {(cuuint32_t)p.unit_cols, (cuuint32_t)(us * bestTW), (cuuint32_t)(us * bestTH), 1};
        cuuint32_t es1[4] = {1, (cuuint32_t)us, (cuuint32_t)us, 1};
        if (duo) { box[1] = 16; es1[1] = 2; }   // store unit e = pixel e of the tile's 8 pairs per row: 16 pixels walked with stride 2
        const CUtensorMapSwizzle oswz = p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                      : p.row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
        CUresult r = enc(&c.tmO, p.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : YSOD_TMAP_16, 4, out, dims, strides,
                         box, es1, CU_TENSOR_MAP_INTERLEAVE_NONE, oswz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            ysod_set_error("ysod_conv_tc_create: cuTensorMapEncodeTiled(O) failed with %d (Cout %d ocs %d unit %d)", (int)r, Cout, ocs,
                           p.unit_cols);
            return YSOD_ERR_CUDA;
        }
    }
    YSOD_CHECK_ARG(!res || Cout % 16 == 0, "ysod_conv_tc_create: residual needs Cout %% 16 == 0");
    YSOD_CHECK_ARG(c.smem <= 227 * 1024, "ysod_conv_tc_create: shared memory plan too large (%zu)", c.smem);
    YSOD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    YSOD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ysod_conv_tc* h = new (std::nothrow) ysod_conv_tc;
    YSOD_CHECK_ARG(h, "ysod_conv_tc_create: out of memory");
    h->c = c;
    owned_guard.keep = true;
    *handle = h;
    return YSOD_OK;
}

int ysod_conv_tc_create(ysod_conv_tc** handle, const void* x, int N, int H, int W, int Cin, int xcs, const void* wgt,
                        const float* bias, int Cout, int Cout_pad, int ksize, int stride, void* out, int out_dtype, int ocs,
                        const void* res, int rcs, int act) {
    return ysod_conv_tc_create_ex(handle, x, N, H, W, Cin, xcs, wgt, bias, Cout, Cout_pad, ksize, stride, out, out_dtype, ocs, res, rcs,
                                  act, 0);
}

int ysod_conv_tc_run(ysod_conv_tc* h, cudaStream_t stream) {
    YSOD_CHECK_ARG(h, "ysod_conv_tc_run: null handle");
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = h->c.grid;
    cfg.blockDim = dim3(TC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = h->c.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: prologue overlaps the previous kernel's tail
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (h->c.halo) YSOD_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<true>, h->c.tmA, h->c.tmB, h->c.tmO, h->c.tmA2, h->c.tmO2, h->c.p));
    else YSOD_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<false>, h->c.tmA, h->c.tmB, h->c.tmO, h->c.tmA2, h->c.tmO2, h->c.p));
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

// Fuses the Detect decode into the epilogue of a plan created for a level's final 1x1 head conv (fp32 raw map out, no
// activation, one N tile, channels [0,64) = 4 x 16 DFL bins, [64, 64+nc) = class logits): in addition to the raw map the launch
// writes y (B, 4+nc, A_total) fp32 for the anchors [a_off, a_off + Ho*Wo) of this level. Replaces ysod_dfl_decode for the level.
int ysod_conv_tc_set_decode(ysod_conv_tc* h, float* y, int A_total, int a_off, int nc, float stride) {
    YSOD_CHECK_ARG(h && y, "ysod_conv_tc_set_decode: null");
    TcParams& p = h->c.p;
    YSOD_CHECK_ARG(p.out_f32 && p.act == YSOD_ACT_NONE && p.res == nullptr && p.n_tiles == 1 && !p.up2 && !p.duo,
                   "ysod_conv_tc_set_decode: plan must be an fp32-output, activation-free, single-N-tile conv");
    YSOD_CHECK_ARG(nc > 0 && p.Cout == 64 + nc && p.BN >= 64 + nc, "ysod_conv_tc_set_decode: Cout %d != 64 + nc (%d)", p.Cout, nc);
    YSOD_CHECK_ARG(a_off >= 0 && a_off + p.Ho * p.Wo <= A_total, "ysod_conv_tc_set_decode: anchor range out of bounds");
    YSOD_CHECK_ARG(!p.no_store || !p.stage_split, "ysod_conv_tc_set_decode: YSOD_CONV_NO_STORE needs the single-pass epilogue");
    p.dec_y = y; p.dec_A = A_total; p.dec_off = a_off; p.dec_nc = nc; p.dec_stride = stride;
    return YSOD_OK;
}

// Back-to-back GEMM + Detect decode (TcParams::b2b): the plan must be a 64-output-channel bf16 conv whose staged tile is one 128 B
// swizzled row per pixel (BN = 64, single store unit, no split staging, no fused upsample). w2: [n2][64] bf16 (n2 % 16 == 0, rows
// >= the real channel count zero), bias2: [n2] fp32. kind 1: w2 = Detect cv2[i][2] (64 DFL logits, n2 = 64) -> y[0..4); kind 2: w2 =
// cv3[i][2] (nc class logits, n2 = 16 * ceil(nc / 16)) -> y[4..4+nc). raw (optional): the level's NHWC fp32 raw map, pixel stride
// raw_cs, written at channel offset raw_coff (0 for the box branch, 64 for the class branch). The conv's own output is not stored.
int ysod_conv_tc_set_b2b(ysod_conv_tc* h, const void* w2, const float* bias2, int n2, int kind, float* y, int A_total, int a_off, int nc,
                         float stride, float* raw, int raw_cs, int raw_coff) {
    YSOD_CHECK_ARG(h && w2 && bias2 && y, "ysod_conv_tc_set_b2b: null");
    ConvTc& c = h->c;
    TcParams& p = c.p;
    YSOD_CHECK_ARG(p.BN == 64 && !p.duo && p.n_tiles == 1 && !p.out_f32 && p.n_units == 1 && !p.stage_split && !p.up2 && p.row_bytes == 128 && p.dec_y == nullptr,
                   "ysod_conv_tc_set_b2b: plan must be a 64-channel bf16 conv with a single 128 B staging unit (BN %d, units %d, split %d)", p.BN,
                   p.n_units, p.stage_split);
    YSOD_CHECK_ARG(kind == 1 || kind == 2, "ysod_conv_tc_set_b2b: kind %d", kind);
    YSOD_CHECK_ARG(n2 % 16 == 0 && n2 >= 16 && n2 <= 128 && (kind != 1 || n2 == 64) && (kind != 2 || (nc > 0 && nc <= n2)), "ysod_conv_tc_set_b2b: bad n2 %d / nc %d", n2, nc);
    YSOD_CHECK_ARG(a_off >= 0 && a_off + p.Ho * p.Wo <= A_total, "ysod_conv_tc_set_b2b: anchor range out of bounds");
    YSOD_CHECK_ARG(((uintptr_t)w2 % 16) == 0 && (!raw || (((uintptr_t)raw % 16) == 0 && raw_cs % 4 == 0 && raw_coff % 4 == 0)), "ysod_conv_tc_set_b2b: alignment");
    const size_t extra = 128u * (size_t)n2 + 4u * (size_t)n2 + 16u;
    YSOD_CHECK_ARG(c.smem + extra <= 227 * 1024, "ysod_conv_tc_set_b2b: shared memory plan too large (%zu)", c.smem + extra);
    YSOD_CHECK_ARG(4 * p.BN + 2 * n2 <= 512, "ysod_conv_tc_set_b2b: TMEM columns");
    c.smem += extra;
    p.tmem_cols = 512;
    p.b2b = kind; p.b2_n = n2; p.b2_w = (const __nv_bfloat16*)w2; p.b2_bias = bias2;
    p.b2_raw = raw; p.b2_raw_cs = raw_cs; p.b2_raw_off = raw_coff;
    p.b2_y = y;
    p.dec_A = A_total; p.dec_off = a_off; p.dec_nc = nc; p.dec_stride = stride;
    return YSOD_OK;
}

// Back-to-back GEMM, plain variant: a 1x1 Conv(64 -> 64) + BN + act that consumes ONLY this plan's output (e.g. C2f.cv1 right after
// a Conv layer, block.py:233-248) runs as the second MMA group of the same launch. The plan must have been created with `out` = the
// destination of that second layer (64 channels, bf16): the first layer's own output is never stored. w2: [64][64] bf16 (BN folded).
int ysod_conv_tc_set_b2b_conv(ysod_conv_tc* h, const void* w2, const float* bias2, int act2) {
    YSOD_CHECK_ARG(h && w2 && bias2, "ysod_conv_tc_set_b2b_conv: null");
    ConvTc& c = h->c;
    TcParams& p = c.p;
    YSOD_CHECK_ARG(p.BN == 64 && !p.duo && p.Cout == 64 && p.n_tiles == 1 && !p.out_f32 && p.n_units == 1 && !p.stage_split && !p.up2 && p.row_bytes == 128 &&
                   p.dec_y == nullptr && !p.b2b && !p.no_store,
                   "ysod_conv_tc_set_b2b_conv: plan must be a 64-channel bf16 conv with a single 128 B staging unit (BN %d, units %d, split %d)", p.BN,
                   p.n_units, p.stage_split);
    YSOD_CHECK_ARG(act2 == YSOD_ACT_SILU || act2 == YSOD_ACT_NONE, "ysod_conv_tc_set_b2b_conv: activation %d unsupported", act2);
    YSOD_CHECK_ARG(((uintptr_t)w2 % 16) == 0, "ysod_conv_tc_set_b2b_conv: alignment");
    const size_t extra = 128u * 64u + 4u * 64u + 16u;
    YSOD_CHECK_ARG(c.smem + extra <= 227 * 1024, "ysod_conv_tc_set_b2b_conv: shared memory plan too large (%zu)", c.smem + extra);
    c.smem += extra;
    p.tmem_cols = 512;
    p.b2b = 3; p.b2_n = 64; p.b2_w = (const __nv_bfloat16*)w2; p.b2_bias = bias2; p.b2_act = act2;
    p.b2_raw = nullptr; p.b2_y = nullptr;
    return YSOD_OK;
}

// Back-to-back GEMM over a concatenation: C2f.cv2 (block.py:233-248: `cv2(cat(cv1 halves, bottleneck output))`, a 1x1 Conv(96 -> 64))
// runs inside the launch of the block's last Bottleneck conv (3x3, 32 -> 32, halo plan). Its K = 96 input channels are the staged
// 32-channel tile plus the 64 channels of `x2` (the cv1 output), whose 16 x 8 pixel tile is TMA-loaded per output tile into a two-slot
// ring. w2: [64][96] bf16, columns [0,64) = the x2 channels, [64,96) = the staged channels (= torch.cat order); out2: the 64-channel
// destination (NHWC view, pixel stride out2cs). The Bottleneck conv's own output is never stored.
int ysod_conv_tc_set_b2b_cat(ysod_conv_tc* h, const void* x2, int x2cs, const void* w2, const float* bias2, int act2, void* out2, int out2cs) {
    YSOD_CHECK_ARG(h && x2 && w2 && bias2 && out2, "ysod_conv_tc_set_b2b_cat: null");
    ConvTc& c = h->c;
    TcParams& p = c.p;
    YSOD_CHECK_ARG(c.halo && p.BN == 32 && p.Cout == 32 && p.n_tiles == 1 && !p.out_f32 && p.n_units == 1 && !p.stage_split && !p.up2 && p.row_bytes == 64 &&
                   p.dec_y == nullptr && !p.b2b && !p.no_store && !p.pair,
                   "ysod_conv_tc_set_b2b_cat: plan must be a 3x3 32 -> 32 bf16 conv on the halo kernel (BN %d, units %d, split %d, pair %d)", p.BN,
                   p.n_units, p.stage_split, p.pair);
    YSOD_CHECK_ARG(act2 == YSOD_ACT_SILU || act2 == YSOD_ACT_NONE, "ysod_conv_tc_set_b2b_cat: activation %d unsupported", act2);
    YSOD_CHECK_ARG(((uintptr_t)w2 % 16) == 0 && ((uintptr_t)x2 % 16) == 0 && ((uintptr_t)out2 % 16) == 0 && x2cs % 8 == 0 && out2cs % 8 == 0,
                   "ysod_conv_tc_set_b2b_cat: alignment");
    EncodeTiledFn enc = get_encode();
    YSOD_CHECK_ARG(enc != nullptr, "ysod_conv_tc_set_b2b_cat: cuTensorMapEncodeTiled unavailable");
    for (int which = 0; which < 2; ++which) {
        void* base = which ? out2 : const_cast<void*>(x2);
        const int cs = which ? out2cs : x2cs;
        cuuint64_t dims[4] = {64, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)p.N};
        cuuint64_t strides[3] = {(cuuint64_t)cs * 2, (cuuint64_t)p.Wo * cs * 2, (cuuint64_t)p.Ho * p.Wo * cs * 2};
        cuuint32_t box[4] = {64, 8, 16, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(which ? &c.tmO2 : &c.tmA2, YSOD_TMAP_16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, which ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            ysod_set_error("ysod_conv_tc_set_b2b_cat: cuTensorMapEncodeTiled(%s) failed with %d", which ? "out2" : "x2", (int)r);
            return YSOD_ERR_CUDA;
        }
    }
    // W2 (64 x 128 B + 64 x 64 B) + bias + 6 barriers + alignment + 2 extra-operand tiles + 2 output staging tiles
    const size_t extra = 128u * 64u + 64u * 64u + 4u * 64u + 48u + 1024u + 4u * 16384u;
    // make room: the halo ring gives up slots (two at a time: one per issuer sub-ring; each slot = a_bytes + its barrier pair)
    while (c.smem + extra > 227 * 1024 && p.a_slots >= 2 * p.issuers) {
        p.a_slots -= p.issuers;
        p.a_stages = p.a_slots / p.issuers;
        p.stages = p.a_stages;
        c.smem -= (size_t)p.issuers * (p.a_bytes + 16u);
    }
    YSOD_CHECK_ARG(c.smem + extra <= 227 * 1024, "ysod_conv_tc_set_b2b_cat: shared memory plan too large (%zu)", c.smem + extra);
    c.smem += extra;
    p.tmem_cols = 256;   // 4 x 32 accumulator columns + 2 x 64 for the second layer
    p.b2b = 4; p.b2_n = 64; p.b2_k2 = 64; p.b2_w = (const __nv_bfloat16*)w2; p.b2_bias = bias2; p.b2_act = act2;
    p.b2_raw = nullptr; p.b2_y = nullptr;
    return YSOD_OK;
}

// SE folded into the following conv (smallobj_modules.py:84-92 `x * a` followed by conv.py:37-55): conv(x * a[n]) == conv with
// the weights' input-channel columns scaled by a[n], so the gated activation map is never written / re-read.
// w: [rows][K] fp32 master weights (K ordered (r, s, cin), BN folded), gate: [N][Cin] fp32 -> out: [N][rows][K] bf16.
int ysod_scale_weights(const float* w, int rows, int K, int Cin, const float* gate, int N, void* out, cudaStream_t stream) {
    YSOD_CHECK_ARG(w && gate && out && rows > 0 && K > 0 && Cin > 0 && K % Cin == 0 && N > 0, "ysod_scale_weights: bad args");
    const long long total = (long long)N * rows * K;
    ysod_launch(scale_weights_kernel, ysod_cdiv(total, 256), 256, 0, stream, w, rows, K, Cin, gate, (__nv_bfloat16*)out, total);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

// tiling chosen for a plan: out[0..7] = TH, TW, BN, BK, stages, grid.x, grid.y, smem bytes
int ysod_conv_tc_info(ysod_conv_tc* h, int* out8) {
    YSOD_CHECK_ARG(h && out8, "ysod_conv_tc_info: null");
    out8[0] = h->c.p.TH; out8[1] = h->c.p.TW; out8[2] = h->c.p.BN; out8[3] = h->c.p.BK; out8[4] = h->c.p.stages;
    out8[5] = (int)h->c.grid.x; out8[6] = h->c.halo ? (1000 + 100 * h->c.p.issuers + 10 * h->c.p.b_resident + h->c.p.n_tiles) : (100 * h->c.p.issuers + 10 * h->c.p.kgroup + h->c.p.n_tiles);
    out8[6] += 10000 * h->c.p.pair + 20000 * h->c.p.duo;
    out8[7] = (int)h->c.smem;
    return YSOD_OK;
}

void ysod_conv_tc_destroy(ysod_conv_tc* h) {
    if (h && h->c.owned_w) cudaFree(h->c.owned_w);
    delete h;
}

// Profiling aid: copies the pipeline trace of the last launch with debug bit 32 (mode = 32 << 8) to the host and resets it.
// out: 2 * cap uint64 (tag, clock64); returns the number of records. Synchronises the device.
int ysod_debug_trace(unsigned long long* out, int cap) {
    if (cudaDeviceSynchronize() != cudaSuccess || cap < TRACE_CAP) return -1;
    if (cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * 2 * TRACE_CAP) != cudaSuccess) return -1;
    void* sym = nullptr;
    if (cudaGetSymbolAddress(&sym, g_trace) != cudaSuccess) return -1;
    cudaMemset(sym, 0, sizeof(unsigned long long) * 2 * TRACE_CAP);
    return TRACE_CAP;
}

}  // extern "C"
