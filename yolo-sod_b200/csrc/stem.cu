// Stem convolution (3x3, stride 2, pad 1, Cin = 3 -> Cout in {16, 32, 64}) for the bf16 path, on tensor cores.
//
// Replaces (reference): layer 0 `Conv(3, c, 3, 2)` (ultralytics/nn/modules/conv.py:37-55, BN folded) together with what feeds
// it: either the predictor's float image tensor (B,3,H,W) in [0,1] (engine/predictor.py:116-134, tensor branch), or -- for the
// end-to-end path -- the predictor's raw uint8 BGR HWC frames, in which case `im[..., ::-1].transpose(0,3,1,2) / 255`
// (predictor.py:127-133) is fused into the load.
//
// The op is HBM-bound (K = 27: 1.2-12 B of input and 64 B of output per output pixel) but 864 MACs per pixel make a CUDA-core
// version FMA-bound at ~5x the memory time, so the MACs go to the tensor cores. K = 27 (padded to 32) is far too short for a
// tcgen05/TMEM pipeline to pay off (one M=128,N=32 MMA pair per tile, accumulator round trip through TMEM); warp-level
// mma.sync m16n8k16 with register accumulators is the right tool here: a warp owns 16 output pixels x Cout, gathers its
// im2col A fragments straight from a shared-memory copy of the input patch and keeps the whole weight matrix in registers.
// CTA tile = 4 output rows x 64 output columns; output is staged in shared memory and written as full 64-byte pixel rows.
#include "common.cuh"

namespace {

constexpr int TH = 4, TW = 64;                 // output tile
constexpr int PH = 2 * TH + 1, PW = 2 * TW + 1;  // input patch 9 x 129
constexpr int PWP = 132;                       // padded patch row (bf16 elements)
constexpr int PLANE = PH * PWP;

__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." YSOD_MMA_T "." YSOD_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float silu_tanh(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

// patch offset of im2col column k = (r*3 + s)*3 + c  (k >= 27: padding column, weight is zero; read element 0)
__device__ __forceinline__ int koff(int k) {
    if (k >= 27) return 0;
    const int tap = k / 3, c = k - tap * 3;
    const int r = tap / 3, s = tap - r * 3;
    return c * PLANE + r * PWP + s;
}

// SRC = 0: img is (N,3,H,W) fp32 in [0,1];  SRC = 1: img is (N,H,W,3) uint8 BGR in 0..255 (-> RGB, /255)
template <int COUT, int SRC, bool VEC>
__global__ void __launch_bounds__(256, 2)
stem_mma_kernel(const void* img_, const __nv_bfloat16* __restrict__ wk, const float* __restrict__ bias,
                __nv_bfloat16* __restrict__ out, int H, int W, int Ho, int Wo, int ocs, int act, int indirect, float* __restrict__ psum) {
    ysod_pdl_sync();
    // indirect: img_ is a device slot holding the image pointer (ysod_set_ptr), so a captured graph can read whichever tensor
    // the caller passed to this forward instead of a private staging copy
    if (indirect) img_ = *static_cast<const void* const*>(img_);
    constexpr int NB = COUT / 8;
    __shared__ __align__(16) __nv_bfloat16 patch[3 * PLANE];
    __shared__ __align__(16) __nv_bfloat16 stage[TH * TW * (COUT + 8)];   // +8: keeps the 4-byte fragment stores conflict-free
    // uint8 source: the 256 possible values of bf16(x / 255.0f) (the reference's `im / 255` then the 16-bit cast), built once per CTA
    __shared__ __nv_bfloat16 lut[SRC == 1 ? 256 : 1];
    if (SRC == 1) lut[threadIdx.x] = __float2bfloat16_rn((float)threadIdx.x / 255.0f);   // 256 threads; visible after the first barrier below
    const int n = blockIdx.z;
    const int oh0 = blockIdx.y * TH;
    const int ih0 = 2 * oh0 - 1;
    const int tid = threadIdx.x;
    const int ntx = (Wo + TW - 1) / TW;   // the CTA walks the ntx tiles of its 4-row strip; the next tile's input is in flight while this one is computed

    // ---- weights: B fragments of the [COUT][32] bf16 matrix, held in registers for the whole CTA
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    uint32_t bf[NB][2][2];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const __nv_bfloat16* wp = wk + (size_t)(nb * 8 + g) * 32 + ks * 16 + 2 * t;
            bf[nb][ks][0] = *reinterpret_cast<const uint32_t*>(wp);
            bf[nb][ks][1] = *reinterpret_cast<const uint32_t*>(wp + 8);
        }
    }
    // im2col gather offsets of this thread's 16 A columns (k = ks*16 + {2t, 2t+1, 2t+8, 2t+9})
    int ko[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        ko[ks][0] = koff(ks * 16 + 2 * t);
        ko[ks][1] = koff(ks * 16 + 2 * t + 1);
        ko[ks][2] = koff(ks * 16 + 2 * t + 8);
        ko[ks][3] = koff(ks * 16 + 2 * t + 9);
    }
    float bv[NB][2];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) { bv[nb][0] = __ldg(bias + nb * 8 + 2 * t); bv[nb][1] = __ldg(bias + nb * 8 + 2 * t + 1); }

    // ---- input patch of a tile: the VEC paths hold it in registers between the request (issued one tile ahead) and the conversion to
    //      bf16 in shared memory, exactly like the reference's cast of the image tensor. Patch column j holds image column 2*ow0 - 4 + j
    //      (so rows start 16-byte aligned).
    constexpr int NV0 = 3 * PH * 33, NV1 = PH * 33;             // float4 pieces (fp32 planes) / 12-byte items = 4 BGR pixels (uint8 rows: 132 pixels)
    constexpr int NIT = ((SRC == 0 ? NV0 : NV1) + 255) / 256;
    float4 vf[SRC == 0 ? NIT : 1];
    uint32_t vu[SRC == 1 ? NIT : 1][3];
    auto request = [&](int ow0) {
        if (SRC == 0) {
            const float* img = static_cast<const float*>(img_);
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int i = tid + it * 256;
                vf[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < NV0) {
                    const int cr = i / 33, j = i - cr * 33;
                    const int c = cr / PH, r = cr - c * PH;
                    const int ih = ih0 + r, iw = 2 * ow0 - 4 + 4 * j;
                    if (ih >= 0 && ih < H && iw >= 0 && iw < W) vf[it] = __ldg(reinterpret_cast<const float4*>(img + (((size_t)n * 3 + c) * H + ih) * W + iw));
                }
            }
        } else {
            const uint8_t* img = static_cast<const uint8_t*>(img_);
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int i = tid + it * 256;
                vu[it][0] = vu[it][1] = vu[it][2] = 0u;
                if (i < NV1) {
                    const int r = i / 33, j = i - r * 33;
                    const int ih = ih0 + r;
                    const int b0 = (2 * ow0 - 4) * 3 + 12 * j;         // byte offset inside the image row (a multiple of 4: W % 4 == 0)
                    if (ih >= 0 && ih < H) {
                        const uint32_t* rowp = reinterpret_cast<const uint32_t*>(img + ((size_t)n * H + ih) * W * 3);
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            if (b0 + 4 * k >= 0 && b0 + 4 * k + 3 < W * 3) vu[it][k] = __ldg(rowp + (b0 >> 2) + k);
                    }
                }
            }
        }
    };
    auto commit = [&]() {
        if (SRC == 0) {
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int i = tid + it * 256;
                if (i < NV0) {
                    const int cr = i / 33, j = i - cr * 33;
                    __nv_bfloat162* d = reinterpret_cast<__nv_bfloat162*>(&patch[cr * PWP + 4 * j]);
                    d[0] = __floats2bfloat162_rn(vf[it].x, vf[it].y);
                    d[1] = __floats2bfloat162_rn(vf[it].z, vf[it].w);
                }
            }
        } else {
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int i = tid + it * 256;
                if (i < NV1) {
                    const int r = i / 33, j = i - r * 33;
                    // 12 bytes = B0 G0 R0 B1 G1 R1 B2 G2 R2 B3 G3 R3 -> four consecutive patch columns of each RGB plane (8 B stores)
                    const unsigned short* lt = reinterpret_cast<const unsigned short*>(lut);
#pragma unroll
                    for (int cs = 0; cs < 3; ++cs) {
                        uint32_t e[4];
#pragma unroll
                        for (int px = 0; px < 4; ++px) {
                            const int bi = 3 * px + cs;
                            e[px] = lt[(vu[it][bi >> 2] >> (8 * (bi & 3))) & 0xffu];
                        }
                        *reinterpret_cast<uint2*>(&patch[(2 - cs) * PLANE + r * PWP + 4 * j]) = make_uint2(e[0] | (e[1] << 16), e[2] | (e[3] << 16));
                    }
                }
            }
        }
    };
    if (VEC) request(0);
    if (SRC == 1) __syncthreads();   // the table is complete before the first conversion

    for (int tx = 0; tx < ntx; ++tx) {
        const int ow0 = tx * TW, iw0 = 2 * ow0 - 1;
        if (VEC) {
            commit();
        } else if (SRC == 0) {
            const float* img = static_cast<const float*>(img_);
            for (int i = tid; i < 3 * PH * PW; i += 256) {
                const int c = i / (PH * PW);
                const int rem = i - c * (PH * PW);
                const int r = rem / PW, q = rem - r * PW;
                const int ih = ih0 + r, iw = iw0 + q;
                float v = 0.f;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __ldg(img + (((size_t)n * 3 + c) * H + ih) * W + iw);
                patch[c * PLANE + r * PWP + q + 3] = __float2bfloat16_rn(v);
            }
        } else {
            const uint8_t* img = static_cast<const uint8_t*>(img_);
            for (int i = tid; i < PH * PW * 3; i += 256) {
                const int r = i / (PW * 3);
                const int rem = i - r * (PW * 3);
                const int q = rem / 3, cs = rem - q * 3;
                const int ih = ih0 + r, iw = iw0 + q;
                unsigned int v = 0u;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __ldg(img + (((size_t)n * H + ih) * W + iw) * 3 + cs);
                patch[(2 - cs) * PLANE + r * PWP + q + 3] = lut[v];
            }
        }
        __syncthreads();
        if (VEC && tx + 1 < ntx) request(ow0 + TW);   // travels under this tile's MMAs and stores
        float gs[NB][2];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) gs[nb][0] = gs[nb][1] = 0.f;

        const unsigned short* ps = reinterpret_cast<const unsigned short*>(patch);
        // 16 m-tiles (16 consecutive output columns of one output row each); warp w takes m-tiles 2w, 2w+1
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
            const int mt = warp * 2 + mi;
            const int row = mt >> 2, col0 = (mt & 3) * 16;
            const int base0 = (2 * row) * PWP + 2 * (col0 + g) + 3;    // pixel g of the m-tile (patch column 3 = image column 2*ow0-1)
            const int base1 = base0 + 16;                               // pixel g + 8
            float acc[NB][4];
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) { acc[nb][0] = bv[nb][0]; acc[nb][1] = bv[nb][1]; acc[nb][2] = bv[nb][0]; acc[nb][3] = bv[nb][1]; }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t a[4];
                a[0] = (uint32_t)ps[base0 + ko[ks][0]] | ((uint32_t)ps[base0 + ko[ks][1]] << 16);
                a[1] = (uint32_t)ps[base1 + ko[ks][0]] | ((uint32_t)ps[base1 + ko[ks][1]] << 16);
                a[2] = (uint32_t)ps[base0 + ko[ks][2]] | ((uint32_t)ps[base0 + ko[ks][3]] << 16);
                a[3] = (uint32_t)ps[base1 + ko[ks][2]] | ((uint32_t)ps[base1 + ko[ks][3]] << 16);
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) mma_bf16_16816(acc[nb], a, bf[nb][ks][0], bf[nb][ks][1]);
            }
            // activation -> bf16 -> staging tile [pixel][COUT + 8]
            const int p0 = row * TW + col0 + g;
            const bool ok0 = oh0 + row < Ho && ow0 + col0 + g < Wo, ok1 = oh0 + row < Ho && ow0 + col0 + g + 8 < Wo;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                float v0 = acc[nb][0], v1 = acc[nb][1], v2 = acc[nb][2], v3 = acc[nb][3];
                if (act == YSOD_ACT_SILU) {   // x * sigmoid(x) = h + h * tanh(h), h = x / 2 (same form as the tcgen05 conv epilogue)
                    v0 = silu_tanh(v0); v1 = silu_tanh(v1); v2 = silu_tanh(v2); v3 = silu_tanh(v3);
                }
                const __nv_bfloat162 q0 = __floats2bfloat162_rn(v0, v1), q1 = __floats2bfloat162_rn(v2, v3);
                *reinterpret_cast<__nv_bfloat162*>(&stage[(size_t)p0 * (COUT + 8) + nb * 8 + 2 * t]) = q0;
                *reinterpret_cast<__nv_bfloat162*>(&stage[(size_t)(p0 + 8) * (COUT + 8) + nb * 8 + 2 * t]) = q1;
                if (psum != nullptr) {   // pooling partials of the STORED values (this thread: channels nb*8 + 2t, +1 of two pixels)
                    const float2 f0 = __bfloat1622float2(q0), f1 = __bfloat1622float2(q1);
                    gs[nb][0] += (ok0 ? f0.x : 0.f) + (ok1 ? f1.x : 0.f);
                    gs[nb][1] += (ok0 ? f0.y : 0.f) + (ok1 ? f1.y : 0.f);
                }
            }
        }
        if (psum != nullptr) {
            // sum over the 8 pixel lanes (g) of the warp, then lanes 0..3 (t) hold the warp's sums of channels nb*8 + 2t, +1
            float* part = reinterpret_cast<float*>(patch) ;   // [8 warps][COUT], written after the barrier below (the patch is dead then)
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float a = gs[nb][e];
                    a += __shfl_xor_sync(0xffffffffu, a, 4);
                    a += __shfl_xor_sync(0xffffffffu, a, 8);
                    a += __shfl_xor_sync(0xffffffffu, a, 16);
                    gs[nb][e] = a;
                }
            __syncthreads();   // every warp is done reading the patch
            if (g == 0) {
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) { part[warp * COUT + nb * 8 + 2 * t] = gs[nb][0]; part[warp * COUT + nb * 8 + 2 * t + 1] = gs[nb][1]; }
            }
        }
        __syncthreads();
        // ---- optional: this tile's per-channel sums of the stored (16-bit) values = partial sums of the global average pool that the
        //      SE block after the stem starts with (smallobj_modules.py SE_Block: avg_pool), so the map is not read again for them.
        //      psum[n][tile][COUT], tile = blockIdx.y * ntx + tx; fixed summation order (lanes, then warps 0..7).
        if (psum != nullptr) {
            const float* part = reinterpret_cast<const float*>(patch);
            if (tid < COUT) {
                float t2 = 0.f;
#pragma unroll
                for (int w2 = 0; w2 < 8; ++w2) t2 += part[w2 * COUT + tid];
                psum[((size_t)n * (ntx * gridDim.y) + blockIdx.y * ntx + tx) * COUT + tid] = t2;
            }
            __syncthreads();   // `part` aliases the patch the next tile's conversion overwrites
        }
        // ---- coalesced write-out: 16-byte pieces, consecutive threads -> consecutive channels then pixels
        constexpr int PIECES = COUT / 8;
        for (int i = tid; i < TH * TW * PIECES; i += 256) {
            const int pix = i / PIECES, pc = i - pix * PIECES;
            const int r = pix / TW, q = pix - r * TW;
            const int oh = oh0 + r, ow = ow0 + q;
            if (oh < Ho && ow < Wo) {
                const uint4 v = *reinterpret_cast<const uint4*>(&stage[(size_t)pix * (COUT + 8) + pc * 8]);
                *reinterpret_cast<uint4*>(out + (((size_t)n * Ho + oh) * Wo + ow) * ocs + pc * 8) = v;
            }
        }
        // (the next tile's staging writes come after its own __syncthreads, which every thread reaches only after these reads)
    }
}

__global__ void set_ptr_kernel(const void** slot, const void* value) { *slot = value; }

}  // namespace

// Binds the image a captured forward reads: writes `value` into the device pointer slot the stem kernel dereferences
// (src_fmt | YSOD_STEM_INDIRECT). Stream-ordered, so it takes effect for the launches that follow it on `stream`.
extern "C" int ysod_set_ptr(void* slot, const void* value, cudaStream_t stream) {
    YSOD_CHECK_ARG(slot && ((uintptr_t)slot % 8) == 0, "ysod_set_ptr: slot must be an 8 B aligned device pointer");
    set_ptr_kernel<<<1, 1, 0, stream>>>((const void**)slot, value);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

// img: src_fmt 0 = (N,3,H,W) fp32 in [0,1] (what DetectionModel.forward receives, tasks.py:129);
//      src_fmt 1 = (N,H,W,3) uint8 BGR frames (what BasePredictor.preprocess receives, predictor.py:116-134).
// wk: [Cout][32] bf16, column k = (r*3 + s)*3 + c (RGB channel c), columns 27..31 zero; bias fp32 [Cout] (BN folded).
// out: NHWC bf16 view with pixel stride ocs. 3x3 / stride 2 / pad 1 only; H, W even.
// psum != NULL: the kernel also writes per-tile channel sums of its output, psum[N][S][Cout] fp32 with S = ceil(W/2 / 64) * ceil(H/2 / 4)
// (the layout ysod_se_gate / ysod_gap_partial use), so a following SE block skips its pooling pass.
static int stem_mma_impl(const void* img, int src_fmt, int N, int H, int W, const void* wk, const float* bias, int Cout, void* out,
                         int ocs, int act, float* psum, cudaStream_t stream) {
    YSOD_CHECK_ARG(img && wk && bias && out, "ysod_stem_mma: null pointer");
    YSOD_CHECK_ARG(Cout == 16 || Cout == 32 || Cout == 64, "ysod_stem_mma: Cout %d unsupported (16, 32, 64)", Cout);
    const int indirect = (src_fmt & YSOD_STEM_INDIRECT) ? 1 : 0;
    src_fmt &= ~YSOD_STEM_INDIRECT;
    YSOD_CHECK_ARG(src_fmt == 0 || src_fmt == 1, "ysod_stem_mma: bad source format %d", src_fmt);
    YSOD_CHECK_ARG(H % 2 == 0 && W % 2 == 0 && ocs % 8 == 0 && ((uintptr_t)out % 16) == 0, "ysod_stem_mma: bad geometry / alignment");
    YSOD_CHECK_ARG(act == YSOD_ACT_SILU || act == YSOD_ACT_NONE, "ysod_stem_mma: activation %d unsupported", act);
    const int Ho = H / 2, Wo = W / 2;
    dim3 grid(1, ysod_cdiv(Ho, TH), N);   // a CTA walks the tiles of its 4-row strip
    const __nv_bfloat16* w = (const __nv_bfloat16*)wk;
    __nv_bfloat16* o = (__nv_bfloat16*)out;
    // aligned 16 B / 4 B row loads (an indirect image pointer must be 16 B aligned: the binder checks it)
    const bool vec = (W % 4 == 0) && (indirect || (uintptr_t)img % 16 == 0);
#define LAUNCH(CO, SRC)                                                                                          \
    do {                                                                                                         \
        if (vec) ysod_launch(stem_mma_kernel<CO, SRC, true>, grid, 256, 0, stream, img, w, bias, o, H, W, Ho, Wo, ocs, act, indirect, psum);   \
        else ysod_launch(stem_mma_kernel<CO, SRC, false>, grid, 256, 0, stream, img, w, bias, o, H, W, Ho, Wo, ocs, act, indirect, psum);      \
    } while (0)
    if (Cout == 16) { if (src_fmt) LAUNCH(16, 1); else LAUNCH(16, 0); }
    else if (Cout == 32) { if (src_fmt) LAUNCH(32, 1); else LAUNCH(32, 0); }
    else { if (src_fmt) LAUNCH(64, 1); else LAUNCH(64, 0); }
#undef LAUNCH
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

extern "C" int ysod_stem_mma(const void* img, int src_fmt, int N, int H, int W, const void* wk, const float* bias, int Cout, void* out,
                             int ocs, int act, cudaStream_t stream) {
    return stem_mma_impl(img, src_fmt, N, H, W, wk, bias, Cout, out, ocs, act, nullptr, stream);
}

// ysod_stem_mma + the global-average-pool partial sums of its output (see stem_mma_impl): stem Conv -> SE_Block without a pooling pass.
extern "C" int ysod_stem_mma_gap(const void* img, int src_fmt, int N, int H, int W, const void* wk, const float* bias, int Cout, void* out,
                                 int ocs, int act, float* psum, cudaStream_t stream) {
    YSOD_CHECK_ARG(psum != nullptr, "ysod_stem_mma_gap: null psum");
    return stem_mma_impl(img, src_fmt, N, H, W, wk, bias, Cout, out, ocs, act, psum, stream);
}
