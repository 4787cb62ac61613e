// Fused Detect decode: DFL softmax-expectation + dist2bbox + stride scale + class sigmoid, one pass.
//
// Replaces (reference):
//   ultralytics/nn/modules/head.py:100-131   Detect._inference (view/cat/split/cat)
//   ultralytics/nn/modules/block.py:64-83    DFL  (softmax over 16 bins, expectation via a 16->1 1x1 conv)
//   ultralytics/utils/tal.py:333-357         make_anchors (+0.5 cell offset), dist2bbox (xywh)
// Input : one raw head level, NHWC, channels [0,4*reg_max) = box logits (side-major: l,t,r,b x reg_max bins),
//         [4*reg_max, 4*reg_max+nc) = class logits; pixel stride `cs` elements.
// Output: y (B, 4+nc, A_total) fp32, xywh in pixels + sigmoid(cls), anchors ordered level-major then row-major.
// HBM-bound: reads (4*reg_max+nc) values and writes (4+nc) fp32 per anchor; output writes are coalesced along A.
#include "common.cuh"

namespace {

template <typename T, int REG>
__global__ void dfl_decode_kernel(const T* __restrict__ raw, int H, int W, int cs, int nc, float stride,
                                  float* __restrict__ y, int A_total, int a_off) {
    ysod_pdl_sync();
    const int hw = H * W;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (a >= hw) return;
    const T* p = raw + ((size_t)b * hw + a) * cs;
    float d[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        float v[REG];
#pragma unroll
        for (int k = 0; k < REG; k += 8) ysod_vec8<T>::load(p + s * REG + k, v + k);
        float m = v[0];
#pragma unroll
        for (int k = 1; k < REG; ++k) m = fmaxf(m, v[k]);
        float sum = 0.f, acc = 0.f;
#pragma unroll
        for (int k = 0; k < REG; ++k) {
            const float e = expf(v[k] - m);
            sum += e;
            acc += e * (float)k;
        }
        d[s] = acc / sum;
    }
    const float ax = (float)(a % W) + 0.5f, ay = (float)(a / W) + 0.5f;
    const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
    float* o = y + (size_t)b * (4 + nc) * A_total + a_off + a;
    o[0] = (x1 + x2) * 0.5f * stride;
    o[(size_t)A_total] = (y1 + y2) * 0.5f * stride;
    o[(size_t)2 * A_total] = (x2 - x1) * stride;
    o[(size_t)3 * A_total] = (y2 - y1) * stride;
    for (int c = 0; c < nc; ++c) o[(size_t)(4 + c) * A_total] = ysod_sigmoid(ysod_ld<T>(p + 4 * REG + c));
}

}  // namespace

extern "C" int ysod_dfl_decode(const void* raw, int dtype, int B, int H, int W, int cs, int nc, int reg_max,
                               float stride, float* y, int A_total, int a_off, cudaStream_t stream) {
    YSOD_CHECK_ARG(raw && y, "ysod_dfl_decode: null pointer");
    YSOD_CHECK_ARG(reg_max == 16, "ysod_dfl_decode: reg_max must be 16 (head.py:37)");
    YSOD_CHECK_ARG(cs % 8 == 0 && cs >= 4 * reg_max + nc, "ysod_dfl_decode: pixel stride %d must be a multiple of 8 and >= %d", cs, 4 * reg_max + nc);
    YSOD_CHECK_ARG(a_off >= 0 && a_off + H * W <= A_total, "ysod_dfl_decode: anchor range out of bounds");
    dim3 grid(ysod_cdiv(H * W, 128), B);
    if (dtype == YSOD_F32)
        ysod_launch(dfl_decode_kernel<float, 16>, grid, 128, 0, stream, (const float*)raw, H, W, cs, nc, stride, y, A_total, a_off);
    else if (dtype == YSOD_BF16)
        ysod_launch(dfl_decode_kernel<__nv_bfloat16, 16>, grid, 128, 0, stream, (const __nv_bfloat16*)raw, H, W, cs, nc, stride, y, A_total, a_off);
    else {
        ysod_set_error("ysod_dfl_decode: bad dtype %d", dtype);
        return YSOD_ERR_INVALID;
    }
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}
