// Predictor glue either side of the forward + NMS hot path, on the device.
//
// Replaces (reference):
//   ultralytics/engine/predictor.py:145-164  pre_transform  -> LetterBox(imgsz, auto, stride) per frame. LetterBox itself lives in
//       ultralytics/data/augment.py, which the reference checkout does not contain (.gitignore:11); the algorithm is upstream
//       ultralytics 8.3.63's: cv2.resize(INTER_LINEAR) to round(shape * r), then cv2.copyMakeBorder(..., value=114).
//   ultralytics/utils/ops.py:92-127 scale_boxes + :319-338 clip_boxes, applied per image in
//       ultralytics/models/yolo/detect/predict.py:38-40 after NMS.
// BGR->RGB, HWC->CHW and /255 (predictor.py:127-133) stay fused in the stem kernel (stem.cu, src_fmt 1), which consumes the
// letterboxed uint8 frames this kernel writes.
//
// The resize reproduces OpenCV's 8-bit INTER_LINEAR bit for bit (imgproc/resize.cpp): coordinates fx = (float)((dx + 0.5) *
// scale - 0.5) with scale computed in double, 11-bit fixed-point coefficients via round-half-even, horizontal pass in int32,
// vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2, and the exact-2x shortcut to the 2x2 box filter.
// HBM-bound (reads <= 4 source pixels, writes one per thread); all index math is per-thread, there are no tables or workspace.
#include "common.cuh"

namespace {

struct LinCoef { int i0, i1, c0, c1; };

// horizontal rule: out-of-range source columns collapse onto the border column with fx = 0
__device__ __forceinline__ LinCoef coef_x(int d, int src, double scale) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    LinCoef c;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src - 1) { c.i0 = c.i1 = src - 1; c.c0 = 2048; c.c1 = 0; return c; }
    c.i0 = s; c.i1 = s + 1;
    c.c0 = __float2int_rn((1.0f - f) * 2048.0f);
    c.c1 = __float2int_rn(f * 2048.0f);
    return c;
}
// vertical rule: coefficients are kept, the two row indices are clamped
__device__ __forceinline__ LinCoef coef_y(int d, int src, double scale) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    const int s = (int)floorf(f);
    f -= (float)s;
    LinCoef c;
    c.i0 = min(max(s, 0), src - 1);
    c.i1 = min(max(s + 1, 0), src - 1);
    c.c0 = __float2int_rn((1.0f - f) * 2048.0f);
    c.c1 = __float2int_rn(f * 2048.0f);
    return c;
}

// mode 0: copy, 1: bilinear (cv2.INTER_LINEAR), 2: exact 2x downscale (2x2 box, what cv::resize substitutes for INTER_LINEAR)
__global__ void letterbox_u8_kernel(const uint8_t* __restrict__ src, int H0, int W0, uint8_t* __restrict__ dst, int H, int W, int new_h,
                                    int new_w, int top, int left, int mode, int value, double scale_x, double scale_y) {
    ysod_pdl_sync();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int b = blockIdx.z;
    if (x >= W) return;
    uint8_t* o = dst + (((size_t)b * H + y) * W + x) * 3;
    const int dx = x - left, dy = y - top;
    if (dx < 0 || dx >= new_w || dy < 0 || dy >= new_h) {
        o[0] = o[1] = o[2] = (uint8_t)value;
        return;
    }
    const uint8_t* s = src + (size_t)b * H0 * W0 * 3;
    if (mode == 0) {
        const uint8_t* p = s + ((size_t)dy * W0 + dx) * 3;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    } else if (mode == 2) {
        const uint8_t* p0 = s + ((size_t)(2 * dy) * W0 + 2 * dx) * 3;
        const uint8_t* p1 = p0 + (size_t)W0 * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = (uint8_t)(((int)p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2);
    } else {
        const LinCoef cx = coef_x(dx, W0, scale_x), cy = coef_y(dy, H0, scale_y);
        const uint8_t* r0 = s + (size_t)cy.i0 * W0 * 3;
        const uint8_t* r1 = s + (size_t)cy.i1 * W0 * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int s0 = (int)r0[cx.i0 * 3 + c] * cx.c0 + (int)r0[cx.i1 * 3 + c] * cx.c1;
            const int s1 = (int)r1[cx.i0 * 3 + c] * cx.c0 + (int)r1[cx.i1 * 3 + c] * cx.c1;
            const int v = (((cy.c0 * (s0 >> 4)) >> 16) + ((cy.c1 * (s1 >> 4)) >> 16) + 2) >> 2;
            o[c] = (uint8_t)min(max(v, 0), 255);
        }
    }
}

// det rows [x1,y1,x2,y2,...] of row_stride floats; params per image: gain, pad_x, pad_y, w0, h0 (fp32)
__global__ void scale_boxes_kernel(float* __restrict__ det, int rows_per_img, int row_stride, const float* __restrict__ params, int params_stride,
                                   long long total) {
    ysod_pdl_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int b = (int)(i / rows_per_img);
    const float* q = params + (size_t)b * params_stride;
    const float gain = q[0], px = q[1], py = q[2], w0 = q[3], h0 = q[4];
    float* r = det + i * row_stride;
    // ops.py:118-126: subtract the pad, true fp32 division by the gain, clamp to the original image
    r[0] = fminf(fmaxf(__fdiv_rn(__fsub_rn(r[0], px), gain), 0.f), w0);
    r[1] = fminf(fmaxf(__fdiv_rn(__fsub_rn(r[1], py), gain), 0.f), h0);
    r[2] = fminf(fmaxf(__fdiv_rn(__fsub_rn(r[2], px), gain), 0.f), w0);
    r[3] = fminf(fmaxf(__fdiv_rn(__fsub_rn(r[3], py), gain), 0.f), h0);
}

}  // namespace

extern "C" {

// frames: (B, H0, W0, 3) uint8 (BGR, as cv2 / the reference's loaders deliver them), contiguous, on the device.
// out:    (B, H, W, 3) uint8 letterboxed frames: the H0 x W0 image resized to new_h x new_w (cv2.INTER_LINEAR semantics) at
//         offset (top, left), everything else filled with `value` (114). Geometry comes from the host (LetterBox rules).
int ysod_letterbox_u8(const void* frames, int B, int H0, int W0, void* out, int H, int W, int new_h, int new_w, int top, int left,
                      int value, cudaStream_t stream) {
    YSOD_CHECK_ARG(frames && out, "ysod_letterbox_u8: null pointer");
    YSOD_CHECK_ARG(B > 0 && B <= 65535 && H0 > 0 && W0 > 0 && H > 0 && H <= 65535 && W > 0 && new_h > 0 && new_w > 0, "ysod_letterbox_u8: bad sizes");
    YSOD_CHECK_ARG(top >= 0 && left >= 0 && top + new_h <= H && left + new_w <= W, "ysod_letterbox_u8: resized image does not fit the output");
    int mode = 1;
    if (new_h == H0 && new_w == W0) mode = 0;
    else if (H0 == 2 * new_h && W0 == 2 * new_w) mode = 2;
    const double sx = 1.0 / ((double)new_w / (double)W0), sy = 1.0 / ((double)new_h / (double)H0);
    dim3 grid(ysod_cdiv(W, 128), H, B);
    ysod_launch(letterbox_u8_kernel, grid, 128, 0, stream, (const uint8_t*)frames, H0, W0, (uint8_t*)out, H, W, new_h, new_w, top, left,
                mode, value, sx, sy);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

// det: (B, rows_per_img, row_stride >= 4) fp32 rows [x1,y1,x2,y2,...] (the padded NMS output has row_stride 6), scaled in place.
// params: per image `params_stride` floats, the first five = gain, pad_x, pad_y, orig_w, orig_h (ops.py:111-116).
int ysod_scale_boxes(float* det, int B, int rows_per_img, int row_stride, const float* params, int params_stride, cudaStream_t stream) {
    YSOD_CHECK_ARG(det && params && B > 0 && rows_per_img > 0 && row_stride >= 4 && params_stride >= 5, "ysod_scale_boxes: bad args");
    const long long total = (long long)B * rows_per_img;
    ysod_launch(scale_boxes_kernel, ysod_cdiv(total, 256), 256, 0, stream, det, rows_per_img, row_stride, params, params_stride, total);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

}  // extern "C"
