/* TEST INFRASTRUCTURE ONLY -- plain C restatement of torchvision.ops.nms (CPU kernel).
 *
 * Third-party algorithm (torchvision, pinned ==0.20.1 by the reference's requirements.txt:62;
 * call site ultralytics/utils/ops.py:296). Published algorithm restated from
 * torchvision/csrc/ops/cpu/nms_kernel.cpp::nms_kernel_impl:
 *   - order = stable sort of scores, descending (ties: ascending original index)
 *   - greedy sweep; box j is suppressed by kept box i when
 *         inter / (area_i + area_j - inter) > iou_threshold        (fp32 IoU, double threshold)
 *     with inter = max(0, xx2-xx1) * max(0, yy2-yy1).
 * Built with -ffp-contract=off so no step is fused into an FMA (x86-64 baseline has none either).
 * Pinned against the live torchvision.ops.nms in tests/test_oracle_nms.py.
 * Never linked into the product library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float s; int64_t i; } ysod_ref_key;

static int cmp_desc_stable(const void* a, const void* b) {
    const ysod_ref_key* x = (const ysod_ref_key*)a;
    const ysod_ref_key* y = (const ysod_ref_key*)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i < y->i) ? -1 : (x->i > y->i);
}

/* boxes: n x 4 (x1,y1,x2,y2) fp32; scores: n fp32; keep: out, capacity n; returns #kept.
 * limit > 0: stop after `limit` keeps (prefix of the full answer). */
int64_t ysod_ref_nms(const float* boxes, const float* scores, int64_t n, double iou_threshold,
                     int64_t limit, int64_t* keep) {
    if (n <= 0) return 0;
    ysod_ref_key* order = (ysod_ref_key*)malloc((size_t)n * sizeof(ysod_ref_key));
    float* areas = (float*)malloc((size_t)n * sizeof(float));
    uint8_t* suppressed = (uint8_t*)calloc((size_t)n, 1);
    for (int64_t i = 0; i < n; ++i) {
        order[i].s = scores[i];
        order[i].i = i;
        float w = boxes[4 * i + 2] - boxes[4 * i + 0];
        float h = boxes[4 * i + 3] - boxes[4 * i + 1];
        areas[i] = w * h;
    }
    qsort(order, (size_t)n, sizeof(ysod_ref_key), cmp_desc_stable);
    int64_t nk = 0;
    for (int64_t _i = 0; _i < n; ++_i) {
        int64_t i = order[_i].i;
        if (suppressed[i]) continue;
        keep[nk++] = i;
        if (limit > 0 && nk >= limit) break;
        float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        float iarea = areas[i];
        for (int64_t _j = _i + 1; _j < n; ++_j) {
            int64_t j = order[_j].i;
            if (suppressed[j]) continue;
            float xx1 = ix1 > boxes[4 * j] ? ix1 : boxes[4 * j];
            float yy1 = iy1 > boxes[4 * j + 1] ? iy1 : boxes[4 * j + 1];
            float xx2 = ix2 < boxes[4 * j + 2] ? ix2 : boxes[4 * j + 2];
            float yy2 = iy2 < boxes[4 * j + 3] ? iy2 : boxes[4 * j + 3];
            float w = xx2 - xx1; if (!(w > 0.0f)) w = 0.0f;   /* std::max(0, w): NaN -> 0 */
            float h = yy2 - yy1; if (!(h > 0.0f)) h = 0.0f;
            float inter = w * h;
            float uni = iarea + areas[j];
            uni = uni - inter;
            float ovr = inter / uni;
            if ((double)ovr > iou_threshold) suppressed[j] = 1;
        }
    }
    free(order); free(areas); free(suppressed);
    return nk;
}
