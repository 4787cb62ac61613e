// Batched, host-sync-free NMS for the detection hot path.
//
// Replaces (reference, relative to the checkout):
//   ultralytics/utils/ops.py:167-316  non_max_suppression  (per-image Python loop + host syncs)
//   torchvision.ops.nms (ops.py:296)  greedy IoU suppression
//
// Pipeline, all on one stream, no host round trip:
//   K1 nms_score   grid-wide : best class + confidence gate (+ class filter) per anchor (ops.py:234,274-281)
//   K2 nms_sort    CTA/image : order-preserving compaction, then a stable LSD radix sort on the score
//                              (== stable descending sort: equal scores keep ascending anchor index,
//                              which is what torchvision's `scores.sort(stable=True, descending=True)` yields),
//                              truncated to max_nms (ops.py:285-286)
//   K3 nms_gather  grid-wide : xywh->xyxy (ops.py:416-433) + fp32 class offset cls*max_wh (ops.py:289,295)
//   K4 nms_greedy  CTA/image : lazy greedy sweep. Only kept boxes ever suppress anything and the caller keeps at
//                              most max_det of them (ops.py:297), so instead of the full n^2 bitmask only
//                              <= max_det mask rows are evaluated: the CTA finds the next unsuppressed box in a
//                              shared-memory bitset, evaluates its IoU row (one warp per 32-box word, ballot ->
//                              one OR into the bitset) and repeats. Output prefix is identical to the full sweep.
//
// Bit-exactness contract: every IoU step is an explicitly rounded fp32 op (__fsub_rn/__fmul_rn/__fadd_rn/
// __fdiv_rn: no FMA contraction, IEEE division), max/min are written as the comparisons std::max/std::min
// perform, and the `float IoU > double thr` comparison is done against the largest float <= thr (computed on
// the host), which is equivalent for every float IoU.
#include "common.cuh"

namespace {

constexpr int SORT_THREADS = 1024;
constexpr int GREEDY_THREADS = 1024;

__global__ void nms_score_kernel(const float* __restrict__ pred, int nc, int A, float conf_thres,
                                 const int* __restrict__ classes, int n_classes, int multi_label,
                                 float* __restrict__ conf, int* __restrict__ cls) {
    ysod_pdl_sync();
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (a >= A) return;
    const float* p = pred + ((size_t)b * (4 + nc) + 4) * A + a;
    if (!multi_label) {
        float best = p[0];
        int bi = 0;
        for (int c = 1; c < nc; ++c) {
            float v = p[(size_t)c * A];
            if (v > best) { best = v; bi = c; }  // first maximal index, as torch.max(dim)
        }
        bool pass = best > conf_thres;
        if (pass && n_classes > 0) {
            bool hit = false;
            for (int k = 0; k < n_classes; ++k) hit |= (classes[k] == bi);
            pass = hit;
        }
        conf[(size_t)b * A + a] = pass ? best : 0.0f;
        cls[(size_t)b * A + a] = bi;
    } else {
        // validator path (ops.py:270-272): one candidate per (anchor, class) pair, anchor-major order
        for (int c = 0; c < nc; ++c) {
            float v = p[(size_t)c * A];
            bool pass = v > conf_thres;
            if (pass && n_classes > 0) {
                bool hit = false;
                for (int k = 0; k < n_classes; ++k) hit |= (classes[k] == c);
                pass = hit;
            }
            conf[((size_t)b * A + a) * nc + c] = pass ? v : 0.0f;
        }
    }
}

// order-preserving float <-> uint32 map (negative floats reversed, sign bit flipped)
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// One CTA per image. conf: M entries (0 = not a candidate, >0 = score) unless all_candidates.
__global__ void __launch_bounds__(SORT_THREADS, 1)
nms_sort_kernel(const float* __restrict__ conf_all, int M, int cap, int all_candidates, uint32_t* __restrict__ ws_keys0,
                uint32_t* __restrict__ ws_vals0, uint32_t* __restrict__ ws_keys1, uint32_t* __restrict__ ws_vals1,
                int* __restrict__ order_all, float* __restrict__ sscore_all, int* __restrict__ count_all) {
    ysod_pdl_sync();
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* conf = conf_all + (size_t)b * M;
    uint32_t* kin = ws_keys0 + (size_t)b * M;
    uint32_t* vin = ws_vals0 + (size_t)b * M;
    uint32_t* kout = ws_keys1 + (size_t)b * M;
    uint32_t* vout = ws_vals1 + (size_t)b * M;

    __shared__ int s_warp_cnt[32];
    __shared__ int s_warp_off[32];
    __shared__ int s_chunk_total;
    __shared__ int s_hist[256];
    __shared__ int s_bin_base[256];
    __shared__ int s_wh[32 * 256];

    // ---- phase 1: order-preserving compaction -------------------------------------------------------
    // A thread owns EPT consecutive entries per round (all loads issued together), so a round covers 8192 entries behind ONE
    // warp scan + block scan: 5 rounds of two barriers for A = 34 000 instead of 34 rounds whose barrier each waited out a global
    // load. Positions are entry-order exclusive prefix counts, exactly as before.
    constexpr int EPT = 8;
    int n = 0;
    for (int start = 0; start < M; start += SORT_THREADS * EPT) {
        const int i0 = start + tid * EPT;
        float sv[EPT];
        if (i0 + EPT <= M && ((reinterpret_cast<uintptr_t>(conf + i0) & 15) == 0)) {
            const float4 a0 = *reinterpret_cast<const float4*>(conf + i0), a1 = *reinterpret_cast<const float4*>(conf + i0 + 4);
            sv[0] = a0.x; sv[1] = a0.y; sv[2] = a0.z; sv[3] = a0.w; sv[4] = a1.x; sv[5] = a1.y; sv[6] = a1.z; sv[7] = a1.w;
        } else {
#pragma unroll
            for (int e = 0; e < EPT; ++e) sv[e] = (i0 + e < M) ? conf[i0 + e] : 0.0f;
        }
        unsigned flags = 0;
#pragma unroll
        for (int e = 0; e < EPT; ++e)
            if ((i0 + e < M) && (all_candidates || sv[e] > 0.0f)) flags |= 1u << e;
        const int cnt = __popc(flags);
        int inc = cnt;   // inclusive scan of the per-thread counts inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp_cnt[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int v = s_warp_cnt[lane];
            int winc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            s_warp_off[lane] = winc - v;
            if (lane == 31) s_chunk_total = winc;
        }
        __syncthreads();
        int pos = n + s_warp_off[warp] + inc - cnt;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            if (flags & (1u << e)) {
                kin[pos] = ~float_to_ordered(sv[e]);  // ascending key == descending score (NaN first, like torch.sort)
                vin[pos] = (uint32_t)(i0 + e);
                ++pos;
            }
        }
        n += s_chunk_total;
        __syncthreads();   // s_warp_cnt / s_chunk_total are rewritten by the next round
    }
    __syncthreads();

    // ---- phase 2: stable LSD radix sort, 8-bit digits ------------------------------------------------
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = pass * 8;
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += SORT_THREADS) atomicAdd(&s_hist[(kin[i] >> shift) & 255], 1);
        __syncthreads();
        const int single = __syncthreads_or(tid < 256 && n > 0 && s_hist[tid] == n);
        if (single || n <= 1) continue;  // every key shares this digit: the pass is the identity
        if (warp == 0) {
            int local[8];
            int sum = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { local[k] = sum; sum += s_hist[lane * 8 + k]; }
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const int base = inc - sum;
#pragma unroll
            for (int k = 0; k < 8; ++k) s_bin_base[lane * 8 + k] = base + local[k];
        }
        __syncthreads();
        for (int start = 0; start < n; start += SORT_THREADS) {
            for (int j = tid; j < 32 * 256; j += SORT_THREADS) s_wh[j] = 0;
            __syncthreads();
            const int i = start + tid;
            const bool valid = i < n;
            const uint32_t key = valid ? kin[i] : 0u;
            const uint32_t val = valid ? vin[i] : 0u;
            const uint32_t d = (key >> shift) & 255u;
            const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : (0x100u | (uint32_t)lane));
            const int rank = __popc(peers & ((1u << lane) - 1u));
            if (valid && rank == 0) s_wh[warp * 256 + d] = __popc(peers);
            __syncthreads();
            if (tid < 256) {
                int run = s_bin_base[tid];
#pragma unroll 8
                for (int w = 0; w < 32; ++w) {
                    const int t = s_wh[w * 256 + tid];
                    s_wh[w * 256 + tid] = run;
                    run += t;
                }
                s_bin_base[tid] = run;
            }
            __syncthreads();
            if (valid) {
                const int pos = s_wh[warp * 256 + d] + rank;
                kout[pos] = key;
                vout[pos] = val;
            }
            __syncthreads();
        }
        uint32_t* t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
        __syncthreads();
    }

    const int m = n < cap ? n : cap;  // ops.py:285-286: keep the max_nms highest scores
    int* order = order_all + (size_t)b * cap;
    float* sscore = sscore_all + (size_t)b * cap;
    for (int i = tid; i < m; i += SORT_THREADS) {
        order[i] = (int)vin[i];
        sscore[i] = ordered_to_float(~kin[i]);
    }
    if (tid == 0) count_all[b] = m;
}

__device__ __forceinline__ float4 load_xyxy(const float* __restrict__ pred, int b, int nc, int A, int a) {
    const float* p = pred + (size_t)b * (4 + nc) * A + a;
    const float cx = p[0], cy = p[(size_t)A], w = p[(size_t)2 * A], h = p[(size_t)3 * A];
    const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);  // ops.py:429 wh = x[..., 2:] / 2
    return make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}

__global__ void nms_gather_kernel(const float* __restrict__ pred, int nc, int A, int cap, int multi_label,
                                  const int* __restrict__ order_all, const int* __restrict__ cls_all,
                                  const int* __restrict__ count_all, float class_mult, float4* __restrict__ boxes_all) {
    ysod_pdl_sync();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (r >= count_all[b]) return;
    const int id = order_all[(size_t)b * cap + r];
    const int a = multi_label ? id / nc : id;
    const int c = multi_label ? id % nc : cls_all[(size_t)b * A + a];
    const float4 q = load_xyxy(pred, b, nc, A, a);
    const float off = __fmul_rn((float)c, class_mult);  // ops.py:289 c = x[:, 5:6] * (0 if agnostic else max_wh)
    boxes_all[(size_t)b * cap + r] =
        make_float4(__fadd_rn(q.x, off), __fadd_rn(q.y, off), __fadd_rn(q.z, off), __fadd_rn(q.w, off));
}

// std::max(a, b) == (a < b) ? b : a ; std::min(a, b) == (b < a) ? b : a   (NaN behaviour included)
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }

__global__ void __launch_bounds__(GREEDY_THREADS, 1)
nms_greedy_kernel(const float* __restrict__ pred, int nc, int A, int cap, int multi_label, int max_det,
                  float thr_f, const float4* __restrict__ boxes_all, const int* __restrict__ order_all,
                  const float* __restrict__ sscore_all, const int* __restrict__ cls_all,
                  const int* __restrict__ count_all, float* __restrict__ det_all, int* __restrict__ index_all,
                  int* __restrict__ nkeep_all, int box_cache) {
    ysod_pdl_sync();
    extern __shared__ uint32_t smem_u32[];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = count_all[b];
    const int nwords = (n + 31) >> 5;
    uint32_t* removed = smem_u32;                       // cap/32 words
    int* keep_idx = (int*)(smem_u32 + ((cap + 31) >> 5));  // max_det ints
    // the first `ncache` (score-sorted) boxes live in shared memory: every sweep iteration re-reads them, and an L2 round
    // trip per iteration (~700 cycles) used to dominate the 300-iteration serial chain
    float4* sbox = reinterpret_cast<float4*>(smem_u32 + ((((cap + 31) >> 5) + max_det + 3) & ~3));
    __shared__ int s_next;
    const float4* boxes = boxes_all + (size_t)b * cap;
    const int nc_box = n < box_cache ? n : box_cache;

    for (int w = tid; w < nwords; w += GREEDY_THREADS) removed[w] = 0u;
    for (int j = tid; j < nc_box; j += GREEDY_THREADS) sbox[j] = boxes[j];
    __syncthreads();

    int kept = 0, pos = 0;
    while (kept < max_det && pos < n) {
        if (warp == 0) {
            int found = -1;
            for (int w0 = pos >> 5; w0 < nwords && found < 0; w0 += 32) {
                const int w = w0 + lane;
                uint32_t bits = (w < nwords) ? ~removed[w] : 0u;
                if (w == (pos >> 5)) bits &= 0xffffffffu << (pos & 31);
                const int valid = n - w * 32;
                if (valid < 32) bits &= (valid <= 0) ? 0u : ((1u << valid) - 1u);
                const unsigned bal = __ballot_sync(0xffffffffu, bits != 0u);
                if (bal) {
                    const int src = __ffs(bal) - 1;
                    const uint32_t bb = __shfl_sync(0xffffffffu, bits, src);
                    found = (w0 + src) * 32 + __ffs(bb) - 1;
                }
            }
            if (lane == 0) s_next = found;
        }
        __syncthreads();
        const int i = s_next;
        if (i < 0) break;
        if (tid == 0) keep_idx[kept] = i;
        ++kept;
        pos = i + 1;
        if (kept < max_det) {
            const float4 bi = (i < nc_box) ? sbox[i] : boxes[i];
            const float iarea = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
            // thr * (1 +- 2^-20); thresholds at / near zero take the IEEE division for every pair (a quotient that underflows to 0 is
            // not > 0, which the product form cannot see): +-inf bounds make both shortcut comparisons false
            const bool band = thr_f > 1.0e-6f;
            const float thr_hi = band ? __fmul_rn(thr_f, 1.0000009537f) : INFINITY, thr_lo = band ? __fmul_rn(thr_f, 0.9999990463f) : -INFINITY;
            for (int w = (pos >> 5) + warp; w < nwords; w += GREEDY_THREADS / 32) {
                const int j = w * 32 + lane;
                const uint32_t already = removed[w];   // boxes suppressed earlier need no evaluation (their bit stays set)
                bool sup = false;
                if (j > i && j < n && !((already >> lane) & 1u)) {
                    const float4 bj = (j < nc_box) ? sbox[j] : boxes[j];
                    const float xx1 = std_max(bi.x, bj.x), yy1 = std_max(bi.y, bj.y);
                    const float xx2 = std_min(bi.z, bj.z), yy2 = std_min(bi.w, bj.w);
                    const float ww = std_max(0.0f, __fsub_rn(xx2, xx1));
                    const float hh = std_max(0.0f, __fsub_rn(yy2, yy1));
                    const float inter = __fmul_rn(ww, hh);
                    if (inter > 0.0f || !(inter == 0.0f)) {   // disjoint boxes: inter == +0 -> IoU is 0 (or NaN when the union is 0): never > thr
                        const float jarea = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
                        const float uni = __fsub_rn(__fadd_rn(iarea, jarea), inter);
                        // Division-free classification with a 2^-20 guard band around the threshold (the rounded quotient of
                        // torchvision's `inter / union > thr` can only differ from the exact one by 2^-24 relative); the IEEE
                        // division is evaluated only inside the band or for non-positive / non-finite unions.
                        if (uni > 1.0e-30f && uni < 3.0e38f && inter > __fmul_rn(thr_hi, uni)) sup = true;
                        else if (uni > 1.0e-30f && uni < 3.0e38f && inter < __fmul_rn(thr_lo, uni)) sup = false;
                        else sup = __fdiv_rn(inter, uni) > thr_f;
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, sup);
                if (lane == 0 && bal) removed[w] = already | bal;  // this warp is the only writer of word w this round
            }
        }
        __syncthreads();
    }
    __syncthreads();

    // emit detections [x1,y1,x2,y2,conf,cls] (un-offset boxes) + the candidate index of each keep
    float* det = det_all + (size_t)b * max_det * 6;
    int* index = index_all + (size_t)b * max_det;
    for (int t = tid; t < max_det; t += GREEDY_THREADS) {
        if (t < kept) {
            const int r = keep_idx[t];
            const int id = order_all[(size_t)b * cap + r];
            index[t] = id;
            if (pred == nullptr) continue;  // plain box-NMS mode: keep indices only
            const int a = multi_label ? id / nc : id;
            const int c = multi_label ? id % nc : cls_all[(size_t)b * A + a];
            const float4 q = load_xyxy(pred, b, nc, A, a);
            det[t * 6 + 0] = q.x; det[t * 6 + 1] = q.y; det[t * 6 + 2] = q.z; det[t * 6 + 3] = q.w;
            det[t * 6 + 4] = sscore_all[(size_t)b * cap + r];
            det[t * 6 + 5] = (float)c;
        } else {
            index[t] = -1;
            if (pred == nullptr) continue;
#pragma unroll
            for (int k = 0; k < 6; ++k) det[t * 6 + k] = 0.0f;
        }
    }
    if (tid == 0) nkeep_all[b] = kept;
}

// Exact pairwise test of torchvision's nms_kernel_impl: suppress j when inter / (area_i + area_j - inter) > thr, every step a
// separately rounded fp32 operation. Most pairs are classified without the IEEE division (2^-20 guard band, see above).
__device__ __forceinline__ bool iou_exceeds(const float4 bi, const float iarea, const float4 bj, const float thr_f, const float thr_hi,
                                            const float thr_lo) {
    const float xx1 = std_max(bi.x, bj.x), yy1 = std_max(bi.y, bj.y);
    const float xx2 = std_min(bi.z, bj.z), yy2 = std_min(bi.w, bj.w);
    const float ww = std_max(0.0f, __fsub_rn(xx2, xx1));
    const float hh = std_max(0.0f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(ww, hh);
    if (inter == 0.0f) return false;   // IoU is +-0 or NaN: never > thr (thr >= 0)
    const float jarea = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
    const float uni = __fsub_rn(__fadd_rn(iarea, jarea), inter);
    if (uni > 1.0e-30f && uni < 3.0e38f) {
        if (inter > __fmul_rn(thr_hi, uni)) return true;
        if (inter < __fmul_rn(thr_lo, uni)) return false;
    }
    return __fdiv_rn(inter, uni) > thr_f;
}

// Greedy NMS in "kept-list" form for small max_det (the detection path keeps <= 300 boxes, ops.py:297): candidates are visited
// in score order, 32 at a time. A chunk is (a) tested against every box kept so far (1024 threads = 32 candidates x 32 slices of
// the kept list), (b) tested pairwise inside the chunk (one pair per thread), (c) resolved serially by one thread with bit
// operations. A candidate is dropped iff an earlier KEPT box overlaps it -- the same pairwise tests as the sequential sweep, so
// the result is identical; the work is (#candidates visited until max_det are kept) x (#kept) instead of max_det x n.
__global__ void __launch_bounds__(GREEDY_THREADS, 1)
nms_chunk_kernel(const float* __restrict__ pred, int nc, int A, int cap, int multi_label, int max_det, float thr_f,
                 const float4* __restrict__ boxes_all, const int* __restrict__ order_all, const float* __restrict__ sscore_all,
                 const int* __restrict__ cls_all, const int* __restrict__ count_all, float* __restrict__ det_all,
                 int* __restrict__ index_all, int* __restrict__ nkeep_all) {
    ysod_pdl_sync();
    extern __shared__ uint32_t smem_u32[];
    float4* kbox = reinterpret_cast<float4*>(smem_u32);                    // max_det kept boxes
    float* karea = reinterpret_cast<float*>(kbox + max_det);               // their areas
    int* keep_idx = reinterpret_cast<int*>(karea + max_det);               // their sorted positions
    __shared__ uint32_t s_sup, s_row[32];
    __shared__ int s_kept;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = count_all[b];
    const float4* boxes = boxes_all + (size_t)b * cap;
    // thr * (1 +- 2^-20); thresholds at / near zero take the IEEE division for every pair (see nms_greedy_kernel)
    const bool band = thr_f > 1.0e-6f;
    const float thr_hi = band ? __fmul_rn(thr_f, 1.0000009537f) : INFINITY, thr_lo = band ? __fmul_rn(thr_f, 0.9999990463f) : -INFINITY;
    // The chunk's 32 candidate boxes live in shared memory, double-buffered: the next chunk's boxes are requested at the top of an
    // iteration and parked at its end, so neither the per-chunk tests nor the serial resolution wait on a global load (the sweep is a
    // chain of dependent chunks: its length x the latency per chunk IS the kernel time).
    __shared__ float4 s_cbox[2][32];
    if (tid == 0) s_kept = 0;
    if (tid < 32) s_cbox[0][tid] = tid < n ? boxes[tid] : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    int kept = 0;
    for (int c0 = 0, cur = 0; c0 < n && kept < max_det; c0 += 32, cur ^= 1) {
        const int m = min(32, n - c0);
        const bool ldn = tid < 32 && c0 + 32 + tid < n;
        float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ldn) nxt = boxes[c0 + 32 + tid];      // in flight during this chunk
        if (tid == 0) s_sup = 0u;
        if (tid < 32) s_row[tid] = 0u;
        __syncthreads();
        // candidate `lane` of the chunk
        const bool cvalid = lane < m;
        const float4 cj = cvalid ? s_cbox[cur][lane] : make_float4(0.f, 0.f, 0.f, 0.f);
        // (a) against the kept list: warp w takes kept boxes w, w + 32, ...
        bool sup = false;
        if (cvalid) {
            for (int i = warp; i < kept && !sup; i += 32) sup = iou_exceeds(kbox[i], karea[i], cj, thr_f, thr_hi, thr_lo);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, sup);
        if (lane == 0 && bal) atomicOr(&s_sup, bal);
        // (b) inside the chunk: warp a = earlier candidate, lane = later candidate
        {
            const int a = warp;
            bool s2 = false;
            if (a < m && cvalid && lane > a) {
                const float4 ba = s_cbox[cur][a];
                const float aarea = __fmul_rn(__fsub_rn(ba.z, ba.x), __fsub_rn(ba.w, ba.y));
                s2 = iou_exceeds(ba, aarea, cj, thr_f, thr_hi, thr_lo);
            }
            const unsigned row = __ballot_sync(0xffffffffu, s2);
            if (lane == 0) s_row[a] = row;
        }
        __syncthreads();
        // (c) serial resolution of the chunk
        if (tid == 0) {
            uint32_t alive = ~s_sup & (m == 32 ? 0xffffffffu : ((1u << m) - 1u));
            int k = kept;
            while (alive && k < max_det) {
                const int a = __ffs(alive) - 1;
                alive &= ~(1u << a);
                alive &= ~s_row[a];
                const float4 ba = s_cbox[cur][a];
                kbox[k] = ba;
                karea[k] = __fmul_rn(__fsub_rn(ba.z, ba.x), __fsub_rn(ba.w, ba.y));
                keep_idx[k] = c0 + a;
                ++k;
            }
            s_kept = k;
        }
        if (tid < 32) s_cbox[cur ^ 1][tid] = nxt;
        __syncthreads();
        kept = s_kept;
    }
    // emit detections [x1,y1,x2,y2,conf,cls] (un-offset boxes) + the candidate index of each keep
    float* det = det_all + (size_t)b * max_det * 6;
    int* index = index_all + (size_t)b * max_det;
    for (int t = tid; t < max_det; t += GREEDY_THREADS) {
        if (t < kept) {
            const int r = keep_idx[t];
            const int id = order_all[(size_t)b * cap + r];
            index[t] = id;
            if (pred == nullptr) continue;  // plain box-NMS mode: keep indices only
            const int a = multi_label ? id / nc : id;
            const int c = multi_label ? id % nc : cls_all[(size_t)b * A + a];
            const float4 q = load_xyxy(pred, b, nc, A, a);
            det[t * 6 + 0] = q.x; det[t * 6 + 1] = q.y; det[t * 6 + 2] = q.z; det[t * 6 + 3] = q.w;
            det[t * 6 + 4] = sscore_all[(size_t)b * cap + r];
            det[t * 6 + 5] = (float)c;
        } else {
            index[t] = -1;
            if (pred == nullptr) continue;
#pragma unroll
            for (int k = 0; k < 6; ++k) det[t * 6 + k] = 0.0f;
        }
    }
    if (tid == 0) nkeep_all[b] = kept;
}

struct NmsWs {
    float* conf; int* cls; uint32_t *k0, *v0, *k1, *v1; int* order; float* sscore; int* count; float4* boxes;
    size_t total;
};

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

NmsWs carve(char* base, int B, int nc, int A, int cap, int multi_label) {
    NmsWs w;
    const size_t M = multi_label ? (size_t)A * nc : (size_t)A;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align256(bytes); return p; };
    w.conf = (float*)take(sizeof(float) * B * M);
    w.cls = (int*)take(sizeof(int) * (size_t)B * A);
    w.k0 = (uint32_t*)take(4 * B * M);
    w.v0 = (uint32_t*)take(4 * B * M);
    w.k1 = (uint32_t*)take(4 * B * M);
    w.v1 = (uint32_t*)take(4 * B * M);
    w.order = (int*)take(sizeof(int) * (size_t)B * cap);
    w.sscore = (float*)take(sizeof(float) * (size_t)B * cap);
    w.count = (int*)take(sizeof(int) * (size_t)B);
    w.boxes = (float4*)take(sizeof(float4) * (size_t)B * cap);
    w.total = off;
    return w;
}

}  // namespace

extern "C" {

long long ysod_nms_workspace_bytes(int B, int nc, int A, int max_nms, int multi_label) {
    const long long M = multi_label ? (long long)A * nc : (long long)A;
    const int cap = (int)(M < max_nms ? M : max_nms);
    return (long long)carve(nullptr, B, nc, A, cap > 0 ? cap : 1, multi_label).total;
}

// pred: (B, 4+nc, A) fp32 device, xywh + class scores (what Detect._inference returns, head.py:100-131).
// thr_f: largest float <= the (double) IoU threshold. classes: device int array or NULL.
// out_det: (B, max_det, 6) fp32; out_index: (B, max_det) int32 candidate ids (anchor, or anchor*nc+cls if
// multi_label); out_count: (B) int32. Asynchronous on `stream`; no host sync; no global state.
int ysod_nms_batched(const float* pred, int B, int nc, int A, float conf_thres, float thr_f, const int* classes,
                     int n_classes, int agnostic, int multi_label, int max_det, int max_nms, float max_wh,
                     float* out_det, int* out_index, int* out_count, void* workspace, long long workspace_bytes,
                     cudaStream_t stream) {
    YSOD_CHECK_ARG(pred && out_det && out_index && out_count && workspace, "ysod_nms_batched: null pointer");
    YSOD_CHECK_ARG(B > 0 && nc > 0 && A > 0 && max_det > 0 && max_nms > 0, "ysod_nms_batched: bad sizes");
    YSOD_CHECK_ARG(conf_thres >= 0.0f && conf_thres <= 1.0f, "Invalid Confidence threshold %f", conf_thres);
    const long long M = multi_label ? (long long)A * nc : (long long)A;
    const int cap = (int)(M < max_nms ? M : max_nms);
    NmsWs w = carve((char*)workspace, B, nc, A, cap, multi_label);
    if ((long long)w.total > workspace_bytes) {
        ysod_set_error("ysod_nms_batched: workspace too small (%lld < %zu)", workspace_bytes, w.total);
        return YSOD_ERR_WORKSPACE;
    }
    const size_t greedy_base = (size_t)((((cap + 31) >> 5) + max_det + 3) & ~3) * 4;
    YSOD_CHECK_ARG(greedy_base <= 96 * 1024, "ysod_nms_batched: max_nms/max_det too large for shared memory");
    const int box_cache = cap < 6144 ? cap : 6144;   // up to 96 KB of sorted boxes cached per image
    const size_t greedy_smem = greedy_base + (size_t)box_cache * 16;
    {
        dim3 grid(ysod_cdiv(A, 256), B);
        ysod_launch(nms_score_kernel, grid, 256, 0, stream, pred, nc, A, conf_thres, classes, n_classes, multi_label, w.conf, w.cls);
        YSOD_LAUNCH_CHECK();
    }
    ysod_launch(nms_sort_kernel, B, SORT_THREADS, 0, stream, w.conf, (int)M, cap, 0, w.k0, w.v0, w.k1, w.v1, w.order, w.sscore, w.count);
    YSOD_LAUNCH_CHECK();
    {
        dim3 grid(ysod_cdiv(cap, 256), B);
        ysod_launch(nms_gather_kernel, grid, 256, 0, stream, pred, nc, A, cap, multi_label, w.order, w.cls, w.count,
                                                    agnostic ? 0.0f : max_wh, w.boxes);
        YSOD_LAUNCH_CHECK();
    }
    if (max_det <= 1990) {   // kept-list formulation (24 B per kept box of dynamic shared memory + 1.2 KB static <= 48 KB): the detection path (max_det = 300)
        ysod_launch(nms_chunk_kernel, B, GREEDY_THREADS, (size_t)max_det * 24, stream, pred, nc, A, cap, multi_label, max_det, thr_f, w.boxes, w.order,
                                                                              w.sscore, w.cls, w.count, out_det, out_index, out_count);
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
    if (greedy_smem > 48 * 1024) {
        YSOD_CUDA(cudaFuncSetAttribute(nms_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)greedy_smem));
    }
    ysod_launch(nms_greedy_kernel, B, GREEDY_THREADS, greedy_smem, stream, pred, nc, A, cap, multi_label, max_det, thr_f, w.boxes,
                                                                  w.order, w.sscore, w.cls, w.count, out_det, out_index,
                                                                  out_count, box_cache);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

// Plain greedy NMS on explicit boxes: the drop-in for `torchvision.ops.nms(boxes, scores, iou)` (ops.py:296).
// boxes (n,4) xyxy fp32, scores (n) fp32, both on the device. keep_out: int32[max_keep] indices into boxes,
// score-descending (stable), -1 padded; nkeep_out: int32[1]. workspace >= ysod_nms_boxes_workspace_bytes(n).
long long ysod_nms_boxes_workspace_bytes(int n) {
    const size_t m = (size_t)(n > 0 ? n : 1);
    return (long long)(4 * align256(4 * m) + align256(4 * m) + align256(4 * m) + align256(4) + align256(16 * m));
}

__global__ void nms_reorder_boxes_kernel(const float4* __restrict__ boxes, const int* __restrict__ order, const int* __restrict__ count,
                                         float4* __restrict__ sorted) {
    ysod_pdl_sync();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < count[0]) sorted[r] = boxes[order[r]];
}

int ysod_nms_boxes(const float* boxes, const float* scores, int n, float thr_f, int max_keep, int* keep_out, int* nkeep_out,
                   void* workspace, long long workspace_bytes, cudaStream_t stream) {
    YSOD_CHECK_ARG(keep_out && nkeep_out && workspace && max_keep > 0, "ysod_nms_boxes: null pointer");
    if (n <= 0) {
        YSOD_CUDA(cudaMemsetAsync(nkeep_out, 0, sizeof(int), stream));
        YSOD_CUDA(cudaMemsetAsync(keep_out, 0xff, sizeof(int) * (size_t)max_keep, stream));
        return YSOD_OK;
    }
    YSOD_CHECK_ARG(boxes && scores, "ysod_nms_boxes: null pointer");
    YSOD_CHECK_ARG(((uintptr_t)boxes % 16) == 0, "ysod_nms_boxes: boxes must be 16 B aligned");
    if (ysod_nms_boxes_workspace_bytes(n) > workspace_bytes) {
        ysod_set_error("ysod_nms_boxes: workspace too small");
        return YSOD_ERR_WORKSPACE;
    }
    char* base = (char*)workspace;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base + off; off += align256(bytes); return p; };
    uint32_t* k0 = (uint32_t*)take(4 * (size_t)n);
    uint32_t* v0 = (uint32_t*)take(4 * (size_t)n);
    uint32_t* k1 = (uint32_t*)take(4 * (size_t)n);
    uint32_t* v1 = (uint32_t*)take(4 * (size_t)n);
    int* order = (int*)take(4 * (size_t)n);
    float* sscore = (float*)take(4 * (size_t)n);
    int* count = (int*)take(4);
    float4* sorted = (float4*)take(16 * (size_t)n);
    const size_t greedy_base = (size_t)((((n + 31) >> 5) + max_keep + 3) & ~3) * 4;
    YSOD_CHECK_ARG(greedy_base <= 128 * 1024, "ysod_nms_boxes: n/max_keep too large for shared memory");
    const int box_cache = n < 6144 ? n : 6144;
    const size_t greedy_smem = greedy_base + (size_t)box_cache * 16;
    ysod_launch(nms_sort_kernel, 1, SORT_THREADS, 0, stream, scores, n, n, 1, k0, v0, k1, v1, order, sscore, count);
    YSOD_LAUNCH_CHECK();
    ysod_launch(nms_reorder_boxes_kernel, ysod_cdiv(n, 256), 256, 0, stream, (const float4*)boxes, order, count, sorted);
    YSOD_LAUNCH_CHECK();
    if (max_keep <= 1990) {
        ysod_launch(nms_chunk_kernel, 1, GREEDY_THREADS, (size_t)max_keep * 24, stream, nullptr, 0, 0, n, 0, max_keep, thr_f, sorted, order, sscore, nullptr,
                                                                               count, nullptr, keep_out, nkeep_out);
        YSOD_LAUNCH_CHECK();
        return YSOD_OK;
    }
    if (greedy_smem > 48 * 1024) {
        YSOD_CUDA(cudaFuncSetAttribute(nms_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)greedy_smem));
    }
    ysod_launch(nms_greedy_kernel, 1, GREEDY_THREADS, greedy_smem, stream, nullptr, 0, 0, n, 0, max_keep, thr_f, sorted, order, sscore, nullptr,
                                                                  count, nullptr, keep_out, nkeep_out, box_cache);
    YSOD_LAUNCH_CHECK();
    return YSOD_OK;
}

}  // extern "C"
