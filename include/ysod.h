/* ysod.h -- C ABI of libysod.so: hand-written sm_100a kernels for the YOLOv12-SOD inference hot path.
 *
 * The reference (quitedob/yolo-sod) is 100% Python and has no C/FFI/plugin ABI; its seams are the
 * Python objects `DetectionModel.forward` and `ops.non_max_suppression`. This header is therefore the boundary a
 * maintainer would bind from Python with ctypes (INTEGRATION.md shows the stub); each entry point cites the reference
 * code it replaces (paths relative to the reference checkout).
 *
 * Conventions: plain pointers and sizes, no torch types. Every function returns 0 on success, non-zero on error
 * (ysod_last_error() gives the text), is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, keeps no
 * global state, performs no hidden host synchronisation and allocates nothing -- the caller owns all buffers.
 * Device only: there is no CPU fallback. Activations are NHWC "views": element (n,h,w,c) at base[((n*H+h)*W+w)*cs + c]
 * with `cs` = pixel stride in elements (>= C), which is how producers write into channel slices of a consumer's buffer.
 * dtype: 0 = fp32, 1 = bf16 (storage; all math is fp32 / fp32-accumulate). act: 0 none 1 SiLU 2 GELU(erf) 3 ReLU
 * 4 sigmoid 5 h-sigmoid. Channel counts and pixel strides of vectorised kernels must be multiples of 8.
 */
#ifndef YSOD_H
#define YSOD_H
#ifdef __cplusplus
extern "C" {
#endif

int ysod_version(void);
const char* ysod_last_error(void);
int ysod_compiled_arch(void); /* 100 = sm_100a */
/* 16-bit storage type of the loaded build: 1 = bf16 (libysod.so), 2 = IEEE fp16 (libysod_f16.so: the same sources compiled with
 * -DYSOD_HALF=1, the reference's `half=True` / `model.half()` mode, nn/autobackend.py:154). dtype code 1 in every entry point below
 * means "this build's 16-bit type". */
int ysod_storage_dtype(void);

/* ---- NMS: ultralytics/utils/ops.py:167-316 non_max_suppression + torchvision.ops.nms (ops.py:296) ---------------- */
long long ysod_nms_workspace_bytes(int B, int nc, int A, int max_nms, int multi_label);
/* pred (B,4+nc,A) fp32 xywh+scores; thr_f = largest float <= IoU threshold; classes = device int[n_classes] or NULL.
 * out_det (B,max_det,6) [x1,y1,x2,y2,conf,cls]; out_index (B,max_det) kept candidate ids; out_count (B). */
int ysod_nms_batched(const float* pred, int B, int nc, int A, float conf_thres, float thr_f, const int* classes,
                     int n_classes, int agnostic, int multi_label, int max_det, int max_nms, float max_wh,
                     float* out_det, int* out_index, int* out_count, void* workspace, long long workspace_bytes, void* stream);

/* torchvision.ops.nms(boxes, scores, iou) drop-in (call site ops.py:296): boxes (n,4) xyxy, scores (n), device fp32.
 * keep_out int32[max_keep] (score-descending, -1 padded), nkeep_out int32[1]. */
long long ysod_nms_boxes_workspace_bytes(int n);
int ysod_nms_boxes(const float* boxes, const float* scores, int n, float thr_f, int max_keep, int* keep_out, int* nkeep_out,
                   void* workspace, long long workspace_bytes, void* stream);

/* ---- Detect decode: head.py:100-131 _inference, block.py:64-83 DFL, tal.py:333-357 make_anchors/dist2bbox --------- */
int ysod_dfl_decode(const void* raw, int dtype, int B, int H, int W, int cs, int nc, int reg_max, float stride, float* y,
                    int A_total, int a_off, void* stream);

/* ---- tensor-core implicit-GEMM conv / linear: conv.py:37-55 Conv (BN folded, torch_utils.py:238-265),
 *      block.py:233-356 C2f/Bottleneck, head.py:43-57 Detect stacks, nn.MultiheadAttention / nn.Linear GEMMs ---------- */
typedef struct ysod_conv_tc ysod_conv_tc;
int ysod_conv_tc_create(ysod_conv_tc** handle, const void* x, int N, int H, int W, int Cin, int xcs, const void* wgt,
                        const float* bias, int Cout, int Cout_pad, int ksize, int stride, void* out, int out_dtype, int ocs,
                        const void* res, int rcs, int act);
/* mode: 0 auto, 1 generic per-tap kernel, 2 force the 3x3/s1 halo-reuse kernel; | 0x40 (YSOD_CONV_UP2): fuse the following
 * nn.Upsample(scale_factor=2, mode="nearest") (yaml neck rows `[-1, 1, nn.Upsample, [None, 2, "nearest"]]`) into the store --
 * `out` is then the N x 2Ho x 2Wo x Cout destination view and every output pixel is written to its 2 x 2 block. */
#define YSOD_CONV_UP2 0x40
/* | 0x80 (YSOD_CONV_IMG_WEIGHTS): `wgt` is [N][Cout_pad][K], one weight matrix per image (see ysod_scale_weights). */
#define YSOD_CONV_IMG_WEIGHTS 0x80
/* | 0x10 (YSOD_CONV_NO_SPLIT_STAGING): A/B switch for measurements -- keeps the full-tile epilogue staging buffer instead of the
 * one-unit buffer that frees shared memory for a deeper operand ring on deep-K layers (tc_conv.cu). Results are identical. */
#define YSOD_CONV_NO_SPLIT_STAGING 0x10
/* | 0x04 (YSOD_CONV_NO_PAIR): A/B switch -- disables the tile-pair plan (two raster-adjacent output tiles share every weight fetch:
 * tc_conv.cu TcParams::pair) that deep-K layers use by default. Results are identical. */
#define YSOD_CONV_NO_PAIR 0x04
/* | 0x08 (YSOD_CONV_FORCE_PAIR): A/B switch -- also pairs tiles on the generic (per-tap) kernel for deep-K layers, where it is off
 * by default (measured neutral: the steady-state gain is cancelled by the coarser last wave). */
#define YSOD_CONV_FORCE_PAIR 0x08
/* | 0x20 (YSOD_CONV_NO_STORE): the plan's own output tensor is not written; only meaningful together with
 * ysod_conv_tc_set_decode -- the predict path consumes `y` alone (detect/predict.py:25-32 takes preds[0]), so the fp32 raw maps
 * Detect.forward also returns (head.py:74) need not be materialised. `out` must still be a valid 16 B aligned device address. */
#define YSOD_CONV_NO_STORE 0x20
/* | 0x10000 (YSOD_CONV_NO_DUO): A/B switch -- 3x3 / stride-1 convs with 32 input and 32 output channels (C2f Bottlenecks at P2,
 * block.py:343-356) normally run the halo kernel's pixel-duo plan (one MMA row = two adjacent output pixels, N = 64, 128 B operand
 * rows: tc_conv.cu TcParams::duo); this bit keeps the 32-channel plan (N = 32, 64 B rows), which ysod_conv_tc_set_b2b_cat needs. */
#define YSOD_CONV_NO_DUO 0x10000
int ysod_conv_tc_create_ex(ysod_conv_tc** handle, const void* x, int N, int H, int W, int Cin, int xcs, const void* wgt,
                           const float* bias, int Cout, int Cout_pad, int ksize, int stride, void* out, int out_dtype, int ocs,
                           const void* res, int rcs, int act, int mode);
/* head.py:100-131 Detect._inference + block.py:64-83 DFL + tal.py:333-357 make_anchors/dist2bbox fused into the epilogue of the
 * level's final 1x1 head conv (plan: fp32 raw map out, no activation, Cout = 64 + nc): besides the raw map the launch writes
 * y (B, 4+nc, A_total) for anchors [a_off, a_off + Ho*Wo). Equivalent to running ysod_dfl_decode on the raw map afterwards. */
int ysod_conv_tc_set_decode(ysod_conv_tc* h, float* y, int A_total, int a_off, int nc, float stride);
/* Back-to-back GEMM: head.py:43-57 `Conv(c, c, 3)` followed by `nn.Conv2d(c, 4*reg_max | nc, 1)` (Detect cv2[i][1..2] / cv3[i][1..2])
 * and the level's share of the decode (head.py:100-131) in ONE launch: the staged bf16 tile of the 3x3 conv is the A operand of a
 * second tcgen05.mma group against w2 [n2][64] bf16 (+ bias2), whose epilogue writes y (kind 1: DFL box branch -> y[0..4);
 * kind 2: class branch -> sigmoid -> y[4..4+nc)) and, when `raw` is given, the fp32 raw map channels [raw_coff, ...). The 3x3 conv's own
 * output is never stored. Plan requirements: Cout 64, bf16 output, no fused upsample. */
int ysod_conv_tc_set_b2b(ysod_conv_tc* h, const void* w2, const float* bias2, int n2, int kind, float* y, int A_total, int a_off, int nc,
                         float stride, float* raw, int raw_cs, int raw_coff);
/* Plain back-to-back GEMM: a following `Conv(64, 64, 1)` + BN + act whose only input is this plan's output (block.py:233-248 C2f.cv1
 * right after a Conv layer) as the second MMA group of the same launch; create the plan with `out` = that second layer's destination. */
int ysod_conv_tc_set_b2b_conv(ysod_conv_tc* h, const void* w2, const float* bias2, int act2);
/* Back-to-back GEMM over a concatenation: block.py:233-248 `C2f.cv2(cat(cv1 output, bottleneck output))` (1x1, 96 -> 64) inside the
 * launch of the block's last Bottleneck conv (3x3 32 -> 32, halo plan): K = 64 channels of x2 (TMA tile per output tile) + the staged
 * 32-channel tile. w2 [64][96] bf16 in torch.cat order; out2 = the C2f output view. */
int ysod_conv_tc_set_b2b_cat(ysod_conv_tc* h, const void* x2, int x2cs, const void* w2, const float* bias2, int act2, void* out2, int out2cs);
/* SE (smallobj_modules.py:57-92) folded into the conv that consumes it: conv(x * a[n]) == conv with input-channel columns of the
 * weights scaled by a[n]. w: [rows][K] fp32 (K ordered (r,s,cin), BN folded); gate: [N][Cin] fp32; out: [N][rows][K] bf16. */
int ysod_scale_weights(const float* w, int rows, int K, int Cin, const float* gate, int N, void* out, void* stream);
int ysod_conv_tc_run(ysod_conv_tc* handle, void* stream);
int ysod_conv_tc_info(ysod_conv_tc* handle, int* out8);
void ysod_conv_tc_destroy(ysod_conv_tc* handle);
/* profiling aid (not part of the drop-in surface): pipeline trace of the last launch created with mode = 32 << 8 */
int ysod_debug_trace(unsigned long long* out, int cap);

/* ---- CUDA-core convs: stem (Cin=3, NCHW fp32 image in), depthwise, grouped, and the fp32 parity mode ---------------- */
int ysod_conv_direct(const void* x, int dtype, int N, int H, int W, int Cin, int xcs, const void* w, const float* bias,
                     int Cout, int k, int s, int pad, int groups, void* out, int out_dtype, int ocs, const void* res, int rcs,
                     int act, void* stream);
int ysod_dwconv(const void* x, int dtype, int N, int H, int W, int C, int xcs, const void* w, const float* bias, int k, int s,
                int pad, void* out, int ocs, const void* res, int rcs, int act, void* stream);
int ysod_stem_conv(const float* img, int N, int H, int W, const float* w, const float* bias, int Cout, int k, int s, int pad,
                   void* out, int out_dtype, int ocs, int act, void* stream);

/* tensor-core stem (3x3 / stride 2 / pad 1, Cout 16|32|64, bf16 out). src_fmt 0: img = (N,3,H,W) fp32 in [0,1] (tasks.py:129);
 * src_fmt 1: img = (N,H,W,3) uint8 BGR frames, with BasePredictor.preprocess (engine/predictor.py:116-134: BGR->RGB, HWC->CHW,
 * /255) fused into the load. wk = [Cout][32] bf16, column (r*3+s)*3+c, columns 27..31 zero. */
#define YSOD_STEM_INDIRECT 0x10 /* src_fmt | 0x10: `img` is a device slot (void**) holding the image pointer, see ysod_set_ptr */
int ysod_stem_mma(const void* img, int src_fmt, int N, int H, int W, const void* wk, const float* bias, int Cout, void* out, int ocs,
                  int act, void* stream);
/* The same stem conv that also leaves the global-average-pool partial sums of its output for the SE_Block that follows it in the SOD
 * YAMLs (smallobj_modules.py SE_Block.forward: avg_pool -> fc1 -> ReLU -> fc2 -> sigmoid): psum[N][S][Cout] fp32, S = ceil(W/2 / 64) *
 * ceil(H/2 / 4) tiles per image, each the sum of the tile's stored 16-bit outputs; ysod_se_gate(psum, N, S, ...) consumes it, so the SE
 * block does not read the stem's map again for its pooling. */
int ysod_stem_mma_gap(const void* img, int src_fmt, int N, int H, int W, const void* wk, const float* bias, int Cout, void* out,
                      int ocs, int act, float* psum, void* stream);

/* Binds the input of a captured forward without a staging copy (the reference's forward reads the caller's tensor in place,
 * nn/tasks.py:129-163): stores `value` in the device pointer slot that ysod_stem_mma(..., src_fmt | YSOD_STEM_INDIRECT) reads. */
int ysod_set_ptr(void* slot, const void* value, void* stream);

/* ---- predictor glue ---------------------------------------------------------------------------
 * engine/predictor.py:145-164 pre_transform -> LetterBox(imgsz, auto, stride): resize (cv2.INTER_LINEAR, bit-exact 8-bit fixed
 * point) + constant border; LetterBox's source (ultralytics/data/augment.py) is absent from the reference checkout, the geometry
 * follows upstream 8.3.63 and is computed by the host. frames: (B,H0,W0,3) uint8 BGR; out: (B,H,W,3) uint8 (feeds ysod_stem_mma). */
int ysod_letterbox_u8(const void* frames, int B, int H0, int W0, void* out, int H, int W, int new_h, int new_w, int top, int left,
                      int value, void* stream);
/* utils/ops.py:92-127 scale_boxes + :319-338 clip_boxes on the padded NMS output (models/yolo/detect/predict.py:38-40), in place.
 * det: (B, rows_per_img, row_stride >= 4) fp32 rows [x1,y1,x2,y2,...]; params: per image params_stride (>= 5) floats = gain, pad_x,
 * pad_y, orig_w, orig_h. */
int ysod_scale_boxes(float* det, int B, int rows_per_img, int row_stride, const float* params, int params_stride, void* stream);

/* ---- SE: smallobj_modules.py:57-92 ; CBAM: cbam_block.py:8-55 ; CoordAtt: ca_block.py:16-59 ------------------------- */
int ysod_gap_partial(const void* x, int dtype, int N, int HW, int C, int xcs, int S, float* psum, float* pmax, void* stream);
/* pool + gate in one launch (the last CTA of an image runs the gate MLP): kind 0 = SE (smallobj_modules.py:84-91), kind 1 = CBAM
 * channel attention (cbam_block.py:14-23). counter: N zero-initialised uint32, left at zero by every call. */
int ysod_gap_gate(const void* x, int dtype, int N, int HW, int C, int xcs, int S, float* psum, float* pmax, void* counter, int kind,
                  const float* w1, const float* b1, const float* w2, const float* b2, int hid, float* gate, void* stream);
int ysod_se_gate(const float* psum, int N, int S, int HW, int C, const float* w1, const float* b1, const float* w2,
                 const float* b2, int hid, float* gate, void* stream);
int ysod_cbam_gate(const float* psum, const float* pmax, int N, int S, int HW, int C, const float* w1, const float* w2, int hid,
                   float* gate, void* stream);
int ysod_scale_channels(const void* x, int dtype, int N, int HW, int C, int xcs, const float* gate, void* out, int ocs, void* stream);
int ysod_cbam_stats(const void* x, int dtype, int N, int HW, int C, int xcs, const float* gate, float* stats, void* stream);
int ysod_cbam_apply(const void* x, int dtype, int N, int H, int W, int C, int xcs, const float* gate, const float* stats,
                    const float* wsp, int ks, void* out, int ocs, void* stream);
/* CBAM spatial attention in one pass (channel statistics of the tile + halo, 7x7 conv, apply): with ysod_gap_partial + ysod_cbam_gate
 * the block reads the map twice and writes it once; ysod_cbam_stats + ysod_cbam_apply remain as the general / A-B path. */
int ysod_cbam_spatial(const void* x, int dtype, int N, int H, int W, int C, int xcs, const float* gate, const float* wsp, int ks, void* out,
                      int ocs, void* stream);
/* workspace: ysod_ca_pool_workspace_floats() floats (single pass over the map: row means + per-row-block column partials, then a
 * fixed-order column reduction), or NULL for the two-pass kernel. */
long long ysod_ca_pool_workspace_floats(int N, int H, int W, int C);
int ysod_ca_pool(const void* x, int dtype, int N, int H, int W, int C, int xcs, float* pooled, float* workspace, void* stream);
int ysod_ca_gate(const float* pooled, int N, int H, int W, int C, int mip, const float* w1, const float* b1, const float* wh,
                 const float* bh, const float* ww, const float* bw, float* att, void* stream);
int ysod_ca_apply(const void* x, int dtype, int N, int H, int W, int C, int xcs, const float* att, void* out, int ocs, void* stream);

/* ---- SPPF pooling: block.py:178-197 ; nn.Upsample(nearest)+Concat: tasks.py:184, conv.py:323-334 -------------------- */
int ysod_sppf_pool(const void* y0, int dtype, int N, int H, int W, int C, int xcs, int k, void* o1, void* o2, void* o3, int ocs,
                   void* stream);
int ysod_upsample_copy(const void* x, int dtype, int N, int H, int W, int C, int xcs, int scale, void* out, int ocs, void* stream);

/* ---- MambaBlock with its GLU fallback (blocks_mamba.py:84-103, 167-236; what the reference runs when mamba_ssm is missing):
 *      F.avg_pool2d(y, r, r) ; sigmoid(g) * a of `pw1(x).chunk(2, 1)` ; x + F.interpolate(y, size=(H, W), mode="nearest").
 *      The 1x1 / depthwise convs of the block go through ysod_conv_tc_* / ysod_dwconv. ------------------------------------- */
int ysod_avgpool2d(const void* x, int dtype, int N, int H, int W, int C, int xcs, int r, void* out, int ocs, void* stream);
int ysod_glu(const void* x, int dtype, long long npix, int hid, int xcs, void* out, int ocs, void* stream);
int ysod_upsample_add(const void* y, int dtype, int N, int Hh, int Wh, int C, int ycs, const void* res, int rcs, int H, int W, void* out,
                      int ocs, void* stream);

/* ---- Swin window attention plumbing: blocks_transformer.py:8-79,98-131 ; A2_Attn pooling/upsample: a2_attn.py:44-60 - */
int ysod_layernorm(const void* x, int dtype, long long rows, int C, int ldx, const float* gamma, const float* beta, float eps,
                   void* out, int ldo, void* stream);
int ysod_window_partition_ln(const void* x, int dtype, int N, int H, int W, int C, int xcs, int wh, int ww, int nWh, int nWw,
                             const float* gamma, const float* beta, float eps, void* raw_out, void* norm_out, int ldo, void* stream);
int ysod_window_reverse(const void* tok, int dtype, int ldt, int N, int H, int W, int C, int wh, int ww, int nWh, int nWw, void* out,
                        int ocs, void* stream);
int ysod_adaptive_pool_rows(const void* x, int dtype, int N, int H, int W, int C, int xcs, int OH, void* out, int ocs, void* stream);
int ysod_bilinear_rows(const void* x, int dtype, int N, int IH, int W, int C, int xcs, int OH, void* out, int ocs, void* stream);

/* SwinBlock with the tokens kept in NHWC pixel order (swin_nhwc.cu; blocks_transformer.py:133-171 around the linears):
 * ysod_dwconv3_ln = dw 3x3 (no bias) + LayerNorm 1 in one pass (y = raw tokens for the residual, yn = normalised tokens; w [3][3][C]
 * fp32, C 256 | 512, 16-bit storage); ysod_mha_window_nhwc = the window attention core reading q / k / v rows and writing its output
 * rows through the window -> pixel map (ws*ws <= 64 tokens, head_dim 32 | 64); kpad / vpad [heads*D] = key / value of a zero-padded
 * window token = in_proj_{k,v}(LayerNorm(0) = beta) + bias. Replace ysod_dwconv + ysod_window_partition_ln ... ysod_window_reverse. */
int ysod_dwconv3_ln(const void* x, int N, int H, int W, int C, int xcs, const float* w, const float* gamma, const float* beta, float eps,
                    void* y, int ycs, void* yn, int ncs, void* stream);
int ysod_mha_window_nhwc(const void* q, const void* k, const void* v, int ld, int N, int H, int W, int ws, int heads, int D, const void* kpad,
                         const void* vpad, float scale, void* out, int ldo, void* stream);

/* Fully fused SwinBlock (blocks_transformer.py:133-171) for C = 64, 2 heads, 7x7 windows (the P2 level): dw3x3 -> window partition
 * (zero padded) -> x + MHA(LN x) -> x + MLP(LN x) -> window reverse / crop -> pw1x1 + BN + SiLU + identity, one kernel.
 * wbf16 (37440 bf16): dw[3][3][64] | in_proj_weight[192][64] | out_proj.weight[64][64] | mlp.0.weight[128][64] | mlp.2.weight[64][128] |
 * pw.weight with BN folded [64][64].  pf32 (768 fp32): norm1.weight | norm1.bias | in_proj_bias | out_proj.bias | norm2.weight |
 * norm2.bias | mlp.0.bias | mlp.2.bias | folded BN bias. */
/* Pre-folded by the caller: norm1 gamma / beta into in_proj, norm2's into mlp.0, log2(e) / sqrt(head_dim) into in_proj's Q rows; the
 * norm slots of pf32 are ignored (swin_fused.cu). */
int ysod_swin64_fused(const void* x, int N, int H, int W, int xcs, const void* wbf16, const float* pf32, void* out, int ocs, int window,
                      int heads, void* stream);
/* The same block (same blobs, same rounding points) with every contraction on tcgen05 / TMEM (swin_tc.cu), the default: two windows per
 * M = 128 tile, thread = token row = TMEM lane, in_proj / QK^T / PV / out_proj / MLP / pw as tcgen05.mma groups from shared-memory operands,
 * LayerNorm / softmax / GELU / SiLU on tcgen05.ld registers as packed fp32 pairs (fma.rn.f32x2), the 9 x 9 input patches, the identity rows
 * and the output tile moved by TMA (out-of-bounds fill = conv padding, store clipping = window_reverse's crop). x / out are 16 B aligned
 * NHWC views (any pixel stride that is a multiple of 8). The K and V thirds of the in_proj bias (pf32[192..320)) must be zero: the caller
 * drops the key bias (softmax-invariant) and folds the value bias into out_proj's (bo += Wo bv). ysod_swin64_fused (mma.sync) stays as the A/B baseline. */
int ysod_swin64_tc(const void* x, int N, int H, int W, int xcs, const void* wbf16, const float* pf32, void* out, int ocs, int window,
                   int heads, void* stream);
/* profiling only: enable / disable the stage trace of the tcgen05 SwinBlock kernel and (host_out != NULL) read back the previous launch's
 * 8 x 24 clock64 stamps of CTA 0 (tools/prof_swin.py). Replaces nothing in the reference. */
int ysod_swin64_tc_trace(int enable, long long* host_out);

/* ---- softmax attention core: nn.MultiheadAttention internals (blocks_transformer.py:116, a2_attn.py:53) and the manual
 *      path of AAttn (block.py:1348-1357). q/k/v addressed as ptr + batch*bs + token*ld + head*D (elements). ----------- */
int ysod_mha_core(const void* q, const void* k, const void* v, int dtype, int batch, int L, int heads, int D, int ldq, int ldk,
                  int ldv, long long bsq, long long bsk, long long bsv, float scale, void* out, int ldo, long long bso, void* stream);
/* impl 0 = auto: QK^T and PV as tcgen05.mma with TMEM accumulators (attention_tc.cu) for 16-bit storage and head_dim 32 / 64 --
 * two <= 64-token windows packed into one M = 128 tile, or 128-query tiles streaming 128-key tiles with an online softmax -- else
 * the mma.sync / CUDA-core kernels; impl 1 = the mma.sync / CUDA-core kernels only (A/B baseline); impl 2 = tcgen05 required. */
int ysod_mha_core_ex(const void* q, const void* k, const void* v, int dtype, int batch, int L, int heads, int D, int ldq, int ldk,
                     int ldv, long long bsq, long long bsk, long long bsv, float scale, void* out, int ldo, long long bso, int impl,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YSOD_H */
