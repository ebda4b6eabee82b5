/* nanovs.h -- C ABI of libnanovs.so: sm_100a kernels for the KP2DTiny perception hot path.
 *
 * The reference (ETH-PBL/Nano-VS-SLAM) is pure Python/PyTorch and has no FFI of its own; the
 * boundary it exposes for this path is the nn.Module surface of KP2DTinyV2/V3
 * (src/kp2dtiny/models/kp2dtiny.py:284,650) plus two third-party call sites
 * (faiss.IndexFlatL2, src/evaluation/global_descriptor.py:55-60; cv2.BFMatcher,
 * src/visual_odometry/feature_matcher.py:248).  Each entry point below names the reference
 * statement(s) it replaces.  The Python package nano_vs_slam_b200 binds these with ctypes and
 * re-creates that module surface on top.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 NCHW-contiguous data unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing synchronises
 *     and nothing allocates: outputs and workspaces are supplied by the caller;
 *   - return value: NVS_OK (0) or a negative NVS_ERR_* code; nvs_last_error() gives the text;
 *   - functions are re-entrant across streams (no hidden global state except the error text).
 */
#ifndef NANOVS_H_
#define NANOVS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NVS_ABI_VERSION 3

#define NVS_OK 0
#define NVS_ERR_ARG (-1)         /* bad shape / null pointer / unsupported size */
#define NVS_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed */
#define NVS_ERR_UNSUPPORTED (-3) /* configuration outside what the kernels implement */
#define NVS_ERR_NO_DEVICE (-4)   /* no sm_100 device visible */

/* activation codes for the conv epilogue */
#define NVS_ACT_NONE 0
#define NVS_ACT_LRELU 1        /* LeakyReLU(0.01)            modules/base.py:33 */
#define NVS_ACT_RELU 2         /* ReLU                       modules/base.py:35 */
#define NVS_ACT_SIGMOID 3      /* score                      kp2dtiny.py:574 */
#define NVS_ACT_TANH 4         /* centre shift               kp2dtiny.py:575 */
#define NVS_ACT_SIGMOID_TANH 5 /* V3 fused head: ch0 sigmoid, rest tanh   kp2dtiny.py:927-935 */
#define NVS_ACT_GELU 6         /* exact erf GELU             modules/segformer.py:184 */

/* where the conv epilogue writes */
#define NVS_OUT_PLAIN 0     /* dst[b, dst_c_off + co, y, x] */
#define NVS_OUT_POOL 1      /* only MaxPool2d(2,2) of the result -> dst2   encoders.py:111 */
#define NVS_OUT_BOTH 2      /* full result -> dst and pooled -> dst2       encoders.py:121-123 */
#define NVS_OUT_SHUFFLE 3   /* PixelShuffle(2): dst[b, dst_c_off + co/4, 2y + (co%4)/2, 2x + co%2]  heads.py:98 */

/* how the conv reads its input */
#define NVS_IN_PLAIN 0
#define NVS_IN_S2D 1        /* 2x2 stride-2 conv expressed as 1x1 over space-to-depth(2): segformer.py:93-95 */
#define NVS_IN_U8_HWC 2     /* src0 = uint8 (B,H,W,3) camera frames; value = (u8 / 255 - 0.5) * 2 on load
                               (visual_odometry.py:283 + frontend.py:79).  Stem layer only: ksize 3, c0 = c0_total = 3,
                               cout 16, NVS_OUT_PLAIN, dst_nhwc */
#define NVS_IN_UNIT 3       /* src0 = fp32 NCHW frames in [0,1]; value = (x - 0.5) * 2 on load (frontend.py:79).  Stem
                               layer only, same constraints as NVS_IN_U8_HWC */

const char* nvs_last_error(void);
int nvs_abi_version(void);
/* 0 when an sm_100 device is current and the kernels can launch, else NVS_ERR_NO_DEVICE. */
int nvs_device_ok(void);

/* ---- conv3x3 / conv1x1 + folded BN + activation (+ pool / pixel-shuffle / concat-read) -----------
 * Replaces AnnotatedConvBnReLUModel.forward (modules/base.py:39-46), the biased plain convs
 * (heads.py:33,96,102; segmentation.py:154,339-342), MaxPool2d (encoders.py:111,123;
 * segmentation.py:132,448), PixelShuffle + torch.cat (heads.py:98-99; segmentation.py:139-149) and
 * the 1x1 / 2x2-s2 projections of the attention block (segformer.py:105-106,136,193-198).
 * The input is the channel-concatenation of slice [c0_off, c0_off+c0) of src0 and (optionally)
 * slice [c1_off, c1_off+c1) of src1, both (B, c*_total, H, W).
 * weight: packed [cin_pad][ksize*ksize][cout_pad] with BN scale folded in; bias: [cout_pad].
 * cin_pad = round_up(c0+c1, chunk), cout_pad = round_up(cout, nvs_conv_cout_tile(cout)). */
typedef struct NvsConvArgs {
  const float* src0;
  const float* src1; /* may be NULL */
  const float* weight;
  const float* bias;
  float* dst;  /* may be NULL for NVS_OUT_POOL */
  float* dst2; /* pooled output for NVS_OUT_POOL / NVS_OUT_BOTH */
  int32_t c0_total, c0_off, c0;
  int32_t c1_total, c1_off, c1;
  int32_t dst_c_total, dst_c_off;
  int32_t dst2_c_total, dst2_c_off;
  int32_t B, H, W; /* OUTPUT plane of the conv (before pool/shuffle); for NVS_IN_S2D the input is (2H+r, 2W+r) */
  int32_t in_H, in_W; /* input plane size (== H, W unless NVS_IN_S2D) */
  int32_t cout;
  int32_t ksize;   /* 3 (pad 1) or 1 */
  int32_t act;     /* NVS_ACT_* */
  int32_t out_mode;/* NVS_OUT_* */
  int32_t in_mode; /* NVS_IN_* */
  int32_t dst_nhwc;  /* 0: NCHW; 1: dst is (B,H,W,dst_c_total) fp32 channels-last (input layout of the 3xTF32 nvs_conv_tc
                        kernels); 2: the SPLIT channels-last format of the 3xFP16 kernels (NvsConvTcArgs.flags bit 4):
                        same footprint, per pixel dst_c_total fp16 values a_hi = fp16(a) followed by dst_c_total fp16
                        values a_lo = fp16(a - a_hi); plain outputs only, dst_c_total and dst_c_off multiples of 8 */
  int32_t dst2_nhwc; /* same for the pooled output */
} NvsConvArgs;

int nvs_conv_cout_tile(int32_t cout);   /* output-channel tile the kernel uses for `cout` */
int nvs_conv_cin_chunk(int32_t cin);    /* input-channel chunk (4 or 8) the kernel uses for `cin` */
int nvs_conv(const NvsConvArgs* args, void* stream);

/* ---- tensor-core 3x3 conv (tcgen05, "3xTF32": fp32-grade accuracy, see csrc/conv_tc.cu) ---------------
 * Same reference statements as nvs_conv, for layers whose input channel counts are multiples of 32 and
 * cout <= 128.  Activations are channels-last: src (B,H,W,c_total) fp32; weights packed as
 * w_hi / w_lo [9][cout_pad][c0+c1] (tap-major, K contiguous; w_hi has the low 13 mantissa bits cleared,
 * w_lo = w - w_hi; BN folded), bias [cout_pad], cout_pad = nvs_conv_tc_cout_pad(cout).
 * dst_mode 0: no full-resolution output, 1: plain, 2: PixelShuffle(2) into NHWC (B,2H,2W,cout/4 at dst_c_off),
 * 3: keypoint-head split (cout = 3): sigmoid(ch 0) -> dst (B,1,H,W), tanh(ch 1,2) -> dst_pool (B,2,H,W), both NCHW
 *    (kp2dtiny.py:574-575, 927-935);
 * dst_layout 0: NHWC, 1: NCHW (plain only); dst_pool (optional): MaxPool2d(2,2) of the result, NHWC.
 * The plan (TMA descriptors + parameters) lives in caller memory of nvs_conv_tc_plan_bytes() bytes. */
typedef struct NvsConvTcArgs {
  const float* src0;
  const float* src1; /* may be NULL */
  const float* w_hi;
  const float* w_lo;
  const float* bias;
  float* dst;      /* may be NULL at plan time when it is supplied per call (dst_override) */
  float* dst_pool; /* may be NULL */
  int32_t c0_total, c0_off, c0;
  int32_t c1_total, c1_off, c1;
  int32_t dst_c_total, dst_c_off, dst_layout, dst_mode;
  int32_t pool_c_total, pool_c_off;
  int32_t B, H, W, cout, act;
  int32_t flags; /* bit 0: single MMA issuer = fixed fp32 accumulation order (bit-reproducible, slower);
                    bit 1: c0 == 16, c1 == 0, cout <= 32 only -- w_hi / w_lo are given in the paired-tap layout
                    [5][cout_pad][32], K row of step t = [tap 2t ch 0-15 | tap 2t+1 ch 0-15], tenth tap zero:
                    a tile then takes 5 pipeline steps instead of 9;
                    bit 2: cout <= 32, c0 and c1 multiples of 32, dst_mode 1 or 3, no pooled output -- the
                    row-stationary kernel: a pipeline step is one kernel ROW, its three taps are one N = 3 x 32 MMA
                    operand ([9][32][cin] weights read as [3][96][cin]) and the epilogue sums the three shifted
                    partial results; 3 pipeline steps per tile and chunk instead of 9;
                    bit 4: the "3xFP16" row-stationary kernel (csrc/conv_rs.cu), cout <= 64, c0 and c1 multiples of 32
                    (or c0 == c0_total == 16, c1 == 0): w_hi / w_lo point to FP16 arrays [3 ky][3 kx x cout_pad][cin]
                    (cout_pad = 32 or 64, cin = c0 + c1 with a 16-channel source padded to 32) holding the fp16 hi /
                    lo parts of w * 2^t, and w_scale = 2^-t; any dst_mode / dst_layout / dst_pool combination.
                    src0 / src1 and every channels-last output (dst with dst_layout 0, dst_pool) are in the SPLIT
                    format (see NvsConvArgs.dst_nhwc = 2, nvs_split16): per pixel c_total fp16 a_hi, then c_total
                    fp16 a_lo, in the 4 c_total bytes of the fp32 layout; NCHW outputs stay fp32.  bit 0 then selects
                    one MMA-issuing thread instead of three (bit-reproducible) */
  int32_t c0_real, c1_real; /* 0, or the number of leading channels of the c0 / c1 window that can be non-zero (the
                    rest is zero padding with zero weights, e.g. 24 real channels in a 32-channel row for the N
                    letters): MMA k-steps that would only multiply padding are skipped */
  float w_scale; /* flags bit 4 only: factor applied to the accumulator before the bias (undoes the weights' 2^t) */
} NvsConvTcArgs;
int32_t nvs_conv_tc_cout_pad(int32_t cout);
int32_t nvs_conv_tc_supported(int32_t c0, int32_t c1, int32_t cout);
size_t nvs_conv_tc_plan_bytes(void);
int nvs_conv_tc_plan_init(void* plan, const NvsConvTcArgs* args);
/* flags bit 4 kernels only: 1 if an activation written since the last reset left the fp16 range (|x| >= 60000: the next
 * layer's operands were then not finite; rerun with the 3xTF32 kernels), 0 if not, -1 on error.  Synchronises. */
int nvs_conv_rs_range_flag(int32_t reset);
/* fp32 channels-last (n_pixels, C) <-> split format, C a multiple of 8 (tests, and callers that feed a 3xFP16 conv from
 * their own fp32 data); in and out must be different buffers */
int nvs_split16(const float* in, float* out, int64_t n_pixels, int32_t C, void* stream);
int nvs_unsplit16(const float* in, float* out, int64_t n_pixels, int32_t C, void* stream);
/* debugging aid: CTA 0 of every following flags-bit-4 launch writes 3 x 256 clock64 stamps (epilogue tiles, converter
 * (unused), MMA chunks) into dev_buf (768 int64, device memory); NULL switches it off */
void nvs_conv_rs_debug_buffer(long long* dev_buf);
/* dst_override / dst2_override (may be NULL) replace dst / dst_pool of the plan for this launch. */
int nvs_conv_tc_run(const void* plan, float* dst_override, float* dst2_override, void* stream);

/* 3x3 conv with 1..4 output channels from a channels-last input (score / location heads,
 * heads.py:33): src (B,H,W,cin) NHWC, weight [9][cout][cin], bias [cout], dst (B,cout,H,W) NCHW;
 * act: NVS_ACT_NONE / SIGMOID / TANH. */
int nvs_conv_small(const float* src, const float* weight, const float* bias, float* dst, int32_t B, int32_t H,
                   int32_t W, int32_t cin, int32_t cout, int32_t act, void* stream);

/* depthwise 3x3 + bias (DsConv2d first half, modules/segformer.py:46-54). w: (C,3,3), bias: (C). */
int nvs_dwconv3x3(const float* src, const float* w, const float* bias, float* dst,
                  int32_t B, int32_t C, int32_t H, int32_t W, void* stream);

/* channel LayerNorm (x-mean)/(std+eps)*g+b, biased variance (modules/segformer.py:70-73). */
int nvs_channel_layernorm(const float* src, const float* g, const float* b, float* dst,
                          int32_t B, int32_t C, int32_t HW, float eps, void* stream);

/* softmax over channels, Softmax2d (kp2dtiny.py:942-943). */
int nvs_softmax_channels(const float* src, float* dst, int32_t B, int32_t C, int32_t HW, void* stream);

/* efficient self-attention core (modules/segformer.py:113-133): q (B,C,Nq), kv (B,2C,Nk) with
 * k = kv[:, :C], v = kv[:, C:], heads split the channel axis; out (B,C,Nq) =
 * softmax(q k^T * (C/heads)^-1/2) v per head, streaming softmax (sim is never materialised). */
int nvs_attention(const float* q, const float* kv, float* out, int32_t B, int32_t C, int32_t heads,
                  int32_t Nq, int32_t Nk, void* stream);

/* NetVLAD (modules/aggregators/netvlad.py:79-106 and the loop form :158-193):
 * x (B,C,S) -> vlad (B, K*C).  w_assign (K,C) = conv.weight, centroids (K,C).
 * workspace: nvs_netvlad_workspace_bytes(B,C,K,S) bytes. */
size_t nvs_netvlad_workspace_bytes(int32_t B, int32_t C, int32_t K, int32_t S);
int nvs_netvlad(const float* x, const float* w_assign, const float* centroids, float* vlad,
                void* workspace, size_t workspace_bytes, int32_t B, int32_t C, int32_t K, int32_t S,
                void* stream);

/* Input side of the path (visual_odometry.py:281-291 process_image + frontend.py:79):
 * uint8 HWC frames (B,Hin,Win,3) -> float / 255 -> bilinear resize to (Hout,Wout) when the sizes differ
 * (kornia.geometry.transform.resize = F.interpolate(mode="bilinear", align_corners=False), no antialias)
 * -> (x - 0.5) * 2 -> fp32 NCHW (B,3,Hout,Wout). */
int nvs_preprocess_u8(const uint8_t* img, float* out, int32_t B, int32_t Hin, int32_t Win, int32_t Hout, int32_t Wout,
                      void* stream);

/* GeM over PixelUnshuffle(4) (modules/aggregators/gem.py:21-33, VPRHead method "gem", vpr.py:70-72):
 * x (B,C,H,W) with H, W multiples of 4 -> out (B, 16*C), out[b, c*16 + (y%4)*4 + x%4] =
 * (mean over the 4x4-strided cells of max(x, eps)^p)^(1/p).  No final normalisation (the reference has none). */
int nvs_gem(const float* x, float* out, int32_t B, int32_t C, int32_t H, int32_t W, float p, float eps, void* stream);

/* ConvAP (modules/aggregators/convap.py:29-37, VPRHead method "convap", vpr.py:73-76): 1x1 conv Cin->Cout with bias,
 * AdaptiveAvgPool2d((s1,s2)), flatten channel-major, L2 normalise (eps 1e-12).  weight (Cout,Cin), out (B, Cout*s1*s2).
 * workspace: nvs_convap_workspace_bytes(B,Cin,s1,s2) bytes. */
size_t nvs_convap_workspace_bytes(int32_t B, int32_t Cin, int32_t s1, int32_t s2);
int nvs_convap(const float* x, const float* weight, const float* bias, float* out, void* workspace,
               size_t workspace_bytes, int32_t B, int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t s1,
               int32_t s2, void* stream);
/* L2 normalisation over channels (VPRHead only_encoder path, decoders/vpr.py:85-86). */
int nvs_l2norm_channels(const float* src, float* dst, int32_t B, int32_t C, int32_t HW, void* stream);

/* keypoint decode = KP2DTinyV*.post_processing with training False (kp2dtiny.py:593-647, 959-1015):
 * border mask, cell coordinates, bilinear descriptor sampling (align_corners=True, zero pad) + L2 norm.
 * score (B,1,Hc,Wc), shift (B,2,Hc,Wc) raw tanh, feat (B,D,Hf,Wf)  ->
 * out_score (B,1,Hc,Wc), out_coord (B,2,Hc,Wc) pixels, out_feat (B,D,Hc,Wc) unit norm. */
int nvs_decode(const float* score, const float* shift, const float* feat, float* out_score,
               float* out_coord, float* out_feat, int32_t B, int32_t Hc, int32_t Wc, int32_t D,
               int32_t Hf, int32_t Wf, int32_t H, int32_t W, int32_t cell, float cross_ratio,
               void* stream);

/* seg argmax over classes -> int64 (kp2dtiny.py:639, 1007).  If coord != NULL the class map is first
 * nearest-sampled at the keypoint coordinates (sample_segmentation, kp2dtiny.py:634-637) and the
 * output is (B,1,Hc,Wc); otherwise (B,1,Hs,Ws). */
int nvs_seg_argmax(const float* seg, const float* coord, int64_t* out, int32_t B, int32_t C,
                   int32_t Hs, int32_t Ws, int32_t Hc, int32_t Wc, int32_t H, int32_t W, void* stream);

/* keypoint selection = frontend.py:94-126 per frame: score > thresh (strict), optional class filter,
 * top-k by score (ties -> lowest cell index); outputs compacted in ascending cell order.
 * seg_cells: (B, n_cells) int64 per-cell labels or NULL; filter_classes: device int32[n_filter] or NULL.
 * out_pts (B,k,2) xy, out_desc (B,k,D), out_score (B,k), out_cell (B,k) int32, out_label (B,k) int64
 * (only if seg_cells), out_count (B) int32.  n_cells <= 65536*4. */
int nvs_select_keypoints(const float* score, const float* coord, const float* feat,
                         const int64_t* seg_cells, const int32_t* filter_classes, int32_t n_filter,
                         float thresh, int32_t top_k, float* out_pts, float* out_desc,
                         float* out_score, int32_t* out_cell, int64_t* out_label, int32_t* out_count,
                         int32_t B, int32_t n_cells, int32_t D, void* stream);

/* descriptor matching = BfFeatureMatcher.match (feature_matcher.py:89-98) + goodMatchesOneToOne
 * (:179-209), or mutual nearest neighbour (cv2 crossCheck=True, evaluation/descriptor.py:221).
 * des1 (n1,D) query, des2 (n2,D) train, row-major.
 * mode 0: ratio test + one-to-one (literal reference semantics; needs n2 >= 2);
 * mode 1: mutual NN;  mode 2: raw 2-NN -> out_idx1 = idx (n1,2) int32, out_dist = dist (n1,2) L2 (not
 * squared), out_idx2 / out_count untouched.
 * outputs for modes 0/1 (capacity n1): out_idx1, out_idx2 int32, out_dist float, out_count int32[1].
 * D in {32, 64, 128}.  workspace: nvs_match_workspace_bytes(n1, n2). */
size_t nvs_match_workspace_bytes(int32_t n1, int32_t n2);
int nvs_match(const float* des1, const float* des2, int32_t n1, int32_t n2, int32_t D, double ratio,
              int32_t mode, int32_t* out_idx1, int32_t* out_idx2, float* out_dist, int32_t* out_count,
              void* workspace, size_t workspace_bytes, void* stream);

/* Batched matcher: P pairs of frames in one call.  des (F, kmax, D): the descriptors of all frames as written by
 * nvs_select_keypoints (rows >= counts[f] are ignored); counts (F), pair_a / pair_b (P) are DEVICE arrays, so nothing
 * returns to the host between selection and matching (visual_odometry.py:314-322 matches frame t with t-1;
 * evaluation/descriptor.py:221-229 matches an image with its warped copy).  mode 0 / 1 as nvs_match.  Outputs are
 * (P, kmax) with out_count (P); a pair whose train frame has fewer than 2 keypoints yields no match in mode 0. */
size_t nvs_match_batch_workspace_bytes(int32_t n_pairs, int32_t kmax);
int nvs_match_batch(const float* des, const int32_t* counts, int32_t n_frames, int32_t kmax, int32_t D,
                    const int32_t* pair_a, const int32_t* pair_b, int32_t n_pairs, double ratio, int32_t mode,
                    int32_t* out_idx1, int32_t* out_idx2, float* out_dist, int32_t* out_count, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ---- batched relative pose (visual_odometry.py:383-412: cv2.findEssentialMat + cv2.recoverPose per pair) ----
 * For each of P frame pairs: matched keypoints are unprojected with the pinhole intrinsics ((u-cx)/fx, (v-cy)/fy,
 * camera.unproject_points; pass fx=fy=1, cx=cy=0 for normalised coordinates), `iters` five-point (Nister) hypotheses
 * are drawn with a counter-based generator from `seed`, scored by the truncated squared Sampson distance
 * (inlier iff <= threshold^2, the reference passes 0.0003), and the best E is decomposed; the (R, t) whose
 * triangulated points lie in front of both cameras for most matches is returned (x_b ~ R x_a + t, |t| = 1).
 * Inputs are what nvs_select_keypoints / nvs_match_batch leave on the device: pts (F, kmax, 2); pair_a (current
 * frame) / pair_b (reference frame) (P); idx1 / idx2 (P, kmax) = match m -> keypoint index in frame a / b (both
 * NULL = identity, i.e. pts rows are already matched); count (P) matches per pair.  All DEVICE arrays.
 * Outputs: out_E (P,9) row-major with p_b^T E p_a = 0, out_R (P,9), out_t (P,3), out_mask (P,kmax) uint8,
 * out_inliers (P).  A pair with fewer than 5 matches, or one whose samples are all degenerate
 * (two identical frames: every [t]x fits), yields E = 0, R = I, t = 0 and no inliers.
 * refine = 0: the result is the best minimal-sample model (what cv2.RANSAC returns); refine = n > 0: up to n
 * Gauss-Newton steps on the essential manifold over the consensus set, kept while the truncated cost decreases (the
 * role of local optimisation + final polishing in cv2.USAC_MSAC, the method the reference asks for first), after
 * which mask and inlier count are those of the refined E.
 * Deterministic for a given seed (integer score accumulation).  workspace: nvs_pose_workspace_bytes, 256-aligned. */
size_t nvs_pose_workspace_bytes(int32_t n_pairs, int32_t kmax, int32_t iters);
int nvs_pose_batch(const float* pts, int32_t n_frames, int32_t kmax, const int32_t* pair_a, const int32_t* pair_b,
                   const int32_t* idx1, const int32_t* idx2, const int32_t* count, int32_t n_pairs, float fx,
                   float fy, float cx, float cy, float threshold, int32_t iters, uint64_t seed, int32_t refine,
                   float* out_E, float* out_R, float* out_t, uint8_t* out_mask, int32_t* out_inliers, void* workspace,
                   size_t workspace_bytes, void* stream);
/* The same with a data-dependent sample count (cv2.findEssentialMat's prob = 0.999, visual_odometry.py:392): samples are
 * drawn in rounds of round_size; after each round a pair whose sample count has reached log(1 - confidence) /
 * log(1 - w^5) -- w = the inlier ratio implied by its best truncated cost so far (a lower bound) -- is finished and
 * skipped by the following rounds, up to max_iters samples.  out_iters (P, may be NULL): samples evaluated per pair.
 * Same sample sequence as nvs_pose_batch: a pair that stops after m samples returns what nvs_pose_batch returns with
 * iters = m.  Workspace as nvs_pose_workspace_bytes(n_pairs, kmax, max_iters). */
int nvs_pose_batch_adaptive(const float* pts, int32_t n_frames, int32_t kmax, const int32_t* pair_a,
                            const int32_t* pair_b, const int32_t* idx1, const int32_t* idx2, const int32_t* count,
                            int32_t n_pairs, float fx, float fy, float cx, float cy, float threshold, int32_t max_iters,
                            uint64_t seed, int32_t refine, float confidence, int32_t round_size, float* out_E,
                            float* out_R, float* out_t, uint8_t* out_mask, int32_t* out_inliers, int32_t* out_iters,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---- exact L2 top-k retrieval = faiss.IndexFlatL2.add / .search (evaluation/global_descriptor.py:55-60) ----
 * add:    nvs_flat_prepare converts the fp32 rows (n,d) to fp16 rows padded to nvs_flat_padded_dim(d) columns (the
 *         tcgen05 GEMM operand; one power-of-two scale for the whole shard), stores |x|^2 (fp32) and four shard
 *         statistics (max |x|^2, max |x - fp16(x)|^2, scale exponent, max |component|) that the search turns into a
 *         proven bound of the half-precision error.  Caller owns all four buffers (stats: 4 floats).
 * search: d2 = |q|^2 + |x|^2 - 2 q.x.  A fp16 tcgen05/TMEM GEMM with a fused per-row running list screens the rows;
 *         every row whose screened value is within the error bound of the k-th smallest is re-ranked with the fp32
 *         formula, and a query whose lists could have dropped such a row is answered by an exact fp32 scan instead
 *         (csrc/retrieval.cu header): the result is the fp32 result, independent of scheduling.
 *         out_D (nq,k) ascending, ties -> lower id; out_I (nq,k) int64 = row index + id_offset (id_offset = first
 *         global row of this shard); fewer than k reachable rows (NaN query) -> (+inf, -1).  k <= nvs_flat_max_k().
 *         ev_gemm_start / ev_gemm_stop: optional cudaEvent_t (may be NULL) recorded on `stream` around the
 *         GEMM+list kernel so callers can time it (bench.py roofline).
 * merge:  nvs_topk_merge combines `parts` (<= 16) sorted (nq,k) lists into the global top-k.  Part s of the distances
 *         starts part_stride_d floats after part s-1, part s of the labels part_stride_i int64s (0 = nq*k: the plain
 *         [parts][nq][k] layout); with strides, distances and labels may live in ONE gathered buffer, i.e. one NCCL
 *         allgather per search. */
int32_t nvs_flat_padded_dim(int32_t d);
int32_t nvs_flat_max_k(void);
/* host-only: number of per-query list slots (= most clusters that can share one group of query blocks) and, through
 * *cluster_size (may be NULL), the CTAs per cluster a search of this shape uses; 0 for invalid arguments */
int32_t nvs_flat_list_slots(int64_t n_db, int32_t nq, int32_t d, int32_t k, int32_t* cluster_size);
/* debugging aid (tools/retr_waits.py): every following GEMM launch writes 16 int64 per CTA into dev_buf (device memory,
 * 16 x 148 entries) -- cycles of the MMA role, producer wait, MMA wait for operands, MMA wait for a drained accumulator,
 * epilogue wait, epilogue cycles, k-blocks issued, unused; NULL switches it off */
void nvs_flat_debug_buffer(long long* dev_buf);
int nvs_flat_prepare(const float* x, int64_t n, int32_t d, void* x_f16, float* norms, float* stats, void* stream);
size_t nvs_flat_search_workspace_bytes(int64_t n_db, int32_t nq, int32_t d, int32_t k);
int nvs_flat_search(const float* db, const void* db_f16, const float* db_norms, const float* db_stats, int64_t n_db,
                    const float* q, int32_t nq, int32_t d, int32_t k, int64_t id_offset, float* out_D, int64_t* out_I,
                    void* workspace, size_t workspace_bytes, void* ev_gemm_start, void* ev_gemm_stop,
                    void* stream);
/* The same search in two calls, for a database sharded over several GPUs (queries replicated): _begin runs the GEMM
 * and publishes, per query, out_bounds[q][0..k) = upper bounds of the EXACT distances of this shard's k best listed rows
 * (unordered, +inf padding); the caller gathers them from all shards (one allgather of nq*k floats per shard) and
 * nvs_flat_bound_merge takes the k-th smallest of the union per query: an upper bound of the GLOBAL k-th distance.
 * _end then re-ranks only the rows that can still be among the global k nearest: the fp32 re-rank (a gather of whole
 * rows) shrinks with the number of shards like the GEMM does.  A shard that holds fewer than k such rows pads its list
 * with (+inf, -1); nvs_topk_merge of the shard lists is the exact global result.  Same workspace for _begin and _end
 * (it carries the per-row lists), same stream order. */
int nvs_flat_search_begin(const float* db, const void* db_f16, const float* db_norms, const float* db_stats, int64_t n_db,
                          const float* q, int32_t nq, int32_t d, int32_t k, float* out_bounds, void* workspace,
                          size_t workspace_bytes, void* ev_gemm_start, void* ev_gemm_stop, void* stream);
int nvs_flat_bound_merge(const float* bounds /* (parts, nq, k) */, int32_t parts, int32_t nq, int32_t k,
                         float* out_bound /* (nq) */, void* stream);
int nvs_flat_search_end(const float* db, const void* db_f16, const float* db_norms, const float* db_stats, int64_t n_db,
                        const float* q, int32_t nq, int32_t d, int32_t k, int64_t id_offset, const float* global_bound,
                        float* out_D, int64_t* out_I, void* workspace, size_t workspace_bytes, void* stream);
int nvs_topk_merge(const float* D_parts, const int64_t* I_parts, int32_t parts, int32_t nq, int32_t k,
                   int64_t part_stride_d, int64_t part_stride_i, float* out_D, int64_t* out_I, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NANOVS_H_ */
