/*
 * tdrn_b200.h -- C ABI of libtdrn_b200.so: the B200 (sm_100a) implementation of the TDRN /
 * DualRefineDet inference hot path.
 *
 * Every entry point is extern "C", takes plain pointers and sizes (no torch / THC types), returns
 * 0 on success or a negative TDRN_E* code, never allocates device memory (the caller owns every
 * buffer, including scratch), launches on the stream it is given and does not synchronise unless
 * the comment says so.  tdrn_last_error() returns the thread-local message of the last failure.
 *
 * Each declaration cites the reference interface it replaces (paths relative to the upstream
 * SeanChenxy/TDRN checkout).  INTEGRATION.md shows the reference-side bindings.
 *
 * Layout conventions: "NCHW" tensors are the reference's; "NHWC" are the library's internal
 * activations.  Pointers are DEVICE pointers unless the parameter name ends in _host.
 */
#ifndef TDRN_B200_H_
#define TDRN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDRN_OK            0
#define TDRN_EINVAL       -1   /* bad argument / shape check failed (reference: THArgCheck -> RuntimeError) */
#define TDRN_ECUDA        -2   /* CUDA runtime / launch error (reference only printf'd these) */
#define TDRN_EWORKSPACE   -3   /* caller-provided workspace too small */
#define TDRN_EUNSUPPORTED -4   /* configuration not implemented by this kernel */

#define TDRN_F32  0
#define TDRN_BF16 1
#define TDRN_BF16_SPLIT 2   /* out_dtype of tdrn_conv_first only: [.., 2*C] bf16, C high parts then C low parts (x = hi + lo to 16
                               mantissa bits) -- the operand format of tdrn_conv2d_tc's split3 mode */
#define TDRN_F16  3          /* IEEE half activations / weights (same tensor-core rate as bf16, 11 significand bits instead of 8): accepted by
                               tdrn_conv_first (out), tdrn_dwconv3x3_io, tdrn_conv2d_tc (in_dtype: weights packed as half too; out_dtype) and
                               tdrn_l2norm_io -- the MobileNet trunks (model/dualrefinedet_mobilenet.py:139-152), whose 27 stacked layers
                               exceed the 2e-2 bar with bf16 storage.  Conversions to half saturate at +-65504. */

typedef void *tdrn_stream_t;   /* cudaStream_t */

const char *tdrn_last_error(void);
int tdrn_version(void);
/* Number of kernels this library has launched in the calling process (bench.py "gpu_launches"). */
long long tdrn_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * A5  PriorBox.forward        layers/functions/prior_box.py:33-64  (host, float64 -> fp32)
 * ars is [n_levels][4] (unused entries ignored), n_ar[k] = len(aspect_ratios[k]); sizes and aspect ratios are doubles:
 * the reference takes any Python number (SSD-512's min_sizes are 35.84, 76.8, ...; MOT_300's aspect ratios 0.5, 1/3, 0.25);
 * max_sizes_host may be NULL (cfg['max_sizes'] == []).  out_host may be NULL to query *num_priors.
 * ------------------------------------------------------------------------------------------ */
int tdrn_prior_box(int image_size, int n_levels, const int *feature_maps_host, const int *steps_host,
                   const double *min_sizes_host, const double *max_sizes_host, const int *n_ar_host,
                   const double *ars_host, int flip, int clip, float *out_host, int *num_priors);

/* ------------------------------------------------------------------------------------------
 * A3  deform_conv_forward_cuda   utils/deformconv/deform_conv_cuda.h:1-7, deform_conv_cuda.c:98-213
 * Same argument meaning as the reference (NCHW fp32 input/offset/output, weight [Cout,Cin,kH,kW],
 * no bias) minus the caller-owned `columns`/`ones` scratch, which no longer exists: the sampled
 * columns never leave the SM.  fp32 exact path (SIMT); shape errors return TDRN_EINVAL.
 * ------------------------------------------------------------------------------------------ */
int tdrn_deform_conv_forward(const float *input, const float *weight, const float *offset, float *output,
                             int B, int Cin, int H, int W, int Cout, int kW, int kH, int dW, int dH,
                             int padW, int padH, int dilationH, int dilationW, int deformable_group,
                             tdrn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A8  NMS   utils/nms_wrapper.py:23-31 -> utils/nms/cpu_nms.pyx:17-68 (the rule Detect uses);
 *           replaces `_nms` utils/nms/gpu_nms.hpp:1-2 (which used `>` and needed pre-sorted input).
 * dets [n,5] = (x1,y1,x2,y2,score) fp32, any order.  Sorts (descending score, ties -> lower
 * index), suppresses j when IoU(+1 convention) >= thresh (compared in double like the Cython
 * code), stops after max_keep boxes are kept (max_keep <= 0: no limit).  keep [<= n] int32 indices
 * into dets in score order, *num_keep on device.  workspace >= tdrn_nms_workspace_bytes(n).
 * tdrn_nms_host: host pointers, synchronous, allocates its own scratch (drop-in for `_nms`).
 * ------------------------------------------------------------------------------------------ */
size_t tdrn_nms_workspace_bytes(int n);
int tdrn_nms(const float *dets, int n, double thresh, int max_keep, int *keep, int *num_keep,
             void *workspace, size_t workspace_bytes, tdrn_stream_t stream);
int tdrn_nms_host(int *keep_out_host, int *num_out_host, const float *dets_host, int boxes_num,
                  int boxes_dim, double thresh, int device_id);
/* The same two entry points with the suppression rule explicit.  rule 0: `ovr >= thresh`, threshold compared as a double
   (utils/nms/cpu_nms.pyx:65: what Detect and nms(..., force_cpu=True) run); rule 1: `ovr > (float)thresh` (the reference's
   GPU kernel, utils/nms/nms_kernel.cu:71: what nms(dets, thresh) runs with the default force_cpu=False).  The two differ
   only for a pair whose IoU equals the threshold exactly. */
int tdrn_nms_rule(const float *dets, int n, double thresh, int rule, int max_keep, int *keep, int *num_keep,
                  void *workspace, size_t workspace_bytes, tdrn_stream_t stream);
int tdrn_nms_host_rule(int *keep_out_host, int *num_out_host, const float *dets_host, int boxes_num,
                       int boxes_dim, double thresh, int device_id, int rule);

/* ------------------------------------------------------------------------------------------
 * A6  decode / center_size   layers/box_utils.py:176-195, :16-25 as applied by
 *     layers/functions/detection.py:43-48.  loc [B,P,4], priors [P,4] (cx,cy,w,h),
 *     arm_loc [B,P,4] or NULL -> boxes [B,P,4] (x1,y1,x2,y2 normalised).
 * ------------------------------------------------------------------------------------------ */
int tdrn_decode(const float *loc, const float *priors, const float *arm_loc, int B, int P, float *boxes,
                tdrn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A8  Detect.forward   layers/functions/detection.py:25-70
 * loc [B,P,4], conf [B*P,C] (softmax scores), priors [P,4], arm_loc [B,P,4] or NULL,
 * scale_host[4] -> out [B,C,top_k,5] = (score,x1,y1,x2,y2), class 0 rows zero.
 * Candidate test `score > conf_thresh` in fp32; NMS as tdrn_nms with max_keep = top_k.
 * ------------------------------------------------------------------------------------------ */
size_t tdrn_detect_workspace_bytes(int B, int P, int C, int top_k);
int tdrn_detect(const float *loc, const float *conf, const float *priors, const float *arm_loc,
                const float *scale_host, int B, int P, int C, int top_k, float conf_thresh,
                double nms_thresh, float *out, void *workspace, size_t workspace_bytes,
                tdrn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A1/A2/A2m/A9 layer operators (replace the cuDNN calls behind nn.Conv2d / ConvTranspose2d /
 * BatchNorm2d / MaxPool2d / Softmax in model/networks.py:136-163,736-745,
 * model/dualrefinedet_vggbn.py:119-206, and L2Norm layers/modules/l2norm.py:17-21).
 * Activations are NHWC in `dtype` (TDRN_F32 or TDRN_BF16); accumulation is always fp32.
 * ------------------------------------------------------------------------------------------ */
typedef struct tdrn_conv_desc {
    int B, H, W, Cin;          /* input  [B,H,W,Cin] NHWC                                          */
    int Cout, kh, kw;          /* weight packed [kh*kw*Cin][Cout] (tap-major, then cin), BN folded  */
    int stride, pad, dil;
    int relu;                  /* apply max(0,.) after bias (+ residual)                            */
    int deconv2x2;             /* 1: ConvTranspose2d k2 s2 (weight packed [Cin][4*Cout], n=(ij,co)) */
    int dg;                    /* >0: deformable conv with dg offset groups (offsets NHWC fp32
                                  [B,Ho,Wo,dg*2*kh*kw]); sampler = deform_conv_cuda_kernel.cu:16-51 */
    int in_dtype, out_dtype;   /* TDRN_F32 / TDRN_BF16 (tdrn_conv2d_tc also: TDRN_F16, see above)    */
    /* output addressing (elements): out[b*out_sb + (y*Wo+x)*out_sp + co]; lets heads write straight
       into the NHWC-flattened [B,P,4] / [B,P,C] tensors the reference builds with permute+cat.    */
    long long out_sb, out_sp;
    long long in_sb;           /* input batch stride in elements; 0 = H*W*Cin (lets the 1x1 offset conv read
                                  the ARM regression straight out of the flattened [B,P,4] tensor)        */
    int pool2x2;               /* 1: fuse the following MaxPool2d(2,2) (vgg() 'M'/'C', networks.py:141-143) into
                                  the epilogue; output is [B,Ho/2,Wo/2,Cout]. tdrn_conv2d_tc only, needs
                                  Wo % 16 == 0 and Ho % 8 == 0                                             */
    int split3;                /* 1 (tdrn_conv2d_tc only): fp32-accurate tensor-core mode.  `in` is the SPLIT tensor
                                  [B,H,W,2*Cin] bf16 written by tdrn_split_bf16 (per pixel: Cin high parts, then Cin low
                                  parts, x = hi + lo to 16 mantissa bits); `weight` is packed [Cout_pad][kh*kw][2*Cin] =
                                  (W_hi | W_lo) per tap.  The kernel accumulates hi*W_hi + hi*W_lo + lo*W_hi in fp32
                                  (the reference's convs are fp32, model/networks.py:136-163; the dropped lo*lo term is
                                  2^-18 relative).  Needs Cin % 64 == 0.                                      */
    int split_out;             /* g > 0 (split3, bf16 out, out_sp == 2*Cout, g % 16 == 0, Cout % g == 0): the fp32 result is
                                  written as (hi | lo) bf16 pairs in groups of g channels -- output channel n -> hi at
                                  n + (n / g) * g, lo g further -- the projection tensor of the fp32-accurate deformable
                                  heads (tdrn_deform_head_desc.split)                                           */
} tdrn_conv_desc;

/* fp32-accurate SIMT implicit GEMM (also bf16 in/out with fp32 accumulate). bias/residual may be NULL;
   residual has the output's addressing and dtype (pass residual == out to accumulate, as the
   multihead `l(ob,f) + l2(ob,f2)` does).  offsets must be non-NULL iff d->dg > 0. */
int tdrn_conv2d(const tdrn_conv_desc *d, const void *in, const float *weight_f32, const float *bias,
                const void *residual, const float *offsets, void *out, tdrn_stream_t stream);

/* tcgen05/TMEM/TMA implicit GEMM, bf16 in, fp32 accumulate, bf16 or fp32 out.
   weight_bf16 packed [Cout_pad][kh*kw*Cin_pad] K-major (Cout_pad = Cout rounded up to 16, Cin_pad = Cin rounded up
   to 64 with zero weights for the padding channels).
   Requires Cin % 8 == 0 and stride 1 or 2.  Returns TDRN_EUNSUPPORTED otherwise (caller picks tdrn_conv2d). */
int tdrn_conv2d_tc(const tdrn_conv_desc *d, const void *in_bf16, const void *weight_bf16, const float *bias,
                   const void *residual, void *out, tdrn_stream_t stream);

/* Depthwise 3x3 (+folded BN bias, ReLU): conv_dw first half, model/networks.py:738-740.
   weight packed [9][C] fp32. */
int tdrn_dwconv3x3(const void *in, const float *weight, const float *bias, void *out, int B, int H, int W,
                   int C, int stride, int relu, int dtype, tdrn_stream_t stream);
/* Same with separate input / output formats (TDRN_BF16 or TDRN_F16 each; C % 8 == 0, stride 1 or 2). */
int tdrn_dwconv3x3_io(const void *in, const float *weight, const float *bias, void *out, int B, int H, int W,
                      int C, int stride, int relu, int in_dtype, int out_dtype, tdrn_stream_t stream);

/* One conv_dw block of the MobileNet trunks (model/networks.py:736-745: Conv2d(inp, inp, 3, stride, 1, groups=inp) + BN + ReLU,
   Conv2d(inp, oup, 1) + BN + ReLU; used by dualrefinedet_mobilenet.py:23-35 and ssd4scale_mobile.py) in one kernel: the
   depthwise result is produced on the CUDA cores straight into the shared-memory operand of the pointwise tcgen05 GEMM and
   never reaches HBM.  bf16 NHWC in [B,H,W,Cin] / out [B,Ho,Wo,Cout]; dw_weight [9][Cin] fp32 and dw_bias [Cin] (BN folded,
   as for tdrn_dwconv3x3); pw_weight bf16 [Cout_pad16][Cin_pad64] K-major and pw_bias [Cout] (as for tdrn_conv2d_tc).
   Bit-identical to tdrn_dwconv3x3 followed by tdrn_conv2d_tc.  Needs Cin % 8 == 0, Cout % 8 == 0, stride 1 or 2. */
typedef struct {
    int B, H, W, Cin, Cout;
    int stride;                /* of the depthwise conv (pad 1) */
    int relu_dw, relu_pw;
} tdrn_dwpw_desc;
int tdrn_conv_dwpw(const tdrn_dwpw_desc *d, const void *in_bf16, const float *dw_weight, const float *dw_bias,
                   const void *pw_weight_bf16, const float *pw_bias, void *out_bf16, tdrn_stream_t stream);

/* First conv (Cin=3) reading the reference's NCHW fp32 image directly, writing NHWC `out_dtype`.
   weight packed [27][Cout] fp32 (tap-major, then cin). stride 1 (VGG conv1_1) or 2 (MobileNet).
   out_dtype TDRN_BF16_SPLIT (Cout = 64, stride 1, maps that tile as 64x2 / 32x4 / 16x8; TDRN_EUNSUPPORTED otherwise): the layer
   runs on the tensor cores in split precision (x and w as hi + lo bf16 pairs, three products) and writes [B,H,W,128]. */
int tdrn_conv_first(const float *x_nchw, const float *weight, const float *bias, void *out, int B, int H,
                    int W, int Cout, int stride, int relu, int out_dtype, tdrn_stream_t stream);

/* MaxPool2d(2,2, ceil_mode) NHWC. */
int tdrn_maxpool2x2(const void *in, void *out, int B, int H, int W, int C, int ceil_mode, int dtype,
                    tdrn_stream_t stream);

/* L2Norm: out = weight[c] * x / (sqrt(sum_c x^2) + 1e-10), NHWC. */
int tdrn_l2norm(const void *in, const float *weight, void *out, long long pixels, int C, int dtype,
                tdrn_stream_t stream);
/* Same with separate 16-bit input / output formats (TDRN_BF16 or TDRN_F16 each; C in {256, 512, 1024}, pixels % 4 == 0): the
   L2Norm that hands a half-precision MobileNet source to the bf16 ARM heads / TCB. */
int tdrn_l2norm_io(const void *in, const float *weight, void *out, long long pixels, int C, int in_dtype, int out_dtype,
                   tdrn_stream_t stream);

/* VGG conv1_1 (3 -> 64, reads the reference's NCHW fp32 image) and conv1_2 (64 -> 64) + optional MaxPool2d(2,2) in one
   kernel (model/networks.py:136-163, cfg entries 64, 64, 'M'; folded BN + ReLU after each conv): conv1_1's output -- the
   largest activation of the network -- never reaches HBM.  w1 [27][64] fp32 (k = (i*3+j)*3 + c), w2 bf16 [64][9*64] K-major
   (k = tap*64 + c, as for tdrn_conv2d_tc), out bf16 NHWC [B,H/2,W/2,64] (pool) or [B,H,W,64].  Needs W % 8 == 0 and
   H % 16 == 0 (TDRN_EUNSUPPORTED otherwise: the caller uses tdrn_conv_first + tdrn_conv2d_tc, same results bit for bit). */
int tdrn_conv_stem_pair(const float *x, const float *w1, const float *b1, const void *w2, const float *b2, void *out,
                        int B, int H, int W, int relu1, int relu2, int pool, tdrn_stream_t stream);

/* L2Norm and the MaxPool2d(2,2) of the SAME input in one pass (conv4_3 / conv5_3 feed both, model/
   dualrefinedet_vggbn.py:130-148): out_norm [B,H,W,C], out_pool [B,H/2,W/2,C].  bf16, C in {256,512,1024}, even H and W;
   TDRN_EUNSUPPORTED otherwise (caller uses tdrn_l2norm + tdrn_maxpool2x2). */
int tdrn_l2norm_pool2x2(const void *in, const float *weight, void *out_norm, void *out_pool, int B, int H, int W,
                        int C, int dtype, tdrn_stream_t stream);

/* Row softmax over C classes, fp32 [rows, C] (nn.Softmax(dim=1), dualrefinedet_vggbn.py:196). In place ok. */
int tdrn_softmax(const float *in, float *out, long long rows, int C, tdrn_stream_t stream);

/* NHWC (dtype) -> NCHW fp32 (to hand offset maps back in the reference layout) and the reverse. */
int tdrn_nhwc_to_nchw_f32(const void *in, float *out, int B, int H, int W, int C, int dtype, tdrn_stream_t stream);

/* The `offset.k` / `offset2.k` 1x1 convs (12 -> 18*dg and 12 -> 50*dg) of all pyramid levels in one launch
   (model/dualrefinedet_vggbn.py:160-164, model/ssd4scale_vgg.py:72-76: `self.offset[k](arm_loc_k)`): arm_loc is the flattened
   ARM regression [B,P,4] (level k: prior_off + (y*W+x)*3 + a), w1 [c1][12] / w2 [c2][12] fp32 row-major (the reference's
   [Cout,12,1,1] weights as they lie), b1 / b2 may be NULL; out1 [B,H,W,c1], out2 [B,H,W,c2] NHWC fp32 (what the deformable
   heads read), out1_nchw [B,c1,H,W] or NULL (the maps the reference's forward returns).  c2 == 0: no second head. */
#define TDRN_MAX_OFFSET_LEVELS 6
typedef struct tdrn_offset_level {
    int H, W, prior_off;
    const float *w1, *b1, *w2, *b2;
    float *out1, *out2, *out1_nchw;
} tdrn_offset_level;
int tdrn_offset_convs(const float *arm_loc, int B, int P, int n_levels, const tdrn_offset_level *levels, int c1, int c2,
                      tdrn_stream_t stream);

/* fp32 NHWC activations [pixels][C] -> the split operand of the fp32-accurate tensor-core convs (tdrn_conv_desc.split3):
   out [pixels][2*C] bf16, out[p][c] = bf16(x), out[p][C + c] = bf16(x - float(out[p][c])).  C % 8 == 0. */
int tdrn_split_bf16(const float *in, void *out, long long pixels, int C, tdrn_stream_t stream);
int tdrn_nchw_f32_to_nhwc(const float *in, void *out, int B, int C, int H, int W, int dtype, tdrn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused deformable ODM / TDRN head (replaces the per-sample im2col + SGEMM loop,
 * deform_conv_cuda.c:157-193, for loc and conf together; model/dualrefinedet_vggbn.py:181-197).
 * bf16 tcgen05 kernel: the bilinear-sampled im2col tile is built in shared memory and consumed by
 * tcgen05.mma in place (the fp32-exact equivalent is tdrn_conv2d with dg > 0).
 *   feat    [B,H,W,Cin] NHWC bf16                    offsets  [B,H,W,dg*2*kh*kw] NHWC fp32
 *   weight  loc||conf rows concatenated, N = 12 + 3*C: [N_pad][kh*kw*Cin] bf16 K-major (N_pad = N up to 16)
 *   Optional second head (multihead 5x5): feat shared, offsets2/weight2 with kh2=kw2=5, pad 2;
 *   its result is added to the first (l(ob,f)+l2(ob,f2)).
 *   loc_out  [B,P,4] fp32 and conf_out [B,P,C] fp32 are written at prior offset `prior_off`
 *   (prior index = prior_off + (y*W+x)*3 + a); if softmax != 0 conf_out holds softmax over C.
 * ------------------------------------------------------------------------------------------ */
typedef struct tdrn_deform_head_desc {
    int B, H, W, Cin, num_classes, dg;
    int kh, pad;               /* first head (square kernel, stride 1, dilation 1) */
    int kh2, pad2;             /* second head, kh2 == 0: absent                    */
    int P, prior_off;
    int softmax;
    int split;                 /* tdrn_deform_head_sample only: 1 = the projections are (hi | lo) bf16 pairs,
                                  proj [B,H,W,taps,2*g] with n_pad = 2*g, 12 + 3C <= g (fp32-accurate heads: both halves
                                  are sampled and added in fp32 before the softmax);
                                  2 = the projections are IEEE half (TDRN_F16 output of tdrn_conv2d_tc) instead of bf16:
                                  same bytes, 11 instead of 8 significand bits in the sampled values                */
} tdrn_deform_head_desc;

int tdrn_deform_head(const tdrn_deform_head_desc *d, const void *feat, const float *offsets,
                     const void *weight, const float *offsets2, const void *weight2,
                     float *loc_out, float *conf_out, tdrn_stream_t stream);

/* "Project, then sample" form of the same head for narrow heads with one deformable group (12 + 3C much smaller
 * than Cin): bilinear sampling commutes with the 1x1 channel contraction, so the caller first computes the per-tap
 * projections with one dense GEMM (tdrn_conv2d_tc, 1x1, Cout = (kh*kh + kh2*kh2) * n_pad, bf16 out)
 *   proj [B,H,W,taps,n_pad] bf16,  proj[b,y,x,t,o] = sum_c W_t[o,c] * feat[b,y,x,c]
 * (t runs over the taps of head 1 then head 2, o over loc rows then conf rows, zero-padded to n_pad % 8 == 0),
 * and this entry point samples them with the reference's sampler (deform_conv_cuda_kernel.cu:16-51,195-203),
 * sums over taps, applies the softmax and writes loc_out / conf_out exactly like tdrn_deform_head.
 * Replaces the same reference calls as tdrn_deform_head (deform_conv_cuda.c:98-213 for loc and conf). */
int tdrn_deform_head_sample(const tdrn_deform_head_desc *d, const void *proj, int n_pad,
                            const float *offsets, const float *offsets2, float *loc_out, float *conf_out,
                            tdrn_stream_t stream);
/* The same for all pyramid levels of a detector in one launch (`for k in range(4): odm_loc[k](...), odm_conf[k](...)`,
   dualrefinedet_vggbn.py:180-189): descs[k] / projs[k] / offsets[k] / offsets2[k] (offsets2 may be NULL) describe level k,
   n_pad, loc_out [B,P,4] and conf_out [B,P,C] are shared (descs[k].prior_off places the level).  n_levels <= 6. */
int tdrn_deform_head_sample_group(int n_levels, const tdrn_deform_head_desc *descs, const void *const *projs, int n_pad,
                                  const float *const *offsets, const float *const *offsets2, float *loc_out, float *conf_out,
                                  tdrn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (next row) Result scatter of the evaluation drivers: evaluate.py:469-483, evaluate_coco.py:140-159
 * (per image and class: masked_select(score > 0), boxes * (w,h,w,h), .cpu()).  One call for the whole batch:
 *   det [B,C,top_k,5] (Detect output), wh [B,2] = original (width, height) of every image (device, fp32)
 *   -> out_rows [n,7] = (image, class, x1*w, y1*h, x2*w, y2*h, score), ordered by image, class, rank;
 *      *count = n on the device (rows beyond max_rows are counted but not written).  Class 0 is skipped.
 * workspace >= tdrn_collect_workspace_bytes(B, C).
 * ------------------------------------------------------------------------------------------ */
size_t tdrn_collect_workspace_bytes(int B, int C);
int tdrn_collect_detections(const float *det, const float *wh, int B, int C, int top_k, float *out_rows, int max_rows,
                            int *count, void *workspace, size_t workspace_bytes, tdrn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (next row) Pre-processing of the evaluation / video drivers on the device:
 *   base_transform (data/__init__.py:7-12): cv2.resize(image, (size, size)) [INTER_LINEAR, 8-bit fixed point],
 *   .astype(float32), -= mean; then the callers' `img[:, :, (2,1,0)]` (data/voc0712.py:466-467; swap_rb = 1) and
 *   HWC -> CHW (`permute(2,0,1)`, voc0712.py:468; test_video_trn.py:91 has no channel swap: swap_rb = 0).
 *   frames [B,Hs,Ws,3] uint8 (device, cv2 channel order), mean3 = host float[3] in the order of the SOURCE channels
 *   -> out [B,3,size,size] fp32 NCHW, the `x` that net(x) takes.
 * ------------------------------------------------------------------------------------------ */
int tdrn_preprocess(const unsigned char *frames, int B, int Hs, int Ws, int size, const float *mean3, int swap_rb,
                    int flip_lr /* cv2.flip(image, 1) first: multi_eval.py:541-544 */, float *out, tdrn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (next row) Merge step of multi-scale / flip testing: multi_eval.py:557-640 (gather the detections of the
 * K passes per class: score > 0, un-flip, scale to pixels, per-scale size rule) + bbox_vote (multi_eval.py:453-494).
 *   dets [K,C,top_k,5] = Detect outputs of the K passes of ONE image (device); flip[k] != 0: pass k saw the mirrored image;
 *   rule[k] 0: keep boxes whose longer side (+1 convention) > rule_thr[k], 1: whose shorter side < rule_thr[k];
 *   (w, h) original image size; vote_thresh 0.45f
 *   -> out [C,max_out,5] rows (x1,y1,x2,y2,score) in pixels, in voting order; out_count[C] (rows beyond max_out are counted,
 *      not written; K*top_k always suffices).  Class 0 is skipped.  NumPy float32 arithmetic order (oracle/multi_scale_ref.py).
 * ------------------------------------------------------------------------------------------ */
size_t tdrn_multiscale_vote_workspace_bytes(int K, int C, int top_k);
int tdrn_multiscale_vote(const float *dets, const int *flip, const int *rule, const float *rule_thr, int K, int C, int top_k,
                         float w, float h, float vote_thresh, float *out, int *out_count, int max_out, void *workspace,
                         size_t workspace_bytes, tdrn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TDRN_B200_H_ */
