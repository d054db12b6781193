/*
 * dbnet_b200.h — C-ABI of the B200-native DBNet++ detection hot path.
 *
 * The reference (ZenHKD/ocr-vi-invoice) is pure Python and has no FFI of its own; every entry
 * point below replaces one torch / torchvision / cv2 call site of the reference's detection path
 * (file:line relative to the reference tree).  Plain pointers and sizes only, no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; b200_last_error() gives the text.
 *     Nothing throws or aborts (reference only has try/except around unclip, src/det/test.py:87-95).
 *   - device pointers are owned by the caller; the library allocates nothing on the device.
 *   - `stream` is a cudaStream_t passed as void*.  Entry points are re-entrant per stream.
 *   - activations are NHWC bf16 unless stated; fp32 maps are NCHW as in the reference's dict.
 */
#ifndef DBNET_B200_H
#define DBNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 3

/* ---- status ------------------------------------------------------------------------------- */
int         b200_abi_version(void);
const char* b200_last_error(void);
/* 0 when device `dev` is an sm_100 part this library can run on. */
int         b200_device_check(int dev);

/* ---- implicit-GEMM convolution family (tcgen05 + TMEM + TMA) --------------------------------
 * Replaces: nn.Conv2d + BatchNorm2d + ReLU chains of torchvision Bottleneck (resnet.py:143-163 via
 * model/det/backbone.py:34-37), ConvBnRelu (model/det/layers.py:13-18), the lateral / smoothing
 * convs (model/det/neck.py:16,21,31-39), offset_mask_conv (model/det/dcn.py:17-25,42),
 * ConvTranspose2d 2x2 s2 (model/det/head.py:13,21) and, in B200_IGEMM_DCN mode,
 * torchvision.ops.deform_conv2d (model/det/dcn.py:48-57).
 */
enum {
    B200_IGEMM_CONV = 0, /* 1x1 / 3x3, stride 1 or 2, zero padding (k-1)/2; A operand by TMA     */
    B200_IGEMM_STEM = 1, /* 7x7 s2 p3 conv on the space-to-depth buffer of b200_stem_pack; A rows
                            (128 contiguous bytes each) copied by producer warps                 */
    B200_IGEMM_DCN  = 2, /* modulated deformable conv v2, 3x3; A operand gathered + blended      */
    B200_IGEMM_DCN_TMEM = 3 /* the same operator for layers with out_c == n_pad in {128, 256} (ResNet-50 c3 / c4):
                            sampled A operand written to tensor memory (tcgen05.st) and read by tcgen05.mma
                            from there, sampling records staged by TMA.  The weights must be packed with the
                            K permutation of ops.pack_dcn_weight (inside each 64-channel chunk, K position
                            16j + 4m + r holds channel 32(j>>1) + 8m + 4(j&1) + r); flags: RELU only       */
};

enum {
    B200_EPI_RELU       = 1,  /* max(v, 0)                                                        */
    B200_EPI_RES        = 2,  /* v += residual[same pixel]               (Bottleneck identity)   */
    B200_EPI_RES_UP2    = 4,  /* v += residual[(y>>1, x>>1)]             (neck.py:36-38 nearest) */
    B200_EPI_OUT_F32    = 8,  /* store fp32 instead of bf16              (offset/mask conv)      */
    B200_EPI_SHUFFLE2   = 16, /* N = 4*shuf_c, column (dy*2+dx)*shuf_c+co goes to pixel
                                 (2y+dy, 2x+dx), channel co               (ConvTranspose2d 2x2) */
    B200_EPI_HEAD_DOT   = 64, /* DB head: with SHUFFLE2, shuf_c = 64, out_c = 256.  The 64-channel map of the first
                               * ConvTranspose2d is not stored; BN + ReLU + bf16 rounding + the LAST
                               * ConvTranspose2d(64 -> 1, k2 s2) (head.py:15,23) run in the epilogue and y receives
                               * the fp32 logit map (batch, 4*in_h, 4*in_w).  `offsets` = float[257]: tail weights
                               * w[c*4 + dy*2+dx], then the bias.  The BN shift is added by an extra UMMA step and the
                               * tail is a second UMMA (weights rounded to bf16): `scale` must be all ones -- fold the BN
                               * scale into the weights.                                                      */
    B200_EPI_GELU       = 128, /* v = GELU(v) (erf form) BEFORE the residual add: the recogniser's Linear / conv +
                                  GELU layers (model/rec2/svtrv2.py:33,51,54,117); bf16 outputs, out_c % 64 == 0 */
    B200_EPI_CHAIN      = 256, /* a SECOND 1x1 GEMM chained onto the epilogue (torchvision Bottleneck: conv3 + bn3 + identity + ReLU
                                  of one block, resnet.py:157-163, followed by conv1 + bn1 + ReLU of the NEXT block, :146-148):
                                  y = the normal output (bf16, 256 channels, with RES + RELU as flagged); every staged 64-channel
                                  chunk of y is also the A operand of y2[pixel][chain_n] += chunk x chain_w^T, and after the tile's
                                  last chunk y2 = relu?(y2 * chain_scale + chain_shift) is stored as bf16: the block output is never
                                  re-read.  CONV mode, 1x1, out_c == n_pad == 256, chain_n 64 or 128 (residual optional).    */
    B200_EPI_DCN_COORDS = 32  /* offset_mask_conv only (out_c = 32): instead of the raw 27 channels, store
                                 the b200_dcn_coord records of the 9 taps, y = [pixel][9] records, so
                                 the DCN producers do no coordinate arithmetic (dcn.py:44-46 fused)    */
};

/* Sampling record of one (output pixel, tap) of the modulated deformable conv (torchvision deform_conv2d
 * semantics): `off` = element offset ((row*W + col)*x_ld) of the top-left pixel of a 2x2 block clamped inside
 * the image; w[4] = bf16 bilinear weights of the block's pixels (0,0) (0,1) (1,0) (1,1), already multiplied by
 * sigmoid(mask) and zero for corners torchvision treats as outside. */
typedef struct b200_dcn_coord {
    int32_t  off;
    uint16_t w[4];
    uint32_t pad;
} b200_dcn_coord;

typedef struct b200_igemm_desc {
    int32_t mode;          /* B200_IGEMM_*                                                        */
    int32_t batch;
    int32_t in_h, in_w;    /* input spatial size (STEM: the ORIGINAL image H, W)                  */
    int32_t in_c;          /* input channels, multiple of 64 (STEM: ignored)                      */
    int32_t out_c;         /* real output channels (GEMM N); weights are padded to n_pad rows     */
    int32_t n_pad;         /* rows of the weight matrix, multiple of block_n                      */
    int32_t ksize;         /* 1 or 3 (STEM: 7)                                                    */
    int32_t stride;        /* 1 or 2                                                              */
    int32_t flags;         /* B200_EPI_*                                                          */
    int32_t out_ld;        /* elements between consecutive output pixels                          */
    int32_t res_ld;        /* elements between consecutive residual pixels                        */
    int32_t shuf_c;        /* channels per sub-pixel for B200_EPI_SHUFFLE2                        */
    int32_t block_n;       /* 32, 64, 128 or 256; 0 = choose                                      */
    int32_t x_ld;          /* elements between consecutive input pixels; 0 = in_c (dense NHWC)    */
    int32_t stride_w;      /* 0 = same as `stride`; 1 with stride == 2 = stride (2, 1): rows halved, width kept
                              (PatchMerging of the recogniser, model/rec2/svtrv2.py:128); CONV mode, 3x3 only  */
    const void*  x;        /* input activations, NHWC bf16                                        */
    const void*  w;        /* weights [n_pad][taps*in_c] bf16, K order (64-channel chunk, tap, channel in chunk),
                              i.e. [n_pad][in_c/64][taps][64] -- what ops.pack_conv_weight() produces (STEM: see
                              ops.pack_stem_weight; DCN v2 launches: ops.pack_dcn_weight)                    */
    const float* scale;    /* [n_pad] per-column multiplier (folded BN gamma/sqrt(var+eps)), or 1 */
    const float* shift;    /* [n_pad] per-column addend (folded BN beta-mean*scale, or bias)      */
    const void*  residual; /* NHWC bf16 or NULL                                                   */
    void*        y;        /* output                                                              */
    const void*  offsets;  /* DCN: b200_dcn_coord [batch*out_h*out_w][9], written by an offset conv with
                              B200_EPI_DCN_COORDS or by b200_dcn_coords() from raw offsets        */
    /* B200_EPI_CHAIN only (NULL / 0 otherwise) */
    const void*  chain_w;     /* [chain_n][out_c] bf16, K-major (what ops.pack_conv_weight gives for the next 1x1 conv) */
    const float* chain_scale; /* [chain_n] */
    const float* chain_shift; /* [chain_n] */
    void*        chain_y;     /* second output, NHWC bf16, chain_ld elements per pixel               */
    int32_t      chain_n;     /* 64 or 128                                                           */
    int32_t      chain_ld;
    int32_t      chain_relu;  /* ReLU on the second output                                           */
    int32_t      chain_reserved;
    /* Second K segment (NULL / 0 = none): a 1x1 stride-1 CONV whose K dimension continues over a second NHWC tensor,
       y = epi([x | x2] . w^T) with w = [n_pad][in_c + in_c2].  A Bottleneck's first block uses it to run the projection
       shortcut INSIDE conv3 (torchvision resnet.py:155-161: out = bn3(conv3(t)) + downsample(x) is one GEMM over [t | x]
       once the BN scales are folded into the weights), instead of writing the projection and reading it back as a
       residual.  x2 is sampled at (stride2 * oy, stride2 * ox).                                                         */
    const void*  x2;          /* NHWC bf16, spatial size (in_h * stride2, in_w * stride2)             */
    int32_t      in_c2;       /* channels of x2, multiple of 64                                       */
    int32_t      x2_ld;       /* elements between consecutive pixels of x2; 0 = in_c2                 */
    int32_t      stride2;     /* 1 or 2                                                               */
    int32_t      x2_reserved;
} b200_igemm_desc;

typedef struct b200_igemm b200_igemm; /* opaque: TMA descriptors + launch geometry */

int b200_igemm_create(const b200_igemm_desc* desc, b200_igemm** out);
int b200_igemm_launch(const b200_igemm* op, void* stream);
int b200_igemm_destroy(b200_igemm* op);
/* geometry a created op settled on (for tests / DESIGN tables) */
int b200_igemm_info(const b200_igemm* op, int32_t* out_h, int32_t* out_w, int32_t* tile_w,
                    int32_t* tile_h, int32_t* block_n, int32_t* num_tiles, int32_t* smem_bytes);

/* raw offset_mask_conv output [batch*out_h*out_w][32] fp32 (ch 2k = dy_k, 2k+1 = dx_k, 18+k = mask logit,
 * dcn.py:44-46) -> b200_dcn_coord [pixel][9] for a 3x3, pad 1 DCN over an (in_h, in_w) image with x_ld
 * elements per pixel.  (The engine fuses this into the offset conv's epilogue; this entry point serves callers
 * that bring their own offsets.) */
int b200_dcn_coords(const float* raw, void* coords, int batch, int in_h, int in_w, int x_ld, int stride,
                    void* stream);

/* ---- bandwidth-bound kernels ---------------------------------------------------------------- */
/* fp32 NCHW (b,3,H,W) normalised image -> zero-padded space-to-depth bf16 buffer
 * [b][H/2+3][W/2+3][16] (channel (dy*2+dx)*3+c of block (row-2, col-2); 12..15 zero) consumed by
 * B200_IGEMM_STEM with weights [n_pad][4 rows][4 blocks][16] (replaces the layout half of
 * torchvision conv1 7x7 s2, resnet.py:197 via model/det/backbone.py:34). */
int b200_stem_pack(const float* img, void* packed, int batch, int h, int w, void* stream);
/* Same, from the decoded image: uint8 HWC RGB (b, H, W, 3) -> (v/255 - mean[c]) / std[c] -> packed bf16 buffer.
 * Replaces the host-side normalisation + HWC->CHW permute of src/pipeline/pipeline2.py:312-314 (mean3 / std3 are
 * HOST pointers to 3 floats). */
int b200_stem_pack_u8(const void* img_hwc, void* packed, int batch, int h, int w, const float* mean3,
                      const float* std3, void* stream);
/* cv2.resize(img, (dst_w, dst_h)) with the default INTER_LINEAR for an 8-bit 3-channel HWC image, bit for bit (OpenCV's
 * fixed-point path): replaces the host resize of resize_image_for_det (src/pipeline/pipeline2.py:34-40) and resize_image
 * (src/det/test.py:108-121).  Pitches are bytes per row; dst may be one slot of a batch buffer. */
int b200_resize_u8_linear(const void* src_hwc, int src_h, int src_w, int src_pitch, void* dst_hwc, int dst_h, int dst_w,
                          int dst_pitch, void* stream);
/* sigmoid + differentiable-binarisation step (model/det/head.py:28-48) from the two fp32 logit maps (n values each):
 * binary = sigmoid(bin), thresh = sigmoid(thr), thresh_binary = 1 / (1 + exp(-k (binary - thresh))); NULL outputs are
 * skipped; bin_logits_out / thresh_logits_out receive copies of the inputs when given. */
int b200_db_maps(const float* bin_logits, const float* thresh_logits, float k, float* binary, float* thresh,
                 float* thresh_binary, float* bin_logits_out, float* thresh_logits_out, int64_t n, void* stream);
/* MaxPool2d(3, stride 2, pad 1) on NHWC bf16 (torchvision resnet.py:200). */
int b200_maxpool3x3s2(const void* x, void* y, int batch, int h, int w, int c, void* stream);
/* ScaleFeatureSelection.forward (model/det/neck.py:57-79): bilinear(align_corners=True) upsample of
 * p3,p4,p5 to p2's size, 1x1 conv 1024->4 (+bias), softmax over the 4 levels, weighted sum.
 * p2..p5 NHWC bf16 with 256 channels; w_att [4][1024] fp32; b_att [4]; out NHWC bf16. */
int b200_asf_fwd(const void* p2, const void* p3, const void* p4, const void* p5,
                 const float* w_att, const float* b_att, void* out,
                 int batch, int h2, int w2, int h3, int w3, int h4, int w4, int h5, int w5,
                 void* stream);
/* Same module, split form used by the engine: z_l = W_l p_l (the 4 attention-logit contributions of level l,
 * conv_atten's weight columns [256 l, 256 (l+1)), no bias) computed at each level's native resolution by the
 * implicit-GEMM kernel, fp32 [pixel][zld]; this call interpolates the z maps (bilinear, align_corners=True, exact
 * because upsampling commutes with the 1x1 conv), adds b_att, takes the softmax and blends p2..p5. */
int b200_asf_combine(const void* p2, const void* p3, const void* p4, const void* p5,
                     const float* z2, const float* z3, const float* z4, const float* z5, int zld,
                     const float* b_att, void* out, int batch, int h2, int w2, int h3, int w3,
                     int h4, int w4, int h5, int w5, void* stream);
/* The same blend on the tensor cores (asf_mma.cu): per 16 x 8 output tile one 128 x 256 x 112 UMMA chain
 * out = S * P, where P holds the (at most 10 x 6 / 6 x 4 / 4 x 3) source pixels of p3 / p4 / p5 the tile can touch --
 * loaded by TMA and used as an MN-major operand as they land -- and S the softmax x bilinear weights (bf16); p2 enters
 * in the epilogue.  A plan bakes the TMA descriptors for fixed tensors; same arguments and result as
 * b200_asf_combine (up to the bf16 rounding of the twelve weights per pixel).  b200_asf_plan_supported() says
 * whether the level sizes fit the kernel's windows (each level about half the previous one, none smaller than
 * its window); otherwise use b200_asf_combine. */
typedef struct b200_asf_plan b200_asf_plan;
int b200_asf_plan_supported(int h2, int w2, int h3, int w3, int h4, int w4, int h5, int w5);
int b200_asf_plan_create(const void* p2, const void* p3, const void* p4, const void* p5,
                         const float* z2, const float* z3, const float* z4, const float* z5, int zld,
                         const float* b_att, void* out, int batch, int h2, int w2, int h3, int w3,
                         int h4, int w4, int h5, int w5, b200_asf_plan** plan);
int b200_asf_plan_launch(const b200_asf_plan* plan, void* stream);
int b200_asf_plan_destroy(b200_asf_plan* plan);
/* DBHead tail (model/det/head.py:15,23,28-48): last ConvTranspose2d(64->1, 2x2 s2) of both branches,
 * sigmoid, differentiable-binarisation step.  t_bin/t_thr: NHWC bf16, 64 channels each (already
 * BN+ReLU'd), `ld` elements between consecutive pixels (64 for two dense tensors, 128 when both branches
 * are interleaved in one tensor); w2_* [64][4] fp32 (ci, dy*2+dx); maps: fp32 (b,1,2h,2w) each; any map
 * may be NULL. */
int b200_db_head_fwd(const void* t_bin, const void* t_thr, const float* w2_bin, float b2_bin,
                     const float* w2_thr, float b2_thr, float k,
                     float* binary, float* thresh, float* thresh_binary,
                     float* bin_logits, float* thresh_logits,
                     int batch, int h, int w, int ld, void* stream);

/* ---- post-processing (replaces DBPostProcessor.__call__, src/det/test.py:55-106) ------------
 * threshold (test.py:57) -> 8-connected component labelling (replaces cv2.findContours, test.py:60)
 * -> per-component bounding box, pixel count and fixed-point probability sum (replaces
 * box_score_fast, test.py:20-34) -> box_thresh / min_area filters (test.py:77,81) -> closed-form
 * un-clip of the bounding rectangle by d = area * ratio / perimeter (replaces unclip, test.py:37-43).
 * All arithmetic after the threshold is integer or IEEE double without contraction, so the result
 * is bit-exact against oracle/postproc_oracle.py. */
typedef struct b200_box {
    int32_t x0, y0, x1, y1;     /* un-clipped (expanded) rectangle, pixel-centre corners            */
    int32_t bx0, by0, bx1, by1; /* tight bounding box of the component, inclusive                   */
    int32_t label;              /* canonical label: rank (1-based) of the component's first pixel   */
    int32_t npix;               /* pixels in the component                                          */
    uint64_t qsum;              /* sum over pixels of (uint)(prob * 2^24 + 0.5): score = qsum/(npix*2^24) */
} b200_box;

/* bytes of device scratch b200_postproc needs for a batch of (h, w) maps; -1 on a bad shape */
int64_t b200_postproc_workspace_bytes(int batch, int h, int w, int max_candidates);
/* prob: fp32 (b, h, w) device.  labels: int32 (b, h, w) device, receives the canonical label map
 * (0 = background; label L = the component whose first pixel in raster order is the L-th such).
 * boxes: device array [batch][max_candidates]; counts: device int32 [batch] = boxes written per
 * image (in label order); ncomp: device int32 [batch] = components found per image.
 * Only the first max_candidates components (by label) are examined (test.py:66-67). */
int b200_postproc(const float* prob, int32_t* labels, b200_box* boxes, int32_t* counts,
                  int32_t* ncomp, void* workspace, int64_t workspace_bytes,
                  int batch, int h, int w, double thresh, double box_thresh, double unclip_ratio,
                  int min_area, int max_candidates, void* stream);

/* ---- detection -> recognition hand-over -------------------------------
 * Replaces, for a whole batch of pages in one launch and without the boxes leaving the device:
 *   the box rescale loop            src/pipeline/pipeline2.py:323-328 (float64 quotient, truncated toward zero)
 *   crop_image                      src/det/test.py:123-130           (cv2.boundingRect + clamps + slice)
 *   preprocess_for_recognition      src/pipeline/pipeline2.py:92-128  (cv2.resize to height th, white pad to tw,
 *                                                                      /255, mean, std, HWC -> CHW fp32)
 *   the blank tensor of recognize_text_batch for an empty crop (pipeline2.py:153-157).
 * Bit-exact against oracle/recprep_oracle.py, which is pinned to the reference's own functions. */
#define B200_REC_MAX_IMAGES 32
enum {
    B200_REC_OK = 0,
    B200_REC_EMPTY = 1,      /* crop.size == 0: the tensor is all zeros (pipeline2.py:155-157)                 */
    B200_REC_DEGENERATE = 2  /* int(w * th / h) == 0: cv2.resize raises in the reference; zeros are written
                                and the host side of this library raises unless told otherwise             */
};
typedef struct b200_image {  /* HOST-side descriptor of one original page resident on the device */
    const void* pixels;      /* device, uint8 HWC RGB, `pitch` bytes per row                      */
    int32_t h, w, pitch, reserved;
    double scale_h, scale_w; /* (new_h / h, new_w / w) of resize_image_for_det (pipeline2.py:40): boxes are DIVIDED by them */
} b200_image;
typedef struct b200_rec_rect {
    int32_t image;           /* index into `images`                                               */
    int32_t x, y, w, h;      /* the slice crop_image takes: page[y:y+h, x:x+w]; w or h <= 0 = empty */
    int32_t status;          /* B200_REC_*                                                        */
} b200_rec_rect;
/* boxes / counts: the device arrays b200_postproc wrote for `n_images` maps ([n_images][max_candidates], [n_images]).
 * Crop slots are the kept boxes of image 0, then image 1, ... (label order inside an image).
 * out: device fp32 [capacity][3][th][tw]; rects: device [capacity]; total: device int32[1] = number of boxes
 * (slots past `capacity` are not written).  Nothing is read back: the caller may queue the recogniser behind it. */
int b200_rec_prep_boxes(const b200_image* images, int n_images, const b200_box* boxes, const int32_t* counts,
                        int max_candidates, float* out, b200_rec_rect* rects, int32_t* total, int capacity,
                        int th, int tw, const float* mean3, const float* std3, void* stream);
/* the same from explicit rectangles (device array, fields image / x / y / w / h; a rectangle that is not inside its
 * page is treated as empty): preprocess_for_recognition for crops the caller chose. */
int b200_rec_prep_rects(const b200_image* images, int n_images, const b200_rec_rect* rects_in, int n, float* out,
                        b200_rec_rect* rects, int32_t* total, int capacity, int th, int tw,
                        const float* mean3, const float* std3, void* stream);

/* ---- greedy CTC decode (tail of the recognition path) --------------------
 * Replaces SVTRv2.decode_probs, model/rec2/svtrv2.py:550-569: argmax over the class axis of (T, B, C) fp32
 * log-probabilities (first index of the maximum, NaN = maximum, as torch.argmax), then per sequence the frames whose
 * class is neither `blank_id` nor the previous frame's class.  Element (t, b, c) is at log_probs[t*stride_t + b*stride_b
 * + c] (strides in elements).  tokens: device int32 [B][T], the kept classes then -1 padding; lengths: device int32 [B].
 * The id -> character join of Tokenizer.decode (model/rec2/tokenizer.py:54-79) is host string work. */
int b200_ctc_greedy(const float* log_probs, int T, int B, int C, int64_t stride_t, int64_t stride_b, int blank_id,
                    int32_t* tokens, int32_t* lengths, void* stream);

/* ---- SVTRv2 recogniser: the kernels around the tcgen05 GEMMs ------------
 * Every nn.Linear / nn.Conv2d of model/rec2/svtrv2.py runs through b200_igemm_* (tokens are the pixels of an NHWC image);
 * these cover the rest.  All activations bf16, channels contiguous; statistics and accumulation in fp32. */
/* ConvStem.conv1 + bn1 + GELU (svtrv2.py:112-114,120): x fp32 NCHW (batch, 3, h, w), w fp32 [mid][27] in (ci, ky, kx) order,
 * scale/shift = folded BatchNorm + conv bias, y bf16 NHWC (batch, h/2, w/2, mid). */
int b200_svtr_stem1(const float* x, const float* w, const float* scale, const float* shift, void* y,
                    int batch, int h, int wd, int mid, void* stream);
/* nn.LayerNorm over the last axis (svtrv2.py:93,95,173-187,446): rows of d channels, leading dimensions in elements. */
int b200_layernorm(const void* x, void* y, const float* gamma, const float* beta, int64_t rows, int d,
                   int x_ld, int y_ld, float eps, void* stream);
/* y = GELU(x) (erf form) [+ residual when given] on n bf16 elements; y may equal x. */
int b200_gelu(const void* x, const void* residual, void* y, int64_t n, void* stream);
/* softmax(q k^T * scale) v per (sequence, head), head_dim 32 (GlobalMixing svtrv2.py:79-86, FRM rows :208-215).
 * qkv: [n_seq][n][3*d], channel = which*d + head*32 + e; out: [n_seq][n][d], channel = head*32 + e. */
int b200_mhsa(const void* qkv, void* out, int n_seq, int n, int d, float scale, void* stream);
/* FRM vertical cross-attention of the selecting token (svtrv2.py:236-243): q fp32 [d] (parameters only),
 * kv: [batch][h][w][2*d] (k | v), out: [batch][w][d]. */
int b200_frm_select(const float* q, const void* kv, void* out, int batch, int h, int w, int d, float scale, void* stream);
/* rows 0, 2, 4, .. of an NHWC tensor: the stride (2, 1) of PatchMerging (svtrv2.py:128) applied to a stride-1 conv output. */
int b200_decimate_rows(const void* x, void* y, int batch, int in_h, int w, int c, void* stream);
/* F.log_softmax over the classes + permute to (T, B, C) (svtrv2.py:536-537): logits fp32 [batch*t_len][ld]. */
int b200_log_softmax_tbc(const float* logits, float* out, int batch, int t_len, int c, int ld, void* stream);

/* ---- plain CUDA-core cross-check kernels (slow; used by the GPU tests, never by the engine) -- */
int b200_ref_conv2d(const void* x, const void* w, const float* scale, const float* shift,
                    const void* residual, void* y, int batch, int in_h, int in_w, int in_c,
                    int out_c, int w_ld, int ksize, int stride, int flags, int out_ld, int res_ld,
                    int shuf_c, void* stream);
/* offsets here are the RAW [pixel][32] fp32 offset_mask_conv output */
int b200_ref_dcn2d(const void* x, const float* offsets, const void* w, const float* scale,
                   const float* shift, void* y, int batch, int in_h, int in_w, int in_c,
                   int out_c, int stride, int flags, int out_ld, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DBNET_B200_H */
