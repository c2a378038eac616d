/* wedetect_b200.h — C ABI of libwedetect_b200.so (sm_100a).
 *
 * The reference (WeChatCV/WeDetect) has no FFI: its boundary is Python.  This ABI
 * is what sits directly underneath our Python mirror of that boundary.  Each entry point cites the
 * reference function(s) whose arithmetic it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - 16-bit GEMM operands come in two formats.  Plane stride 0: one bf16 plane (the opt-in fast mode).  Plane stride
 *     > 0 (the default, parity-grade mode): two fp16 planes `plane_stride` elements apart starting at the given pointer,
 *     hi + lo = value * s with s a power of two: WD_ACT_PLANE_SCALE for every activation the library writes, a
 *     per-matrix value chosen by the host for weights (undone by the GEMM's `acc_scale`);
 *   - plain C types only; all pointers in `wd_op.p[]` are DEVICE pointers owned by the caller;
 *   - every call enqueues work on the caller's `cudaStream_t` (passed as void*), never syncs;
 *   - return 0 on success, negative on failure; `wd_last_error()` returns a thread-local message;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Execution model: the host (Python) lowers a model into a flat array of `wd_op` records (one per
 * kernel-level operation), `wd_program_create` validates them and pre-builds TMA descriptors,
 * `wd_program_run` replays the whole list on a stream (optionally as a captured CUDA graph).
 */
#ifndef WEDETECT_B200_H
#define WEDETECT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WD_OP_NI 48
#define WD_OP_NF 8
#define WD_OP_NP 16
#define WD_ACT_PLANE_SCALE 4.0f

typedef struct wd_op {
    int32_t kind;          /* enum wd_op_kind */
    int32_t i[WD_OP_NI];   /* integer fields, meaning depends on kind (see below) */
    float f[WD_OP_NF];     /* float fields */
    void* p[WD_OP_NP];     /* device pointers */
} wd_op;

typedef struct wd_program wd_program;

enum wd_op_kind {
    /* Tensor-core GEMM / implicit-GEMM convolution with fused epilogue (tcgen05 + TMEM + TMA).
     * Replaces nn.Linear / 1x1 / 3x3 s1 / 2x2-s2-transposed convolutions, BatchNorm (folded),
     * ReLU/SiLU/GELU, LayerScale+residual and DFL of:
     *   wedetect/models/backbones/mm_backbone.py:112-125 (Block.forward pwconv1/act/pwconv2/gamma)
     *   wedetect/models/necks/yolo_world_pafpn.py:40-68,195-208,587-605,631-647,692-715
     *   wedetect/models/dense_heads/yolo_world_head.py:90-108,263-294
     *   transformers XLMRobertaModel Linear layers (mm_backbone.py:382-386)
     * i: 0..2 D0,D1,D2 (row space, row m = (d2*D1+d1)*D0+d0)   3..5 E0,E1,E2 (tile extents, E0*E1*E2<=128)
     *    6 Kc (K per tap, %64==0)  7 ntaps (1|9)  8 N  9..11 A strides of d0,d1,d2 (elements)
     *    12 ldb  13 block_n (64|128|256; fp16 hi/lo mode: 64|128)  14 out_dtype (0 16-bit: bf16, or fp16 hi/lo planes when planes = 2; 1 f32)  15 act (wd_act)
     *    16 resid_dtype (0 none, 1 16-bit (format as A), 2 f32)  17 ld_res  18 group_cols  19 n_groups
     *    20..22 C strides of d0,d1,d2 (elements)  23 C stride of group  24 epi_mode (0 store, 1 DFL)
     *    25 tap_w (3 for 3x3)  26 pad (1 for 3x3)  27 group_valid (columns of a group present in memory, 0 = group_cols)
     *    28 K_valid (channels of A present in memory, 0 = Kc; TMA zero-fills up to Kc)
     *    29 BK_valid (columns of B present in memory, 0 = ntaps*Kc)
     *    36 no_pair (1: do not use 2-CTA clusters (cta_group::2 UMMA) for 256-wide tiles)
     *    38, 39 input width / height of a stride-2 3x3 convolution (both > 0: A is walked with TMA element strides 2, the
     *           tile / D extents then describe the OUTPUT map); 0 = stride 1
     *    37 no_warp_store (1: one 128-row TMA store per epilogue warpgroup instead of one 32-row store per warp; A/B switch)
     *    35 exact_act (1: erf-GELU / exp-SiLU instead of the MUFU.TANH forms used for bf16 outputs of the fast path)
     *    30 planes (0|1 bf16 fast mode, 2 fp16 hi/lo)  31 A plane stride  32 B plane stride  33 C plane stride  34 resid plane stride
     *    40 lblk (fp16 hi/lo mode: 64-wide k-blocks (1|2) accumulated in TMEM before the partial sum moves to fp32 registers)
     *    41 no_trunc_comp (1: do not compensate the tensor pipe's truncating accumulation; measurement switch)
     *    42 no_red_store (1: an in-place fp32 residual with alpha == 1 is loaded and added in registers instead of being added by the
     *       TMA store in the L2; same bits either way; measurement switch)
     * f: 0 resid_alpha  1 acc_scale (fp16 hi/lo mode: 1 / (A scale * B scale), 0 = 1)
     * p: 0 A  1 B [N, ntaps*Kc]  2 C  3 bias f32[N]  4 gamma f32[N]  5 resid
     * out = resid*alpha + gamma * act(acc * acc_scale + bias)        (each term optional) */
    WD_OP_GEMM = 1,
    /* Row LayerNorm over C (biased variance, eps inside sqrt): mm_backbone.py:145-155, F.layer_norm.
     * i: 0 rows 1 C
     *    4 s2d (0 | 1: write 2x2 space-to-depth layout for the stride-2 patchify conv, :193-198)
     *    5 W 6 H (only for s2d) 7 ld_in 8 ld_out
     *    30 out plane stride
     * f: 0 eps   p: 0 in f32  1 out bf16 (nullable)  2 weight f32[C]  3 bias f32[C]  6 out f32 (nullable) */
    WD_OP_LN_ROWS = 2,
    /* Depthwise 7x7 (pad 3) + bias + LayerNorm(C) on an NHWC fp32 tensor -> bf16 rows.
     * mm_backbone.py:114-116 (Block.forward dwconv/permute/norm).
     * i: 0 B 1 H 2 W 3 C 4 ld_out (0 = C) 5 tx 6 ty (block = 8tx x 4ty pixels; used with the scratch) 30 out plane stride
     * f: 0 eps
     * p: 0 in f32 [B,H,W,C]  1 out bf16 [B*H*W,ld_out]  2 w f32[49,C]  3 b f32[C]  4 ln_w  5 ln_b
     *    7 scratch f32 [B*H*W, C] (non-null selects the shared-memory tiled kernel) */
    WD_OP_DWCONV_LN = 3,
    /* Stem patchify: image -> bf16 rows [B*(H/4)*(W/4), 64] (48 valid = (dy,dx,c), rest 0).
     * mm_backbone.py:188-191 (Conv2d k4 s4 input gather); data_preprocessor.py:35-36 (mean/std are
     * folded into the stem weights by the host).
     * i: 0 B 1 H 2 W 3 in_dtype (0 u8, 2 f32) 4 layout (0 NCHW) 30 out plane stride   f: 0 scale
     * p: 0 in  1 out bf16 */
    WD_OP_STEM_PATCH = 4,
    /* im2col for 3x3 stride-2 pad-1 conv on NHWC bf16: rows [B*Ho*Wo, 9*C] (tap-major).
     * yolo_world_pafpn.py:704-709,1062-1082 (downsample ConvBNReLU k3 s2).
     * i: 0 B 1 H 2 W 3 C 4 ld_in 30 out plane stride 31 in plane stride   p: 0 in bf16  1 out bf16 */
    WD_OP_IM2COL_S2 = 5,
    /* f32 -> bf16 row cast (backbone outputs c1..c4 to neck operands). i: 0 rows 1 C 2 ld_in 3 ld_out 30 out plane stride
     * p: 0 in f32  1 out bf16 */
    WD_OP_CAST_BF16 = 6,
    /* XLM-R embeddings: word + position + token_type, LayerNorm (transformers
     * modeling_xlm_roberta.py embeddings; position ids = cumsum(mask)*mask + pad_idx).
     * i: 0 S (sequences) 1 L (tokens/seq) 2 Hd 3 pad_idx 30 out plane stride   f: 0 eps
     * p: 0 ids i32[S,L] 1 mask i32[S,L] 2 word f32[V,Hd] 3 pos f32[P,Hd] 4 type f32[Hd] 5 ln_w 6 ln_b
     *    7 out f32 [S*L,Hd]  8 out bf16 */
    WD_OP_TEXT_EMBED = 7,
    /* Short-sequence multi-head self-attention: L <= 32 one warp per (sequence, head); 32 < L <= 128 one block per (sequence, head).
     * i: 0 S 1 L 2 heads 3 head_dim 4 ld_qkv (= 3*Hd) 30 out plane stride   f: 0 scale
     * p: 0 qkv f32 [S*L, 3*Hd]  1 mask i32[S,L]  2 out bf16 [S*L,Hd] */
    WD_OP_ATTN_SMALL = 8,
    /* CLS pooling + L2 normalise: out[s,:] = x[s,:] / max(||x[s,:]||, 1e-12)  (F.normalize,
     * mm_backbone.py:387).  i: 0 S 1 C 2 ld_in   p: 0 in f32  1 out f32 */
    WD_OP_L2NORM_ROWS = 9,
    /* gather rows: out[s,:] = in[idx0 + s*stride,:] as bf16 (CLS token rows, mm_backbone.py:385).
     * i: 0 S 1 C 2 row_stride 3 ld_in 30 out plane stride  p: 0 in f32  1 out bf16 */
    WD_OP_GATHER_ROWS = 10,
    /* Fold BNContrastiveHead into a GEMM weight: W'[k,c] = t[k,c]/max(||t[k]||,eps) * g[c] * exp(s),
     * b'[k] = exp(s) * sum_c h[c]*tn[k,c] + bias (yolo_world_head.py:90-108; Uni variant without
     * text normalisation: generate_proposal.py:1129-1131).
     * i: 0 K 1 C 2 normalize (0|1) 3 K_pad 30 W' plane stride    f: 0 power-of-two scale of W' (fp16 hi/lo planes; 0 = 1)
     * p: 0 text f32[K,C] 1 bn_g f32[C] 2 bn_h f32[C] 3 logit_scale f32[1] 4 bias f32[1]
     *    5 W' bf16 [K_pad,C]  6 b' f32[K_pad] */
    WD_OP_FOLD_TEXT = 11,
    /* Detection post-process for a batch: sigmoid, score threshold, top-k (nms_pre), box decode,
     * rescale, class-aware greedy NMS, keep max_per_img.  Bit-exact integer/index semantics.
     *   yolo_world_head.py:619-749 (predict_by_feat), generate_proposal.py:85-131,1000-1048,1150-1218,
     *   mmdet filter_scores_and_topk / batched_nms.
     * see wd_pp_params below (passed through p[0] as a HOST pointer copied at create time). */
    WD_OP_POSTPROCESS = 12,
    /* Gather kept proposals' embedding rows: out[b, j, :] = BN(embed[b, anchor(b,j), :]) f32.
     * generate_proposal.py:1129,1209-1212.  i: 0 B 1 A 2 C 3 max_keep  4 nlevels 5..7 level sizes
     *    30..32 embed plane strides
     * p: 0..2 embed bf16 per level [B*HW_l, C]  3 keep_anchor i32[B,max]  4 counts i32[B]
     *    5 bn_g f32[3*C] 6 bn_h f32[3*C] 7 out f32 [B,max,C]
     *    optional (all four or none; eval_retrieval/extract_embedding.py:1181-1190,1238-1260 `scales` / `bias`):
     *    8 logit_scale per level f32[nlevels]  9 bias per level f32[nlevels]  10 out scales f32[B,max]  11 out bias f32[B,max] */
    WD_OP_GATHER_EMBED = 13,
    /* Row-scaled cast: out[r,:] = bf16(in[r,:] * exp(scale[r])) for rows r = b*P + j with j < counts[b], else 0.
     * First stage of retrieval scoring (eval_retrieval/retrieval_metric.py:365-372): the per-proposal exp(scale) is
     * applied to the embedding row so that the class logits are one GEMM against the text matrix.
     * i: 0 B 1 P (rows per image) 2 C 30 out plane stride
     * p: 0 in f32 [B*P, C]  1 scale f32 [B*P] (nullable: 1)  2 counts i32 [B] (nullable: all rows)  3 out bf16 [B*P, C] */
    WD_OP_SCALE_ROWS = 14,
    /* Retrieval score reduce: out[b,k] = max_{j < counts[b]} sigmoid(z[b*P+j, k] + bias[b*P+j]); 0 when counts[b] == 0.
     * eval_retrieval/retrieval_metric.py:372-373 (sigmoid, max over proposals).
     * i: 0 B 1 P 2 K 3 ldz
     * p: 0 z f32 [B*P, ldz]  1 bias f32 [B*P] (nullable: 0)  2 counts i32 [B] (nullable: P)  3 out f32 [B, K] */
    WD_OP_RETR_REDUCE = 15,
    /* Letterbox a batch of decoded RGB images into the detector's planar uint8 input: PIL-exact BILINEAR resize
     * (Pillow Resample.c: 22-bit fixed-point triangle filter, horizontal pass into an 8-bit intermediate, then vertical)
     * + centred paste on a `pad`-grey canvas.  generate_proposal.py:17-82 (letterbox), :1087-1101 (forward).
     * i: 0 B 1 H 2 W (canvas) 3 pad value (114)
     * p: 0 src u8: the images back to back, each [h, w, 3] RGB interleaved
     *    1 desc i32 [B, 16] per image: 0,1 byte offset of the image in src (lo, hi)  2 src_w  3 src_h  4 new_w  5 new_h
     *      6 left  7 top (paste position)  8 first source row the vertical pass needs  9 rows of the intermediate image
     *      10,11 byte offset of the intermediate image in tmp (lo, hi)  12 offset (int32 words) of the image's tables in coef
     *      13 ksize_h  14 ksize_v  15 vertical_first (Pillow >= 12 for h > 100 w when the height shrinks: the intermediate
     *      image is then [new_h, src_w] and desc[8] is 0).   new_w == 0: the whole canvas is padding (unused batch slot)
     *    2 coef i32: per image  bounds_h [new_w, 2] (first, count), kk_h [new_w, ksize_h], bounds_v [new_h, 2] (first row is
     *      relative to desc[8]), kk_v [new_h, ksize_v]; weights are PIL's (int)(0.5 + w * 2^22)
     *    3 tmp u8 workspace (sum over images of rows * new_w * 3 bytes)  4 out u8 [B, 3, H, W]
     * desc / coef / src are re-filled by the host before every run (the pointers are fixed, so the op can sit in a graph). */
    WD_OP_LETTERBOX = 16,
    /* ConvNeXt block MLP in one kernel (fast path, C = 128 / hidden 512 only: WeDetect-Base stage 0):
     *   x[m,:] += gamma * (W2 . GELU(W1 . t[m,:] + b1) + b2)      mm_backbone.py:117-124 (pwconv1, act, pwconv2, gamma, residual)
     * The 4C-wide hidden activation stays on chip (TMEM -> bf16 shared-memory operand), weights are resident in the
     * shared memory of a CTA pair.  Same arithmetic as two WD_OP_GEMM records (bf16 operands, fp32 accumulation, bf16 hidden).
     * i: 0 M 1 C 2 H 3 ld of t (0 = C)
     * p: 0 t bf16 [M, C]  1 W1 bf16 [H, C]  2 W2 bf16 [C, H]  3 b1 f32[H]  4 b2 f32[C]  5 gamma f32[C]  6 x f32 [M, C] (in place) */
    WD_OP_MLP_FUSED = 17,
    /* The mmcv test pipeline of infer_wedetect.py / test.py on the device: cv2.resize (OpenCV 4.x arithmetic, bit-exact) of a batch
     * of decoded uint8 BGR images + paste at (left, top) on a `pad`-grey canvas, into the detector's planar uint8 input.
     * WeDetectKeepRatioResize (transforms.py:94-123: INTER_AREA when the ratio is < 1, else INTER_LINEAR) followed by
     * WeDetectLetterResize (transforms.py:180-272: pad 114, top/left = round(pad // 2 - 0.1)); mmcv.imresize / impad = cv2.resize /
     * cv2.copyMakeBorder (mmcv 2.1.0, un-vendored).
     * i: 0 B 1 H 2 W (canvas) 3 pad value
     * p: 0 src u8: the images back to back, each [h, w, 3] interleaved (channel order is passed through)
     *    1 desc i32 [B, 16] per image: 0,1 byte offset in src (lo, hi)  2 src_w  3 src_h  4 new_w  5 new_h  6 left  7 top
     *      8 mode (0 copy, 1 INTER_AREA with tables, 2 INTER_AREA integer boxes, 3 INTER_LINEAR)  9 offset (int32 words) of the image's
     *      tables in coef  10 kx 11 ky 12 bits of the float 1.f / (kx ky) (mode 2)  13 xmax (mode 3: first column that replicates the
     *      last source column).  new_w == 0: the whole canvas is padding
     *    2 coef i32: mode 1: xidx [new_w + 1], yidx [new_h + 1] (entry ranges), x entries si [nx], alpha f32 [nx], y entries si [ny],
     *      beta f32 [ny] (OpenCV's computeResizeAreaTab);  mode 3: xofs [new_w], (alpha0 | alpha1 << 16) [new_w], yofs [new_h],
     *      (beta0 | beta1 << 16) [new_h] (11-bit fixed point, OpenCV's resizeGeneric_ tables)
     *    3 unused (nullable)  4 out u8 [B, 3, H, W] */
    WD_OP_CV_RESIZE_PAD = 18,
};

enum wd_act { WD_ACT_NONE = 0, WD_ACT_RELU = 1, WD_ACT_SILU = 2, WD_ACT_GELU = 3 };

/* Parameters of WD_OP_POSTPROCESS (host struct; device pointers inside). */
typedef struct wd_pp_params {
    int32_t B;              /* images */
    int32_t K;              /* classes / prompts */
    int32_t nlevels;        /* <= 4 */
    int32_t lvl_h[4], lvl_w[4], lvl_stride[4];
    int32_t ld_logit[4];    /* row stride (elements) of each level's logit matrix */
    const float* logits[4]; /* f32 [B*H_l*W_l, ld_logit] pre-sigmoid */
    const float* dist[4];   /* f32 [B*H_l*W_l, 4] ltrb in stride units (DFL output) */
    float score_thr;        /* keep score > thr (strict), mmdet filter_scores_and_topk */
    int32_t nms_pre;        /* top-k before NMS */
    float iou_thr;          /* suppress iff IoU > thr (strict) */
    int32_t max_per_img;
    int32_t nms_mode;       /* 0: mmcv batched_nms (coordinate offsets, always)
                               1: torchvision batched_nms (offsets iff 4*n <= 20000, else per class) */
    int32_t tv_numel_thr;   /* nms_mode 1 only: torchvision's switch (20000 on CUDA devices, 4000 on CPU) */
    int32_t multi_label;    /* 1: every (anchor,class) pair is a candidate; 0: argmax class per anchor */
    const float* img_meta;  /* f32 [B, 8]: pre_sub_x, pre_sub_y, pre_div_x, pre_div_y (applied before
                               NMS: mmdet rescale), post_sub_x, post_sub_y, post_div (after NMS: Uni
                               un-letterbox), unused */
    const float* clamp_wh;  /* f32 [B,2] (ori_w, ori_h); clamp applied last */
    /* outputs (caller allocated) */
    float* out_boxes;       /* [B, max_per_img, 4] */
    float* out_scores;      /* [B, max_per_img] */
    int32_t* out_labels;    /* [B, max_per_img] */
    int32_t* out_anchor;    /* [B, max_per_img] flat anchor index (level-concatenated) */
    int32_t* out_counts;    /* [B] */
    /* workspace */
    void* workspace;
    uint64_t workspace_bytes;
} wd_pp_params;

/* ---- library-level ---------------------------------------------------------------------- */
const char* wd_last_error(void);
int wd_version(void);
/* WD_ACT_PLANE_SCALE the library was built with (host mirrors must agree). */
float wd_act_plane_scale(void);
/* number of CUDA kernels launched by this library in this process (bench `gpu_launches`). */
uint64_t wd_launch_count(void);
/* device query: returns 0 and fills sm count / cc; fails loudly (negative) without a GPU. */
int wd_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ---- single op (used by parity tests; builds descriptors, launches, frees) --------------- */
int wd_op_run(const wd_op* op, void* stream);

/* ---- programs ---------------------------------------------------------------------------- */
int wd_program_create(const wd_op* ops, int n_ops, wd_program** out);
int wd_program_run(wd_program* prog, void* stream);
/* capture once into a CUDA graph and replay (falls back to an error, never to eager silently). */
int wd_program_capture(wd_program* prog, void* stream);
int wd_program_replay(wd_program* prog, void* stream);
int wd_program_num_launches(const wd_program* prog);
int wd_program_num_ops(const wd_program* prog);
/* measurement helper: eager run with CUDA events between ops; ms_per_op has wd_program_num_ops entries. */
int wd_program_run_timed(wd_program* prog, void* stream, float* ms_per_op);
/* debug helper: run op by op; *stuck_op = index of the first op not finished after timeout_ms (-1: none). */
int wd_program_find_stuck_op(wd_program* prog, void* stream, int timeout_ms, int* stuck_op);
void wd_program_destroy(wd_program* prog);

/* ---- JPEG decode on the device (optional) ---------------------------------
 * Replaces the reference's per-image CPU decode (LoadImageFromFile -> mmcv.imfrombytes -> cv2.imdecode, config/wedetect_base.py:112,
 * infer_wedetect.py:111; Image.open().convert("RGB"), generate_proposal.py:1089-1090) with nvJPEG writing interleaved pixels into
 * device memory (the `src` buffer of WD_OP_CV_RESIZE_PAD / WD_OP_LETTERBOX).  libnvjpeg is loaded lazily: wd_jpeg_open fails loudly
 * where it is missing, nothing else depends on it.  Not bit-identical to libjpeg-turbo on chroma-subsampled files (opt-in).
 *   wd_jpeg_info    width / height (and component count, nvjpegChromaSubsampling_t value) of a compressed image in HOST memory
 *   wd_jpeg_decode  decode into dst (DEVICE, [h, pitch] bytes, 3 interleaved channels; bgr != 0: B,G,R order) on `stream` */
typedef struct wd_jpeg wd_jpeg;
int wd_jpeg_open(wd_jpeg** out);
int wd_jpeg_info(wd_jpeg* j, const uint8_t* data, uint64_t len, int* width, int* height, int* components, int* subsampling);
int wd_jpeg_decode(wd_jpeg* j, const uint8_t* data, uint64_t len, uint8_t* dst, uint64_t pitch, int bgr, void* stream);
void wd_jpeg_close(wd_jpeg* j);

/* workspace size needed by WD_OP_POSTPROCESS for (B, anchors, K). */
uint64_t wd_pp_workspace_bytes(int B, int anchors, int K, int nms_pre);

#ifdef __cplusplus
}
#endif
#endif /* WEDETECT_B200_H */
