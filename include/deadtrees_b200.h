/* deadtrees_b200 — C-ABI of the B200 (sm_100a) hot path of cwerner/deadtrees.
 *
 * The reference is pure Python and has NO native/FFI boundary for this path; the
 * drop-in boundary is the set of Python signatures in `deadtrees_b200/` (mirroring `deadtrees/`).
 * Those call this library through ctypes.  Every entry point below names the reference code it
 * replaces (paths relative to the reference repo).
 *
 * Conventions
 *  - all data pointers are DEVICE pointers owned by the caller (PyTorch); the library never
 *    allocates or frees device memory; `stream` is a cudaStream_t; every call is asynchronous
 *  - return value: DT_OK or a negative DT_ERR_*; `dt_last_error` returns the thread-local message
 *  - there is no CPU fallback: on a device that is not compute capability 10.x every compute
 *    entry point returns DT_ERR_UNSUPPORTED_ARCH
 *  - activations are NHWC; "act dtype" 0 = bf16 (tensor-core path), 1 = fp32 (check mode)
 */
#ifndef DEADTREES_B200_H_
#define DEADTREES_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DT_VERSION 100

enum {
  DT_OK = 0,
  DT_ERR_BAD_SHAPE = -1,
  DT_ERR_BAD_ALIGN = -2,
  DT_ERR_UNSUPPORTED_ARCH = -3,
  DT_ERR_CUDA = -4,
  DT_ERR_UNSUPPORTED = -5
};

enum { DT_BF16 = 0, DT_F32 = 1 };

typedef void* dt_stream_t; /* cudaStream_t */

int dt_version(void);
/* copies the calling thread's last error message into buf (NUL-terminated); returns its length */
int dt_last_error(char* buf, size_t n);
/* DT_OK iff the current CUDA device is sm_100-class */
int dt_device_check(void);

/* ---- T1/T2: block split / merge --------------------------------------------------------------
 * deadtrees/utils/data_handling.py:9-19  make_blocks_vectorized(x, d):
 *   src (p, m, n) -> dst (m/d * n/d, p, d, d), block = row_block * (n/d) + col_block.
 * deadtrees/utils/data_handling.py:22-34 unmake_blocks_vectorized(x, d, m, n):
 *   src (m/d * n/d, d, d) -> dst (m, n).
 * Pure index permutations of `elem_size`-byte elements (1, 2, 4 or 8): bit-exact. */
int dt_make_blocks(const void* src, int p, int m, int n, int d, int elem_size, void* dst, dt_stream_t stream);
int dt_unmake_blocks(const void* src, int d, int m, int n, int elem_size, void* dst, dt_stream_t stream);

/* ---- K1: tile gather + normalise ---------------------------------------------------------------
 * Replaces Tiler.get_batches (deadtrees/deployment/tiler.py:142-145) followed by the per-tile
 * val_transform loop (scripts/inference.py:94-96, deadtrees/data/deadtreedata.py:148-154) and the
 * RGB slice (deadtrees/deployment/inference.py:57-59).
 * Mosaic: uint8, element (y, x, c) at mosaic + y*row_stride + x*pix_stride + c*chan_stride (bytes),
 * so both band-planar (rasterio) and interleaved layouts are accepted.  Tile t (0 <= t < ntiles)
 * is grid cell (tile0 + t) of a gy x gx grid, origin (ty*step, tx*step); pixels outside H x W read
 * as raw 0 (the reference zero-pads the uint8 array, tiler.py:106-111).
 * out[t, y, x, c] = (u8 - offset[c]) * scale[c] for c < C, 0 for C <= c < c_out; fp32 arithmetic
 * (subtract, then multiply), rounded to bf16 when out_dtype == DT_BF16.  offset/scale are HOST
 * pointers to 4 floats.  out_pad == 3 writes each tile into a (tile+6) x (tile+8) pixel frame at offset (3, 3)
 * whose border the caller has zeroed: the input layout of the TMA-im2col stem (DT_CONV_X_PAD3). */
int dt_tile_gather_normalize(const uint8_t* mosaic, int H, int W, int C, int64_t row_stride, int64_t pix_stride,
                             int64_t chan_stride, int tile, int step, int gx, int tile0, int ntiles,
                             const float* offset, const float* scale, int c_out, int out_dtype, int out_pad, void* out,
                             dt_stream_t stream);

/* NCHW fp32 batch (what callers hand to `PyTorchInference.run` / `self.model(img)`) -> NHWC with 4
 * channels in `out_dtype`; keeps the first C of C_src channels (the RGB slice of
 * deadtrees/deployment/inference.py:57-59), zero-fills the rest. */
int dt_pack_input_nchw(const float* x, int N, int C_src, int C, int H, int W, int out_dtype, void* out,
                       dt_stream_t stream);
/* The same conversion into the bf16 stem frame (N, H+6, W+8, 4) of dt_conv2d_fwd's DT_CONV_X_PAD3 layout: interior at
 * offset (3, 3), the zero border written by the kernel (the caller need not clear the frame). */
int dt_pack_input_nchw_frame(const float* x, int N, int C_src, int C, int H, int W, void* out, dt_stream_t stream);

/* ---- K12: stitching ----------------------------------------------------------------------------
 * dt_stitch_mask_u8: Tiler.put_batches at overlap 0 (deadtrees/deployment/tiler.py:147-170 ->
 * unmake_blocks_vectorized): per-tile class ids (ntiles, T, T) uint8 -> mosaic rows, cropped to H x W.
 * dt_stitch_blend_argmax (extension): overlapping tiles, logits (ntiles_total, T, T, K)
 * in `dtype`; for every mosaic pixel of rows [row0, row0+nrows): blended = sum_t w*logit / sum_t w
 * over covering tiles in (ty, tx) order, w = win[y]*win[x]; mask = first-max argmax; optional fp32
 * blended output (H, W, K).  `win` is a device array of T floats.  `ty_base` is the tile row stored
 * first in `logits` (0 for a whole mosaic; > 0 for a tile-row shard of a multi-GPU run). */
int dt_stitch_mask_u8(const uint8_t* tile_masks, int T, int gx, int tile0, int ntiles, uint8_t* mosaic_mask, int H,
                      int W, int64_t row_stride, dt_stream_t stream);
int dt_stitch_blend_argmax(const void* logits, int dtype, int K, int T, int overlap, int gy, int gx, int ty_base,
                           const float* win, uint8_t* mosaic_mask, float* blended, int H, int W, int row0,
                           int nrows, dt_stream_t stream);

/* ---- K2-K9: convolutions -----------------------------------------------------------------------
 * One fused conv = smp `Conv2dReLU` / torchvision BasicBlock conv with eval-mode BatchNorm folded:
 *   y = act( scale[co] * conv(x)[co] + shift[co] (+ residual) ).
 * Input is the *virtual* tensor cat([up2(x) if upsample else x, skip], C) — the nearest-x2 upsample
 * and the concat of the smp UnetDecoder block (deadtrees/network/extra/resunet/decoder.py:41-43 shows
 * the convention) are never materialised.
 * Layouts: x (N, H/(upsample?2:1), W/(..), C_x), skip (N, H, W, C_in - C_x), y (N, Ho, Wo, C_out) NHWC.
 * Weights (packed by the host package):
 *   dtype DT_F32  : float [R*S][C_in][C_out]
 *   dtype DT_BF16 : bf16 [C_out][Kpad], k = (r*S + s)*C_in + ci, Kpad = roundup(R*S*C_in, 64);
 *                   stem (C_in == 4, R == S == 7): k = r*32 + s*4 + ci, Kpad = 256. */
typedef struct {
  int32_t N, H, W;      /* batch, spatial size of the (virtual) conv input */
  int32_t C_in;         /* channels of the virtual input */
  int32_t C_x;          /* channels taken from x (== C_in when there is no skip) */
  int32_t upsample;     /* 1: x is stored at half resolution and nearest-upsampled x2 */
  int32_t C_out, R, S, stride, pad;
  int32_t relu;         /* apply ReLU */
  int32_t has_residual; /* add residual (N, Ho, Wo, C_out) before the ReLU */
  int32_t dtype;        /* DT_BF16: tcgen05 implicit GEMM; DT_F32: CUDA-core check mode */
  int32_t flags;        /* DT_CONV_* */
} dt_conv_desc;

enum {
  DT_CONV_FORCE_GATHER = 1, /* A operand through the generic gather producer even if TMA-eligible */
  DT_CONV_FORCE_DIRECT = 2, /* CUDA-core direct kernel for bf16 tensors (debug/validation) */
  DT_CONV_X_PAD3 = 8,       /* 7x7 stem: x is a zero-bordered (N, H+6, W+8, 4) frame (see dt_tile_gather_normalize);
                               im2col through a TMA tensor map with overlapping strides (conv_stem.cu) */
  DT_CONV_NO_HALO = 4,      /* 3x3/s1 layers: per-tap TMA boxes (conv_tc.cu) instead of the smem-resident halo
                               patches (conv_halo.cu, the default where the shape fits) */
  DT_CONV_NO_QUAD = 32,     /* up-sample + concat layers with C_out <= 64: one parity class per tile instead of the
                               class-fused tiles that share the five patches of a region (same results; A/B testing) */
  DT_CONV_UPS_FOLDED = 64,  /* up-sampled input: the nearest-x2 up-sampling is folded into the weights of the x operand -
                               for output parity class (a,b) the nine taps touch only the 2 x 2 low-res pixels
                               (a-1+ey, b-1+ex), and w holds the per-class sum of the taps sharing a pixel: bf16
                               [C_out][16*C_x + 9*C_s] with k = ((a*2+b)*4 + ey*2+ex)*C_x + ci for the x operand,
                               followed by k = 16*C_x + tap*C_s + cs for the skip operand (C_s = C_in - C_x).  Exact in
                               real arithmetic; in bf16 the summed weight is rounded once instead of each tap's weight
                               (4 instead of 9 tap-MMAs per class).  Shapes: no skip, C_in <= 64, C_out <= 32
                               (resident-weight kernel); with skip, C_x and C_s multiples of 64, C_out 32 or 64
                               (class-fused kernel) */
  DT_CONV_PAIR = 128,       /* 3x3/s1 layers with C_in % 64 == 0, C_out % 128 == 0, H % 16 == 0, W % 16 == 0: CTA pairs
                               (tcgen05.mma.cta_group::2, M = 256; conv_pair.cu) instead of the single-CTA halo kernel;
                               same results */
  DT_CONV_NO_ROW = 256,     /* 3x3/s1 layers with C_in, C_out <= 64: the 8 x 16 tile kernels (conv_res.cu) instead of the
                               row-streaming kernel whose vertical taps ride in the MMA's N dimension (conv_row.cu, the
                               default where the width is 64 or a multiple of 128); same results */
  DT_CONV_TRANSPOSED = 16   /* data gradient of a stride-2 conv: desc.H, W = size of the OUTPUT (the conv's input), x = gy
                               (N, Ho, Wo, C_in) at the conv's output size, out[h][w] = sum over taps with (h + pad - r)
                               even of gy[(h + pad - r) / 2][..] * w; weights in the dt_conv2d_fwd packing with
                               k = (r*S + s) * C_in + c (dt_pack_conv_weight mode 4) */
};

int dt_conv2d_fwd(const dt_conv_desc* desc, const void* x, const void* skip, const void* w, const float* scale,
                  const float* shift, const void* residual, void* y, dt_stream_t stream);

/* K3: maxpool 3x3 stride 2 pad 1 (torchvision resnet `maxpool`), NHWC, dtype as above. */
int dt_maxpool3x3s2(const void* x, int N, int H, int W, int C, int dtype, void* y, dt_stream_t stream);

/* Stem + maxpool in one launch (bf16 path, zero-bordered input frame, DT_CONV_X_PAD3): y = ReLU(BN(conv7x7/s2(x))) =
 * `encoder.conv1/bn1/relu` and pooled = `encoder.maxpool(y)` of smp's ResNetEncoder.forward (the first two stages of
 * segmodel.py:214 `self.model(x)`); the pooling reads the stem rows from shared memory instead of HBM.  d as for
 * dt_conv2d_fwd.  Returns DT_ERR_UNSUPPORTED unless W/2 == 128 and H/2 % 8 == 0: call dt_conv2d_fwd + dt_maxpool3x3s2 then.
 * Both outputs are bit-identical to those two calls. */
int dt_stem_pool_fwd(const dt_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift, void* y,
                     void* pooled, dt_stream_t stream);

/* ---- K9-K11: segmentation head -------------------------------------------------------------------
 * smp SegmentationHead conv 3x3 (C -> K, bias) fused with the consumers of the logits:
 *   logits_nchw (N, K, H, W) fp32   — what `self.model(img)` returns (segmodel.py:214)
 *   logits_nhwc (N, H, W, K) in `x_dtype` — input of dt_stitch_blend_argmax
 *   mask (N, H, W) uint8            — `out.argmax(dim=1)` (deployment/inference.py:62), first max wins
 * any of the three outputs may be NULL.  w: float [9][C][K], bias: float [K]; K <= 4. */
int dt_head_fwd(const void* x, int x_dtype, int N, int H, int W, int C, int K, const float* w, const float* bias,
                float* logits_nchw, void* logits_nhwc, uint8_t* mask, dt_stream_t stream);

/* The same head on the tensor cores (bf16 path, C == 16, H % 16 == 0, W % 8 == 0): w_packed = bf16 [16][192],
 * row k < K = class k in the dt_conv2d_fwd packing (k = tap*16 + c), rows >= K zero; bias16 = float[16]. */
int dt_head_fwd_tc(const void* x, int N, int H, int W, int K, const void* w_packed, const float* bias16,
                   float* logits_nchw, void* logits_nhwc, uint8_t* mask, dt_stream_t stream);

/* Decoder tail in one launch (bf16 path): decoder.blocks.4.conv2 (3x3, 16 -> 16, folded BatchNorm + ReLU; smp
 * DecoderBlock.conv2, the last Conv2dReLU before `segmentation_head`, segmodel.py:214 via smp.Unet.forward) followed by the
 * head above; the 16-channel intermediate stays in shared memory.  x: bf16 NHWC (N, H, W, 16) = the output of
 * decoder.blocks.4.conv1; w2_packed / wh_packed: bf16 [16][192] in the dt_conv2d_fwd packing; scale2 / shift2: float[16].
 * W must be 128 or 256 (DT_ERR_UNSUPPORTED otherwise: call dt_conv2d_fwd + dt_head_fwd_tc).  Outputs as dt_head_fwd, bit-
 * identical to the two-launch path. */
int dt_tail_fused(const void* x, int N, int H, int W, int K, const void* w2_packed, const float* scale2, const float* shift2,
                  const void* wh_packed, const float* bias16, float* logits_nchw, void* logits_nhwc, uint8_t* mask,
                  dt_stream_t stream);

/* argmax over the class dim of NCHW fp32 logits -> uint8 (first max wins) */
int dt_argmax_nchw(const float* logits, int N, int K, int H, int W, uint8_t* mask, dt_stream_t stream);

/* Majority vote of an ensemble (PyTorchEnsembleInference.run, deadtrees/deployment/inference.py:96-116: torch.mode over
 * the models' argmax masks): masks = M stacked uint8 class-id masks of n pixels each (M * n bytes, M <= 15);
 * out[p] = the most frequent class of pixel p, the smallest one on ties (torch.mode), as int64 (out_int64 != 0, the
 * reference's return type) or uint8. */
int dt_mode_vote(const uint8_t* masks, int M, int64_t n, int out_int64, void* out, dt_stream_t stream);

/* ---- K11: losses and metric ----------------------------------------------------------------------
 * One pass over logits (N, K, H, W) fp32 + labels (N, H, W) int64 computing softmax in registers and
 * the partial sums every reference loss needs (deadtrees/loss/losses.py:226-247 DiceLoss, :273-291
 * FocalLoss, deadtrees/loss/gdl.py:10-27 GeneralizedDiceLoss, smp Fscore at segmodel.py:145-149).
 * sums: double [N][K][4] = { sum p*t, sum p, sum t, sum (1-p)^2 * t * log(p + 1e-10) }
 * counts: int64 [K][3]  = { tp, sum_pr, sum_gt } with pr = (p > 0.5)   (whole batch)
 * bad_label: int32 flag set when a label is outside [0, K) (class2one_hot's assert, losses.py:129).
 * Caller zeroes sums/counts/bad_label. */
int dt_seg_loss_partials(const float* logits, const int64_t* labels, int N, int K, int H, int W, double* sums,
                         int64_t* counts, int32_t* bad_label, dt_stream_t stream);
/* Tiny single-block finalize: turns the partial sums into the reference's scalars and the
 * coefficient tables the backward pass needs.
 * dice_mode: 0 none, 1 DiceLoss(idc = 1..K-1) (losses.py:226-247), 2 GeneralizedDiceLoss (gdl.py:10-27);
 * use_focal: FocalLoss(idc = 0..K-1, gamma = 2) (losses.py:273-291).
 * out: float [8] = { dice_loss, focal_loss, total_loss, fscore_without_bg, fscore_with_bg, 0, 0, 0 }
 *      (total = dice + focal as SemSegment.calculate_loss, segmodel.py:169-200; Fscore eps 1e-7)
 * coef: float [N][K][2] (a, b) with d(dice)/dp[n,k,x] = a*t + b;  focal_scale: float [1] = 1/(sum t + 1e-10)
 * (0 when focal is off). */
int dt_seg_loss_finalize(const double* sums, const int64_t* counts, int N, int K, int dice_mode, int use_focal,
                         float* out, float* coef, float* focal_scale, dt_stream_t stream);
/* d(total_loss)/d(logits): grad_p = a*t + b - focal_scale * t * (-2(1-p)log(p+1e-10) + (1-p)^2/(p+1e-10)),
 * then the softmax Jacobian dz_k = p_k (g_k - sum_j g_j p_j); scaled by `upstream`.  grad_logits (N,K,H,W) fp32. */
int dt_seg_loss_backward(const float* logits, const int64_t* labels, int N, int K, int H, int W, const float* coef,
                         const float* focal_scale, float upstream, float* grad_logits, dt_stream_t stream);

/* Generalized Wasserstein Dice loss, weighting "default" (deadtrees/loss/gwdl.py:84-138; "GWDICE" in SemSegment,
 * segmodel.py:118-124, 176-178).  logits (N, K, H, W) fp32, labels (N, H, W) int64, dist_matrix: HOST float[K*K] class
 * distances (normalised to a maximum of 1 as the module does).  softmax_twice != 0 reproduces the reference call path, which
 * hands the module softmax probabilities that it soft-maxes again.  As in the reference, the "generalised true positives"
 * of a sample sum (1 - W) over ALL samples of the batch (the (B,1,S) x (B,S) broadcast of gwdl.py:187-205).
 * loss_out: float[1]; coef: float[2*N + H*W] backward terms consumed by dt_gwdl_loss_backward, which ADDS
 * weight * d(loss)/d(logits) to grad_logits.  workspace: dt_gwdl_workspace(N, H, W) bytes. */
int64_t dt_gwdl_workspace(int N, int H, int W);
int dt_gwdl_loss(const float* logits, const int64_t* labels, int N, int K, int H, int W, const float* dist_matrix,
                 int softmax_twice, void* workspace, int64_t workspace_bytes, float* loss_out, float* coef,
                 dt_stream_t stream);
int dt_gwdl_loss_backward(const float* logits, const int64_t* labels, int N, int K, int H, int W, const float* dist_matrix,
                          int softmax_twice, const float* coef, float weight, float* grad_logits, dt_stream_t stream);

/* The dataloader's train_transform for one batch (deadtrees/data/deadtreedata.py:132-146 and transform() :156-189):
 * OneOf(HorizontalFlip, VerticalFlip) -> RandomRotate90 -> RandomBrightnessContrast(brightness_by_max=False) -> Normalize ->
 * ToTensorV2, then image[0:out_channels], mask.long(), lu.long() and mask > 1 -> 1 when merge_classes != 0.  The random draws
 * are the caller's: geom int32 [N][2] = {flip (0 none, 1 horizontal, 2 vertical), rot (np.rot90 quarter turns 0..3, applied
 * after the flip)}, bc double [N][2] = {alpha, beta} of the brightness / contrast table
 * lut[v] = uint8(clip(float32(v) * alpha + float32(beta * mean(image)), 0, 255)); {1, 0} leaves the bytes unchanged.
 * images (N, H, W, C) uint8 with H == W, masks / lus (N, H, W) uint8 or NULL; offset / scale: HOST float[out_channels] of the
 * normalisation (u8 - offset) * scale; sums: N uint64 of workspace; out_img (N, out_channels, H, W) fp32, out_mask / out_lu
 * (N, H, W) int64.  All pointers except offset / scale are device memory. */
int dt_train_transform(const uint8_t* images, const uint8_t* masks, const uint8_t* lus, int N, int H, int W, int C,
                       int out_channels, const int32_t* geom, const double* bc, const float* offset, const float* scale,
                       int merge_classes, uint64_t* sums, float* out_img, int64_t* out_mask, int64_t* out_lu,
                       dt_stream_t stream);

/* Confusion matrices of a validation / test epoch (SemSegment.validation_epoch_end / test_epoch_end,
 * deadtrees/network/segmodel.py:291-407: torchmetrics confusion_matrix(prediction, target) and the same on the pixels of
 * the forest mask, `lu == 1`).  pred: n predictions, uint8 (pred_elem 1) or int64 (8); target: n int64 labels; lu: n land-use
 * flags (lu_elem 1, 4 or 8 bytes) or NULL.  counts: int64 [2][K][K], rows = target, columns = prediction; [0] all pixels,
 * [1] the pixels with lu == 1 (zero when lu is NULL).  The call ADDS to counts (zero it before the first batch of an epoch);
 * bad_label: int32 [1], set to 1 when a value lies outside [0, K) (such pixels are not counted).  K <= 16. */
int dt_confusion_matrix(const void* pred, int pred_elem, const int64_t* target, const void* lu, int lu_elem, int64_t n,
                        int K, int64_t* counts, int32_t* bad_label, dt_stream_t stream);

/* The dataloader's signed distance maps for the boundary loss (one_hot2dist, deadtrees/loss/losses.py:159-178, called at
 * deadtrees/data/deadtreedata.py:182-185): labels (N, H, W) int64 -> out (N, K, H, W) fp32 with, for every class k that has
 * a pixel in image n, edt(not k) outside the class and -(edt(k) - 1) inside it (scipy's exact Euclidean distance transform,
 * including its convention for a class that covers the whole image); absent classes stay 0.  truncate != 0: values
 * truncated towards zero as by the reference's assignment into the int32 one-hot dtype; 0: rounded to fp32
 * (dtype=np.float32).  workspace: dt_one_hot2dist_workspace() bytes. */
int64_t dt_one_hot2dist_workspace(int N, int K, int H, int W);
int dt_one_hot2dist(const int64_t* labels, int N, int K, int H, int W, int truncate, float* out, void* workspace,
                    int64_t workspace_bytes, dt_stream_t stream);

/* Boundary (surface) loss on the logits (SurfaceLoss / BoundaryLoss, deadtrees/loss/losses.py:250-270, added to the total in
 * SemSegment.calculate_loss, segmodel.py:188-191): loss = mean over (b, k in idc, h, w) of softmax(logits)_k * dist_k.
 * logits, dist: (N, K, H, W) fp32; idc_mask: bit k set = class k in idc; workspace: 64 * N doubles.
 * dt_boundary_loss_backward ADDS weight * d(loss)/d(logits) to grad_logits (weight = upstream gradient times the ramp
 * factor alpha of BOUNDARY-RAMPED, segmodel.py:157-160). */
int dt_boundary_loss(const float* logits, const float* dist, int N, int K, int H, int W, unsigned idc_mask,
                     double* workspace, float* loss_out, dt_stream_t stream);
int dt_boundary_loss_backward(const float* logits, const float* dist, int N, int K, int H, int W, unsigned idc_mask,
                              float weight, float* grad_logits, dt_stream_t stream);

/* The reference's loss callables take softmax probabilities and an int32 one-hot target
 * (segmodel.py:215-216); these three entry points serve that API:
 * dt_class2one_hot: losses.py:124-141 (int32 scatter; bad_label flag = its label-range assert)
 * dt_softmax_nchw : logits.softmax(dim=1) (segmodel.py:216)
 * dt_prob_loss_partials: per (n, k) plane { sum p*t, sum p, sum t, sum (1-p)^gamma * t * log(p+1e-10) }
 *   into double sums [N][K][4] (caller zeroes); target is int32 one-hot or, with target_is_float,
 *   a float map (the distance maps of SurfaceLoss, losses.py:250-270). */
int dt_class2one_hot(const int64_t* labels, int N, int K, int H, int W, int32_t* onehot, int32_t* bad_label,
                     dt_stream_t stream);
int dt_softmax_nchw(const float* logits, int N, int K, int H, int W, float* probs, dt_stream_t stream);
int dt_prob_loss_partials(const float* probs, const void* target, int target_is_float, int N, int K, int H, int W,
                          float gamma, double* sums, dt_stream_t stream);

/* ---- S1 / K13: training step ----------------------------------------------------------------------
 * Autograd of smp.Unet(resnet34) in train mode as SemSegment.training_step drives it
 * (deadtrees/network/segmodel.py:210-229).  Activations / gradients are NHWC in `dtype`
 * (DT_BF16 or DT_F32); M = N*H*W pixels; per-channel vectors are float[C].
 *
 * dt_bn_train_stats: nn.BatchNorm2d in train mode on the raw conv output y (M, C): batch mean / biased variance
 *   -> scale = gamma*invstd, shift = beta - mean*scale, saved (mean, invstd); running stats updated with
 *   `momentum` and the unbiased variance (running_* may be NULL).  workspace: float[2 * dt_reduce_blocks() * C].
 * dt_bn_apply: out = [relu]( y*scale + shift (+ residual) ).
 * dt_bn_train_bwd: gz = g * [a > 0] (a == NULL: no ReLU); dbeta = sum gz; dgamma = sum gz*yhat;
 *   gy = scale * (gz - dbeta/M - yhat*dgamma/M); optional gz_out (gradient of the identity branch of a
 *   BasicBlock).  workspace: float[2 * dt_reduce_blocks() * C + 2 * C].
 * dt_reduce_blocks: number of partial-sum rows the two calls above use (or DT_ERR_BAD_SHAPE). */
int dt_reduce_blocks(int64_t M, int C, int dtype);
int dt_bn_train_stats(const void* y, int64_t M, int C, int dtype, const float* gamma, const float* beta, float eps,
                      float momentum, float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                      float* invstd, float* workspace, dt_stream_t stream);
int dt_bn_apply(const void* y, int64_t M, int C, int dtype, const float* scale, const float* shift, const void* residual,
                int relu, void* out, dt_stream_t stream);
int dt_bn_train_bwd(const void* g, const void* a, const void* y, int64_t M, int C, int dtype, const float* mean,
                    const float* invstd, const float* scale, float* dgamma, float* dbeta, void* gy, void* gz_out,
                    float* workspace, dt_stream_t stream);
/* dt_bn_train_bwd for a layer whose forward was relu(y*scale + shift) WITHOUT a residual: the ReLU mask is recomputed from
 * y (a > 0 <=> fma(y, scale, shift) > 0, the value dt_bn_apply clamped), so the activation tensor is not read. */
int dt_bn_train_bwd_relu(const void* g, const void* y, int64_t M, int C, int dtype, const float* mean, const float* invstd,
                         const float* scale, const float* shift, float* dgamma, float* dbeta, void* gy, void* gz_out,
                         float* workspace, dt_stream_t stream);
/* out = a + b over n elements (gradient merges) */
int dt_add(const void* a, const void* b, int64_t n, int dtype, void* out, dt_stream_t stream);
/* MaxPool2d(3, 2, 1) backward: gradient goes to the FIRST maximum of each window (ATen semantics);
 * x (N, H, W, C) is the pool input, gout (N, Ho, Wo, C); gx = addend (may be NULL) + pooled gradient. */
int dt_maxpool3x3s2_bwd(const void* x, const void* gout, const void* addend, int N, int H, int W, int C, int dtype,
                        void* gx, dt_stream_t stream);
/* Index form used by the training step (C a multiple of 8 for bf16 / 4 for fp32): the forward also stores, per output
 * element, the window position dy*3+dx (0..8) of its first maximum in idx (N, Ho, Wo, C) uint8; the backward gathers
 * gout through idx without re-reading the pool input.  Same results as dt_maxpool3x3s2 / dt_maxpool3x3s2_bwd. */
int dt_maxpool3x3s2_idx(const void* x, int N, int H, int W, int C, int dtype, void* y, uint8_t* idx, dt_stream_t stream);
int dt_maxpool3x3s2_bwd_idx(const uint8_t* idx, const void* gout, const void* addend, int N, int H, int W, int C, int dtype,
                            void* gx, dt_stream_t stream);
/* out (N, 2Ho, 2Wo, C) = gy (N, Ho, Wo, C) at the even positions, zero elsewhere: the data gradient of a stride-2
 * convolution is dt_conv2d_fwd (stride 1) of this tensor with the dt_pack_conv_weight mode-3 weights. */
int dt_zero_insert2x(const void* gy, int N, int Ho, int Wo, int C, int dtype, void* out, dt_stream_t stream);
/* cat([nearest_x2(x_low), skip], C) materialised for the training path, and its backward
 * (g_x_low = 2x2 block sums of g_cat[..., :Cx]; g_skip = g_cat[..., Cx:]).  H, W: full resolution. */
int dt_upsample_concat(const void* x_low, const void* skip, int N, int H, int W, int Cx, int Cs, int dtype, void* out,
                       dt_stream_t stream);
int dt_upsample_concat_bwd(const void* g_cat, int N, int H, int W, int Cx, int Cs, int dtype, void* g_x_low, void* g_skip,
                           dt_stream_t stream);
/* (N, K, H, W) fp32 -> (N, H, W, Kp) `dtype`, channels >= K zero (gradient of the logits into the head backward) */
int dt_nchw_to_nhwc(const float* x, int N, int K, int H, int W, int Kp, int dtype, void* out, dt_stream_t stream);
/* out[k] = sum over the M pixels of g[pixel][k] for k < K <= 4 (g: (M, C) NHWC rows) — the bias gradient of the head.
 * Two fixed-order stages (no atomics): the same bits on every run.  workspace: 2048 floats. */
int dt_channel_sum(const void* g, int64_t M, int C, int K, int dtype, float* workspace, float* out, dt_stream_t stream);
/* fp32 OIHW master weights -> kernel layouts.  mode 0: float [tap][C_in_p][C_out]; 1: bf16 [C_out][Kpad]
 * (dt_conv2d_fwd); 2: bf16 stem packing [C_out][256]; 3: bf16 [C_in][Kpad], k = (R*S-1-tap)*Cop + co — the
 * weights with which dt_conv2d_fwd computes the data gradient of a stride-1 convolution; 4: the same without the tap
 * flip, k = tap*Cop + co (DT_CONV_TRANSPOSED).  Modes 3 / 4 take Cop (>= C_out, channel stride of gy) in `C_in_p`. */
int dt_pack_conv_weight(const float* w_oihw, int C_out, int C_in, int R, int S, int mode, int C_in_p, int Kpad, void* out,
                        dt_stream_t stream);
/* The same for many tensors in one launch.  jobs_device: DEVICE array of njobs records sorted by `start` (the running
 * sum of the output element counts, jobs[0].start = 0); total = sum of all output element counts.  Arguments per
 * record as dt_pack_conv_weight (not validated here: build the records from calls that passed its checks). */
typedef struct dt_pack_job {
  const float* w;
  void* out;
  int32_t C_out, C_in, R, S, mode, C_in_p, Kpad, reserved;
  int64_t start;
} dt_pack_job;
int dt_pack_conv_weights_batched(const dt_pack_job* jobs_device, int njobs, int64_t total, dt_stream_t stream);
/* Generic (CUDA-core) data / weight gradients of conv2d, any stride: fp32 check mode and the layer shapes the
 * tcgen05 kernels do not cover.  x / gx: (N, H, W, C_x) with C_in <= C_x real channels; gy: (N, Ho, Wo, C_out);
 * weights and dw: fp32 OIHW (C_out, C_in, R, S); gx = addend (may be NULL) + dgrad; dbias (may be NULL): float[C_out];
 * round_weights (bf16 tensors only): use the weights rounded to bf16, as the tensor-core kernels do. */
int dt_conv2d_dgrad_direct(const void* gy, const float* w_oihw, const void* addend, int N, int H, int W, int C_in, int C_x,
                           int C_out, int R, int S, int stride, int pad, int dtype, int round_weights, void* gx,
                           dt_stream_t stream);
int dt_conv2d_wgrad_direct(const void* x, const void* gy, int N, int H, int W, int C_in, int C_x, int C_out, int R, int S,
                           int stride, int pad, int dtype, float* dw_oihw, float* dbias, dt_stream_t stream);

/* Weight gradient on the tensor cores (tcgen05, MN-major operands straight from the NHWC tensors, fp32 accumulation in
 * TMEM, deterministic split reduction): 3x3 / pad 1 with stride 1 or 2, and 1x1 / stride 2 / pad 0.
 * x (N, stride*Ho, stride*Wo, x_cstride) and gy (N, Ho, Wo, gy_cstride) bf16 with C_in <= x_cstride, C_out <= gy_cstride real
 * channels (channels beyond them must be zero or are ignored); dw fp32 OIHW (C_out, C_in, k, k) overwritten.
 * Needs Wo % 8 == 0, Ho % 16 == 0 (or Ho == 8 with N even), C_in % 4 == 0, channel strides multiples of 8;
 * returns DT_ERR_UNSUPPORTED otherwise (use the direct kernel).  workspace: dt_conv2d_wgrad_tc_workspace() bytes
 * (that function returns DT_ERR_UNSUPPORTED for unsupported shapes). */
int64_t dt_conv2d_wgrad_tc_workspace(int N, int Ho, int Wo, int C_in, int C_out, int ksize, int stride);
int dt_conv2d_wgrad_tc(const void* x, const void* gy, int N, int Ho, int Wo, int C_in, int x_cstride, int C_out,
                       int gy_cstride, int ksize, int stride, float* dw_oihw, float* workspace, int64_t workspace_bytes,
                       dt_stream_t stream);

/* Weight gradient of the 7x7 / stride-2 / pad-3 stem (encoder.conv1) on the tensor cores.  x_frame is the zero-bordered
 * bf16 frame (N, H+6, W+8, 4) the forward stem reads (DT_CONV_X_PAD3); the SAME overlapping-stride im2col tensor map is
 * the B operand, gy (N, H/2, W/2, 64) bf16 the A operand (both MN-major, K = output pixels); seven 64 x 32 fp32
 * accumulators (one per filter row) stay in TMEM over all tiles of a CTA, per-CTA partials are reduced in a fixed order
 * (deterministic).  dw fp32 OIHW (64, C_in, 7, 7) overwritten, C_in <= 4.  DT_ERR_UNSUPPORTED when the output grid does
 * not tile into 128-pixel boxes (as dt_conv2d_fwd's stem path). */
int64_t dt_stem_wgrad_tc_workspace(int N, int H, int W);
int dt_stem_wgrad_tc(const void* x_frame, const void* gy, int N, int H, int W, int C_in, float* dw_oihw, float* workspace,
                     int64_t workspace_bytes, dt_stream_t stream);

/* ---- O1: optimizer ------------------------------------------------------------------------------
 * torch.optim.Adam step (segmodel.py:420-425 defaults) with the Lightning global-norm clip
 * (configs/trainer/default.yaml:18) folded in: g *= min(1, max_norm / (norm + 1e-6)).
 * sumsq: device double[1] holding the sum of squares of all grads (dt_sumsq accumulates into it;
 * the caller zeroes it); max_norm <= 0 disables clipping (sumsq may then be NULL). */
int dt_sumsq(const float* g, int64_t n, double* sumsq, dt_stream_t stream);
int dt_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                 float eps, int step, const double* sumsq, float max_norm, dt_stream_t stream);

/* The same update with every piece of step state on the device, so the launch can be captured in a CUDA graph:
 * state = float[2] {step count, learning rate} (step incremented here; the caller updates lr between replays);
 * loss (may be NULL): the step is skipped - parameters, moments and the step count untouched - when *loss is not finite
 * (SemSegment.training_step returns None for a NaN / Inf loss, segmodel.py:219-221); scratch4: float[4] work area. */
int dt_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float* state, float beta1, float beta2,
                     float eps, const double* sumsq, float max_norm, const float* loss, float* scratch4,
                     dt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DEADTREES_B200_H_ */
