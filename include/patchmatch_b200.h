/*
 * patchmatch_b200.h -- C ABI of the B200-native (sm_100a) learned-PatchMatch hot path.
 *
 * This is the drop-in boundary: plain device pointers, sizes and a CUDA stream;
 * no torch types.  The reference (FangjinhuaWang/PatchmatchNet) has no native
 * layer at all -- its hot path is a chain of ATen calls inside
 * models/patchmatch.py / models/module.py -- so each entry point below names the
 * reference Python lines it replaces; INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all tensors are fp32, dense, in the layout written beside them;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     kernels are enqueued, never synchronised -- the calls are CUDA-graph capturable;
 *   - return value: 0 on success, negative PMB200_E* on a rejected argument,
 *     positive cudaError_t if the launch failed.  pmb200_last_error() returns a
 *     thread-local message for the last non-zero return.  Nothing aborts.
 *   - re-entrant: no global mutable state; safe to call from one host thread per
 *     GPU (the reference's nn.DataParallel threading model, train.py:282).
 *     The caller makes the right device current (cudaSetDevice / torch.cuda.device).
 */
#ifndef PATCHMATCH_B200_H_
#define PATCHMATCH_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMB200_ABI_VERSION 1

#define PMB200_EINVAL (-1)       /* bad size / null pointer */
#define PMB200_EUNSUPPORTED (-2) /* combination the reference raises NotImplementedError for */
#define PMB200_EIO (-3)          /* map files: open / read / write failed (message carries strerror and the path) */
#define PMB200_EFORMAT (-4)      /* map files: not a PFM file / malformed header / payload does not match the header */

#define PMB200_MAX_VIEWS 16
#define PMB200_MAX_NEIGHBORS 32
#define PMB200_MAX_HYPOTHESES 256

int pmb200_abi_version(void);
const char *pmb200_last_error(void);

/* Measurement aid (tools/kbench.py, A/B tests): override a launch-configuration knob of the library for this process.
 * Keys: "ka_gen" (3 | 4: generation of the fused warp+correlation kernel), "ka3_dc", "ka3_dc_vw", "ka3_pipe", "ka3_minb"
 * (generation-3 rows per pass / gather pipeline / resident CTAs), "ka4_nw" (4 | 8 consumer warps = tile rows), "ka4_ctas"
 * (resident CTAs per SM the window ring is sized for), "ka4_cap" (texels per ring slot), "ka4_stages" (ring depth 2..4), "ka4_grid" (persistent CTAs),
 * "kb_tp", "kb_dy" (block shape of the adaptive-evaluation kernel), "stem_ppt" (output pixels per thread of the fused
 * conv0 -> conv1 kernel: 2 or 4); 0 = the built-in default.  "reset" restores every
 * default.  Results never depend on these knobs, only launch shapes do.  Returns 0, or PMB200_EINVAL for an unknown key.
 * The reference has no counterpart (its launch shapes are ATen's). */
int pmb200_set_tuning(const char *key, int value);

/* ------------------------------------------------------------------------------------
 * K-D5: channels-last convolution on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), fp32-accurate (3xTF32 split).
 * Replaces, for the FLOP-bound layers in the fp32-accurate mode, the nn.Conv2d calls of reference models/net.py:9-70 (FeatureNet
 * conv2..10), models/patchmatch.py:288-311 (stage-2/3 offset convs), models/net.py:73-122 (Refinement).
 *   x_nhwc      [N,H,W,Cin] fp32, 16-byte aligned, Cin in {8,16,32,64}
 *   filter_tc5  pmb200_conv2d_tc5_filter_floats(Cin,Cout,KS) floats: [tap ky*KS+kx][Cin/32 blocks (1 if Cin <= 32)]
 *               [2 Npad rows: w_hi rows then w_lo rows, Npad = Cout rounded up to 16][min(Cin,32) channels]; every
 *               [2 Npad][row bytes] tile is ONE K-major tcgen05 operand (its first Npad rows double as the N = Npad operand)
 *               stored in the tensor core's shared-memory swizzle for that row size (16-byte chunk j of row r sits at chunk
 *               j ^ (r % 8) for 128-byte rows, j ^ ((r / 2) % 4) for 64-byte rows, j ^ ((r / 4) % 2) for 32-byte rows);
 *               w_hi = weight rounded to TF32, w_lo = (weight - w_hi) rounded to TF32; rows >= Cout of each half are zero
 *   y_nhwc      [N,Ho,Wo,y_channel_stride]; channels y_channel_offset .. +Cout are written
 * Ho = (H + 2 pad - dil (KS-1) - 1) / stride + 1.  bias may be NULL.  relu != 0 applies max(.,0).
 * pmb200_conv2d_tc5_supported: 1 when (Cin, Cout, KS, stride) is served (KS in {1,3,5}, stride in {1,2}, Cout <= 64). */
int pmb200_conv2d_tc5_supported(int Cin, int Cout, int KS, int stride);
int pmb200_conv2d_tc5_filter_floats(int Cin, int Cout, int KS);
int pmb200_conv2d_tc5(const float *x_nhwc, const float *filter_tc5, const float *bias, float *y_nhwc, int N, int H, int W, int Cin,
                      int Cout, int KS, int stride, int pad, int dil, int relu, int y_channel_stride, int y_channel_offset,
                      void *stream);
/* Stride-1 form of the same convolution (K-D5h): the input halo of an 8 x 16 output tile is fetched ONCE (one dense TMA box),
 * split once into channel-chunk planes [Cin/4][rows][cols][4 floats]; every filter tap is a shifted view of them (no-swizzle
 * K-major tcgen05 operand: 8 consecutive pixels of a plane are one core matrix).  Same arguments minus the stride.
 *   filter_tc5h  same float count as filter_tc5, layout [tap][Cin/4 chunks][w_hi rows Npad | w_lo rows Npad][4 floats] */
int pmb200_conv2d_tc5h(const float *x_nhwc, const float *filter_tc5h, const float *bias, float *y_nhwc, int N, int H, int W, int Cin,
                       int Cout, int KS, int pad, int dil, int relu, int y_channel_stride, int y_channel_offset, void *stream);
/* Debugging aid (not a reference interface): a device buffer of 256 int64 that following K-D5h launches stamp with clock64
 * for the four roles of CTA 0 ([role][tile 0..15][event 0..3]); NULL switches it off. */
int pmb200_debug_conv5h_trace(long long *device_buffer_256);

/* ------------------------------------------------------------------------------------
 * K-S: the two full-resolution layers of FeatureNet in one launch, exact fp32 (plain FFMA, weights in the constant bank):
 *     y = relu(conv1(relu(conv0(x))))   conv0 3 -> 8, conv1 8 -> 8, both 3x3 / pad 1, BatchNorm folded into weight and bias.
 * Replaces `self.conv1(self.conv0(x))` at reference models/net.py:44 (layers :18-19) in eval mode.
 *   x_nchw   device [N,3,H,W] contiguous (the caller's image planes are read in place: no layout conversion)
 *   y_nhwc   device [N,H,W,8], 16-byte aligned
 *   host_w0  HOST memory [8][3][3][3] (PyTorch layout), host_b0 [8]; host_w1 HOST memory [8][8][3][3], host_b1 [8]:
 *            copied into the kernel parameter block at launch (so a captured CUDA graph keeps the values it was captured with). */
int pmb200_conv_stem(const float *x_nchw, const float *host_w0, const float *host_b0, const float *host_w1, const float *host_b1,
                     float *y_nhwc, int N, int H, int W, void *stream);

/* ------------------------------------------------------------------------------------
 * K-R: Refinement (reference models/net.py:73-122) in two launches, exact fp32 (plain FFMA, weights in the constant bank).
 * pmb200_refine_low  (half resolution; :104-110):  d = (depth - depth_min) / (depth_max - depth_min);
 *                    low = relu(conv2(relu(conv1(d))))   conv1 1 -> 8, conv2 8 -> 8, 3x3 / pad 1, BatchNorm folded.
 * pmb200_refine_full (full resolution; :103, :112-120):  up = relu(bn(deconv(low))), c0 = relu(conv0(img)),
 *                    c3 = relu(conv3(cat(up, c0))), res = conv_res(c3), out = (nearest_up2x(d) + res) * span + depth_min.
 *   depth_half [N,1,h,w], depth_min / depth_max [N], low_nhwc [N,h,w,8] (16-byte aligned), img_nchw [N,3,2h,2w] contiguous,
 *   depth_out [N,1,2h,2w]: device.   Every host_* argument is HOST memory in PyTorch layout with BatchNorm folded:
 *   host_w1 [8][1][3][3], host_w2 [8][8][3][3], host_wd [8 in][8 out][3][3] (ConvTranspose2d), host_w0 [8][3][3][3],
 *   host_w3 [8][16][3][3], host_wr [1][8][3][3] (no bias), biases [8].  They travel in the kernel parameter block. */
int pmb200_refine_low(const float *depth_half, const float *depth_min, const float *depth_max, const float *host_w1, const float *host_b1,
                      const float *host_w2, const float *host_b2, float *low_nhwc, int N, int h, int w, void *stream);
int pmb200_refine_full(const float *low_nhwc, const float *img_nchw, const float *depth_half, const float *depth_min, const float *depth_max,
                       const float *host_wd, const float *host_bd, const float *host_w0, const float *host_b0, const float *host_w3,
                       const float *host_b3, const float *host_wr, float *depth_out, int N, int H, int W, void *stream);

/* ------------------------------------------------------------------------------------
 * Relative projections for every (source view, batch element):
 *     P = src_proj . inverse(ref_proj);  rt = [P[0,0..2], P[1,0..2], P[2,0..2], P[0..2,3]]
 * Replaces torch.matmul(src_proj, torch.inverse(ref_proj)) at models/module.py:148-150,
 * which the reference re-evaluates (with a host sync inside linalg.inv) once per source
 * view per PatchMatch iteration; here it runs once per stage with no host sync.
 * The 4x4 inverse is done in fp64 on the device and rounded to fp32.
 *   ref_proj        [B,4,4], row stride 4, batch stride `ref_batch_stride` floats
 *   src_projs_host  host array of V device pointers, each [B,4,4] with batch stride `src_batch_stride`
 *   rt_out          [V,B,12]
 */
int pmb200_relative_projection(const float *ref_proj, int64_t ref_batch_stride,
                               const float *const *src_projs_host, int64_t src_batch_stride,
                               int V, int B, float *rt_out, void *stream);

/* ------------------------------------------------------------------------------------
 * Pack n NCHW feature maps into one channels-last buffer [n,B,H,W,C] (one launch).
 * The fused kernels read features channels-last so that each bilinear tap is one
 * contiguous C-vector.  Not needed when the producer already emits channels-last.
 *   maps_host  host array of n device pointers, each [B,C,H,W] contiguous
 */
int pmb200_pack_nhwc(const float *const *maps_host, int n, int B, int C, int H, int W,
                     float *out_nhwc, void *stream);

/* ------------------------------------------------------------------------------------
 * Caller-side helper (feature pyramid top-down path, models/net.py:60-66): channels-last
 *     out = bilinear_upsample_x2(x) + y (+ bias[c])     x [N,h,w,C], y/out [N,2h,2w,C], C % 4 == 0
 * same sampling as F.interpolate(scale_factor=2, mode="bilinear", align_corners=False).
 * bias: NULL, or the [C] bias of the lateral 1x1 conv that produced y (so that conv can run bias-free).
 * Replaces ATen's channels-last bilinear kernel + separate bias/add passes (12 % + 4 % of the forward before).
 */
int pmb200_upsample2x_add_nhwc(const float *x_nhwc, const float *y_nhwc, const float *bias, float *out_nhwc,
                               int N, int h, int w, int C, void *stream);

/* ------------------------------------------------------------------------------------
 * Channels-last 2-D convolution for the small learned convs of the model (1..64 channels), tensor cores
 * (TF32 implicit GEMM), bias + ReLU fused:
 *     y[n,oy,ox,yco+co] = act(bias[co] + sum_{ky,kx,ci} x[n, oy*S-pad+ky*dil, ox*S-pad+kx*dil, ci] * w[co,ci,ky,kx])
 * Replaces, on the hot path, the offset convs propa_conv / eval_conv (models/patchmatch.py:288-311, called at
 * :486 and :498: nn.Conv2d 3x3, dilation = padding = propagation range) and, either side of it (SURVEY.md 8f rows
 * f1 / f3), ConvBnReLU (models/module.py:11-40, BatchNorm folded by the caller) and the plain convs of FeatureNet
 * (models/net.py:9-70) and Refinement (models/net.py:73-122, including its ConvTranspose2d: transposed2x = 1 runs the
 * equivalent stride-1 conv over the virtually zero-stuffed input, caller passes the flipped filter and pad = KS-1-pad_t).
 *   x            [N,H,W,Cin]      (16-byte aligned when Cin % 4 == 0)
 *   filter_frag  the filter in tensor-core fragment order for the chosen precision,
 *                pmb200_conv2d_filter_floats(Cin,Cout,KS,precision) floats, 16-byte aligned:
 *                [tap = ky*KS+kx][ks = 0..Cin'/8)[nt = 0..NT)[lane = 0..32)[2 or 4], lane = 4*g + t holding
 *                (b0, b1) = (w[nt*8+g][ks*8+2t][ky][kx], w[nt*8+g][ks*8+2t+1][ky][kx]) -- the MMA's k slots (t, t+4)
 *                are mapped to the adjacent channels (2t, 2t+1) -- zero where an index is outside the filter;
 *                Cin' = Cin rounded up to 8/16/32/64, NT = ceil(Cout/8) rounded up to 1,2,3,4 or 8.
 *                precision 1: the two values rounded to TF32 (nearest, ties away from zero);
 *                precision 3: four values (b0_hi, b1_hi, b0_lo, b1_lo), hi = tf32(w), lo = tf32(w - hi)
 *   bias         [Cout] or NULL
 *   add_up2x     NULL, or a coarser map [N,Ho/2,Wo/2,Cout] whose bilinear x2 upsample (F.interpolate scale_factor=2,
 *                align_corners=False) is added before the activation: the top-down step of the feature pyramid
 *                (models/net.py:60-66) fused into the lateral conv's epilogue (Ho, Wo, Cout even)
 *   y            [N,Ho,Wo,y_channel_stride], written at channels y_channel_offset .. +Cout (stride 0 = Cout: dense)
 *   precision    1: TF32 operands (the library's behaviour under torch.backends.cudnn.allow_tf32, torch's default:
 *                filter rounded by the host, activations truncated by the tensor core as for tcgen05 kind::tf32);
 *                3: error-compensated 3xTF32, fp32-accurate
 *   rows_per_warp 0 = choose (tile = 16 columns x 4*rows_per_warp rows per 4-warp CTA); 1, 2 or 4 to force
 */
int pmb200_conv2d_filter_floats(int Cin, int Cout, int KS, int precision);
int pmb200_conv2d_nhwc(const float *x, const float *filter_frag, const float *bias, const float *add_up2x, float *y,
                       int N, int H, int W, int Cin, int Cout, int KS, int stride, int pad, int dil,
                       int relu, int precision, int transposed2x, int y_channel_stride, int y_channel_offset,
                       int rows_per_warp, void *stream);

/* ------------------------------------------------------------------------------------
 * Geometric-consistency filtering of one reference depth map against V source depth maps, one launch
 * (SURVEY.md 8f row f4).  Replaces reproject_with_depth (eval.py:86-146), check_geometric_consistency (eval.py:149-190)
 * and the per-reference-view accumulation of filter_depth (eval.py:220, :226-256): numpy + cv2.remap on one CPU thread.
 *   ref_depth [H,W], confidence [H,W], src_depths [V,Hs,Ws]   float32
 *   cams [V,60] float64 (device): per source view, row-major, composed exactly as the reference composes them in float32
 *       and then widened:  inv(K_ref) (9) | (E_src . inv(E_ref))[:3,:4] (12) | K_src (9) | inv(K_src) (9) |
 *       (E_ref . inv(E_src))[:3,:4] (12) | K_ref (9)        (eval.py:116-139)
 *   geo_mask_sum_out [H,W] int32: source views consistent with the pixel (dist < geo_pixel_thres and relative depth
 *       difference < geo_depth_thres, eval.py:180-187); photo_mask_out / final_mask_out [H,W] uint8 (0/1):
 *       confidence > photo_thres, and that AND geo_mask_sum >= geo_mask_thres (eval.py:220, :254-255);
 *   depth_avg_out [H,W] float64: (sum of the consistent reprojected depths + ref_depth) / (geo_mask_sum + 1) (eval.py:252).
 * The source-depth sample reproduces cv2.remap(INTER_LINEAR, constant border 0): coordinates rounded to 1/32 pixel,
 * float32 weights, no fused multiply-add. */
int pmb200_geometric_filter(const float *ref_depth, const float *confidence, const float *src_depths,
                            const double *cams, int V, int H, int W, int Hs, int Ws, double geo_pixel_thres,
                            float geo_depth_thres, float photo_thres, int geo_mask_thres, int *geo_mask_sum_out,
                            unsigned char *photo_mask_out, unsigned char *final_mask_out, double *depth_avg_out,
                            void *stream);

/* Point-cloud half of the fusion for one reference view (reference eval.py:273-296): the pixels of final_mask, in row-major
 * order, back-projected with the averaged depth into the world frame, with their colours, written as the 15-byte vertex
 * records of the reference's fused.ply (binary little endian: float32 x, y, z, uint8 red, green, blue).
 *   final_mask  [H*W] bytes (0 / non-0), depth_avg [H*W] float64 (both as pmb200_geometric_filter writes them)
 *   ref_img_hwc [H*W*3] float32 in [0,1] (the reference's read_image output)
 *   cam25       25 doubles: inverse(ref_intrinsics) row-major (9) then inverse(ref_extrinsics) row-major (16), the float32
 *               inverses numpy computes, widened (ops.compose_fusion_camera)
 *   ply_body_out  capacity H*W*15 bytes; count_out: ONE int, the number of vertices written; block_scratch: ceil(H*W/256) ints
 * All pointers are device memory.  Three launches on `stream` (count, scan, write); no host synchronisation. */
int pmb200_fuse_points(const unsigned char *final_mask, const double *depth_avg, const float *ref_img_hwc, const double *cam25,
                       int H, int W, unsigned char *ply_body_out, int *count_out, int *block_scratch, void *stream);

/* ------------------------------------------------------------------------------------
 * Depth / confidence map files either side of the path: PFM and COLMAP .bin, read straight into /
 * written straight from caller-owned host buffers (pinned, so the next step is one cudaMemcpyAsync).  Host code, no
 * stream.  Replaces reference datasets/data_io.py read_pfm :257-288, save_pfm :291-322, read_bin :165-191,
 * save_bin :194-223 -- identical bytes on disk, identical values in memory.  In-memory layout is [H,W,C] row-major, top
 * row first (what read_pfm / read_bin return); the PFM bottom-up row order and the planar [C][H][W] payload of .bin are
 * handled inside the I/O.  Messages of PMB200_EFORMAT are the reference's exception texts ("Not a PFM file.",
 * "Malformed PFM header.", numpy's "cannot reshape array of size N into shape (..)"). */
#define PMB200_MAP_PFM 1
#define PMB200_MAP_COLMAP_BIN 2

typedef struct {
    int format;             /* PMB200_MAP_* */
    int width, height, channels;
    int big_endian;         /* PFM: non-negative scale line (data_io.py:275-279); .bin: 0 */
    double scale;           /* PFM: |scale line| (data_io.py:277); .bin: 1 */
    int64_t data_offset;    /* first payload byte */
    int64_t payload_floats; /* whole float32 items after the header (np.fromfile) */
} pmb200_map_info;

/* Parse the header only (to size the destination buffer). */
int pmb200_map_probe(const char *path, int format, pmb200_map_info *info);
/* Read the whole map into out_host[H*W*C]; `info_out` may be NULL.  capacity_floats < H*W*C -> PMB200_EINVAL. */
int pmb200_map_read(const char *path, int format, float *out_host, int64_t capacity_floats, pmb200_map_info *info_out);
/* Write data_host[H,W,C] (C = 1 or 3).  PFM: the scale line is printf("%f", -scale) -- the reference negates the scale on a
 * little-endian host (data_io.py:315-318); .bin ignores `scale`. */
int pmb200_map_write(const char *path, int format, const float *data_host, int height, int width, int channels, double scale);

/* Caller-side helper (models/net.py:289-299): photometric confidence = probability mass of the four
 * hypotheses around the regressed hypothesis index, nearest-resized to [H_out, W_out].
 *   prob [B,D,h,w] (the last PatchMatch stage's probabilities)   confidence_out [B,H_out,W_out] */
int pmb200_photometric_confidence(const float *prob, float *confidence_out,
                                  int B, int D, int h, int w, int H_out, int W_out, void *stream);

/* ------------------------------------------------------------------------------------
 * K-A: fused homography warp + bilinear gather + group-wise correlation
 *      (+ view-weighted aggregation).
 * Replaces, per source view, differentiable_warping (models/module.py:130-181), the
 * broadcast multiply + mean at models/patchmatch.py:199-203 and -- when view_weights is
 * given -- the accumulation/normalisation at models/patchmatch.py:192-194,213-217.
 * The [B,C,D,H,W] warped tensor is never materialised.
 *   ref_nhwc      [B,H,W,C]
 *   src_nhwc      [V,B,Hs,Ws,C]
 *   rt            [V,B,12] from pmb200_relative_projection
 *   depth         [B,D,H,W]
 *   view_weights  [B,V,H,W] or NULL
 *   out           view_weights == NULL : per-view similarity [V,B,G,D,H,W]
 *                 view_weights != NULL : sum_v(sim_v * w_v) / (1e-5 + sum_v w_v)  [B,G,D,H,W]
 * Fast path for (C,G) in {(64,8),(32,8),(16,4)}; any other C % G == 0 runs a generic kernel.
 */
int pmb200_warp_corr(const float *ref_nhwc, const float *src_nhwc, const float *rt,
                     const float *depth, const float *view_weights, float *out,
                     int V, int B, int C, int G, int H, int W, int Hs, int Ws, int D,
                     void *stream);

/* ------------------------------------------------------------------------------------
 * View-weighted aggregation of stored per-view similarities (first iteration on the
 * coarsest stage, where the weights come from PixelwiseNet run on those similarities):
 *     out = sum_v(sims[v] * w[:,v]) / (1e-5 + sum_v w[:,v])
 * Replaces models/patchmatch.py:192-194,213-217.
 *   sims [V,B,G,D,H,W]   view_weights [B,V,H,W]   out [B,G,D,H,W]
 */
int pmb200_aggregate_views(const float *sims, const float *view_weights, float *out,
                           int V, int B, int G, int D, int H, int W, void *stream);

/* ------------------------------------------------------------------------------------
 * K-A': reference-feature self-correlation at the learned evaluation neighbours.
 * Replaces PatchMatch.get_grid(evaluation) (models/patchmatch.py:361-426) and the
 * grid_sample + multiply + mean of FeatureWeightNet.forward (models/patchmatch.py:613-622).
 * Neighbour k samples at (x + dx_k + offsets[2k], y + dy_k + offsets[2k+1]) with the
 * reference's normalise(align_corners=True)/sample(align_corners=False) mismatch and
 * border padding reproduced.
 *   offsets  [B,2K,H,W] raw output of eval_conv;  K in {9,17};  dilation = propagation range
 *   out      [B,G,K,H,W]
 */
int pmb200_offset_corr(const float *ref_nhwc, const float *offsets, int offsets_channels_last, float *out,
                       int B, int C, int G, int H, int W, int K, int dilation, void *stream);
/* offsets_channels_last (here and below): 0 = planar [B,2K,H,W]; 1 = channels-last [B,H,W,2K], i.e. the memory a
 * conv2d on channels-last input produces -- consumed in place, no re-layout copy. */

/* ------------------------------------------------------------------------------------
 * Eval-mode fusion of the learned 1x1x1 heads (SURVEY.md 8f "f2").
 * A head is conv3d(G->16) + BN + ReLU, conv3d(16->8) + BN + ReLU, conv3d(8->1) + bias
 * (ConvBnReLU3D, models/module.py:43-72).  With BatchNorm in eval mode the BN folds into the
 * conv; the caller passes the folded weights (HOST pointer, copied into the kernel parameter
 * block, i.e. constant memory) and the kernels apply the head in their epilogue, so the
 * [B,G,D,H,W] similarity tensor (62-73 % of K-A's bytes) is never written.
 * Inference only: training-mode BatchNorm needs batch statistics and keeps the unfused path.
 */
typedef struct pmb200_mlp {
    float w0[16 * 8]; /* [16][G], rows padded to 8 inputs */
    float b0[16];
    float w1[8 * 16]; /* [8][16] */
    float b1[8];
    float w2[8];
    float b2;
} pmb200_mlp;

/* K-A + SimilarityNet head (models/patchmatch.py:547-549,570): weighted view average -> MLP.
 *   score_out [B,D,H,W];  view_weights [B,V,H,W] required.  (C,G) in {(64,8),(32,8),(16,4)}. */
int pmb200_warp_corr_score(const float *ref_nhwc, const float *src_nhwc, const float *rt,
                           const float *depth, const float *view_weights,
                           const pmb200_mlp *head_host, float *score_out, int score_stride,
                           int V, int B, int C, int G, int H, int W, int Hs, int Ws, int D,
                           void *stream);
/* score_stride: element stride of score_out.  1 = dense [B,D,H,W]; 2 = write the .y lanes of an interleaved
 * (xnorm, score) buffer [B,D,H,W,2] whose .x lanes pmb200_init_propagate filled (xnorm_stride = 2), which
 * pmb200_adaptive_eval then gathers with ONE 8-byte load per tap (xnorm_score argument). */

/* K-A + PixelwiseNet (models/patchmatch.py:690-702): per view, max over hypotheses of
 * sigmoid(MLP(similarity)).  view_weights_out [B,V,H,W] (zeroed by the call, then atomic max).
 * sims_out: NULL, or [V,B,G,D,H,W] to also keep the per-view similarities so that the weighted
 * aggregation that follows (pmb200_aggregate_views_score) does not have to recompute them. */
int pmb200_warp_corr_view_weights(const float *ref_nhwc, const float *src_nhwc, const float *rt,
                                  const float *depth, const pmb200_mlp *head_host,
                                  float *view_weights_out, float *sims_out,
                                  int V, int B, int C, int G, int H, int W, int Hs, int Ws, int D,
                                  void *stream);

/* View-weighted aggregation of stored per-view similarities + SimilarityNet head:
 *   score_out [B,D,H,W] = MLP( sum_v sims[v]*w[:,v] / (1e-5 + sum_v w[:,v]) ).  G in {4,8}. */
int pmb200_aggregate_views_score(const float *sims, const float *view_weights,
                                 const pmb200_mlp *head_host, float *score_out, int score_stride,
                                 int V, int B, int G, int D, int H, int W, void *stream);

/* K-A' + FeatureWeightNet head (models/patchmatch.py:597-601,624): sigmoid(MLP(correlation)).
 *   weight_out [B,K,H,W] */
int pmb200_offset_corr_weight(const float *ref_nhwc, const float *offsets, int offsets_channels_last,
                              const pmb200_mlp *head_host, float *weight_out,
                              int B, int C, int G, int H, int W, int K, int dilation, void *stream);

/* ------------------------------------------------------------------------------------
 * K-C: hypothesis initialisation + adaptive propagation + sort, one launch.
 * Replaces DepthInitialization.forward (models/patchmatch.py:53-94),
 * PatchMatch.get_grid(propagation) (:331-360,:396-426) and Propagation.forward (:115-124).
 *   mode 0  random init : `seed_map` is U[0,1) noise [B,48,H,W] (the caller draws it with
 *                          torch.rand so the generator stream matches the reference); Ns = 48
 *   mode 1  perturbation: `seed_map` is the current depth [B,1,H,W]; Ns samples at
 *                          inverse-depth steps (1/dmin - 1/dmax) * interval_scale, clamped
 *   mode 2  passthrough : Ns == 1, hypotheses = current depth
 *   offsets [B,2Kp,H,W] raw output of propa_conv, Kp in {0,4,8,16}; with Kp == 0 nothing is
 *           propagated and the samples keep initialisation order (no sort), as in the reference.
 *   out       [B,Ns+Kp,H,W]  (ascending along dim 1 when Kp > 0)
 *   xnorm_out [B,Ns+Kp,H,W] or NULL: (1/out - 1/dmax) / (1/dmin - 1/dmax), the normalised inverse
 *             depth that depth_weight gathers (models/patchmatch.py:655-657), for pmb200_adaptive_eval
 */
int pmb200_init_propagate(const float *seed_map, const float *offsets, int offsets_channels_last,
                          const float *depth_min, const float *depth_max, float *out, float *xnorm_out,
                          int xnorm_stride, int mode, int B, int H, int W, int Ns, int Kp, int dilation,
                          float interval_scale, void *stream);

/* ------------------------------------------------------------------------------------
 * K-B: adaptive evaluation tail, one launch.
 * Replaces depth_weight (models/patchmatch.py:650-669), the feature-weight product and
 * normalisation (:509-510), SimilarityNet's neighbour gather + weighted sum (:569-577),
 * softmax (:221) and depth regression (:226-237).
 *   score0          [B,D,H,W] per-hypothesis score from the 1x1x1 MLP
 *   depth_sample    [B,D,H,W]
 *   xnorm           [B,D,H,W] normalised inverse depth from pmb200_init_propagate, or NULL
 *                   (then it is recomputed from depth_sample at every tap: same result, slower)
 *   offsets         [B,2K,H,W] raw output of eval_conv
 *   feature_weight  [B,K,H,W]
 *   prob_out        [B,D,H,W]   depth_out [B,H,W]
 *   is_inverse      stage-1 last-iteration inverse-depth index regression (:227-234)
 */
int pmb200_adaptive_eval(const float *score0, const float *depth_sample, const float *xnorm,
                         const float *xnorm_score, /* NULL, or interleaved [B,D,H,W,2] replacing score0 + xnorm */
                         const float *offsets, int offsets_channels_last,
                         const float *feature_weight, const float *depth_min, const float *depth_max,
                         float *prob_out, float *depth_out,
                         int B, int D, int H, int W, int K, int dilation,
                         float interval_scale, int is_inverse, void *stream);

/* ====================================================================================
 * Backward entry points (training configuration).  Which gradients exist follows the
 * reference's graph: the warp grid is built under no_grad
 * (models/module.py:147), depth_weight and the incoming depth are detached
 * (models/patchmatch.py:74,85,503,506,669), FeatureWeightNet sees a detached reference
 * feature (:475).  All d_* outputs have the shape of the tensor they differentiate.
 * ==================================================================================== */

/* K-A backward: gradients w.r.t. the reference feature and the source features.
 *   grad_out  view_weights == NULL: [V,B,G,D,H,W] (per-view similarities)
 *             view_weights != NULL: [B,G,D,H,W]   (weighted average; the weights get no gradient here,
 *                                                  they arrive detached on every iteration that uses this mode)
 *   d_ref_nhwc [B,H,W,C]   d_src_nhwc [V,B,Hs,Ws,C] (zeroed by the call, then 128-bit vector atomics) */
int pmb200_warp_corr_backward(const float *ref_nhwc, const float *src_nhwc, const float *rt,
                              const float *depth, const float *view_weights, const float *grad_out,
                              float *d_ref_nhwc, float *d_src_nhwc,
                              int V, int B, int C, int G, int H, int W, int Hs, int Ws, int D,
                              void *stream);

/* aggregate_views backward: d_sims [V,B,G,D,H,W], d_view_weights [B,V,H,W]. */
int pmb200_aggregate_views_backward(const float *sims, const float *view_weights, const float *grad_out,
                                    float *d_sims, float *d_view_weights,
                                    int V, int B, int G, int D, int H, int W, void *stream);

/* K-A' backward: gradient of the self-correlation [B,G,K,H,W] w.r.t. the raw evaluation offsets [B,2K,H,W]. */
int pmb200_offset_corr_backward(const float *ref_nhwc, const float *offsets, const float *grad_out,
                                float *d_offsets, int B, int C, int G, int H, int W, int K, int dilation,
                                void *stream);

/* K-C backward: gradient of the sorted hypotheses [B,Ns+Kp,H,W] w.r.t. the raw propagation offsets [B,2Kp,H,W]. */
int pmb200_init_propagate_backward(const float *seed_map, const float *offsets,
                                   const float *depth_min, const float *depth_max,
                                   const float *grad_out, float *d_offsets,
                                   int mode, int B, int H, int W, int Ns, int Kp, int dilation,
                                   float interval_scale, void *stream);

/* K-B backward.  prob is the forward's probability output; grad_depth [B,H,W] and/or grad_prob [B,D,H,W]
 * (either may be NULL).  d_score0 [B,D,H,W] (zeroed by the call, atomics), d_depth_sample [B,D,H,W],
 * d_offsets [B,2K,H,W], d_feature_weight [B,K,H,W]. */
int pmb200_adaptive_eval_backward(const float *score0, const float *depth_sample, const float *xnorm,
                                  const float *offsets, const float *feature_weight,
                                  const float *depth_min, const float *depth_max, const float *prob,
                                  const float *grad_depth, const float *grad_prob,
                                  float *d_score0, float *d_depth_sample, float *d_offsets,
                                  float *d_feature_weight,
                                  int B, int D, int H, int W, int K, int dilation,
                                  float interval_scale, int is_inverse, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PATCHMATCH_B200_H_ */
