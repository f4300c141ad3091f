/* dsr_b200 - C-ABI of the B200-native hot path of the image-guided depth-enhancement training step.
 *
 * Drop-in boundary.  The reference (neeek2303/Depth-Enhancement-and-Super-Resolution) is pure
 * Python/PyTorch and has no FFI of its own; every entry point below replaces the ATen call(s) the
 * reference makes at the cited file:line (citations into /root/reference).  INTEGRATION.md shows
 * the ctypes binding a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only (no torch types); every pointer is DEVICE memory
 * owned by the caller and outliving the call unless the name says `host`; `stream` is a
 * cudaStream_t passed as void*; calls only enqueue work, never synchronise and never allocate;
 * return 0 on success, a negative DSR_ERR_* otherwise (message via dsr_last_error_string(),
 * thread-local).  The caller selects the device (cudaSetDevice) before calling.
 * Layouts: "NCHW planes" = the reference's layout for depth / normals / masks; "NHWC" = the
 * library's internal activation layout (torch channels_last memory).
 */
#ifndef DSR_B200_H
#define DSR_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define DSR_B200_VERSION 100

#define DSR_PAD_ZERO 0
#define DSR_PAD_REFLECT 1      /* nn.ReflectionPad2d            models/networks.py:378,413,453 */
#define DSR_PAD_REPLICATE 2    /* Conv2d(padding_mode='replicate') models/translation_network.py:472 */

#define DSR_ACT_NONE 0
#define DSR_ACT_RELU 1         /* nn.ReLU(True)                 models/networks.py:381,548 */
#define DSR_ACT_LRELU 2        /* nn.LeakyReLU(0.2, False)      models/networks.py:546 */
#define DSR_ACT_TANH 3         /* nn.Tanh                       models/networks.py:415,557 */

int dsr_version(void);
const char* dsr_last_error_string(void);
int dsr_device_sm_count(void);
int dsr_device_arch(void);      /* 100 on B200 */

/* ---- loss stack: depth-derived stencils and reductions (NCHW planes, fp32) ------------------ */

/* hole = 1[d <= border], valid = 1 - dilate3x3(hole).  models/main_model.py:208-230.  hole may be NULL. */
int dsr_hole_valid_masks(const float* depth, int B, int H, int W, float border, float* hole, float* valid,
                         void* stream);
/* out = (depth < thr) ? 0 : 1: the valid-depth mask of the Image Guidance step, models/I2D_model.py:223,226. */
int dsr_below_mask(const float* depth, long n, float thr, float* out, void* stream);
/* random rectangle holes.  models/main_model.py:257-298 (+ :354-357, :396 for `extra`).
 * rects int32 [B][max_rects][4] = (x, y, size_x, size_y), counts int32 [B] (both device).
 * gt_mask u8 {0,1}; masked = gt ? depth : -1; extra = 1[(masked < extra_border) || !gt] (may be NULL). */
int dsr_rect_holes(const float* valid, const float* depth, const int* rects, const int* counts, int max_rects,
                   int B, int H, int W, float extra_border, unsigned char* gt_mask, float* masked, float* extra,
                   void* stream);
/* image-space normals x scale, fwd / bwd.  models/norms.py:185-190 (caller's x100: main_model.py:345-349). */
int dsr_normals_old_fwd(const float* depth, int B, int H, int W, float scale, float* out /*B,3,H,W*/, void* stream);
int dsr_normals_old_bwd(const float* depth, const float* gout, int B, int H, int W, float scale, float* gdepth,
                        void* stream);
/* camera-space normals, fwd / bwd.  models/norms.py:103-108, :75-101, :29-73.
 * cams: double [B][11] = K^-1 row-major (9), w0 + shift, h0 + shift.  A plane whose K^-1 has the last row (0, 0, 1) (every
 * pin-hole K) runs the closed fp32 form of csrc/stencil_math.cuh (aff_*: the reference's fp64 point map cancelled
 * analytically, within 2e-6 of it); any other camera runs the reference's fp64 arithmetic per pixel.  Chosen per plane on the
 * device. */
int dsr_normals_new_fwd(const float* depth, const double* cams, int B, int H, int W, float* out, void* stream);
int dsr_normals_new_bwd(const float* depth, const float* gout, const double* cams, int B, int H, int W,
                        float* gdepth, void* stream);
/* tv_loss: *out_sum += sum of squared forward differences.  models/main_model.py:15-19.
 * bwd: gx = coef * (*gscale) * d tv / dx. */
int dsr_tv_fwd(const float* x, long planes, int H, int W, double* out_sum, void* stream);
int dsr_tv_bwd(const float* x, long planes, int H, int W, const float* gscale, float coef, float* gx, void* stream);
/* masked L1 / L2: t = (a*m1)*m2 - (b*m1)*m2; out2[0] += sum|t|, out2[1] += sum t^2.  a,b (B,C,plane),
 * masks (B,1,plane), m2 may be NULL.  L1Loss/MSELoss call sites models/main_model.py:352,371-372,383-398.
 * bwd: gb = -(c1*(*g_l1)*sign(t) + c2*(*g_l2)*2t) * m. */
int dsr_masked_diff_fwd(const float* a, const float* b, const float* m1, const float* m2, int B, int C, long plane,
                        double* out2, void* stream);
int dsr_masked_diff_bwd(const float* a, const float* b, const float* m1, const float* m2, int B, int C, long plane,
                        const float* g_l1, const float* g_l2, float c1, float c2, float* gb, void* stream);
/* monitoring sums: out3 += (sum d*m, sum p*m, sum |d*m - p*m|).  models/main_model.py:308-318. */
int dsr_masked_sums(const float* d, const float* p, const float* m, long total, double* out3, void* stream);
/* smoothness loss pieces.  models/main_model.py:22-73 (F.upsample bilinear align_corners=True :34). */
int dsr_bilinear_ac_fwd(const float* x, long planes, int H, int W, int nh, int nw, float* out, void* stream);
int dsr_bilinear_ac_bwd(const float* g, long planes, int H, int W, int nh, int nw, float* gx /* += */, void* stream);
int dsr_smooth_level_fwd(const float* d, const float* img, int B, int C, int h, int w, double* out2, void* stream);
int dsr_smooth_level_bwd(const float* d, const float* img, int B, int C, int h, int w, const float* gscale,
                         float cx, float cy, float* gd, int accumulate, void* stream);
/* Row-ring forms of the two calls above for large plane sets (csrc/stencil_ring.cu: image rows streamed through a shared-memory
 * ring by bulk copies, one producer warp + consumer warps).  dsr_smooth_level_fwd / _bwd dispatch here on their own when
 * dsr_smooth_ring_suits() says so (>= 2 M pixels, w % 4 == 0, w >= 64, rows fit shared memory); calling them directly on a
 * shape that does not suit returns DSR_ERR_ARG.  Same results as the register kernels up to fp32 summation order. */
int dsr_smooth_level_fwd_ring(const float* d, const float* img, int B, int C, int h, int w, double* out2, void* stream);
int dsr_smooth_level_bwd_ring(const float* d, const float* img, int B, int C, int h, int w, const float* gscale,
                              float cx, float cy, float* gd, int accumulate, void* stream);
int dsr_smooth_ring_suits(int B, int C, int h, int w);
/* SSIM, 11x11 Gaussian sigma 1.5, zero padding: *out_sum += sum of the SSIM map; map may be NULL.
 * models/pytorch_ssim/__init__.py:17-37. */
int dsr_ssim_fwd(const float* a, const float* b, long planes, int H, int W, double* out_sum, float* map,
                 void* stream);
/* resampling of the super-resolution step, tensors [N][H][W][C] (C = 1: NCHW planes with N = B*C).
 * F.interpolate(mode='bicubic') (align_corners=False, A=-0.75): models/main_sr_model.py:279-293, :361, :368-372,
 * :396-398; its adjoint (gx += W^T gy, gx zeroed by the caller) for :361; F.interpolate(mode='nearest'): :394-395,
 * :452, :459. */
int dsr_bicubic_fwd(const float* x, int N, int H, int W, int C, int Ho, int Wo, float* y, void* stream);
int dsr_bicubic_bwd(const float* gy, int N, int H, int W, int C, int Ho, int Wo, float* gx /* += */, void* stream);
int dsr_nearest_fwd(const float* x, int N, int H, int W, int C, int Ho, int Wo, float* y, void* stream);

/* on-disk formats either side of the path.  Input (data/my_main_dataset.py:35-52): uint16 depth in mm -> min(d, max_mm) /
 * max_mm * 2 - 1; HWC uint8 RGB -> CHW (x - 127.5) / 127.5.  Output (models/main_model.py:321-333, --save_all):
 * uint16(clip((pred + 1) / 2, 0, 1) * scale) of rows [crop, H - crop). */
int dsr_u16_to_depth(const unsigned short* in, long n, int max_mm, float* out, void* stream);
int dsr_u8_to_image(const unsigned char* in, int N, int H, int W, int C, float* out, void* stream);
int dsr_depth_to_u16(const float* pred, int N, int H, int W, int crop, float scale, unsigned short* out, void* stream);

/* GPU augmentation stage of the dataset (data/my_main_dataset.py:56-90 `trasform`; albumentations 0.4.6 + OpenCV):
 * dsr_resize_area_int = cv2.resize(INTER_AREA) for integer down-scale factors (A.Resize(interpolation=3), :69);
 * dsr_augment_gather = A.Rotate(+-30) [cv2.warpAffine INTER_LINEAR / BORDER_REFLECT_101, bit-exact fixed-point coordinates]
 * -> A.RandomCrop or A.PadIfNeeded(512, 640) [top / left < 0: reflect-101 padding] -> A.HorizontalFlip -> np.clip(-1, 1)
 * (:71-86) of N samples with C planes each; minv double [N][6] = inverse rotation map, ipar int [N][4] = {rotate?, top, left,
 * flip} (drawn on the host in the library's `random` call order, like the rectangle tables). */
int dsr_resize_area_int(const float* src, long planes, int H, int W, int fy, int fx, float* dst, void* stream);
int dsr_augment_gather(const float* src, int N, int C, int Hs, int Ws, const double* minv, const int* ipar, float* dst, int Hd,
                       int Wd, void* stream);

/* batch evaluator of new_metrics.py (:115-191): per-image fp64 sums, out double [B][16] =
 * {n, sum|d|, sum d^2} over ~target_hole | the same over ~target_hole & hole | over ~(hole | target_hole) |
 * {3 n, sum |dn|^2} of the first-order normals outside the dilated target holes (kinv: double [B][9] = K^-1, may be NULL) |
 * {n, sum} of the 'valid' 11x11 SSIM map (with_ssim).  hole = input < thr, target_hole = target < thr (:224-225). */
int dsr_eval_metric_sums(const float* pred, const float* target, const float* input, const double* kinv, int B, int H, int W,
                         float hole_threshold, double max_depth, int with_ssim, double* out, void* stream);

/* translation block only.  Field-of-view normals translation_network.SurfaceNormals (models/translation_network.py:329-360),
 * fwd / bwd (gdepth zeroed by the caller, +=); CosSimLoss (:310-316): *out_sum += sum over pixels of 1 - cos(x, y) along the
 * channel dimension, bwd: gx = coef * (*gscale) * d sum / dx. */
int dsr_fov_normals_fwd(const float* depth, int B, int H, int W, float* out /*B,3,H,W*/, void* stream);
int dsr_fov_normals_bwd(const float* depth, const float* gout, int B, int H, int W, float* gdepth /* += */, void* stream);
int dsr_cos_sim_fwd(const float* x, const float* y, int B, int C, long plane, double* out_sum, void* stream);
int dsr_cos_sim_bwd(const float* x, const float* y, int B, int C, long plane, const float* gscale, float coef, float* gx,
                    void* stream);
/* MaskedCosSimLoss (models/translation_network.py:320-327; call site translation_model.py:225): the same with a per-pixel
 * weight mask [B][plane]: *out_sum += sum mask * (1 - cos).  (The caller applies the reference's 1 / (sum(mask) + 1e+6).) */
int dsr_cos_sim_masked_fwd(const float* x, const float* y, const float* mask, int B, int C, long plane, double* out_sum,
                           void* stream);
int dsr_cos_sim_masked_bwd(const float* x, const float* y, const float* mask, int B, int C, long plane, const float* gscale,
                           float coef, float* gx, void* stream);
/* MaskedMeanDif (models/translation_network.py:288-293; call sites translation_model.py:243-247): sums (zeroed by the caller,
 * double [B][2]) += per-sample (sum (y - x) * mask, sum mask); *loss = mean_b |sums[b][0] / (sums[b][1] + 1e-6)|;
 * bwd: gx = d loss / dx * (*gscale). */
int dsr_masked_mean_dif_fwd(const float* x, const float* y, const float* mask, int B, long plane, double* sums, float* loss,
                            void* stream);
int dsr_masked_mean_dif_bwd(const float* mask, const double* sums, int B, long plane, const float* gscale, float* gx, void* stream);

/* ---- network plumbing (NHWC fp32) ------------------------------------------------------------ */
int dsr_nchw_to_nhwc(const float* x, float* y, int N, int C, long P, void* stream);
int dsr_nhwc_to_nchw(const float* x, float* y, int N, int C, long P, void* stream);
/* dst[p][doff + c] (=|+=) src[p][soff + c]: torch.cat(dim=1) and its backward.  models/networks.py:629,
 * models/main_model.py:302-306, models/translation_network.py:549. */
int dsr_copy_channels(const float* src, int srcC, int soff, float* dst, int dstC, int doff, int nC, long npix,
                      int accumulate, void* stream);
int dsr_pad2d_fwd(const float* x, float* y, int N, int H, int W, int C, int pad, int mode, void* stream);
int dsr_pad2d_bwd(const float* gy, float* gx, int N, int H, int W, int C, int pad, int mode, void* stream);
/* the same with an explicit row pitch of gy in pixels (wpitch >= W + 2 pad): the grouped data gradient of the 7x7 heads
 * (dsr_b200/ops.py _tc_dgrad_group) computes rows rounded up to whole groups of 4 pixels.  networks.py:378, :413. */
int dsr_pad2d_bwd_pitch(const float* gy, float* gx, int N, int H, int W, int C, int pad, int mode, int wpitch, void* stream);
/* ... plus a second gradient of the same (N, H, W, C) tensor added in the same pass (NULL = none): the skip connection of a
 * residual block hands its gradient to the block's first convolution, whose padding adjoint is the last pass of the block's
 * backward (out = x + conv_block(x), models/networks.py:478-480; autograd would add the two with a separate kernel). */
int dsr_pad2d_bwd_pitch_add(const float* gy, const float* add, float* gx, int N, int H, int W, int C, int pad, int mode,
                            int wpitch, void* stream);
int dsr_act_fwd(const float* x, float* y, long n, int kind, float slope, void* stream);
int dsr_act_bwd(const float* ref, const float* gy, float* gx, long n, int kind, float slope, void* stream);
/* per-(n,c) (sum, sum of squares) over P pixels, accumulated into sums (double [N][C][2], pre-zeroed). */
int dsr_channel_sums(const float* x, int N, long P, int C, double* sums, void* stream);
/* sums -> prm float [3][N][C] = (mean, scale, shift).  groups 0: InstanceNorm2d(affine=False)
 * models/networks.py:30;  groups 8: GroupNorm(8,C,affine) models/translation_network.py:46. */
int dsr_norm_finalize(const double* sums, int N, int C, long P, int groups, const float* gamma, const float* beta,
                      float eps, float* prm, void* stream);
/* y = act((x - mean) * scale + shift) (+ res).  act in {NONE, RELU}.  Residual: models/networks.py:480. */
int dsr_norm_apply_fwd(const float* x, const float* prm, const float* res, float* y, int N, long P, int C, int act,
                       void* stream);
/* dsr_norm_finalize + dsr_norm_apply_fwd in ONE launch: every block derives its sample's (mean, scale, shift) from the raw
 * sums (the finalize kernel's own arithmetic: bit-identical) and prm_out receives them for the backward pass.  Same reference
 * call sites: models/networks.py:30, :380-381, :480 (InstanceNorm2d [+ReLU] [+ residual]), translation_network.py:46. */
int dsr_norm_apply_fwd_fin(const float* x, const double* sums, int groups, const float* gamma, const float* beta, float eps,
                           float* prm_out, const float* res, float* y, int N, long P, int C, int act, void* stream);
int dsr_in_bwd_sums(const float* x, const float* dy, const float* prm, int N, long P, int C, int act, double* sums2,
                    void* stream);
int dsr_in_bwd_apply(const float* x, const float* dy, const float* prm, const double* sums2, float* dx, int N,
                     long P, int C, int act, void* stream);
/* GroupNorm(groups, C, affine) [+ReLU] backward, translation_network.py:46 (the translation generators).  prm = (mean, rstd, 0)
 * per (n, c) from dsr_norm_finalize with gamma = beta = NULL.  sums2 double [N][C][2] (pre-zeroed) = (sum g, sum g * xhat);
 * finalize: coef float [N][C][2] = the two group means of the apply pass, dgamma / dbeta float [C] (=|+=);
 * apply: dx = rstd * (gamma * g - coef0 - xhat * coef1). */
int dsr_gn_bwd_sums(const float* x, const float* dy, const float* prm, const float* gamma, const float* beta, int N, long P,
                    int C, int act, double* sums2, void* stream);
int dsr_gn_bwd_finalize(const double* sums2, const float* gamma, int N, int C, int groups, long P, float* coef, float* dgamma,
                        float* dbeta, int accumulate, void* stream);
int dsr_gn_bwd_apply(const float* x, const float* dy, const float* prm, const float* gamma, const float* beta, const float* coef,
                     float* dx, int N, long P, int C, int act, void* stream);
/* 4-D parameter [D0][D1][R][S] <-> GEMM operand [(r*S+s)*Ck + ck][Co]; kdim selects which of D0/D1 is ck. */
int dsr_pack_weight(const float* w, int D0, int D1, int R, int S, int kdim, float* out, void* stream);
int dsr_unpack_weight(const float* packed, int D0, int D1, int R, int S, int kdim, float* w, int accumulate,
                      void* stream);
int dsr_cvt_f64_f32(const double* in, long stride_in, float* out, long n, float scale, int accumulate, void* stream);
/* out (+)= (float) sum over the `reps` replica rows of a double accumulator [reps][n]: the bias gradient
 * (sum of dY over batch and pixels; every nn.Conv2d / nn.ConvTranspose2d with bias, networks.py:379-415) from the replicated
 * channel sums dsr_tc_prep takes on the way. */
int dsr_sum_reps_f64_f32(const double* in, int reps, long n, float* out, int accumulate, void* stream);
/* Loss assembly (models/main_model.py:393-417: loss_G = sum of ~14 weighted scalar terms, times scale_G) in ONE launch:
 * out = scale * sum_k sum_j weights[k][j] * terms[k][j] over n <= 32 tiny device tensors of counts[k] <= 4 elements each
 * (`terms`, `counts`, `weights` are HOST arrays; weights is flat, 4 per term).  The backward form writes
 * grads[4 k + j] = g * scale * weights[k][j] for every term at once.  Replaces ~40 scalar multiply / add kernels per pass. */
int dsr_loss_sum_fwd(const float* const* terms, const int* counts, const float* weights, int n, float scale, float* out,
                     int* nonfinite /* device, may be NULL: incremented when the sum is NaN / Inf (a guard readable outside a graph) */,
                     void* stream);
int dsr_loss_sum_bwd(const float* g, const float* weights, int n, float scale, float* grads, void* stream);

/* ---- convolutions ---------------------------------------------------------------------------- */
/* generic fp32 implicit GEMM (CUDA cores): out[m][co] = bias[co] + sum_k G(m,k) Wk[k][co].
 * nn.Conv2d / nn.ConvTranspose2d call sites: models/networks.py:379,385,406,414,453,544,553,605,612;
 * models/translation_network.py:472,478,495,508,563,568. */
int dsr_conv_simt(const float* G, const float* Wk, const float* bias, float* out, int N, int Hg, int Wg, int Cg,
                  int Ho, int Wo, int Co, int R, int S, int stride, int pad, int transposed, int act_out,
                  void* stream);
int dsr_wgrad_simt(const float* G, const float* D, float* dWk, int N, int Hg, int Wg, int Cg, int Ho, int Wo, int Cd,
                   int R, int S, int stride, int pad, void* stream);

/* one-output-channel heads on CUDA cores (csrc/conv_out1.cu): Conv2d R x S stride 1 (w = [1][C][R][S]) or, with
 * transposed = 1, ConvTranspose2d 4x4 stride 2 padding 1 (w = [C][1][4][4]); fp32 NHWC input read once through shared
 * memory with the fused (mean, scale, shift) + ReLU / LeakyReLU prologue and the padding mode; bias + optional tanh.
 * models/translation_network.py:495 (64 -> 1, 7x7), models/networks.py:553 (128 -> 1).  out = N x Ho x Wo (x 1). */
int dsr_conv_out1(const float* x, int N, int H, int W, int C, const float* prm, int act_in, float slope,
                  const float* w, const float* bias, int R, int S, int pad, int pad_mode, int transposed,
                  int act_out, float* out, void* stream);
/* data gradient of a stride-1 Conv2d with ONE output channel (weights [1][C][R][S]): gx[n][i][j][c] = sum_{r,s}
 * g[n][i + off - r][j + off - s] * W[c][r][s], gx (N, Hx, Wx, C) NHWC; off = the conv's implicit zero padding (0 when gx is
 * the gradient of an explicitly padded input).  models/translation_network.py:495 (generator head), :773 (discriminator head). */
int dsr_conv1_dgrad(const float* g, int N, int Ho, int Wo, const float* w, int C, int R, int S, int off, float* gx, int Hx, int Wx,
                    void* stream);
/* border term of the data gradient of a stride-2, padding-1 Conv2d with replicate / reflect padding (weights [Co][Ci][R][S]):
 * gx (N, H, W, Ci) already holds the zero-padding data gradient; this ADDS the one-pixel frame of the padded-input gradient
 * onto the source pixels the padding mode read it from.  models/translation_network.py:478 (encoder 4x4 stride-2 convs,
 * padding_mode='replicate'). */
int dsr_conv_s2_border_dgrad(const float* g, int N, int Ho, int Wo, int Co, const float* w, int Ci, int R, int S, int pad_mode,
                             float* gx, int H, int W, void* stream);

/* tcgen05 / TMEM / TMA implicit GEMM (csrc/conv_tc.cu).  Three calls per convolution:
 *   dsr_tc_prep        fp32 NHWC activation -> arranged bf16 hi(+lo) operand [N][Ha][Wa][Ca]; fuses the preceding
 *                      norm-apply (prm = (mean, scale, shift) or NULL), ReLU / LeakyReLU, and the padding
 *                      (nn.ReflectionPad2d / padding_mode='replicate' / zeros).  layout: NORMAL (Ca >= Cp >= C),
 *                      PAIR (g = Ca/Cp horizontally adjacent pixels side by side: g = 2 for C = 32 layers, g = 8 for the
 *                      1..8-channel first layers, so one 128-byte swizzle row holds a full K block), S2D (space-to-depth of the padded
 *                      input for stride-2 convs, Ca = 4*Cp).
 *   dsr_tc_pack_weight 4-D fp32 parameter -> bf16 hi(+lo) [Cout][T*Ca] K-major for the matching variant.
 *   dsr_tc_gemm        out[n, h*os+ph, w*os+pw, co] = bias[co] + sum_t sum_c A[n, h+ah+dr[t], w+aw+ds[t], c] W[co][t*Ca+c]
 *                      npass 1 = single 16-bit pass, 2 = A hi+lo, 3 = A and W hi+lo; tap tables are HOST arrays;
 *                      f16 = 0: bf16 operands (8-bit significands; hi+lo = 16 bits), f16 = 1: IEEE half operands
 *                      (11-bit; hi+lo = 22 bits, fp32-class products) with weights pre-scaled by the power of two
 *                      `wscale` at pack time and out_scale = 1/wscale applied to the accumulator;
 *                      split_k: 1 = off, -1 = auto (tiny-M layers), >1 = that many K splits (out must not alias);
 *                      nphase: 1, or 4 = all four output phases of a stride-2 transposed conv in one launch (os = 2,
 *                      ph = pw = 0; phase (a, b) adds (a, b) to the A offsets and to the output offsets and reads the
 *                      weight rows [phase*Cout, (phase+1)*Cout) of W = four stacked [Cout][T*Ca] matrices). */
#define DSR_TC_LAYOUT_NORMAL 0
#define DSR_TC_LAYOUT_PAIR 1
#define DSR_TC_LAYOUT_S2D 2
#define DSR_TC_W_CONV 0
#define DSR_TC_W_CONV_PAIR 1
#define DSR_TC_W_CONV_S2D 2
#define DSR_TC_W_CONVT_PH 3
#define DSR_TC_W_CONV_DGRAD 4
#define DSR_TC_W_CONV_DGRAD_PAIR 5   /* stride-1 data gradient producing 2 adjacent pixels x Cin outputs per GEMM row (N = 2*Cin) */
int dsr_tc_prep(const float* x, int N, int H, int W, int C, const float* prm, int act, float slope, int pad,
                int pad_mode, int layout, int Cp, void* A_hi, void* A_lo,
                void* A_bf /* optional: the same operand once more as plain bf16 (read by the weight-gradient GEMM later) */,
                int Ha, int Wa, int Ca, int f16,
                double* csum /* optional: += per-channel sums of x (bias gradient of a dY), pre-zeroed, double [csum_reps][C] */,
                int csum_reps /* replica rows of csum (block b adds into row b % csum_reps; >= 4 keeps the wide grid): the
                                 reader sums them with dsr_sum_reps_f64_f32 */,
                void* stream);
/* dsr_tc_prep with dsr_norm_finalize folded in (sums double [N][C][2] of the RAW input, usually taken by the producing GEMM's
 * epilogue): one launch less per normalised layer on the step's critical path.  prm_out (optional) = float [3][N][C] as
 * dsr_norm_finalize writes it.  models/networks.py:380-381, :478-480 (conv -> InstanceNorm -> ReLU -> conv chains). */
int dsr_tc_prep_fin(const float* x, int N, int H, int W, int C, const double* sums, int groups, const float* gamma,
                    const float* beta, float eps, float* prm_out, int act, float slope, int pad, int pad_mode, int layout,
                    int Cp, void* A_hi, void* A_lo, void* A_bf, int Ha, int Wa, int Ca, int f16, double* csum, int csum_reps,
                    void* stream);
/* The normalisation layer that closes a residual block AND the operand of the next convolution in one pass:
 * y_out = act(norm(x)) + res (fp32 NHWC, res optional; what dsr_norm_apply_fwd_fin writes) and A = the arranged 16-bit planes
 * of y_out (what dsr_tc_prep makes of it).  NORMAL / S2D layouts, C % 8 == 0, Ca / 8 a power of two <= 256; other shapes
 * return DSR_ERR_UNSUPPORTED (the caller runs the two passes).  models/networks.py:478-480 (out = x + conv_block(x)),
 * translation_network.py:46 (GroupNorm variant). */
int dsr_tc_prep_norm_res(const float* x, int N, int H, int W, int C, const double* sums, int groups, const float* gamma,
                         const float* beta, float eps, float* prm_out, int act, const float* res, float* y_out, int pad,
                         int pad_mode, int layout, int Cp, void* A_hi, void* A_lo, void* A_bf, int Ha, int Wa, int Ca, int f16,
                         void* stream);
/* InstanceNorm2d(affine=False) [+ReLU] backward apply AND the arranged dY operand of the convolution in front of the norm layer
 * in one pass: dx (fp32 NHWC, bit-identical to dsr_in_bwd_apply), A = hi (+lo) planes of dx with a zero frame of `pad` pixels
 * (what dsr_tc_prep would make of dx: NORMAL / S2D layouts), csum (optional, [csum_reps][C], pre-zeroed) += per-channel sums
 * of dx = the bias gradient of that convolution.  x, dy, dx: (N, H, W, C); prm from dsr_norm_finalize, sums2 from
 * dsr_in_bwd_sums.  Shapes outside the fast scheme return DSR_ERR_UNSUPPORTED (the caller runs the two passes).
 * models/networks.py:30, :380-381, :480 (the InstanceNorm2d layers) with :378-379, :413-414, :553 (the convolutions they follow). */
int dsr_tc_prep_in_bwd(const float* x, const float* dy, const float* prm, const double* sums2, float* dx, int N, int H, int W,
                       int C, int act, int pad, int layout, int Cp, void* A_hi, void* A_lo, int Ha, int Wa, int Ca, int f16,
                       double* csum, int csum_reps, void* stream);
/* dsr_tc_prep over torch.cat((x0, x1, x2, x3), dim=1) (NULL / 0 = absent) without materialising the concatenation:
 * models/main_model.py:305-306 (the 261-channel Task input), models/networks.py:629 (U-Net skip connections). */
int dsr_tc_prep_cat(const float* x0, int C0, const float* x1, int C1, const float* x2, int C2, const float* x3, int C3, int N,
                    int H, int W, int pad, int pad_mode, int layout, int Cp, void* A_hi, void* A_lo, void* A_bf, int Ha, int Wa,
                    int Ca, int f16, void* stream);
int dsr_tc_pack_weight(const float* w, int D0, int D1, int R, int S, int variant, int Cp, int phase_a, int phase_b,
                       int pad, int Cout, int T, int Ca, void* W_hi, void* W_lo, int f16, float wscale, void* stream);
int dsr_tc_gemm(const void* A_hi, const void* A_lo, int N, int Ha, int Wa, int Ca, const void* W_hi, const void* W_lo,
                int Cout, int T, const int* tap_dr, const int* tap_ds, int a_off_h, int a_off_w, int Ht, int Wt,
                const float* bias, float* out, int Ho, int Wo, int os, int ph, int pw, int nphase, int act, int npass,
                int split_k, int f16, float out_scale, void* stream);

/* second-generation GEMM (csrc/conv_tc2.cu): same contract as dsr_tc_gemm without split-K; the A operand of an 8x16-pixel
 * output tile is ONE shared-memory patch per 64-channel block that every tap reads at a shifted descriptor address;
 * persistent CTAs, double-buffered TMEM accumulators.  stats (may be NULL): double [N][Cout][2], pre-zeroed by the
 * caller; the epilogue adds the per-(n, c) sum and sum of squares of the outputs (InstanceNorm2d / GroupNorm statistics,
 * models/networks.py:30, models/translation_network.py:46).  Returns DSR_ERR_UNSUPPORTED (-3) for shapes it does not
 * cover (tap window wider than 9, outputs smaller than 128 pixels or narrower than 8). */
int dsr_tc_gemm2(const void* A_hi, const void* A_lo, int N, int Ha, int Wa, int Ca, const void* W_hi, const void* W_lo,
                 int Cout, int T, const int* tap_dr, const int* tap_ds, int a_off_h, int a_off_w, int Ht, int Wt,
                 const float* bias, float* out, int Ho, int Wo, int os, int ph, int pw, int nphase, int act, int npass,
                 int f16, float out_scale, double* stats, int a_mode, void* stream);
/* a_mode 1 = compact first-layer operand: A is an 8-channel arranged tensor (Ca = 8, zero-padded channels, made by
 * dsr_tc_prep with layout NORMAL); one tap per kernel row, its K block = 8 adjacent pixels x 8 channels read straight
 * from the 16-byte pixel rows (no 8x expansion in memory); W rows hold 64 K values per tap (variant CONV_PAIR, Cp = 8). */

/* third-generation GEMM for Cout >= 128 (csrc/conv_tc3.cu): channel-major accumulator (M = 128 output channels,
 * N = 8 x TH <= 256 pixels per MMA, the shape that runs at the tensor floor with both operands in shared memory),
 * 32-channel K blocks (SWIZZLE_64B), tile height picked per layer to fill whole waves of SMs, coalesced NHWC stores and
 * thread-local norm statistics.  Same contract as dsr_tc_gemm2; needs Ht >= 16 and Wt >= 8. */
int dsr_tc_gemm3(const void* A_hi, const void* A_lo, int N, int Ha, int Wa, int Ca, const void* W_hi, const void* W_lo,
                 int Cout, int T, const int* tap_dr, const int* tap_ds, int a_off_h, int a_off_w, int Ht, int Wt,
                 const float* bias, float* out, int Ho, int Wo, int os, int ph, int pw, int nphase, int act, int npass,
                 int f16, float out_scale, double* stats, void* stream);
int dsr_tc3_set_debug(long long* counters);

/* profiling aid: counters = device long long [grid][16] (or NULL to switch off): per-CTA clock64 ticks the roles of the
 * dsr_tc_gemm2 kernel spent waiting (see csrc/conv_tc2.cu).  Synchronous (cudaMemcpyToSymbol). */
int dsr_tc2_set_debug(long long* counters);

/* weight gradient on the tcgen05 path: dWp[cm][t*Ca + c] = sum_p M[p + moff][cm] * A[p + aoff + tap_t][c] over the
 * base grid (Hb x Wb x N); M = arranged dY (Conv2d) or x (ConvTranspose2d), A = the arranged operand of the forward
 * GEMM.  dWp is fp32 [Cm_real][T*Ca]; dsr_tc_unpack_wgrad scatters it back into the (D0, D1, R, S) parameter layout.
 * f16: bit 0 = M is IEEE half (else bf16), bit 1 = A is IEEE half (else bf16); the hardware requires both operands of
 * one kind::f16 MMA to have the same format (mixed formats raise an illegal-instruction fault), so pass 0 or 3. */
int dsr_tc_wgrad(const void* M_hi, const void* M_lo, int N, int Hm, int Wm, int Cm, int Cm_real, int m_off_h,
                 int m_off_w, const void* A_hi, const void* A_lo, int Ha, int Wa, int Ca, int T, const int* tap_dr,
                 const int* tap_ds, int a_off_h, int a_off_w, int Hb, int Wb, float* dWp, int npass, int f16,
                 float out_scale, int split_k, void* stream);
/* second-generation schedule of the same weight-gradient GEMM (csrc/wgrad_tc2.cu): eight 64-column blocks (taps / channel
 * blocks) per CTA in all 512 TMEM columns, N = 256 MMAs, 32-pixel K tiles, vector reductions for split-K; one 16-bit pass
 * (high planes only - the gate margins do not feel the weight-gradient operand rounding).  Same packed output. */
int dsr_tc_wgrad2(const void* M_hi, int N, int Hm, int Wm, int Cm, int Cm_real, int m_off_h, int m_off_w, const void* A_hi,
                  int Ha, int Wa, int Ca, int T, const int* tap_dr, const int* tap_ds, int a_off_h, int a_off_w, int Hb, int Wb,
                  float* dWp, int f16, float out_scale, int split_k, void* stream);
/* the same GEMM with partial = 1: dWp holds dsr_tc_wgrad2_splits(...) slabs of [Cm_real][T*Ca]; K split z STORES slab z (no
 * memset, no atomics) and dsr_tc_unpack_wgrad_splits sums the slabs in a fixed order on the way into the parameter layout:
 * the weight gradient of every nn.Conv2d / nn.ConvTranspose2d (networks.py:379-415, :544-616) becomes bit-reproducible and the
 * 30-50-way red.add fan-in of the short-K layers is gone.  partial = 0 is dsr_tc_wgrad2. */
int dsr_tc_wgrad2p(const void* M_hi, int N, int Hm, int Wm, int Cm, int Cm_real, int m_off_h, int m_off_w, const void* A_hi,
                   int Ha, int Wa, int Ca, int T, const int* tap_dr, const int* tap_ds, int a_off_h, int a_off_w, int Hb, int Wb,
                   float* dWp, int f16, float out_scale, int split_k, int partial, void* stream);
/* host-only query: the number of K splits (slabs) the launch above will use for this shape; < 0 = bad arguments */
int dsr_tc_wgrad2_splits(int N, int Hb, int Wb, int Cm_real, int Ca, int T, int split_k);
int dsr_tc_unpack_wgrad(const float* dWp, int D0, int D1, int R, int S, int variant, int Cp, int T, int Ca,
                        float* grad, int accumulate, void* stream);
int dsr_tc_unpack_wgrad_splits(const float* dWp, int nsplit, int D0, int D1, int R, int S, int variant, int Cp, int T, int Ca,
                               float* grad, int accumulate, void* stream);

/* ---- optimizer ------------------------------------------------------------------------------- */
/* torch.optim.Adam (defaults) over one flat arena.  models/main_model.py:176, :429. */
int dsr_adam_step(float* p, const float* g, float* m, float* v, long n, double lr, double b1, double b2, double eps,
                  int step, float grad_scale, void* stream);

/* same update with the step counter (int, incremented by the call) and (lr, beta1, beta2, eps) as doubles in DEVICE memory:
 * launch parameters never change, so the whole training step can be captured once and replayed as a CUDA graph. */
int dsr_adam_step_dev(float* p, const float* g, float* m, float* v, long n, const double* hyper, int* step,
                      float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif
