/* rpnet_b200 — C ABI of the B200 (sm_100a) RP-Net hot path.
 *
 * This is the drop-in boundary: every entry point is `extern "C"`, takes plain device pointers, sizes
 * and a CUDA stream (`cudaStream_t` passed as `void*`), allocates nothing, launches asynchronously on
 * the given stream, never throws, and returns 0 on success or a negative error code
 * (-1 CUDA error, -2 bad argument, -3 driver entry point missing) with the message available from
 * rpnet_last_error() (thread local).  Pointers are device pointers unless stated otherwise.
 *
 * The reference (uci-cbcl/RP-Net @169a0268) is pure PyTorch; the "FFI" for its hot path is the set of
 * ATen ops its nn.Modules call.  Each entry point cites the reference call site (file:line, relative to
 * the reference root) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes stub that binds
 * these symbols from the reference's Python modules.
 *
 * Layouts: activations are fp16 NHWC ("pixel major": [n][y][x][c], c contiguous) unless stated; images,
 * masks, prototypes and logits are fp32 NCHW / [n][y][x] exactly like the reference tensors.
 */
#ifndef RPNET_B200_H_
#define RPNET_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ABI version of this header (bumped on any signature change). */
int rpnet_abi_version(void);   /* currently 8 */

/* Message of the last failing call on this thread ("" if none). */
const char* rpnet_last_error(void);

/* Tap-list convolution as an implicit GEMM on tcgen05 tensor cores (fp16 operands, fp32 accumulate)
 * with fused per-channel affine (+ReLU), optional fused 2x2 max-pool, channel concat of two sources and
 * strided output placement (sub-pixel form of "nearest upsample x2 then 3x3 conv").
 * Replaces: nn.Conv2d + nn.BatchNorm2d(eval) + nn.ReLU          net/modules.py:47-54, :66-71
 *           nn.MaxPool2d(2, 2) after a conv_block                net/unet.py:397,442-455
 *           torch.cat((skip, up), dim=1) feeding a conv_block    net/unet.py:460,464
 *           nn.Upsample(scale_factor=2) feeding a conv           net/modules.py:67
 *           cre.w_k / cre.w_q / cre.q convs                      net/rp_net.py:50-59, :65-69
 *           VGG conv(+ReLU) incl. dilation 2                     net/vgg.py:53-56
 *   y[n, Y, X, co] = act( scale[co] * sum_{t, ci} w[t][co][ci] * x[n, y + dy[t], x + dx[t], ci] + shift[co] )
 * src0/src1: fp16 NHWC [n][h][w][c0] / [n][h][w][c1] (input channels = concat(src0, src1); c1 may be 0),
 *            c0, c1 multiples of 64; reads outside the h x w grid are zero (conv zero padding).
 * wpack:     fp16 [ntaps][cout][c0 + c1]; tap_dy/tap_dx: HOST int arrays of length ntaps (1..9).
 * cout:      multiple of 64.  scale/shift: fp32 [cout].
 * out_f16:   optional fp16 NHWC [n][out_h][out_w][out_c]; conv pixel (y, x) is stored at
 *            (y*oy_mul + oy_off, x*ox_mul + ox_off), channels [out_coff, out_coff + cout).
 * out_pool_f16: optional fp16 NHWC [n][h/2][w/2][cout] = 2x2/stride-2 max-pool of the activated output.
 * out_f32:   optional fp32 NHWC [n][h][w][cout].
 * Kernel selection (same results bit for bit, tests/test_gpu_kernels.py::test_conv_variants_bit_identical): cout % 256 == 0 runs as
 * CTA pairs (tcgen05 cta_group::2, M = 256; RPNET_CONV_2CTA=0 disables), cout == 64 with <= 9 k-blocks and many pixel tiles runs
 * weights-stationary (RPNET_CONV_NO_WS=1 disables). */
int rpnet_conv_igemm_f16(const void* src0, int c0, const void* src1, int c1, int n, int h, int w,
                         const void* wpack, int ntaps, const int* tap_dy, const int* tap_dx, int cout,
                         const float* scale, const float* shift, int relu, void* out_f16, int out_h, int out_w,
                         int out_c, int out_coff, int oy_mul, int oy_off, int ox_mul, int ox_off,
                         void* out_pool_f16, float* out_f32, void* stream);

/* A 64-output-channel tap-list conv (same operands as rpnet_conv_igemm_f16) whose epilogue also evaluates calDist
 * (net/rp_net.py:353-363) on the activated output: pred[i][p][pixel] = scaler * cos(y[i,pixel,:], protos[i % proto_sets][p][:])
 * with torch's per-norm 1e-8 clamp.  protos fp32 [proto_sets][n_protos][64], pred fp32 [n][n_protos][h*w]; out_f32 (optional)
 * fp32 NHWC [n][h][w][64] = the features themselves.  Fuses cre.q (net/rp_net.py:65-69,81) with the prototype match
 * (:287-303) in the eval forward: the 64-channel features never travel to HBM. */
int rpnet_conv_cos_f16(const void* src0, int c0, const void* src1, int c1, int n, int h, int w, const void* wpack, int ntaps,
                       const int* tap_dy, const int* tap_dx, const float* scale, const float* shift, int relu,
                       const float* protos, int n_protos, int proto_sets, float scaler, float* pred, float* out_f32, void* stream);

/* "Next" row N3 (ResNet18 backbone, net/rp_net.py:19-42).
 * rpnet_conv_res_f16: tap-list conv (as rpnet_conv_igemm_f16, one source, dense output) with a residual added after the affine
 *   and before the ReLU: out = act(scale * conv(src) + shift + res) — torchvision BasicBlock's `out += identity; relu`.
 *   res fp16 NHWC [n][h][w][cout] (null = no residual).
 * rpnet_conv7x7s2_stem_f16: torchvision resnet18 conv1 (7x7, stride 2, pad 3, no bias) + folded bn1 + ReLU on a 3-channel fp32
 *   NCHW image -> fp16 NHWC [n][(h-1)/2+1][(w-1)/2+1][64].  weight fp32 [64][3][7][7]. */
int rpnet_conv_res_f16(const void* src, int cin, int n, int h, int w, const void* wpack, int ntaps, const int* tap_dy,
                       const int* tap_dx, int cout, const float* scale, const float* shift, const void* res_f16, int relu,
                       void* out_f16, void* stream);
int rpnet_conv7x7s2_stem_f16(const float* img, int n, int h, int w, const float* weight, const float* scale, const float* shift,
                             int relu, void* out_f16, void* stream);

/* First encoder conv: fp32 NCHW image [n][cin][h][w] (cin 1 or 3) -> 64 channels, 3x3 pad 1, fused
 * affine (+ReLU), fp16 NHWC out [n][h][w][64].  weight fp32 [64][cin][3][3] (PyTorch layout).
 * Replaces encoder.Conv1.conv.0-2 (net/modules.py:48-50 via net/unet.py:405) and VGG features.0.0
 * (net/vgg.py:53-56). */
int rpnet_conv3x3_first_f16(const float* img, int n, int cin, int h, int w, const float* weight,
                            const float* scale, const float* shift, int relu, void* out_f16, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Split-fp16 ("fp32-class") convolution.  The tensor pipe multiplies 11-bit significands (fp16, bf16 and tf32 alike); one
 * rounding of the weights plus one of the activations per layer puts the logits of the 25-layer path 2e-3 .. 4e-3 (rel-Linf) away
 * from the reference's fp32 result — outside the 1e-3 the parity tests hold.  These entry points carry every encoder activation
 * and weight as a PAIR of fp16 tensors x = hi + lo (hi = fp16(x), lo = fp16(x - hi): 21+ significant bits) and accumulate
 *     hi.Wh + lo.Wh + hi.Wl        (the dropped lo.Wl term is 2^-22 relative)
 * in fp32 inside the same tcgen05 kernel: three k-block passes per tap instead of one.  Same replaced reference ops as
 * rpnet_conv_igemm_f16 (nn.Conv2d + BatchNorm2d + ReLU (+ MaxPool2d, cat, Upsample), net/modules.py:42-75, net/unet.py:435-467).
 *
 * rpnet_conv_split_f16: src*_hi / src*_lo fp16 NHWC (lo may be null: the source is used as plain fp16); wpack fp16
 *   [ntaps][cout][(c0 + c1) * (w_split ? 2 : 1)] = Wh | Wl; out_hi / out_lo, out_pool_hi / out_pool_lo: the activated output and
 *   its residual plane (lo pointers may be null), placed like rpnet_conv_igemm_f16's out_f16 / out_pool_f16.
 *   sums != null (train mode, scale = 1, shift = 0, relu = 0): also accumulates the BatchNorm statistics of the fp32
 *   accumulators like rpnet_conv_bnstats_f16 (group_start HOST array; keep_sums = 1 adds to existing sums: the four phase launches
 *   of a sub-pixel up-conv).
 * rpnet_conv3x3_first_split_f16: rpnet_conv3x3_first_f16 (fp32 arithmetic) that also writes the residual plane of its output.
 * rpnet_bn_stats_split_f16 / rpnet_bn_apply_split_f16: the train-mode BatchNorm passes on z = z_hi + z_lo, writing y (and the
 *   2x2 max-pooled y) as hi / lo planes.
 * rpnet_pack_conv_weight_split / rpnet_pack_upconv_weight_split: split = 1 packs the forward weights as
 *   [taps][cout][Wh (cin) | Wl (cin)], split = 2 as [taps][cout][Wh (cin) | fp8 corrections (2 * cin bytes)] (the data-gradient
 *   packs are unchanged: the backward runs single-term).
 *
 * fp8 corrections (the default precision, `b200_precision: split8`): x.w = hi.Wh + 2^-15 (lo8.Wh8 + x8.Wl8) — the main term on fp16
 *   operands, the two first-order corrections on e4m3 operands at twice the MMA rate (kind::f8f6f4) into the same fp32 accumulator,
 *   which the first main-term MMA scales by 2^-15 (tcgen05.mma scale-input-d).  A correction is 2^-11 of the main term, so its e4m3
 *   rounding (2^-4) lands at 2^-15 of the product: the same logits error as the three-pass form for two thirds of the
 *   tensor time.  Fixed scales: lo8 = e4m3((x - hi) * 2^9), x8 = e4m3(x * 2^-2), Wh8 = e4m3(Wh * 2^6), Wl8 = e4m3((w - Wh) * 2^17), saturating
 *   (exact corrections for |x| < 1792, |w| < 7; beyond, that element falls back to single-term accuracy).
 *   A "c8 plane" replaces the fp16 residual plane of an activation (same size): per pixel and 64-channel group 128 bytes = lo8 of the
 *   64 channels | x8 of the 64 channels; the second half of a weight pack row holds, per 64 input channels, Wh8 (64) | Wl8 (64).
 *   `lo_fmt` arguments: 0 = fp16 residual plane, 1 = c8 plane (channel counts must be multiples of 64).
 *   `w_split` of the conv entry points: 0 fp16 weights, 1 Wh | Wl pack, 2 fp8-correction pack with c8 source planes (and residual
 *   plane) writing c8 output planes, 3 the same inputs writing fp16 residual planes (the pre-BatchNorm z of the train path, whose
 *   |mean| >> std needs the full residual). */
int rpnet_conv_split_f16(const void* src0_hi, const void* src0_lo, int c0, const void* src1_hi, const void* src1_lo, int c1,
                         int n, int h, int w, const void* wpack, int w_split, int ntaps, const int* tap_dy, const int* tap_dx,
                         int cout, const float* scale, const float* shift, int relu, void* out_hi, void* out_lo, int out_h,
                         int out_w, int out_c, int out_coff, int oy_mul, int oy_off, int ox_mul, int ox_off,
                         void* out_pool_hi, void* out_pool_lo, float* out_f32, const int* group_start, int groups,
                         double* sums, int keep_sums, void* stream);
/* rpnet_conv_split_res_f16: rpnet_conv_split_f16 with a residual (hi / lo planes, lo may be null) added after the affine and before
 * the ReLU — torchvision BasicBlock's `out += identity; relu` (ResNet18 backbone, net/rp_net.py:19-42) in split precision.
 * rpnet_conv7x7s2_stem_split_f16: rpnet_conv7x7s2_stem_f16 that also writes the residual plane of its output. */
int rpnet_conv_split_res_f16(const void* src0_hi, const void* src0_lo, int c0, const void* src1_hi, const void* src1_lo, int c1,
                             int n, int h, int w, const void* wpack, int w_split, int ntaps, const int* tap_dy, const int* tap_dx,
                             int cout, const float* scale, const float* shift, const void* res_hi, const void* res_lo, int relu,
                             void* out_hi, void* out_lo, int out_h, int out_w, int out_c, int out_coff, int oy_mul, int oy_off,
                             int ox_mul, int ox_off, void* out_pool_hi, void* out_pool_lo, float* out_f32, const int* group_start,
                             int groups, double* sums, int keep_sums, void* stream);
int rpnet_conv7x7s2_stem_split_f16(const float* img, int n, int h, int w, const float* weight, const float* scale,
                                   const float* shift, int relu, void* out_f16, void* out_lo_f16, int lo_fmt, void* stream);
int rpnet_conv3x3_first_split_f16(const float* img, int n, int cin, int h, int w, const float* weight, const float* scale,
                                  const float* shift, int relu, void* out_f16, void* out_lo_f16, int lo_fmt, void* stream);
int rpnet_bn_stats_split_f16(const void* z_hi, const void* z_lo, int n, int h, int w, int c, const int* group_start, int groups,
                             double* sums, void* stream);
int rpnet_bn_apply_split_f16(const void* z_hi, const void* z_lo, const float* stats, int n, int h, int w, int c,
                             const int* group_start, int groups, int relu, void* y_f16, void* y_lo_f16, void* y_pool_f16,
                             void* y_pool_lo_f16, float* y_f32, int lo_fmt, void* stream);
int rpnet_pack_conv_weight_split(const float* w, int cout, int cin_real, int ntaps, int hole_start, int hole_len,
                                 void* w_fwd_f16, int split, void* w_dgrad_bf16, void* stream);
int rpnet_pack_upconv_weight_split(const float* w, int cout, int cin, void* wf_f16, int split, void* w16_bf16, void* stream);
/* rpnet_pack_conv_weight_split for every conv of the model in ONE launch (the packs follow every optimizer step): descs_host is
 * a HOST array of n (<= 40) layer descriptors with the arguments of rpnet_pack_conv_weight_split (device pointers inside). */
typedef struct {
  const float* w; void* w_fwd_f16; void* w_dgrad_bf16;
  int cout, cin_real, ntaps, hole_start, hole_len, split;
} rpnet_pack_desc;
int rpnet_pack_conv_weights(const rpnet_pack_desc* descs_host, int n, void* stream);
/* rpnet_maxpool_f16 on a split-fp16 activation: the maximum of hi + lo, written back as hi / lo planes (VGG pools between split convs). */
int rpnet_maxpool_split_f16(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int lo_fmt, int n, int h, int w, int c,
                            int k, int stride, int pad, void* stream);

/* F.avg_pool2d(mask[:, None], s): fp32 [n][h][w] -> fp32 [n][h/s][w/s].  net/rp_net.py:270,272. */
int rpnet_avgpool_mask_f32(const float* in, float* out, int n, int h, int w, int s, void* stream);

/* x_fg = x * m, x_bg = x * (1 - m); x fp16 NHWC with `pixels` = n*h*w pixels of c channels, m fp32 per
 * pixel.  net/rp_net.py:275,283 (the two arguments of self.cre). */
int rpnet_premask_f16(const void* x, const float* mask, void* x_fg, void* x_bg, long long pixels, int c,
                      void* stream);

/* Correlation(fmap1, fmap2, r) (net/rp_net.py:153-181) in its local zero-padded window form:
 *   out[n,y,x, a*(2r+1)+b] = 1/sqrt(c) * sum_ch f1[n,y,x,ch] * f2[n, y+(b-r), x+(a-r), ch]
 * f1, f2 fp16 NHWC [n][h][w][c]; out fp16 NHWC [n][h][w][out_c], channels >= (2r+1)^2 are zero. */
int rpnet_local_corr_f16(const void* f1, const void* f2, void* out, int n, int h, int w, int c, int radius,
                         int out_c, void* stream);

/* Fused relation head of the eval forward: Correlation(f1, f2, r) -> cat([corr, f1]) -> cre.q 1x1 conv + folded BN + ReLU ->
 * calDist against the prototypes (net/rp_net.py:79-84, 287-303), one tcgen05 kernel; only pred leaves the SM.
 * f1, f2 fp16 NHWC [n][h][w][c]; wq_pack fp16 [1][64][128 + c] (the corr block of the input channels padded to 128);
 * scale / shift fp32 [64]; protos fp32 [proto_sets][n_protos][64] (image i uses set i % proto_sets); pred fp32 [n][n_protos][h*w].
 * Returns -2 for shapes the fused kernel does not cover (callers then run rpnet_local_corr_f16 + rpnet_conv_cos_f16). */
int rpnet_relation_head_f16(const void* f1, const void* f2, const void* wq_pack, const float* scale, const float* shift,
                            const float* protos, int n_protos, int proto_sets, float scaler, float* pred, int n, int h, int w,
                            int c, int radius, void* stream);

/* getFeatures (net/rp_net.py:366-376) for two masks at once:
 *   out[i][k][ch] = sum_{Y,X} bilinear_up(feat[i])[ch,Y,X] * mask_k[i][Y,X] / (sum mask_k[i] + 1e-5)
 * feat fp32 NHWC [n][h][w][c] (c <= 64), mask0/mask1 fp32 [n][mask_h][mask_w], out fp32 [n][2][c]. */
int rpnet_masked_avg_pool_f32(const float* feat, const float* mask0, const float* mask1, float* out, int n,
                              int h, int w, int c, int mask_h, int mask_w, void* stream);

/* getPrototype (net/rp_net.py:379-391): raw fp32 [ways][shots][batch][2][c] (0 = fg, 1 = bg pooled
 * features) -> protos fp32 [batch][1 + ways][c] ordered [bg, fg_1 .. fg_ways] (net/rp_net.py:299). */
int rpnet_proto_finalize_f32(const float* raw, float* protos, int ways, int shots, int batch, int c,
                             void* stream);

/* calDist (net/rp_net.py:353-363): pred[b][p][pix] = scaler * cosine(feat[b][pix][:], protos[b][p][:])
 * with torch's per-norm eps clamp 1e-8.  feat fp32 NHWC [batch][hw][c] (c == 64), pred fp32 [batch][p][hw]. */
int rpnet_cos_sim_f32(const float* feat, const float* protos, float* pred, int batch, int hw, int c,
                      int n_protos, float scaler, void* stream);

/* Refinement tail (net/rp_net.py:303-312): logits = bilinear_up(pred, x scale, align_corners=False);
 * p_fg = sum_{k>=1} softmax(logits)[k]; m = soft_mask ? p_fg : (p_fg > 0.5); mask_out = avg_pool2d(m, scale).
 * pred fp32 [batch][p][h][w]; logits fp32 [batch][p][h*scale][w*scale]; mask_out fp32 [batch][h][w].
 * scale 4 or 8. */
int rpnet_upsample_tail_f32(const float* pred, float* logits, float* mask_out, int batch, int n_protos,
                            int h, int w, int scale, int soft_mask, void* stream);

/* nn.MaxPool2d(k, stride, pad) on fp16 NHWC [n][h][w][c] -> [n][ho][wo][c], ho = (h + 2*pad - k)/stride + 1
 * (implicit -inf padding).  VGG pools, net/vgg.py:24-30. */
int rpnet_maxpool_f16(const void* in, void* out, int n, int h, int w, int c, int k, int stride, int pad,
                      void* stream);

/* ===================================================================================================
 * Training path.  The reference trains through torch autograd (it ships no backward code and no train
 * script — SURVEY D9); each entry point cites the forward op whose forward/backward it restates.
 * Activation gradients are bf16 NHWC; weight gradients, BN statistics, prototypes and losses are fp32.
 * BatchNorm "call groups": images [group_start[g], group_start[g+1]) of a batched launch form one
 * nn.BatchNorm2d call of the reference (own batch statistics and running-stat update);
 * group_start is a HOST int array of groups + 1 entries.
 * =================================================================================================== */

/* Same contract as rpnet_conv_igemm_f16 with bf16 operands (src0/src1/wpack) and bf16 outputs: the data-gradient
 * of a conv is the conv of dZ with the transposed weights [tap][cin][cout] and the negated tap list. */
int rpnet_conv_igemm_bf16(const void* src0, int c0, const void* src1, int c1, int n, int h, int w,
                          const void* wpack, int ntaps, const int* tap_dy, const int* tap_dx, int cout,
                          const float* scale, const float* shift, int relu, void* out_bf16, int out_h, int out_w,
                          int out_c, int out_coff, int oy_mul, int oy_off, int ox_mul, int ox_off,
                          void* out_pool_bf16, float* out_f32, void* stream);

/* Bytes of scratch rpnet_conv_wgrad needs for this shape (negative = bad argument). */
long long rpnet_conv_wgrad_workspace_bytes(int c0, int c1, int n, int h, int w, int ntaps, int cout);

/* fp16 -> bf16 copy of n elements (n % 8 == 0): the weight-gradient GEMM needs both operands in one format. */
int rpnet_cvt_f16_to_bf16(const void* in_f16, void* out_bf16, long long n, void* stream);

/* Weight gradient of a tap-list conv on tcgen05 tensor cores (MN-major operands, split-K over pixels, deterministic):
 *   grad[co][ci][tap] (+)= sum_{n,y,x} dz[n,y,x,co] * x[n, y+dy[tap], x+dx[tap], ci]        (nn.Conv2d weight layout)
 * x0/x1: NHWC activations, bf16 (x_bf16 = 1) or fp16 (x_bf16 = 0: one tcgen05.mma takes a single operand format, so the
 * epilogue warps convert the x boxes to bf16 in shared memory as they land — no separate conversion pass), channel concat
 * like the forward; dz_bf16: bf16 NHWC [n][h][w][cout].
 * Packed input channels [hole_start, hole_start+hole_len) are padding and are skipped in `grad`.
 * Replaces autograd's conv weight gradient for net/modules.py:47-54,66-71 and net/rp_net.py:50-69. */
int rpnet_conv_wgrad(const void* x0, int c0, const void* x1, int c1, int x_bf16, const void* dz_bf16, int n, int h,
                     int w, int ntaps, const int* tap_dy, const int* tap_dx, int cout, float* grad, int hole_start,
                     int hole_len, int accumulate, void* workspace, long long workspace_bytes, void* stream);

/* up_conv (nn.Upsample(x2, nearest) + 3x3 conv, net/modules.py:61-75) in sub-pixel form, train mode: output parity phase
 * (py, px) of z is a 2x2-tap conv of the LOW-resolution input with row/column-summed weights (2.25x fewer MACs, and the
 * up-sampled map is never materialised).
 *   rpnet_pack_upconv_weight: w fp32 [cout][cin][3][3] -> wf fp16 [4 phases][4 taps][cout][cin] (forward) and
 *                             w16 bf16 [16][cin][cout] (data gradient as a 4x4 stride-2 conv of dZ).
 *   rpnet_upconv_phase_bnstats_f16: one phase of z [n][2h][2w][cout] from x_low [n][h][w][cin] with wphase = wf[py*2+px];
 *                             BatchNorm statistics of z accumulate over the four launches (keep_sums = 0 on the first).
 *                             -2 when the statistics cannot be fused for the shape (maps smaller than a pixel tile).
 *   rpnet_upconv_dgrad_bf16:  dx_low [n][h][w][out_c] (channels [out_coff, +cin)) from dz bf16 [n][2h][2w][cout].
 *   rpnet_upconv_wgrad:       grad [cout][cin][3][3] (=|+=) from x_low fp16/bf16 [n][h][w][cin] and dz bf16 [n][2h][2w][cout]. */
int rpnet_pack_upconv_weight(const float* w, int cout, int cin, void* wf_f16, void* w16_bf16, void* stream);
int rpnet_upconv_phase_bnstats_f16(const void* x_low, int cin, int n, int h, int w, const void* wphase, int py, int px, int cout,
                                   const float* ones, const float* zeros, void* z_f16, const int* group_start, int groups,
                                   double* sums, int keep_sums, void* stream);
int rpnet_upconv_dgrad_bf16(const void* dz, int cout, int n, int h, int w, const void* w16, int cin, void* out, int out_c,
                            int out_coff, const float* ones, const float* zeros, void* stream);
long long rpnet_upconv_wgrad_workspace_bytes(int cin, int n, int h, int w, int cout);
int rpnet_upconv_wgrad(const void* x_low, int x_bf16, const void* dz_bf16, int n, int h, int w, int cin, int cout, float* grad,
                       int accumulate, void* workspace, long long workspace_bytes, void* stream);

/* Weight gradient of the Cin = 1 first conv: grad[64][1][3][3] += sum dz * img.  net/unet.py:405 (encoder.Conv1.conv.0).
 * scratch576: 576 doubles.  Determinism (this and every reduction of the training path below): per-thread / per-warp / per-block
 * fp32 partial sums are formed in a fixed order and then accumulated in FP64 — adding fp32 addends into a double is exact, hence
 * independent of the order in which warps and blocks arrive, unless an addend is below 2^-30 of the running sum — so two
 * identical train steps produce bit-identical gradients. */
int rpnet_conv3x3_first_wgrad(const float* img, const void* dz_bf16, int n, int h, int w, float* grad, double* scratch576,
                              void* stream);

/* fp32 [cout][cin_real][taps] -> fp16 [taps][cout][cin] (forward pack) and/or bf16 [taps][cin][cout] (dgrad pack);
 * cin = cin_real + hole_len with zero padding channels at [hole_start, hole_start+hole_len). */
int rpnet_pack_conv_weight(const float* w, int cout, int cin_real, int ntaps, int hole_start, int hole_len,
                           void* w_fwd_f16, void* w_dgrad_bf16, void* stream);

/* Train-mode nn.BatchNorm2d (net/modules.py:49,52,69; net/rp_net.py:52,57,67), statistics pass:
 * sums[g][c] = {sum z, sum z^2} over the images of call group g, in fp64 (the variance is a difference of nearly equal
 * numbers for channels whose mean dwarfs their spread).  z fp16 NHWC. */
int rpnet_bn_stats_f16(const void* z, int n, int h, int w, int c, const int* group_start, int groups, double* sums,
                       void* stream);

/* Train-mode conv: z = conv(src0 | src1) without bias (it cancels inside batch-statistics BatchNorm) stored as fp16 NHWC
 * [n][h][w][cout], and the BatchNorm statistics of z in the same launch: sums[g][cout] = {sum z, sum z^2} over call group g,
 * accumulated from the fp32 accumulators in the conv epilogue (per-CTA shared-memory partials, one atomic flush per group;
 * when a pixel tile would straddle two call groups — maps smaller than a tile — the statistics pass runs as
 * rpnet_bn_stats_f16 instead).  ones / zeros: fp32 [cout] constant vectors (the epilogue's affine is the identity).
 * Replaces nn.Conv2d + the statistics half of nn.BatchNorm2d(train): net/modules.py:47-54,66-71, net/rp_net.py:50-69. */
int rpnet_conv_bnstats_f16(const void* src0, int c0, const void* src1, int c1, int n, int h, int w, const void* wpack,
                           int ntaps, const int* tap_dy, const int* tap_dx, int cout, const float* ones, const float* zeros,
                           void* z_f16, const int* group_start, int groups, double* sums, void* stream);

/* stats[g][c] = {mean, rstd, a = rstd*gamma, b = beta - mean*a}; running_mean/var (momentum, unbiased var) updated once
 * per call group in order, num_batches_tracked += groups.  conv_bias: the bias the conv kernel dropped (it cancels in
 * train-mode BN but is part of the running mean).  hw = pixels per image. */
int rpnet_bn_finalize_f32(const double* sums, const int* group_start, int groups, int c, int hw, const float* gamma,
                          const float* beta, const float* conv_bias, float eps, float momentum, float* running_mean,
                          float* running_var, long long* num_batches_tracked, float* stats, void* stream);

/* y = relu?(a*z + b): fp16 NHWC and/or fp32 NHWC and/or the 2x2 max-pooled fp16 copy (nn.MaxPool2d(2,2), net/unet.py:397). */
int rpnet_bn_apply_f16(const void* z, const float* stats, int n, int h, int w, int c, const int* group_start, int groups,
                       int relu, void* y_f16, void* y_pool_f16, float* y_f32, void* stream);

/* Backward of BatchNorm(batch stats) + ReLU: dz (bf16 NHWC) from the gradient of the activation, which is the sum of
 *   g_direct : bf16 (fp32 if d_is_f32) NHWC with pixel pitch d_ld, channel offset d_off          (may be null)
 *   g_pool   : bf16 NHWC [n][h/2][w/2] gradient of the max-pooled copy, routed to the window's first max (may be null)
 *   g_up     : bf16 NHWC [n][2h][2w] gradient of the nearest-x2 upsampled copy, summed per 2x2     (may be null)
 * dgamma[c] += sum dy*x_hat, dbeta[c] += sum dy (fp32, may be null).  scratch: 8-byte aligned, groups*c*6 floats
 * (fp64 sums [groups][c][2] followed by fp32 coefficients [groups][c][2]). */
int rpnet_bn_bwd(const void* z, const float* stats, int n, int h, int w, int c, const int* group_start, int groups, int relu,
                 const void* g_direct, int d_ld, int d_off, int d_is_f32, const void* g_pool_bf16, int p_ld, int p_off,
                 const void* g_up_bf16, int u_ld, int u_off, float* dgamma, float* dbeta, float* scratch,
                 void* dz_bf16, void* stream);

/* nn.Upsample(scale_factor=2) nearest, fp16 NHWC [n][h][w][c] -> [n][2h][2w][c].  net/modules.py:67. */
int rpnet_upsample2x_f16(const void* x, void* y, int n, int h, int w, int c, void* stream);

/* Backward of the pre-mask (net/rp_net.py:275,283) summed over `iters` uses of the same features:
 * dx[p][c] = sum_i dxfg[i][p][c]*mask[i][p] + dxbg[i][p][c]*(1 - mask[i][p]);  bf16 in / out, mask fp32 [iters][pixels]. */
int rpnet_premask_bwd_bf16(const void* dxfg, const void* dxbg, const float* mask, int iters, long long pixels, int c,
                           void* dx, void* stream);

/* Backward pieces of the VGG stack (net/vgg.py:22-58: conv + bias + ReLU, nn.MaxPool2d(3, stride, 1); no normalisation), for
 * training RP_Net with `backbone: vgg` (the reference cannot run that backbone through RP_Net at all).
 * rpnet_relu_bias_bwd: g = dy * [y > 0] (bf16 NHWC; y_f16 = null: no ReLU) and dbias[c] += sum_p g[p][c]; scratch_c: c doubles.
 * rpnet_maxpool_idx_f16: rpnet_maxpool_split_f16 that also records the window position (dy * k + dx, uint8 [n][ho][wo][c]) of the first
 *   maximum; rpnet_maxpool_bwd_bf16 routes dy back through those positions (every input pixel gathers from the windows covering it).
 * rpnet_conv3x3_first_wgrad_cin: rpnet_conv3x3_first_wgrad for a Cin-channel (<= 4) image [n][cin][h][w]: grad [64][cin][3][3] +=. */
int rpnet_relu_bias_bwd(const void* dy_bf16, const void* y_f16, long long pixels, int c, void* g_bf16, float* dbias, double* scratch_c,
                        void* stream);
int rpnet_maxpool_idx_f16(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int lo_fmt, void* idx_u8, int n, int h, int w,
                          int c, int k, int stride, int pad, void* stream);
int rpnet_maxpool_bwd_bf16(const void* dy_bf16, const void* idx_u8, void* dx_bf16, int n, int h, int w, int c, int k, int stride,
                           int pad, void* stream);

/* ---- ResNet18 backbone training (net/rp_net.py:19-42; torchvision BasicBlock: y = relu(bn2(conv2(relu(bn1(conv1(x))))) + identity)).
 * rpnet_bn_apply_res_f16: rpnet_bn_apply_split_f16 with the identity (hi plane + optional lo plane in `lo_fmt`) added before the ReLU.
 * rpnet_add_relu_mask_bf16: out = (a + b) where y > 0, else 0 (b, y optional; bf16 gradients, fp16 activation, `elems` a multiple
 *   of 8): the ReLU mask in front of both branches of a block and the sum of the two branches' input gradients.
 * rpnet_conv7x7s2_stem_wgrad: grad [64][3][7][7] += weight gradient of the stem conv (7x7, stride 2, padding 3) from the fp32 NCHW
 *   images and dz bf16 [n][(h-1)/2+1][(w-1)/2+1][64]; scratch9408: 64 * 147 doubles. */
int rpnet_bn_apply_res_f16(const void* z_hi, const void* z_lo, const float* stats, int n, int h, int w, int c, const int* group_start,
                           int groups, int relu, const void* res_f16, const void* res_lo, void* y_f16, void* y_lo_f16,
                           void* y_pool_f16, void* y_pool_lo_f16, float* y_f32, int lo_fmt, void* stream);
int rpnet_add_relu_mask_bf16(const void* a_bf16, const void* b_bf16, const void* y_f16, void* out_bf16, long long elems, void* stream);
int rpnet_conv7x7s2_stem_wgrad(const float* img, const void* dz_bf16, int n, int h, int w, float* grad, double* scratch9408,
                               void* stream);
int rpnet_conv3x3_first_wgrad_cin(const float* img, int cin, const void* dz_bf16, int n, int h, int w, float* grad,
                                  double* scratch576, void* stream);

/* `soft_mask: True` training (net/rp_net.py:308-311 without the threshold): the recurrent mask m_{i+1} = avg_pool2d(p_fg(logits_i), scale)
 * stays in the autograd graph.
 * rpnet_premask_mask_bwd: dmask[p] = sum_c (dxfg[p][c] - dxbg[p][c]) * x[p][c] — the gradient of x_fg = x * m, x_bg = x * (1 - m) (:283)
 *   w.r.t. m; dxfg / dxbg bf16 NHWC, x fp16 NHWC (`pixels` x c), dmask fp32 [pixels].
 * rpnet_soft_mask_bwd_f32: dlogits[b][k][Y][X] += dmask[b][Y/scale][X/scale] / scale^2 * softmax_k * ([k >= 1] - p_fg): through
 *   F.avg_pool2d and softmax(dim=1)[:, 1] (sum of the foreground classes for more than one way) into the logits of the previous
 *   iteration.  logits / dlogits fp32 [batch][classes][h*scale][w*scale], dmask fp32 [batch][h][w]. */
int rpnet_premask_mask_bwd(const void* dxfg_bf16, const void* dxbg_bf16, const void* x_f16, long long pixels, int c, float* dmask,
                           void* stream);
int rpnet_soft_mask_bwd_f32(const float* logits, const float* dmask, int batch, int n_classes, int h, int w, int scale,
                            float* dlogits, void* stream);

/* Backward of Correlation (net/rp_net.py:153-181).  dq_bf16 NHWC [n][h][w][ld]: channels [0,(2r+1)^2) = d corr,
 * [add_off, add_off+c) = the direct gradient of fm1 from cat([corr, fm1]) (net/rp_net.py:81), added into df1. */
/* workspace (optional): rpnet_local_corr_bwd_workspace_bytes(n, h, w, radius) bytes of scratch enable the tensor-core
 * band-GEMM path (c % 64 == 0, maps of at least (8+2r) x (16+2r) pixels); without it the CUDA-core kernels run. */
long long rpnet_local_corr_bwd_workspace_bytes(int n, int h, int w, int radius);
int rpnet_local_corr_bwd(const void* f1_f16, const void* f2_f16, const void* dq_bf16, int ld, int add_off, void* df1_bf16,
                         void* df2_bf16, int n, int h, int w, int c, int radius, void* workspace, long long workspace_bytes,
                         void* stream);

/* Backward of calDist (net/rp_net.py:353-363).  feat fp32 [n][hw][64]; protos fp32 [proto_sets][p][64], image i uses
 * set i % proto_sets; dpred fp32 [n][p][hw]; dfeat (= or += when accumulate); dprotos (may be null) = the prototype gradient,
 * accumulated in the fp64 scratch dprotos_acc [proto_sets][p][64] (required with dprotos). */
int rpnet_cos_sim_bwd_f32(const float* feat, const float* protos, const float* dpred, int n, int hw, int c, int n_protos,
                          int proto_sets, float scaler, float* dfeat, int accumulate, float* dprotos, double* dprotos_acc, void* stream);

/* Adjoint of F.interpolate(bilinear, align_corners=False): out[n][i][j] = sum_{Y,X} wy(Y,i) wx(X,j) in[n][Y][X].
 * (a) masked-average-pool weights U^T mask (net/rp_net.py:373-376), (b) backward of the logit upsample (:303,337).
 * sums (optional) [n] = sum of in[n]. */
int rpnet_bilinear_adjoint_f32(const float* in, float* out, float* sums, int n, int in_h, int in_w, int out_h, int out_w,
                               void* stream);

/* getFeatures for a fore and a back mask from their adjoint maps: out[n][k][c] = sum_p feat[n][p][c]*wmap_k[n][p] /
 * (msum_k[n] + 1e-5).  feat fp32 [n][hw][c], c <= 64.  net/rp_net.py:366-376. */
int rpnet_weighted_pool_f32(const float* feat, const float* wmap0, const float* wmap1, const float* msum0, const float* msum1,
                            float* out, int n, int hw, int c, void* stream);
int rpnet_weighted_pool_bwd_f32(const float* dout, const float* wmap0, const float* wmap1, const float* msum0,
                                const float* msum1, float* dfeat, int accumulate, int n, int hw, int c, void* stream);

/* Backward of getPrototype (net/rp_net.py:379-391): dprotos [batch][1+ways][c] -> draw [ways][shots][batch][2][c]. */
int rpnet_proto_finalize_bwd_f32(const float* dprotos, float* draw, int ways, int shots, int batch, int c, void* stream);

/* dice_ce (net/rp_net.py:87-127) for `groups` logit tensors sharing the labels: logits fp32 [groups][batch][classes][hw],
 * labels int64 [batch][hw]; loss[g] = dice + CE; dlogits (optional) = grad_scale * dloss_g/dlogits.
 * sums: fp64 scratch [groups][2*classes+1]. */
int rpnet_dice_ce_f32(const float* logits, const long long* labels, int groups, int batch, int n_classes, long long hw,
                      float eps, float grad_scale, double* sums, float* dlogits, float* loss, void* stream);

/* alignLoss pieces (net/rp_net.py:394-440).  class_pool: argmax over pred [batch][classes][hw] -> per-class masked mean
 * of feat [batch][hw][64] -> qproto [batch][classes][64], counts [batch][classes], amax int32 [batch][hw]. */
int rpnet_class_pool_f32(const float* feat, const float* pred, int batch, int hw, int c, int n_classes, float* qproto,
                         float* counts, void* amax_i32, void* stream);
int rpnet_class_pool_bwd_f32(const float* dqproto, const float* counts, const void* amax_i32, int batch, int hw, int c, int n_classes,
                             float* dfeat, void* stream);
/* protos_s[w][s][b] = [qproto[b][0], qproto[b][1+w]]; weight[w][s][b] = (counts[b][1+w] > 0) * scaler / (shots*ways*batch). */
int rpnet_align_gather_f32(const float* qproto, const float* counts, int ways, int shots, int batch, float scaler,
                           float* protos_s, float* weight, void* stream);
int rpnet_align_scatter_f32(const float* dprotos_s, int ways, int shots, int batch, float* dqproto, void* stream);
/* Cross entropy with ignore over 2-class logits [n][2][hw]; label = 1 where fore == 1, else 0 where back == 1, else
 * ignored; *loss = sum_n weight[n] * mean_valid(nll_n); dlogits (optional) = grad_scale * dloss.  sums: fp64 scratch [n][2]. */
int rpnet_ce_mask_f32(const float* logits, const float* fore, const float* back, const float* weight, int n, long long hw,
                      float grad_scale, double* sums, float* dlogits, float* loss, void* stream);
/* F.interpolate(bilinear, align_corners=False) of fp32 maps [n][h][w] -> [n][out_h][out_w]  (net/rp_net.py:430). */
int rpnet_bilinear_up_f32(const float* in, float* out, int n, int h, int w, int out_h, int out_w, void* stream);

/* torch.optim.Adam step (L2 weight decay added to the gradient; yamls/example.yml:64-67) on flat fp32 buffers;
 * the gradient is multiplied by grad_scale first (1/world_size after a sum all-reduce). */
int rpnet_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

/* ===================================================================================================
 * "Next" row N1: batched affine registration in front of the hot path.
 * =================================================================================================== */

/* AffineRegistration.train_registraion (net/registration.py:316-357) as driven by get_registration_field
 * (dataset/few_shot_reader.py:109-198) for n slice pairs in one launch: theta = identity, then `iters` times
 *   warped = F.grid_sample(moving, F.affine_grid(theta, size))   (bilinear, zero padding, align_corners=False)
 *   loss = mean((warped - fixed)^2); Adam(lr, beta1, beta2, eps).step() on the six parameters.
 * moving / fixed fp32 [n][h][w]; theta fp32 [n][2][3] (out); loss_curve (optional) fp32 [n][iters]. */
int rpnet_affine_register_f32(const float* moving, const float* fixed, int n, int h, int w, int iters, float lr, float beta1,
                              float beta2, float eps, float* theta, float* loss_curve, void* stream);

/* AffineRegistration.forward (net/registration.py:337-344): out[n][c] = grid_sample(x[n][c], affine_grid(theta[n])).
 * x, out fp32 [n][c][h][w]; theta fp32 [n][2][3]. */
int rpnet_affine_warp_f32(const float* x, const float* theta, float* out, int n, int c, int h, int w, void* stream);

/* NCC (net/registration.py:157-160; printed by the eval driver, test_rpnet.py:229-230) of two fp32 arrays of n elements:
 * -cov(f, m) / sqrt(var(f) * var(m) * n^2 ... + 1e-10) exactly as the reference's sums; scratch5: 5 doubles; out: 1 float. */
int rpnet_ncc_f32(const float* moving, const float* fixed, long long n, double* scratch5, float* out, void* stream);

/* Deformable half of get_registration_field (`do_deformable: True`, dataset/few_shot_reader.py:137-170): DemonsRegistration with
 * Diffeomorphic(10) — exp(flow) by scaling and squaring —, the NCC loss, torch.optim.Adam(lr) on the flow and the Gaussian
 * regulariser after every step (net/registration.py:16-160, 190-313), for ALL slices in one launch (one CTA per slice runs every
 * iteration: forward chain, NCC reduction, hand-derived backward, Adam, smoothing).
 *   moving (already affinely warped), fixed: fp32 [n][h][w] in [0, 1];  gauss_host: HOST array [gauss_h][gauss_w] = the regulariser's
 *   kernel (net/registration.py:16-51; 9 x 9 for sigma 2), odd sides up to 15;  scaling: number of compositions (10).
 *   flow [n][2][h][w] (out): the trained DemonsRegistration.flow (channel 0 = x, normalised units);  disp [n][2][h][w] (out):
 *   exp(flow), the displacement DemonsRegistration.forward adds to the identity grid;  loss_curve [n][iters] (optional).
 *   workspace: rpnet_demons_workspace_bytes(n, h, w, scaling) bytes.
 * rpnet_demons_warp_f32: DemonsRegistration.forward (:244-258) with that displacement: out = grid_sample(x, grid + disp), x / out
 *   fp32 [n][c][h][w], torch's grid_sample defaults (bilinear, zeros, align_corners=False) on the corner-aligned compute_grid. */
long long rpnet_demons_workspace_bytes(int n, int h, int w, int scaling);
int rpnet_demons_register_f32(const float* moving, const float* fixed, int n, int h, int w, int iters, float lr, float beta1,
                              float beta2, float eps, int scaling, const float* gauss_host, int gauss_h, int gauss_w,
                              float* flow, float* disp, float* loss_curve, void* workspace, long long workspace_bytes,
                              void* stream);
int rpnet_demons_warp_f32(const float* x, const float* disp, float* out, int n, int c, int h, int w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RPNET_B200_H_ */
