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
int rpnet_abi_version(void);

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
 * out_f32:   optional fp32 NHWC [n][h][w][cout]. */
int rpnet_conv_igemm_f16(const void* src0, int c0, const void* src1, int c1, int n, int h, int w,
                         const void* wpack, int ntaps, const int* tap_dy, const int* tap_dx, int cout,
                         const float* scale, const float* shift, int relu, void* out_f16, int out_h, int out_w,
                         int out_c, int out_coff, int oy_mul, int oy_off, int ox_mul, int ox_off,
                         void* out_pool_f16, float* out_f32, void* stream);

/* First encoder conv: fp32 NCHW image [n][cin][h][w] (cin 1 or 3) -> 64 channels, 3x3 pad 1, fused
 * affine (+ReLU), fp16 NHWC out [n][h][w][64].  weight fp32 [64][cin][3][3] (PyTorch layout).
 * Replaces encoder.Conv1.conv.0-2 (net/modules.py:48-50 via net/unet.py:405) and VGG features.0.0
 * (net/vgg.py:53-56). */
int rpnet_conv3x3_first_f16(const float* img, int n, int cin, int h, int w, const float* weight,
                            const float* scale, const float* shift, int relu, void* out_f16, void* stream);

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

#ifdef __cplusplus
}
#endif
#endif /* RPNET_B200_H_ */
