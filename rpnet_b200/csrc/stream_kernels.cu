// HBM-bound / CUDA-core kernels of the RP-Net hot path (everything that is not a dense contraction):
// first conv (Cin <= 4), mask pooling / pre-masking, local (2r+1)^2 correlation, masked-average-pool
// prototypes, cosine matching and the fused bilinear-upsample + softmax + threshold + avg-pool tail.
// Each kernel cites the reference op it replaces (paths relative to the reference root).
#include "common.cuh"

#include <cstdarg>
#include <cstring>
#include <string>

namespace rpnet {

// ---- error plumbing (C ABI: no exceptions cross the boundary) -------------------------------------
static thread_local std::string g_last_error;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return RPNET_ERR_CUDA;
}

__device__ __forceinline__ uint4 pack8(const float* v) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
  return u;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

// ---------------------------------------------------------------------------------------------------
// First conv of an encoder: fp32 NCHW image with Cin <= 4 -> 64 channels, 3x3 pad 1, fused
// scale/shift(+ReLU), fp16 NHWC out.  K = 9*Cin is far too small for tensor cores; the layer is a
// pure streaming write (128 B per pixel).  8 threads per pixel x 8 channels each => every warp store
// instruction writes 512 contiguous bytes.   Replaces net/modules.py:48-50 for encoder.Conv1.conv.0
// (Cin=1, net/unet.py:405) and net/vgg.py:53-56 for features.0.0 (Cin=3).
// ---------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256)
conv3x3_first_kernel(const float* __restrict__ img, const float* __restrict__ wgt /*[64][CIN][3][3]*/,
                     const float* __restrict__ scale, const float* __restrict__ shift, int relu, __half* __restrict__ out,
                     __half* __restrict__ out_lo, int lo_fmt, int N, int H, int W) {
  __shared__ float s_w[9 * CIN][64];      // [tap*CIN + ci][co]
  __shared__ float s_sc[64], s_sh[64];
  for (int i = threadIdx.x; i < 64 * CIN * 9; i += blockDim.x) {
    const int co = i / (CIN * 9), rem = i % (CIN * 9);
    const int ci = rem / 9, tap = rem % 9;
    s_w[tap * CIN + ci][co] = wgt[i];
  }
  if (threadIdx.x < 64) { s_sc[threadIdx.x] = scale[threadIdx.x]; s_sh[threadIdx.x] = shift[threadIdx.x]; }
  __syncthreads();
  const long long total = (long long)N * H * W * 8;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(gid & 7);
    const long long pix = gid >> 3;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float* plane = img + ((long long)n * CIN + ci) * H * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = x + kx - 1;
          const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(plane + (long long)yy * W + xx) : 0.f;
          const float* wrow = &s_w[(ky * 3 + kx) * CIN + ci][cg * 8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wrow[j], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = fmaf(acc[j], s_sc[cg * 8 + j], s_sh[cg * 8 + j]);
      acc[j] = relu ? fmaxf(t, 0.f) : t;
    }
    const uint4 hi = pack8(acc);
    *reinterpret_cast<uint4*>(out + pix * 64 + cg * 8) = hi;
    if (out_lo) lo8_store(out_lo, lo_fmt, (size_t)pix * 64 + cg * 8, cg * 8, acc, hi);
  }
}

// Cin == 1 specialisation (encoder.Conv1.conv.0, net/unet.py:405): the 8 output channels a thread owns never change
// (grid stride is a multiple of 8), so its 72 weights (scale folded in) live in registers; each thread produces two
// horizontally adjacent pixels from a 3 x 4 image patch: 12 loads, 144 FMAs, two 16-byte stores.
__global__ void __launch_bounds__(256)
conv3x3_first_c1_kernel(const float* __restrict__ img, const float* __restrict__ wgt /*[64][1][3][3]*/,
                        const float* __restrict__ scale, const float* __restrict__ shift, int relu, __half* __restrict__ out,
                        __half* __restrict__ out_lo, int lo_fmt, int N, int H, int W) {
  const int cg = threadIdx.x & 7;
  float wr[9][8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float sc = __ldg(scale + cg * 8 + j);
    sh[j] = __ldg(shift + cg * 8 + j);
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[t][j] = __ldg(wgt + (cg * 8 + j) * 9 + t) * sc;
  }
  const unsigned Wp = (unsigned)(W + 1) >> 1;                    // pixel pairs per row
  const unsigned total = (unsigned)N * H * Wp;                   // pixel pairs (< 2^31, host-checked)
  for (unsigned pp = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; pp < total; pp += (gridDim.x * blockDim.x) >> 3) {
    const int x = (int)(pp % Wp) * 2;
    const int y = (int)((pp / Wp) % (unsigned)H);
    const long long n = pp / (Wp * (unsigned)H);
    const float* plane = img + n * H * W;
    float v[3][4];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      const bool yok = yy >= 0 && yy < H;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int xx = x + kx - 1;
        v[ky][kx] = (yok && xx >= 0 && xx < W) ? __ldg(plane + (long long)yy * W + xx) : 0.f;
      }
    }
    float a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a0[j] = sh[j]; a1[j] = sh[j]; }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a0[j] = fmaf(v[ky][kx], wr[ky * 3 + kx][j], a0[j]);
          a1[j] = fmaf(v[ky][kx + 1], wr[ky * 3 + kx][j], a1[j]);
        }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { a0[j] = fmaxf(a0[j], 0.f); a1[j] = fmaxf(a1[j], 0.f); }
    }
    const long long pix = (n * H + y) * W + x;
    const uint4 h0 = pack8(a0), h1 = pack8(a1);
    *reinterpret_cast<uint4*>(out + pix * 64 + cg * 8) = h0;
    if (x + 1 < W) *reinterpret_cast<uint4*>(out + (pix + 1) * 64 + cg * 8) = h1;
    if (out_lo) {                                                 // split-fp16 residual plane
      lo8_store(out_lo, lo_fmt, (size_t)pix * 64 + cg * 8, cg * 8, a0, h0);
      if (x + 1 < W) lo8_store(out_lo, lo_fmt, (size_t)(pix + 1) * 64 + cg * 8, cg * 8, a1, h1);
    }
  }
}

// ResNet stem (torchvision resnet18.conv1 + bn1 + relu, reached through net/rp_net.py:19-37): 7x7 stride-2 pad-3 conv of a
// 3-channel fp32 NCHW image -> 64 channels, folded BN + ReLU, fp16 NHWC [n][H/2][W/2][64].  K = 147 per output: CUDA cores;
// weights live in shared memory as [tap][64], 8 threads per output pixel x 8 channels each.
__global__ void __launch_bounds__(256)
conv7x7s2_stem_kernel(const float* __restrict__ img, const float* __restrict__ wgt /*[64][3][7][7]*/, const float* __restrict__ scale,
                      const float* __restrict__ shift, int relu, __half* __restrict__ out, __half* __restrict__ out_lo, int lo_fmt, int N, int H, int W) {
  __shared__ float s_w[147][64];
  for (int i = threadIdx.x; i < 64 * 147; i += blockDim.x) s_w[i % 147][i / 147] = wgt[i];
  __syncthreads();
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  const int cg = threadIdx.x & 7;
  const unsigned total = (unsigned)N * Ho * Wo;
  for (unsigned p = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; p < total; p += (gridDim.x * blockDim.x) >> 3) {
    const int xo = (int)(p % Wo), yo = (int)((p / Wo) % Ho), n = (int)(p / ((unsigned)Wo * Ho));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int ci = 0; ci < 3; ++ci) {
      const float* plane = img + ((size_t)n * 3 + ci) * H * W;
      for (int ky = 0; ky < 7; ++ky) {
        const int yy = yo * 2 + ky - 3;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int xx = xo * 2 + kx - 3;
          const float v = (xx >= 0 && xx < W) ? __ldg(plane + (size_t)yy * W + xx) : 0.f;
          const float* wr = &s_w[(ci * 7 + ky) * 7 + kx][cg * 8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wr[j], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = fmaf(acc[j], __ldg(scale + cg * 8 + j), __ldg(shift + cg * 8 + j));
      acc[j] = relu ? fmaxf(t, 0.f) : t;
    }
    const uint4 hi = pack8(acc);
    *reinterpret_cast<uint4*>(out + (size_t)p * 64 + cg * 8) = hi;
    if (out_lo) lo8_store(out_lo, lo_fmt, (size_t)p * 64 + cg * 8, cg * 8, acc, hi);
  }
}

// ---------------------------------------------------------------------------------------------------
// F.avg_pool2d(mask[:, None], s)  (net/rp_net.py:270,272).  fp32 [N,H,W] -> fp32 [N,H/s,W/s].
// ---------------------------------------------------------------------------------------------------
__global__ void avgpool_mask_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W, int s) {
  const int ho = H / s, wo = W / s;
  const long long total = (long long)N * ho * wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % wo), y = (int)((i / wo) % ho), n = (int)(i / ((long long)wo * ho));
    const float* src = in + ((long long)n * H + (long long)y * s) * W + (long long)x * s;
    float acc = 0.f;
    for (int dy = 0; dy < s; ++dy)
      for (int dx = 0; dx < s; ++dx) acc += __ldg(src + (long long)dy * W + dx);
    out[i] = acc / (float)(s * s);
  }
}

// ---------------------------------------------------------------------------------------------------
// fts * mask and fts * (1 - mask)  (net/rp_net.py:275,283).  x fp16 NHWC [P, C], m fp32 [P].
// ---------------------------------------------------------------------------------------------------
__global__ void premask_kernel(const uint4* __restrict__ x, const float* __restrict__ m, uint4* __restrict__ xfg,
                               uint4* __restrict__ xbg, long long pixels, int c8 /* C/8 */) {
  const long long total = pixels * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float mk = __ldg(m + i / c8);
    const float mb = 1.f - mk;
    float v[8], a[8], b[8];
    unpack8(__ldg(x + i), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = v[j] * mk; b[j] = v[j] * mb; }
    xfg[i] = pack8(a);
    xbg[i] = pack8(b);
  }
}

// ---------------------------------------------------------------------------------------------------
// Local correlation (Correlation(), net/rp_net.py:153-181, restated as the zero-padded (2r+1)^2 window
// it actually computes — SURVEY D7):
//   out[n, y, x, a*(2r+1)+b] = 1/sqrt(C) * sum_c f1[n,y,x,c] * f2[n, y+(b-r), x+(a-r), c]
// (channel index: a <-> column offset, b <-> row offset — the RAFT x/y quirk).
// fp16 NHWC in, fp32 accumulate, fp16 NHWC out with `out_c` >= (2r+1)^2 channels (tail zero filled so
// the following 1x1 tcgen05 conv can treat it as a 64-aligned K range).
// Block = 8x16 pixel tile; thread = 4 consecutive pixels x all 2r+1 column offsets for one row offset.
// ---------------------------------------------------------------------------------------------------
constexpr int kCorrTH = 8, kCorrTW = 16, kCorrCC = 32;   // tile and channel chunk (fp16 channels)

template <int R>
__global__ void __launch_bounds__((2 * R + 1) * 32)
local_corr_kernel(const __half* __restrict__ f1, const __half* __restrict__ f2, __half* __restrict__ out, int N, int H, int W,
                  int C, int out_c, float scale) {
  constexpr int K = 2 * R + 1;
  constexpr int HW_ = kCorrTW + 2 * R, HH_ = kCorrTH + 2 * R;
  constexpr int HP = HW_ + 1;                         // padded halo row pitch (pixels)
  constexpr int kThreads = K * 32;
  extern __shared__ __align__(16) uint8_t smem_corr[];
  __half2* s2 = reinterpret_cast<__half2*>(smem_corr);              // [CC/2][HH_*HP]
  __half2* s1 = s2 + (kCorrCC / 2) * HH_ * HP;                        // [CC/2][128]
  const int tiles_x = (W + kCorrTW - 1) / kCorrTW, tiles_y = (H + kCorrTH - 1) / kCorrTH;
  const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y, n = blockIdx.x / (tiles_x * tiles_y);
  const int x0 = tx * kCorrTW, y0 = ty * kCorrTH;
  const int tid = threadIdx.x;
  const int dyi = tid >> 5;                            // row-offset index b (0..K-1)
  const int g = tid & 31;                              // pixel group: 8 rows x 4 groups of 4 pixels
  const int gy = g >> 2, gx = (g & 3) * 4;
  float acc[4][K];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int a = 0; a < K; ++a) acc[j][a] = 0.f;

  for (int cc = 0; cc < C; cc += kCorrCC) {
    __syncthreads();
    // f2 halo: HH_ x HW_ pixels x 32 channels (4 x uint4 per pixel), transposed to [c2][pixel]
    for (int i = tid; i < HH_ * HW_ * 4; i += kThreads) {
      const int part = i & 3, pix = i >> 2;
      const int hx = pix % HW_, hy = pix / HW_;
      const int gx_ = x0 + hx - R, gy_ = y0 + hy - R;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (gx_ >= 0 && gx_ < W && gy_ >= 0 && gy_ < H)
        u = __ldg(reinterpret_cast<const uint4*>(f2 + ((size_t)(n * H + gy_) * W + gx_) * C + cc + part * 8));
      const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) s2[(part * 4 + k) * (HH_ * HP) + hy * HP + hx] = hp[k];
    }
    for (int i = tid; i < kCorrTH * kCorrTW * 4; i += kThreads) {
      const int part = i & 3, pix = i >> 2;
      const int px = pix % kCorrTW, py = pix / kCorrTW;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (x0 + px < W && y0 + py < H)
        u = __ldg(reinterpret_cast<const uint4*>(f1 + ((size_t)(n * H + y0 + py) * W + x0 + px) * C + cc + part * 8));
      const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) s1[(part * 4 + k) * 128 + pix] = hp[k];
    }
    __syncthreads();
#pragma unroll 2
    for (int c2 = 0; c2 < kCorrCC / 2; ++c2) {
      float2 a1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) a1[j] = __half22float2(s1[c2 * 128 + gy * kCorrTW + gx + j]);
      const __half2* row = s2 + c2 * (HH_ * HP) + (gy + dyi) * HP + gx;
      float2 b2[K + 3];
#pragma unroll
      for (int t = 0; t < K + 3; ++t) b2[t] = __half22float2(row[t]);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int a = 0; a < K; ++a) acc[j][a] = fmaf(a1[j].x, b2[j + a].x, fmaf(a1[j].y, b2[j + a].y, acc[j][a]));
    }
  }
  __syncthreads();
  // stage the 128 x out_c fp16 output tile in shared memory, then coalesced 16-byte stores
  __half* so = reinterpret_cast<__half*>(smem_corr);                   // [128][out_c]
  for (int i = tid; i < 128 * out_c / 2; i += kThreads) reinterpret_cast<uint32_t*>(so)[i] = 0u;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int a = 0; a < K; ++a) so[(gy * kCorrTW + gx + j) * out_c + a * K + dyi] = __float2half_rn(acc[j][a] * scale);
  __syncthreads();
  const int vec_per_pix = out_c / 8;
  for (int i = tid; i < 128 * vec_per_pix; i += kThreads) {
    const int pix = i / vec_per_pix, v = i % vec_per_pix;
    const int px = pix % kCorrTW, py = pix / kCorrTW;
    if (x0 + px < W && y0 + py < H)
      *reinterpret_cast<uint4*>(out + ((size_t)(n * H + y0 + py) * W + x0 + px) * out_c + v * 8) =
          *reinterpret_cast<const uint4*>(so + pix * out_c + v * 8);
  }
}

// ---------------------------------------------------------------------------------------------------
// Masked average pooling (getFeatures, net/rp_net.py:366-376) without materialising the upsampled
// feature map:  sum_{Y,X} up(f)[c,Y,X] * m[Y,X]  ==  sum_{y,x} f[c,y,x] * (U^T m)[y,x]   (U = bilinear
// align_corners=False interpolation matrix), so one block builds the adjoint-pooled mask in shared
// memory and reduces the h' x w' feature map against it.  out[n][which][c], which in {0: m0, 1: m1}.
// feat fp32 NHWC [N, h, w, C] (C <= 64), masks fp32 [N, H, W].
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src(int o, float rscale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  // ATen upsample_bilinear2d, align_corners=False (area_pixel_compute_source_index, cubic=false)
  float src = rscale * (o + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  i0 = i0 < in_size - 1 ? i0 : in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256)
masked_avg_pool_kernel(const float* __restrict__ feat, const float* __restrict__ mask0, const float* __restrict__ mask1,
                       float* __restrict__ out, int h, int w, int C, int H, int W) {
  extern __shared__ float s_map[];            // [h*w] adjoint-pooled mask, then [4][64] partials
  __shared__ float s_red[256];
  const int n = blockIdx.x, which = blockIdx.y;
  const float* mask = (which == 0 ? mask0 : mask1) + (size_t)n * H * W;
  const int tid = threadIdx.x;
  const float rsy = (float)h / (float)H, rsx = (float)w / (float)W;
  const int sy = (H + h - 1) / h, sx = (W + w - 1) / w;
  // mask sum (fixed-order tree reduction => deterministic)
  float ms = 0.f;
  for (int i = tid; i < H * W; i += 256) ms += __ldg(mask + i);
  s_red[tid] = ms;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) s_red[tid] += s_red[tid + s];
    __syncthreads();
  }
  const float msum = s_red[0];
  __syncthreads();
  for (int o = tid; o < h * w; o += 256) {
    const int i = o / w, j = o % w;
    float acc = 0.f;
    const int Y0 = max(0, sy * (i - 1)), Y1 = min(H, sy * (i + 2));
    const int X0 = max(0, sx * (j - 1)), X1 = min(W, sx * (j + 2));
    for (int Y = Y0; Y < Y1; ++Y) {
      int i0, i1; float l0, l1;
      bilinear_src(Y, rsy, h, i0, i1, l0, l1);
      const float wy = (i0 == i ? l0 : 0.f) + (i1 == i ? l1 : 0.f);
      if (wy == 0.f) continue;
      float racc = 0.f;
      for (int X = X0; X < X1; ++X) {
        int j0, j1; float m0, m1;
        bilinear_src(X, rsx, w, j0, j1, m0, m1);
        const float wx = (j0 == j ? m0 : 0.f) + (j1 == j ? m1 : 0.f);
        racc = fmaf(wx, __ldg(mask + (size_t)Y * W + X), racc);
      }
      acc = fmaf(wy, racc, acc);
    }
    s_map[o] = acc;
  }
  __syncthreads();
  const int c = tid & 63, part = tid >> 6;
  float acc = 0.f;
  if (c < C) {
    const float* f = feat + (size_t)n * h * w * C + c;
    for (int o = part; o < h * w; o += 4) acc = fmaf(__ldg(f + (size_t)o * C), s_map[o], acc);
  }
  s_red[tid] = acc;
  __syncthreads();
  if (tid < 64 && tid < C) {
    const float t = (s_red[tid] + s_red[tid + 64]) + (s_red[tid + 128] + s_red[tid + 192]);
    out[((size_t)n * 2 + which) * C + tid] = t / (msum + 1e-5f);
  }
}

// getPrototype (net/rp_net.py:379-391): raw [Wa][Sh][B][2][C] (0 = fg, 1 = bg) -> protos [B][1+Wa][C]
__global__ void proto_finalize_kernel(const float* __restrict__ raw, float* __restrict__ protos, int Wa, int Sh, int B, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  float bg = 0.f;
  for (int w = 0; w < Wa; ++w) {
    float fg = 0.f, bgw = 0.f;
    for (int s = 0; s < Sh; ++s) {
      const float* r = raw + ((((size_t)w * Sh + s) * B + b) * 2) * C + c;
      fg += r[0];
      bgw += r[C];
    }
    protos[((size_t)b * (1 + Wa) + 1 + w) * C + c] = fg / (float)Sh;
    bg += bgw / (float)Sh;
  }
  protos[((size_t)b * (1 + Wa)) * C + c] = bg / (float)Wa;
}

// ---------------------------------------------------------------------------------------------------
// calDist (net/rp_net.py:353-363): pred[b][p][pix] = scaler * cos(feat[b,pix,:], proto[b][p][:]),
// torch semantics: each norm clamped at eps = 1e-8.  feat fp32 NHWC [B, hw, 64]; 16 lanes per pixel.
// ---------------------------------------------------------------------------------------------------
constexpr int kMaxProtos = 8;
__global__ void __launch_bounds__(256)
cos_sim_kernel(const float* __restrict__ feat, const float* __restrict__ protos, float* __restrict__ pred, int B, int hw, int P,
               float scaler) {
  __shared__ float s_p[kMaxProtos][64];
  __shared__ float s_pn[kMaxProtos];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < P * 64; i += blockDim.x) s_p[i / 64][i % 64] = protos[(size_t)b * P * 64 + i];
  __syncthreads();
  if (threadIdx.x < P) {
    float s = 0.f;
    for (int c = 0; c < 64; ++c) s = fmaf(s_p[threadIdx.x][c], s_p[threadIdx.x][c], s);
    s_pn[threadIdx.x] = fmaxf(sqrtf(s), 1e-8f);
  }
  __syncthreads();
  const int sub = threadIdx.x & 15;
  const int pix = blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4);
  const bool ok = pix < hw;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) v = __ldg(reinterpret_cast<const float4*>(feat + ((size_t)b * hw + pix) * 64) + sub);
  float nn = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  float dots[kMaxProtos];
#pragma unroll
  for (int p = 0; p < kMaxProtos; ++p) {
    dots[p] = 0.f;
    if (p < P) {
      const float* pp = &s_p[p][sub * 4];
      dots[p] = v.x * pp[0] + v.y * pp[1] + v.z * pp[2] + v.w * pp[3];
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    nn += __shfl_xor_sync(0xffffffffu, nn, o);
#pragma unroll
    for (int p = 0; p < kMaxProtos; ++p) dots[p] += __shfl_xor_sync(0xffffffffu, dots[p], o);
  }
  if (ok && sub == 0) {
    const float xn = fmaxf(sqrtf(nn), 1e-8f);
    for (int p = 0; p < P; ++p) pred[((size_t)b * P + p) * hw + pix] = scaler * (dots[p] / (xn * s_pn[p]));
  }
}

// ---------------------------------------------------------------------------------------------------
// Refinement tail (net/rp_net.py:303-312): bilinear upsample xS (align_corners=False) of the
// (1+Wa)-class prediction -> logits (kept: refinement[i]) -> softmax fg probability -> (>0.5 | soft)
// -> avg_pool2d(S) -> next query mask.  One thread per output row segment of S pixels: warp dy of the block owns row dy of 32
// consecutive S x S blocks (a warp stores 32 * S * 4 contiguous bytes per class), the S row sums of a block meet in shared
// memory.  The six source values a segment can touch (3 columns x 2 rows) are loaded once per class.
// pred fp32 [B][P][h][w]; logits fp32 [B][P][h*S][w*S]; mask_out fp32 [B][h][w].
// ---------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(32 * S)
upsample_tail_kernel(const float* __restrict__ pred, float* __restrict__ logits, float* __restrict__ mask_out, int B, int P, int h,
                     int w, int soft) {
  __shared__ float s_row[S][32];
  const int H = h * S, W = w * S;
  const long long total = (long long)B * h * w;
  const int lane = threadIdx.x & 31, dy = threadIdx.x >> 5;
  const long long idx = (long long)blockIdx.x * 32 + lane;
  const bool ok = idx < total;
  float msum = 0.f;
  if (ok) {
    const int j = (int)(idx % w), i = (int)((idx / w) % h), b = (int)(idx / ((long long)w * h));
    // exact xS upsample, align_corners=False (ATen area_pixel_compute_source_index): output d of block j reads
    // src = j + f(d), f(d) = (d + 0.5) / S - 0.5: f < 0 -> (j-1, j) with weights (-f, 1 + f), else (j, j+1) with (1 - f, f);
    // the clamp at 0 makes the first half of block 0 read (src[0], src[1]) with weights (1, 0).  All weights are exact in fp32.
    const float fy = ((float)dy + 0.5f) / (float)S - 0.5f;
    const bool up = fy < 0.f;
    const int i0 = up ? (i > 0 ? i - 1 : 0) : i;
    const int i1 = up ? (i > 0 ? i : (h > 1 ? 1 : 0)) : (i < h - 1 ? i + 1 : h - 1);
    const float ly1 = up ? (i > 0 ? 1.f + fy : 0.f) : fy, ly0 = 1.f - ly1;
    const int jm = j > 0 ? j - 1 : 0, jp = j < w - 1 ? j + 1 : w - 1;
    const bool left = j == 0;
    const int Y = i * S + dy;
    float val[kMaxProtos][S];
#pragma unroll
    for (int p = 0; p < kMaxProtos; ++p) {
      if (p < P) {
        const float* src = pred + ((size_t)b * P + p) * h * w;
        const float a0 = __ldg(src + i0 * w + jm), a1 = __ldg(src + i0 * w + j), a2 = __ldg(src + i0 * w + jp);
        const float c0 = __ldg(src + i1 * w + jm), c1 = __ldg(src + i1 * w + j), c2 = __ldg(src + i1 * w + jp);
#pragma unroll
        for (int dx = 0; dx < S; ++dx) {
          constexpr float inv = 1.f / (float)S;
          const float fx = ((float)dx + 0.5f) * inv - 0.5f;          // compile-time per dx
          float v;
          if (dx < S / 2) {
            const float lx1 = left ? 0.f : 1.f + fx, lx0 = 1.f - lx1;
            const float t0 = left ? a1 : a0, t1 = left ? a2 : a1, u0 = left ? c1 : c0, u1 = left ? c2 : c1;
            v = ly0 * (lx0 * t0 + lx1 * t1) + ly1 * (lx0 * u0 + lx1 * u1);
          } else {
            v = ly0 * ((1.f - fx) * a1 + fx * a2) + ly1 * ((1.f - fx) * c1 + fx * c2);
          }
          val[p][dx] = v;
        }
        float* dst = logits + (((size_t)b * P + p) * H + Y) * W + (size_t)j * S;
#pragma unroll
        for (int dx = 0; dx < S; dx += 4)
          __stcs(reinterpret_cast<float4*>(dst + dx), make_float4(val[p][dx], val[p][dx + 1], val[p][dx + 2], val[p][dx + 3]));
      }
    }
    // softmax(dim=1)[:, 1] for Wa == 1 (sum over the fg classes for Wa > 1: oracle-ext)
    if (P == 2) {
      // two classes: the larger logit contributes exp(0) = 1, so one expf and one division per pixel give the same bits
#pragma unroll
      for (int dx = 0; dx < S; ++dx) {
        const float d = val[1][dx] - val[0][dx];
        const float e = expf(-fabsf(d));
        const float den = 1.f + e;
        const float prob = (d > 0.f ? 1.f : e) / den;
        msum += soft ? prob : (prob > 0.5f ? 1.f : 0.f);
      }
    } else {
      float mx[S], den[S], fg[S];
#pragma unroll
      for (int dx = 0; dx < S; ++dx) { mx[dx] = -INFINITY; den[dx] = 0.f; fg[dx] = 0.f; }
#pragma unroll
      for (int p = 0; p < kMaxProtos; ++p)
        if (p < P) {
#pragma unroll
          for (int dx = 0; dx < S; ++dx) mx[dx] = fmaxf(mx[dx], val[p][dx]);
        }
#pragma unroll
      for (int p = 0; p < kMaxProtos; ++p) {
        if (p < P) {
#pragma unroll
          for (int dx = 0; dx < S; ++dx) {
            const float e = expf(val[p][dx] - mx[dx]);
            den[dx] += e;
            if (p >= 1) fg[dx] += e;
          }
        }
      }
#pragma unroll
      for (int dx = 0; dx < S; ++dx) {
        const float prob = fg[dx] / den[dx];
        msum += soft ? prob : (prob > 0.5f ? 1.f : 0.f);
      }
    }
  }
  s_row[dy][lane] = msum;
  __syncthreads();
  if (dy == 0 && ok) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < S; ++k) t += s_row[k][lane];
    mask_out[idx] = t / (float)(S * S);
  }
}


// ---------------------------------------------------------------------------------------------------
// nn.MaxPool2d(k, stride, padding) with implicit -inf padding on fp16 NHWC (VGG: k3 s2 p1 and k3 s1 p1,
// net/vgg.py:24-30).  One thread per output pixel x 8 channels.
// ---------------------------------------------------------------------------------------------------
__global__ void maxpool_f16_kernel(const uint4* __restrict__ in, const uint4* __restrict__ in_lo, uint4* __restrict__ out,
                                   uint4* __restrict__ out_lo, uint2* __restrict__ idx, int lo_fmt, int N, int H, int W, int c8, int Ho, int Wo, int k,
                                   int stride, int pad) {
  const long long total = (long long)N * Ho * Wo * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % c8);
    long long pix = i / c8;
    const int xo = (int)(pix % Wo), yo = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    float best[8];
    unsigned arg[8];                       // window position dy * k + dx of the FIRST maximum (nn.MaxPool2d's backward routing)
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = 0u; }
    for (int dy = 0; dy < k; ++dy) {
      const int y = yo * stride - pad + dy;
      if (y < 0 || y >= H) continue;
      for (int dx = 0; dx < k; ++dx) {
        const int x = xo * stride - pad + dx;
        if (x < 0 || x >= W) continue;
        float v[8];
        unpack8(__ldg(in + ((long long)(n * H + y) * W + x) * c8 + cv), v);
        if (in_lo)                                // split-fp16: the value is hi + lo (exact in fp32; c8 plane: lo8 * 2^-11)
          lo8_add(in_lo, lo_fmt, (size_t)(((long long)(n * H + y) * W + x) * c8 + cv) * 8, (cv * 8) & 63, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v[j] > best[j]) { best[j] = v[j]; arg[j] = (unsigned)(dy * k + dx); }
      }
    }
    const uint4 hi = pack8(best);
    out[i] = hi;
    if (out_lo) lo8_store(out_lo, lo_fmt, (size_t)i * 8, (cv * 8) & 63, best, hi);
    if (idx) idx[i] = make_uint2(arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24), arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24));
  }
}

static int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace rpnet

using namespace rpnet;

RPNET_API const char* rpnet_last_error(void) { return g_last_error.c_str(); }

RPNET_API int rpnet_abi_version(void) { return 8; }

RPNET_API int rpnet_conv3x3_first_split_f16(const float* img, int n, int cin, int h, int w, const float* weight, const float* scale,
                                             const float* shift, int relu, void* out_f16, void* out_lo_f16, int lo_fmt, void* stream_);

RPNET_API int rpnet_conv3x3_first_f16(const float* img, int n, int cin, int h, int w, const float* weight, const float* scale,
                                       const float* shift, int relu, void* out_f16, void* stream_) {
  return rpnet_conv3x3_first_split_f16(img, n, cin, h, w, weight, scale, shift, relu, out_f16, nullptr, 0, stream_);
}

RPNET_API int rpnet_conv3x3_first_split_f16(const float* img, int n, int cin, int h, int w, const float* weight, const float* scale,
                                             const float* shift, int relu, void* out_f16, void* out_lo_f16, int lo_fmt, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(img && weight && scale && shift && out_f16, "conv3x3_first: null pointer argument");
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0, "conv3x3_first: bad shape %d x %d x %d", n, h, w);
  RPNET_REQUIRE(cin >= 1 && cin <= 3, "conv3x3_first: cin must be 1, 2 or 3 (got %d)", cin);
  RPNET_REQUIRE(lo_fmt == 0 || lo_fmt == 1, "conv3x3_first: lo_fmt %d", lo_fmt);
  const long long total = (long long)n * h * w * 8;
  const int grid = grid_for(total, 256);
  if (cin == 1)
    conv3x3_first_c1_kernel<<<grid_for((long long)n * h * ((w + 1) / 2) * 8, 256), 256, 0, stream>>>(
        img, weight, scale, shift, relu, static_cast<__half*>(out_f16), static_cast<__half*>(out_lo_f16), lo_fmt, n, h, w);
  else if (cin == 2)      // mask_feature_map: x (net/unet.py:401-402, 437-438): image + mask channel
    conv3x3_first_kernel<2><<<grid, 256, 0, stream>>>(img, weight, scale, shift, relu, static_cast<__half*>(out_f16),
                                                      static_cast<__half*>(out_lo_f16), lo_fmt, n, h, w);
  else
    conv3x3_first_kernel<3><<<grid, 256, 0, stream>>>(img, weight, scale, shift, relu, static_cast<__half*>(out_f16),
                                                      static_cast<__half*>(out_lo_f16), lo_fmt, n, h, w);
  return check_cuda(cudaGetLastError(), "conv3x3_first launch");
}

RPNET_API int rpnet_conv7x7s2_stem_split_f16(const float* img, int n, int h, int w, const float* weight, const float* scale,
                                              const float* shift, int relu, void* out_f16, void* out_lo_f16, int lo_fmt, void* stream_);

RPNET_API int rpnet_conv7x7s2_stem_f16(const float* img, int n, int h, int w, const float* weight, const float* scale,
                                        const float* shift, int relu, void* out_f16, void* stream_) {
  return rpnet_conv7x7s2_stem_split_f16(img, n, h, w, weight, scale, shift, relu, out_f16, nullptr, 0, stream_);
}

RPNET_API int rpnet_conv7x7s2_stem_split_f16(const float* img, int n, int h, int w, const float* weight, const float* scale,
                                              const float* shift, int relu, void* out_f16, void* out_lo_f16, int lo_fmt, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(img && weight && scale && shift && out_f16, "conv7x7s2_stem: null pointer argument");
  RPNET_REQUIRE(n > 0 && h >= 7 && w >= 7 && (long long)n * h * w < (1LL << 31), "conv7x7s2_stem: bad shape %d x %d x %d", n, h, w);
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  conv7x7s2_stem_kernel<<<grid_for((long long)n * ho * wo * 8, 256), 256, 0, stream>>>(img, weight, scale, shift, relu,
                                                                                    static_cast<__half*>(out_f16),
                                                                                    static_cast<__half*>(out_lo_f16), lo_fmt, n, h, w);
  return check_cuda(cudaGetLastError(), "conv7x7s2_stem launch");
}

RPNET_API int rpnet_avgpool_mask_f32(const float* in, float* out, int n, int h, int w, int s, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(in && out, "avgpool_mask: null pointer argument");
  RPNET_REQUIRE(n > 0 && s > 0 && h >= s && w >= s, "avgpool_mask: bad shape n=%d h=%d w=%d s=%d", n, h, w, s);
  const long long total = (long long)n * (h / s) * (w / s);
  avgpool_mask_kernel<<<grid_for(total, 256), 256, 0, stream>>>(in, out, n, h, w, s);
  return check_cuda(cudaGetLastError(), "avgpool_mask launch");
}

RPNET_API int rpnet_premask_f16(const void* x, const float* mask, void* x_fg, void* x_bg, long long pixels, int c,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(x && mask && x_fg && x_bg, "premask: null pointer argument");
  RPNET_REQUIRE(pixels > 0 && c > 0 && c % 8 == 0, "premask: bad shape pixels=%lld c=%d", pixels, c);
  premask_kernel<<<grid_for(pixels * (c / 8), 256), 256, 0, stream>>>(static_cast<const uint4*>(x), mask, static_cast<uint4*>(x_fg),
                                                                        static_cast<uint4*>(x_bg), pixels, c / 8);
  return check_cuda(cudaGetLastError(), "premask launch");
}

namespace rpnet {
int relation_head_tc(const void* f1, const void* f2, const void* wq, const float* scale, const float* shift, const float* protos,
                     int P, int sets, float cos_scaler, float* pred, int n, int h, int w, int c, int radius, cudaStream_t stream);
int local_corr_tc(const void* f1, const void* f2, void* out, int n, int h, int w, int c, int radius, int out_c, cudaStream_t stream);
}

template <int R>
static int launch_corr(const void* f1, const void* f2, void* out, int n, int h, int w, int c, int out_c, cudaStream_t stream) {
  constexpr int K = 2 * R + 1;
  constexpr int HW_ = kCorrTW + 2 * R, HH_ = kCorrTH + 2 * R, HP = HW_ + 1;
  size_t smem = (size_t)(kCorrCC / 2) * HH_ * HP * 4 + (size_t)(kCorrCC / 2) * 128 * 4;
  const size_t smem_out = (size_t)128 * out_c * 2;
  if (smem_out > smem) smem = smem_out;
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(local_corr_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  RPNET_REQUIRE(smem <= 96 * 1024, "local_corr: shared memory %zu too large", smem);
  const int tiles = ((w + kCorrTW - 1) / kCorrTW) * ((h + kCorrTH - 1) / kCorrTH) * n;
  local_corr_kernel<R><<<tiles, K * 32, smem, stream>>>(static_cast<const __half*>(f1), static_cast<const __half*>(f2),
                                                        static_cast<__half*>(out), n, h, w, c, out_c, 1.0f / sqrtf((float)c));
  return check_cuda(cudaGetLastError(), "local_corr launch");
}

RPNET_API int rpnet_local_corr_f16(const void* f1, const void* f2, void* out, int n, int h, int w, int c, int radius,
                                    int out_c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(f1 && f2 && out, "local_corr: null pointer argument");
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % kCorrCC == 0, "local_corr: bad shape n=%d h=%d w=%d c=%d (c %% 32 == 0)", n, h, w, c);
  const int k = 2 * radius + 1;
  RPNET_REQUIRE(out_c >= k * k && out_c % 8 == 0, "local_corr: out_c=%d must be >= %d and a multiple of 8", out_c, k * k);
  if (c % 64 == 0 && w >= 8 + 2 * radius && h >= 16 + 2 * radius) {      // tcgen05 banded-GEMM path (local_corr_tc.cu)
    const int rc = local_corr_tc(f1, f2, out, n, h, w, c, radius, out_c, stream);
    if (rc <= 0) return rc;
  }
  switch (radius) {
    case 1: return launch_corr<1>(f1, f2, out, n, h, w, c, out_c, stream);
    case 2: return launch_corr<2>(f1, f2, out, n, h, w, c, out_c, stream);
    case 3: return launch_corr<3>(f1, f2, out, n, h, w, c, out_c, stream);
    case 4: return launch_corr<4>(f1, f2, out, n, h, w, c, out_c, stream);
    case 5: return launch_corr<5>(f1, f2, out, n, h, w, c, out_c, stream);
    default:
      set_error("local_corr: radius %d not supported (1..5)", radius);
      return RPNET_ERR_ARG;
  }
}

RPNET_API int rpnet_relation_head_f16(const void* f1, const void* f2, const void* wq_pack, const float* scale, const float* shift,
                                       const float* protos, int n_protos, int proto_sets, float scaler, float* pred, int n, int h,
                                       int w, int c, int radius, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(f1 && f2 && wq_pack && scale && shift && protos && pred, "relation_head: null pointer argument");
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0, "relation_head: bad shape");
  if (w >= 8 + 2 * radius && h >= 16 + 2 * radius) {
    const int rc = relation_head_tc(f1, f2, wq_pack, scale, shift, protos, n_protos, proto_sets, scaler, pred, n, h, w, c, radius, stream);
    if (rc <= 0) return rc;
  }
  set_error("relation_head: shape not supported by the fused kernel (c %% 64 == 0, c <= 256, radius 3 or 5, maps >= (8+2r) x (16+2r)): "
            "run rpnet_local_corr_f16 + rpnet_conv_cos_f16 instead");
  return RPNET_ERR_ARG;
}

RPNET_API int rpnet_masked_avg_pool_f32(const float* feat, const float* mask0, const float* mask1, float* out, int n, int h,
                                         int w, int c, int mask_h, int mask_w, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(feat && mask0 && mask1 && out, "masked_avg_pool: null pointer argument");
  RPNET_REQUIRE(n > 0 && c > 0 && c <= 64 && h > 0 && w > 0, "masked_avg_pool: bad shape n=%d h=%d w=%d c=%d (c <= 64)", n, h, w, c);
  RPNET_REQUIRE(mask_h >= h && mask_w >= w, "masked_avg_pool: mask %d x %d smaller than features %d x %d", mask_h, mask_w, h, w);
  const size_t smem = (size_t)h * w * 4;
  RPNET_REQUIRE(smem <= 160 * 1024, "masked_avg_pool: feature map %d x %d too large", h, w);
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(masked_avg_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  masked_avg_pool_kernel<<<dim3(n, 2), 256, smem, stream>>>(feat, mask0, mask1, out, h, w, c, mask_h, mask_w);
  return check_cuda(cudaGetLastError(), "masked_avg_pool launch");
}

RPNET_API int rpnet_proto_finalize_f32(const float* raw, float* protos, int ways, int shots, int batch, int c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(raw && protos, "proto_finalize: null pointer argument");
  RPNET_REQUIRE(ways > 0 && shots > 0 && batch > 0 && c > 0, "proto_finalize: bad shape");
  proto_finalize_kernel<<<(batch * c + 127) / 128, 128, 0, stream>>>(raw, protos, ways, shots, batch, c);
  return check_cuda(cudaGetLastError(), "proto_finalize launch");
}

RPNET_API int rpnet_cos_sim_f32(const float* feat, const float* protos, float* pred, int batch, int hw, int c, int n_protos,
                                 float scaler, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(feat && protos && pred, "cos_sim: null pointer argument");
  RPNET_REQUIRE(c == 64, "cos_sim: feature width must be 64 (got %d)", c);
  RPNET_REQUIRE(n_protos >= 1 && n_protos <= kMaxProtos, "cos_sim: n_protos %d out of range [1, %d]", n_protos, kMaxProtos);
  RPNET_REQUIRE(batch > 0 && hw > 0, "cos_sim: bad shape");
  cos_sim_kernel<<<dim3((hw + 15) / 16, batch), 256, 0, stream>>>(feat, protos, pred, batch, hw, n_protos, scaler);
  return check_cuda(cudaGetLastError(), "cos_sim launch");
}

RPNET_API int rpnet_upsample_tail_f32(const float* pred, float* logits, float* mask_out, int batch, int n_protos, int h, int w,
                                       int scale, int soft_mask, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(pred && logits && mask_out, "upsample_tail: null pointer argument");
  RPNET_REQUIRE(n_protos >= 2 && n_protos <= kMaxProtos, "upsample_tail: n_protos %d out of range [2, %d]", n_protos, kMaxProtos);
  RPNET_REQUIRE(batch > 0 && h > 0 && w > 0, "upsample_tail: bad shape");
  const long long total = (long long)batch * h * w;
  const int grid = (int)((total + 31) / 32);
  if (scale == 4)
    upsample_tail_kernel<4><<<grid, 128, 0, stream>>>(pred, logits, mask_out, batch, n_protos, h, w, soft_mask);
  else if (scale == 8)
    upsample_tail_kernel<8><<<grid, 256, 0, stream>>>(pred, logits, mask_out, batch, n_protos, h, w, soft_mask);
  else {
    set_error("upsample_tail: scale %d not supported (4 or 8)", scale);
    return RPNET_ERR_ARG;
  }
  return check_cuda(cudaGetLastError(), "upsample_tail launch");
}

RPNET_API int rpnet_maxpool_split_f16(const void* in, const void* in_lo, void* out, void* out_lo, int lo_fmt, int n, int h, int w, int c,
                                       int k, int stride, int pad, void* stream_);
RPNET_API int rpnet_maxpool_idx_f16(const void* in, const void* in_lo, void* out, void* out_lo, int lo_fmt, void* idx_u8, int n, int h, int w,
                                     int c, int k, int stride, int pad, void* stream_);

RPNET_API int rpnet_maxpool_f16(const void* in, void* out, int n, int h, int w, int c, int k, int stride, int pad, void* stream_) {
  return rpnet_maxpool_split_f16(in, nullptr, out, nullptr, 0, n, h, w, c, k, stride, pad, stream_);
}

RPNET_API int rpnet_maxpool_split_f16(const void* in, const void* in_lo, void* out, void* out_lo, int lo_fmt, int n, int h, int w, int c,
                                       int k, int stride, int pad, void* stream_) {
  return rpnet_maxpool_idx_f16(in, in_lo, out, out_lo, lo_fmt, nullptr, n, h, w, c, k, stride, pad, stream_);
}

RPNET_API int rpnet_maxpool_idx_f16(const void* in, const void* in_lo, void* out, void* out_lo, int lo_fmt, void* idx_u8, int n, int h, int w,
                                     int c, int k, int stride, int pad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(k <= 15, "maxpool: window %d too large for the uint8 argmax positions", k);
  RPNET_REQUIRE(in && out, "maxpool: null pointer argument");
  RPNET_REQUIRE((in_lo == nullptr) == (out_lo == nullptr), "maxpool: residual planes go in and out together");
  RPNET_REQUIRE(lo_fmt == 0 || (lo_fmt == 1 && c % 64 == 0), "maxpool: lo_fmt %d (c8 planes need c %% 64 == 0, c = %d)", lo_fmt, c);
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "maxpool: bad shape n=%d h=%d w=%d c=%d", n, h, w, c);
  RPNET_REQUIRE(k >= 1 && stride >= 1 && pad >= 0 && 2 * pad <= k, "maxpool: bad window k=%d stride=%d pad=%d", k, stride, pad);
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  const long long total = (long long)n * ho * wo * (c / 8);
  maxpool_f16_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const uint4*>(in), static_cast<const uint4*>(in_lo),
                                                                static_cast<uint4*>(out), static_cast<uint4*>(out_lo),
                                                                static_cast<uint2*>(idx_u8), lo_fmt, n, h, w, c / 8, ho, wo, k, stride, pad);
  return check_cuda(cudaGetLastError(), "maxpool launch");
}
