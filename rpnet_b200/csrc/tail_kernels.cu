// Training kernels of the prototype-match / refinement tail and the losses (CUDA cores, fp32 math):
// local-correlation backward, cosine-similarity backward, the bilinear adjoint (masked-average-pool
// weights and upsample backward), weighted pooling fwd/bwd, prototype averaging backward, dice_ce and the
// PANet-style alignment loss pieces.  Each kernel cites the reference forward op whose gradient (taken by
// torch autograd in the reference) it restates.
#include <cooperative_groups.h>

#include "common.cuh"

namespace rpnet {

static int grid_for(long long total, int block, int cap_mult = 16) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * cap_mult;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// ---------------------------------------------------------------------------------------------------
// Backward of the local correlation (Correlation, net/rp_net.py:153-181; forward in stream_kernels.cu):
//   corr[p, a*K+b] = s * sum_c f1[p,c] * f2[p + (b-r, a-r), c]          (row offset b-r, column offset a-r)
//   d f1[p,c]  = s * sum_{a,b} dcorr[p, a*K+b] * f2[p + (b-r, a-r), c]  (+ add[p,c]: the direct path of cat([corr, fm1]))
//   d f2[q,c]  = s * sum_{a,b} dcorr[q - (b-r, a-r), a*K+b] * f1[q - (b-r, a-r), c]
// dq: bf16 NHWC [n,h,w,ld]; channels [0,K*K) = dcorr, [add_off, add_off+C) = direct gradient of fm1.
// Block = 8x16 pixel tile, 512 threads = 128 pixels x 4 groups of 8 channels; 32-channel chunks staged in smem.
// ---------------------------------------------------------------------------------------------------
constexpr int kCbTH = 8, kCbTW = 16, kCbCC = 32;

template <int R, int WHICH>     // WHICH 1: d f1, 2: d f2
__global__ void __launch_bounds__(512)
local_corr_bwd_kernel(const __half* __restrict__ f1, const __half* __restrict__ f2, const __nv_bfloat16* __restrict__ dq, int ld,
                      int add_off, __nv_bfloat16* __restrict__ out, int N, int H, int W, int C, float scale) {
  constexpr int K = 2 * R + 1, KK = K * K;
  constexpr int HW_ = kCbTW + 2 * R, HH_ = kCbTH + 2 * R, NH = HW_ * HH_;
  extern __shared__ __align__(16) uint8_t smem_cb[];
  // WHICH 1: dcorr of the tile, fp32 [128][KK+?]; WHICH 2: dcorr of the halo, bf16 [NH][KKP]
  constexpr int KKP = (KK + 1) | 1;                       // odd pitch (elements)
  float* s_dc32 = reinterpret_cast<float*>(smem_cb);
  __nv_bfloat16* s_dc16 = reinterpret_cast<__nv_bfloat16*>(smem_cb);
  constexpr size_t dc_bytes = WHICH == 1 ? (size_t)128 * KKP * 4 : (((size_t)NH * KKP * 2 + 15) & ~(size_t)15);
  uint4* s_f = reinterpret_cast<uint4*>(smem_cb + dc_bytes);   // halo features [NH][4] uint4 (32 fp16 channels per pixel)

  const int tiles_x = (W + kCbTW - 1) / kCbTW, tiles_y = (H + kCbTH - 1) / kCbTH;
  const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y, n = blockIdx.x / (tiles_x * tiles_y);
  const int x0 = tx * kCbTW, y0 = ty * kCbTH;
  const int tid = threadIdx.x;
  const int g = tid & 3, pix = tid >> 2;
  const int px = pix % kCbTW, py = pix / kCbTW;
  const bool inside = (x0 + px < W) && (y0 + py < H);

  if (WHICH == 1) {
    for (int i = tid; i < 128 * KK; i += 512) {
      const int p = i / KK, k = i % KK;
      const int gx = x0 + p % kCbTW, gy = y0 + p / kCbTW;
      float v = 0.f;
      if (gx < W && gy < H) v = __bfloat162float(dq[((size_t)(n * H + gy) * W + gx) * ld + k]) * scale;
      s_dc32[p * KKP + k] = v;
    }
  } else {
    for (int i = tid; i < NH * KK; i += 512) {
      const int p = i / KK, k = i % KK;
      const int gx = x0 + p % HW_ - R, gy = y0 + p / HW_ - R;
      __nv_bfloat16 v = __float2bfloat16_rn(0.f);
      if (gx >= 0 && gx < W && gy >= 0 && gy < H) v = dq[((size_t)(n * H + gy) * W + gx) * ld + k];
      s_dc16[p * KKP + k] = v;
    }
  }
  const __half* fsrc = WHICH == 1 ? f2 : f1;
  for (int cc = 0; cc < C; cc += kCbCC) {
    __syncthreads();
    for (int i = tid; i < NH * 4; i += 512) {
      const int part = i & 3, p = i >> 2;
      const int gx = x0 + p % HW_ - R, gy = y0 + p / HW_ - R;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (gx >= 0 && gx < W && gy >= 0 && gy < H)
        u = __ldg(reinterpret_cast<const uint4*>(fsrc + ((size_t)(n * H + gy) * W + gx) * C + cc + part * 8));
      s_f[i] = u;
    }
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int a = 0; a < K; ++a) {
#pragma unroll
      for (int b = 0; b < K; ++b) {
        int hp;
        float dc;
        if (WHICH == 1) {
          hp = (py + b) * HW_ + (px + a);                               // p + (b-r, a-r) in halo coordinates
          dc = s_dc32[pix * KKP + a * K + b];
        } else {
          hp = (py + 2 * R - b) * HW_ + (px + 2 * R - a);               // q - (b-r, a-r) in halo coordinates
          dc = __bfloat162float(s_dc16[hp * KKP + a * K + b]);
        }
        float f[8];
        unpack8_f16(s_f[hp * 4 + g], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(dc, f[j], acc[j]);
      }
    }
    if (inside) {
      const size_t o = ((size_t)(n * H + y0 + py) * W + x0 + px);
      if (WHICH == 1) {
        float d[8];
        unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(dq + o * ld + add_off + cc + g * 8)), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += d[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= scale;
      }
      *reinterpret_cast<uint4*>(out + o * C + cc + g * 8) = pack8_bf16(acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Backward of calDist (net/rp_net.py:353-363): pred = scaler * <x/max(|x|,eps), p/max(|p|,eps)>.
// feat fp32 [n][hw][64]; protos fp32 [n_p][P][64] with image i using prototype set (i % n_p);
// dpred fp32 [n][P][hw].  dfeat (=|+=), dprotos += (atomics; caller zeroes).
// ---------------------------------------------------------------------------------------------------
constexpr int kMaxP = 8;
__global__ void __launch_bounds__(256)
cos_sim_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ protos, const float* __restrict__ dpred, int hw, int P,
                   int n_p, float scaler, float* __restrict__ dfeat, int accumulate, double* __restrict__ dprotos) {
  __shared__ float s_p[kMaxP][64];
  __shared__ float s_pn[kMaxP];
  __shared__ float s_dp[16][kMaxP][64];
  const int b = blockIdx.y;
  const int pb = b % n_p;
  for (int i = threadIdx.x; i < P * 64; i += blockDim.x) s_p[i / 64][i % 64] = protos[(size_t)pb * P * 64 + i];
  __syncthreads();
  if (threadIdx.x < P) {
    float s = 0.f;
    for (int c = 0; c < 64; ++c) s = fmaf(s_p[threadIdx.x][c], s_p[threadIdx.x][c], s);
    s_pn[threadIdx.x] = sqrtf(s);
  }
  __syncthreads();
  const int sub = threadIdx.x & 15, grp = threadIdx.x >> 4;
  float dpa[kMaxP][4];
#pragma unroll
  for (int p = 0; p < kMaxP; ++p)
#pragma unroll
    for (int j = 0; j < 4; ++j) dpa[p][j] = 0.f;
  for (int pix = blockIdx.x * 16 + grp; pix < hw; pix += gridDim.x * 16) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(feat + ((size_t)b * hw + pix) * 64) + sub);
    const float x[4] = {v.x, v.y, v.z, v.w};
    float nn = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    float dots[kMaxP];
#pragma unroll
    for (int p = 0; p < kMaxP; ++p) {
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
      for (int p = 0; p < kMaxP; ++p) dots[p] += __shfl_xor_sync(0xffffffffu, dots[p], o);
    }
    const float xn = sqrtf(nn);
    const float xnc = fmaxf(xn, 1e-8f);
    float dx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int p = 0; p < kMaxP; ++p) {
      if (p < P) {
        const float gp = scaler * __ldg(dpred + ((size_t)b * P + p) * hw + pix);
        const float pn = s_pn[p], pnc = fmaxf(pn, 1e-8f);
        const float inv = 1.f / (xnc * pnc);
        const float kx = xn >= 1e-8f ? dots[p] / (xn * xn * xn * pnc) : 0.f;
        const float kp = pn >= 1e-8f ? dots[p] / (xnc * pn * pn * pn) : 0.f;
        const float* pp = &s_p[p][sub * 4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dx[j] += gp * (pp[j] * inv - kx * x[j]);
          dpa[p][j] += gp * (x[j] * inv - kp * pp[j]);
        }
      }
    }
    float4* d = reinterpret_cast<float4*>(dfeat + ((size_t)b * hw + pix) * 64) + sub;
    if (accumulate) {
      const float4 o = *d;
      *d = make_float4(o.x + dx[0], o.y + dx[1], o.z + dx[2], o.w + dx[3]);
    } else {
      *d = make_float4(dx[0], dx[1], dx[2], dx[3]);
    }
  }
  if (dprotos) {
#pragma unroll
    for (int p = 0; p < kMaxP; ++p)
      if (p < P)
#pragma unroll
        for (int j = 0; j < 4; ++j) s_dp[grp][p][sub * 4 + j] = dpa[p][j];
    __syncthreads();
    for (int i = threadIdx.x; i < P * 64; i += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) s += s_dp[k][i / 64][i % 64];
      atomicAdd(dprotos + (size_t)pb * P * 64 + i, (double)s);    // fp32 partials in fp64: exact, order-independent
    }
  }
}

__global__ void cvt_f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

// ---------------------------------------------------------------------------------------------------
// Adjoint of F.interpolate(., size=(H,W), mode='bilinear', align_corners=False):
//   out[n][i][j] = sum_{Y,X} wy(Y,i) * wx(X,j) * in[n][Y][X]
// Used (a) on the support masks: getFeatures' sum_{Y,X} up(f)*m == sum_{y,x} f * (U^T m)  (net/rp_net.py:373-376),
// (b) as the backward of the logits upsample (net/rp_net.py:303,337).  Optional sums[n] = sum of in[n] (needs H = S*h).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilin_src(int o, float rscale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = rscale * (o + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  i0 = i0 < in_size - 1 ? i0 : in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(128)
bilinear_adjoint_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ sums, int H, int W, int h, int w) {
  const int n = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const float* src = in + (size_t)n * H * W;
  const float rsy = (float)h / (float)H, rsx = (float)w / (float)W;
  const int sy = (H + h - 1) / h, sx = (W + w - 1) / w;
  float own = 0.f;
  if (o < h * w) {
    const int i = o / w, j = o % w;
    float acc = 0.f;
    const int Y0 = max(0, sy * (i - 1)), Y1 = min(H, sy * (i + 2));
    const int X0 = max(0, sx * (j - 1)), X1 = min(W, sx * (j + 2));
    for (int Y = Y0; Y < Y1; ++Y) {
      int i0, i1; float l0, l1;
      bilin_src(Y, rsy, h, i0, i1, l0, l1);
      const float wy = (i0 == i ? l0 : 0.f) + (i1 == i ? l1 : 0.f);
      const bool mine_y = (Y >= sy * i) && (Y < sy * (i + 1));
      if (wy == 0.f && !(sums && mine_y)) continue;
      float racc = 0.f;
      for (int X = X0; X < X1; ++X) {
        int j0, j1; float m0, m1;
        bilin_src(X, rsx, w, j0, j1, m0, m1);
        const float wx = (j0 == j ? m0 : 0.f) + (j1 == j ? m1 : 0.f);
        const float v = __ldg(src + (size_t)Y * W + X);
        racc = fmaf(wx, v, racc);
        if (mine_y && X >= sx * j && X < sx * (j + 1)) own += v;
      }
      acc = fmaf(wy, racc, acc);
    }
    out[(size_t)n * h * w + o] = acc;
  }
  if (sums) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) own += __shfl_xor_sync(0xffffffffu, own, s);
    if ((threadIdx.x & 31) == 0 && own != 0.f) atomicAdd(sums + n, own);
  }
}

// Integer scale S (every use on the hot path: S = 4): output (i, j) can only receive from the 2S x 2S input window starting at
// (S*i - S/2, S*j - S/2) — i0 == i for inputs [S*i + S/2, S*i + 3S/2), i1 == i for [S*i - S/2, S*i + S/2), and the border
// clamps keep their inputs inside that window — so the column weights are computed once per thread and the row weight once
// per input row instead of one source-index computation per (output, input) pair.
template <int S>
__global__ void __launch_bounds__(128)
bilinear_adjoint_int_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ sums, int H, int W, int h, int w) {
  const int n = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const float* src = in + (size_t)n * H * W;
  const float rs = 1.f / (float)S;
  float own = 0.f;
  if (o < h * w) {
    const int i = o / w, j = o % w;
    const int X0 = S * j - S / 2, Y0 = S * i - S / 2;
    float wx[2 * S];
#pragma unroll
    for (int t = 0; t < 2 * S; ++t) {
      const int X = X0 + t;
      int j0, j1; float m0, m1;
      bilin_src(X < 0 ? 0 : X, rs, w, j0, j1, m0, m1);
      wx[t] = (X >= 0 && X < W) ? (j0 == j ? m0 : 0.f) + (j1 == j ? m1 : 0.f) : 0.f;
    }
    float acc = 0.f;
#pragma unroll 2
    for (int ty = 0; ty < 2 * S; ++ty) {
      const int Y = Y0 + ty;
      if (Y < 0 || Y >= H) continue;
      int i0, i1; float l0, l1;
      bilin_src(Y, rs, h, i0, i1, l0, l1);
      const float wy = (i0 == i ? l0 : 0.f) + (i1 == i ? l1 : 0.f);
      const bool mine_y = sums != nullptr && ty >= S / 2 && ty < S / 2 + S;
      const float* row = src + (size_t)Y * W + X0;
      float racc = 0.f, rown = 0.f;
#pragma unroll
      for (int t = 0; t < 2 * S; ++t) {
        const float v = (X0 + t >= 0 && X0 + t < W) ? __ldg(row + t) : 0.f;
        racc = fmaf(wx[t], v, racc);
        if (t >= S / 2 && t < S / 2 + S) rown += v;
      }
      acc = fmaf(wy, racc, acc);
      if (mine_y) own += rown;
    }
    out[(size_t)n * h * w + o] = acc;
  }
  if (sums) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) own += __shfl_xor_sync(0xffffffffu, own, s);
    if ((threadIdx.x & 31) == 0 && own != 0.f) atomicAdd(sums + n, own);
  }
}

// out[n][k][c] = sum_p feat[n][p][c] * wmap_k[n][p] / (msum_k[n] + 1e-5), k in {0,1}   (getFeatures for fore & back mask)
// One block of 1024 threads per image: 16 lanes x float4 cover the (<= 64) channels of a pixel, 64 pixels per step, both
// masks in the same pass over feat; ordered shared-memory tree (deterministic).
__global__ void __launch_bounds__(1024)
weighted_pool_kernel(const float* __restrict__ feat, const float* __restrict__ wmap0, const float* __restrict__ wmap1,
                     const float* __restrict__ msum0, const float* __restrict__ msum1, float* __restrict__ out, int hw, int C) {
  __shared__ float4 s_red[2][1024];
  const int n = blockIdx.x;
  const int v = threadIdx.x & 15, part = threadIdx.x >> 4;         // channel quad, pixel lane (64 lanes)
  const float* w0 = wmap0 + (size_t)n * hw;
  const float* w1 = wmap1 + (size_t)n * hw;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  if (v * 4 < C) {
    const float* f = feat + (size_t)n * hw * C + v * 4;
    for (int o = part; o < hw; o += 64) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(f + (size_t)o * C));
      const float m0 = __ldg(w0 + o), m1 = __ldg(w1 + o);
      a0.x = fmaf(x.x, m0, a0.x); a0.y = fmaf(x.y, m0, a0.y); a0.z = fmaf(x.z, m0, a0.z); a0.w = fmaf(x.w, m0, a0.w);
      a1.x = fmaf(x.x, m1, a1.x); a1.y = fmaf(x.y, m1, a1.y); a1.z = fmaf(x.z, m1, a1.z); a1.w = fmaf(x.w, m1, a1.w);
    }
  }
  s_red[0][threadIdx.x] = a0;
  s_red[1][threadIdx.x] = a1;
  __syncthreads();
  for (int s = 32; s > 0; s >>= 1) {
    if (part < s) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float4 x = s_red[k][threadIdx.x];
        const float4 y = s_red[k][threadIdx.x + s * 16];
        x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
        s_red[k][threadIdx.x] = x;
      }
    }
    __syncthreads();
  }
  if (part == 0 && v * 4 < C) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float inv = 1.f / ((k == 0 ? msum0 : msum1)[n] + 1e-5f);
      const float4 x = s_red[k][v];
      float* o = out + ((size_t)n * 2 + k) * C + v * 4;
      o[0] = x.x * inv; if (v * 4 + 1 < C) o[1] = x.y * inv; if (v * 4 + 2 < C) o[2] = x.z * inv; if (v * 4 + 3 < C) o[3] = x.w * inv;
    }
  }
}

// dfeat[n][p][c] (=|+=) sum_k dout[n][k][c] * wmap_k[n][p] / (msum_k[n] + 1e-5)
__global__ void weighted_pool_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ wmap0, const float* __restrict__ wmap1,
                                         const float* __restrict__ msum0, const float* __restrict__ msum1, float* __restrict__ dfeat,
                                         int accumulate, int N, int hw, int C) {
  const int c4 = C / 4;
  const long long total = (long long)N * hw * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % c4);
    const long long np = i / c4;
    const int n = (int)(np / hw);
    const float w0 = __ldg(wmap0 + np) / (__ldg(msum0 + n) + 1e-5f);
    const float w1 = __ldg(wmap1 + np) / (__ldg(msum1 + n) + 1e-5f);
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(dout + ((size_t)n * 2) * C) + v);
    const float4 d1 = __ldg(reinterpret_cast<const float4*>(dout + ((size_t)n * 2 + 1) * C) + v);
    float4 r = make_float4(d0.x * w0 + d1.x * w1, d0.y * w0 + d1.y * w1, d0.z * w0 + d1.z * w1, d0.w * w0 + d1.w * w1);
    float4* d = reinterpret_cast<float4*>(dfeat) + i;
    if (accumulate) { const float4 o = *d; r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
    *d = r;
  }
}

// Backward of getPrototype (net/rp_net.py:379-391): draw[w][s][b][0=fg] = dprotos[b][1+w]/Sh, [1=bg] = dprotos[b][0]/(Sh*Wa)
__global__ void proto_finalize_bwd_kernel(const float* __restrict__ dprotos, float* __restrict__ draw, int Wa, int Sh, int B, int C) {
  const long long total = (long long)Wa * Sh * B * 2 * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int k = (int)((i / C) % 2);
    const int b = (int)((i / (2 * C)) % B);
    const int w = (int)(i / ((long long)2 * C * B * Sh));
    draw[i] = k == 0 ? dprotos[((size_t)b * (1 + Wa) + 1 + w) * C + c] / (float)Sh
                     : dprotos[((size_t)b * (1 + Wa)) * C + c] / (float)(Sh * Wa);
  }
}

// ---------------------------------------------------------------------------------------------------
// dice_ce (net/rp_net.py:87-127), multi-class branch, for G independent logit tensors that share the labels:
//   loss_g = 1 - mean_k( 2 I_k / (Card_k + eps) ) + mean_pix( -log softmax(logits)[label] )
// logits fp32 [G][B][P][HW], labels int64 [B][HW].  sums[g] = {I_0..I_{P-1}, Card_0..Card_{P-1}, ce_sum}.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dice_ce_reduce_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int P, long long HW,
                      double* __restrict__ sums) {
  __shared__ float s_red[8][2 * kMaxP + 1];
  const int g = blockIdx.y;
  const float* lg = logits + (size_t)g * B * P * HW;
  float I[kMaxP], Cd[kMaxP], ce = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxP; ++k) { I[k] = 0.f; Cd[k] = 0.f; }
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, pix = i % HW;
    const int lab = (int)__ldg(labels + i);
    float v[kMaxP], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kMaxP; ++k)
      if (k < P) { v[k] = __ldg(lg + ((size_t)b * P + k) * HW + pix); mx = fmaxf(mx, v[k]); }
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxP; ++k)
      if (k < P) { v[k] = expf(v[k] - mx); den += v[k]; }
    const float inv = 1.f / den;
#pragma unroll
    for (int k = 0; k < kMaxP; ++k)
      if (k < P) {
        const float pk = v[k] * inv;
        const float oh = (k == lab) ? 1.f : 0.f;
        I[k] += pk * oh;
        Cd[k] += pk + oh;
        if (k == lab) ce -= logf(fmaxf(pk, 1e-38f));
      }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < kMaxP; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      I[k] += __shfl_xor_sync(0xffffffffu, I[k], o);
      Cd[k] += __shfl_xor_sync(0xffffffffu, Cd[k], o);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ce += __shfl_xor_sync(0xffffffffu, ce, o);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kMaxP; ++k) { s_red[warp][k] = I[k]; s_red[warp][kMaxP + k] = Cd[k]; }
    s_red[warp][2 * kMaxP] = ce;
  }
  __syncthreads();
  if (threadIdx.x < 2 * kMaxP + 1) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += s_red[wv][threadIdx.x];
    const int k = threadIdx.x;
    // fp32 block partials accumulated in fp64 (exact for fp32 addends: the result does not depend on the block order)
    if (k < kMaxP) { if (k < P) atomicAdd(sums + (size_t)g * (2 * P + 1) + k, (double)s); }
    else if (k < 2 * kMaxP) { if (k - kMaxP < P) atomicAdd(sums + (size_t)g * (2 * P + 1) + P + (k - kMaxP), (double)s); }
    else atomicAdd(sums + (size_t)g * (2 * P + 1) + 2 * P, (double)s);
  }
}

// dlogits = grad_scale * d loss_g / d logits;  loss[g] written by block (0, g).
__global__ void __launch_bounds__(256)
dice_ce_grad_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, const double* __restrict__ sums, int B,
                    int P, long long HW, float eps, float grad_scale, float* __restrict__ dlogits, float* __restrict__ loss) {
  const int g = blockIdx.y;
  const float* lg = logits + (size_t)g * B * P * HW;
  float* dl = dlogits + (size_t)g * B * P * HW;
  const double* sm = sums + (size_t)g * (2 * P + 1);
  const float npix = (float)((long long)B * HW);
  float cA[kMaxP], cB[kMaxP];        // d dice / d p_k = -(1/P) * (2 oh_k / (Card_k+eps) - 2 I_k / (Card_k+eps)^2) = oh_k * cA[k] + cB[k]
  float dice = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxP; ++k) {
    cA[k] = 0.f; cB[k] = 0.f;
    if (k < P) {
      const float I = (float)sm[k], Cd = (float)sm[P + k] + eps;
      cA[k] = -2.f / (Cd * (float)P);
      cB[k] = 2.f * I / (Cd * Cd * (float)P);
      dice += 2.f * I / Cd;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && loss) loss[g] = 1.f - dice / (float)P + (float)sm[2 * P] / npix;
  const long long total = dlogits ? (long long)B * HW : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, pix = i % HW;
    const int lab = (int)__ldg(labels + i);
    float v[kMaxP], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kMaxP; ++k)
      if (k < P) { v[k] = __ldg(lg + ((size_t)b * P + k) * HW + pix); mx = fmaxf(mx, v[k]); }
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxP; ++k)
      if (k < P) { v[k] = expf(v[k] - mx); den += v[k]; }
    const float inv = 1.f / den;
    float dot = 0.f, gk[kMaxP];
#pragma unroll
    for (int k = 0; k < kMaxP; ++k)
      if (k < P) {
        v[k] *= inv;
        gk[k] = ((k == lab) ? cA[k] : 0.f) + cB[k];
        dot += gk[k] * v[k];
      }
#pragma unroll
    for (int k = 0; k < kMaxP; ++k)
      if (k < P) {
        const float oh = (k == lab) ? 1.f : 0.f;
        dl[((size_t)b * P + k) * HW + pix] = grad_scale * (v[k] * (gk[k] - dot) + (v[k] - oh) / npix);
      }
  }
}

// ---------------------------------------------------------------------------------------------------
// Alignment loss pieces (alignLoss, net/rp_net.py:394-440).
// (1) class_pool: argmax over the low-resolution prediction -> per-class masked average of the query features.
// ---------------------------------------------------------------------------------------------------
// One thread-block cluster of kCpCluster CTAs per image: every CTA reduces a contiguous pixel range in a fixed order into its
// own shared memory, rank 0 then adds the partials rank by rank through distributed shared memory (deterministic, no scratch).
constexpr int kCpCluster = 8;

__global__ void __cluster_dims__(kCpCluster, 1, 1) __launch_bounds__(256)
class_pool_kernel(const float* __restrict__ feat /*[B][hw][64]*/, const float* __restrict__ pred /*[B][P][hw]*/, int hw, int P,
                  float* __restrict__ qproto /*[B][P][64]*/, float* __restrict__ counts /*[B][P]*/, int* __restrict__ amax /*[B][hw]*/) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float s_acc[4][kMaxP][64];
  __shared__ float s_part[kMaxP][64];
  __shared__ float s_cnt[kMaxP];
  const int b = blockIdx.y;
  const int rank = (int)cluster.block_rank();
  const int chunk = (hw + kCpCluster - 1) / kCpCluster;
  const int p_lo = rank * chunk, p_hi = min(hw, p_lo + chunk);
  for (int i = threadIdx.x; i < 4 * kMaxP * 64; i += 256) (&s_acc[0][0][0])[i] = 0.f;
  if (threadIdx.x < kMaxP) s_cnt[threadIdx.x] = 0.f;
  __syncthreads();
  for (int p = p_lo + threadIdx.x; p < p_hi; p += 256) {
    int best = 0;
    float bv = __ldg(pred + ((size_t)b * P) * hw + p);
    for (int k = 1; k < P; ++k) {
      const float v = __ldg(pred + ((size_t)b * P + k) * hw + p);
      if (v > bv) { bv = v; best = k; }
    }
    amax[(size_t)b * hw + p] = best;
    atomicAdd(&s_cnt[best], 1.f);                       // integer-valued: exact in any order
  }
  __syncthreads();
  const int c = threadIdx.x & 63, part = threadIdx.x >> 6;
  for (int p = p_lo + part; p < p_hi; p += 4) {
    const int k = amax[(size_t)b * hw + p];
    s_acc[part][k][c] += __ldg(feat + ((size_t)b * hw + p) * 64 + c);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P * 64; i += 256) {
    const int k = i / 64, cc = i % 64;
    s_part[k][cc] = (s_acc[0][k][cc] + s_acc[1][k][cc]) + (s_acc[2][k][cc] + s_acc[3][k][cc]);
  }
  cluster.sync();
  if (rank == 0) {
    for (int i = threadIdx.x; i < P * 64; i += 256) {
      const int k = i / 64, cc = i % 64;
      float t = 0.f, n = 0.f;
      for (int r = 0; r < kCpCluster; ++r) {
        t += cluster.map_shared_rank(&s_part[0][0], r)[k * 64 + cc];
        n += cluster.map_shared_rank(&s_cnt[0], r)[k];
      }
      qproto[((size_t)b * P + k) * 64 + cc] = t / (n + 1e-5f);
      if (cc == 0) counts[(size_t)b * P + k] = n;
    }
  }
  cluster.sync();                                        // the partials stay mapped until rank 0 has read them
}

__global__ void class_pool_bwd_kernel(const float* __restrict__ dqproto, const float* __restrict__ counts, const int* __restrict__ amax,
                                      int B, int hw, int P, float* __restrict__ dfeat) {
  const long long total = (long long)B * hw * 16;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i & 15);
    const long long bp = i >> 4;
    const int b = (int)(bp / hw);
    const int k = amax[bp];
    const float inv = 1.f / (counts[(size_t)b * P + k] + 1e-5f);
    const float4 d = __ldg(reinterpret_cast<const float4*>(dqproto + ((size_t)b * P + k) * 64) + v);
    float4* o = reinterpret_cast<float4*>(dfeat) + i;
    const float4 cur = *o;
    *o = make_float4(cur.x + d.x * inv, cur.y + d.y * inv, cur.z + d.z * inv, cur.w + d.w * inv);
  }
}

// (2) per support image (w, s, b): prototype pair [query bg proto, query fg_w proto] and the loss weight
//     active(w, b) * scaler / (Sh * Wa * B)   (skip_ways: a way whose class is absent from the prediction, :414).
__global__ void align_gather_kernel(const float* __restrict__ qproto, const float* __restrict__ counts, int Wa, int Sh, int B, float scaler,
                                    float* __restrict__ protos_s /*[Wa][Sh][B][2][64]*/, float* __restrict__ weight /*[Wa][Sh][B]*/) {
  const int total = Wa * Sh * B * 2 * 64;
  const int P = 1 + Wa;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % 64, k = (i / 64) % 2, b = (i / 128) % B, w = i / (128 * B * Sh);
    protos_s[i] = qproto[((size_t)b * P + (k == 0 ? 0 : w + 1)) * 64 + c];
    if (c == 0 && k == 0) weight[i / 128] = (counts[(size_t)b * P + w + 1] > 0.f ? 1.f : 0.f) * scaler / (float)(Sh * Wa * B);
  }
}

__global__ void align_scatter_kernel(const float* __restrict__ dprotos_s, int Wa, int Sh, int B, float* __restrict__ dqproto /*[B][P][64]*/) {
  const int P = 1 + Wa;
  const int total = B * P * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % 64, k = (i / 64) % P, b = i / (64 * P);
    float s = 0.f;
    for (int w = 0; w < Wa; ++w) {
      if (k != 0 && k != w + 1) continue;
      for (int sh = 0; sh < Sh; ++sh) s += dprotos_s[((((size_t)w * Sh + sh) * B + b) * 2 + (k == 0 ? 0 : 1)) * 64 + c];
    }
    dqproto[i] = s;
  }
}

// (3) cross entropy with ignore_index over 2-class logits [n][2][HW]; label 1 where fore == 1, else 0 where back == 1,
//     else ignored (net/rp_net.py:432-438).  sums[n] = {sum nll, valid count}.
__global__ void __launch_bounds__(256)
ce_mask_reduce_kernel(const float* __restrict__ logits, const float* __restrict__ fore, const float* __restrict__ back, long long HW,
                      double* __restrict__ sums) {
  __shared__ float s_red[8][2];
  const int n = blockIdx.y;
  float nll = 0.f, cnt = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    const float f = __ldg(fore + (size_t)n * HW + i), bk = __ldg(back + (size_t)n * HW + i);
    const int lab = (f == 1.f) ? 1 : ((bk == 1.f) ? 0 : 255);
    if (lab == 255) continue;
    const float l0 = __ldg(logits + ((size_t)n * 2) * HW + i), l1 = __ldg(logits + ((size_t)n * 2 + 1) * HW + i);
    const float mx = fmaxf(l0, l1);
    const float lse = mx + logf(expf(l0 - mx) + expf(l1 - mx));
    nll += lse - (lab ? l1 : l0);
    cnt += 1.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nll += __shfl_xor_sync(0xffffffffu, nll, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5][0] = nll; s_red[threadIdx.x >> 5][1] = cnt; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += s_red[wv][threadIdx.x];
    atomicAdd(sums + (size_t)n * 2 + threadIdx.x, (double)s);      // fp32 partials in fp64: exact, order-independent
  }
}

__global__ void __launch_bounds__(256)
ce_mask_grad_kernel(const float* __restrict__ logits, const float* __restrict__ fore, const float* __restrict__ back,
                    const double* __restrict__ sums, const float* __restrict__ weight, int N, long long HW, float grad_scale,
                    float* __restrict__ dlogits, float* __restrict__ loss) {
  const int n = blockIdx.y;
  const float cnt = (float)sums[(size_t)n * 2 + 1];
  const float wgt = weight[n];
  const float k = (cnt > 0.f) ? grad_scale * wgt / cnt : 0.f;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && loss) {
    float s = 0.f;
    for (int i = 0; i < N; ++i) {
      const float c = (float)sums[(size_t)i * 2 + 1];
      if (weight[i] != 0.f && c > 0.f) s += weight[i] * (float)sums[(size_t)i * 2] / c;
    }
    *loss = s;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    const float f = __ldg(fore + (size_t)n * HW + i), bk = __ldg(back + (size_t)n * HW + i);
    const int lab = (f == 1.f) ? 1 : ((bk == 1.f) ? 0 : 255);
    float d0 = 0.f, d1 = 0.f;
    if (lab != 255 && k != 0.f) {
      const float l0 = __ldg(logits + ((size_t)n * 2) * HW + i), l1 = __ldg(logits + ((size_t)n * 2 + 1) * HW + i);
      const float mx = fmaxf(l0, l1);
      const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
      const float inv = 1.f / (e0 + e1);
      d0 = k * (e0 * inv - (lab == 0 ? 1.f : 0.f));
      d1 = k * (e1 * inv - (lab == 1 ? 1.f : 0.f));
    }
    dlogits[((size_t)n * 2) * HW + i] = d0;
    dlogits[((size_t)n * 2 + 1) * HW + i] = d1;
  }
}

// Plain bilinear upsample of fp32 maps [n][h][w] -> [n][h*S][w*S] (align loss: supp_pred upsample, :430)
__global__ void bilinear_up_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int h, int w, int H, int W) {
  const long long total = (long long)N * H * W;
  const float rsy = (float)h / (float)H, rsx = (float)w / (float)W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % W), Y = (int)((i / W) % H);
    const long long n = i / ((long long)W * H);
    int i0, i1, j0, j1; float ly0, ly1, lx0, lx1;
    bilin_src(Y, rsy, h, i0, i1, ly0, ly1);
    bilin_src(X, rsx, w, j0, j1, lx0, lx1);
    const float* s = in + n * h * w;
    out[i] = ly0 * (lx0 * __ldg(s + i0 * w + j0) + lx1 * __ldg(s + i0 * w + j1)) +
             ly1 * (lx0 * __ldg(s + i1 * w + j0) + lx1 * __ldg(s + i1 * w + j1));
  }
}

}  // namespace rpnet

using namespace rpnet;

template <int R>
static int launch_corr_bwd(const void* f1, const void* f2, const void* dq, int ld, int add_off, void* df1, void* df2, int n, int h,
                           int w, int c, cudaStream_t stream) {
  constexpr int K = 2 * R + 1, KK = K * K, KKP = (KK + 1) | 1;
  constexpr int NH = (kCbTW + 2 * R) * (kCbTH + 2 * R);
  const size_t smem1 = (size_t)128 * KKP * 4 + (size_t)NH * 64;
  const size_t smem2 = (((size_t)NH * KKP * 2 + 15) & ~(size_t)15) + (size_t)NH * 64;
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(local_corr_bwd_kernel<R, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    RPNET_CUDA_OK(cudaFuncSetAttribute(local_corr_bwd_kernel<R, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    attr_set = true;
  }
  const int tiles = ((w + kCbTW - 1) / kCbTW) * ((h + kCbTH - 1) / kCbTH) * n;
  const float scale = 1.0f / sqrtf((float)c);
  local_corr_bwd_kernel<R, 1><<<tiles, 512, smem1, stream>>>(static_cast<const __half*>(f1), static_cast<const __half*>(f2),
                                                             static_cast<const __nv_bfloat16*>(dq), ld, add_off,
                                                             static_cast<__nv_bfloat16*>(df1), n, h, w, c, scale);
  RPNET_CUDA_OK(cudaGetLastError());
  local_corr_bwd_kernel<R, 2><<<tiles, 512, smem2, stream>>>(static_cast<const __half*>(f1), static_cast<const __half*>(f2),
                                                             static_cast<const __nv_bfloat16*>(dq), ld, add_off,
                                                             static_cast<__nv_bfloat16*>(df2), n, h, w, c, scale);
  return check_cuda(cudaGetLastError(), "local_corr_bwd launch");
}

namespace rpnet {
int local_corr_bwd_tc(const void* f1, const void* f2, const void* dq, int ld, int add_off, void* df1, void* df2, void* scratch,
                      int n, int h, int w, int c, int radius, cudaStream_t stream);
}

RPNET_API long long rpnet_local_corr_bwd_workspace_bytes(int n, int h, int w, int radius) {
  (void)radius;
  return (long long)n * h * w * 128 * 2;        // re-indexed correlation gradient, rows padded to 128 bf16
}

RPNET_API int rpnet_local_corr_bwd(const void* f1_f16, const void* f2_f16, const void* dq_bf16, int ld, int add_off, void* df1_bf16,
                                    void* df2_bf16, int n, int h, int w, int c, int radius, void* workspace,
                                    long long workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(f1_f16 && f2_f16 && dq_bf16 && df1_bf16 && df2_bf16, "local_corr_bwd: null pointer argument");
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % kCbCC == 0, "local_corr_bwd: bad shape n=%d h=%d w=%d c=%d", n, h, w, c);
  const int k = 2 * radius + 1;
  RPNET_REQUIRE(ld % 8 == 0 && add_off % 8 == 0 && add_off >= k * k && add_off + c <= ld,
                "local_corr_bwd: bad gradient layout ld=%d add_off=%d", ld, add_off);
  if (workspace && workspace_bytes >= rpnet_local_corr_bwd_workspace_bytes(n, h, w, radius) && c % 64 == 0 && w >= 8 + 2 * radius &&
      h >= 16 + 2 * radius) {                                       // tcgen05 band-GEMM path (local_corr_tc.cu)
    const int rc = local_corr_bwd_tc(f1_f16, f2_f16, dq_bf16, ld, add_off, df1_bf16, df2_bf16, workspace, n, h, w, c, radius, stream);
    if (rc <= 0) return rc;
  }
  switch (radius) {
    case 1: return launch_corr_bwd<1>(f1_f16, f2_f16, dq_bf16, ld, add_off, df1_bf16, df2_bf16, n, h, w, c, stream);
    case 2: return launch_corr_bwd<2>(f1_f16, f2_f16, dq_bf16, ld, add_off, df1_bf16, df2_bf16, n, h, w, c, stream);
    case 3: return launch_corr_bwd<3>(f1_f16, f2_f16, dq_bf16, ld, add_off, df1_bf16, df2_bf16, n, h, w, c, stream);
    case 4: return launch_corr_bwd<4>(f1_f16, f2_f16, dq_bf16, ld, add_off, df1_bf16, df2_bf16, n, h, w, c, stream);
    case 5: return launch_corr_bwd<5>(f1_f16, f2_f16, dq_bf16, ld, add_off, df1_bf16, df2_bf16, n, h, w, c, stream);
    default:
      set_error("local_corr_bwd: radius %d not supported (1..5)", radius);
      return RPNET_ERR_ARG;
  }
}

RPNET_API int rpnet_cos_sim_bwd_f32(const float* feat, const float* protos, const float* dpred, int n, int hw, int c, int n_protos,
                                     int proto_sets, float scaler, float* dfeat, int accumulate, float* dprotos, double* dprotos_acc,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(feat && protos && dpred && dfeat, "cos_sim_bwd: null pointer argument");
  RPNET_REQUIRE(!dprotos || dprotos_acc, "cos_sim_bwd: dprotos needs the fp64 accumulator scratch");
  RPNET_REQUIRE(c == 64, "cos_sim_bwd: feature width must be 64 (got %d)", c);
  RPNET_REQUIRE(n_protos >= 1 && n_protos <= kMaxP, "cos_sim_bwd: n_protos %d out of range [1, %d]", n_protos, kMaxP);
  RPNET_REQUIRE(n > 0 && hw > 0 && proto_sets > 0 && n % proto_sets == 0, "cos_sim_bwd: bad shape n=%d sets=%d", n, proto_sets);
  int bx = (hw + 16 * 8 - 1) / (16 * 8);
  if (bx < 1) bx = 1;
  const int np = proto_sets * n_protos * 64;
  if (dprotos) RPNET_CUDA_OK(cudaMemsetAsync(dprotos_acc, 0, (size_t)np * sizeof(double), stream));
  cos_sim_bwd_kernel<<<dim3(bx, n), 256, 0, stream>>>(feat, protos, dpred, hw, n_protos, proto_sets, scaler, dfeat, accumulate,
                                                      dprotos ? dprotos_acc : nullptr);
  RPNET_CUDA_OK(cudaGetLastError());
  if (dprotos) cvt_f64_to_f32_kernel<<<(np + 255) / 256, 256, 0, stream>>>(dprotos_acc, dprotos, np);
  return check_cuda(cudaGetLastError(), "cos_sim_bwd launch");
}

RPNET_API int rpnet_bilinear_adjoint_f32(const float* in, float* out, float* sums, int n, int in_h, int in_w, int out_h, int out_w,
                                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(in && out, "bilinear_adjoint: null pointer argument");
  RPNET_REQUIRE(n > 0 && out_h > 0 && out_w > 0 && in_h >= out_h && in_w >= out_w, "bilinear_adjoint: bad shape");
  RPNET_REQUIRE(!sums || (in_h % out_h == 0 && in_w % out_w == 0), "bilinear_adjoint: sums need an integer scale");
  if (sums) RPNET_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)n * sizeof(float), stream));
  const dim3 grid((out_h * out_w + 127) / 128, n);
  if (in_h == 4 * out_h && in_w == 4 * out_w)
    bilinear_adjoint_int_kernel<4><<<grid, 128, 0, stream>>>(in, out, sums, in_h, in_w, out_h, out_w);
  else if (in_h == 8 * out_h && in_w == 8 * out_w)
    bilinear_adjoint_int_kernel<8><<<grid, 128, 0, stream>>>(in, out, sums, in_h, in_w, out_h, out_w);
  else
    bilinear_adjoint_kernel<<<grid, 128, 0, stream>>>(in, out, sums, in_h, in_w, out_h, out_w);
  return check_cuda(cudaGetLastError(), "bilinear_adjoint launch");
}

RPNET_API int rpnet_weighted_pool_f32(const float* feat, const float* wmap0, const float* wmap1, const float* msum0, const float* msum1,
                                       float* out, int n, int hw, int c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(feat && wmap0 && wmap1 && msum0 && msum1 && out, "weighted_pool: null pointer argument");
  RPNET_REQUIRE(n > 0 && hw > 0 && c > 0 && c <= 64 && c % 4 == 0, "weighted_pool: bad shape n=%d hw=%d c=%d (c <= 64, c %% 4 == 0)", n, hw, c);
  weighted_pool_kernel<<<n, 1024, 0, stream>>>(feat, wmap0, wmap1, msum0, msum1, out, hw, c);
  return check_cuda(cudaGetLastError(), "weighted_pool launch");
}

RPNET_API int rpnet_weighted_pool_bwd_f32(const float* dout, const float* wmap0, const float* wmap1, const float* msum0,
                                           const float* msum1, float* dfeat, int accumulate, int n, int hw, int c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(dout && wmap0 && wmap1 && msum0 && msum1 && dfeat, "weighted_pool_bwd: null pointer argument");
  RPNET_REQUIRE(n > 0 && hw > 0 && c > 0 && c % 4 == 0, "weighted_pool_bwd: bad shape");
  weighted_pool_bwd_kernel<<<grid_for((long long)n * hw * (c / 4), 256), 256, 0, stream>>>(dout, wmap0, wmap1, msum0, msum1, dfeat,
                                                                                          accumulate, n, hw, c);
  return check_cuda(cudaGetLastError(), "weighted_pool_bwd launch");
}

RPNET_API int rpnet_proto_finalize_bwd_f32(const float* dprotos, float* draw, int ways, int shots, int batch, int c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(dprotos && draw, "proto_finalize_bwd: null pointer argument");
  RPNET_REQUIRE(ways > 0 && shots > 0 && batch > 0 && c > 0, "proto_finalize_bwd: bad shape");
  proto_finalize_bwd_kernel<<<grid_for((long long)ways * shots * batch * 2 * c, 256), 256, 0, stream>>>(dprotos, draw, ways, shots, batch, c);
  return check_cuda(cudaGetLastError(), "proto_finalize_bwd launch");
}

RPNET_API int rpnet_dice_ce_f32(const float* logits, const long long* labels, int groups, int batch, int n_classes, long long hw,
                                 float eps, float grad_scale, double* sums, float* dlogits, float* loss, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(logits && labels && sums && loss, "dice_ce: null pointer argument");
  RPNET_REQUIRE(groups > 0 && batch > 0 && hw > 0 && n_classes >= 2 && n_classes <= kMaxP, "dice_ce: bad shape (classes 2..%d)", kMaxP);
  RPNET_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)groups * (2 * n_classes + 1) * sizeof(double), stream));
  int bx = grid_for((long long)batch * hw, 256 * 4, 8);
  bx = (bx + groups - 1) / groups;
  if (bx < 1) bx = 1;
  dice_ce_reduce_kernel<<<dim3(bx, groups), 256, 0, stream>>>(logits, labels, batch, n_classes, hw, sums);
  RPNET_CUDA_OK(cudaGetLastError());
  if (dlogits) {
    dice_ce_grad_kernel<<<dim3(bx, groups), 256, 0, stream>>>(logits, labels, sums, batch, n_classes, hw, eps, grad_scale, dlogits, loss);
  } else {
    dice_ce_grad_kernel<<<dim3(1, groups), 32, 0, stream>>>(logits, labels, sums, batch, n_classes, hw, eps, grad_scale, nullptr, loss);
  }
  return check_cuda(cudaGetLastError(), "dice_ce launch");
}

RPNET_API int rpnet_class_pool_f32(const float* feat, const float* pred, int batch, int hw, int c, int n_classes, float* qproto,
                                    float* counts, void* amax_i32, void* stream_) {
  int* amax = static_cast<int*>(amax_i32);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(feat && pred && qproto && counts && amax, "class_pool: null pointer argument");
  RPNET_REQUIRE(c == 64 && n_classes >= 1 && n_classes <= kMaxP && batch > 0 && hw > 0, "class_pool: bad shape (c == 64)");
  class_pool_kernel<<<dim3(kCpCluster, batch), 256, 0, stream>>>(feat, pred, hw, n_classes, qproto, counts, amax);
  return check_cuda(cudaGetLastError(), "class_pool launch");
}

RPNET_API int rpnet_class_pool_bwd_f32(const float* dqproto, const float* counts, const void* amax_i32, int batch, int hw, int c,
                                        int n_classes, float* dfeat, void* stream_) {
  const int* amax = static_cast<const int*>(amax_i32);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(dqproto && counts && amax && dfeat, "class_pool_bwd: null pointer argument");
  RPNET_REQUIRE(c == 64 && batch > 0 && hw > 0 && n_classes >= 1, "class_pool_bwd: bad shape (c == 64)");
  class_pool_bwd_kernel<<<grid_for((long long)batch * hw * 16, 256), 256, 0, stream>>>(dqproto, counts, amax, batch, hw, n_classes, dfeat);
  return check_cuda(cudaGetLastError(), "class_pool_bwd launch");
}

RPNET_API int rpnet_align_gather_f32(const float* qproto, const float* counts, int ways, int shots, int batch, float scaler,
                                      float* protos_s, float* weight, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(qproto && counts && protos_s && weight, "align_gather: null pointer argument");
  RPNET_REQUIRE(ways > 0 && shots > 0 && batch > 0, "align_gather: bad shape");
  align_gather_kernel<<<grid_for((long long)ways * shots * batch * 128, 256), 256, 0, stream>>>(qproto, counts, ways, shots, batch, scaler,
                                                                                              protos_s, weight);
  return check_cuda(cudaGetLastError(), "align_gather launch");
}

RPNET_API int rpnet_align_scatter_f32(const float* dprotos_s, int ways, int shots, int batch, float* dqproto, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(dprotos_s && dqproto, "align_scatter: null pointer argument");
  RPNET_REQUIRE(ways > 0 && shots > 0 && batch > 0, "align_scatter: bad shape");
  align_scatter_kernel<<<grid_for((long long)batch * (1 + ways) * 64, 256), 256, 0, stream>>>(dprotos_s, ways, shots, batch, dqproto);
  return check_cuda(cudaGetLastError(), "align_scatter launch");
}

RPNET_API int rpnet_ce_mask_f32(const float* logits, const float* fore, const float* back, const float* weight, int n, long long hw,
                                 float grad_scale, double* sums, float* dlogits, float* loss, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(logits && fore && back && weight && sums && loss, "ce_mask: null pointer argument");
  RPNET_REQUIRE(n > 0 && hw > 0, "ce_mask: bad shape");
  RPNET_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)n * 2 * sizeof(double), stream));
  int bx = grid_for(hw, 256 * 4, 2);
  ce_mask_reduce_kernel<<<dim3(bx, n), 256, 0, stream>>>(logits, fore, back, hw, sums);
  RPNET_CUDA_OK(cudaGetLastError());
  if (dlogits) ce_mask_grad_kernel<<<dim3(bx, n), 256, 0, stream>>>(logits, fore, back, sums, weight, n, hw, grad_scale, dlogits, loss);
  else         ce_mask_grad_kernel<<<dim3(1, 1), 32, 0, stream>>>(logits, fore, back, sums, weight, n, 0, grad_scale, nullptr, loss);
  return check_cuda(cudaGetLastError(), "ce_mask launch");
}

RPNET_API int rpnet_bilinear_up_f32(const float* in, float* out, int n, int h, int w, int out_h, int out_w, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(in && out, "bilinear_up: null pointer argument");
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0 && out_h >= h && out_w >= w, "bilinear_up: bad shape");
  bilinear_up_kernel<<<grid_for((long long)n * out_h * out_w, 256), 256, 0, stream>>>(in, out, n, h, w, out_h, out_w);
  return check_cuda(cudaGetLastError(), "bilinear_up launch");
}

// =====================================================================================================================
// "Next" row N1: the per-slice affine registration that produces the warped support image / label in front of
// the hot path — AffineRegistration (net/registration.py:316-357) driven by get_registration_field
// (dataset/few_shot_reader.py:109-198): theta = identity; 50 x { warped = grid_sample(moving, affine_grid(theta));
// loss = mean((warped - fixed)^2); loss.backward(); Adam(lr 0.01).step() }, one slice after the other, ~10 launches per
// iteration.  Here: ONE launch for all slices, one CTA per slice runs every iteration — the bilinear warp, the loss, the
// analytic gradient of the six parameters (the backward of grid_sample o affine_grid, align_corners = False, zero
// padding), an ordered block reduction and the Adam update stay inside the kernel.
// =====================================================================================================================
namespace rpnet {

constexpr int kRegThreads = 1024;

struct Bilin {
  float v, dvdx, dvdy;      // sample and its derivative w.r.t. the (unnormalised) source coordinates
};

// F.grid_sample(bilinear, padding_mode='zeros', align_corners=False) at source pixel coordinates (ix, iy)
__device__ __forceinline__ Bilin bilinear_zero(const float* __restrict__ img, int H, int W, float ix, float iy) {
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float tx = ix - fx, ty = iy - fy;
  const bool xin0 = x0 >= 0 && x0 < W, xin1 = x1 >= 0 && x1 < W, yin0 = y0 >= 0 && y0 < H, yin1 = y1 >= 0 && y1 < H;
  const float v00 = (xin0 && yin0) ? __ldg(img + y0 * W + x0) : 0.f;
  const float v01 = (xin1 && yin0) ? __ldg(img + y0 * W + x1) : 0.f;
  const float v10 = (xin0 && yin1) ? __ldg(img + y1 * W + x0) : 0.f;
  const float v11 = (xin1 && yin1) ? __ldg(img + y1 * W + x1) : 0.f;
  Bilin r;
  r.v = (v00 * (1.f - tx) + v01 * tx) * (1.f - ty) + (v10 * (1.f - tx) + v11 * tx) * ty;
  r.dvdx = (v01 - v00) * (1.f - ty) + (v11 - v10) * ty;
  r.dvdy = (v10 - v00) * (1.f - tx) + (v11 - v01) * tx;
  return r;
}

__global__ void __launch_bounds__(kRegThreads)
affine_register_kernel(const float* __restrict__ moving, const float* __restrict__ fixed, int H, int W, int iters, float lr,
                       float beta1, float beta2, float eps, float* __restrict__ theta_out /*[n][6]*/,
                       float* __restrict__ loss_out /*[n][iters] or null*/) {
  __shared__ float s_theta[6], s_m[6], s_v[6];
  __shared__ float s_part[kRegThreads / 32][7];
  const int n = blockIdx.x;
  const float* mov = moving + (size_t)n * H * W;
  const float* fix = fixed + (size_t)n * H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 6) {
    s_theta[threadIdx.x] = (threadIdx.x == 0 || threadIdx.x == 4) ? 1.f : 0.f;       // identity (net/registration.py:320-322)
    s_m[threadIdx.x] = 0.f;
    s_v[threadIdx.x] = 0.f;
  }
  __syncthreads();
  const int total = H * W;
  const float inv_n = 1.f / (float)total;
  for (int it = 1; it <= iters; ++it) {
    const float t0 = s_theta[0], t1 = s_theta[1], t2 = s_theta[2], t3 = s_theta[3], t4 = s_theta[4], t5 = s_theta[5];
    float acc[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[k] = 0.f;
    for (int p = threadIdx.x; p < total; p += kRegThreads) {
      const int i = p / W, j = p - i * W;
      const float xn = (2.f * j + 1.f) / (float)W - 1.f, yn = (2.f * i + 1.f) / (float)H - 1.f;   // affine_grid, align_corners=False
      const float gx = t0 * xn + t1 * yn + t2, gy = t3 * xn + t4 * yn + t5;
      const float ix = ((gx + 1.f) * (float)W - 1.f) * 0.5f, iy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
      const Bilin s = bilinear_zero(mov, H, W, ix, iy);
      const float r = s.v - __ldg(fix + p);
      const float g = 2.f * r * inv_n;
      const float gxg = g * s.dvdx * (0.5f * (float)W), gyg = g * s.dvdy * (0.5f * (float)H);
      acc[0] = fmaf(gxg, xn, acc[0]); acc[1] = fmaf(gxg, yn, acc[1]); acc[2] += gxg;
      acc[3] = fmaf(gyg, xn, acc[3]); acc[4] = fmaf(gyg, yn, acc[4]); acc[5] += gyg;
      acc[6] = fmaf(r, r, acc[6]);
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 7; ++k) s_part[warp][k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 7) {
      float s = 0.f;
      for (int wv = 0; wv < kRegThreads / 32; ++wv) s += s_part[wv][threadIdx.x];        // fixed order: deterministic
      if (threadIdx.x == 6) {
        if (loss_out) loss_out[(size_t)n * iters + it - 1] = s * inv_n;
      } else {
        // torch.optim.Adam (no weight decay, no amsgrad)
        const float m = beta1 * s_m[threadIdx.x] + (1.f - beta1) * s;
        const float v = beta2 * s_v[threadIdx.x] + (1.f - beta2) * s * s;
        s_m[threadIdx.x] = m;
        s_v[threadIdx.x] = v;
        const float bc1 = 1.f - powf(beta1, (float)it), bc2 = 1.f - powf(beta2, (float)it);
        const float denom = sqrtf(v) / sqrtf(bc2) + eps;
        s_theta[threadIdx.x] -= (lr / bc1) * (m / denom);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < 6) theta_out[(size_t)n * 6 + threadIdx.x] = s_theta[threadIdx.x];
}

// out[n][c][i][j] = grid_sample(x[n][c], affine_grid(theta[n]))  (bilinear, zeros, align_corners=False): AffineRegistration.forward
__global__ void affine_warp_kernel(const float* __restrict__ x, const float* __restrict__ theta, float* __restrict__ out, int N, int C,
                                   int H, int W) {
  const long long total = (long long)N * C * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % W), i = (int)((idx / W) % H);
    const long long nc = idx / ((long long)W * H);
    const int n = (int)(nc / C);
    const float* th = theta + (size_t)n * 6;
    const float xn = (2.f * j + 1.f) / (float)W - 1.f, yn = (2.f * i + 1.f) / (float)H - 1.f;
    const float gx = th[0] * xn + th[1] * yn + th[2], gy = th[3] * xn + th[4] * yn + th[5];
    const float ix = ((gx + 1.f) * (float)W - 1.f) * 0.5f, iy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
    out[idx] = bilinear_zero(x + nc * H * W, H, W, ix, iy).v;
  }
}

}  // namespace rpnet

RPNET_API int rpnet_affine_register_f32(const float* moving, const float* fixed, int n, int h, int w, int iters, float lr, float beta1,
                                         float beta2, float eps, float* theta, float* loss_curve, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(moving && fixed && theta, "affine_register: null pointer argument");
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0 && iters >= 0 && (long long)h * w < (1LL << 30), "affine_register: bad shape");
  rpnet::affine_register_kernel<<<n, rpnet::kRegThreads, 0, stream>>>(moving, fixed, h, w, iters, lr, beta1, beta2, eps, theta, loss_curve);
  return check_cuda(cudaGetLastError(), "affine_register launch");
}

RPNET_API int rpnet_affine_warp_f32(const float* x, const float* theta, float* out, int n, int c, int h, int w, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(x && theta && out, "affine_warp: null pointer argument");
  RPNET_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "affine_warp: bad shape");
  rpnet::affine_warp_kernel<<<grid_for((long long)n * c * h * w, 256), 256, 0, stream>>>(x, theta, out, n, c, h, w);
  return check_cuda(cudaGetLastError(), "affine_warp launch");
}

// ---------------------------------------------------------------------------------------------------
// NCC (net/registration.py:157-160), the similarity the eval driver prints next to Dice (test_rpnet.py:229-230):
//   -sum((f - mean f) * (m - mean m)) / sqrt(sum((f - mean f)^2) * sum((m - mean m)^2) + 1e-10)
// one pass: the five raw moments in fp64 (per-thread fp32 partials over a few elements, fp64 across threads / blocks).
// ---------------------------------------------------------------------------------------------------
namespace rpnet {
__global__ void __launch_bounds__(256)
ncc_moments_kernel(const float* __restrict__ m, const float* __restrict__ f, long long n, double* __restrict__ sums /*[5]*/) {
  __shared__ double s_red[8][5];
  double acc[5] = {0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double a = (double)__ldg(m + i), b = (double)__ldg(f + i);
    acc[0] += a; acc[1] += b; acc[2] += a * a; acc[3] += b * b; acc[4] += a * b;
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 5; ++k) s_red[threadIdx.x >> 5][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0;
    for (int wv = 0; wv < 8; ++wv) s += s_red[wv][threadIdx.x];
    atomicAdd(sums + threadIdx.x, s);
  }
}
__global__ void ncc_finalize_kernel(const double* __restrict__ sums, long long n, float* __restrict__ out) {
  const double N = (double)n, sm = sums[0], sf = sums[1];
  const double cov = sums[4] - sm * sf / N, vm = sums[2] - sm * sm / N, vf = sums[3] - sf * sf / N;
  *out = (float)(-cov / sqrt(vf * vm + 1e-10));
}
}  // namespace rpnet

RPNET_API int rpnet_ncc_f32(const float* moving, const float* fixed, long long n, double* scratch5, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(moving && fixed && scratch5 && out && n > 0, "ncc: bad argument");
  RPNET_CUDA_OK(cudaMemsetAsync(scratch5, 0, 5 * sizeof(double), stream));
  rpnet::ncc_moments_kernel<<<grid_for(n, 256 * 8, 4), 256, 0, stream>>>(moving, fixed, n, scratch5);
  RPNET_CUDA_OK(cudaGetLastError());
  rpnet::ncc_finalize_kernel<<<1, 1, 0, stream>>>(scratch5, n, out);
  return check_cuda(cudaGetLastError(), "ncc launch");
}
