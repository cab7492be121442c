// Train-mode streaming kernels of the RP-Net conv stacks (HBM-bound, CUDA cores): batch-statistics
// BatchNorm2d (+ReLU, + fused 2x2 max-pool) forward and backward, weight packing, nearest x2 upsample,
// pre-mask backward, the Cin=1 first-conv weight gradient and the fused Adam step.
// The reference trains through torch autograd over nn.BatchNorm2d / nn.ReLU / nn.MaxPool2d / nn.Upsample
// (net/modules.py:42-75, net/unet.py:435-467); these kernels restate those ops' forward/backward on
// NHWC fp16 activations and NHWC bf16 activation gradients.
//
// BatchNorm call groups: images [start[g], start[g+1]) of a batched launch form one BatchNorm *call* of the
// reference (own batch statistics, own running-stat update — SURVEY D14).
#include "common.cuh"

#include <cstdlib>

extern "C" {
// include/rpnet_b200.h
typedef struct {
  const float* w; void* w_fwd_f16; void* w_dgrad_bf16;
  int cout, cin_real, ntaps, hole_start, hole_len, split;
} rpnet_pack_desc;
}

namespace rpnet {

constexpr int kMaxGroups = 64;
struct Groups {
  int G;
  int start[kMaxGroups + 1];
  int bstart[kMaxGroups + 1];     // reduction kernels: blocks [bstart[g], bstart[g+1]) of the 1-D grid work on call group g
};
// group / block-in-group / blocks-of-group of this block in a reduction launch planned by plan_group_blocks()
__device__ __forceinline__ void group_block(const Groups& gr, int& g, int& bx, int& nbx) {
  g = 0;
  while (g + 1 < gr.G && (int)blockIdx.x >= gr.bstart[g + 1]) ++g;
  bx = (int)blockIdx.x - gr.bstart[g];
  nbx = gr.bstart[g + 1] - gr.bstart[g];
}
__device__ __forceinline__ int group_of(const Groups& gr, int n) {
  int g = 0;
  while (g + 1 < gr.G && n >= gr.start[g + 1]) ++g;
  return g;
}

static int make_groups(Groups* gr, const int* group_start, int groups, int n) {
  RPNET_REQUIRE(groups >= 1 && groups <= kMaxGroups && group_start, "bn: groups %d out of range [1, %d]", groups, kMaxGroups);
  gr->G = groups;
  for (int g = 0; g <= groups; ++g) gr->start[g] = group_start[g];
  RPNET_REQUIRE(gr->start[0] == 0 && gr->start[groups] == n, "bn: group_start must span [0, %d]", n);
  for (int g = 0; g < groups; ++g) RPNET_REQUIRE(gr->start[g + 1] > gr->start[g], "bn: empty BatchNorm call group %d", g);
  return 0;
}

// Blocks per call group proportional to the group's size (a support pass of 80 images next to a query pass of 16 must not
// get the same number of blocks), at most `budget` in total; returns the grid size.
static int plan_group_blocks(Groups* gr, long long units_per_img, long long units_per_block, int budget = 0) {
  if (budget <= 0) {
    static int env_budget = -1;
    if (env_budget < 0) {
      const char* e = getenv("RPNET_BN_BLOCKS");          // tuning knob (blocks of a reduction launch); default 3 per SM
      env_budget = e ? atoi(e) : 148 * 2;
      if (env_budget <= 0) env_budget = 148 * 2;
    }
    budget = env_budget;
  }
  long long total_units = 0, want_sum = 0;
  for (int g = 0; g < gr->G; ++g) {
    const long long u = (long long)(gr->start[g + 1] - gr->start[g]) * units_per_img;
    total_units += u;
    want_sum += (u + units_per_block - 1) / units_per_block;
  }
  int acc = 0;
  for (int g = 0; g < gr->G; ++g) {
    const long long u = (long long)(gr->start[g + 1] - gr->start[g]) * units_per_img;
    long long b = (u + units_per_block - 1) / units_per_block;
    if (want_sum > budget) b = (long long)budget * u / (total_units > 0 ? total_units : 1);
    if (b < 1) b = 1;
    gr->bstart[g] = acc;
    acc += (int)b;
  }
  gr->bstart[gr->G] = acc;
  return acc;
}

// blocks of an element-wise (apply) launch: a few resident waves of 2 blocks per SM; RPNET_BN_APPLY_BLOCKS overrides
static int apply_blocks() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RPNET_BN_APPLY_BLOCKS");
    v = e ? atoi(e) : 148 * 4;
    if (v <= 0) v = 148 * 4;
  }
  return v;
}

static int grid_for(long long total, int block, int cap_mult = 16) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * cap_mult;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// ---------------------------------------------------------------------------------------------------
// bn_stats: per (call group, channel) sum and sum of squares of the raw conv output z (fp16 NHWC).
// Block = (C/8) channel vectors x (256 / (C/8)) pixel lanes; fp32 per-thread partials, shared-memory
// tree over the pixel lanes, one atomicAdd per channel per block.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_stats_kernel(const uint4* __restrict__ z, const uint4* __restrict__ z_lo, Groups gr, int HW, int c8,
                double* __restrict__ sums /*[G][C][2]*/) {
  __shared__ float s_red[256 * 16];
  int g, bx, nbx;
  group_block(gr, g, bx, nbx);
  const int lanes = 256 / c8;
  const int v = threadIdx.x % c8, pl = threadIdx.x / c8;
  const long long p0 = (long long)gr.start[g] * HW, p1 = (long long)gr.start[g + 1] * HW;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  constexpr int U = 4;                          // pixels per thread and iteration: all loads go out before the math
  const long long stride = (long long)nbx * lanes;
  for (long long pb = p0 + (long long)bx * lanes + pl; pb < p1; pb += U * stride) {
    uint4 zh[U], zl[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long p = pb + k * stride;
      if (p < p1) {
        zh[k] = __ldg(z + p * c8 + v);
        if (z_lo) zl[k] = __ldg(z_lo + p * c8 + v);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (pb + k * stride >= p1) break;
      float f[8];
      unpack8_f16(zh[k], f);
      if (z_lo) {                                 // split-fp16 z = hi + lo
        float l[8];
        unpack8_f16(zl[k], l);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += l[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1[j] += f[j]; s2[j] = fmaf(f[j], f[j], s2[j]); }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { s_red[threadIdx.x * 16 + j] = s1[j]; s_red[threadIdx.x * 16 + 8 + j] = s2[j]; }
  __syncthreads();
  for (int s = lanes >> 1; s > 0; s >>= 1) {
    if (pl < s) {
#pragma unroll
      for (int j = 0; j < 16; ++j) s_red[threadIdx.x * 16 + j] += s_red[(threadIdx.x + s * c8) * 16 + j];
    }
    __syncthreads();
  }
  if (pl == 0) {
    double* dst = sums + ((size_t)g * c8 * 8 + v * 8) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(dst + 2 * j, (double)s_red[threadIdx.x * 16 + j]);
      atomicAdd(dst + 2 * j + 1, (double)s_red[threadIdx.x * 16 + 8 + j]);
    }
  }
}

// bn_finalize: mean / rstd / folded (a, b) per (group, channel) + the running-statistics update of
// nn.BatchNorm2d (momentum, unbiased running_var), applied once per call group IN ORDER.  `conv_bias` is the
// bias the conv kernel dropped (it cancels inside train-mode BN but belongs to the running mean).
__global__ void bn_finalize_kernel(const double* __restrict__ sums, Groups gr, int C, int HW, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ conv_bias, float eps, float momentum,
                                   float* running_mean, float* running_var, long long* num_batches_tracked,
                                   float* __restrict__ stats /*[G][C][4]*/) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += gr.G;
  if (c >= C) return;
  float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 0.f;
  const float ga = gamma[c], be = beta[c], cb = conv_bias ? conv_bias[c] : 0.f;
  for (int g = 0; g < gr.G; ++g) {
    const float cnt = (float)(gr.start[g + 1] - gr.start[g]) * (float)HW;
    // E[z^2] - E[z]^2 in fp64: channels whose mean dwarfs their standard deviation lose all fp32 digits in the subtraction
    const double dmean = sums[((size_t)g * C + c) * 2] / (double)cnt;
    const double dvar = sums[((size_t)g * C + c) * 2 + 1] / (double)cnt - dmean * dmean;
    const float mean = (float)dmean;
    float var = (float)dvar;
    var = var > 0.f ? var : 0.f;
    const float rstd = rsqrtf(var + eps);
    float* st = stats + ((size_t)g * C + c) * 4;
    st[0] = mean; st[1] = rstd; st[2] = rstd * ga; st[3] = be - mean * rstd * ga;
    rm = (1.f - momentum) * rm + momentum * (mean + cb);
    rv = (1.f - momentum) * rv + momentum * (cnt > 1.f ? var * cnt / (cnt - 1.f) : var);
  }
  if (running_mean) running_mean[c] = rm;
  if (running_var) running_var[c] = rv;
}

// bn_apply: y = act(a * z + b) -> fp16 NHWC (optional), fp32 NHWC (optional), 2x2 max-pooled fp16 (optional).
// Thread (pl, v): pixel lane pl, fixed 8-channel vector v (its (a, b) stay in registers; reloaded when the image moves to
// another call group — pixels are visited in increasing order, so the group only moves forward).  U units per iteration
// with all loads issued before the math; 32-bit index arithmetic (n*h*w < 2^31).
template <bool POOL>
__global__ void __launch_bounds__(256, 2)
bn_apply_kernel(const uint4* __restrict__ z, const uint4* __restrict__ z_lo, const float* __restrict__ stats, Groups gr, int N, int H,
                int W, int c8, int relu, uint4* __restrict__ y16, uint4* __restrict__ y16_lo, float* __restrict__ y32,
                uint4* __restrict__ ypool, uint4* __restrict__ ypool_lo, int lo_fmt, const uint4* __restrict__ res,
                const void* __restrict__ res_lo) {
  constexpr int U = POOL ? 2 : 4;
  const int C = c8 * 8;
  const int lanes = 256 / c8;
  const int v = threadIdx.x % c8, pl = threadIdx.x / c8;
  const unsigned Hs = POOL ? H / 2 : H, Ws = POOL ? W / 2 : W;
  const unsigned units = (unsigned)N * Hs * Ws;
  const unsigned stride = gridDim.x * lanes;
  int g = 0;
  unsigned g_end = (unsigned)gr.start[1] * Hs * Ws;          // first unit of the next call group
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 st = __ldg(reinterpret_cast<const float4*>(stats) + (size_t)v * 8 + j);
    a[j] = st.z; b[j] = st.w;
  }
  for (unsigned u0 = blockIdx.x * lanes + pl; u0 < units; u0 += U * stride) {
    uint4 zin[U][POOL ? 4 : 1], zlo[U][POOL ? 4 : 1];
    unsigned pix[U][POOL ? 4 : 1];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const unsigned u = u0 + k * stride;
      if (u < units) {
        if (POOL) {
          const unsigned xs = u % Ws, t = u / Ws;
          const unsigned ys = t % Hs, n = t / Hs;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            pix[k][q] = (n * H + ys * 2 + (q >> 1)) * W + xs * 2 + (q & 1);
            zin[k][q] = __ldg(z + (size_t)pix[k][q] * c8 + v);
            if (z_lo) zlo[k][q] = __ldg(z_lo + (size_t)pix[k][q] * c8 + v);
          }
        } else {
          pix[k][0] = u;
          zin[k][0] = __ldg(z + (size_t)u * c8 + v);
          if (z_lo) zlo[k][0] = __ldg(z_lo + (size_t)u * c8 + v);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const unsigned u = u0 + k * stride;
      if (u >= units) break;
      if (u >= g_end) {
        while (u >= g_end) { ++g; g_end = (unsigned)gr.start[g + 1] * Hs * Ws; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 st = __ldg(reinterpret_cast<const float4*>(stats) + (size_t)g * C + v * 8 + j);
          a[j] = st.z; b[j] = st.w;
        }
      }
      float best[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) best[j] = -INFINITY;
#pragma unroll
      for (int q = 0; q < (POOL ? 4 : 1); ++q) {
        float f[8];
        unpack8_f16(zin[k][q], f);
        if (z_lo) {                               // split-fp16 z = hi + lo
          float l[8];
          unpack8_f16(zlo[k][q], l);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += l[j];
        }
        float idn[8];                             // BasicBlock: y = relu(bn2(z) + identity); identity as hi (+ lo) planes
#pragma unroll
        for (int j = 0; j < 8; ++j) idn[j] = 0.f;
        if (!POOL && res) {
          unpack8_f16(__ldg(res + (size_t)pix[k][q] * c8 + v), idn);
          if (res_lo) lo8_add(res_lo, lo_fmt, ((size_t)pix[k][q] * c8 + v) * 8, (v * 8) & 63, idn);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float r = fmaf(f[j], a[j], b[j]) + idn[j];
          r = relu ? fmaxf(r, 0.f) : r;
          f[j] = r;
          best[j] = fmaxf(best[j], r);
        }
        if (y16) {
          const uint4 hi = pack8_f16(f);
          y16[(size_t)pix[k][q] * c8 + v] = hi;
          if (y16_lo) lo8_store(y16_lo, lo_fmt, ((size_t)pix[k][q] * c8 + v) * 8, (v * 8) & 63, f, hi);
        }
        if (y32) {
          float4* d = reinterpret_cast<float4*>(y32 + ((size_t)pix[k][q] * c8 + v) * 8);
          d[0] = make_float4(f[0], f[1], f[2], f[3]);
          d[1] = make_float4(f[4], f[5], f[6], f[7]);
        }
      }
      if (POOL) {
        const uint4 hi = pack8_f16(best);
        ypool[(size_t)u * c8 + v] = hi;
        if (ypool_lo) lo8_store(ypool_lo, lo_fmt, ((size_t)u * c8 + v) * 8, (v * 8) & 63, best, hi);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// BatchNorm + ReLU (+ max-pool / nearest-upsample consumers) backward.
// The gradient w.r.t. the activation y = relu(a*z+b) arrives from up to three places:
//   direct : bf16 (or fp32) NHWC [n,h,w] with pixel pitch d_ld, channel offset d_off  (next conv's dgrad / concat slice)
//   pooled : bf16 NHWC [n,h/2,w/2]: gradient of the 2x2 max-pooled copy, routed to the FIRST max of the window
//            (nn.MaxPool2d backward; y is recomputed from z so the argmax is the forward's)
//   up     : bf16 NHWC [n,2h,2w]: gradient of the nearest x2 upsampled copy, summed over its 2x2 block.
// ---------------------------------------------------------------------------------------------------
struct GradSrc {
  const void* direct; int d_ld, d_off, d_f32;
  const __nv_bfloat16* pooled; int p_ld, p_off;
  const __nv_bfloat16* up; int u_ld, u_off;
};

__device__ __forceinline__ void load_grad8(const GradSrc& s, int n, int y, int x, int H, int W, int c, float* g) {
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = 0.f;
  const long long pix = ((long long)n * H + y) * W + x;
  if (s.direct) {
    if (s.d_f32) {
      const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(s.direct) + pix * s.d_ld + s.d_off + c);
      const float4 a = __ldg(p), b = __ldg(p + 1);
      g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
    } else {
      unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(s.direct) + pix * s.d_ld + s.d_off + c)), g);
    }
  }
  if (s.up) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long up = ((long long)n * (2 * H) + 2 * y + (q >> 1)) * (2 * W) + 2 * x + (q & 1);
      float t[8];
      unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(s.up + up * s.u_ld + s.u_off + c)), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += t[j];
    }
  }
}

// Work unit: one 2x2 pixel window (WIN) or one pixel (!WIN) x 8 channels.
// MODE 0 (reduce): per (group, channel) S1 = sum dy_hat, S2 = sum dy_hat * (z - mean)   [x_hat = (z - mean) * rstd]
// MODE 1 (apply):  dz = a * (dy_hat - m1 - x_hat * m2) = a * dy_hat + k1 * z + k0 with per-channel constants
//                  k1 = -a * m2 * rstd, k0 = -a * m1 - k1 * mean   (m1 = S1 / cnt, m2 = rstd * S2 / cnt)
// where dy_hat = relu'(a*z+b) * (sum of the incoming gradients).  The 8-channel vector of a thread is fixed, so the
// per-channel constants live in registers and are reloaded only when the image moves to another call group.
template <bool WIN, int MODE>
__global__ void __launch_bounds__(256, WIN ? 2 : 3)
bn_bwd_kernel(const uint4* __restrict__ z, const float* __restrict__ stats, const float* __restrict__ coef, Groups gr, int N, int H,
              int W, int c8, int relu, GradSrc src, double* __restrict__ sums, uint4* __restrict__ dz) {
  __shared__ float s_red[MODE == 0 ? 256 * 16 : 1];
  const int C = c8 * 8;
  const int Hs = WIN ? H / 2 : H, Ws = WIN ? W / 2 : W;
  const int lanes = 256 / c8;
  const int v = threadIdx.x % c8, pl = threadIdx.x / c8;
  int n_begin = 0, n_end = N;
  int bg = 0, bx = blockIdx.x, nbx = gridDim.x;
  if (MODE == 0) { group_block(gr, bg, bx, nbx); n_begin = gr.start[bg]; n_end = gr.start[bg + 1]; }   // a block stays inside one group
  const unsigned units = (unsigned)(n_end - n_begin) * Hs * Ws;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  int cur_g = -1;
  float a[8], b[8], c1[8], c0[8];          // MODE 0: c1 = mean (c0 unused); MODE 1: c1 = k1, c0 = k0
  for (unsigned u = (unsigned)bx * lanes + pl; u < units; u += (unsigned)nbx * lanes) {
    unsigned t = u;
    const int xs = (int)(t % (unsigned)Ws);  t /= (unsigned)Ws;
    const int ys = (int)(t % (unsigned)Hs);
    const int n = n_begin + (int)(t / (unsigned)Hs);
    const int g = MODE == 0 ? bg : group_of(gr, n);
    if (g != cur_g) {
      cur_g = g;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 st = __ldg(reinterpret_cast<const float4*>(stats) + (size_t)g * C + v * 8 + j);
        a[j] = st.z; b[j] = st.w;
        if (MODE == 0) {
          c1[j] = st.x; c0[j] = 0.f;
        } else {
          const float2 cf = __ldg(reinterpret_cast<const float2*>(coef) + (size_t)g * C + v * 8 + j);   // (m1, m2)
          c1[j] = -st.z * cf.y * st.y;
          c0[j] = -st.z * cf.x - c1[j] * st.x;
        }
      }
    }
    uint4 zin[WIN ? 4 : 1];
    uint32_t win_arg = 0;                  // 2 bits per channel: which pixel of the 2x2 window holds the (first) maximum
    float pg[8];
    if (WIN) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        zin[q] = __ldg(z + (((long long)n * H + ys * 2 + (q >> 1)) * W + xs * 2 + (q & 1)) * c8 + v);
#pragma unroll
      for (int j = 0; j < 8; ++j) pg[j] = 0.f;
      if (src.pooled) {
        const long long pp = ((long long)n * Hs + ys) * Ws + xs;
        unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(src.pooled + pp * src.p_ld + src.p_off + v * 8)), pg);
      }
      float best[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) best[j] = -INFINITY;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float f[8];
        unpack8_f16(zin[q], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float r = fmaf(f[j], a[j], b[j]);
          r = relu ? fmaxf(r, 0.f) : r;
          if (r > best[j]) { best[j] = r; win_arg = (win_arg & ~(3u << (2 * j))) | ((uint32_t)q << (2 * j)); }
        }
      }
    } else {
      zin[0] = __ldg(z + (((long long)n * H + ys) * W + xs) * c8 + v);
    }
#pragma unroll
    for (int q = 0; q < (WIN ? 4 : 1); ++q) {
      const int y = WIN ? ys * 2 + (q >> 1) : ys, x = WIN ? xs * 2 + (q & 1) : xs;
      float gy[8], f[8];
      load_grad8(src, n, y, x, H, W, v * 8, gy);
      unpack8_f16(zin[q], f);
      float out[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float d = gy[j];
        if (WIN && ((win_arg >> (2 * j)) & 3u) == (uint32_t)q) d += pg[j];
        if (relu && !(fmaf(f[j], a[j], b[j]) > 0.f)) d = 0.f;
        if (MODE == 0) {
          s1[j] += d;
          s2[j] = fmaf(d, f[j] - c1[j], s2[j]);
        } else {
          out[j] = fmaf(a[j], d, fmaf(c1[j], f[j], c0[j]));
        }
      }
      if (MODE == 1) dz[(((long long)n * H + y) * W + x) * c8 + v] = pack8_bf16(out);
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { s_red[threadIdx.x * 16 + j] = s1[j]; s_red[threadIdx.x * 16 + 8 + j] = s2[j]; }
    __syncthreads();
    for (int s = lanes >> 1; s > 0; s >>= 1) {
      if (pl < s) {
#pragma unroll
        for (int j = 0; j < 16; ++j) s_red[threadIdx.x * 16 + j] += s_red[(threadIdx.x + s * c8) * 16 + j];
      }
      __syncthreads();
    }
    if (pl == 0) {
      // fp32 block partials (fixed summation order inside the block) accumulated in fp64: the addition of fp32 addends into
      // a double is exact — hence independent of the order the blocks arrive in — unless an addend is below 2^-30 of the
      // running sum; two identical runs give bit-identical gradients (tests/test_gpu_train.py::test_backward_is_deterministic)
      double* dst = sums + ((size_t)bg * C + v * 8) * 2;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(dst + 2 * j, (double)s_red[threadIdx.x * 16 + j]);
        atomicAdd(dst + 2 * j + 1, (double)s_red[threadIdx.x * 16 + 8 + j]);
      }
    }
  }
}

// Direct-only specialisation (the gradient arrives as one bf16 NHWC tensor / channel slice: every conv that feeds another
// conv): U pixels per thread and iteration with all 2U 16-byte loads issued before the math (memory-level parallelism),
// 32-bit index arithmetic, call group tracked by pixel boundaries (pixels are visited in increasing order).
template <int MODE>
__global__ void __launch_bounds__(256, 2)
bn_bwd_direct_kernel(const uint4* __restrict__ z, const float* __restrict__ stats, const float* __restrict__ coef, Groups gr, int N,
                     int HW, int c8, int relu, const __nv_bfloat16* __restrict__ gdir, int d_ld, int d_off,
                     double* __restrict__ sums, uint4* __restrict__ dz) {
  constexpr int U = 4;
  __shared__ float s_red[MODE == 0 ? 256 * 16 : 1];
  const int C = c8 * 8;
  const int lanes = 256 / c8;
  const int v = threadIdx.x % c8, pl = threadIdx.x / c8;
  int bg = 0, bx = blockIdx.x, nbx = gridDim.x;
  unsigned p_begin = 0, p_end = (unsigned)N * HW;
  if (MODE == 0) { group_block(gr, bg, bx, nbx); p_begin = (unsigned)gr.start[bg] * HW; p_end = (unsigned)gr.start[bg + 1] * HW; }
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  int g = MODE == 0 ? bg : -1;
  unsigned g_end = MODE == 0 ? p_end : 0u;
  float a[8], b[8], c1[8], c0[8];          // MODE 0: c1 = mean; MODE 1: c1 = k1, c0 = k0
  if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 st = __ldg(reinterpret_cast<const float4*>(stats) + (size_t)bg * C + v * 8 + j);
      a[j] = st.z; b[j] = st.w; c1[j] = st.x; c0[j] = 0.f;
    }
  }
  const unsigned stride = (unsigned)nbx * lanes;
  for (unsigned p0 = p_begin + (unsigned)bx * lanes + pl; p0 < p_end; p0 += U * stride) {
    uint4 zr[U], gq[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const unsigned p = p0 + k * stride;
      if (p < p_end) {
        zr[k] = __ldg(z + (size_t)p * c8 + v);
        gq[k] = __ldg(reinterpret_cast<const uint4*>(gdir + (size_t)p * d_ld + d_off + v * 8));
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const unsigned p = p0 + k * stride;
      if (p >= p_end) break;
      if (MODE == 1 && p >= g_end) {
        do { ++g; g_end = (unsigned)gr.start[g + 1] * HW; } while (p >= g_end);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 st = __ldg(reinterpret_cast<const float4*>(stats) + (size_t)g * C + v * 8 + j);
          const float2 cf = __ldg(reinterpret_cast<const float2*>(coef) + (size_t)g * C + v * 8 + j);   // (m1, m2)
          a[j] = st.z; b[j] = st.w;
          c1[j] = -st.z * cf.y * st.y;
          c0[j] = -st.z * cf.x - c1[j] * st.x;
        }
      }
      float f[8], d[8], out[8];
      unpack8_f16(zr[k], f);
      unpack8_bf16(gq[k], d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dd = (relu && !(fmaf(f[j], a[j], b[j]) > 0.f)) ? 0.f : d[j];
        if (MODE == 0) {
          s1[j] += dd;
          s2[j] = fmaf(dd, f[j] - c1[j], s2[j]);
        } else {
          out[j] = fmaf(a[j], dd, fmaf(c1[j], f[j], c0[j]));
        }
      }
      if (MODE == 1) dz[(size_t)p * c8 + v] = pack8_bf16(out);
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { s_red[threadIdx.x * 16 + j] = s1[j]; s_red[threadIdx.x * 16 + 8 + j] = s2[j]; }
    __syncthreads();
    for (int s = lanes >> 1; s > 0; s >>= 1) {
      if (pl < s) {
#pragma unroll
        for (int j = 0; j < 16; ++j) s_red[threadIdx.x * 16 + j] += s_red[(threadIdx.x + s * c8) * 16 + j];
      }
      __syncthreads();
    }
    if (pl == 0) {
      // fp32 block partials (fixed summation order inside the block) accumulated in fp64: the addition of fp32 addends into
      // a double is exact — hence independent of the order the blocks arrive in — unless an addend is below 2^-30 of the
      // running sum; two identical runs give bit-identical gradients (tests/test_gpu_train.py::test_backward_is_deterministic)
      double* dst = sums + ((size_t)bg * C + v * 8) * 2;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(dst + 2 * j, (double)s_red[threadIdx.x * 16 + j]);
        atomicAdd(dst + 2 * j + 1, (double)s_red[threadIdx.x * 16 + 8 + j]);
      }
    }
  }
}

// Pool-only specialisation (encoder.Conv1/Conv2 second convs: the activation feeds nothing but the 2x2 max-pool): the
// incoming gradient is non-zero only at the window's (first) maximum, so the reduce pass touches one value per window and
// the apply pass is one FMA per element plus the routed term.  Same MODE semantics as bn_bwd_kernel.
template <int MODE>
__global__ void __launch_bounds__(256, MODE == 0 ? 3 : 2)
bn_bwd_pool_kernel(const uint4* __restrict__ z, const float* __restrict__ stats, const float* __restrict__ coef, Groups gr, int N,
                   int H, int W, int c8, int relu, const __nv_bfloat16* __restrict__ pooled, int p_ld, int p_off,
                   double* __restrict__ sums, uint4* __restrict__ dz) {
  __shared__ float s_red[MODE == 0 ? 256 * 16 : 1];
  const int C = c8 * 8;
  const int Hs = H / 2, Ws = W / 2;
  const int lanes = 256 / c8;
  const int v = threadIdx.x % c8, pl = threadIdx.x / c8;
  int n_begin = 0, n_end = N;
  int bg = 0, bx = blockIdx.x, nbx = gridDim.x;
  if (MODE == 0) { group_block(gr, bg, bx, nbx); n_begin = gr.start[bg]; n_end = gr.start[bg + 1]; }
  const unsigned units = (unsigned)(n_end - n_begin) * Hs * Ws;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  int cur_g = -1;
  float a[8], b[8], c1[8], c0[8];
  for (unsigned u = (unsigned)bx * lanes + pl; u < units; u += (unsigned)nbx * lanes) {
    unsigned t = u;
    const int xs = (int)(t % (unsigned)Ws);  t /= (unsigned)Ws;
    const int ys = (int)(t % (unsigned)Hs);
    const int n = n_begin + (int)(t / (unsigned)Hs);
    const int g = MODE == 0 ? bg : group_of(gr, n);
    if (g != cur_g) {
      cur_g = g;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 st = __ldg(reinterpret_cast<const float4*>(stats) + (size_t)g * C + v * 8 + j);
        a[j] = st.z; b[j] = st.w;
        if (MODE == 0) {
          c1[j] = st.x; c0[j] = 0.f;
        } else {
          const float2 cf = __ldg(reinterpret_cast<const float2*>(coef) + (size_t)g * C + v * 8 + j);
          c1[j] = -st.z * cf.y * st.y;
          c0[j] = -st.z * cf.x - c1[j] * st.x;
        }
      }
    }
    uint4 zin[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      zin[q] = __ldg(z + (((long long)n * H + ys * 2 + (q >> 1)) * W + xs * 2 + (q & 1)) * c8 + v);
    float pg[8];
    unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(pooled + (((long long)n * Hs + ys) * Ws + xs) * p_ld + p_off + v * 8)), pg);
    float best[8], zbest[8];
    int win[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; zbest[j] = 0.f; win[j] = 0; }
    float f[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      unpack8_f16(zin[q], f[q]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float r = fmaf(f[q][j], a[j], b[j]);
        r = relu ? fmaxf(r, 0.f) : r;
        if (r > best[j]) { best[j] = r; zbest[j] = f[q][j]; win[j] = q; }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = (relu && !(best[j] > 0.f)) ? 0.f : pg[j];       // relu'(winner) * routed gradient
      if (MODE == 0) {
        s1[j] += d;
        s2[j] = fmaf(d, zbest[j] - c1[j], s2[j]);
      } else {
        pg[j] = a[j] * d;
      }
    }
    if (MODE == 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float out[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = fmaf(c1[j], f[q][j], c0[j]) + (win[j] == q ? pg[j] : 0.f);
        dz[(((long long)n * H + ys * 2 + (q >> 1)) * W + xs * 2 + (q & 1)) * c8 + v] = pack8_bf16(out);
      }
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { s_red[threadIdx.x * 16 + j] = s1[j]; s_red[threadIdx.x * 16 + 8 + j] = s2[j]; }
    __syncthreads();
    for (int s = lanes >> 1; s > 0; s >>= 1) {
      if (pl < s) {
#pragma unroll
        for (int j = 0; j < 16; ++j) s_red[threadIdx.x * 16 + j] += s_red[(threadIdx.x + s * c8) * 16 + j];
      }
      __syncthreads();
    }
    if (pl == 0) {
      // fp32 block partials (fixed summation order inside the block) accumulated in fp64: the addition of fp32 addends into
      // a double is exact — hence independent of the order the blocks arrive in — unless an addend is below 2^-30 of the
      // running sum; two identical runs give bit-identical gradients (tests/test_gpu_train.py::test_backward_is_deterministic)
      double* dst = sums + ((size_t)bg * C + v * 8) * 2;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(dst + 2 * j, (double)s_red[threadIdx.x * 16 + j]);
        atomicAdd(dst + 2 * j + 1, (double)s_red[threadIdx.x * 16 + 8 + j]);
      }
    }
  }
}

// coef[g][c] = (m1, m2) = (S1 / cnt, rstd * S2 / cnt);  dgamma[c] += sum_g rstd * S2, dbeta[c] += sum_g S1
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ stats, Groups gr, int C, int HW,
                                       float* dgamma, float* dbeta, float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float dg = 0.f, db = 0.f;
  for (int g = 0; g < gr.G; ++g) {
    const float cnt = (float)(gr.start[g + 1] - gr.start[g]) * (float)HW;
    const float rstd = stats[((size_t)g * C + c) * 4 + 1];
    const float S1 = (float)sums[((size_t)g * C + c) * 2], S2 = rstd * (float)sums[((size_t)g * C + c) * 2 + 1];
    coef[((size_t)g * C + c) * 2] = S1 / cnt;
    coef[((size_t)g * C + c) * 2 + 1] = S2 / cnt;
    db += S1;
    dg += S2;
  }
  if (dgamma) dgamma[c] += dg;
  if (dbeta) dbeta[c] += db;
}

// ---------------------------------------------------------------------------------------------------
// nn.Upsample(scale_factor=2) (nearest), fp16 NHWC.  net/modules.py:67 (train path: the up-sampled map is
// materialised so that the weight gradient is a plain 3x3 tap-list GEMM).
// ---------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int c8) {
  const long long total = (long long)N * 2 * H * 2 * W * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % c8);
    long long t = i / c8;
    const int xo = (int)(t % (2 * W));  t /= 2 * W;
    const int yo = (int)(t % (2 * H));
    const int n = (int)(t / (2 * H));
    y[i] = __ldg(x + (((long long)n * H + (yo >> 1)) * W + (xo >> 1)) * c8 + v);
  }
}

// fp16 -> bf16 copy: tcgen05 kind::f16 wants one element format for both operands, and the weight-gradient GEMM's
// other operand is the bf16 activation gradient.
__global__ void cvt_f16_bf16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    unpack8_f16(__ldg(in + i), f);
    out[i] = pack8_bf16(f);
  }
}

// d(x) for x_fg = x*m, x_bg = x*(1-m) (net/rp_net.py:275,283), summed over `iters` uses of the same x:
//   dx[p][c] = sum_i dxfg[i][p][c] * m[i][p] + dxbg[i][p][c] * (1 - m[i][p])
__global__ void premask_bwd_kernel(const uint4* __restrict__ dxfg, const uint4* __restrict__ dxbg, const float* __restrict__ m,
                                   int iters, long long pixels, int c8, uint4* __restrict__ dx) {
  const long long total = pixels * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / c8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int it = 0; it < iters; ++it) {
      const float mk = __ldg(m + (long long)it * pixels + pix);
      float a[8], b[8];
      unpack8_bf16(__ldg(dxfg + (long long)it * total + i), a);
      unpack8_bf16(__ldg(dxbg + (long long)it * total + i), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a[j] * mk + b[j] * (1.f - mk);
    }
    dx[i] = pack8_bf16(acc);
  }
}

// ---------------------------------------------------------------------------------------------------
// soft_mask: True (net/rp_net.py:308-311 without the threshold): the recurrent mask m_{i+1} = avg_pool2d(p_fg(logits_i), S)
// stays in the graph, so iteration i+1 sends a gradient back into iteration i's logits.
// (1) d m(p) = sum_c (dxfg[p][c] - dxbg[p][c]) * x[p][c]           (x_fg = x * m, x_bg = x * (1 - m), :283)
//     one warp per pixel: lanes stride the 8-channel vectors, shuffle reduce.
// (2) dlogits[b][k][Y][X] += dm[b][Y/S][X/S] / S^2 * s_k * ([k >= 1] - p_fg),  s = softmax(logits[b][:][Y][X]),
//     p_fg = sum_{k>=1} s_k  (the reference's channel 1 for Wa == 1; the oracle-ext sum for Wa > 1).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
premask_mask_bwd_kernel(const uint4* __restrict__ dxfg, const uint4* __restrict__ dxbg, const uint4* __restrict__ x, long long pixels,
                        int c8, float* __restrict__ dm) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = warp; p < pixels; p += nwarps) {
    float acc = 0.f;
    for (int v = lane; v < c8; v += 32) {
      float a[8], b[8], f[8];
      unpack8_bf16(__ldg(dxfg + p * c8 + v), a);
      unpack8_bf16(__ldg(dxbg + p * c8 + v), b);
      unpack8_f16(__ldg(x + p * c8 + v), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(a[j] - b[j], f[j], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) dm[p] = acc;
  }
}

constexpr int kSoftMaxP = 8;
__global__ void soft_mask_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ dm, int B, int P, int H, int W, int S,
                                     float* __restrict__ dlogits) {
  const long long total = (long long)B * H * W;
  const int h = H / S, w = W / S;
  const float inv = 1.f / (float)(S * S);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % W), Y = (int)((i / W) % H), b = (int)(i / ((long long)W * H));
    const float g = __ldg(dm + ((long long)b * h + Y / S) * w + X / S) * inv;
    float v[kSoftMaxP], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kSoftMaxP; ++k)
      if (k < P) { v[k] = __ldg(logits + (((long long)b * P + k) * H + Y) * W + X); mx = fmaxf(mx, v[k]); }
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < kSoftMaxP; ++k)
      if (k < P) { v[k] = expf(v[k] - mx); den += v[k]; }
    const float rden = 1.f / den;
    const float pfg = 1.f - v[0] * rden;
#pragma unroll
    for (int k = 0; k < kSoftMaxP; ++k)
      if (k < P) {
        float* d = dlogits + (((long long)b * P + k) * H + Y) * W + X;
        *d += g * v[k] * rden * ((k >= 1 ? 1.f : 0.f) - pfg);
      }
  }
}

// ---------------------------------------------------------------------------------------------------
// Weight repack (every optimizer step): fp32 [cout][cin_real][taps] ->
//   forward  fp16 [taps][cout][cin]   (K-major B operand of conv_igemm)
//   dgrad    bf16 [taps][cin][cout]   (transposed; used with the negated tap list)
// `cin` = cin_real + hole_len: packed input channels [hole_start, hole_start+hole_len) are zero padding.
// ---------------------------------------------------------------------------------------------------
// One block per 32 (cout) x 32 (cin) tile: the fp32 weights of the tile (32 x 32 x taps, contiguous runs of 32 * taps floats
// per cout) are staged in shared memory, so that both packs leave as coalesced rows — wf rows run along cin, the transposed
// wd rows along cout.
// one weight of a split forward pack row [Wh (cin) | second half]: split 1 -> Wl fp16 (cin); split 2 -> the fp8 corrections, per 64
// input channels 128 bytes = Wh8 (64) | Wl8 (64) (common.cuh, "c8")
__device__ __forceinline__ void store_split_weight(__half* row, int cin, int ci, float val, int split) {
  const __half hi = __float2half_rn(val);
  const float lo = val - __half2float(hi);
  row[ci] = hi;
  if (split == 2) {
    uint8_t* g = reinterpret_cast<uint8_t*>(row + cin) + (ci >> 6) * 128 + (ci & 63);
    g[0] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(hi) * kC8WhScale, __NV_SATFINITE, __NV_E4M3);
    g[64] = (uint8_t)__nv_cvt_float_to_fp8(lo * kC8WlScale, __NV_SATFINITE, __NV_E4M3);
  } else {
    row[cin + ci] = __float2half_rn(lo);
  }
}

__device__ __forceinline__ void pack_conv_weight_tile(float (*s_w)[32 * 9 + 1], int tile, const float* __restrict__ w, int cout, int cin,
                                                      int ntaps, int hole_start, int hole_len, __half* __restrict__ wf,
                                                      __nv_bfloat16* __restrict__ wd, int split) {
  const int cin_real = cin - hole_len;
  const int tiles_ci = (cin + 31) / 32;
  const int co0 = (tile / tiles_ci) * 32, ci0 = (tile % tiles_ci) * 32;
  // stage: thread -> (co, element of the run); packed channel ci maps to the real channel (padding hole = zeros)
  for (int i = threadIdx.x; i < 32 * 32 * ntaps; i += 256) {
    const int co = i / (32 * ntaps), r = i % (32 * ntaps);
    const int ci = ci0 + r / ntaps, t = r % ntaps;
    float val = 0.f;
    if (co0 + co < cout && ci < cin && (ci < hole_start || ci >= hole_start + hole_len)) {
      const int cr = ci < hole_start ? ci : ci - hole_len;
      val = __ldg(w + ((size_t)(co0 + co) * cin_real + cr) * ntaps + t);
    }
    s_w[co][r] = val;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, rowi = threadIdx.x >> 5;     // 8 rows per pass
  for (int t = 0; t < ntaps; ++t) {
    for (int r = rowi; r < 32; r += 8) {
      if (wf && co0 + r < cout && ci0 + lane < cin) {              // wf[t][co][ci]: lanes along ci
        const float val = s_w[r][lane * ntaps + t];
        if (split)                                                 // split pack [t][co][Wh (cin) | Wl (cin) or fp8 corrections]
          store_split_weight(wf + ((size_t)t * cout + co0 + r) * (2 * cin), cin, ci0 + lane, val, split);
        else
          wf[((size_t)t * cout + co0 + r) * cin + ci0 + lane] = __float2half_rn(val);
      }
      if (wd && ci0 + r < cin && co0 + lane < cout)                // wd[t][ci][co]: lanes along co
        wd[((size_t)t * cin + ci0 + r) * cout + co0 + lane] = __float2bfloat16_rn(s_w[lane][r * ntaps + t]);
    }
  }
}

__global__ void __launch_bounds__(256)
pack_conv_weight_kernel(const float* __restrict__ w, int cout, int cin, int ntaps, int hole_start, int hole_len,
                        __half* __restrict__ wf, __nv_bfloat16* __restrict__ wd, int split) {
  __shared__ float s_w[32][32 * 9 + 1];                           // [co][ci * ntaps + t]
  pack_conv_weight_tile(s_w, blockIdx.x, w, cout, cin, ntaps, hole_start, hole_len, wf, wd, split);
}

// Every conv of the model in ONE launch (the packs run after each optimizer step: twenty small launches were launch bound):
// block b works on tile b - first_tile[l] of layer l.
constexpr int kMaxPackLayers = 40;
struct PackLayer { const float* w; __half* wf; __nv_bfloat16* wd; int cout, cin, ntaps, hole_start, hole_len, split, first_tile; };
struct PackTable { int n; PackLayer l[kMaxPackLayers]; };
__global__ void __launch_bounds__(256)
pack_conv_weights_kernel(const __grid_constant__ PackTable t) {
  __shared__ float s_w[32][32 * 9 + 1];
  int l = 0;
  while (l + 1 < t.n && (int)blockIdx.x >= t.l[l + 1].first_tile) ++l;
  const PackLayer& L = t.l[l];
  pack_conv_weight_tile(s_w, (int)blockIdx.x - L.first_tile, L.w, L.cout, L.cin, L.ntaps, L.hole_start, L.hole_len, L.wf, L.wd, L.split);
}

// ---------------------------------------------------------------------------------------------------
// Sub-pixel packs of an up_conv weight (nn.Upsample(x2) + 3x3 conv, net/modules.py:61-75), every optimizer step:
//   wf  fp16 [4 phases][4 taps][cout][cin] : forward, phase (py, px) = py * 2 + px, tap = ty * 2 + tx, row/column-summed
//   w16 bf16 [16][cin][cout]               : data gradient (4x4 stride-2 form), tap = (oy + 1) * 4 + (ox + 1)
// ---------------------------------------------------------------------------------------------------
__global__ void pack_upconv_weight_kernel(const float* __restrict__ w, int cout, int cin, __half* __restrict__ wf,
                                          __nv_bfloat16* __restrict__ w16, int split) {
  const long long cc = (long long)cout * cin;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < cc; e += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(e % cin), co = (int)(e / cin);
    float k[3][3];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i / 3][i % 3] = __ldg(w + e * 9 + i);
    // forward phases: rows(py=0) = {ky 0 | ky 1+2}, rows(py=1) = {ky 0+1 | ky 2}
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px)
#pragma unroll
        for (int ty = 0; ty < 2; ++ty)
#pragma unroll
          for (int tx = 0; tx < 2; ++tx) {
            float s = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const int ry = py == 0 ? (ky == 0 ? 0 : 1) : (ky == 2 ? 1 : 0), rx = px == 0 ? (kx == 0 ? 0 : 1) : (kx == 2 ? 1 : 0);
                if (ry == ty && rx == tx) s += k[ky][kx];
              }
            if (split)                                             // [phase][tap][co][Wh (cin) | Wl (cin) or fp8 corrections]
              store_split_weight(wf + (((size_t)(py * 2 + px) * 4 + ty * 2 + tx) * cout + co) * (2 * cin), cin, ci, s, split);
            else
              wf[(((size_t)(py * 2 + px) * 4 + ty * 2 + tx) * cout + co) * cin + ci] = __float2half_rn(s);
          }
    // data gradient: S(-1) = {2}, S(0) = {1, 2}, S(1) = {0, 1}, S(2) = {0}
#pragma unroll
    for (int oy = -1; oy <= 2; ++oy)
#pragma unroll
      for (int ox = -1; ox <= 2; ++ox) {
        float s = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const bool iny = (oy == -1 && ky == 2) || (oy == 0 && ky >= 1) || (oy == 1 && ky <= 1) || (oy == 2 && ky == 0);
            const bool inx = (ox == -1 && kx == 2) || (ox == 0 && kx >= 1) || (ox == 1 && kx <= 1) || (ox == 2 && kx == 0);
            if (iny && inx) s += k[ky][kx];
          }
        w16[((size_t)((oy + 1) * 4 + ox + 1) * cin + ci) * cout + co] = __float2bfloat16_rn(s);
      }
  }
}

// ---------------------------------------------------------------------------------------------------
// Weight gradient of the Cin = 1 first conv (encoder.Conv1.conv.0): grad[co][ky][kx] += sum_p dz[p][co] * img[p + tap].
// 8 threads per pixel, 8 output channels each (one 16-byte dz load), 72 register accumulators per thread; the four
// pixel lanes of a warp are folded with shuffles, warps through shared-memory atomics, blocks through global atomics.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv3x3_first_wgrad_kernel(const float* __restrict__ img /* plane ci of image 0 */, unsigned img_n_stride /* cin * H * W */,
                           const uint4* __restrict__ dz, int N, int H, int W, double* __restrict__ acc64 /*[64][9], zeroed*/) {
  __shared__ double s_acc[64 * 9];      // fp32 warp partials added in fp64: exact, so the arrival order of warps / blocks is immaterial
  for (int i = threadIdx.x; i < 64 * 9; i += 256) s_acc[i] = 0.0;
  __syncthreads();
  const int cg = threadIdx.x & 7;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  const unsigned total = (unsigned)N * H * W;                    // pixels (< 2^31, host-checked)
  const unsigned stride = (gridDim.x * blockDim.x) >> 3;
  constexpr int U = 2;                                           // pixels per iteration: all 2 x (1 + 9) loads go out before the math
  for (unsigned p0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; p0 < total; p0 += U * stride) {
    uint4 dq[U];
    float v[U][9];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned p = p0 + u * stride;
      const bool ok = p < total;
      const unsigned x = p % (unsigned)W, y = (p / (unsigned)W) % (unsigned)H;
      const unsigned base = (p / ((unsigned)H * (unsigned)W)) * img_n_stride;   // this image's plane
      dq[u] = ok ? __ldg(dz + (size_t)p * 8 + cg) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = (int)y + ky - 1;
        const bool yok = ok && yy >= 0 && yy < H;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = (int)x + kx - 1;
          v[u][ky * 3 + kx] = (yok && xx >= 0 && xx < W) ? __ldg(img + base + (unsigned)(yy * W + xx)) : 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float d[8];
      unpack8_bf16(dq[u], d);
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(d[j], v[u][t], acc[t][j]);
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = acc[t][j];
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      if ((threadIdx.x & 31) < 8) atomicAdd(&s_acc[(cg * 8 + j) * 9 + t], (double)a);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 9; i += 256) atomicAdd(acc64 + i, (double)(float)s_acc[i]);
}

__global__ void add_f64_to_f32_kernel(const double* __restrict__ acc, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += (float)acc[i];
}
// grad[co][ci][t] += acc[co][t] for one input plane ci of a Cin-channel first conv
__global__ void add_first_wgrad_kernel(const double* __restrict__ acc, float* __restrict__ grad, int cin, int ci) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 64 * 9) grad[((i / 9) * cin + ci) * 9 + i % 9] += (float)acc[i];
}

// ---------------------------------------------------------------------------------------------------
// Backward pieces of the VGG stack (net/vgg.py:22-58: conv + bias + ReLU, nn.MaxPool2d(3, stride, 1), no normalisation).
// (1) g = dy * [y > 0] (bf16) and dbias[c] += sum_p g[p][c]  (y = null: the last conv has no ReLU).  Thread = fixed 8-channel
//     vector, fp32 partials, shared-memory tree over the pixel lanes, fp64 accumulation across blocks (exact, order independent).
// (2) nn.MaxPool2d backward from the argmax positions recorded by the forward (uint8 window position of the first maximum):
//     every input pixel gathers from the (at most ceil(k / stride)^2) windows that cover it.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
relu_bias_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, long long pixels, int c8, uint4* __restrict__ g,
                     double* __restrict__ dbias_acc) {
  __shared__ float s_red[256 * 8];
  const int lanes = 256 / c8;
  const int v = threadIdx.x % c8, pl = threadIdx.x / c8;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  for (long long p = (long long)blockIdx.x * lanes + pl; p < pixels; p += (long long)gridDim.x * lanes) {
    float d[8];
    unpack8_bf16(__ldg(dy + p * c8 + v), d);
    if (y) {
      float f[8];
      unpack8_f16(__ldg(y + p * c8 + v), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = f[j] > 0.f ? d[j] : 0.f;
    }
    g[p * c8 + v] = pack8_bf16(d);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += d[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s_red[threadIdx.x * 8 + j] = s[j];
  __syncthreads();
  for (int st = lanes >> 1; st > 0; st >>= 1) {
    if (pl < st) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s_red[threadIdx.x * 8 + j] += s_red[(threadIdx.x + st * c8) * 8 + j];
    }
    __syncthreads();
  }
  if (pl == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dbias_acc + v * 8 + j, (double)s_red[threadIdx.x * 8 + j]);
  }
}

// BasicBlock backward glue: out = (a + b) masked by y > 0 (b and y optional): the ReLU of `y = relu(bn2(z2) + identity)` in front
// of both branches, and the sum of the main-branch and identity-branch input gradients.
__global__ void add_relu_mask_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const uint4* __restrict__ y,
                                     uint4* __restrict__ out, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float d[8];
    unpack8_bf16(__ldg(a + i), d);
    if (b) {
      float e[8];
      unpack8_bf16(__ldg(b + i), e);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] += e[j];
    }
    if (y) {
      float f[8];
      unpack8_f16(__ldg(y + i), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = f[j] > 0.f ? d[j] : 0.f;
    }
    out[i] = pack8_bf16(d);
  }
}

// ---------------------------------------------------------------------------------------------------
// Weight gradient of the ResNet stem conv (7x7, stride 2, padding 3, 3 -> 64; net/rp_net.py:19-23 = torchvision resnet18.conv1):
//   grad[co][ci][ky][kx] = sum over (n, oy, ox) of dz[n][oy][ox][co] * img[n][ci][2 oy + ky - 3][2 ox + kx - 3]
// One block per 8 x 8 tile of output pixels: the 21 x 21 x 3 input patch and the 64 x 64 dz tile are staged in shared memory;
// thread t owns output channel t % 64 and the taps {t / 64 + 4 i} (37 of the 147 per thread) in registers across the tiles of its
// block, block partials (fp32, fixed order) are accumulated in fp64 (deterministic, like every other reduction of the path).
// ---------------------------------------------------------------------------------------------------
constexpr int kStemTaps = 147, kStemTapsPerThread = 37, kStemTile = 8, kStemPatch = 2 * kStemTile + 5;
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const float* __restrict__ img, const __nv_bfloat16* __restrict__ dz, int N, int H, int W, int Ho, int Wo,
                  double* __restrict__ acc64) {
  __shared__ float s_in[3][kStemPatch][kStemPatch + 1];
  __shared__ float s_dz[kStemTile * kStemTile][64];
  const int co = threadIdx.x & 63, kg = threadIdx.x >> 6;
  const int tiles_x = (Wo + kStemTile - 1) / kStemTile, tiles_y = (Ho + kStemTile - 1) / kStemTile;
  const long long tiles = (long long)N * tiles_y * tiles_x;
  float acc[kStemTapsPerThread];
  int t_ci[kStemTapsPerThread], t_off[kStemTapsPerThread];
#pragma unroll
  for (int i = 0; i < kStemTapsPerThread; ++i) {
    acc[i] = 0.f;
    const int t = kg + 4 * i;                       // tap index ci * 49 + ky * 7 + kx (>= 147: idle slot)
    const int tt = t < kStemTaps ? t : 0;
    t_ci[i] = tt / 49;
    t_off[i] = ((tt % 49) / 7) * (kStemPatch + 1) + (tt % 7);
  }
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tx = (int)(tile % tiles_x), ty = (int)((tile / tiles_x) % tiles_y), n = (int)(tile / ((long long)tiles_x * tiles_y));
    const int oy0 = ty * kStemTile, ox0 = tx * kStemTile;
    const int iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * kStemPatch * kStemPatch; i += 256) {
      const int ci = i / (kStemPatch * kStemPatch), r = i % (kStemPatch * kStemPatch);
      const int y = iy0 + r / kStemPatch, x = ix0 + r % kStemPatch;
      s_in[ci][r / kStemPatch][r % kStemPatch] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(img + ((size_t)(n * 3 + ci) * H + y) * W + x) : 0.f;
    }
    for (int i = threadIdx.x; i < kStemTile * kStemTile * 64; i += 256) {
      const int p = i >> 6, c = i & 63;
      const int oy = oy0 + p / kStemTile, ox = ox0 + p % kStemTile;
      s_dz[p][c] = (oy < Ho && ox < Wo) ? __bfloat162float(dz[((size_t)(n * Ho + oy) * Wo + ox) * 64 + c]) : 0.f;
    }
    __syncthreads();
    for (int p = 0; p < kStemTile * kStemTile; ++p) {
      const float g = s_dz[p][co];
      const float* base = &s_in[0][2 * (p / kStemTile)][2 * (p % kStemTile)];
#pragma unroll
      for (int i = 0; i < kStemTapsPerThread; ++i)
        acc[i] = fmaf(g, base[t_ci[i] * kStemPatch * (kStemPatch + 1) + t_off[i]], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < kStemTapsPerThread; ++i) {
    const int t = kg + 4 * i;
    if (t < kStemTaps) atomicAdd(acc64 + (size_t)co * kStemTaps + t, (double)acc[i]);
  }
}

__global__ void maxpool_bwd_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx, uint4* __restrict__ dx, int N, int H, int W,
                                   int c8, int Ho, int Wo, int k, int stride, int pad) {
  const long long total = (long long)N * H * W * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % c8);
    long long pix = i / c8;
    const int x = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / ((long long)W * H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // windows (yo, xo) with yo * stride - pad <= y <= yo * stride - pad + k - 1
    int yo0 = y + pad - k + 1;  yo0 = yo0 <= 0 ? 0 : (yo0 + stride - 1) / stride;
    int xo0 = x + pad - k + 1;  xo0 = xo0 <= 0 ? 0 : (xo0 + stride - 1) / stride;
    const int yo1 = min((y + pad) / stride, Ho - 1), xo1 = min((x + pad) / stride, Wo - 1);
    for (int yo = yo0; yo <= yo1; ++yo)
      for (int xo = xo0; xo <= xo1; ++xo) {
        const int pos = (y - (yo * stride - pad)) * k + (x - (xo * stride - pad));
        const long long o = ((long long)(n * Ho + yo) * Wo + xo) * c8 + cv;
        const uint2 id = __ldg(idx + o);
        float d[8];
        unpack8_bf16(__ldg(dy + o), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const unsigned b = ((j < 4 ? id.x : id.y) >> (8 * (j & 3))) & 0xffu;
          if ((int)b == pos) acc[j] += d[j];
        }
      }
    dx[i] = pack8_bf16(acc);
  }
}

// ---------------------------------------------------------------------------------------------------
// torch.optim.Adam (weight_decay = L2 added to the gradient, yamls/example.yml:64-67) on flat fp32 buffers.
// ---------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                            float grad_scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i] * grad_scale);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

}  // namespace rpnet

using namespace rpnet;

RPNET_API int rpnet_bn_stats_split_f16(const void* z, const void* z_lo, int n, int h, int w, int c, const int* group_start, int groups,
                                        double* sums, void* stream_);
RPNET_API int rpnet_bn_apply_split_f16(const void* z, const void* z_lo, const float* stats, int n, int h, int w, int c,
                                        const int* group_start, int groups, int relu, void* y_f16, void* y_lo_f16, void* y_pool_f16,
                                        void* y_pool_lo_f16, float* y_f32, int lo_fmt, void* stream_);

RPNET_API int rpnet_bn_stats_f16(const void* z, int n, int h, int w, int c, const int* group_start, int groups, double* sums,
                                  void* stream_) {
  return rpnet_bn_stats_split_f16(z, nullptr, n, h, w, c, group_start, groups, sums, stream_);
}

RPNET_API int rpnet_bn_stats_split_f16(const void* z, const void* z_lo, int n, int h, int w, int c, const int* group_start, int groups,
                                        double* sums, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(z && sums, "bn_stats: null pointer argument");
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0 && c >= 64 && c % 8 == 0 && 256 % (c / 8) == 0, "bn_stats: bad shape n=%d h=%d w=%d c=%d (c in {64..2048}, power of two)", n, h, w, c);
  Groups gr;
  int rc = make_groups(&gr, group_start, groups, n);
  if (rc) return rc;
  RPNET_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)groups * c * 2 * sizeof(double), stream));
  const int lanes = 256 / (c / 8);
  const int grid = plan_group_blocks(&gr, (long long)h * w, (long long)lanes * 8);
  bn_stats_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint4*>(z), static_cast<const uint4*>(z_lo), gr, h * w, c / 8, sums);
  return check_cuda(cudaGetLastError(), "bn_stats launch");
}

RPNET_API int rpnet_bn_finalize_f32(const double* sums, const int* group_start, int groups, int c, int hw, const float* gamma,
                                     const float* beta, const float* conv_bias, float eps, float momentum, float* running_mean,
                                     float* running_var, long long* num_batches_tracked, float* stats, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(sums && gamma && beta && stats, "bn_finalize: null pointer argument");
  RPNET_REQUIRE(c > 0 && hw > 0, "bn_finalize: bad shape");
  Groups gr;
  int rc = make_groups(&gr, group_start, groups, group_start ? group_start[groups > 0 ? groups : 0] : 0);
  if (rc) return rc;
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, stream>>>(sums, gr, c, hw, gamma, beta, conv_bias, eps, momentum, running_mean,
                                                          running_var, num_batches_tracked, stats);
  return check_cuda(cudaGetLastError(), "bn_finalize launch");
}

RPNET_API int rpnet_bn_apply_f16(const void* z, const float* stats, int n, int h, int w, int c, const int* group_start, int groups,
                                  int relu, void* y_f16, void* y_pool_f16, float* y_f32, void* stream_) {
  return rpnet_bn_apply_split_f16(z, nullptr, stats, n, h, w, c, group_start, groups, relu, y_f16, nullptr, y_pool_f16, nullptr, y_f32,
                                  0, stream_);
}

RPNET_API int rpnet_bn_apply_res_f16(const void* z, const void* z_lo, const float* stats, int n, int h, int w, int c,
                                      const int* group_start, int groups, int relu, const void* res_f16, const void* res_lo, void* y_f16,
                                      void* y_lo_f16, void* y_pool_f16, void* y_pool_lo_f16, float* y_f32, int lo_fmt, void* stream_);

RPNET_API int rpnet_bn_apply_split_f16(const void* z, const void* z_lo, const float* stats, int n, int h, int w, int c,
                                        const int* group_start, int groups, int relu, void* y_f16, void* y_lo_f16, void* y_pool_f16,
                                        void* y_pool_lo_f16, float* y_f32, int lo_fmt, void* stream_) {
  return rpnet_bn_apply_res_f16(z, z_lo, stats, n, h, w, c, group_start, groups, relu, nullptr, nullptr, y_f16, y_lo_f16, y_pool_f16,
                                y_pool_lo_f16, y_f32, lo_fmt, stream_);
}

RPNET_API int rpnet_bn_apply_res_f16(const void* z, const void* z_lo, const float* stats, int n, int h, int w, int c,
                                      const int* group_start, int groups, int relu, const void* res_f16, const void* res_lo, void* y_f16,
                                      void* y_lo_f16, void* y_pool_f16, void* y_pool_lo_f16, float* y_f32, int lo_fmt, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE((!res_lo || res_f16) && (!res_f16 || !y_pool_f16), "bn_apply: the residual input goes with the unpooled output only");
  RPNET_REQUIRE((!y_lo_f16 || y_f16) && (!y_pool_lo_f16 || y_pool_f16), "bn_apply: a residual output needs its main output");
  RPNET_REQUIRE(lo_fmt == 0 || (lo_fmt == 1 && c % 64 == 0), "bn_apply: lo_fmt %d (c8 planes need c %% 64 == 0, c = %d)", lo_fmt, c);
  RPNET_REQUIRE(z && stats, "bn_apply: null pointer argument");
  RPNET_REQUIRE(y_f16 || y_pool_f16 || y_f32, "bn_apply: no output requested");
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "bn_apply: bad shape n=%d h=%d w=%d c=%d", n, h, w, c);
  RPNET_REQUIRE(!y_pool_f16 || (h % 2 == 0 && w % 2 == 0), "bn_apply: fused 2x2 max-pool needs even H, W (got %d x %d)", h, w);
  Groups gr;
  int rc = make_groups(&gr, group_start, groups, n);
  if (rc) return rc;
  RPNET_REQUIRE(256 % (c / 8) == 0, "bn_apply: c = %d must be a power of two in [8, 2048]", c);
  RPNET_REQUIRE((long long)n * h * w < (1LL << 31), "bn_apply: more than 2^31 pixels");
  const int lanes = 256 / (c / 8);
  const long long units = y_pool_f16 ? (long long)n * (h / 2) * (w / 2) : (long long)n * h * w;
  const int per_block = lanes * (y_pool_f16 ? 2 : 4);
  long long grid = (units + per_block - 1) / per_block;
  const long long cap_a = (long long)apply_blocks() * (units >= (1LL << 21) ? 2 : 1);    // large maps: more blocks in flight
  if (grid > cap_a) grid = cap_a;
  if (y_pool_f16)
    bn_apply_kernel<true><<<(unsigned)grid, 256, 0, stream>>>(static_cast<const uint4*>(z), static_cast<const uint4*>(z_lo), stats, gr, n,
                                                             h, w, c / 8, relu, static_cast<uint4*>(y_f16), static_cast<uint4*>(y_lo_f16),
                                                             y_f32, static_cast<uint4*>(y_pool_f16), static_cast<uint4*>(y_pool_lo_f16), lo_fmt, nullptr, nullptr);
  else
    bn_apply_kernel<false><<<(unsigned)grid, 256, 0, stream>>>(static_cast<const uint4*>(z), static_cast<const uint4*>(z_lo), stats, gr, n,
                                                              h, w, c / 8, relu, static_cast<uint4*>(y_f16),
                                                              static_cast<uint4*>(y_lo_f16), y_f32, nullptr, nullptr, lo_fmt, static_cast<const uint4*>(res_f16), res_lo);
  return check_cuda(cudaGetLastError(), "bn_apply launch");
}

// Backward of BatchNorm(batch stats)+ReLU for dz, in three launches: reduce -> finalize (dgamma, dbeta) -> apply.
RPNET_API int rpnet_bn_bwd(const void* z, const float* stats, int n, int h, int w, int c, const int* group_start, int groups, int relu,
                           const void* g_direct, int d_ld, int d_off, int d_is_f32, const void* g_pool_bf16, int p_ld, int p_off,
                           const void* g_up_bf16, int u_ld, int u_off, float* dgamma, float* dbeta, float* scratch /*[G][C][6]*/,
                           void* dz_bf16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(z && stats && scratch && dz_bf16, "bn_bwd: null pointer argument");
  RPNET_REQUIRE(g_direct || g_pool_bf16 || g_up_bf16, "bn_bwd: no gradient source");
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0 && c >= 64 && c % 8 == 0 && 256 % (c / 8) == 0, "bn_bwd: bad shape n=%d h=%d w=%d c=%d", n, h, w, c);
  RPNET_REQUIRE(!g_pool_bf16 || (h % 2 == 0 && w % 2 == 0), "bn_bwd: max-pool routing needs even H, W (got %d x %d)", h, w);
  RPNET_REQUIRE(d_ld % 8 == 0 && d_off % 8 == 0 && p_ld % 8 == 0 && p_off % 8 == 0 && u_ld % 8 == 0 && u_off % 8 == 0,
                "bn_bwd: gradient pitches / offsets must be multiples of 8 channels");
  Groups gr;
  int rc = make_groups(&gr, group_start, groups, n);
  if (rc) return rc;
  GradSrc src;
  src.direct = g_direct; src.d_ld = d_ld; src.d_off = d_off; src.d_f32 = d_is_f32;
  src.pooled = static_cast<const __nv_bfloat16*>(g_pool_bf16); src.p_ld = p_ld; src.p_off = p_off;
  src.up = static_cast<const __nv_bfloat16*>(g_up_bf16); src.u_ld = u_ld; src.u_off = u_off;
  RPNET_REQUIRE(reinterpret_cast<uintptr_t>(scratch) % 8 == 0, "bn_bwd: scratch must be 8-byte aligned");
  double* sums = reinterpret_cast<double*>(scratch);      // [G][C][2] fp64
  float* coef = scratch + (size_t)groups * c * 4;         // [G][C][2] fp32
  RPNET_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)groups * c * 2 * sizeof(double), stream));
  const bool win = g_pool_bf16 != nullptr;
  const int c8 = c / 8, lanes = 256 / c8;
  // reduction launches: 1-D grid, blocks per call group proportional to the group's size
  RPNET_REQUIRE((long long)n * h * w < (1LL << 31), "bn_bwd: more than 2^31 pixels");
  const int grid0 = plan_group_blocks(&gr, win ? (long long)(h / 2) * (w / 2) : (long long)h * w, (long long)lanes * (win ? 4 : 8));
  const uint4* zz = static_cast<const uint4*>(z);
  if (!win && g_direct && !d_is_f32 && !g_up_bf16) {       // direct-only bf16 gradient: specialised kernels
    const __nv_bfloat16* gd = static_cast<const __nv_bfloat16*>(g_direct);
    bn_bwd_direct_kernel<0><<<grid0, 256, 0, stream>>>(zz, stats, nullptr, gr, n, h * w, c8, relu, gd, d_ld, d_off,
                                                                            sums, nullptr);
    RPNET_CUDA_OK(cudaGetLastError());
    bn_bwd_finalize_kernel<<<(c + 127) / 128, 128, 0, stream>>>(sums, stats, gr, c, h * w, dgamma, dbeta, coef);
    RPNET_CUDA_OK(cudaGetLastError());
    long long nb1 = ((long long)n * h * w + lanes * 4 - 1) / (lanes * 4);
    const long long cap1 = (long long)apply_blocks() * ((long long)n * h * w >= (1LL << 21) ? 2 : 1);
    if (nb1 > cap1) nb1 = cap1;
    bn_bwd_direct_kernel<1><<<(unsigned)nb1, 256, 0, stream>>>(zz, stats, coef, gr, n, h * w, c8, relu, gd, d_ld, d_off, nullptr,
                                                              static_cast<uint4*>(dz_bf16));
    return check_cuda(cudaGetLastError(), "bn_bwd(direct) launch");
  }
  if (win && !g_direct && !g_up_bf16) {       // pool-only consumers: specialised kernels
    bn_bwd_pool_kernel<0><<<grid0, 256, 0, stream>>>(zz, stats, nullptr, gr, n, h, w, c8, relu, src.pooled,
                                                                             p_ld, p_off, sums, nullptr);
    RPNET_CUDA_OK(cudaGetLastError());
    bn_bwd_finalize_kernel<<<(c + 127) / 128, 128, 0, stream>>>(sums, stats, gr, c, h * w, dgamma, dbeta, coef);
    RPNET_CUDA_OK(cudaGetLastError());
    const long long units_all = (long long)n * (h / 2) * (w / 2);
    long long nb = (units_all + lanes - 1) / lanes;
    if (nb > 148LL * 12) nb = 148LL * 12;
    bn_bwd_pool_kernel<1><<<(unsigned)nb, 256, 0, stream>>>(zz, stats, coef, gr, n, h, w, c8, relu, src.pooled, p_ld, p_off, nullptr,
                                                           static_cast<uint4*>(dz_bf16));
    return check_cuda(cudaGetLastError(), "bn_bwd(pool) launch");
  }
  if (win) bn_bwd_kernel<true, 0><<<grid0, 256, 0, stream>>>(zz, stats, nullptr, gr, n, h, w, c8, relu, src, sums, nullptr);
  else     bn_bwd_kernel<false, 0><<<grid0, 256, 0, stream>>>(zz, stats, nullptr, gr, n, h, w, c8, relu, src, sums, nullptr);
  RPNET_CUDA_OK(cudaGetLastError());
  bn_bwd_finalize_kernel<<<(c + 127) / 128, 128, 0, stream>>>(sums, stats, gr, c, h * w, dgamma, dbeta, coef);
  RPNET_CUDA_OK(cudaGetLastError());
  const long long units = (long long)n * (win ? (h / 2) * (w / 2) : h * w);
  long long blocks2 = (units + lanes - 1) / lanes;
  if (blocks2 > 148LL * 16) blocks2 = 148LL * 16;
  if (win) bn_bwd_kernel<true, 1><<<(unsigned)blocks2, 256, 0, stream>>>(zz, stats, coef, gr, n, h, w, c8, relu, src, nullptr, static_cast<uint4*>(dz_bf16));
  else     bn_bwd_kernel<false, 1><<<(unsigned)blocks2, 256, 0, stream>>>(zz, stats, coef, gr, n, h, w, c8, relu, src, nullptr, static_cast<uint4*>(dz_bf16));
  return check_cuda(cudaGetLastError(), "bn_bwd launch");
}

RPNET_API int rpnet_upsample2x_f16(const void* x, void* y, int n, int h, int w, int c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(x && y, "upsample2x: null pointer argument");
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "upsample2x: bad shape n=%d h=%d w=%d c=%d", n, h, w, c);
  const long long total = (long long)n * 4 * h * w * (c / 8);
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const uint4*>(x), static_cast<uint4*>(y), n, h, w, c / 8);
  return check_cuda(cudaGetLastError(), "upsample2x launch");
}

RPNET_API int rpnet_cvt_f16_to_bf16(const void* in_f16, void* out_bf16, long long n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(in_f16 && out_bf16, "cvt_f16_to_bf16: null pointer argument");
  RPNET_REQUIRE(n > 0 && n % 8 == 0, "cvt_f16_to_bf16: element count must be a positive multiple of 8 (got %lld)", n);
  cvt_f16_bf16_kernel<<<grid_for(n / 8, 256), 256, 0, stream>>>(static_cast<const uint4*>(in_f16), static_cast<uint4*>(out_bf16), n / 8);
  return check_cuda(cudaGetLastError(), "cvt_f16_to_bf16 launch");
}

RPNET_API int rpnet_premask_bwd_bf16(const void* dxfg, const void* dxbg, const float* mask, int iters, long long pixels, int c,
                                      void* dx, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(dxfg && dxbg && mask && dx, "premask_bwd: null pointer argument");
  RPNET_REQUIRE(iters >= 1 && pixels > 0 && c > 0 && c % 8 == 0, "premask_bwd: bad shape iters=%d pixels=%lld c=%d", iters, pixels, c);
  premask_bwd_kernel<<<grid_for(pixels * (c / 8), 256), 256, 0, stream>>>(static_cast<const uint4*>(dxfg), static_cast<const uint4*>(dxbg),
                                                                            mask, iters, pixels, c / 8, static_cast<uint4*>(dx));
  return check_cuda(cudaGetLastError(), "premask_bwd launch");
}

RPNET_API int rpnet_pack_conv_weight_split(const float* w, int cout, int cin_real, int ntaps, int hole_start, int hole_len,
                                            void* w_fwd_f16, int split, void* w_dgrad_bf16, void* stream_);

RPNET_API int rpnet_premask_mask_bwd(const void* dxfg_bf16, const void* dxbg_bf16, const void* x_f16, long long pixels, int c, float* dmask,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(dxfg_bf16 && dxbg_bf16 && x_f16 && dmask, "premask_mask_bwd: null pointer argument");
  RPNET_REQUIRE(pixels > 0 && c > 0 && c % 8 == 0, "premask_mask_bwd: bad shape pixels=%lld c=%d", pixels, c);
  premask_mask_bwd_kernel<<<grid_for(pixels * 32, 256), 256, 0, stream>>>(static_cast<const uint4*>(dxfg_bf16), static_cast<const uint4*>(dxbg_bf16),
                                                                          static_cast<const uint4*>(x_f16), pixels, c / 8, dmask);
  return check_cuda(cudaGetLastError(), "premask_mask_bwd launch");
}

RPNET_API int rpnet_soft_mask_bwd_f32(const float* logits, const float* dmask, int batch, int n_classes, int h, int w, int scale,
                                       float* dlogits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(logits && dmask && dlogits, "soft_mask_bwd: null pointer argument");
  RPNET_REQUIRE(batch > 0 && n_classes >= 2 && n_classes <= kSoftMaxP && h > 0 && w > 0 && scale >= 1, "soft_mask_bwd: bad shape");
  soft_mask_bwd_kernel<<<grid_for((long long)batch * h * scale * w * scale, 256), 256, 0, stream>>>(logits, dmask, batch, n_classes, h * scale,
                                                                                                    w * scale, scale, dlogits);
  return check_cuda(cudaGetLastError(), "soft_mask_bwd launch");
}

RPNET_API int rpnet_pack_conv_weight(const float* w, int cout, int cin_real, int ntaps, int hole_start, int hole_len, void* w_fwd_f16,
                                      void* w_dgrad_bf16, void* stream_) {
  return rpnet_pack_conv_weight_split(w, cout, cin_real, ntaps, hole_start, hole_len, w_fwd_f16, 0, w_dgrad_bf16, stream_);
}

RPNET_API int rpnet_pack_conv_weight_split(const float* w, int cout, int cin_real, int ntaps, int hole_start, int hole_len,
                                            void* w_fwd_f16, int split, void* w_dgrad_bf16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(w && (w_fwd_f16 || w_dgrad_bf16), "pack_conv_weight: null pointer argument");
  RPNET_REQUIRE(cout > 0 && cin_real > 0 && ntaps > 0 && hole_len >= 0 && hole_start >= 0 && hole_start <= cin_real,
                "pack_conv_weight: bad shape");
  const int cin = cin_real + hole_len;
  RPNET_REQUIRE(ntaps <= 9, "pack_conv_weight: at most 9 taps (got %d)", ntaps);
  RPNET_REQUIRE(split >= 0 && split <= 2 && (split != 2 || cin % 64 == 0), "pack_conv_weight: split %d (fp8 corrections need cin %% 64 == 0)", split);
  const int tiles = ((cout + 31) / 32) * ((cin + 31) / 32);
  pack_conv_weight_kernel<<<tiles, 256, 0, stream>>>(w, cout, cin, ntaps, hole_start, hole_len, static_cast<__half*>(w_fwd_f16),
                                                     static_cast<__nv_bfloat16*>(w_dgrad_bf16), split);
  return check_cuda(cudaGetLastError(), "pack_conv_weight launch");
}

// See include/rpnet_b200.h for the contract.
RPNET_API int rpnet_pack_conv_weights(const rpnet_pack_desc* descs_host, int n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(descs_host && n >= 1 && n <= kMaxPackLayers, "pack_conv_weights: 1..%d layers (got %d)", kMaxPackLayers, n);
  PackTable t;
  t.n = n;
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    const rpnet_pack_desc& d = descs_host[i];
    RPNET_REQUIRE(d.w && (d.w_fwd_f16 || d.w_dgrad_bf16) && d.cout > 0 && d.cin_real > 0 && d.ntaps > 0 && d.ntaps <= 9 && d.hole_len >= 0 &&
                  d.hole_start >= 0 && d.hole_start <= d.cin_real, "pack_conv_weights: bad layer %d", i);
    PackLayer& L = t.l[i];
    L.w = d.w; L.wf = static_cast<__half*>(d.w_fwd_f16); L.wd = static_cast<__nv_bfloat16*>(d.w_dgrad_bf16);
    L.cout = d.cout; L.cin = d.cin_real + d.hole_len; L.ntaps = d.ntaps; L.hole_start = d.hole_start; L.hole_len = d.hole_len;
    RPNET_REQUIRE(d.split >= 0 && d.split <= 2 && (d.split != 2 || L.cin % 64 == 0), "pack_conv_weights: layer %d: split %d", i, d.split);
    L.split = d.split; L.first_tile = tiles;
    tiles += ((L.cout + 31) / 32) * ((L.cin + 31) / 32);
  }
  pack_conv_weights_kernel<<<tiles, 256, 0, stream>>>(t);
  return check_cuda(cudaGetLastError(), "pack_conv_weights launch");
}

RPNET_API int rpnet_conv3x3_first_wgrad_cin(const float* img, int cin, const void* dz_bf16, int n, int h, int w, float* grad,
                                             double* scratch576, void* stream_);

RPNET_API int rpnet_conv3x3_first_wgrad(const float* img, const void* dz_bf16, int n, int h, int w, float* grad, double* scratch576,
                                         void* stream_) {
  return rpnet_conv3x3_first_wgrad_cin(img, 1, dz_bf16, n, h, w, grad, scratch576, stream_);
}

RPNET_API int rpnet_conv3x3_first_wgrad_cin(const float* img, int cin, const void* dz_bf16, int n, int h, int w, float* grad,
                                             double* scratch576, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(img && dz_bf16 && grad && scratch576, "conv3x3_first_wgrad: null pointer argument");
  RPNET_REQUIRE(cin >= 1 && cin <= 4, "conv3x3_first_wgrad: cin %d out of range [1, 4]", cin);
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0 && (long long)n * h * w < (1LL << 31), "conv3x3_first_wgrad: bad shape");
  const long long total = (long long)n * h * w * 8;
  long long blocks = (total + 256 * 16 - 1) / (256 * 16);
  if (blocks > 148LL * 4) blocks = 148LL * 4;
  if (blocks < 1) blocks = 1;
  for (int ci = 0; ci < cin; ++ci) {                 // one pass per input plane: grad[co][ci][:] += sum dz[.., co] * img[.., ci, shifted]
    RPNET_CUDA_OK(cudaMemsetAsync(scratch576, 0, 576 * sizeof(double), stream));
    conv3x3_first_wgrad_kernel<<<(unsigned)blocks, 256, 0, stream>>>(img + (size_t)ci * h * w, (unsigned)(cin * h * w),
                                                                     static_cast<const uint4*>(dz_bf16), n, h, w, scratch576);
    RPNET_CUDA_OK(cudaGetLastError());
    add_first_wgrad_kernel<<<3, 192, 0, stream>>>(scratch576, grad, cin, ci);
    RPNET_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

RPNET_API int rpnet_relu_bias_bwd(const void* dy_bf16, const void* y_f16, long long pixels, int c, void* g_bf16, float* dbias,
                                   double* scratch_c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(dy_bf16 && g_bf16 && dbias && scratch_c, "relu_bias_bwd: null pointer argument");
  RPNET_REQUIRE(pixels > 0 && c >= 64 && c % 8 == 0 && 256 % (c / 8) == 0, "relu_bias_bwd: bad shape pixels=%lld c=%d", pixels, c);
  RPNET_CUDA_OK(cudaMemsetAsync(scratch_c, 0, (size_t)c * sizeof(double), stream));
  const int lanes = 256 / (c / 8);
  long long blocks = (pixels + lanes * 8 - 1) / (lanes * 8);
  if (blocks > 148LL * 4) blocks = 148LL * 4;
  relu_bias_bwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const uint4*>(dy_bf16), static_cast<const uint4*>(y_f16), pixels, c / 8,
                                                            static_cast<uint4*>(g_bf16), scratch_c);
  RPNET_CUDA_OK(cudaGetLastError());
  add_f64_to_f32_kernel<<<(c + 255) / 256, 256, 0, stream>>>(scratch_c, dbias, c);
  return check_cuda(cudaGetLastError(), "relu_bias_bwd launch");
}

// See include/rpnet_b200.h for the contracts.
RPNET_API int rpnet_add_relu_mask_bf16(const void* a_bf16, const void* b_bf16, const void* y_f16, void* out_bf16, long long elems,
                                        void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(a_bf16 && out_bf16 && elems > 0 && elems % 8 == 0, "add_relu_mask: bad arguments (elems = %lld)", elems);
  add_relu_mask_kernel<<<grid_for(elems / 8, 256), 256, 0, stream>>>(static_cast<const uint4*>(a_bf16), static_cast<const uint4*>(b_bf16),
                                                                    static_cast<const uint4*>(y_f16), static_cast<uint4*>(out_bf16), elems / 8);
  return check_cuda(cudaGetLastError(), "add_relu_mask launch");
}

RPNET_API int rpnet_conv7x7s2_stem_wgrad(const float* img, const void* dz_bf16, int n, int h, int w, float* grad, double* scratch9408,
                                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(img && dz_bf16 && grad && scratch9408, "conv7x7s2_stem_wgrad: null pointer argument");
  RPNET_REQUIRE(n > 0 && h >= 7 && w >= 7, "conv7x7s2_stem_wgrad: bad shape %d x %d x %d", n, h, w);
  RPNET_REQUIRE(reinterpret_cast<uintptr_t>(scratch9408) % 8 == 0, "conv7x7s2_stem_wgrad: scratch must be 8-byte aligned");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  RPNET_CUDA_OK(cudaMemsetAsync(scratch9408, 0, 64 * kStemTaps * sizeof(double), stream));
  const long long tiles = (long long)n * ((ho + kStemTile - 1) / kStemTile) * ((wo + kStemTile - 1) / kStemTile);
  const long long blocks = tiles < 148LL * 4 ? tiles : 148LL * 4;
  stem_wgrad_kernel<<<(unsigned)blocks, 256, 0, stream>>>(img, static_cast<const __nv_bfloat16*>(dz_bf16), n, h, w, ho, wo, scratch9408);
  RPNET_CUDA_OK(cudaGetLastError());
  add_f64_to_f32_kernel<<<(64 * kStemTaps + 255) / 256, 256, 0, stream>>>(scratch9408, grad, 64 * kStemTaps);
  return check_cuda(cudaGetLastError(), "conv7x7s2_stem_wgrad launch");
}

RPNET_API int rpnet_maxpool_bwd_bf16(const void* dy_bf16, const void* idx_u8, void* dx_bf16, int n, int h, int w, int c, int k, int stride,
                                      int pad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(dy_bf16 && idx_u8 && dx_bf16, "maxpool_bwd: null pointer argument");
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && k >= 1 && k <= 15 && stride >= 1 && pad >= 0 && 2 * pad <= k,
                "maxpool_bwd: bad shape / window");
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  maxpool_bwd_kernel<<<grid_for((long long)n * h * w * (c / 8), 256), 256, 0, stream>>>(
      static_cast<const uint4*>(dy_bf16), static_cast<const uint2*>(idx_u8), static_cast<uint4*>(dx_bf16), n, h, w, c / 8, ho, wo, k, stride, pad);
  return check_cuda(cudaGetLastError(), "maxpool_bwd launch");
}

RPNET_API int rpnet_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(param && grad && exp_avg && exp_avg_sq, "adam: null pointer argument");
  RPNET_REQUIRE(n > 0 && step >= 1, "adam: bad arguments n=%lld step=%d", n, step);
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<grid_for(n, 256), 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, bc1,
                                                    sqrtf(bc2), grad_scale);
  return check_cuda(cudaGetLastError(), "adam launch");
}

RPNET_API int rpnet_pack_upconv_weight_split(const float* w, int cout, int cin, void* wf_f16, int split, void* w16_bf16, void* stream_);

RPNET_API int rpnet_pack_upconv_weight(const float* w, int cout, int cin, void* wf_f16, void* w16_bf16, void* stream_) {
  return rpnet_pack_upconv_weight_split(w, cout, cin, wf_f16, 0, w16_bf16, stream_);
}

RPNET_API int rpnet_pack_upconv_weight_split(const float* w, int cout, int cin, void* wf_f16, int split, void* w16_bf16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(w && wf_f16 && w16_bf16 && cout > 0 && cin > 0, "pack_upconv_weight: bad argument");
  RPNET_REQUIRE(split >= 0 && split <= 2 && (split != 2 || cin % 64 == 0), "pack_upconv_weight: split %d (fp8 corrections need cin %% 64 == 0)", split);
  pack_upconv_weight_kernel<<<grid_for((long long)cout * cin, 256), 256, 0, stream>>>(w, cout, cin, static_cast<__half*>(wf_f16),
                                                                                    static_cast<__nv_bfloat16*>(w16_bf16), split);
  return check_cuda(cudaGetLastError(), "pack_upconv_weight launch");
}
