// Batched deformable ("demons") registration of slices, the `do_deformable: True` half of get_registration_field
// (dataset/few_shot_reader.py:109-198): DemonsRegistration with Diffeomorphic(10) (scaling and squaring), the NCC loss,
// torch.optim.Adam(lr 0.01) on the flow and GaussianRegulariser(sigma 2) after every step
// (net/registration.py:16-160, 190-313).  "Next" row N1 of SURVEY §8(f).
//
// The reference registers one slice at a time: 50 iterations, each ~25 grid_sample / elementwise launches forward, the
// same again through autograd, Adam and a conv2d.  Here ONE launch registers every slice of a volume: one CTA of 1024
// threads owns one slice and runs all iterations — forward composition chain, NCC reduction, the hand-derived backward
// through the chain, Adam and the 9x9 Gaussian smoothing — with block-level barriers between the phases; the
// intermediate displacement fields live in a per-slice global-memory workspace (L2 resident: 33 fields of h*w floats).
//
// Conventions reproduced from the reference (they matter for parity):
//   * the identity grid is corner aligned, g = 2 * (i / (n - 1) - 0.5)        (compute_grid, net/registration.py:171-186)
//   * F.grid_sample runs with its defaults: bilinear, zeros padding, align_corners=False, i.e. pixel = ((g + 1) * n - 1) / 2
//   * flow channel 0 is the x displacement, channel 1 the y displacement (normalised units)
//   * exp(flow): d = flow / 2^10; ten times d <- d + sample(d, g + d)           (Diffeomorphic.diffeomorphic_2D, :201-211)
#include "common.cuh"

namespace rpnet {

constexpr int kDemonsThreads = 1024;
constexpr int kMaxGauss = 15;             // Gaussian kernel side (sigma 2 -> 9)

struct GaussKernel {
  int ky, kx;
  float w[kMaxGauss * kMaxGauss];
};

struct Bilin {
  int x0, y0;
  float tx, ty;
  bool in00, in01, in10, in11;          // (y0, x0), (y0, x0 + 1), (y0 + 1, x0), (y0 + 1, x0 + 1) inside the image
};

// grid_sampler_unnormalize (align_corners = false) + corner bookkeeping of ATen's bilinear grid_sampler_2d
__device__ __forceinline__ Bilin bilin_at(float gx, float gy, int H, int W) {
  Bilin b;
  const float ix = ((gx + 1.f) * (float)W - 1.f) * 0.5f;
  const float iy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
  const float fx = floorf(ix), fy = floorf(iy);
  b.x0 = (int)fx; b.y0 = (int)fy;
  b.tx = ix - fx; b.ty = iy - fy;
  const bool x0ok = b.x0 >= 0 && b.x0 < W, x1ok = b.x0 + 1 >= 0 && b.x0 + 1 < W;
  const bool y0ok = b.y0 >= 0 && b.y0 < H, y1ok = b.y0 + 1 >= 0 && b.y0 + 1 < H;
  b.in00 = y0ok && x0ok; b.in01 = y0ok && x1ok; b.in10 = y1ok && x0ok; b.in11 = y1ok && x1ok;
  return b;
}

__device__ __forceinline__ void corners(const float* __restrict__ f, const Bilin& b, int W, float& nw, float& ne, float& sw, float& se) {
  const int base = b.y0 * W + b.x0;
  nw = b.in00 ? f[base] : 0.f;
  ne = b.in01 ? f[base + 1] : 0.f;
  sw = b.in10 ? f[base + W] : 0.f;
  se = b.in11 ? f[base + W + 1] : 0.f;
}

__device__ __forceinline__ float interp(const Bilin& b, float nw, float ne, float sw, float se) {
  // ATen: nw * (ix_se - ix) * (iy_se - iy) + ne * (ix - ix_sw) * (iy_sw - iy) + sw * (ix_ne - ix) * (iy - iy_ne) + se * ...
  return nw * (1.f - b.tx) * (1.f - b.ty) + ne * b.tx * (1.f - b.ty) + sw * (1.f - b.tx) * b.ty + se * b.tx * b.ty;
}

__device__ __forceinline__ float grid_x(int j, int W) { return 2.f * ((float)j / (float)(W - 1) - 0.5f); }

// block-wide sum of `v` (double), result broadcast to every thread; s_red: 32 doubles
__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                                  // s_red may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t += s_red[i];       // fixed order: deterministic
  return t;
}

// d_out = exp(flow): the ten compositions, D[0..scaling] are kept (the backward needs them).
__device__ void exp_flow(const float* __restrict__ flow, float* __restrict__ D, int H, int W, int scaling) {
  const int hw = H * W;
  const float inv = 1.f / (float)(1 << scaling);
  for (int p = threadIdx.x; p < 2 * hw; p += blockDim.x) D[p] = flow[p] * inv;      // displacement / 2^scaling
  __syncthreads();
  for (int k = 0; k < scaling; ++k) {
    const float* d = D + (size_t)k * 2 * hw;
    float* o = D + (size_t)(k + 1) * 2 * hw;
    for (int p = threadIdx.x; p < hw; p += blockDim.x) {
      const int i = p / W, j = p % W;
      const float dx = d[p], dy = d[hw + p];
      const Bilin b = bilin_at(grid_x(j, W) + dx, grid_x(i, H) + dy, H, W);
      float nw, ne, sw, se;
      corners(d, b, W, nw, ne, sw, se);
      o[p] = dx + interp(b, nw, ne, sw, se);
      corners(d + hw, b, W, nw, ne, sw, se);
      o[hw + p] = dy + interp(b, nw, ne, sw, se);
    }
    __syncthreads();
  }
}

// workspace per slice (floats): D [(scaling + 1)][2][hw] | G [2][2][hw] | m [2][hw] | v [2][hw] | tmp [2][hw]
__host__ __device__ inline size_t demons_ws_floats(int hw, int scaling) { return (size_t)((scaling + 1) * 2 + 4 + 2 + 2 + 2) * hw; }

__global__ void __launch_bounds__(kDemonsThreads, 1)
demons_register_kernel(const float* __restrict__ moving, const float* __restrict__ fixed, int H, int W, int iters, int scaling, float lr,
                       float beta1, float beta2, float eps, GaussKernel gk, float* __restrict__ flow_out, float* __restrict__ disp_out,
                       float* __restrict__ workspace, float* __restrict__ loss_curve) {
  __shared__ double s_red[32];
  __shared__ float s_g[kMaxGauss * kMaxGauss];
  const int n = blockIdx.x, hw = H * W;
  const float* mov = moving + (size_t)n * hw;
  const float* fix = fixed + (size_t)n * hw;
  float* flow = flow_out + (size_t)n * 2 * hw;
  float* ws = workspace + (size_t)n * demons_ws_floats(hw, scaling);
  float* D = ws;
  float* G = D + (size_t)(scaling + 1) * 2 * hw;
  float* am = G + (size_t)4 * hw;
  float* av = am + (size_t)2 * hw;
  float* tmp = av + (size_t)2 * hw;
  for (int i = threadIdx.x; i < gk.ky * gk.kx; i += blockDim.x) s_g[i] = gk.w[i];
  for (int p = threadIdx.x; p < 2 * hw; p += blockDim.x) { flow[p] = 0.f; am[p] = 0.f; av[p] = 0.f; }     // flow.data.fill_(0), :234
  // fixed image statistics (constant over the iterations)
  double sf = 0.0;
  for (int p = threadIdx.x; p < hw; p += blockDim.x) sf += (double)fix[p];
  const float mean_f = (float)(block_sum(sf, s_red) / hw);
  double sB = 0.0;
  for (int p = threadIdx.x; p < hw; p += blockDim.x) { const float f = fix[p] - mean_f; sB += (double)(f * f); }
  const float Bsum = (float)block_sum(sB, s_red);
  __syncthreads();
  const float gmx = 0.5f * (float)W, gmy = 0.5f * (float)H;                      // d pixel / d normalised coordinate
  float b1t = 1.f, b2t = 1.f;
  for (int it = 0; it < iters; ++it) {
    // ---------------- forward: exp(flow), warp, NCC (net/registration.py:244-258, 157-160)
    exp_flow(flow, D, H, W, scaling);
    const float* dl = D + (size_t)scaling * 2 * hw;
    float* warped = tmp;                                                          // [hw]
    double sw_ = 0.0;
    for (int p = threadIdx.x; p < hw; p += blockDim.x) {
      const int i = p / W, j = p % W;
      const Bilin b = bilin_at(grid_x(j, W) + dl[p], grid_x(i, H) + dl[hw + p], H, W);
      float nw, ne, sw, se;
      corners(mov, b, W, nw, ne, sw, se);
      const float wv = interp(b, nw, ne, sw, se);
      warped[p] = wv;
      sw_ += (double)wv;
    }
    const float mean_w = (float)(block_sum(sw_, s_red) / hw);
    double sA = 0.0, sC = 0.0;
    for (int p = threadIdx.x; p < hw; p += blockDim.x) {
      const float fm = fix[p] - mean_f, mm = warped[p] - mean_w;
      sA += (double)(fm * mm);
      sC += (double)(mm * mm);
    }
    const float A = (float)block_sum(sA, s_red);
    const float C = (float)block_sum(sC, s_red);
    const float Dn = sqrtf(Bsum * C + 1e-10f);
    if (loss_curve && threadIdx.x == 0) loss_curve[(size_t)n * iters + it] = -A / Dn;
    // ---------------- backward
    // d loss / d warped = -fm / Dn + A * B * mm / Dn^3 (the mean-subtraction terms vanish: sum fm = sum mm = 0);
    // d warped / d location through the bilinear weights of `moving` -> G_scaling
    const float kA = A * Bsum / (Dn * Dn * Dn), kF = -1.f / Dn;
    float* Gn = G;                                                                // gradient w.r.t. D[k + 1]
    float* Gk = G + (size_t)2 * hw;                                               // gradient w.r.t. D[k]
    for (int p = threadIdx.x; p < hw; p += blockDim.x) {
      const int i = p / W, j = p % W;
      const float gw = kF * (fix[p] - mean_f) + kA * (warped[p] - mean_w);
      const Bilin b = bilin_at(grid_x(j, W) + dl[p], grid_x(i, H) + dl[hw + p], H, W);
      float nw, ne, sw, se;
      corners(mov, b, W, nw, ne, sw, se);
      Gn[p] = gmx * gw * ((ne - nw) * (1.f - b.ty) + (se - sw) * b.ty);
      Gn[hw + p] = gmy * gw * ((sw - nw) * (1.f - b.tx) + (se - ne) * b.tx);
    }
    __syncthreads();
    for (int k = scaling - 1; k >= 0; --k) {
      // D[k+1](p) = D[k](p) + sum_q w_q(p) D[k](q), q = corners of the location g(p) + D[k](p)
      const float* d = D + (size_t)k * 2 * hw;
      // (a) own-pixel terms: identity + the dependence of the location on D[k](p)
      for (int p = threadIdx.x; p < hw; p += blockDim.x) {
        const int i = p / W, j = p % W;
        const float g0 = Gn[p], g1 = Gn[hw + p];
        const Bilin b = bilin_at(grid_x(j, W) + d[p], grid_x(i, H) + d[hw + p], H, W);
        float nw, ne, sw, se, gix, giy;
        corners(d, b, W, nw, ne, sw, se);
        gix = g0 * ((ne - nw) * (1.f - b.ty) + (se - sw) * b.ty);
        giy = g0 * ((sw - nw) * (1.f - b.tx) + (se - ne) * b.tx);
        corners(d + hw, b, W, nw, ne, sw, se);
        gix += g1 * ((ne - nw) * (1.f - b.ty) + (se - sw) * b.ty);
        giy += g1 * ((sw - nw) * (1.f - b.tx) + (se - ne) * b.tx);
        Gk[p] = g0 + gmx * gix;
        Gk[hw + p] = g1 + gmy * giy;
      }
      __syncthreads();
      // (b) the sampled values: scatter through the bilinear weights (grid_sampler backward w.r.t. its input)
      for (int p = threadIdx.x; p < hw; p += blockDim.x) {
        const int i = p / W, j = p % W;
        const float g0 = Gn[p], g1 = Gn[hw + p];
        const Bilin b = bilin_at(grid_x(j, W) + d[p], grid_x(i, H) + d[hw + p], H, W);
        const int base = b.y0 * W + b.x0;
        const float w00 = (1.f - b.tx) * (1.f - b.ty), w01 = b.tx * (1.f - b.ty), w10 = (1.f - b.tx) * b.ty, w11 = b.tx * b.ty;
        if (b.in00) { atomicAdd(Gk + base, w00 * g0); atomicAdd(Gk + hw + base, w00 * g1); }
        if (b.in01) { atomicAdd(Gk + base + 1, w01 * g0); atomicAdd(Gk + hw + base + 1, w01 * g1); }
        if (b.in10) { atomicAdd(Gk + base + W, w10 * g0); atomicAdd(Gk + hw + base + W, w10 * g1); }
        if (b.in11) { atomicAdd(Gk + base + W + 1, w11 * g0); atomicAdd(Gk + hw + base + W + 1, w11 * g1); }
      }
      __syncthreads();
      float* t = Gn; Gn = Gk; Gk = t;
    }
    // ---------------- Adam on the flow (torch.optim.Adam defaults but lr; few_shot_reader.py:148), gradient = G_0 / 2^scaling
    b1t *= beta1; b2t *= beta2;
    const float bc1 = 1.f - b1t, bc2s = sqrtf(1.f - b2t), ginv = 1.f / (float)(1 << scaling);
    for (int p = threadIdx.x; p < 2 * hw; p += blockDim.x) {
      const float g = Gn[p] * ginv;
      const float m = beta1 * am[p] + (1.f - beta1) * g;
      const float v = beta2 * av[p] + (1.f - beta2) * g * g;
      am[p] = m; av[p] = v;
      flow[p] -= (lr / bc1) * (m / (sqrtf(v) / bc2s + eps));
    }
    __syncthreads();
    // ---------------- GaussianRegulariser: flow <- conv2d(flow, gaussian, zero padding, groups = 2)  (net/registration.py:128-133)
    const int ry = gk.ky / 2, rx = gk.kx / 2;
    for (int p = threadIdx.x; p < 2 * hw; p += blockDim.x) {
      const int c = p / hw, q = p % hw, i = q / W, j = q % W;
      const float* f = flow + (size_t)c * hw;
      float acc = 0.f;
      for (int a = 0; a < gk.ky; ++a) {
        const int y = i + a - ry;
        if (y < 0 || y >= H) continue;
        for (int bb = 0; bb < gk.kx; ++bb) {
          const int x = j + bb - rx;
          if (x < 0 || x >= W) continue;
          acc = fmaf(s_g[a * gk.kx + bb], f[y * W + x], acc);
        }
      }
      tmp[p] = acc;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < 2 * hw; p += blockDim.x) flow[p] = tmp[p];
    __syncthreads();
  }
  // the displacement the trained module applies: exp(final flow)
  exp_flow(flow, D, H, W, scaling);
  const float* dl = D + (size_t)scaling * 2 * hw;
  float* out = disp_out + (size_t)n * 2 * hw;
  for (int p = threadIdx.x; p < 2 * hw; p += blockDim.x) out[p] = dl[p];
}

// out[n][c] = grid_sample(x[n][c], g + disp[n]) — DemonsRegistration.forward with the displacement already exponentiated.
__global__ void demons_warp_kernel(const float* __restrict__ x, const float* __restrict__ disp, float* __restrict__ out, int N, int C,
                                   int H, int W) {
  const int hw = H * W;
  const long long total = (long long)N * hw;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(t / hw), p = (int)(t % hw), i = p / W, j = p % W;
    const float* d = disp + (size_t)n * 2 * hw;
    const Bilin b = bilin_at(grid_x(j, W) + d[p], grid_x(i, H) + d[hw + p], H, W);
    for (int c = 0; c < C; ++c) {
      float nw, ne, sw, se;
      corners(x + ((size_t)n * C + c) * hw, b, W, nw, ne, sw, se);
      out[((size_t)n * C + c) * hw + p] = interp(b, nw, ne, sw, se);
    }
  }
}

}  // namespace rpnet

using namespace rpnet;

RPNET_API long long rpnet_demons_workspace_bytes(int n, int h, int w, int scaling) {
  if (n <= 0 || h <= 1 || w <= 1 || scaling < 0 || scaling > 16) return -2;
  return (long long)n * (long long)demons_ws_floats(h * w, scaling) * 4;
}

// See include/rpnet_b200.h for the contract.
RPNET_API int rpnet_demons_register_f32(const float* moving, const float* fixed, int n, int h, int w, int iters, float lr, float beta1,
                                         float beta2, float eps, int scaling, const float* gauss_host, int gauss_h, int gauss_w,
                                         float* flow, float* disp, float* loss_curve, void* workspace, long long workspace_bytes,
                                         void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(moving && fixed && flow && disp && workspace && gauss_host, "demons_register: null pointer argument");
  RPNET_REQUIRE(n > 0 && h > 1 && w > 1 && iters >= 0 && scaling >= 0 && scaling <= 16, "demons_register: bad shape n=%d h=%d w=%d", n, h, w);
  RPNET_REQUIRE(gauss_h >= 1 && gauss_w >= 1 && gauss_h <= kMaxGauss && gauss_w <= kMaxGauss && (gauss_h & 1) && (gauss_w & 1),
                "demons_register: Gaussian kernel %d x %d not supported (odd sides up to %d)", gauss_h, gauss_w, kMaxGauss);
  RPNET_REQUIRE(workspace_bytes >= rpnet_demons_workspace_bytes(n, h, w, scaling), "demons_register: workspace too small (%lld bytes)", workspace_bytes);
  GaussKernel gk;
  gk.ky = gauss_h; gk.kx = gauss_w;
  for (int i = 0; i < gauss_h * gauss_w; ++i) gk.w[i] = gauss_host[i];
  demons_register_kernel<<<n, kDemonsThreads, 0, stream>>>(moving, fixed, h, w, iters, scaling, lr, beta1, beta2, eps, gk, flow, disp,
                                                          static_cast<float*>(workspace), loss_curve);
  return check_cuda(cudaGetLastError(), "demons_register launch");
}

RPNET_API int rpnet_demons_warp_f32(const float* x, const float* disp, float* out, int n, int c, int h, int w, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(x && disp && out, "demons_warp: null pointer argument");
  RPNET_REQUIRE(n > 0 && c > 0 && h > 1 && w > 1, "demons_warp: bad shape");
  const long long total = (long long)n * h * w;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  demons_warp_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, disp, out, n, c, h, w);
  return check_cuda(cudaGetLastError(), "demons_warp launch");
}
