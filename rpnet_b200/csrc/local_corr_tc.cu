// Local (2r+1)^2 correlation of the context-relation encoder on the 5th-gen tensor cores.
//
//   reference: Correlation (net/rp_net.py:153-181) — all-pairs bmm + integer-coordinate grid_sample, i.e. exactly
//     out[n,y,x, a*(2r+1)+b] = 1/sqrt(C) * sum_c f1[n,y,x,c] * f2[n, y+(b-r), x+(a-r), c]      (zero outside the map;
//     a -> column offset, b -> row offset: the RAFT x/y channel order).
//
// GEMM view per pixel tile (8 wide x 16 tall = 128 pixels): D[128 px][NH halo px] = F1[128 px][C] * F2halo[NH][C]^T with the
// halo = the (8+2r) x (16+2r) window of f2 around the tile; the wanted band (|dx|,|dy| <= r) is 121 of the NH = 468 columns
// (r = 5), i.e. 26 % of the MMA work is useful — still ~30x faster than the CUDA-core FMA form because the contraction
// (C = 256) runs on tcgen05.  TMA out-of-bounds zero fill is the zero padding of the correlation window.
//   warp 0: TMA producer (f1 tile box + f2 halo box per 64-channel chunk), warp 1: MMA issuer (fp32 accumulators: NH <= 480
//   TMEM columns), warps 2..5: epilogue — every lane owns one pixel, walks the accumulator columns its warp's rows can need
//   (static register indices; the band test and the output channel are lane arithmetic), stages the fp16 result in shared
//   memory and the tile leaves as coalesced 16-byte stores.
#include "common.cuh"

namespace rpnet {

constexpr int kLcTW = 8, kLcTH = 16;        // pixel tile

// 2-byte store through the shared window (32-bit address arithmetic, STS instead of a generic ST)
__device__ __forceinline__ void sts_f16(uint32_t saddr, float v) {
  const __half h = __float2half_rn(v);
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(*reinterpret_cast<const unsigned short*>(&h)) : "memory");
}
constexpr int kLcThreads = 192;
constexpr int kLcStages = 2;
constexpr int kLcMaxOutC = 128;

template <int R>
struct LcCfg {
  static constexpr int K = 2 * R + 1;
  static constexpr int HW = kLcTW + 2 * R, HH = kLcTH + 2 * R;
  static constexpr int NH = HW * HH;                              // halo pixels = accumulator columns in use
  static constexpr int NPAD = (NH + 31) / 32 * 32;
  static constexpr int NMMA = NPAD > 256 ? NPAD / 2 : NPAD;       // columns per tcgen05.mma (<= 256, % 16 == 0)
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = NPAD * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 512;                           // the epilogue reads 32 columns from every halo row start
  static constexpr int kOutPitch = kLcMaxOutC / 2 + 1;            // 32-bit words per staged pixel row (odd: no bank conflicts)
  static constexpr int kSmemBytes = kLcStages * kStageBytes + 1024 + 128 * kOutPitch * 4 + 256;
};

struct LcParams {
  int N, H, W, chunks;
  int tiles_x, tiles_y;
  int out_c;
  float scale;
  __half* out;
};

template <int R>
__global__ void __launch_bounds__(kLcThreads, 1)
local_corr_tc_kernel(const __grid_constant__ CUtensorMap tm_f1, const __grid_constant__ CUtensorMap tm_f2, const LcParams p) {
  using Cfg = LcCfg<R>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint32_t* s_out = reinterpret_cast<uint32_t*>(tiles + kLcStages * Cfg::kStageBytes);       // [128][kOutPitch]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_out + 128 * Cfg::kOutPitch);
  uint64_t* empty_bar = full_bar + kLcStages;
  uint64_t* tfull_bar = empty_bar + kLcStages;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_x * p.tiles_y * p.N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_f1);
    tma_prefetch_desc(&tm_f2);
    for (int i = 0; i < kLcStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int t = tile;
        const int tx = t % p.tiles_x;  t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int n = t / p.tiles_y;
        const int x0 = tx * kLcTW, y0 = ty * kLcTH;
        for (int kc = 0; kc < p.chunks; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = tiles + stage * Cfg::kStageBytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kABytes + Cfg::NH * 128);
          tma_load_4d(&tm_f1, &full_bar[stage], a_dst, kc * 64, x0, y0, n);
          tma_load_4d(&tm_f2, &full_bar[stage], a_dst + Cfg::kABytes, kc * 64, x0 - R, y0 - R, n);
          if (++stage == kLcStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(128, Cfg::NMMA);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        mbar_wait(tempty_bar, (it & 1) ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < p.chunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + stage * Cfg::kStageBytes);
          const uint64_t a_desc = umma_desc_sw128(a_addr, 1024);
#pragma unroll
          for (int half = 0; half < Cfg::NPAD / Cfg::NMMA; ++half) {
            const uint64_t b_desc = umma_desc_sw128(a_addr + Cfg::kABytes + half * Cfg::NMMA * 128, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + half * Cfg::NMMA, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kLcStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar);
      }
    }
  } else {
    const int q = warp & 3;                                  // TMEM lane quarter: accumulator rows 32q .. 32q+31
    const int row = q * 32 + lane;
    const int px = row & (kLcTW - 1), py = row >> 3;         // py = 4q + (lane >> 3)
    const int et = threadIdx.x - 64;
    const int pairs = p.out_c >> 1;                          // 32-bit words per output pixel
    const uint32_t s_row = smem_u32(s_out + row * Cfg::kOutPitch);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int t = tile;
      const int tx = t % p.tiles_x;  t /= p.tiles_x;
      const int ty = t % p.tiles_y;
      const int n = t / p.tiles_y;
      const int x0 = tx * kLcTW, y0 = ty * kLcTH;
      // padding channels [K*K, out_c) are zero by contract
      for (int c = Cfg::K * Cfg::K; c < p.out_c; ++c) sts_f16(s_row + 2 * c, 0.f);
      mbar_wait(tfull_bar, it & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      // the rows of this warp (py = 4q .. 4q+3) can need halo rows 4q .. 4q + 3 + 2R; one 32-column load per halo row
      // (columns hy*HW .. +HW-1 are that row): hx = j is a static register index, the band test is lane arithmetic
      // two halo rows per step: both TMEM loads are in flight before the (independent) selection code of either row runs
#pragma unroll 1
      for (int hy = 4 * q; hy < 4 * q + 4 + 2 * R; hy += 2) {
        float v0[32], v1[32];
        tmem_ld32(t_addr + hy * Cfg::HW, v0);
        tmem_ld32(t_addr + (hy + 1) * Cfg::HW, v1);
        tmem_ld_wait();
        const unsigned b0 = (unsigned)(hy - py), b1 = b0 + 1u;
        const uint32_t dst0 = s_row + 2 * ((int)b0 - px * Cfg::K);   // + 2 * a * K with a = j - px
        const bool r0 = b0 < (unsigned)Cfg::K, r1 = b1 < (unsigned)Cfg::K;
#pragma unroll
        for (int j = 0; j < Cfg::HW; ++j) {
          const bool in = (unsigned)(j - px) < (unsigned)Cfg::K;
          if (in && r0) sts_f16(dst0 + 2 * j * Cfg::K, v0[j] * p.scale);
          if (in && r1) sts_f16(dst0 + 2 + 2 * j * Cfg::K, v1[j] * p.scale);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);               // accumulator drained: the next tile's MMAs may start
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // coalesced copy-out: 128 pixels x out_c fp16
      for (int i = et; i < 128 * pairs; i += 128) {
        const int r = i / pairs, w = i - r * pairs;
        const int x = x0 + (r & (kLcTW - 1)), y = y0 + (r >> 3);
        if (x < p.W && y < p.H)
          reinterpret_cast<uint32_t*>(p.out + ((size_t)(n * p.H + y) * p.W + x) * p.out_c)[w] = s_out[r * Cfg::kOutPitch + w];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int R>
static int launch_corr_tc(const void* f1, const void* f2, void* out, int n, int h, int w, int c, int out_c, cudaStream_t stream) {
  using Cfg = LcCfg<R>;
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(local_corr_tc_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  CUtensorMap t1, t2;
  const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
  const uint64_t str[3] = {(uint64_t)c, (uint64_t)c * w, (uint64_t)c * w * h};
  const uint32_t box1[4] = {64u, (uint32_t)kLcTW, (uint32_t)kLcTH, 1u};
  const uint32_t box2[4] = {64u, (uint32_t)Cfg::HW, (uint32_t)Cfg::HH, 1u};
  int rc = make_tmap_2b(&t1, f1, 4, dims, str, box1, false);
  if (rc) return rc;
  rc = make_tmap_2b(&t2, f2, 4, dims, str, box2, false);
  if (rc) return rc;
  LcParams p{};
  p.N = n; p.H = h; p.W = w; p.chunks = c / 64;
  p.tiles_x = (w + kLcTW - 1) / kLcTW; p.tiles_y = (h + kLcTH - 1) / kLcTH;
  p.out_c = out_c;
  p.scale = 1.0f / sqrtf((float)c);
  p.out = static_cast<__half*>(out);
  const int tiles = p.tiles_x * p.tiles_y * n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  local_corr_tc_kernel<R><<<grid, kLcThreads, Cfg::kSmemBytes, stream>>>(t1, t2, p);
  return check_cuda(cudaGetLastError(), "local_corr_tc launch");
}

// Tensor-core path of rpnet_local_corr_f16 (dispatch in stream_kernels.cu).  Returns 1 if the shape is not eligible.
int local_corr_tc(const void* f1, const void* f2, void* out, int n, int h, int w, int c, int radius, int out_c, cudaStream_t stream) {
  if (c % 64 != 0 || out_c > kLcMaxOutC || out_c % 8 != 0) return 1;
  switch (radius) {
    case 1: return launch_corr_tc<1>(f1, f2, out, n, h, w, c, out_c, stream);
    case 2: return launch_corr_tc<2>(f1, f2, out, n, h, w, c, out_c, stream);
    case 3: return launch_corr_tc<3>(f1, f2, out, n, h, w, c, out_c, stream);
    case 4: return launch_corr_tc<4>(f1, f2, out, n, h, w, c, out_c, stream);
    case 5: return launch_corr_tc<5>(f1, f2, out, n, h, w, c, out_c, stream);
    default: return 1;
  }
}


// =====================================================================================================================
// Fused relation head of the eval forward (net/rp_net.py:77-84 tail + :287-303): local correlation -> cat([corr, fm1]) ->
// cre.q 1x1 conv + folded BN + ReLU -> calDist against the prototypes, one kernel, nothing but the (1+Wa) cosine maps leaves
// the SM.  Per 8 x 16 pixel tile:
//   phase 1  D1[128 px][NH] = F1 * F2halo^T over the 64-channel chunks (as local_corr_tc_kernel);
//   extract  the band of D1 -> fp16 K-major swizzled operand tile in shared memory (the corr block of the concat);
//   phase 2  D2[128 px][64] = [corr | fm1] * Wq^T: the f1 chunks are re-fetched (L2-hot) and Wq arrives into the two pipeline
//            slots the main loop has just drained; D2 reuses TMEM columns 0..63;
//   epilogue affine + ReLU + cosine vs prototypes -> pred[n][p][pixel].
// =====================================================================================================================
struct RhParams {
  int N, H, W, chunks;
  int tiles_x, tiles_y;
  float corr_scale;
  const float* scale;       // folded BN of cre.q: y = relu(acc * scale + shift)
  const float* shift;
  const float* protos;      // [sets][P][64]
  int P, sets;
  float cos_scaler;
  float* pred;              // [n][P][h*w]
};

constexpr int kRhThreads = 320;             // warp 0 TMA, warp 1 MMA, warps 2..9: two epilogue warps per TMEM lane quarter

template <int R>
__global__ void __launch_bounds__(kRhThreads, 1)
relation_head_kernel(const __grid_constant__ CUtensorMap tm_f1, const __grid_constant__ CUtensorMap tm_f2,
                     const __grid_constant__ CUtensorMap tm_w, const RhParams p) {
  using Cfg = LcCfg<R>;
  constexpr int kMaxP = 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_corr = tiles + kLcStages * Cfg::kStageBytes;                     // 2 x [128 rows x 64 K] fp16, K-major SW128
  uint32_t* s_stage = reinterpret_cast<uint32_t*>(s_corr + 2 * 16384);        // [128 rows][65 words]: channel-linear fp16 rows
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_stage + 128 * Cfg::kOutPitch);
  uint64_t* empty_bar = full_bar + kLcStages;
  uint64_t* tfull_bar = empty_bar + kLcStages;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint64_t* corr_bar = tempty_bar + 1;
  uint64_t* d2_bar = corr_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_x * p.tiles_y * p.N;
  const int kq = 2 + p.chunks;                                               // 64-wide K chunks of the 1x1 conv

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_f1);
    tma_prefetch_desc(&tm_f2);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < kLcStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);
    mbar_init(corr_bar, 1);
    mbar_init(d2_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int t = tile;
        const int tx = t % p.tiles_x;  t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int n = t / p.tiles_y;
        const int x0 = tx * kLcTW, y0 = ty * kLcTH;
        for (int kc = 0; kc < p.chunks; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = tiles + stage * Cfg::kStageBytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kABytes + Cfg::NH * 128);
          tma_load_4d(&tm_f1, &full_bar[stage], a_dst, kc * 64, x0, y0, n);
          tma_load_4d(&tm_f2, &full_bar[stage], a_dst + Cfg::kABytes, kc * 64, x0 - R, y0 - R, n);
          if (++stage == kLcStages) { stage = 0; phase ^= 1; }
        }
        // phase 2, slot A: all f1 chunks of the tile again (the A operand of the fm1 half of the 1x1 conv)
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], p.chunks * Cfg::kABytes);
        for (int kc = 0; kc < p.chunks; ++kc)
          tma_load_4d(&tm_f1, &full_bar[stage], tiles + stage * Cfg::kStageBytes + kc * Cfg::kABytes, kc * 64, x0, y0, n);
        if (++stage == kLcStages) { stage = 0; phase ^= 1; }
        // phase 2, slot B: the packed 1x1 weights [64 cout][64 K] per K chunk
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], kq * 8192);
        for (int kc = 0; kc < kq; ++kc)
          tma_load_3d(&tm_w, &full_bar[stage], tiles + stage * Cfg::kStageBytes + kc * 8192, kc * 64, 0, 0);
        if (++stage == kLcStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc_f16(128, Cfg::NMMA);
      const uint32_t idesc2 = umma_idesc_f16(128, 64);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        mbar_wait(tempty_bar, (it & 1) ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < p.chunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + stage * Cfg::kStageBytes);
          const uint64_t a_desc = umma_desc_sw128(a_addr, 1024);
#pragma unroll
          for (int half = 0; half < Cfg::NPAD / Cfg::NMMA; ++half) {
            const uint64_t b_desc = umma_desc_sw128(a_addr + Cfg::kABytes + half * Cfg::NMMA * 128, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + half * Cfg::NMMA, a_desc + 2 * k, b_desc + 2 * k, idesc1, (kc | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kLcStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar);
        // ---- phase 2
        const int sa = stage;
        const uint32_t pa = phase;
        if (++stage == kLcStages) { stage = 0; phase ^= 1; }
        const int sb = stage;
        const uint32_t pb = phase;
        if (++stage == kLcStages) { stage = 0; phase ^= 1; }
        mbar_wait(&full_bar[sa], pa);
        mbar_wait(&full_bar[sb], pb);
        mbar_wait(corr_bar, it & 1);                     // band extracted (D1 fully read) and staged as an operand tile
        tc_fence_after();
        const uint32_t f1_addr = smem_u32(tiles + sa * Cfg::kStageBytes);
        const uint32_t w_addr = smem_u32(tiles + sb * Cfg::kStageBytes);
        const uint32_t c_addr = smem_u32(s_corr);
        for (int kc = 0; kc < kq; ++kc) {
          const uint64_t a_desc = umma_desc_sw128(kc < 2 ? c_addr + kc * 16384 : f1_addr + (kc - 2) * Cfg::kABytes, 1024);
          const uint64_t b_desc = umma_desc_sw128(w_addr + kc * 8192, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc2, (kc | k) != 0);
        }
        umma_commit(&empty_bar[sa]);
        umma_commit(&empty_bar[sb]);
        umma_commit(d2_bar);
      }
    }
  } else {
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;                        // the two warps of a lane quarter split the halo rows and the chunks
    const int row = q * 32 + lane;
    const int px = row & (kLcTW - 1), py = row >> 3;
    const uint32_t s_crow = smem_u32(s_corr) + row * 128;
    const uint32_t s_srow = smem_u32(s_stage + row * Cfg::kOutPitch);
    const int sw = row & 7;
    if (part == 0)
      for (int ch = Cfg::K * Cfg::K; ch < 128; ++ch) sts_f16(s_srow + 2 * ch, 0.f);   // padding channels [K*K, 128): zero, never rewritten
    constexpr int kSteps = (4 + 2 * R + 1) / 2, kSteps0 = (kSteps + 1) / 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int t = tile;
      const int tx = t % p.tiles_x;  t /= p.tiles_x;
      const int ty = t % p.tiles_y;
      const int n = t / p.tiles_y;
      const int x = tx * kLcTW + px, y = ty * kLcTH + py;
      const bool valid = x < p.W && y < p.H;
      mbar_wait(tfull_bar, it & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      // band extraction into the pixel's staging row (channel-linear fp16, odd word pitch: conflict-free) ...
      const int hy_lo = 4 * q + (part ? 2 * kSteps0 : 0), hy_hi = 4 * q + (part ? 4 + 2 * R : 2 * kSteps0);
#pragma unroll 1
      for (int hy = hy_lo; hy < hy_hi; hy += 2) {
        float v0[32], v1[32];
        tmem_ld32(t_addr + hy * Cfg::HW, v0);
        tmem_ld32(t_addr + (hy + 1) * Cfg::HW, v1);
        tmem_ld_wait();
        const unsigned b0 = (unsigned)(hy - py), b1 = b0 + 1u;
        const uint32_t dst0 = s_srow + 2 * ((int)b0 - px * Cfg::K);   // channel = (j - px) * K + b
        const bool r0 = b0 < (unsigned)Cfg::K, r1 = b1 < (unsigned)Cfg::K;
#pragma unroll
        for (int j = 0; j < Cfg::HW; ++j) {
          const bool in = (unsigned)(j - px) < (unsigned)Cfg::K;
          if (in && r0) sts_f16(dst0 + 2 * j * Cfg::K, v0[j] * p.corr_scale);
          if (in && r1) sts_f16(dst0 + 2 + 2 * j * Cfg::K, v1[j] * p.corr_scale);
        }
      }
      tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");          // both halves of every staging row are written
      // ... then the row moves into the K-major 128B-swizzled operand tile as 16-byte stores (8 lanes cover the 8 swizzle
      // phases: conflict-free); each warp of the pair moves one 64-channel K block
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const int ck = part * 8 + c8;
        uint32_t w0, w1, w2, w3;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(s_srow + ck * 16) : "memory");
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(s_srow + ck * 16 + 4) : "memory");
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w2) : "r"(s_srow + ck * 16 + 8) : "memory");
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w3) : "r"(s_srow + ck * 16 + 12) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(s_crow + part * 16384 + ((c8 ^ sw) << 4)), "r"(w0), "r"(w1),
                     "r"(w2), "r"(w3) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 64) mbar_arrive(corr_bar);
      if (part) continue;                                     // the second warp of the pair goes on to wait for the next tile
      // ---- phase-2 epilogue: affine + ReLU + cosine against the prototypes of this image
      mbar_wait(d2_bar, it & 1);
      tc_fence_after();
      // D2 leaves TMEM first (64 fp32 per lane) and the accumulator is handed back at once: the affine / cosine math below
      // overlaps the next tile's correlation MMAs
      float v[64];
      tmem_ld32(t_addr, v);
      tmem_ld32(t_addr + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
      float nn = 0.f, pn[kMaxP], dot[kMaxP];
#pragma unroll
      for (int k = 0; k < kMaxP; ++k) { pn[k] = 0.f; dot[k] = 0.f; }
      const float* pr_base = p.protos + (size_t)(n % p.sets) * p.P * 64;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        v[j] = fmaxf(fmaf(v[j], __ldg(p.scale + j), __ldg(p.shift + j)), 0.f);
        nn = fmaf(v[j], v[j], nn);
      }
#pragma unroll
      for (int k = 0; k < kMaxP; ++k) {
        if (k < p.P) {
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            const float4 pr = __ldg(reinterpret_cast<const float4*>(pr_base + k * 64 + j));
            dot[k] = fmaf(v[j], pr.x, fmaf(v[j + 1], pr.y, fmaf(v[j + 2], pr.z, fmaf(v[j + 3], pr.w, dot[k]))));
            pn[k] = fmaf(pr.x, pr.x, fmaf(pr.y, pr.y, fmaf(pr.z, pr.z, fmaf(pr.w, pr.w, pn[k]))));
          }
        }
      }
      if (valid) {
        const float xn = fmaxf(sqrtf(nn), 1e-8f);
        const size_t hw = (size_t)p.H * p.W;
#pragma unroll
        for (int k = 0; k < kMaxP; ++k)
          if (k < p.P) p.pred[((size_t)n * p.P + k) * hw + (size_t)y * p.W + x] = p.cos_scaler * (dot[k] / (xn * fmaxf(sqrtf(pn[k]), 1e-8f)));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int R>
static int launch_relation_head(const void* f1, const void* f2, const void* wq, const float* scale, const float* shift,
                                const float* protos, int P, int sets, float cos_scaler, float* pred, int n, int h, int w, int c,
                                cudaStream_t stream) {
  using Cfg = LcCfg<R>;
  constexpr int kSmem = kLcStages * Cfg::kStageBytes + 2 * 16384 + 128 * Cfg::kOutPitch * 4 + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(relation_head_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set = true;
  }
  const int chunks = c / 64;
  if (chunks * Cfg::kABytes > Cfg::kStageBytes || (2 + chunks) * 8192 > Cfg::kStageBytes) return 1;   // phase-2 tiles must fit a slot
  CUtensorMap t1, t2, tw;
  const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
  const uint64_t str[3] = {(uint64_t)c, (uint64_t)c * w, (uint64_t)c * w * h};
  const uint32_t box1[4] = {64u, (uint32_t)kLcTW, (uint32_t)kLcTH, 1u};
  const uint32_t box2[4] = {64u, (uint32_t)Cfg::HW, (uint32_t)Cfg::HH, 1u};
  int rc = make_tmap_2b(&t1, f1, 4, dims, str, box1, false);
  if (rc) return rc;
  rc = make_tmap_2b(&t2, f2, 4, dims, str, box2, false);
  if (rc) return rc;
  const uint64_t cin = 128 + (uint64_t)c;
  const uint64_t wdims[3] = {cin, 64, 1};
  const uint64_t wstr[2] = {cin, cin * 64};
  const uint32_t wbox[3] = {64u, 64u, 1u};
  rc = make_tmap_2b(&tw, wq, 3, wdims, wstr, wbox, false);
  if (rc) return rc;
  RhParams p{};
  p.N = n; p.H = h; p.W = w; p.chunks = chunks;
  p.tiles_x = (w + kLcTW - 1) / kLcTW; p.tiles_y = (h + kLcTH - 1) / kLcTH;
  p.corr_scale = 1.0f / sqrtf((float)c);
  p.scale = scale; p.shift = shift; p.protos = protos; p.P = P; p.sets = sets; p.cos_scaler = cos_scaler; p.pred = pred;
  const int tiles = p.tiles_x * p.tiles_y * n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  relation_head_kernel<R><<<grid, kRhThreads, kSmem, stream>>>(t1, t2, tw, p);
  return check_cuda(cudaGetLastError(), "relation_head launch");
}

// Dispatch for rpnet_relation_head_f16 (stream_kernels.cu).  Returns 1 if the shape is not eligible.
int relation_head_tc(const void* f1, const void* f2, const void* wq, const float* scale, const float* shift, const float* protos,
                     int P, int sets, float cos_scaler, float* pred, int n, int h, int w, int c, int radius, cudaStream_t stream) {
  if (c % 64 != 0 || P < 1 || P > 8 || (2 * radius + 1) * (2 * radius + 1) > 128) return 1;
  switch (radius) {
    case 3: return launch_relation_head<3>(f1, f2, wq, scale, shift, protos, P, sets, cos_scaler, pred, n, h, w, c, stream);
    case 5: return launch_relation_head<5>(f1, f2, wq, scale, shift, protos, P, sets, cos_scaler, pred, n, h, w, c, stream);
    default: return 1;
  }
}

// =====================================================================================================================
// Backward of the local correlation on tensor cores.
//   df1[p, c] = scale * sum_{a,b} dcorr[p, (a,b)] * f2[p + (b-r, a-r), c]  (+ the direct gradient of fm1 from cat([corr, fm1]))
//   df2[p',c] = scale * sum_{a,b} dcorr[p' - (b-r, a-r), (a,b)] * f1[p' - (b-r, a-r), c]
// Both are the same "band GEMM"  D[128 px][64 ch] = Band[128 px][NH halo px] * Fhalo[NH][64 ch]  once the gradient of the
// second form is re-indexed per destination pixel (corr_grad_transpose_kernel):
//   dcorrT[p', (a',b')] = dcorr[p' + (b'-r, a'-r), (2r-a', 2r-b')]   (0 outside the map)   =>   df2 = BandGEMM(dcorrT, f1).
// Band[m][n] is non-zero only for the (2r+1)^2 halo pixels in the window of pixel m: the epilogue warps build it in shared
// memory as the K-major, 128B-swizzled A operand (fp16, scaled by a per-tile power of two so that tiny gradients stay
// normal numbers), the f halo tile arrives by TMA in slices of 8 halo rows and is consumed as the MN-major B operand.
// =====================================================================================================================
template <int R>
struct LbCfg {
  static constexpr int K = 2 * R + 1;
  static constexpr int HW = kLcTW + 2 * R, HH = kLcTH + 2 * R;
  static constexpr int NH = HW * HH;
  static constexpr int kSlices = (HH + 7) / 8;                     // TMA slices of 8 halo rows per 64-channel chunk
  static constexpr int kSliceK = 8 * HW;                           // K (halo pixels) per full slice, % 16 == 0
  static constexpr int kLastRows = HH - 8 * (kSlices - 1);
  static constexpr int kLastK = (kLastRows * HW + 15) / 16 * 16;   // K consumed from the last slice (padding rows are zero)
  static constexpr int KTOT = (kSlices - 1) * kSliceK + kLastK;
  static constexpr int kAChunks = (KTOT + 63) / 64;                // 64-element K chunks of the band operand
  static constexpr int kABytes = kAChunks * 128 * 128;
  static constexpr int kSliceBytes = kSliceK * 128;
  static constexpr int kSmemBytes = kABytes + kSlices * kSliceBytes + 1024 + 256;
};

struct LbParams {
  int N, H, W, chunks;
  int tiles_x, tiles_y;
  const __nv_bfloat16* dc; int ldc;          // band source: [n][h][w][ldc], channels [0, K*K)
  const __nv_bfloat16* add; int ld_add, add_off;   // optional direct gradient added to the result
  __nv_bfloat16* out;                        // [n][h][w][chunks * 64]
  float scale;
};

template <int R>
__global__ void __launch_bounds__(kLcThreads, 1)
local_corr_band_kernel(const __grid_constant__ CUtensorMap tm_full, const __grid_constant__ CUtensorMap tm_last, const LbParams p) {
  using Cfg = LbCfg<R>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* s_a = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_b = s_a + Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_b + Cfg::kSlices * Cfg::kSliceBytes);
  uint64_t* empty_bar = full_bar + Cfg::kSlices;
  uint64_t* tfull_bar = empty_bar + Cfg::kSlices;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                // [2]
  uint64_t* afull_bar = tempty_bar + 2;
  uint64_t* aempty_bar = afull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + 1);
  float* s_max = reinterpret_cast<float*>(tmem_slot + 1);          // [4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_x * p.tiles_y * p.N;

  // the slice slots are zero-filled once: rows the short last slice never writes must be finite (they meet zero band entries)
  for (int i = threadIdx.x; i < Cfg::kSlices * Cfg::kSliceBytes / 16; i += kLcThreads)
    reinterpret_cast<uint4*>(s_b)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_full);
    tma_prefetch_desc(&tm_last);
    for (int i = 0; i < Cfg::kSlices; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    mbar_init(afull_bar, 1);
    mbar_init(aempty_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int t = tile;
        const int tx = t % p.tiles_x;  t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int n = t / p.tiles_y;
        const int x0 = tx * kLcTW - R, y0 = ty * kLcTH - R;
        for (int kc = 0; kc < p.chunks; ++kc) {
#pragma unroll
          for (int s = 0; s < Cfg::kSlices; ++s) {          // slot == slice index: the ring advances one chunk at a time
            mbar_wait(&empty_bar[s], phase ^ 1);
            if (s + 1 < Cfg::kSlices) {
              mbar_expect_tx(&full_bar[s], Cfg::kSliceBytes);
              tma_load_4d(&tm_full, &full_bar[s], s_b + s * Cfg::kSliceBytes, kc * 64, x0, y0 + 8 * s, n);
            } else {
              mbar_expect_tx(&full_bar[s], Cfg::kLastRows * Cfg::HW * 128);
              tma_load_4d(&tm_last, &full_bar[s], s_b + s * Cfg::kSliceBytes, kc * 64, x0, y0 + 8 * s, n);
            }
          }
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(128, 64) | (1u << 16);       // A K-major (band), B MN-major (halo slice)
      const uint32_t a_base = smem_u32(s_a);
      uint32_t phase = 0;
      int it = 0, g = 0;                                                  // g: running chunk counter (accumulator buffers)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        mbar_wait(afull_bar, it & 1);                                     // band operand of this tile is in shared memory
        tc_fence_after();
        for (int kc = 0; kc < p.chunks; ++kc, ++g) {
          const int as = g & 1;
          mbar_wait(&tempty_bar[as], ((g >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * 64;
#pragma unroll
          for (int s = 0; s < Cfg::kSlices; ++s) {
            mbar_wait(&full_bar[s], phase);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(s_b + s * Cfg::kSliceBytes);
            const int ksteps = (s + 1 < Cfg::kSlices ? Cfg::kSliceK : Cfg::kLastK) / 16;
            for (int j = 0; j < ksteps; ++j) {
              const int koff = s * Cfg::kSliceK + 16 * j;                 // K offset (halo pixel index) of this step
              const uint64_t a_desc = umma_desc_sw128(a_base + (koff >> 6) * 16384, 1024) + 2 * ((koff & 63) >> 4);
              const uint64_t b_desc = umma_desc_sw128_mn(b_addr + j * 2048, Cfg::kSliceBytes, 1024);
              umma_f16(d_tmem, a_desc, b_desc, idesc, (s | j) != 0);
            }
            umma_commit(&empty_bar[s]);
          }
          phase ^= 1;
          umma_commit(&tfull_bar[as]);
        }
        umma_commit(aempty_bar);                                          // all MMAs reading this tile's band have retired
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int px = row & (kLcTW - 1), py = row >> 3;
    const int et = threadIdx.x - 64;
    const uint32_t s_a32 = smem_u32(s_a);
    int it = 0, g = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int t = tile;
      const int tx = t % p.tiles_x;  t /= p.tiles_x;
      const int ty = t % p.tiles_y;
      const int n = t / p.tiles_y;
      const int x = tx * kLcTW + px, y = ty * kLcTH + py;
      const bool valid = x < p.W && y < p.H;
      const size_t pix = (size_t)(n * p.H + y) * p.W + x;
      // ---- build the band operand -------------------------------------------------------------------------------
      mbar_wait(aempty_bar, (it & 1) ^ 1);                                // previous tile's MMAs are done with s_a
      for (int i = et; i < Cfg::kABytes / 16; i += 128) reinterpret_cast<uint4*>(s_a)[i] = make_uint4(0, 0, 0, 0);
      // this pixel's (2r+1)^2 gradients: 16-byte loads (ldc % 8 == 0, rows 16-byte aligned), kept packed in registers
      constexpr int kVec = (Cfg::K * Cfg::K + 7) / 8;
      uint4 raw[kVec];
      const uint4* src = reinterpret_cast<const uint4*>(p.dc + pix * p.ldc);
      float amax = 0.f;
      if (valid) {
#pragma unroll
        for (int i = 0; i < kVec; ++i) raw[i] = __ldg(src + i);
#pragma unroll
        for (int i = 0; i < Cfg::K * Cfg::K; ++i) {
          const __nv_bfloat16 e = reinterpret_cast<const __nv_bfloat16*>(raw)[i];
          amax = fmaxf(amax, fabsf(__bfloat162float(e)));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      if (lane == 0) s_max[q] = amax;
      asm volatile("bar.sync 1, 128;" ::: "memory");                      // zero fill + per-warp maxima visible
      amax = fmaxf(fmaxf(s_max[0], s_max[1]), fmaxf(s_max[2], s_max[3]));
      // power-of-two tile scale: the largest |gradient| lands in [2^13, 2^14)
      int e = 0;
      if (amax > 0.f) { frexpf(amax, &e); e = 14 - e; }
      e = e > 120 ? 120 : (e < -120 ? -120 : e);
      const float up = ldexpf(1.f, e), down = ldexpf(p.scale, -e);
      if (valid) {
#pragma unroll
        for (int a = 0; a < Cfg::K; ++a) {
#pragma unroll
          for (int b = 0; b < Cfg::K; ++b) {
            const float v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(raw)[a * Cfg::K + b]) * up;
            const int nn = (py + b) * Cfg::HW + px + a;                    // halo pixel index = K index
            sts_f16(s_a32 + (nn >> 6) * 16384 + row * 128 + ((((nn & 63) >> 3) ^ (row & 7)) << 4) + (nn & 7) * 2, v);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy writes -> visible to the MMA
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) mbar_arrive(afull_bar);
      // ---- per 64-channel chunk: accumulator -> (+ direct gradient) -> bf16 ----------------------------------------
      for (int kc = 0; kc < p.chunks; ++kc, ++g) {
        const int as = g & 1;
        mbar_wait(&tfull_bar[as], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 64;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          float v[32];
          tmem_ld32(t_addr + c0, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= down;
            if (p.add) {
              const uint4* ap = reinterpret_cast<const uint4*>(p.add + pix * p.ld_add + p.add_off + kc * 64 + c0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float f[8];
                unpack8_bf16(__ldg(ap + j), f);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[j * 8 + k] += f[k];
              }
            }
            uint4* op = reinterpret_cast<uint4*>(p.out + pix * (size_t)(p.chunks * 64) + kc * 64 + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) op[j] = pack8_bf16(v + j * 8);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

// dcorrT[p', a'*K + b'] = dcorr[p' + (b'-r, a'-r), (2r-a')*K + (2r-b')]  (0 when that pixel is outside the map): the gradient
// of the correlation volume re-indexed by the f2 pixel it touches.  One block per 8 x 16 pixel tile; the source halo tile is
// staged in shared memory with 16-byte loads, every thread then gathers 8 consecutive output channels and stores 16 bytes.
// Output rows are padded to kLcTPitch channels (16-byte aligned rows for the band kernel's vector loads).
constexpr int kLcTPitch = 128;
template <int R>
__global__ void __launch_bounds__(256)
corr_grad_transpose_kernel(const __nv_bfloat16* __restrict__ dc, int ldc, __nv_bfloat16* __restrict__ out, int H, int W,
                           int tiles_x, int tiles_y) {
  constexpr int K = 2 * R + 1, KK = K * K;
  constexpr int kVec = (KK + 7) / 8, KP = kVec * 8 + 8;          // staged row pitch (bf16), 16-byte multiple, bank-spread
  constexpr int HW = kLcTW + 2 * R, HH = kLcTH + 2 * R;
  extern __shared__ __align__(16) __nv_bfloat16 s_t[];           // [HH * HW][KP]
  int t = blockIdx.x;
  const int tx = t % tiles_x;  t /= tiles_x;
  const int ty = t % tiles_y;
  const int n = t / tiles_y;
  const int x0 = tx * kLcTW - R, y0 = ty * kLcTH - R;
  for (int i = threadIdx.x; i < HH * HW * kVec; i += 256) {
    const int hp = i / kVec, vi = i - hp * kVec;
    const int hy = hp / HW, hx = hp - hy * HW;
    const int y = y0 + hy, x = x0 + hx;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (x >= 0 && x < W && y >= 0 && y < H) v = __ldg(reinterpret_cast<const uint4*>(dc + ((size_t)(n * H + y) * W + x) * ldc) + vi);
    *reinterpret_cast<uint4*>(s_t + hp * KP + vi * 8) = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * kVec; i += 256) {
    const int m = i / kVec, vi = i - m * kVec;
    const int px = m & (kLcTW - 1), py = m >> 3;
    const int x = tx * kLcTW + px, y = ty * kLcTH + py;
    if (x >= W || y >= H) continue;
    __nv_bfloat16 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = vi * 8 + j;                                  // destination channel (a', b')
      const int a = ch / K, b = ch - a * K;
      o[j] = __float2bfloat16(0.f);
      if (ch < KK) o[j] = s_t[((py + b) * HW + (px + a)) * KP + (2 * R - a) * K + (2 * R - b)];
    }
    *reinterpret_cast<uint4*>(out + ((size_t)(n * H + y) * W + x) * kLcTPitch + vi * 8) = *reinterpret_cast<const uint4*>(o);
  }
}

template <int R>
static int launch_band(const void* f, const __nv_bfloat16* dc, int ldc, const __nv_bfloat16* add, int ld_add, int add_off,
                       void* out, int n, int h, int w, int c, cudaStream_t stream) {
  using Cfg = LbCfg<R>;
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(local_corr_band_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  CUtensorMap tf, tl;
  const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
  const uint64_t str[3] = {(uint64_t)c, (uint64_t)c * w, (uint64_t)c * w * h};
  const uint32_t box_f[4] = {64u, (uint32_t)Cfg::HW, 8u, 1u};
  const uint32_t box_l[4] = {64u, (uint32_t)Cfg::HW, (uint32_t)Cfg::kLastRows, 1u};
  int rc = make_tmap_2b(&tf, f, 4, dims, str, box_f, false);
  if (rc) return rc;
  rc = make_tmap_2b(&tl, f, 4, dims, str, box_l, false);
  if (rc) return rc;
  LbParams p{};
  p.N = n; p.H = h; p.W = w; p.chunks = c / 64;
  p.tiles_x = (w + kLcTW - 1) / kLcTW; p.tiles_y = (h + kLcTH - 1) / kLcTH;
  p.dc = dc; p.ldc = ldc; p.add = add; p.ld_add = ld_add; p.add_off = add_off;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.scale = 1.0f / sqrtf((float)c);
  const int tiles = p.tiles_x * p.tiles_y * n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  local_corr_band_kernel<R><<<grid, kLcThreads, Cfg::kSmemBytes, stream>>>(tf, tl, p);
  return check_cuda(cudaGetLastError(), "local_corr_band launch");
}

template <int R>
static int launch_corr_bwd_tc(const void* f1, const void* f2, const void* dq, int ld, int add_off, void* df1, void* df2, void* scratch,
                              int n, int h, int w, int c, cudaStream_t stream) {
  constexpr int K = 2 * R + 1, KK = K * K;
  const __nv_bfloat16* dqb = static_cast<const __nv_bfloat16*>(dq);
  // df1 = Band(dcorr) x f2 halo + direct gradient
  int rc = launch_band<R>(f2, dqb, ld, dqb, ld, add_off, df1, n, h, w, c, stream);
  if (rc) return rc;
  // df2 = Band(dcorrT) x f1 halo
  __nv_bfloat16* dct = static_cast<__nv_bfloat16*>(scratch);
  const int tiles_x = (w + kLcTW - 1) / kLcTW, tiles_y = (h + kLcTH - 1) / kLcTH;
  const size_t smem = (size_t)(kLcTW + 2 * R) * (kLcTH + 2 * R) * ((KK + 7) / 8 * 8 + 8) * 2;
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(corr_grad_transpose_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  corr_grad_transpose_kernel<R><<<tiles_x * tiles_y * n, 256, smem, stream>>>(dqb, ld, dct, h, w, tiles_x, tiles_y);
  RPNET_CUDA_OK(cudaGetLastError());
  return launch_band<R>(f1, dct, kLcTPitch, nullptr, 0, 0, df2, n, h, w, c, stream);
}

// Tensor-core path of rpnet_local_corr_bwd (dispatch in tail_kernels.cu).  scratch: bf16 [n*h*w*(2r+1)^2].  Returns 1 if the
// shape is not eligible.
int local_corr_bwd_tc(const void* f1, const void* f2, const void* dq, int ld, int add_off, void* df1, void* df2, void* scratch,
                      int n, int h, int w, int c, int radius, cudaStream_t stream) {
  if (c % 64 != 0 || !scratch) return 1;
  switch (radius) {
    case 3: return launch_corr_bwd_tc<3>(f1, f2, dq, ld, add_off, df1, df2, scratch, n, h, w, c, stream);
    case 5: return launch_corr_bwd_tc<5>(f1, f2, dq, ld, add_off, df1, df2, scratch, n, h, w, c, stream);
    default: return 1;
  }
}

}  // namespace rpnet
