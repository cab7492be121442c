// Weight gradient of the tap-list convolutions on the 5th-gen tensor cores (tcgen05.mma kind::f16, fp32
// accumulation in TMEM), reading BOTH operands straight from the NHWC tensors as MN-major UMMA tiles:
//
//   dW[tap][ci][co] = sum_{pixels p} X[p + tap][ci] * dZ[p][co]
//
// GEMM view: M = (tap, ci) rows — one M tile is two "X boxes" (tap, 64-channel chunk) = 128 rows;
//            N = co (BN per tile);  K = pixels (64 per stage: a bn x bh x bw box of the activation).
// A operand: X (forward activations as bf16, optional two-source channel concat), TMA box at the tap-shifted
//            pixel origin — out-of-bounds zero fill is the conv zero padding, exactly like the forward.
// B operand: dZ (bf16 gradient of the conv output), TMA box at the unshifted origin.
// A pixel row of 64 channels is one 128-byte swizzle row, so the TMA tile IS the canonical MN-major
// SWIZZLE_128B layout ((8,n),(8,k)):((1,LBO),(8,SBO)): LBO = box bytes (next 64-channel chunk),
// SBO = 1024 B (next 8 pixels).  Both operands must share one element format (a mixed fp16 x bf16 descriptor
// traps with an illegal instruction on B200), so the fp16 forward activations are converted to bf16 first.
//
// The K (pixel) range is split over CTAs (few output tiles, millions of pixels); each work item writes an
// fp32 partial tile, and wgrad_reduce_kernel sums the splits in a fixed order (deterministic) into the
// PyTorch-layout gradient [cout][cin][kh][kw] (+= so that a weight used by several calls accumulates).
//
// Replaces autograd's conv2d weight gradient for nn.Conv2d in net/modules.py:47-54,66-71 and
// net/rp_net.py:50-69 (the reference trains through torch autograd; it ships no backward code).
#include "common.cuh"

#include <cstdlib>

namespace rpnet {

#ifndef RPNET_WG_PIX
#define RPNET_WG_PIX 64
#endif
constexpr int kWgPix = RPNET_WG_PIX;               // pixels (K) per stage
constexpr int kWgBoxBytes = kWgPix * 128;          // one [64 ch x 64 px] box = 8 KB
constexpr int kWgThreads = 192;
constexpr int kWgMaxTaps = 9;
constexpr int kWgSmemBudget = 200 * 1024;

struct WgradParams {
  int N, H, W;
  int bw_log2, bh_log2;
  int tiles_x, tiles_y, tiles_n, ptiles;
  int chunks0, chunks1;
  int ntaps;
  int dy[kWgMaxTaps], dx[kWgMaxTaps];
  int n_boxes, n_mtiles, n_ntiles, splits;
  int rows;                                        // ntaps * cin
  int cout;
  int x_bf16, dz_bf16;
  float* partial;                                  // [splits][rows][cout]
};

constexpr int kWgAcc = 2;                          // 128-row accumulators per work item (they share every dZ stage)
constexpr int kWgBoxesPerItem = 2 * kWgAcc;        // M tile = 256 rows = 4 (tap, 64-channel chunk) boxes

// In-place fp16 -> bf16 conversion of an operand region in shared memory by the 128 epilogue threads (they idle during the
// K loop): the forward activations are fp16, the gradients bf16, and one tcgen05.mma takes a single operand format.
// Element-wise, so the 128-byte swizzle of the TMA tile is irrelevant.  et = 0..127.
__device__ __forceinline__ void cvt_smem_f16_to_bf16(uint8_t* base, int bytes, int et) {
  uint4* v = reinterpret_cast<uint4*>(base);
  for (int i = et; i < bytes / 16; i += 128) {
    const uint4 u = v[i];
    float f[8];
    unpack8_f16(u, f);
    v[i] = pack8_bf16(f);
  }
}

template <int BN>
struct WgradCfg {
  static constexpr int kABytes = kWgBoxesPerItem * kWgBoxBytes;
  static constexpr int kBBytes = (BN / 64) * kWgBoxBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kWgSmemBudget / kStageBytes) > 8 ? 8 : (kWgSmemBudget / kStageBytes);
  static constexpr int kTmemCols = kWgAcc * BN;     // single-buffered: the K loop of an item is long, its epilogue short
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 512;
};

template <int BN>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tm_x0, const __grid_constant__ CUtensorMap tm_x1,
                  const __grid_constant__ CUtensorMap tm_dz, const WgradParams p) {
  using Cfg = WgradCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* conv_bar = tempty_bar + 2;                   // [kStages] fp16 -> bf16 conversion of the x boxes done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(conv_bar + kStages);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // lane-0 broadcast: provably warp-uniform role branches
  const int lane = threadIdx.x & 31;
  const int tiles_mn = p.n_mtiles * p.n_ntiles;
  const int num_items = tiles_mn * p.splits;
  const int nch = p.chunks0 + p.chunks1;
  const int bw = 1 << p.bw_log2, bh = 1 << p.bh_log2;
  const int bn = kWgPix >> (p.bw_log2 + p.bh_log2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x0);
    tma_prefetch_desc(&tm_x1);
    tma_prefetch_desc(&tm_dz);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
      mbar_init(&conv_bar[i], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int split = item / tiles_mn;
        const int rem = item % tiles_mn;
        const int mt = rem / p.n_ntiles, nt = rem % p.n_ntiles;
        const int k0 = (int)((long long)split * p.ptiles / p.splits);
        const int k1 = (int)((long long)(split + 1) * p.ptiles / p.splits);
        // everything that does not depend on the pixel tile is resolved once per item: the producer is ONE thread, and the per-stage
        // divisions / parameter loads of its inner loop were what bounded the kernel (tensor pipe 60 % active)
        const CUtensorMap* box_tm[kWgBoxesPerItem];
        int box_c[kWgBoxesPerItem], box_dx[kWgBoxesPerItem], box_dy[kWgBoxesPerItem];
#pragma unroll
        for (int j = 0; j < kWgBoxesPerItem; ++j) {                 // ragged box count: duplicate the first box (rows discarded)
          const int b = (kWgBoxesPerItem * mt + j < p.n_boxes) ? kWgBoxesPerItem * mt + j : kWgBoxesPerItem * mt;
          const int tap = b / nch, c = b % nch;
          box_tm[j] = c < p.chunks0 ? &tm_x0 : &tm_x1;
          box_c[j] = (c < p.chunks0 ? c : c - p.chunks0) * 64;
          box_dx[j] = p.dx[tap];
          box_dy[j] = p.dy[tap];
        }
        int tx = k0 % p.tiles_x, ty = (k0 / p.tiles_x) % p.tiles_y, tn = k0 / (p.tiles_x * p.tiles_y);
        for (int pt = k0; pt < k1; ++pt) {
          const int x0 = tx * bw, y0 = ty * bh, n0 = tn * bn;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = tiles + stage * Cfg::kStageBytes;
          uint8_t* b_dst = a_dst + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
          for (int j = 0; j < kWgBoxesPerItem; ++j)
            tma_load_4d(box_tm[j], &full_bar[stage], a_dst + j * kWgBoxBytes, box_c[j], x0 + box_dx[j], y0 + box_dy[j], n0);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(&tm_dz, &full_bar[stage], b_dst + j * kWgBoxBytes, nt * BN + j * 64, x0, y0, n0);
          if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++tn; } }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp runs the loop (uniform operands), one elected lane issues =====================
    {
      const uint32_t idesc = umma_idesc_f16(128, BN) | (1u << 15) | (1u << 16) | (1u << 7) | (1u << 10);   // both operands bf16 at MMA time
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int split = item / tiles_mn;
        const int k0 = (int)((long long)split * p.ptiles / p.splits);
        const int k1 = (int)((long long)(split + 1) * p.ptiles / p.splits);
        const uint32_t aphase = it & 1;
        mbar_wait(&tempty_bar[0], aphase ^ 1);            // the epilogue has drained the accumulators of the previous item
        tc_fence_after();
        for (int pt = k0; pt < k1; ++pt) {
          mbar_wait(p.x_bf16 ? &full_bar[stage] : &conv_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kWgPix / 16; ++k) {
            // 16 pixels (K) = 16 swizzle rows = 2048 B further into every box
            const uint64_t b_desc = umma_desc_sw128_mn(b_addr + k * 2048, kWgBoxBytes, 1024);
#pragma unroll
            for (int h = 0; h < kWgAcc; ++h) {
              const uint64_t a_desc = umma_desc_sw128_mn(a_addr + h * 2 * kWgBoxBytes + k * 2048, kWgBoxBytes, 1024);
              if (elect_one()) umma_f16(tmem_base + h * BN, a_desc, b_desc, idesc, (pt != k0 || k != 0) ? 1u : 0u);
            }
          }
          if (elect_one()) umma_commit(&empty_bar[stage]);
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(&tfull_bar[0]);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps (2..5): TMEM -> fp32 partial tile =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int it = 0;
    int cstage = 0;
    uint32_t cphase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int split = item / tiles_mn;
      const int rem = item % tiles_mn;
      const int mt = rem / p.n_ntiles, nt = rem % p.n_ntiles;
      const uint32_t aphase = it & 1;
      if (!p.x_bf16) {                                   // K loop: convert the x boxes of every stage as they land
        const int k0 = (int)((long long)split * p.ptiles / p.splits);
        const int k1 = (int)((long long)(split + 1) * p.ptiles / p.splits);
        for (int pt = k0; pt < k1; ++pt) {
          mbar_wait(&full_bar[cstage], cphase);
          cvt_smem_f16_to_bf16(tiles + cstage * Cfg::kStageBytes, Cfg::kABytes, threadIdx.x - 64);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&conv_bar[cstage]);
          if (++cstage == kStages) { cstage = 0; cphase ^= 1; }
        }
      }
      mbar_wait(&tfull_bar[0], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < kWgAcc; ++h) {
        const int r = (mt * kWgAcc + h) * 128 + row;
        const bool valid = r < p.rows;
        float* dst = p.partial + ((size_t)split * p.rows + (valid ? r : 0)) * p.cout + nt * BN;
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + h * BN;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          float v[32];
          tmem_ld32(t_addr + c0, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[0]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------
// Halo-sharing variant for 3x3 (dilation 1) convs with few output channels (cout <= 128: the full-resolution layers,
// where the generic kernel is operand-feed bound: every tap is its own 8 KB box and dZ is re-read per M tile).
// Work item = (64-channel input chunk, 64-wide cout tile, pixel split).  Per 8 x 8 pixel tile (K = 64) the stage holds
//   three x boxes [64 ch x 8 px x 10 rows], one per column offset dx (shifted by TMA, out-of-bounds = zero padding), and
//   one dZ box [64 cout x 8 x 8];
// the row offset dy of a tap is a 1024-byte (8-pixel) shift of the MN-major descriptor inside its dx box, so all nine
// taps are served by 30 KB instead of 72 KB, and all nine share the one dZ box.  M = 9 taps x 64 channels = five 128-row
// accumulators (taps in (dx, dy) order so that the second 64-row chunk of a descriptor lies at a positive offset; the last
// accumulator holds tap 8 twice, the duplicate rows are dropped).
// ---------------------------------------------------------------------------------------------------
constexpr int kWhBoxA = 64 * 2 * 8 * 10;            // 10240 B: one dx box (80 pixel rows of 128 B)
constexpr int kWhBoxB = 64 * 2 * 64;                // 8192 B
constexpr int kWhStageBytes = 3 * kWhBoxA + kWhBoxB;  // 38912 B (multiple of 1024)
constexpr int kWhStages = 5;
constexpr int kWhSmemBytes = kWhStages * kWhStageBytes + 1024 + 512;

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tm_x0, const __grid_constant__ CUtensorMap tm_x1,
                       const __grid_constant__ CUtensorMap tm_dz, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + kWhStages * kWhStageBytes);
  uint64_t* empty_bar = full_bar + kWhStages;
  uint64_t* tfull_bar = empty_bar + kWhStages;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint64_t* conv_bar = tempty_bar + 1;                   // [kWhStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(conv_bar + kWhStages);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform role branches
  const int nch = p.chunks0 + p.chunks1;
  const int tiles_mn = nch * p.n_ntiles;                // (input chunk, cout tile) pairs
  const int num_items = tiles_mn * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x0);
    tma_prefetch_desc(&tm_x1);
    tma_prefetch_desc(&tm_dz);
    for (int i = 0; i < kWhStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); mbar_init(&conv_bar[i], 4); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);
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
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int split = item / tiles_mn;
        const int rem = item % tiles_mn;
        const int kc = rem / p.n_ntiles, nt = rem % p.n_ntiles;
        const int k0 = (int)((long long)split * p.ptiles / p.splits);
        const int k1 = (int)((long long)(split + 1) * p.ptiles / p.splits);
        const CUtensorMap* tmx = kc < p.chunks0 ? &tm_x0 : &tm_x1;
        const int cx = (kc < p.chunks0 ? kc : kc - p.chunks0) * 64;
        int tx = k0 % p.tiles_x, ty = (k0 / p.tiles_x) % p.tiles_y, tn = k0 / (p.tiles_x * p.tiles_y);   // no division per stage
        for (int pt = k0; pt < k1; ++pt) {
          const int x0 = tx * 8, y0 = ty * 8;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = tiles + stage * kWhStageBytes;
          mbar_expect_tx(&full_bar[stage], kWhStageBytes);
#pragma unroll
          for (int d = 0; d < 3; ++d) tma_load_4d(tmx, &full_bar[stage], a_dst + d * kWhBoxA, cx, x0 + d - 1, y0 - 1, tn);
          tma_load_4d(&tm_dz, &full_bar[stage], a_dst + 3 * kWhBoxA, nt * 64, x0, y0, tn);
          if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++tn; } }
          if (++stage == kWhStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {   // the whole warp runs the loop (uniform operands), one elected lane issues
      const uint32_t idesc = umma_idesc_f16(128, 64) | (1u << 15) | (1u << 16) | (1u << 7) | (1u << 10);   // MN-major, bf16
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int split = item / tiles_mn;
        const int k0 = (int)((long long)split * p.ptiles / p.splits);
        const int k1 = (int)((long long)(split + 1) * p.ptiles / p.splits);
        mbar_wait(tempty_bar, (it & 1) ^ 1);
        tc_fence_after();
        for (int pt = k0; pt < k1; ++pt) {
          mbar_wait(p.x_bf16 ? &full_bar[stage] : &conv_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + stage * kWhStageBytes);
          const uint32_t b_addr = a_addr + 3 * kWhBoxA;
#pragma unroll
          for (int k = 0; k < 4; ++k) {                                   // 16 pixels (two image rows of the tile) per MMA
            const uint64_t b_desc = umma_desc_sw128_mn(b_addr + k * 2048, kWhBoxB, 1024);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              // sorted tap s = dxi * 3 + dyi lives at dxi * kWhBoxA + dyi * 1024 (row offset dy = one 8-pixel group)
              const int sa = 2 * j, sb = 2 * j + 1 < 9 ? 2 * j + 1 : 2 * j;
              const uint32_t off_a = (sa / 3) * kWhBoxA + (sa % 3) * 1024, off_b = (sb / 3) * kWhBoxA + (sb % 3) * 1024;
              const uint64_t a_desc = umma_desc_sw128_mn(a_addr + off_a + k * 2048, off_b - off_a, 1024);
              if (elect_one()) umma_f16(tmem_base + j * 64, a_desc, b_desc, idesc, (pt != k0 || k != 0) ? 1u : 0u);
            }
          }
          if (elect_one()) umma_commit(&empty_bar[stage]);
          __syncwarp();
          if (++stage == kWhStages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(tfull_bar);
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int cin = nch * 64;
    int it = 0;
    int cstage = 0;
    uint32_t cphase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int split = item / tiles_mn;
      const int rem = item % tiles_mn;
      const int kc = rem / p.n_ntiles, nt = rem % p.n_ntiles;
      if (!p.x_bf16) {                                   // K loop: convert the three x boxes of every stage as they land
        const int k0 = (int)((long long)split * p.ptiles / p.splits);
        const int k1 = (int)((long long)(split + 1) * p.ptiles / p.splits);
        for (int pt = k0; pt < k1; ++pt) {
          mbar_wait(&full_bar[cstage], cphase);
          cvt_smem_f16_to_bf16(tiles + cstage * kWhStageBytes, 3 * kWhBoxA, threadIdx.x - 64);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&conv_bar[cstage]);
          if (++cstage == kWhStages) { cstage = 0; cphase ^= 1; }
        }
      }
      mbar_wait(tfull_bar, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 5; ++j) {
        const int s = 2 * j + (row >> 6);                               // sorted tap of this accumulator row
        const bool valid = s < 9;
        const int tap = valid ? (s % 3) * 3 + s / 3 : 0;                // natural tap index (dy + 1) * 3 + (dx + 1)
        float* dst = p.partial + ((size_t)split * p.rows + (size_t)tap * cin + kc * 64 + (row & 63)) * p.cout + nt * 64;
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j * 64;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          float v[32];
          tmem_ld32(t_addr + c0, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// grad[co][ci_real][tap] (beta * old +) = sum_split partial[split][tap * cin + ci][co].
// Packed input channels [hole_start, hole_start + hole_len) are padding (skipped); later channels shift down.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grad, int splits, int ntaps, int cin, int cout,
                    int hole_start, int hole_len, int accumulate) {
  const long long total = (long long)ntaps * cin * cout;
  const int cin_real = cin - hole_len;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    const int r = (int)(i / cout);
    const int ci = r % cin, tap = r / cin;
    if (ci >= hole_start && ci < hole_start + hole_len) continue;
    const int cr = ci < hole_start ? ci : ci - hole_len;
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += __ldg(partial + (size_t)sp * total + i);
    float* g = grad + ((size_t)co * cin_real + cr) * ntaps + tap;
    *g = accumulate ? (*g + s) : s;
  }
}

template <int BN>
static int launch_wgrad(const CUtensorMap& tx0, const CUtensorMap& tx1, const CUtensorMap& tdz, const WgradParams& p, int grid,
                        cudaStream_t stream) {
  using Cfg = WgradCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  conv_wgrad_kernel<BN><<<grid, kWgThreads, Cfg::kSmemBytes, stream>>>(tx0, tx1, tdz, p);
  return check_cuda(cudaGetLastError(), "conv_wgrad_kernel launch");
}

struct WgradPlan {
  int BN, bw, bh, bn, tiles_x, tiles_y, tiles_n, ptiles, n_boxes, n_mtiles, n_ntiles, splits, rows;
};

static WgradPlan plan_wgrad(int c0, int c1, int n, int h, int w, int ntaps, int cout) {
  WgradPlan pl;
  pl.BN = (cout % 256 == 0) ? 256 : (cout % 128 == 0 ? 128 : 64);
  pl.bw = pow2_floor(w < 16 ? w : 16);
  pl.bh = pow2_floor(h < kWgPix / pl.bw ? h : kWgPix / pl.bw);
  pl.bn = kWgPix / (pl.bw * pl.bh);
  pl.tiles_x = (w + pl.bw - 1) / pl.bw;
  pl.tiles_y = (h + pl.bh - 1) / pl.bh;
  pl.tiles_n = (n + pl.bn - 1) / pl.bn;
  pl.ptiles = pl.tiles_x * pl.tiles_y * pl.tiles_n;
  const int cin = c0 + c1;
  pl.rows = ntaps * cin;
  pl.n_boxes = ntaps * (cin / 64);
  pl.n_mtiles = (pl.n_boxes + kWgBoxesPerItem - 1) / kWgBoxesPerItem;
  pl.n_ntiles = cout / pl.BN;
  const int tiles_mn = pl.n_mtiles * pl.n_ntiles;
  // Split K (pixels) over CTAs.  Work items run in waves of one per SM, so the cost of a choice is
  //   waves(s) * ceil(ptiles / s) stages  +  the fp32 partial tiles every split writes and the reduce kernel re-reads;
  // pick the split count that minimises it (a fixed "2 waves" rule left up to a third of the SMs idle in the last wave).
  const int sms = num_sms();
  const int max_splits = pl.ptiles / 8 > 0 ? pl.ptiles / 8 : 1;     // >= 8 pixel tiles (512 px) per item
  const double stage_cyc = kWgAcc * (kWgPix / 16.0) * (pl.BN / 2.0) / 0.65;     // kWgPix / 16 MMAs of K=16 per accumulator and stage
  const double partial_cyc = (double)pl.rows * cout * 8.0 / 2100.0; // write + re-read of one split's partial tile, chip-wide
  int splits = 1;
  double best = 1e300;
  for (int sp = 1; sp <= max_splits && sp <= 4096; ++sp) {
    const long long items = (long long)tiles_mn * sp;
    const long long waves = (items + sms - 1) / sms;
    const double cost = (double)waves * ((pl.ptiles + sp - 1) / sp) * stage_cyc + sp * partial_cyc;
    if (cost < best) { best = cost; splits = sp; }
  }
  pl.splits = splits;
  return pl;
}

// Plan of the halo-sharing variant (8 x 8 pixel tiles, items = (input chunk, 64-wide cout tile, split)).
static bool halo_eligible(int n, int h, int w, int ntaps, int cout) {
  return ntaps == 9 && cout <= 128 && h >= 8 && w >= 8 && n > 0;
}
static WgradPlan plan_wgrad_halo(int c0, int c1, int n, int h, int w, int cout) {
  WgradPlan pl{};
  pl.BN = 64; pl.bw = 8; pl.bh = 8; pl.bn = 1;
  pl.tiles_x = (w + 7) / 8; pl.tiles_y = (h + 7) / 8; pl.tiles_n = n;
  pl.ptiles = pl.tiles_x * pl.tiles_y * pl.tiles_n;
  const int cin = c0 + c1;
  pl.rows = 9 * cin;
  pl.n_boxes = 9 * (cin / 64);
  pl.n_mtiles = cin / 64;
  pl.n_ntiles = cout / 64;
  const int tiles_mn = pl.n_mtiles * pl.n_ntiles;
  const int sms = num_sms();
  const int max_splits = pl.ptiles / 8 > 0 ? pl.ptiles / 8 : 1;
  const double stage_cyc = 5 * 4.0 * 32.0 / 0.65;
  const double partial_cyc = (double)pl.rows * cout * 8.0 / 2100.0 / (tiles_mn > 0 ? tiles_mn : 1) * tiles_mn;
  int splits = 1;
  double best = 1e300;
  for (int sp = 1; sp <= max_splits && sp <= 4096; ++sp) {
    const long long items = (long long)tiles_mn * sp;
    const long long waves = (items + sms - 1) / sms;
    const double cost = (double)waves * ((pl.ptiles + sp - 1) / sp) * stage_cyc + sp * partial_cyc;
    if (cost < best) { best = cost; splits = sp; }
  }
  pl.splits = splits;
  return pl;
}
static bool taps_are_3x3(int ntaps, const int* dy, const int* dx) {
  if (ntaps != 9) return false;
  for (int t = 0; t < 9; ++t)
    if (dy[t] != t / 3 - 1 || dx[t] != t % 3 - 1) return false;
  return true;
}

}  // namespace rpnet

using namespace rpnet;

RPNET_API long long rpnet_conv_wgrad_workspace_bytes(int c0, int c1, int n, int h, int w, int ntaps, int cout) {
  if (c0 <= 0 || c0 % 64 || c1 < 0 || c1 % 64 || n <= 0 || h <= 0 || w <= 0 || ntaps < 1 || ntaps > kWgMaxTaps || cout < 64 ||
      cout % 64) {
    set_error("conv_wgrad_workspace_bytes: bad shape");
    return RPNET_ERR_ARG;
  }
  const WgradPlan pl = plan_wgrad(c0, c1, n, h, w, ntaps, cout);
  long long bytes = (long long)pl.splits * pl.rows * cout * 4;
  if (halo_eligible(n, h, w, ntaps, cout)) {                 // the 3x3 halo-sharing variant may be chosen at call time
    const WgradPlan ph = plan_wgrad_halo(c0, c1, n, h, w, cout);
    const long long b2 = (long long)ph.splits * ph.rows * cout * 4;
    if (b2 > bytes) bytes = b2;
  }
  return bytes;
}

// dz_strides (optional): pixel strides {x, y, image} in elements of a strided view of the gradient tensor (the parity phases
// of a 2x up-sampled map); null = dense [n][h][w][cout].
static int wgrad_impl(const void* x0, int c0, const void* x1, int c1, int x_bf16, const void* dz_bf16, const long long* dz_strides,
                      int n, int h, int w, int ntaps, const int* tap_dy, const int* tap_dx, int cout, float* grad, int hole_start,
                      int hole_len, int accumulate, void* workspace, long long workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(x0 && dz_bf16 && grad && workspace, "conv_wgrad: null pointer argument");
  RPNET_REQUIRE(c0 > 0 && c0 % 64 == 0 && c1 >= 0 && c1 % 64 == 0, "conv_wgrad: channel counts must be multiples of 64 (got %d, %d)", c0, c1);
  RPNET_REQUIRE(c1 == 0 || x1, "conv_wgrad: x1 is null but c1 = %d", c1);
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0, "conv_wgrad: bad grid %d x %d x %d", n, h, w);
  RPNET_REQUIRE(ntaps >= 1 && ntaps <= kWgMaxTaps, "conv_wgrad: ntaps %d out of range [1, %d]", ntaps, kWgMaxTaps);
  RPNET_REQUIRE(cout >= 64 && cout % 64 == 0, "conv_wgrad: cout must be a multiple of 64 (got %d)", cout);
  RPNET_REQUIRE(hole_len >= 0 && hole_start >= 0 && hole_start + hole_len <= c0 + c1, "conv_wgrad: bad padding hole [%d, +%d)", hole_start, hole_len);
  if (!dz_strides && halo_eligible(n, h, w, ntaps, cout) && taps_are_3x3(ntaps, tap_dy, tap_dx) && !getenv("RPNET_WGRAD_NO_HALO")) {
    const WgradPlan ph = plan_wgrad_halo(c0, c1, n, h, w, cout);
    const long long need_h = (long long)ph.splits * ph.rows * cout * 4;
    RPNET_REQUIRE(workspace_bytes >= need_h, "conv_wgrad: workspace too small (%lld < %lld bytes)", workspace_bytes, need_h);
    WgradParams p{};
    p.N = n; p.H = h; p.W = w;
    p.tiles_x = ph.tiles_x; p.tiles_y = ph.tiles_y; p.tiles_n = ph.tiles_n; p.ptiles = ph.ptiles;
    p.chunks0 = c0 / 64; p.chunks1 = c1 / 64; p.ntaps = 9;
    p.n_boxes = ph.n_boxes; p.n_mtiles = ph.n_mtiles; p.n_ntiles = ph.n_ntiles; p.splits = ph.splits;
    p.rows = ph.rows; p.cout = cout; p.x_bf16 = x_bf16 ? 1 : 0; p.dz_bf16 = 1;
    p.partial = static_cast<float*>(workspace);
    CUtensorMap tx0, tx1, tdz;
    const uint32_t box_a[4] = {64u, 8u, 10u, 1u}, box_b[4] = {64u, 8u, 8u, 1u};
    {
      const uint64_t dims[4] = {(uint64_t)c0, (uint64_t)w, (uint64_t)h, (uint64_t)n};
      const uint64_t str[3] = {(uint64_t)c0, (uint64_t)c0 * w, (uint64_t)c0 * w * h};
      int rc = make_tmap_2b(&tx0, x0, 4, dims, str, box_a, x_bf16 != 0);
      if (rc) return rc;
    }
    if (c1 > 0) {
      const uint64_t dims[4] = {(uint64_t)c1, (uint64_t)w, (uint64_t)h, (uint64_t)n};
      const uint64_t str[3] = {(uint64_t)c1, (uint64_t)c1 * w, (uint64_t)c1 * w * h};
      int rc = make_tmap_2b(&tx1, x1, 4, dims, str, box_a, x_bf16 != 0);
      if (rc) return rc;
    } else {
      tx1 = tx0;
    }
    {
      const uint64_t dims[4] = {(uint64_t)cout, (uint64_t)w, (uint64_t)h, (uint64_t)n};
      const uint64_t str[3] = {(uint64_t)cout, (uint64_t)cout * w, (uint64_t)cout * w * h};
      int rc = make_tmap_2b(&tdz, dz_bf16, 4, dims, str, box_b, true);
      if (rc) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
      RPNET_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWhSmemBytes));
      attr_set = true;
    }
    const int items = ph.n_mtiles * ph.n_ntiles * ph.splits;
    const int grid = items < num_sms() ? items : num_sms();
    conv_wgrad_halo_kernel<<<grid, kWgThreads, kWhSmemBytes, stream>>>(tx0, tx1, tdz, p);
    RPNET_CUDA_OK(cudaGetLastError());
    const long long total = (long long)ph.rows * cout;
    long long g = (total + 255) / 256;
    if (g > 148LL * 16) g = 148LL * 16;
    wgrad_reduce_kernel<<<(int)g, 256, 0, stream>>>(p.partial, grad, ph.splits, 9, c0 + c1, cout, hole_start, hole_len, accumulate);
    return check_cuda(cudaGetLastError(), "wgrad_reduce_kernel launch");
  }
  const WgradPlan pl = plan_wgrad(c0, c1, n, h, w, ntaps, cout);
  const long long need = (long long)pl.splits * pl.rows * cout * 4;
  RPNET_REQUIRE(workspace_bytes >= need, "conv_wgrad: workspace too small (%lld < %lld bytes)", workspace_bytes, need);

  WgradParams p{};
  p.N = n; p.H = h; p.W = w;
  p.bw_log2 = ilog2(pl.bw); p.bh_log2 = ilog2(pl.bh);
  p.tiles_x = pl.tiles_x; p.tiles_y = pl.tiles_y; p.tiles_n = pl.tiles_n; p.ptiles = pl.ptiles;
  p.chunks0 = c0 / 64; p.chunks1 = c1 / 64;
  p.ntaps = ntaps;
  for (int i = 0; i < ntaps; ++i) { p.dy[i] = tap_dy[i]; p.dx[i] = tap_dx[i]; }
  p.n_boxes = pl.n_boxes; p.n_mtiles = pl.n_mtiles; p.n_ntiles = pl.n_ntiles; p.splits = pl.splits;
  p.rows = pl.rows; p.cout = cout;
  p.x_bf16 = x_bf16 ? 1 : 0; p.dz_bf16 = 1;
  p.partial = static_cast<float*>(workspace);

  CUtensorMap tx0, tx1, tdz;
  const uint32_t box[4] = {64u, (uint32_t)pl.bw, (uint32_t)pl.bh, (uint32_t)pl.bn};
  {
    const uint64_t dims[4] = {(uint64_t)c0, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t str[3] = {(uint64_t)c0, (uint64_t)c0 * w, (uint64_t)c0 * w * h};
    int rc = make_tmap_2b(&tx0, x0, 4, dims, str, box, x_bf16 != 0);
    if (rc) return rc;
  }
  if (c1 > 0) {
    const uint64_t dims[4] = {(uint64_t)c1, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t str[3] = {(uint64_t)c1, (uint64_t)c1 * w, (uint64_t)c1 * w * h};
    int rc = make_tmap_2b(&tx1, x1, 4, dims, str, box, x_bf16 != 0);
    if (rc) return rc;
  } else {
    tx1 = tx0;
  }
  {
    const uint64_t dims[4] = {(uint64_t)cout, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    uint64_t str[3] = {(uint64_t)cout, (uint64_t)cout * w, (uint64_t)cout * w * h};
    if (dz_strides) { str[0] = (uint64_t)dz_strides[0]; str[1] = (uint64_t)dz_strides[1]; str[2] = (uint64_t)dz_strides[2]; }
    int rc = make_tmap_2b(&tdz, dz_bf16, 4, dims, str, box, true);
    if (rc) return rc;
  }
  const int items = pl.n_mtiles * pl.n_ntiles * pl.splits;
  const int grid = items < num_sms() ? items : num_sms();
  int rc;
  switch (pl.BN) {
    case 256: rc = launch_wgrad<256>(tx0, tx1, tdz, p, grid, stream); break;
    case 128: rc = launch_wgrad<128>(tx0, tx1, tdz, p, grid, stream); break;
    default:  rc = launch_wgrad<64>(tx0, tx1, tdz, p, grid, stream); break;
  }
  if (rc) return rc;
  const long long total = (long long)pl.rows * cout;
  long long g = (total + 255) / 256;
  if (g > 148LL * 16) g = 148LL * 16;
  wgrad_reduce_kernel<<<(int)g, 256, 0, stream>>>(p.partial, grad, pl.splits, ntaps, c0 + c1, cout, hole_start, hole_len,
                                                   accumulate);
  return check_cuda(cudaGetLastError(), "wgrad_reduce_kernel launch");
}

RPNET_API int rpnet_conv_wgrad(const void* x0, int c0, const void* x1, int c1, int x_bf16, const void* dz_bf16, int n, int h,
                               int w, int ntaps, const int* tap_dy, const int* tap_dx, int cout, float* grad, int hole_start,
                               int hole_len, int accumulate, void* workspace, long long workspace_bytes, void* stream_) {
  return wgrad_impl(x0, c0, x1, c1, x_bf16, dz_bf16, nullptr, n, h, w, ntaps, tap_dy, tap_dx, cout, grad, hole_start, hole_len, accumulate,
                    workspace, workspace_bytes, stream_);
}

// ---------------------------------------------------------------------------------------------------
// up_conv (nn.Upsample(x2, nearest) + 3x3 conv, net/modules.py:61-75) in sub-pixel form: output parity phase (py, px) is a
// 2x2 conv of the LOW-resolution input with row/column-summed weights (2.25x fewer MACs than the materialised form).
//   rows(py = 0): tap 0 = offset -1 <- ky {0},   tap 1 = offset 0 <- ky {1, 2}
//   rows(py = 1): tap 0 = offset  0 <- ky {0, 1}, tap 1 = offset +1 <- ky {2}          (same for columns)
// Weight gradient: four 4-tap weight-gradient GEMMs (x = the low-resolution input, dZ = the phase's strided view of the
// high-resolution gradient), then dW[ky][kx] = sum over the four phases of the phase tap that contains (ky, kx).
// ---------------------------------------------------------------------------------------------------
namespace rpnet {
__device__ __host__ inline int upconv_tap_of(int parity, int k) {      // which of the phase's two taps holds kernel row/col k
  return parity == 0 ? (k == 0 ? 0 : 1) : (k == 2 ? 1 : 0);
}
__global__ void upconv_wgrad_combine_kernel(const float* __restrict__ dwp /*[4 phases][cout][cin][4 taps]*/, float* __restrict__ grad,
                                            long long cc /*cout * cin*/, int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cc * 9; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % 9);
    const long long e = i / 9;
    const int ky = k / 3, kx = k % 3;
    float s = 0.f;
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
      const int py = ph >> 1, px = ph & 1;
      s += __ldg(dwp + ((size_t)ph * cc + e) * 4 + upconv_tap_of(py, ky) * 2 + upconv_tap_of(px, kx));
    }
    grad[i] = accumulate ? grad[i] + s : s;
  }
}
}  // namespace rpnet

RPNET_API long long rpnet_upconv_wgrad_workspace_bytes(int cin, int n, int h, int w, int cout) {
  const long long part = rpnet_conv_wgrad_workspace_bytes(cin, 0, n, h, w, 4, cout);
  if (part < 0) return part;
  return part + 4LL * cout * cin * 4 * 4;
}

RPNET_API int rpnet_upconv_wgrad(const void* x_low, int x_bf16, const void* dz_bf16, int n, int h, int w, int cin, int cout, float* grad,
                                 int accumulate, void* workspace, long long workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RPNET_REQUIRE(x_low && dz_bf16 && grad && workspace, "upconv_wgrad: null pointer argument");
  const long long part = rpnet_conv_wgrad_workspace_bytes(cin, 0, n, h, w, 4, cout);
  RPNET_REQUIRE(part >= 0 && workspace_bytes >= part + 4LL * cout * cin * 16, "upconv_wgrad: workspace too small");
  float* dwp = reinterpret_cast<float*>(static_cast<char*>(workspace) + part);
  const long long strides[3] = {2LL * cout, 2LL * (2 * w) * cout, (long long)(2 * h) * (2 * w) * cout};
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    int dy[4], dx[4];
    for (int t = 0; t < 4; ++t) {
      dy[t] = (py == 0 ? -1 : 0) + (t >> 1);
      dx[t] = (px == 0 ? -1 : 0) + (t & 1);
    }
    const __nv_bfloat16* dzp = static_cast<const __nv_bfloat16*>(dz_bf16) + ((size_t)py * (2 * w) + px) * cout;
    int rc = wgrad_impl(x_low, cin, nullptr, 0, x_bf16, dzp, strides, n, h, w, 4, dy, dx, cout, dwp + (size_t)ph * cout * cin * 4, 0, 0, 0,
                        workspace, part, stream_);
    if (rc) return rc;
  }
  const long long cc = (long long)cout * cin;
  long long g = (cc * 9 + 255) / 256;
  if (g > 148LL * 16) g = 148LL * 16;
  upconv_wgrad_combine_kernel<<<(int)g, 256, 0, stream>>>(dwp, grad, cc, accumulate);
  return check_cuda(cudaGetLastError(), "upconv_wgrad_combine launch");
}
