// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, fp32 accumulation in TMEM) for the RP-Net conv stacks: every
// 3x3 / dilated 3x3 / 1x1 / sub-pixel 2x2 "tap list" convolution of the U-Net / VGG / ResNet18 encoders and of the context-relation
// encoder, forward and data gradient.
//
//   reference ops replaced: nn.Conv2d + nn.BatchNorm2d(eval) + nn.ReLU (+ nn.MaxPool2d(2,2), + torch.cat on channels,
//   + nn.Upsample(x2) via sub-pixel taps, + the BasicBlock residual add), or in train mode the conv + the batch statistics
//   net/modules.py:42-75, net/unet.py:435-467, net/vgg.py:22-58, net/rp_net.py:19-42,50-69.
//
// GEMM view: M = output pixels (128 per tile: a bn x bh x bw box of the NHWC activation; 256 for a CTA pair),
//            N = output channels (BN per tile), K = taps x input channels (one 128-byte swizzle row per k-block).
// A operand: NHWC activations; one TMA box load per (tap, 64-channel chunk) at the tap-shifted pixel origin - TMA out-of-bounds
//            zero fill IS the conv zero padding.  B operand: weights packed [tap][cout][cin] (K-major); one TMA box per k-block.
// Both land in 128B-swizzled K-major shared-memory tiles consumed directly by tcgen05.mma.
// Arithmetic (include/rpnet_b200.h): plain fp16 / bf16 operands (kind::f16); split-fp16 hi.Wh + lo.Wh + hi.Wl as three k-block
// segments; or - the default - the fp16 main term plus two e4m3 first-order corrections (kind::f8f6f4, K = 128 per k-block) into the
// SAME accumulator: the e4m3 k-blocks of a tile go first and the first kind::f16 MMA scales the accumulator by 2^-15 (scale-input-d).
// Warp roles (192 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (warp-uniform loop,
// one elected lane issues), warps 2..5 = epilogue (TMEM -> registers, next chunk prefetched -> scale/shift/(+residual)/ReLU or the
// BatchNorm statistics [-> 2x2 max-pool via warp shuffles] -> hi / lo plane stores).  Two TMEM accumulator buffers overlap the
// epilogue of tile i with the main loop of tile i+1; a buffer is handed back as soon as its last chunk sits in registers.
// Variants: CTA pair (cta_group::2), weights-stationary, halo-sharing (one activation box per column offset), both together.
#include "common.cuh"

#include <mutex>
#include <type_traits>

namespace rpnet {

constexpr int kBM = 128;          // pixels per tile (UMMA M)
constexpr int kBK = 64;           // fp16 channels per k-block = one 128-byte swizzle row
constexpr int kMaxTaps = 16;         // 3x3 tap lists, and the 4x4 stride-2 form of the up-conv data gradient (16 taps over 4 sources)
constexpr int kNumThreads = 192;
constexpr int kSmemBudget = 200 * 1024;
constexpr int kMaxBnGroups = 64;
constexpr int kMaxCosP = 8;       // prototypes of the fused calDist epilogue (1 + ways)
constexpr int kMaxBnCout = 1024;  // per-CTA shared accumulators of the fused BatchNorm statistics: [cout][2] doubles
constexpr int kMaxWsHaloKb = 12;  // stages per tile of the weights-stationary halo kernel (3 column offsets x k-blocks per tap)
constexpr int kMaxSeg = 6;        // K segments per tap (split-fp16 convs: hi.Wh, lo.Wh, hi.Wl for up to two concat sources)

struct ConvParams {
  int N, H, W;                    // pixel grid of the conv (input == output grid, stride 1)
  int bw_log2, bh_log2;           // tile box: bw x bh x bn pixels, bn = 128 / (bw * bh)
  int tiles_x, tiles_y, tiles_n;  // pixel tiles
  int n_tiles_c;                  // cout / BN
  // K loop of one tap = `nseg` segments; segment s reads seg_n[s] 64-channel chunks of tensor map seg_map[s] from chunk
  // seg_ach[s] on, against the weight chunks seg_wch[s].. of the packed weights.  Plain conv: (src0 | src1) channel concat.
  // Split-fp16 conv (x = hi + lo, w = Wh + Wl, all fp16): hi.Wh + lo.Wh + hi.Wl accumulated in fp32 — fp32-class products
  // from the fp16 tensor pipe (the dropped lo.Wl term is 2^-22 relative).
  // fp8-correction conv (common.cuh, "c8"): hi.Wh on fp16 operands (kind 0 segments) + 2^-15 (lo8.Wh8 + x8.Wl8) on e4m3 operands
  // (kind 1 segments: one 128-byte k-block of a c8 plane / pack = 64 channels of both corrections).  All kind 1 k-blocks of a tile
  // run first (kb8_per_tap per tap); the first kind 0 MMA then scales the accumulator by 2^-15.
  int nseg, kblocks_per_tap, kb8_per_tap;
  int seg_map[kMaxSeg], seg_ach[kMaxSeg], seg_wch[kMaxSeg], seg_n[kMaxSeg], seg_kind[kMaxSeg];
  int in_c8, out_c8;              // lo planes of the residual input / of the outputs are c8 planes instead of fp16 residuals
  int ntaps;
  int dy[kMaxTaps], dx[kMaxTaps];
  int tap_src[kMaxTaps];          // -1: channel concat of source 0 | source 1 (default); 0..3: the tap reads that source only
  int halo_tap[9];                // HALO kernels: tap index of offset (dy, dx) = (i / 3 - 1, i % 3 - 1) in the weight pack
  // weights-stationary HALO kernel: every 64-unit chunk of every tap of the pack is resident (chunk c of tap t at slot
  // t * ws_wchunks + c); stage kb of a tile (producer order) uses column offset ws_dx[kb] and weight chunk ws_wch[kb]
  int ws_wchunks;
  int ws_dx[kMaxWsHaloKb], ws_wch[kMaxWsHaloKb];
  const float* scale;             // per-cout epilogue: y = acc * scale + shift
  const float* shift;
  int relu;
  int in_bf16;                    // operands (activations AND weights) are bf16 instead of fp16 (dgrad path)
  int out_bf16;                   // `out` / `out_pool` are bf16 instead of fp16
  // fp16 NHWC output; pixel (n, y, x) of the conv grid lands at (n, y*oy_mul+oy_off, x*ox_mul+ox_off)
  __half* out;
  int out_H, out_W, out_C, out_coff;
  int oy_mul, oy_off, ox_mul, ox_off;
  __half* out_pool;               // optional 2x2/stride-2 max-pooled fp16 NHWC output (N, H/2, W/2, pool_C)
  int pool_C;
  // split-fp16 outputs: the residual fp16(v - fp16(v)) of what `out` / `out_pool` hold, same indexing (null = not written)
  __half* out_lo;
  __half* out_pool_lo;
  float* out_f32;                 // optional fp32 NHWC output (N, H, W, f32_C)
  int f32_C;
  // optional residual (torchvision BasicBlock: out = relu(bn2(conv2(.)) + identity)): fp16 NHWC [N][H][W][res_C], added after the
  // affine and before the ReLU
  const __half* res;
  const __half* res_lo;           // split-fp16 residual plane (null: the residual is plain fp16)
  int res_C;
  // optional fused calDist (net/rp_net.py:353-363) on the activated output of a 64-channel conv (the whole feature vector of
  // a pixel sits in one accumulator row): cos_pred[n][p][pixel] = cos_scaler * cos(y[n,pixel,:], cos_protos[n % cos_sets][p][:])
  const float* cos_protos;
  float* cos_pred;
  int cos_P, cos_sets;
  float cos_scaler;
  // optional fused train-mode BatchNorm statistics: bn_sums[g][cout][2] += {sum, sum of squares} of the fp32 accumulators
  // over the valid pixels of call group g (images [bn_start[g], bn_start[g+1])); a tile never straddles two groups
  double* bn_sums;                // fp64: the variance is a difference of nearly equal sums when |mean| >> std
  int bn_groups;
  int bn_start[kMaxBnGroups + 1];
};

// CTAS == 2: a CTA pair runs one M = 256 tile pair (two pixel tiles, the same BN couts); every CTA stages its own A tile and
// HALF of the weight tile, so a stage is 32 KB instead of 48 KB at BN = 256 and the shared-memory traffic per MMA (operand
// reads + TMA writes, what bounds the single-CTA kernel at ~2/3 of the tensor peak) drops by a third.
// WS == 1 (weights-stationary, single-cout-tile convs with at most kWsMaxKb k-blocks — the 64 -> 64 full-resolution layers):
// the whole packed weight tensor is loaded ONCE per CTA into a resident region and the ring carries activations only.  Those
// layers are bound by the L2 -> SM feed (ncu: 9.4 TB/s of TMA reads, a third of them the same 72 KB of weights re-fetched for
// every pixel tile): 1118 -> 964 us on the 96 x 256 x 256 layer.  (Combined with CTA pairs: no further gain — at N = 64 the MMA rate
// itself is the floor.)
// HALO == 1 (full 3x3 tap sets on 16 x 8 pixel tiles, cout tiles of 64 / 128): a stage carries the activation box of ONE column
// offset dx with a one-row halo above and below — (8 + 2) x 16 pixels x 64 channels — plus the three weight tiles of the taps
// (dy, dx), dy = -1, 0, 1; the three row offsets are views of the same box 16 rows (2048 B, swizzle-atom aligned) apart.  A
// pixel tile then pulls 3 x 20 KB of activations per 64-channel chunk through the L2 -> SM path instead of 9 x 16 KB: the
// small-channel full-resolution layers (Conv1.conv.3, Conv2.*) are bound by exactly that feed, three-fold so in split-fp16.
// WS == 1 && HALO == 1: both — the 64 -> 64 full-resolution layers in split precisions and their data gradients.  The halo kernel
// alone is bound by the aggregate L2 -> SM bandwidth (~30 B / clk / SM with all SMs pulling), and more than half of what a pixel
// tile pulls is the same 72 KB (bf16 data gradient) / 144 KB (split packs) of weights: resident, a tile pulls its activation
// boxes only (120 instead of 264 KB in split8).
constexpr int kWsMaxKb = 9;
constexpr int kWsHaloMaxKb = 18;  // resident weight k-blocks of the WS + HALO variant: 9 taps x 2 chunks (Wh | corrections, cin = 64)
constexpr int kHaloRows = (8 + 2) * 16;
template <int BN, int CTAS = 1, int WS = 0, int HALO = 0>
struct ConvCfg {
  static constexpr int kABytes = (HALO ? kHaloRows : kBM) * kBK * 2;   // 16 KB (20 KB with the halo rows)
  static constexpr int kBBytes = (BN / CTAS) * kBK * 2;
  static constexpr int kStageBytes = WS ? kABytes : kABytes + (HALO ? 3 : 1) * kBBytes;
  static constexpr int kResidentBytes = WS ? (HALO ? kWsHaloMaxKb : kWsMaxKb) * kBBytes : 0;
  static constexpr int kBnBytes = (WS ? 64 : kMaxBnCout) * 2 * 8;            // fused BN statistics (WS kernels: cout == 64)
  static constexpr int kBudget = (WS && HALO) ? 227 * 1024 - 4096 - kBnBytes : kSmemBudget;
  static constexpr int kStages = ((kBudget - kResidentBytes) / kStageBytes) > 8 ? 8 : ((kBudget - kResidentBytes) / kStageBytes);
  static constexpr int kTmemCols = 2 * BN;                        // double-buffered accumulator (power of 2 >= 32)
  static constexpr int kTileBytes = kStages * kStageBytes + kResidentBytes;
  static constexpr int kSmemBytes = kTileBytes + 1024 /*align*/ + 2 * BN * 4 * 2 /*scale,shift x2*/ + 256 + kBnBytes;
  static_assert(kSmemBytes <= 227 * 1024 && kStages >= 2, "shared-memory plan does not fit");
};

// The MMAs of one k-block (one pipeline stage).  KIND 0: fp16 / bf16 operands; 1: e4m3 correction block; 2: the first fp16 block after
// the corrections, whose first MMA scales the accumulator by 2^-15.  One straight-line instance per kind: predicating the kinds MMA by
// MMA inside one loop cost the issuing thread a quarter of the tensor throughput.
// b_addr[dyi]: the weight tile of row offset dyi (HALO), b_addr[0] the weight tile otherwise.
template <int CTAS, int HALO, int KIND>
__device__ __forceinline__ void conv_issue_kblock(uint32_t a_addr, const uint32_t (&b_addr)[3], uint32_t d_tmem, uint32_t idesc, uint32_t acc_first) {
  auto mma = [&](uint64_t a_desc, uint64_t b_desc, bool first) {
    const uint32_t acc = first ? acc_first : 1u;
    if (!elect_one()) return;
    if (KIND == 1) {
      if (CTAS == 2) umma_f8_pair(d_tmem, a_desc, b_desc, idesc, acc);
      else           umma_f8(d_tmem, a_desc, b_desc, idesc, acc);
    } else if (KIND == 2 && first) {
      if (CTAS == 2) umma_f16_pair_sd15(d_tmem, a_desc, b_desc, idesc);
      else           umma_f16_sd15(d_tmem, a_desc, b_desc, idesc);
    } else {
      if (CTAS == 2) umma_f16_pair(d_tmem, a_desc, b_desc, idesc, acc);
      else           umma_f16(d_tmem, a_desc, b_desc, idesc, acc);
    }
  };
  if (HALO) {
    // three row offsets of the same halo box: 16 pixel rows = 2048 B apart (whole swizzle atoms), one weight tile each
#pragma unroll
    for (int dyi = 0; dyi < 3; ++dyi) {
      const uint64_t a_desc = umma_desc_sw128(a_addr + dyi * 16 * 128, 1024);
      const uint64_t b_desc = umma_desc_sw128(b_addr[dyi], 1024);
#pragma unroll
      for (int k = 0; k < kBK / 16; ++k) mma(a_desc + 2 * k, b_desc + 2 * k, (dyi | k) == 0);
    }
  } else {
    const uint64_t a_desc = umma_desc_sw128(a_addr, 1024);
    const uint64_t b_desc = umma_desc_sw128(b_addr[0], 1024);
    // advance 16 fp16 (32 e4m3) = 32 bytes along K inside the 128B swizzle row: +2 in the (addr >> 4) field
#pragma unroll
    for (int k = 0; k < kBK / 16; ++k) mma(a_desc + 2 * k, b_desc + 2 * k, k == 0);
  }
}

template <int BN, int CTAS, int WS, int HALO = 0>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tm_src0, const __grid_constant__ CUtensorMap tm_src1,
                  const __grid_constant__ CUtensorMap tm_src2, const __grid_constant__ CUtensorMap tm_src3,
                  const __grid_constant__ CUtensorMap tm_w, const ConvParams p) {
  using Cfg = ConvCfg<BN, CTAS, WS, HALO>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint8_t* w_res = smem + kStages * Cfg::kStageBytes;                             // WS: all k-blocks of the packed weights
  float* s_scale = reinterpret_cast<float*>(smem + Cfg::kTileBytes);              // [2][BN]
  float* s_shift = s_scale + 2 * BN;                                               // [2][BN]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + 2 * BN);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;     // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint64_t* ws_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ws_bar + 1);
  double* s_bn = reinterpret_cast<double*>(smem + Cfg::kTileBytes + 2 * BN * 4 * 2 + 256);   // [cout][2]

  // broadcast from lane 0: the compiler then knows the warp index (and every role branch on it) is warp-uniform
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int kblocks_per_tap = p.kblocks_per_tap;
  const int num_kb = (HALO ? 3 : p.ntaps) * kblocks_per_tap;      // stages per tile
  const int num_m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  // work unit = CTAS consecutive pixel tiles x one cout tile; CTA `cta_rank` of the pair owns pixel tile unit * CTAS + cta_rank
  // (past the end for the odd one out: its loads are out of bounds = zeros and its pixels fail the `valid` test)
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const int num_tiles = ((num_m_tiles + CTAS - 1) / CTAS) * p.n_tiles_c;
  const int tile_first = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = CTAS == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int bw = 1 << p.bw_log2, bh = 1 << p.bh_log2;
  const int bn = kBM >> (p.bw_log2 + p.bh_log2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_src0);
    tma_prefetch_desc(&tm_src1);
    tma_prefetch_desc(&tm_src2);
    tma_prefetch_desc(&tm_src3);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4 * CTAS);     // one arrive per epilogue warp (of both CTAs: the leader's barrier gates the MMAs)
    }
    mbar_init(ws_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    if (CTAS == 2) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
    else           tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncwarp();
  if (CTAS == 2) cluster_sync_all();      // the peer's barriers are initialised before anything arrives on them
  else           __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (WS && HALO) {                                    // every chunk of every tap, once: slot tap * ws_wchunks + chunk
        mbar_expect_tx(ws_bar, 9 * p.ws_wchunks * Cfg::kBBytes);
        for (int t = 0; t < 9; ++t)
          for (int c = 0; c < p.ws_wchunks; ++c)
            tma_load_3d(&tm_w, ws_bar, w_res + (t * p.ws_wchunks + c) * Cfg::kBBytes, c * kBK, 0, t);
      } else if (WS) {                                     // the whole weight tensor, once (n_tiles_c == 1, num_kb <= kWsMaxKb)
        if (CTAS == 2) {                                   // each CTA its half of the couts; both complete on the leader's barrier
          if (cta_rank == 0) mbar_expect_tx(ws_bar, 2 * num_kb * Cfg::kBBytes);
          for (int kb = 0; kb < num_kb; ++kb)
            tma_load_3d_pair(&tm_w, ws_bar, w_res + kb * Cfg::kBBytes, (kb % kblocks_per_tap) * kBK, (int)cta_rank * (BN / 2),
                             kb / kblocks_per_tap);
        } else {
          mbar_expect_tx(ws_bar, num_kb * Cfg::kBBytes);
          for (int kb = 0; kb < num_kb; ++kb)
            tma_load_3d(&tm_w, ws_bar, w_res + kb * Cfg::kBBytes, (kb % kblocks_per_tap) * kBK, 0, kb / kblocks_per_tap);
        }
      }
      for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
        const int ct = tile % p.n_tiles_c;
        int mt = (tile / p.n_tiles_c) * CTAS + (int)cta_rank;
        const int tx = mt % p.tiles_x;  mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int tn = mt / p.tiles_y;
        const int x0 = tx * bw, y0 = ty * bh, n0 = tn * bn;
        if (HALO) {
          for (int ph = p.kb8_per_tap ? 1 : 0; ph >= 0; --ph)
          for (int dxi = 0; dxi < 3; ++dxi) {
            for (int s = 0; s < p.nseg; ++s) {
              if (p.seg_kind[s] != ph) continue;
              const int mi = p.seg_map[s];
              const CUtensorMap* tm = mi == 0 ? &tm_src0 : (mi == 1 ? &tm_src1 : (mi == 2 ? &tm_src2 : &tm_src3));
              const int ach0 = p.seg_ach[s], wch0 = p.seg_wch[s], sn = p.seg_n[s];
              for (int j = 0; j < sn; ++j) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* a_dst = tiles + stage * Cfg::kStageBytes;
                uint8_t* b_dst = a_dst + Cfg::kABytes;
                if (CTAS == 2) {
                  if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
                  tma_load_4d_pair(tm, &full_bar[stage], a_dst, (ach0 + j) * kBK, x0 + dxi - 1, y0 - 1, n0);
#pragma unroll
                  for (int dyi = 0; dyi < 3; ++dyi)
                    tma_load_3d_pair(&tm_w, &full_bar[stage], b_dst + dyi * Cfg::kBBytes, (wch0 + j) * kBK,
                                     ct * BN + (int)cta_rank * (BN / 2), p.halo_tap[dyi * 3 + dxi]);
                } else {
                  mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                  tma_load_4d(tm, &full_bar[stage], a_dst, (ach0 + j) * kBK, x0 + dxi - 1, y0 - 1, n0);
                  if (!WS) {
#pragma unroll
                    for (int dyi = 0; dyi < 3; ++dyi)
                      tma_load_3d(&tm_w, &full_bar[stage], b_dst + dyi * Cfg::kBBytes, (wch0 + j) * kBK, ct * BN, p.halo_tap[dyi * 3 + dxi]);
                  }
                }
                if (++stage == kStages) { stage = 0; phase ^= 1; }
              }
            }
          }
        } else
        for (int ph = p.kb8_per_tap ? 1 : 0; ph >= 0; --ph)
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int xs = x0 + p.dx[tap], ys = y0 + p.dy[tap];
          const int ts = p.tap_src[tap];
          for (int s = 0; s < p.nseg; ++s) {
            if (p.seg_kind[s] != ph) continue;
            const int mi = ts >= 0 ? ts : p.seg_map[s];
            const CUtensorMap* tm = mi == 0 ? &tm_src0 : (mi == 1 ? &tm_src1 : (mi == 2 ? &tm_src2 : &tm_src3));
            const int ach0 = p.seg_ach[s], wch0 = p.seg_wch[s], sn = p.seg_n[s];
            for (int j = 0; j < sn; ++j) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* a_dst = tiles + stage * Cfg::kStageBytes;
              uint8_t* b_dst = a_dst + Cfg::kABytes;
              if (CTAS == 2) {
                // both CTAs' bytes complete on the leader's barrier; only the leader arrives on it
                if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
                tma_load_4d_pair(tm, &full_bar[stage], a_dst, (ach0 + j) * kBK, xs, ys, n0);
                if (!WS) tma_load_3d_pair(&tm_w, &full_bar[stage], b_dst, (wch0 + j) * kBK, ct * BN + (int)cta_rank * (BN / 2), tap);
              } else {
                mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                tma_load_4d(tm, &full_bar[stage], a_dst, (ach0 + j) * kBK, xs, ys, n0);
                if (!WS) tma_load_3d(&tm_w, &full_bar[stage], b_dst, (wch0 + j) * kBK, ct * BN, tap);
              }
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop (warp-uniform control flow and operands: descriptors and addresses stay in uniform registers)
    // and one elected lane issues each tcgen05 instruction.  Entering the loop with a single lane instead makes the compiler
    // wrap every MMA in a vector-to-uniform register hand-over loop (ELECT / R2UR / BRA.U.ANY, ~14 instructions per MMA), which
    // bounds the N = 64 / 128 tiles whose MMAs last 32 / 64 cycles.
    if (cta_rank == 0) {
      const uint32_t idesc8 = umma_idesc_f16(kBM * CTAS, BN);      // a / b format 0 = e4m3 for kind::f8f6f4 (F16 for kind::f16)
      const uint32_t idesc = idesc8 | (p.in_bf16 ? ((1u << 7) | (1u << 10)) : 0u);
      const int nkb8 = (HALO ? 3 : p.ntaps) * p.kb8_per_tap;       // leading e4m3 k-blocks of every tile
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (WS) mbar_wait(ws_bar, 0);
      for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        if (CTAS == 2) mbar_wait_cluster(&tempty_bar[as], aphase ^ 1);
        else           mbar_wait(&tempty_bar[as], aphase ^ 1);           // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        // one pipeline stage: wait for its operands, issue its MMAs (kind chosen at compile time), hand the slot back
        auto kblock = [&](int kb, auto kind) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + stage * Cfg::kStageBytes);
          uint32_t b_addr[3];
          if (WS && HALO) {                                // resident weights: the three taps (dy, ws_dx[kb]) of chunk ws_wch[kb]
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi)
              b_addr[dyi] = smem_u32(w_res + (p.halo_tap[dyi * 3 + p.ws_dx[kb]] * p.ws_wchunks + p.ws_wch[kb]) * Cfg::kBBytes);
          } else {
            const uint32_t b0 = WS ? smem_u32(w_res + kb * Cfg::kBBytes) : a_addr + Cfg::kABytes;
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) b_addr[dyi] = b0 + dyi * Cfg::kBBytes;
          }
          conv_issue_kblock<CTAS, HALO, decltype(kind)::value>(a_addr, b_addr, d_tmem, decltype(kind)::value == 1 ? idesc8 : idesc, kb != 0);
          // smem slot reusable once these MMAs retire (in both CTAs of a pair)
          if (elect_one()) {
            if (CTAS == 2) umma_commit_pair(&empty_bar[stage], 3);
            else           umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        };
        int kb = 0;
        if (nkb8 > 0) {                                    // e4m3 corrections first, then the block that rescales the accumulator
          for (; kb < nkb8; ++kb) kblock(kb, std::integral_constant<int, 1>());
          kblock(kb++, std::integral_constant<int, 2>());
        }
        for (; kb < num_kb; ++kb) kblock(kb, std::integral_constant<int, 0>());
        if (elect_one()) {
          if (CTAS == 2) umma_commit_pair(&tfull_bar[as], 3);  // accumulator complete -> the epilogue of each CTA
          else           umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int q = warp & 3;                                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                          // accumulator row == pixel within the tile
    const int px = row & (bw - 1);
    const int py = (row >> p.bw_log2) & (bh - 1);
    const int pn = row >> (p.bw_log2 + p.bh_log2);
    const int et = threadIdx.x - 64;                        // 0..127
    const int cout_all = p.n_tiles_c * BN;
    int cur_g = -1;                                         // BatchNorm call group of the statistics held in s_bn
    if (p.bn_sums) {
      for (int i = et; i < cout_all * 2; i += 128) s_bn[i] = 0.0;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    int it = 0;
    for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int ct = tile % p.n_tiles_c;
      int mt = (tile / p.n_tiles_c) * CTAS + (int)cta_rank;
      const int tx = mt % p.tiles_x;  mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int tn = mt / p.tiles_y;
      const int x = tx * bw + px, y = ty * bh + py, n = tn * bn + pn;
      const bool valid = (x < p.W) && (y < p.H) && (n < p.N);
      if (p.bn_sums) {
        int g = 0;
        while (g + 1 < p.bn_groups && tn * bn >= p.bn_start[g + 1]) ++g;
        if (g != cur_g) {                                   // tiles are ordered by image: at most bn_groups flushes per CTA
          if (cur_g >= 0) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = et; i < cout_all * 2; i += 128) {
              atomicAdd(p.bn_sums + (size_t)cur_g * cout_all * 2 + i, s_bn[i]);
              s_bn[i] = 0.0;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");     // nobody accumulates the new group before the reset
          }
          cur_g = g;
        }
      }
      // stage this tile's per-channel scale/shift (buffer `as`: the other buffer may still be in use); with a single cout tile
      // they are staged once, before the first tile (the global-load latency + barrier per tile bounded the N = 64 layers)
      float* sc = s_scale + (p.n_tiles_c == 1 ? 0 : as * BN);
      float* sh = s_shift + (p.n_tiles_c == 1 ? 0 : as * BN);
      if (p.n_tiles_c > 1 || it == 0) {
        for (int i = et; i < BN; i += 128) {
          sc[i] = __ldg(p.scale + ct * BN + i);
          sh[i] = __ldg(p.shift + ct * BN + i);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      __half* o16 = nullptr;
      __half* o16lo = nullptr;
      __half* opool = nullptr;
      __half* opoollo = nullptr;
      float* o32 = nullptr;
      if (p.out) {
        const size_t off = (static_cast<size_t>(n * p.out_H + y * p.oy_mul + p.oy_off) * p.out_W + x * p.ox_mul + p.ox_off) *
                               p.out_C + p.out_coff + ct * BN;
        o16 = p.out + off;
        if (p.out_lo) o16lo = p.out_lo + off;
      }
      if (p.out_pool) {
        const size_t off = (static_cast<size_t>(n * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1)) * p.pool_C + ct * BN;
        opool = p.out_pool + off;
        if (p.out_pool_lo) opoollo = p.out_pool_lo + off;
      }
      if (p.out_f32) o32 = p.out_f32 + (static_cast<size_t>(n * p.H + y) * p.W + x) * p.f32_C + ct * BN;
      const bool pool_writer = valid && ((x & 1) == 0) && ((y & 1) == 0);
      float cos_nn = 0.f, cos_pn[kMaxCosP], cos_dot[kMaxCosP];
#pragma unroll
      for (int q2 = 0; q2 < kMaxCosP; ++q2) { cos_pn[q2] = 0.f; cos_dot[q2] = 0.f; }
      const float* cos_pr = p.cos_pred ? p.cos_protos + (size_t)((valid ? n : 0) % p.cos_sets) * p.cos_P * 64 : nullptr;
      // one 32-channel chunk of the accumulator row, already in registers
      auto chunk = [&](float* v, const int c0) {
        if (p.res) {
          float r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0.f;
          if (valid) {
            const size_t roff = (static_cast<size_t>(n * p.H + y) * p.W + x) * p.res_C + ct * BN + c0;
            const uint4* rp = reinterpret_cast<const uint4*>(p.res + roff);
#pragma unroll
            for (int j = 0; j < 4; ++j) unpack8_f16(__ldg(rp + j), r + 8 * j);
            if (p.res_lo) {
#pragma unroll
              for (int j = 0; j < 4; ++j) lo8_add(p.res_lo, p.in_c8, roff + 8 * j, (c0 + 8 * j) & 63, r + 8 * j);
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float t = fmaf(v[j], sc[c0 + j], sh[c0 + j]) + r[j];
            v[j] = p.relu ? fmaxf(t, 0.f) : t;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float t = fmaf(v[j], sc[c0 + j], sh[c0 + j]);
            v[j] = p.relu ? fmaxf(t, 0.f) : t;
          }
        }
        if (BN == 64 && p.cos_pred) {
#pragma unroll
          for (int j = 0; j < 32; ++j) cos_nn = fmaf(v[j], v[j], cos_nn);
#pragma unroll
          for (int q2 = 0; q2 < kMaxCosP; ++q2) {
            if (q2 < p.cos_P) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 pr = __ldg(reinterpret_cast<const float4*>(cos_pr + q2 * 64 + c0 + j));
                cos_dot[q2] = fmaf(v[j], pr.x, fmaf(v[j + 1], pr.y, fmaf(v[j + 2], pr.z, fmaf(v[j + 3], pr.w, cos_dot[q2]))));
                cos_pn[q2] = fmaf(pr.x, pr.x, fmaf(pr.y, pr.y, fmaf(pr.z, pr.z, fmaf(pr.w, pr.w, cos_pn[q2]))));
              }
            }
          }
        }
        if (p.bn_sums) {
          // per-channel sums over the 32 pixels of this warp: butterfly transpose-reduce (31 shuffles per array); lane l
          // ends up with the sum of column c0 + l
          float s1[32], s2[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float t = valid ? v[j] : 0.f;
            s1[j] = t;
            s2[j] = t * t;
          }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < off; ++j) {
              const float k1 = up ? s1[j + off] : s1[j], g1 = up ? s1[j] : s1[j + off];
              const float k2 = up ? s2[j + off] : s2[j], g2 = up ? s2[j] : s2[j + off];
              s1[j] = k1 + __shfl_xor_sync(0xffffffffu, g1, off);
              s2[j] = k2 + __shfl_xor_sync(0xffffffffu, g2, off);
            }
          }
          // fp32 sums of 32 pixels, accumulated across tiles / warps / CTAs in fp64
          atomicAdd(&s_bn[(ct * BN + c0 + lane) * 2], (double)s1[0]);
          atomicAdd(&s_bn[(ct * BN + c0 + lane) * 2 + 1], (double)s2[0]);
        }
        if (o32 && valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(o32 + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (o16 && valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const uint4 hi = p.out_bf16 ? pack8_bf16(v + j) : pack8_f16(v + j);
            *reinterpret_cast<uint4*>(o16 + c0 + j) = hi;
            if (o16lo) lo8_store(p.out_lo, p.out_c8, static_cast<size_t>(o16lo - p.out_lo) + c0 + j, (c0 + j) & 63, v + j, hi);
          }
        }
        if (p.out_pool) {      // 2x2 max over (x, x^1) and (y, y^1): partner lanes lane^1 and lane^bw
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float t = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
            v[j] = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, bw));
          }
          if (pool_writer) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 hi = p.out_bf16 ? pack8_bf16(v + j) : pack8_f16(v + j);
              *reinterpret_cast<uint4*>(opool + c0 + j) = hi;
              if (opoollo) lo8_store(p.out_pool_lo, p.out_c8, static_cast<size_t>(opoollo - p.out_pool_lo) + c0 + j, (c0 + j) & 63, v + j, hi);
            }
          }
        }
      };
      // the TMEM load of the next chunk is in flight while this one is processed; the accumulator buffer goes back to the MMA
      // warp as soon as its last chunk sits in registers (before that chunk's math and stores)
      auto release = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTAS == 2) mbar_arrive_rank0(&tempty_bar[as]);
          else           mbar_arrive(&tempty_bar[as]);
        }
      };
      {
        float va[32], vb[32];
        tmem_ld32(t_addr, va);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 64) {
          tmem_ld_wait();
          tmem_ld32(t_addr + c0 + 32, vb);
          chunk(va, c0);
          tmem_ld_wait();
          if (c0 + 64 < BN) tmem_ld32(t_addr + c0 + 64, va);
          else              release();
          chunk(vb, c0 + 32);
        }
      }
      if (BN == 64 && p.cos_pred && valid) {
        // torch's cosine_similarity: each norm clamped at 1e-8 separately
        const float xn = fmaxf(sqrtf(cos_nn), 1e-8f);
        const size_t hw = (size_t)p.H * p.W;
#pragma unroll
        for (int q2 = 0; q2 < kMaxCosP; ++q2)
          if (q2 < p.cos_P)
            p.cos_pred[((size_t)n * p.cos_P + q2) * hw + (size_t)y * p.W + x] =
                p.cos_scaler * (cos_dot[q2] / (xn * fmaxf(sqrtf(cos_pn[q2]), 1e-8f)));
      }
    }
    if (p.bn_sums && cur_g >= 0) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = et; i < cout_all * 2; i += 128) atomicAdd(p.bn_sums + (size_t)cur_g * cout_all * 2 + i, s_bn[i]);
    }
  }
  tc_fence_before();
  __syncwarp();
  if (CTAS == 2) cluster_sync_all();      // both CTAs are done with the pair's TMEM and with each other's barriers
  else           __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
    else           tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// fp16 / bf16 tensor map with 128B swizzle; dims/box innermost first; strides in elements for dims 1..rank-1.
int make_tmap_2b(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                 const uint32_t* box, bool bf16) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return RPNET_ERR_DRIVER;
  }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_elems[i] * 2;
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return RPNET_ERR_DRIVER;
  }
  return 0;
}

int pow2_floor(int v) { int r = 1; while (r * 2 <= v) r *= 2; return r; }
int ilog2(int v) { int r = 0; while ((1 << r) < v) ++r; return r; }

static int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN>
static int launch(const CUtensorMap& t0, const CUtensorMap& t1, const CUtensorMap& t2, const CUtensorMap& t3, const CUtensorMap& tw,
                  const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<BN, 1, 0>;
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<BN, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles_c;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  conv_igemm_kernel<BN, 1, 0><<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(t0, t1, t2, t3, tw, p);
  return check_cuda(cudaGetLastError(), "conv_igemm_kernel launch");
}

// Weights-stationary launch (cout == 64, at most kWsMaxKb k-blocks).
static int launch_ws(const CUtensorMap& t0, const CUtensorMap& t1, const CUtensorMap& t2, const CUtensorMap& t3, const CUtensorMap& tw,
                     const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<64, 1, 1>;
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<64, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  conv_igemm_kernel<64, 1, 1><<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(t0, t1, t2, t3, tw, p);
  return check_cuda(cudaGetLastError(), "conv_igemm_kernel (weights-stationary) launch");
}

// CTA-pair launch: clusters of two CTAs, one pair per TPC, persistent over the tile-pair units.
template <int BN>
static int launch_pair(const CUtensorMap& t0, const CUtensorMap& t1, const CUtensorMap& t2, const CUtensorMap& t3, const CUtensorMap& tw,
                       const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<BN, 2, 0>;
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<BN, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int units = ((p.tiles_x * p.tiles_y * p.tiles_n + 1) / 2) * p.n_tiles_c;
  const int pairs = units < num_sms() / 2 ? units : num_sms() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs, 1, 1);
  cfg.blockDim = dim3(kNumThreads, 1, 1);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return check_cuda(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BN, 2, 0>, t0, t1, t2, t3, tw, p), "conv_igemm_kernel (CTA pair) launch");
}

// Halo launches: single CTA for 64-wide cout tiles, CTA pair for 128-wide ones (a pair halves the three weight tiles per CTA).
static int launch_halo64(const CUtensorMap& t0, const CUtensorMap& t1, const CUtensorMap& t2, const CUtensorMap& t3, const CUtensorMap& tw,
                         const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<64, 1, 0, 1>;
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<64, 1, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles_c;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  conv_igemm_kernel<64, 1, 0, 1><<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(t0, t1, t2, t3, tw, p);
  return check_cuda(cudaGetLastError(), "conv_igemm_kernel (halo) launch");
}

// Weights-stationary halo launch (cout == 64, one source, at most kWsHaloMaxKb resident weight k-blocks).
static int launch_ws_halo64(const CUtensorMap& t0, const CUtensorMap& t1, const CUtensorMap& t2, const CUtensorMap& t3, const CUtensorMap& tw,
                            const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<64, 1, 1, 1>;
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<64, 1, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  conv_igemm_kernel<64, 1, 1, 1><<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(t0, t1, t2, t3, tw, p);
  return check_cuda(cudaGetLastError(), "conv_igemm_kernel (weights-stationary halo) launch");
}

static int launch_halo128_pair(const CUtensorMap& t0, const CUtensorMap& t1, const CUtensorMap& t2, const CUtensorMap& t3,
                               const CUtensorMap& tw, const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<128, 2, 0, 1>;
  static bool attr_set = false;
  if (!attr_set) {
    RPNET_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<128, 2, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int units = ((p.tiles_x * p.tiles_y * p.tiles_n + 1) / 2) * p.n_tiles_c;
  const int pairs = units < num_sms() / 2 ? units : num_sms() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs, 1, 1);
  cfg.blockDim = dim3(kNumThreads, 1, 1);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return check_cuda(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<128, 2, 0, 1>, t0, t1, t2, t3, tw, p), "conv_igemm_kernel (halo, CTA pair) launch");
}

// RPNET_CONV_2CTA: bit mask of the cout tile widths that run as CTA pairs (1: 256, 2: 128, 4: 64).  Default 1: measured on
// B200, pairs gain 10 % at BN = 256 (1344 -> 1479 TF/s) and nothing at 128 / 64 — those layers are bound by the L2 -> SM feed of
// the nine per-tap activation boxes, not by MMA issue or operand reads.
static bool pair_enabled(int BN) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RPNET_CONV_2CTA"); v = e ? atoi(e) : 1; }
  return (v & (BN == 256 ? 1 : (BN == 128 ? 2 : 4))) != 0;
}

}  // namespace rpnet

using namespace rpnet;

extern "C" int rpnet_bn_stats_f16(const void* z, int n, int h, int w, int c, const int* group_start, int groups, double* sums,
                                  void* stream);
extern "C" int rpnet_bn_stats_split_f16(const void* z_hi, const void* z_lo, int n, int h, int w, int c, const int* group_start,
                                        int groups, double* sums, void* stream);

RPNET_API int rpnet_conv_split_res_f16(const void* src0_hi, const void* src0_lo, int c0, const void* src1_hi, const void* src1_lo, int c1,
                                        int n, int h, int w, const void* wpack, int w_split, int ntaps, const int* tap_dy,
                                        const int* tap_dx, int cout, const float* scale, const float* shift, const void* res_hi,
                                        const void* res_lo, int relu, void* out_hi, void* out_lo, int out_h, int out_w, int out_c,
                                        int out_coff, int oy_mul, int oy_off, int ox_mul, int ox_off, void* out_pool_hi,
                                        void* out_pool_lo, float* out_f32, const int* group_start, int groups, double* sums,
                                        int keep_sums, void* stream_);

namespace {
// One launch of conv_igemm_kernel, as the C-ABI entry points describe it.
struct ConvCall {
  bool bf16 = false;
  const void* src0 = nullptr; int c0 = 0;
  const void* src1 = nullptr; int c1 = 0;
  // split-fp16: residual planes of the sources (same shapes) and a weight pack [ntaps][cout][2 * (c0 + c1)] = Wh | Wl
  const void* src0_lo = nullptr;
  const void* src1_lo = nullptr;
  // w_split: 0 fp16 weights; 1 pack [ntaps][cout][Wh | Wl] (three fp16 passes); 2 / 3 fp8 corrections: pack [ntaps][cout][Wh |
  // per 64 channels (Wh8 | Wl8)], the sources' lo planes are c8 planes (common.cuh); 2 writes c8 lo planes, 3 fp16 residual planes
  int w_split = 0;
  int n = 0, h = 0, w = 0;
  const void* wpack = nullptr; int ntaps = 0; const int* tap_dy = nullptr; const int* tap_dx = nullptr; int cout = 0;
  const float* scale = nullptr; const float* shift = nullptr; int relu = 0;
  void* out = nullptr; void* out_lo = nullptr; int out_h = 0, out_w = 0, out_c = 0, out_coff = 0;
  int oy_mul = 1, oy_off = 0, ox_mul = 1, ox_off = 0;
  void* out_pool = nullptr; void* out_pool_lo = nullptr; float* out_f32 = nullptr;
  void* stream = nullptr;
  const int* bn_group_start = nullptr; int bn_groups = 0; double* bn_sums = nullptr; int* bn_fused = nullptr; int bn_keep_sums = 0;
  const float* cos_protos = nullptr; int cos_P = 0, cos_sets = 0; float cos_scaler = 0.f; float* cos_pred = nullptr;
  const void* const* view_ptr = nullptr; const long long* view_strides = nullptr; const int* tap_src = nullptr;
  const void* res = nullptr;
  const void* res_lo = nullptr;
};
}  // namespace

static int conv_igemm_run(const ConvCall& a) {
  cudaStream_t stream = static_cast<cudaStream_t>(a.stream);
  const int c0 = a.c0, c1 = a.c1, n = a.n, h = a.h, w = a.w, ntaps = a.ntaps, cout = a.cout;
  const bool bf16 = a.bf16;
  RPNET_REQUIRE(a.src0 && a.wpack && a.scale && a.shift, "conv_igemm: null pointer argument");
  RPNET_REQUIRE(c0 > 0 && c0 % kBK == 0 && c1 >= 0 && c1 % kBK == 0, "conv_igemm: channel counts must be multiples of 64 (got %d, %d)", c0, c1);
  RPNET_REQUIRE(c1 == 0 || a.src1, "conv_igemm: src1 is null but c1 = %d", c1);
  RPNET_REQUIRE(n > 0 && h > 0 && w > 0, "conv_igemm: bad grid %d x %d x %d", n, h, w);
  RPNET_REQUIRE(ntaps >= 1 && ntaps <= kMaxTaps, "conv_igemm: ntaps %d out of range [1, %d]", ntaps, kMaxTaps);
  RPNET_REQUIRE(cout >= 64 && cout % 64 == 0, "conv_igemm: cout must be a multiple of 64 (got %d)", cout);
  RPNET_REQUIRE(a.out || a.out_pool || a.out_f32 || a.cos_pred, "conv_igemm: no output requested");
  RPNET_REQUIRE(!a.out_pool || (h % 2 == 0 && w % 2 == 0), "conv_igemm: fused 2x2 max-pool needs even H, W (got %d x %d)", h, w);
  RPNET_REQUIRE((!a.out_lo || a.out) && (!a.out_pool_lo || a.out_pool), "conv_igemm: a residual output needs its main output");
  const bool a_split = a.src0_lo != nullptr;
  RPNET_REQUIRE(!a_split || c1 == 0 || a.src1_lo, "conv_igemm: split sources need both residual planes");
  RPNET_REQUIRE(!(a_split || a.w_split) || (!a.tap_src && !bf16), "conv_igemm: split-fp16 operands are fp16, without per-tap views");
  const int BN = (cout % 256 == 0) ? 256 : (cout % 128 == 0 ? 128 : 64);
  const bool c8 = a.w_split >= 2;
  RPNET_REQUIRE(a.w_split >= 0 && a.w_split <= 3, "conv_igemm: w_split %d out of range [0, 3]", a.w_split);
  RPNET_REQUIRE(!c8 || a_split, "conv_igemm: the fp8-correction pack needs the sources' c8 planes");
  RPNET_REQUIRE(!c8 || !a.out_lo || (a.out_c % 64 == 0 && a.out_coff % 64 == 0), "conv_igemm: c8 output planes need 64-channel groups");

  ConvParams p{};
  p.N = n; p.H = h; p.W = w;
  const int bw = pow2_floor(w < 16 ? w : 16);
  const int bh = pow2_floor(h < kBM / bw ? h : kBM / bw);
  const int bn = kBM / (bw * bh);
  RPNET_REQUIRE(!a.out_pool || (bw >= 2 && bh >= 2), "conv_igemm: fused pooling needs a tile of at least 2x2 pixels");
  p.bw_log2 = ilog2(bw); p.bh_log2 = ilog2(bh);
  p.tiles_x = (w + bw - 1) / bw; p.tiles_y = (h + bh - 1) / bh; p.tiles_n = (n + bn - 1) / bn;
  p.n_tiles_c = cout / BN;
  {
    // K segments of one tap: tensor maps 0 / 1 = the sources, 2 / 3 = their residual planes
    const int n0 = c0 / kBK, n1 = c1 / kBK, ncin = n0 + n1;
    int s = 0;
    auto seg = [&](int map, int wch, int cnt, int kind) {
      if (cnt > 0) { p.seg_map[s] = map; p.seg_ach[s] = 0; p.seg_wch[s] = wch; p.seg_n[s] = cnt; p.seg_kind[s] = kind; ++s; }
    };
    seg(0, 0, n0, 0);
    seg(1, n0, n1, 0);
    if (c8) {                                      // e4m3 corrections: c8 planes against the second half of the pack
      seg(2, ncin, n0, 1);
      seg(3, ncin + n0, n1, 1);
    } else {
      if (a_split) { seg(2, 0, n0, 0); seg(3, n0, n1, 0); }
      if (a.w_split) { seg(0, ncin, n0, 0); seg(1, ncin + n0, n1, 0); }
    }
    p.nseg = s;
    p.kblocks_per_tap = 0; p.kb8_per_tap = 0;
    for (int i = 0; i < s; ++i) { p.kblocks_per_tap += p.seg_n[i]; if (p.seg_kind[i]) p.kb8_per_tap += p.seg_n[i]; }
    p.in_c8 = c8 ? 1 : 0; p.out_c8 = a.w_split == 2 ? 1 : 0;
  }
  p.ntaps = ntaps;
  for (int i = 0; i < ntaps; ++i) { p.dy[i] = a.tap_dy[i]; p.dx[i] = a.tap_dx[i]; p.tap_src[i] = a.tap_src ? a.tap_src[i] : -1; }
  if (a.tap_src) {
    RPNET_REQUIRE(a.view_ptr && a.view_strides && c1 == 0, "conv_igemm: per-tap sources need the four views and no channel concat");
    for (int i = 0; i < ntaps; ++i) RPNET_REQUIRE(a.tap_src[i] >= 0 && a.tap_src[i] < 4, "conv_igemm: tap source %d out of range", a.tap_src[i]);
  }
  p.scale = a.scale; p.shift = a.shift; p.relu = a.relu;
  p.in_bf16 = bf16 ? 1 : 0; p.out_bf16 = bf16 ? 1 : 0;
  p.out = static_cast<__half*>(a.out); p.out_lo = static_cast<__half*>(a.out_lo);
  p.out_H = a.out_h; p.out_W = a.out_w; p.out_C = a.out_c; p.out_coff = a.out_coff;
  p.oy_mul = a.oy_mul; p.oy_off = a.oy_off; p.ox_mul = a.ox_mul; p.ox_off = a.ox_off;
  p.out_pool = static_cast<__half*>(a.out_pool); p.out_pool_lo = static_cast<__half*>(a.out_pool_lo); p.pool_C = cout;
  p.out_f32 = a.out_f32; p.f32_C = cout;
  p.bn_sums = nullptr; p.bn_groups = 0;
  if (a.bn_fused) *a.bn_fused = 0;
  p.res = static_cast<const __half*>(a.res); p.res_lo = static_cast<const __half*>(a.res_lo); p.res_C = cout;
  RPNET_REQUIRE(!a.res_lo || a.res, "conv_igemm: a residual lo plane needs the residual");
  p.cos_pred = nullptr; p.cos_protos = nullptr; p.cos_P = 0; p.cos_sets = 1; p.cos_scaler = 0.f;
  if (a.cos_pred) {
    RPNET_REQUIRE(cout == 64 && a.cos_protos && a.cos_P >= 1 && a.cos_P <= kMaxCosP && a.cos_sets >= 1,
                  "conv_igemm: the fused calDist epilogue needs cout == 64 and 1..%d prototypes (got cout=%d, P=%d)", kMaxCosP, cout, a.cos_P);
    p.cos_pred = a.cos_pred; p.cos_protos = a.cos_protos; p.cos_P = a.cos_P; p.cos_sets = a.cos_sets; p.cos_scaler = a.cos_scaler;
  }
  if (a.bn_sums) {
    // fuse the statistics when no pixel tile straddles two call groups and the per-CTA accumulators fit
    RPNET_REQUIRE(a.bn_groups >= 1 && a.bn_groups <= kMaxBnGroups && a.bn_group_start, "conv_igemm: bn groups %d out of range [1, %d]", a.bn_groups, kMaxBnGroups);
    RPNET_REQUIRE(a.bn_group_start[0] == 0 && a.bn_group_start[a.bn_groups] == n, "conv_igemm: bn group_start must span [0, %d]", n);
    bool ok = cout <= kMaxBnCout;
    for (int g = 1; g < a.bn_groups; ++g) ok = ok && (a.bn_group_start[g] % bn == 0);
    if (ok) {
      p.bn_sums = a.bn_sums; p.bn_groups = a.bn_groups;
      for (int g = 0; g <= a.bn_groups; ++g) p.bn_start[g] = a.bn_group_start[g];
      if (!a.bn_keep_sums) RPNET_CUDA_OK(cudaMemsetAsync(a.bn_sums, 0, (size_t)a.bn_groups * cout * 2 * sizeof(double), stream));
      if (a.bn_fused) *a.bn_fused = 1;
    }
  }
  if (a.out) {
    RPNET_REQUIRE(a.out_c % 8 == 0 && a.out_coff % 8 == 0 && a.out_coff + cout <= a.out_c, "conv_igemm: bad output channel window (%d + %d in %d)", a.out_coff, cout, a.out_c);
    RPNET_REQUIRE((h - 1) * a.oy_mul + a.oy_off < a.out_h && (w - 1) * a.ox_mul + a.ox_off < a.out_w && a.oy_off >= 0 && a.ox_off >= 0,
                  "conv_igemm: output mapping exceeds the %d x %d output", a.out_h, a.out_w);
  }

  // halo variant: a full 3x3 tap set (any order / sign: forward and data-gradient packs) on 16 x 8 pixel tiles, cout tiles 64 / 128
  bool halo = ntaps == 9 && !a.tap_src && bw == 16 && bh == 8 && bn == 1 && h >= bh + 2 && BN <= 128 && !a.res && !getenv("RPNET_CONV_NO_HALO");
  if (halo) {
    for (int i = 0; i < 9; ++i) p.halo_tap[i] = -1;
    for (int t = 0; t < 9; ++t) {
      const int dy = a.tap_dy[t], dx = a.tap_dx[t];
      if (dy < -1 || dy > 1 || dx < -1 || dx > 1) { halo = false; break; }
      p.halo_tap[(dy + 1) * 3 + dx + 1] = t;
    }
    for (int i = 0; i < 9 && halo; ++i) halo = p.halo_tap[i] >= 0;
    if (BN == 128 && p.tiles_x * p.tiles_y * p.tiles_n < 2) halo = false;
  }
  CUtensorMap t0, t1, t2, t3, tw;
  const uint32_t abox[4] = {(uint32_t)kBK, (uint32_t)bw, (uint32_t)(halo ? bh + 2 : bh), (uint32_t)bn};
  if (a.tap_src) {
    // four strided views of one tensor (the parity phases of a 2x up-sampled map): same dims, custom pixel strides
    CUtensorMap* tv[4] = {&t0, &t1, &t2, &t3};
    const uint64_t dims[4] = {(uint64_t)c0, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t str[3] = {(uint64_t)a.view_strides[0], (uint64_t)a.view_strides[1], (uint64_t)a.view_strides[2]};
    for (int v = 0; v < 4; ++v) {
      int rc = make_tmap_2b(tv[v], a.view_ptr[v], 4, dims, str, abox, bf16);
      if (rc) return rc;
    }
  } else {
    auto dense = [&](CUtensorMap* m, const void* ptr, int c) {
      const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
      const uint64_t str[3] = {(uint64_t)c, (uint64_t)c * w, (uint64_t)c * w * h};
      return make_tmap_2b(m, ptr, 4, dims, str, abox, bf16);
    };
    int rc = dense(&t0, a.src0, c0);
    if (rc) return rc;
    if (c1 > 0) { rc = dense(&t1, a.src1, c1); if (rc) return rc; } else t1 = t0;
    if (a_split) {
      rc = dense(&t2, a.src0_lo, c0);
      if (rc) return rc;
      if (c1 > 0) { rc = dense(&t3, a.src1_lo, c1); if (rc) return rc; } else t3 = t2;
    } else {
      t2 = t0;
      t3 = t0;
    }
  }
  const bool ws = !halo && cout == 64 && !a_split && !a.w_split && ntaps * p.kblocks_per_tap <= kWsMaxKb &&
                  p.tiles_x * p.tiles_y * p.tiles_n >= 4 * num_sms() && !getenv("RPNET_CONV_NO_WS");
  const bool pair = halo ? BN == 128 : (!ws && pair_enabled(BN) && p.tiles_x * p.tiles_y * p.tiles_n >= 2);
  {
    const uint64_t cin = (uint64_t)(c0 + c1) * (a.w_split ? 2 : 1);
    const uint64_t dims[3] = {cin, (uint64_t)cout, (uint64_t)ntaps};
    const uint64_t str[2] = {cin, cin * cout};
    const uint32_t box[3] = {(uint32_t)kBK, (uint32_t)(pair ? BN / 2 : BN), 1};
    int rc = make_tmap_2b(&tw, a.wpack, 3, dims, str, box, bf16);
    if (rc) return rc;
  }
  if (halo && BN == 64) {
    // weights-stationary when the whole pack fits beside the activation ring and there are enough tiles to amortise the preload
    const int wchunks = (int)((uint64_t)(c0 + c1) * (a.w_split ? 2 : 1) / kBK);
    const int nkb = 3 * p.kblocks_per_tap;
    if (cout == 64 && c1 == 0 && 9 * wchunks <= kWsHaloMaxKb && nkb <= kMaxWsHaloKb && p.tiles_x * p.tiles_y * p.tiles_n >= 4 * num_sms() &&
        !getenv("RPNET_CONV_NO_WS")) {
      p.ws_wchunks = wchunks;
      int kb = 0;                                          // replay the producer's stage order
      for (int ph = p.kb8_per_tap ? 1 : 0; ph >= 0; --ph)
        for (int dxi = 0; dxi < 3; ++dxi)
          for (int sg = 0; sg < p.nseg; ++sg) {
            if (p.seg_kind[sg] != ph) continue;
            for (int j = 0; j < p.seg_n[sg]; ++j, ++kb) { p.ws_dx[kb] = dxi; p.ws_wch[kb] = p.seg_wch[sg] + j; }
          }
      return launch_ws_halo64(t0, t1, t2, t3, tw, p, stream);
    }
  }
  if (halo) return BN == 128 ? launch_halo128_pair(t0, t1, t2, t3, tw, p, stream) : launch_halo64(t0, t1, t2, t3, tw, p, stream);
  if (pair) {
    switch (BN) {
      case 256: return launch_pair<256>(t0, t1, t2, t3, tw, p, stream);
      case 128: return launch_pair<128>(t0, t1, t2, t3, tw, p, stream);
      default:  return launch_pair<64>(t0, t1, t2, t3, tw, p, stream);
    }
  }
  if (ws) return launch_ws(t0, t1, t2, t3, tw, p, stream);
  switch (BN) {
    case 256: return launch<256>(t0, t1, t2, t3, tw, p, stream);
    case 128: return launch<128>(t0, t1, t2, t3, tw, p, stream);
    default:  return launch<64>(t0, t1, t2, t3, tw, p, stream);
  }
}

// the argument block shared by the plain entry points
static ConvCall plain_call(bool bf16, const void* src0, int c0, const void* src1, int c1, int n, int h, int w, const void* wpack,
                           int ntaps, const int* tap_dy, const int* tap_dx, int cout, const float* scale, const float* shift, int relu,
                           void* stream) {
  ConvCall a;
  a.bf16 = bf16; a.src0 = src0; a.c0 = c0; a.src1 = src1; a.c1 = c1; a.n = n; a.h = h; a.w = w; a.wpack = wpack; a.ntaps = ntaps;
  a.tap_dy = tap_dy; a.tap_dx = tap_dx; a.cout = cout; a.scale = scale; a.shift = shift; a.relu = relu; a.stream = stream;
  a.out_h = h; a.out_w = w; a.out_c = cout;
  return a;
}

// See include/rpnet_b200.h for the contract.
RPNET_API int rpnet_conv_igemm_f16(const void* src0, int c0, const void* src1, int c1, int n, int h, int w,
                                    const void* wpack, int ntaps, const int* tap_dy, const int* tap_dx, int cout,
                                    const float* scale, const float* shift, int relu, void* out_f16, int out_h, int out_w,
                                    int out_c, int out_coff, int oy_mul, int oy_off, int ox_mul, int ox_off,
                                    void* out_pool_f16, float* out_f32, void* stream_) {
  ConvCall a = plain_call(false, src0, c0, src1, c1, n, h, w, wpack, ntaps, tap_dy, tap_dx, cout, scale, shift, relu, stream_);
  a.out = out_f16; a.out_h = out_h; a.out_w = out_w; a.out_c = out_c; a.out_coff = out_coff;
  a.oy_mul = oy_mul; a.oy_off = oy_off; a.ox_mul = ox_mul; a.ox_off = ox_off;
  a.out_pool = out_pool_f16; a.out_f32 = out_f32;
  return conv_igemm_run(a);
}

RPNET_API int rpnet_conv_igemm_bf16(const void* src0, int c0, const void* src1, int c1, int n, int h, int w,
                                     const void* wpack, int ntaps, const int* tap_dy, const int* tap_dx, int cout,
                                     const float* scale, const float* shift, int relu, void* out_bf16, int out_h, int out_w,
                                     int out_c, int out_coff, int oy_mul, int oy_off, int ox_mul, int ox_off,
                                     void* out_pool_bf16, float* out_f32, void* stream_) {
  ConvCall a = plain_call(true, src0, c0, src1, c1, n, h, w, wpack, ntaps, tap_dy, tap_dx, cout, scale, shift, relu, stream_);
  a.out = out_bf16; a.out_h = out_h; a.out_w = out_w; a.out_c = out_c; a.out_coff = out_coff;
  a.oy_mul = oy_mul; a.oy_off = oy_off; a.ox_mul = ox_mul; a.ox_off = ox_off;
  a.out_pool = out_pool_bf16; a.out_f32 = out_f32;
  return conv_igemm_run(a);
}

// See include/rpnet_b200.h for the contract.
RPNET_API int rpnet_conv_split_f16(const void* src0_hi, const void* src0_lo, int c0, const void* src1_hi, const void* src1_lo, int c1,
                                    int n, int h, int w, const void* wpack, int w_split, int ntaps, const int* tap_dy, const int* tap_dx,
                                    int cout, const float* scale, const float* shift, int relu, void* out_hi, void* out_lo, int out_h,
                                    int out_w, int out_c, int out_coff, int oy_mul, int oy_off, int ox_mul, int ox_off,
                                    void* out_pool_hi, void* out_pool_lo, float* out_f32, const int* group_start, int groups,
                                    double* sums, int keep_sums, void* stream_) {
  return rpnet_conv_split_res_f16(src0_hi, src0_lo, c0, src1_hi, src1_lo, c1, n, h, w, wpack, w_split, ntaps, tap_dy, tap_dx, cout, scale, shift,
                                  nullptr, nullptr, relu, out_hi, out_lo, out_h, out_w, out_c, out_coff, oy_mul, oy_off, ox_mul, ox_off,
                                  out_pool_hi, out_pool_lo, out_f32, group_start, groups, sums, keep_sums, stream_);
}

// See include/rpnet_b200.h for the contract.
RPNET_API int rpnet_conv_split_res_f16(const void* src0_hi, const void* src0_lo, int c0, const void* src1_hi, const void* src1_lo, int c1,
                                        int n, int h, int w, const void* wpack, int w_split, int ntaps, const int* tap_dy,
                                        const int* tap_dx, int cout, const float* scale, const float* shift, const void* res_hi,
                                        const void* res_lo, int relu, void* out_hi, void* out_lo, int out_h, int out_w, int out_c,
                                        int out_coff, int oy_mul, int oy_off, int ox_mul, int ox_off, void* out_pool_hi,
                                        void* out_pool_lo, float* out_f32, const int* group_start, int groups, double* sums,
                                        int keep_sums, void* stream_) {
  ConvCall a = plain_call(false, src0_hi, c0, src1_hi, c1, n, h, w, wpack, ntaps, tap_dy, tap_dx, cout, scale, shift, relu, stream_);
  a.res = res_hi; a.res_lo = res_lo;
  a.src0_lo = src0_lo; a.src1_lo = src1_lo; a.w_split = w_split;
  a.out = out_hi; a.out_lo = out_lo; a.out_h = out_h; a.out_w = out_w; a.out_c = out_c; a.out_coff = out_coff;
  a.oy_mul = oy_mul; a.oy_off = oy_off; a.ox_mul = ox_mul; a.ox_off = ox_off;
  a.out_pool = out_pool_hi; a.out_pool_lo = out_pool_lo; a.out_f32 = out_f32;
  if (!sums) return conv_igemm_run(a);
  // train mode: BatchNorm statistics of the fp32 accumulators in the epilogue (dense identity-mapped or sub-pixel output)
  RPNET_REQUIRE(out_hi && group_start, "conv_split: the BatchNorm statistics need the z output and the call groups");
  int fused = 0;
  a.bn_group_start = group_start; a.bn_groups = groups; a.bn_sums = sums; a.bn_fused = &fused; a.bn_keep_sums = keep_sums;
  int rc = conv_igemm_run(a);
  if (rc) return rc;
  if (!fused) {
    RPNET_REQUIRE(oy_mul == 1 && ox_mul == 1 && out_coff == 0 && out_c == cout && !keep_sums,
                  "conv_split: BatchNorm statistics cannot be fused for this %d x %d sub-pixel launch", h, w);
    return out_lo ? rpnet_bn_stats_split_f16(out_hi, out_lo, n, h, w, cout, group_start, groups, sums, stream_)
                  : rpnet_bn_stats_f16(out_hi, n, h, w, cout, group_start, groups, sums, stream_);
  }
  return 0;
}

// See include/rpnet_b200.h for the contract.
RPNET_API int rpnet_conv_bnstats_f16(const void* src0, int c0, const void* src1, int c1, int n, int h, int w, const void* wpack,
                                      int ntaps, const int* tap_dy, const int* tap_dx, int cout, const float* ones,
                                      const float* zeros, void* z_f16, const int* group_start, int groups, double* sums,
                                      void* stream_) {
  RPNET_REQUIRE(z_f16 && sums && group_start, "conv_bnstats: null pointer argument");
  int fused = 0;
  ConvCall a = plain_call(false, src0, c0, src1, c1, n, h, w, wpack, ntaps, tap_dy, tap_dx, cout, ones, zeros, 0, stream_);
  a.out = z_f16;
  a.bn_group_start = group_start; a.bn_groups = groups; a.bn_sums = sums; a.bn_fused = &fused;
  int rc = conv_igemm_run(a);
  if (rc) return rc;
  if (!fused) return rpnet_bn_stats_f16(z_f16, n, h, w, cout, group_start, groups, sums, stream_);   // tiny maps: separate pass
  return 0;
}

// See include/rpnet_b200.h for the contract.
RPNET_API int rpnet_conv_cos_f16(const void* src0, int c0, const void* src1, int c1, int n, int h, int w, const void* wpack, int ntaps,
                                  const int* tap_dy, const int* tap_dx, const float* scale, const float* shift, int relu,
                                  const float* protos, int n_protos, int proto_sets, float scaler, float* pred, float* out_f32,
                                  void* stream_) {
  RPNET_REQUIRE(protos && pred, "conv_cos: null pointer argument");
  ConvCall a = plain_call(false, src0, c0, src1, c1, n, h, w, wpack, ntaps, tap_dy, tap_dx, 64, scale, shift, relu, stream_);
  a.out_f32 = out_f32;
  a.cos_protos = protos; a.cos_P = n_protos; a.cos_sets = proto_sets; a.cos_scaler = scaler; a.cos_pred = pred;
  return conv_igemm_run(a);
}

// ---------------------------------------------------------------------------------------------------
// up_conv in sub-pixel form (see conv_wgrad.cu for the tap tables), train mode.
// ---------------------------------------------------------------------------------------------------
// One output parity phase of z = conv3x3(upsample2x(x)): a 2x2-tap conv of the low-resolution x [n][h][w][cin] whose pixels
// land at (2y + py, 2x + px) of z [n][2h][2w][cout]; the BatchNorm statistics of z accumulate over the four phase launches
// (keep_sums = 0 on the first).  Returns -2 when the statistics cannot be fused for this shape (callers then materialise the
// up-sampled map and run rpnet_conv_bnstats_f16).
RPNET_API int rpnet_upconv_phase_bnstats_f16(const void* x_low, int cin, int n, int h, int w, const void* wphase /*[4][cout][cin]*/,
                                              int py, int px, int cout, const float* ones, const float* zeros, void* z_f16,
                                              const int* group_start, int groups, double* sums, int keep_sums, void* stream_) {
  RPNET_REQUIRE(x_low && wphase && z_f16 && sums && group_start && (py | 1) == 1 && (px | 1) == 1, "upconv_phase: bad argument");
  int dy[4], dx[4];
  for (int t = 0; t < 4; ++t) {
    dy[t] = (py == 0 ? -1 : 0) + (t >> 1);
    dx[t] = (px == 0 ? -1 : 0) + (t & 1);
  }
  int fused = 0;
  ConvCall a = plain_call(false, x_low, cin, nullptr, 0, n, h, w, wphase, 4, dy, dx, cout, ones, zeros, 0, stream_);
  a.out = z_f16; a.out_h = 2 * h; a.out_w = 2 * w; a.oy_mul = 2; a.oy_off = py; a.ox_mul = 2; a.ox_off = px;
  a.bn_group_start = group_start; a.bn_groups = groups; a.bn_sums = sums; a.bn_fused = &fused; a.bn_keep_sums = keep_sums;
  int rc = conv_igemm_run(a);
  if (rc) return rc;
  if (!fused) {
    set_error("upconv_phase: BatchNorm statistics cannot be fused for %d x %d maps (tile straddles call groups)", h, w);
    return RPNET_ERR_ARG;
  }
  return 0;
}

// Data gradient of the same op w.r.t. the LOW-resolution input: a 4x4 stride-2 conv of dZ [n][2h][2w][cout] (bf16),
//   dx[y, x, ci] = sum_{oy, ox in -1..2} sum_co dZ[2y + oy, 2x + ox, co] * W4[oy][ox][ci][co],   W4 = row/column sums of w,
// run as 16 taps over the four parity views of dZ (custom-stride tensor maps; out-of-bounds = zero).
// w16: bf16 [16][cin][cout] with tap index (oy + 1) * 4 + (ox + 1).  out: bf16 [n][h][w][out_c], channels [out_coff, +cin).
RPNET_API int rpnet_upconv_dgrad_bf16(const void* dz, int cout, int n, int h, int w, const void* w16, int cin, void* out, int out_c,
                                       int out_coff, const float* ones, const float* zeros, void* stream_) {
  RPNET_REQUIRE(dz && w16 && out, "upconv_dgrad: null pointer argument");
  int dy[16], dx[16], src[16];
  for (int t = 0; t < 16; ++t) {
    const int oy = t / 4 - 1, ox = t % 4 - 1;
    const int py = oy & 1, px = ox & 1;
    dy[t] = (oy - py) / 2;
    dx[t] = (ox - px) / 2;
    src[t] = py * 2 + px;
  }
  const void* views[4];
  for (int v = 0; v < 4; ++v)
    views[v] = static_cast<const __nv_bfloat16*>(dz) + ((size_t)(v >> 1) * (2 * w) + (v & 1)) * cout;
  const long long strides[3] = {2LL * cout, 2LL * (2 * w) * cout, (long long)(2 * h) * (2 * w) * cout};
  ConvCall a = plain_call(true, views[0], cout, nullptr, 0, n, h, w, w16, 16, dy, dx, cin, ones, zeros, 0, stream_);
  a.out = out; a.out_c = out_c; a.out_coff = out_coff;
  a.view_ptr = views; a.view_strides = strides; a.tap_src = src;
  return conv_igemm_run(a);
}

// See include/rpnet_b200.h for the contract.
RPNET_API int rpnet_conv_res_f16(const void* src, int cin, int n, int h, int w, const void* wpack, int ntaps, const int* tap_dy,
                                  const int* tap_dx, int cout, const float* scale, const float* shift, const void* res_f16, int relu,
                                  void* out_f16, void* stream_) {
  RPNET_REQUIRE(out_f16, "conv_res: null output");
  ConvCall a = plain_call(false, src, cin, nullptr, 0, n, h, w, wpack, ntaps, tap_dy, tap_dx, cout, scale, shift, relu, stream_);
  a.out = out_f16;
  a.res = res_f16;
  return conv_igemm_run(a);
}
