// Shared device/host helpers for the rpnet_b200 sm_100a kernels: error plumbing for the C ABI and
// thin inline-PTX wrappers (mbarrier, TMA, tcgen05/TMEM).  No CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define RPNET_API extern "C" __attribute__((visibility("default")))

namespace rpnet {

// ---- C-ABI error convention: 0 = ok, negative = error, message via rpnet_last_error() -------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
#define RPNET_CUDA_OK(expr)                                   \
  do {                                                        \
    int _rc = ::rpnet::check_cuda((expr), #expr);             \
    if (_rc) return _rc;                                      \
  } while (0)
#define RPNET_REQUIRE(cond, ...)                              \
  do {                                                        \
    if (!(cond)) {                                            \
      ::rpnet::set_error(__VA_ARGS__);                        \
      return -2;                                              \
    }                                                         \
  } while (0)

enum { RPNET_OK = 0, RPNET_ERR_CUDA = -1, RPNET_ERR_ARG = -2, RPNET_ERR_DRIVER = -3 };

// host helpers shared by the tensor-core kernels (defined in conv_igemm.cu)
// 2-byte (fp16 / bf16) tensor map with 128B swizzle; dims/box innermost first; strides in elements for dims 1..rank-1.
int make_tmap_2b(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                 const uint32_t* box, bool bf16);
int num_sms();
int pow2_floor(int v);
int ilog2(int v);

#ifdef __CUDACC__
// ---- generic -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- 8-wide packing: fp32 registers <-> one 16-byte vector of fp16 / bf16 ------------------------
__device__ __forceinline__ uint4 pack8_f16(const float* v) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
  return u;
}
__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
  return u;
}
__device__ __forceinline__ void unpack8_f16(const uint4& u, float* v) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
// split-fp16 representation x = hi + lo: the fp16 residual of 8 fp32 values against their rounded fp16 vector `hi`
__device__ __forceinline__ uint4 residual8_f16(const float* v, const uint4& hi) {
  float r[8];
  unpack8_f16(hi, r);
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = v[i] - r[i];
  return pack8_f16(r);
}
// ---- fp8 correction planes ("c8") --------------------------------------------------------------------
// The tensor-core convs of the default precision compute x.w = hi.Wh + 2^-15 (lo8.Wh8 + x8.Wl8): the main term on fp16
// operands, the two first-order corrections on e4m3 operands (twice the MMA rate; a correction is 2^-11 of the main term, so
// its own 2^-4 rounding lands at 2^-15), with fixed power-of-two scales
//   lo8 = e4m3((x - hi) * 2^9)   x8 = e4m3(x * 2^-2)   Wh8 = e4m3(Wh * 2^6)   Wl8 = e4m3((w - Wh) * 2^17)
// (saturating: exact corrections for |x| < 1792 and |w| < 7, beyond that the element falls back to single-term accuracy)
// so that both products carry 2^15, which tcgen05.mma's scale-input-d removes from the accumulator when the fp16 main term
// starts.  Layout of a c8 plane (same bytes as an fp16 plane of the same shape): per pixel and 64-channel group, 128 bytes
// = lo8 of the 64 channels, then x8 of the 64 channels — one 128-byte swizzle row, K = 128 for the e4m3 MMA.
constexpr float kC8LoScale = 512.f;           // 2^9
constexpr float kC8XScale = 0.25f;            // 2^-2
constexpr float kC8WhScale = 64.f;            // 2^6:  kC8LoScale * kC8WhScale = 2^15
constexpr float kC8WlScale = 131072.f;        // 2^17: kC8XScale * kC8WlScale = 2^15
constexpr int kC8AccShift = 15;               // scale-input-d of the first main-term MMA
__device__ __forceinline__ uint32_t pack4_e4m3(float a, float b, float c, float d) {
  const uint32_t l = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const uint32_t h = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return l | (h << 16);
}
__device__ __forceinline__ uint2 pack8_e4m3(const float* v) {
  return make_uint2(pack4_e4m3(v[0], v[1], v[2], v[3]), pack4_e4m3(v[4], v[5], v[6], v[7]));
}
__device__ __forceinline__ void unpack8_e4m3(const uint2& u, float* v) {
  const uint32_t w[2] = {u.x, u.y};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2_raw r = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)((w[i >> 1] >> (16 * (i & 1))) & 0xffffu), __NV_E4M3);
    const float2 f = __half22float2(__half2(r));
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
// byte address of the lo8 entry of the channel at fp16-element offset `off` (= pixel * C + channel) of a c8 plane; `cg` = that
// channel's index inside its 64-channel group; the x8 entry sits 64 bytes further
__device__ __forceinline__ uint8_t* c8_addr(void* plane, size_t off, int cg) {
  return static_cast<uint8_t*>(plane) + 2 * off - cg;
}
__device__ __forceinline__ const uint8_t* c8_addr(const void* plane, size_t off, int cg) {
  return static_cast<const uint8_t*>(plane) + 2 * off - cg;
}
// the c8 entries of 8 consecutive channels (fp32 values v, `hi` = their fp16 rounding) at dst = c8_addr(...)
__device__ __forceinline__ void c8_store8(uint8_t* dst, const float* v, const uint4& hi) {
  float r[8];
  unpack8_f16(hi, r);
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = (v[i] - r[i]) * kC8LoScale;
  *reinterpret_cast<uint2*>(dst) = pack8_e4m3(r);
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = v[i] * kC8XScale;
  *reinterpret_cast<uint2*>(dst + 64) = pack8_e4m3(r);
}
// lo plane of 8 consecutive channels, whichever format: r[i] += x - hi (c8: lo8 * 2^-9)
__device__ __forceinline__ void lo8_add(const void* plane, int c8, size_t off, int cg, float* r) {
  float l[8];
  if (c8) {
    unpack8_e4m3(__ldg(reinterpret_cast<const uint2*>(c8_addr(plane, off, cg))), l);
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = fmaf(l[i], 1.f / kC8LoScale, r[i]);
  } else {
    unpack8_f16(__ldg(reinterpret_cast<const uint4*>(static_cast<const __half*>(plane) + off)), l);
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] += l[i];
  }
}
// write the lo plane of 8 consecutive channels, whichever format
__device__ __forceinline__ void lo8_store(void* plane, int c8, size_t off, int cg, const float* v, const uint4& hi) {
  if (c8) c8_store8(c8_addr(plane, off, cg), v, hi);
  else *reinterpret_cast<uint4*>(static_cast<__half*>(plane) + off) = residual8_f16(v, hi);
}
__device__ __forceinline__ void unpack8_bf16(const uint4& u, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P;\n"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n\t"
      "@P bra DONE;\n\t"
      "bra WAIT_LOOP;\n"
      "DONE:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- TMA -----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- tcgen05 / TMEM ------------------------------------------------------------------------------
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, fp16/bf16 operands, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// e4m3 operands (kind::f8f6f4, K = 32 per instruction), same descriptors and fp32 accumulator
__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D = A * B + D * 2^-15 (scale-input-d): the first main-term MMA after the 2^15-scaled e4m3 corrections
__device__ __forceinline__ void umma_f16_sd15(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 15;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc)
      : "memory");
}
// mbarrier arrive once all previously issued MMAs of this thread have completed (implies fence::before).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// ---- CTA pair (cta_group::2): two CTAs of a cluster on one TPC execute one M = 256 MMA; each holds its 128 accumulator rows in
// its own TMEM and supplies its own A rows and half of the B rows from its own shared memory.  Only the leader (cluster rank 0)
// issues MMAs and commits; both CTAs issue their TMA loads against the LEADER's mbarrier.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in cluster rank 0
__device__ __forceinline__ void mbar_arrive_rank0(uint64_t* bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(bar)));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P;\n"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%0], %1;\n\t"
      "@P bra DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n"
      "DONE_C:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // shared::cluster address of a barrier -> the same barrier in the even CTA of the pair
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem) {  // one full warp of EACH CTA of the pair (same warp index)
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair_sd15(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p, 15;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in every CTA of `cta_mask` once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t <-> TMEM lane base+t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 128 B, 8-row groups
// `sbo_bytes` apart (1024 B when rows are dense).  Fields per the sm_100 descriptor format:
// [0,14) addr>>4, [16,30) LBO>>4 (unused for swizzled K-major, 1), [32,46) SBO>>4, [46,48) version=1,
// [61,64) layout (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) |
         (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major operand (the M/N index is contiguous: 64 elements = one 128-byte swizzle row per K index), 128-byte swizzle.
// Canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: `lbo_bytes` = distance between consecutive
// 64-element chunks along M/N, `sbo_bytes` = distance between groups of 8 K rows (1024 B when rows are dense).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor: fp16 x fp16 -> fp32, both operands K-major, dense.
// Modifiers: bit 7 / bit 10 = A / B operand is bf16; bit 15 / bit 16 = A / B operand is MN-major.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
#endif  // __CUDACC__

}  // namespace rpnet
