// Common device helpers for the sgdm_b200 kernels (sm_100a only).
//
// Operand precision: GEMM / conv / attention operands are 16-bit floats with fp32
// accumulation.  The default operand type is IEEE fp16: the CPU forecast in
// DESIGN.md ("operand precision") shows bf16 operands miss the north-star tolerance
// (guided eps rel-L2 1.2e-2..1.3e-2 > 1e-2) while fp16 operands give 1.6e-3.  Build
// with -DSGDM_OPERAND_BF16 to get bf16 operands instead (same tcgen05 rate).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#ifdef SGDM_OPERAND_BF16
typedef __nv_bfloat16 op_t;
typedef __nv_bfloat162 op2_t;
#define SGDM_UMMA_FMT 1u /* InstrDescriptor a/b_format: BF16 */
#else
typedef __half op_t;
typedef __half2 op2_t;
#define SGDM_UMMA_FMT 0u /* F16 */
#endif

#include <stdlib.h>

#include <utility>

namespace sgdm {

constexpr int kNumSMs = 148;

// ---- programmatic dependent launch ------------------------------------------------------
// The ~170 kernels of a forward are a strict chain on one stream.  Kernels launched through launch_pdl() carry
// cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may be scheduled while the preceding kernel drains
// (it calls pdl_launch_dependents() first thing), run their private set-up (barrier init, TMEM allocation, tensor-map
// prefetch) and then block in pdl_wait() until the preceding grid has completed and its writes are visible.
// Rules kept by every such kernel: nothing global is read or written before pdl_wait(); every CTA executes pdl_wait()
// (so a grid can never complete before its predecessor: the guarantee is transitive along the chain).
// Both instructions are no-ops in a kernel launched without the attribute.
// Measured on B200: config 1 (batch 16, 32x32: ~170 launches of 5-10 us) 2.245 -> 1.948 ms per guided step; config 2 at
// batch 256 (launches of 0.1-2 ms, one persistent CTA per SM) 48.3 -> 49.0 ms.  So the attribute is a per-replay
// decision of the engine (pdl_mode(): small plans only); SGDM_PDL=0 / 1 forces it off / on everywhere.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
inline bool& pdl_mode() {
  static bool on = false;  // set by the engine around each plan replay; single-kernel calls keep it off
  return on;
}
// (Also measured, round 2: the attribute on the GroupNorm finalise / apply launches only, in every plan — their small
//  CTAs become resident behind the running conv: config 2 at batch 256 50.2 / 50.4 -> 51.0 / 51.0 ms.  Not kept.)
inline bool pdl_enabled() {
  static const int forced = getenv("SGDM_PDL") ? (atoi(getenv("SGDM_PDL")) != 0 ? 1 : 0) : -1;
  return forced >= 0 ? forced == 1 : pdl_mode();
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- operand conversions (saturating: fp16 overflow must never make an inf) --------
__device__ __forceinline__ op_t to_op(float v) {
#ifdef SGDM_OPERAND_BF16
  return __float2bfloat16_rn(v);
#else
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return __ushort_as_half(r);
#endif
}
__device__ __forceinline__ float from_op(op_t v) {
#ifdef SGDM_OPERAND_BF16
  return __bfloat162float(v);
#else
  return __half2float(v);
#endif
}
__device__ __forceinline__ uint32_t pack_op2(float lo, float hi) {
#ifdef SGDM_OPERAND_BF16
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
#else
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
#endif
}
__device__ __forceinline__ float2 unpack_op2(uint32_t v) {
#ifdef SGDM_OPERAND_BF16
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
#else
  return __half22float2(*reinterpret_cast<__half2*>(&v));
#endif
}

// x * sigmoid(x) with one MUFU.EX2 and one MUFU.RCP (an IEEE divide would make the GroupNorm
// apply pass issue-bound instead of HBM-bound); |error| <= ~2 ulp, far below the 16-bit output rounding.
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// Four SiLUs with ONE reciprocal: 1 / d_i = (1 / (d_0 d_1 d_2 d_3)) * prod_{j != i} d_j, d_i = 1 + 2^(-x_i log2 e).
// 1.25 MUFU operations per element instead of 2: the GroupNorm apply pass with a 16-bit source moves 4 B per element
// and was bound by the 16-lane MUFU at the power-capped clocks (4.7 instead of 6.1 TB/s).  The exponent is clamped at
// 2^30 so that the product of four stays finite: for x < -20.8 the result is x * 2^-30 instead of ~x e^x, |error| < 2e-8 |x|.
__device__ __forceinline__ void silu4(const float (&x)[4], float (&y)[4]) {
  float d[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(x[i] * -1.4426950408889634f, 30.0f)));
    d[i] = 1.0f + e;
  }
  const float p01 = d[0] * d[1], p23 = d[2] * d[3];
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p01 * p23));
  const float r01 = r * p23, r23 = r * p01;  // 1 / (d0 d1), 1 / (d2 d3)
  y[0] = x[0] * (r01 * d[1]);
  y[1] = x[1] * (r01 * d[0]);
  y[2] = x[2] * (r23 * d[3]);
  y[3] = x[3] * (r23 * d[2]);
}

// ---- shared-memory address / mbarrier -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU
// (a hang on the shared B200 box costs a strike).  ~2^26 polls is many seconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}

// ---- TMA ------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* d) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(d)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* d, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(d)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* d, uint64_t* bar, void* dst, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(d)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// bulk prefetch of [p, p + bytes) into L2 (16-byte aligned, bytes % 16 == 0); no destination registers
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- tcgen05 / TMEM -------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tcgen05.commit: arrive on an mbarrier when all previously issued MMAs have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; 16-bit operands, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns of fp32: thread t of the warp gets lane (base_lane + t), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t addr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t addr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr)
      : "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 64 16-bit elements (128 B), 8-row
// swizzle atoms 1024 B apart (SBO), LBO unused, descriptor version 1 (sm_100),
// layout_type 2 = SWIZZLE_128B.  `addr` is the shared-space byte address of the tile
// (1024 B aligned) plus k*32 B for the k-th 16-element K slice inside the atom.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// The same split in two: the address-independent upper word for 128-byte rows / SWIZZLE_128B (SBO 1024, layout 2) or
// 64-byte rows / SWIZZLE_64B (SBO 512, layout 4; the K-major tile of a 32-element K block), and the address bits.
__device__ __forceinline__ uint64_t umma_smem_desc_hi(bool sw64) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((sw64 ? 512u : 1024u) >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(sw64 ? 4 : 2) << 61;
  return d;
}
__device__ __forceinline__ uint64_t umma_smem_desc_lo(uint32_t addr) { return static_cast<uint64_t>((addr >> 4) & 0x3FFFu); }
// kind::f16 instruction descriptor: fp32 accumulator, A/B 16-bit (fmt), both K-major, MxN.
__device__ __forceinline__ uint32_t umma_idesc(uint32_t m, uint32_t n) {
  return (1u << 4) | (SGDM_UMMA_FMT << 7) | (SGDM_UMMA_FMT << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---- CTA pair (cta_group::2): two CTAs of a cluster on one TPC issue one M=256 MMA -------------
// The shared::cluster address of a CTA's own shared memory carries the CTA rank in bit 24; clearing
// it addresses the same offset in the even ("leader") CTA of the pair.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the LEADER CTA's copy of `bar` (works from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
// TMA loads of a CTA pair: data lands in the executing CTA's shared memory, the transaction bytes are
// counted on the LEADER's mbarrier.
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* d, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(d)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* d, uint64_t* bar, void* dst, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(d)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// commit of the pair's MMAs: arrives on `bar` (same offset) in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 from each CTA's smem] * B[N rows: N/2 from each CTA's smem]
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

}  // namespace sgdm
