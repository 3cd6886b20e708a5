// HBM-bound kernels: GroupNorm(+FiLM+SiLU+resample), LayerNorm, casts, the fp32 prologue
// (timestep embedding, null substitution, small MLPs), the fused guidance-mix + sampler
// update, uint8 conversion and weight packing.  See kernels.cuh for the reference lines.
#include <stdlib.h>

#include "kernels.cuh"

namespace sgdm {

#define SGDM_LAUNCH_OK() (cudaGetLastError() == cudaSuccess ? 0 : 1)

// =========================================================================== GroupNorm
// Two HBM-bound passes (roofline: read 4 B + read 4 B + write 2 B per element, 2+2+2 for a
// 16-bit source).  Both passes give every thread a FIXED channel slice and walk pixels with
// several independent 128-bit loads in flight; all reduction orders depend on the per-sample
// shape only, so results are bit-identical however samples are batched.
//
// Pass 1: per (sample, chunk of pixels) partial sum / sum-of-squares per group.
// Thread = (4-channel column q, pixel lane pl); fp32 per-thread partials over at most a few
// hundred elements, combined in double -> deterministic and cancellation-safe.
template <bool kHalfIn>
__device__ __forceinline__ float4 gn_ld4(const void* base, long off) {
  if (kHalfIn) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(static_cast<const op_t*>(base) + off));
    const float2 a = unpack_op2(r.x), b = unpack_op2(r.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(base) + off));
}

template <bool kHalfIn>
__global__ void __launch_bounds__(256) gn_stats_kernel(const void* __restrict__ src0, const float* __restrict__ src1,
                                                       int HW, int C0, int C1, int chunks, int PL,
                                                       double* __restrict__ partial) {
  __shared__ float s_sum[1024];
  __shared__ float s_sq[1024];
  const int C = C0 + C1, C4 = C >> 2;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int q = threadIdx.x % C4, pl = threadIdx.x / C4;
  const int ppc = HW / chunks;
  const int c = q * 4;
  const bool first = c < C0;
  const long cs = first ? C0 : C1;
  const long base_off = static_cast<long>(n) * HW * cs + (first ? c : c - C0);
  const void* base = first ? src0 : static_cast<const void*>(src1);
  float4 s[4], ss[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) { s[u] = make_float4(0.f, 0.f, 0.f, 0.f); ss[u] = s[u]; }
  const int px_end = (chunk + 1) * ppc;
  int px = chunk * ppc + pl;
  for (; px + 3 * PL < px_end; px += 4 * PL) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      v[u] = (first && kHalfIn) ? gn_ld4<true>(base, base_off + static_cast<long>(px + u * PL) * cs)
                                : gn_ld4<false>(base, base_off + static_cast<long>(px + u * PL) * cs);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[u].x += v[u].x; s[u].y += v[u].y; s[u].z += v[u].z; s[u].w += v[u].w;
      ss[u].x += v[u].x * v[u].x; ss[u].y += v[u].y * v[u].y; ss[u].z += v[u].z * v[u].z; ss[u].w += v[u].w * v[u].w;
    }
  }
  for (; px < px_end; px += PL) {
    const float4 v = (first && kHalfIn) ? gn_ld4<true>(base, base_off + static_cast<long>(px) * cs)
                                        : gn_ld4<false>(base, base_off + static_cast<long>(px) * cs);
    s[0].x += v.x; s[0].y += v.y; s[0].z += v.z; s[0].w += v.w;
    ss[0].x += v.x * v.x; ss[0].y += v.y * v.y; ss[0].z += v.z * v.z; ss[0].w += v.w * v.w;
  }
  // layout [pl][C]; the 4 unroll slots are folded in a fixed order
  float* ps = s_sum + pl * C + c;
  float* pq = s_sq + pl * C + c;
  ps[0] = (s[0].x + s[1].x) + (s[2].x + s[3].x); ps[1] = (s[0].y + s[1].y) + (s[2].y + s[3].y);
  ps[2] = (s[0].z + s[1].z) + (s[2].z + s[3].z); ps[3] = (s[0].w + s[1].w) + (s[2].w + s[3].w);
  pq[0] = (ss[0].x + ss[1].x) + (ss[2].x + ss[3].x); pq[1] = (ss[0].y + ss[1].y) + (ss[2].y + ss[3].y);
  pq[2] = (ss[0].z + ss[1].z) + (ss[2].z + ss[3].z); pq[3] = (ss[0].w + ss[1].w) + (ss[2].w + ss[3].w);
  __syncthreads();
  if (threadIdx.x < 32) {
    const int g = threadIdx.x, cpg = C / 32;
    double a = 0.0, b = 0.0;
    for (int l = 0; l < PL; ++l)
      for (int k = 0; k < cpg; ++k) {
        a += static_cast<double>(s_sum[l * C + g * cpg + k]);
        b += static_cast<double>(s_sq[l * C + g * cpg + k]);
      }
    double* out = partial + ((static_cast<long>(n) * chunks + chunk) * 32 + g) * 2;
    out[0] = a;
    out[1] = b;
  }
}

// 16-bit source variant: thread = (8-channel column, pixel lane), 16-byte loads.
__global__ void __launch_bounds__(256) gn_stats_op_kernel(const op_t* __restrict__ src, int HW, int C, int chunks, int PL,
                                                          double* __restrict__ partial) {
  __shared__ float s_sum[2048];
  __shared__ float s_sq[2048];
  const int C8 = C >> 3;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int q = threadIdx.x % C8, pl = threadIdx.x / C8;
  const int ppc = HW / chunks;
  const int c = q * 8;
  const op_t* base = src + static_cast<long>(n) * HW * C + c;
  float s[2][8], ss[2][8];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[u][j] = 0.f; ss[u][j] = 0.f; }
  auto acc8 = [&](const uint4& r, float (&a)[8], float (&b)[8]) {
    const float2 p0 = unpack_op2(r.x), p1 = unpack_op2(r.y), p2 = unpack_op2(r.z), p3 = unpack_op2(r.w);
    const float v[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] += v[j]; b[j] += v[j] * v[j]; }
  };
  const int px_end = (chunk + 1) * ppc;
  int px = chunk * ppc + pl;
  for (; px + 3 * PL < px_end; px += 4 * PL) {
    uint4 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) r[u] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long>(px + u * PL) * C));
#pragma unroll
    for (int u = 0; u < 4; ++u) acc8(r[u], s[u & 1], ss[u & 1]);
  }
  for (; px < px_end; px += PL) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long>(px) * C));
    acc8(r, s[0], ss[0]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_sum[pl * C + c + j] = s[0][j] + s[1][j];
    s_sq[pl * C + c + j] = ss[0][j] + ss[1][j];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int g = threadIdx.x, cpg = C / 32;
    double a = 0.0, b = 0.0;
    for (int l = 0; l < PL; ++l)
      for (int k = 0; k < cpg; ++k) {
        a += static_cast<double>(s_sum[l * C + g * cpg + k]);
        b += static_cast<double>(s_sq[l * C + g * cpg + k]);
      }
    double* out = partial + ((static_cast<long>(n) * chunks + chunk) * 32 + g) * 2;
    out[0] = a;
    out[1] = b;
  }
}

struct GnApplyArgs {
  const void* src0; const void* src1;
  int H, W, C0, C1;
  const float* gamma; const float* beta; const float* film; long film_stride;
  int silu, resample, chunks, ppb, PLa;
  const double* partial;
  const float2* final;
  op_t* out; op_t* raw_out; float* pool_out;
  int mod0, mod1;  // source batch modulo (GnDesc::src_mod0 / src_mod1), 0 = none
};

// (Measured, round 2: doing this reduction in the prologue of every gn_apply block instead of a separate launch —
//  49 launches of ~14 us per step — costs more than it saves at batch 256: gn_apply + finalise 10.2 -> 11.4 ms, the
//  two dependent L2 round trips sit in front of every block's first load; config 1 gains 0.13 ms.  Not kept.)
// Fused-statistics path: reduce the conv epilogue's partial sums (per 32-row block and channel
// granule, see ConvDesc::stats) to {mean, rstd} per (sample, group).  One block per sample, one thread
// per channel granule walking the sample's row blocks (consecutive threads read consecutive float2:
// coalesced), double accumulation, then the granules of a group are added in order.  The summation
// order depends on the sample's shape only (batch-invariant bits).
__global__ void __launch_bounds__(512) gn_finalize_kernel(const float2* __restrict__ st0, int C0,
                                                          const float2* __restrict__ st1, int C1, int HW,
                                                          int gran, int RL, float2* __restrict__ out, int mod0, int mod1) {
  __shared__ double s_s[512], s_q[512];
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.x;
  const int C = C0 + C1, cpg = C / 32, epg = cpg / gran, rbs = HW >> 5;
  const int e0 = C0 / gran, e1 = C1 / gran, ne = e0 + e1;  // granules: ne * RL <= 512 threads
  // thread = (row-block lane rl, granule cg): consecutive threads read consecutive float2 of one row block
  const int cg = threadIdx.x % ne, rl = threadIdx.x / ne;
  {
    const bool first = cg < e0;
    const int ns = first ? (mod0 ? n % mod0 : n) : (mod1 ? n % mod1 : n);  // source sample (shared CFG prefix)
    const float2* p = first ? st0 + static_cast<long>(ns) * rbs * e0 + cg : st1 + static_cast<long>(ns) * rbs * e1 + (cg - e0);
    const long ld = first ? e0 : e1;
    double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
    int rb = rl;
    for (; rb + 3 * RL < rbs; rb += 4 * RL) {
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(p + (rb + u * RL) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) { s[u] += static_cast<double>(v[u].x); q[u] += static_cast<double>(v[u].y); }
    }
    for (; rb < rbs; rb += RL) {
      const float2 v = __ldg(p + rb * ld);
      s[0] += static_cast<double>(v.x);
      q[0] += static_cast<double>(v.y);
    }
    s_s[rl * ne + cg] = (s[0] + s[1]) + (s[2] + s[3]);
    s_q[rl * ne + cg] = (q[0] + q[1]) + (q[2] + q[3]);
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int g = threadIdx.x;
    double s = 0.0, q = 0.0;
    for (int l = 0; l < RL; ++l)
      for (int j = 0; j < epg; ++j) { s += s_s[l * ne + g * epg + j]; q += s_q[l * ne + g * epg + j]; }
    const double cnt = static_cast<double>(HW) * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    out[n * 32 + g] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + 1e-5)));
  }
}

template <bool kHalfIn>
__device__ __forceinline__ void gn_load8(const GnApplyArgs& a, int n0, int n1, int pix, int c, float (&v)[8]) {
  const long HW = static_cast<long>(a.H) * a.W;
  if (c < a.C0) {
    const long off = (static_cast<long>(n0) * HW + pix) * a.C0 + c;
    if (kHalfIn) {
      const uint4 r = __ldg(reinterpret_cast<const uint4*>(static_cast<const op_t*>(a.src0) + off));
      const float2 p0 = unpack_op2(r.x), p1 = unpack_op2(r.y), p2 = unpack_op2(r.z), p3 = unpack_op2(r.w);
      v[0] = p0.x; v[1] = p0.y; v[2] = p1.x; v[3] = p1.y; v[4] = p2.x; v[5] = p2.y; v[6] = p3.x; v[7] = p3.y;
      return;
    }
    const float* p = static_cast<const float*>(a.src0) + off;
    const float4 lo = __ldg(reinterpret_cast<const float4*>(p)), hi = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
  } else {
    const long off = (static_cast<long>(n1) * HW + pix) * a.C1 + (c - a.C0);
    if (kHalfIn) {  // both concat sources are 16-bit
      const uint4 r = __ldg(reinterpret_cast<const uint4*>(static_cast<const op_t*>(a.src1) + off));
      const float2 p0 = unpack_op2(r.x), p1 = unpack_op2(r.y), p2 = unpack_op2(r.z), p3 = unpack_op2(r.w);
      v[0] = p0.x; v[1] = p0.y; v[2] = p1.x; v[3] = p1.y; v[4] = p2.x; v[5] = p2.y; v[6] = p3.x; v[7] = p3.y;
      return;
    }
    const float* p = static_cast<const float*>(a.src1) + off;
    const float4 lo = __ldg(reinterpret_cast<const float4*>(p)), hi = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
  }
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 lo = __ldg(reinterpret_cast<const float4*>(p)), hi = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}
__device__ __forceinline__ void store8_op(op_t* dst, const float (&y)[8]) {
  uint4 o = make_uint4(pack_op2(y[0], y[1]), pack_op2(y[2], y[3]), pack_op2(y[4], y[5]), pack_op2(y[6], y[7]));
  *reinterpret_cast<uint4*>(dst) = o;
}
// Split-precision operand layout (engine precision 1, "fp16 x3"): a C-channel activation row becomes 3C channels
// [hi | hi | lo] with hi = op(v), lo = op(v - hi); the matching weight rows are packed [w_hi | w_lo | w_hi]
// (pack_conv_weight_kernel), so the unchanged GEMM accumulates a_hi w_hi + a_hi w_lo + a_lo w_hi in fp32:
// the product of two ~22-bit operands.  `row` points at channel c of the first part, C = channels per part.
__device__ __forceinline__ void store8_split3(op_t* row, int C, const float (&y)[8]) {
  float hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hi[j] = from_op(to_op(y[j]));
    lo[j] = y[j] - hi[j];
  }
  store8_op(row, hi);
  store8_op(row + C, hi);
  store8_op(row + 2 * C, lo);
}
// out[(row) * (kSplit ? 3C : C) + c ...]
template <bool kSplit>
__device__ __forceinline__ void gn_store8(op_t* base, long row, int C, int c, const float (&y)[8]) {
  if (kSplit) store8_split3(base + row * 3 * C + c, C, y);
  else store8_op(base + row * C + c, y);
}

// Pass 2: normalise + affine (+FiLM) (+SiLU), write op_t NHWC, optionally pooled / upsampled.
// Thread = (8-channel slice cg, pixel lane): the folded per-channel scale/offset live in
// registers for the whole block; the block walks `ppb` pixels.
template <bool kHalfIn, int kResample, int kMinBlocks = 3, bool kSplit = false>
__global__ void __launch_bounds__(256, kMinBlocks) gn_apply_kernel(const GnApplyArgs a) {
  __shared__ float s_mean[32], s_rstd[32];
  pdl_launch_dependents();
  pdl_wait();
  const int C = a.C0 + a.C1, C8 = C >> 3, cpg = C / 32;
  const int n = blockIdx.y;
  const int n0 = a.mod0 ? n % a.mod0 : n, n1 = a.mod1 ? n % a.mod1 : n;  // source samples (shared CFG prefix)
  const int HW = a.H * a.W;
  const int cg = threadIdx.x % C8, lane = threadIdx.x / C8;
  const int c = cg * 8;
  // The per-channel parameters do not depend on the statistics: their loads are issued before the statistics are
  // fetched and published, so the two L2 round trips of the block prologue overlap.
  // eight consecutive channels: 16-byte loads (c % 8 == 0; every base is 16-byte aligned, checked at launch)
  const float* f = a.film ? a.film + static_cast<long>(n) * a.film_stride : nullptr;
  float gm[8], bt[8], fs[8], fb[8];
  ld8(a.gamma + c, gm);
  ld8(a.beta + c, bt);
  if (f) { ld8(f + c, fs); ld8(f + C + c, fb); }
  if (a.final) {
    if (threadIdx.x < 32) {
      const float2 mr = a.final[n * 32 + threadIdx.x];
      s_mean[threadIdx.x] = mr.x;
      s_rstd[threadIdx.x] = mr.y;
    }
  } else if (threadIdx.x < 32) {
    double s = 0.0, q = 0.0;
    for (int k = 0; k < a.chunks; ++k) {
      const double* pp = a.partial + ((static_cast<long>(n) * a.chunks + k) * 32 + threadIdx.x) * 2;
      s += pp[0];
      q += pp[1];
    }
    const double cnt = static_cast<double>(HW) * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
  // y = ((v - mean) rstd gamma + beta) (1 + scale) + shift  ==  v * ka + kb
  float ka[8], kb[8];
  {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c + j) / cpg;
      const float ga = gm[j] * s_rstd[g];
      const float be = bt[j] - s_mean[g] * ga;
      const float s1 = f ? 1.0f + fs[j] : 1.0f;
      ka[j] = ga * s1;
      kb[j] = f ? be * s1 + fb[j] : be;
    }
  }
  auto xform = [&](const float (&v)[8], float (&y)[8]) {
    if (a.silu) {  // SiLU with one reciprocal per four elements (common.cuh: silu4)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float t[4], s[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = v[4 * q + j] * ka[4 * q + j] + kb[4 * q + j];
        silu4(t, s);
#pragma unroll
        for (int j = 0; j < 4; ++j) y[4 * q + j] = s[j];
      }
      return;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = v[j] * ka[j] + kb[j];
  };
  const int Ho = kResample == 1 ? a.H >> 1 : a.H, Wo = kResample == 1 ? a.W >> 1 : a.W;
  const int n_iter = Ho * Wo;  // pooled: output pixels; otherwise input pixels
  const int p_begin = blockIdx.x * a.ppb;
  const int p_end = min(p_begin + a.ppb, n_iter);
  if (kResample == 0) {
    int pix = p_begin + lane;
    if (kHalfIn && !kSplit) {
      // 16-bit source(s): eight pixels = 8 x 16-byte loads in flight per thread, kept packed.  A thread's
      // 8-channel slice lies in exactly one of the two concat sources.
      const bool first = c < a.C0;
      const long cs = first ? a.C0 : a.C1;
      const op_t* src = (first ? static_cast<const op_t*>(a.src0) + c : static_cast<const op_t*>(a.src1) + (c - a.C0)) +
                        static_cast<long>(first ? n0 : n1) * HW * cs;
      for (; pix + 7 * a.PLa < p_end; pix += 8 * a.PLa) {
        uint4 r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) r[u] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<long>(pix + u * a.PLa) * cs));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float2 p0 = unpack_op2(r[u].x), p1 = unpack_op2(r[u].y), p2 = unpack_op2(r[u].z), p3 = unpack_op2(r[u].w);
          const float v[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
          float y[8];
          xform(v, y);
          const long o = (static_cast<long>(n) * HW + pix + u * a.PLa) * C + c;
          store8_op(a.out + o, y);
          if (a.raw_out) *reinterpret_cast<uint4*>(a.raw_out + o) = r[u];
        }
      }
    }
    for (; pix + 3 * a.PLa < p_end; pix += 4 * a.PLa) {  // four pixels (8 x 16-byte loads) in flight
      float v[4][8], y[8];
#pragma unroll
      for (int u = 0; u < 4; ++u) gn_load8<kHalfIn>(a, n0, n1, pix + u * a.PLa, c, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xform(v[u], y);
        gn_store8<kSplit>(a.out, static_cast<long>(n) * HW + pix + u * a.PLa, C, c, y);
        if (a.raw_out) gn_store8<kSplit>(a.raw_out, static_cast<long>(n) * HW + pix + u * a.PLa, C, c, v[u]);
      }
    }
    for (; pix < p_end; pix += a.PLa) {
      float v[8], y[8];
      gn_load8<kHalfIn>(a, n0, n1, pix, c, v);
      xform(v, y);
      gn_store8<kSplit>(a.out, static_cast<long>(n) * HW + pix, C, c, y);
      if (a.raw_out) gn_store8<kSplit>(a.raw_out, static_cast<long>(n) * HW + pix, C, c, v);
    }
  } else if (kResample == 1) {
    for (int pix = p_begin + lane; pix < p_end; pix += a.PLa) {
      const int yo = pix / Wo, xo = pix - yo * Wo;
      float v[4][8], y[8], acc[8], racc[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) gn_load8<kHalfIn>(a, n0, n1, (2 * yo + (k >> 1)) * a.W + 2 * xo + (k & 1), c, v[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j] = 0.f; racc[j] = 0.f; }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        xform(v[k], y);
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[j] += y[j]; racc[j] += v[k][j]; }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j] *= 0.25f; racc[j] *= 0.25f; }
      const long o = (static_cast<long>(n) * n_iter + pix) * C + c;
      gn_store8<kSplit>(a.out, static_cast<long>(n) * n_iter + pix, C, c, acc);
      if (a.pool_out) {
        *reinterpret_cast<float4*>(a.pool_out + o) = make_float4(racc[0], racc[1], racc[2], racc[3]);
        *reinterpret_cast<float4*>(a.pool_out + o + 4) = make_float4(racc[4], racc[5], racc[6], racc[7]);
      }
    }
  } else {
    const int W2 = a.W * 2;
    for (int pix = p_begin + lane; pix < p_end; pix += a.PLa) {
      float v[8], y[8];
      gn_load8<kHalfIn>(a, n0, n1, pix, c, v);
      xform(v, y);
      const int yi = pix / a.W, xi = pix - yi * a.W;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        gn_store8<kSplit>(a.out, (static_cast<long>(n) * a.H * 2 + 2 * yi + (k >> 1)) * W2 + 2 * xi + (k & 1), C, c, y);
    }
  }
}

int gn_chunks_for(int B, int HW, int C) {
  // Depends on the per-sample shape ONLY (never on the batch): the fp32 summation order, and with
  // it every output bit, is then independent of how samples are batched together.
  (void)B;
  const int C4 = C / 4;
  const int PL = 256 / C4 > 0 ? 256 / C4 : 1;
  int chunks = 1;
  while (chunks < 16 && HW / (chunks * 2) >= 4 * PL && (HW % (chunks * 2)) == 0) chunks *= 2;
  return chunks;
}

static int gn_check(const GnDesc& d) {
  const int C = d.C0 + d.C1, HW = d.H * d.W;
  if (C % 32 || C > 2048 || d.C0 % 8 || d.C1 % 8 || HW % d.chunks) return 1;  // 2048: the mc = 256 skip concats
  return 0;
}

static bool gn_fused(const GnDesc& d) { return d.stats0 != nullptr && (d.C1 == 0 || d.stats1 != nullptr); }

int gn_finalize_launch(const GnDesc& d, cudaStream_t s) {
  const int C = d.C0 + d.C1, HW = d.H * d.W;
  if (gn_check(d) || !gn_fused(d) || d.final == nullptr || (HW % 32) || (d.stat_gran != 2 && d.stat_gran != 4) ||
      ((C / 32) % d.stat_gran) || (d.C0 % d.stat_gran) || (d.C1 % d.stat_gran))
    return 1;
  // RL row-block lanes per granule (a power of two, fixed by the sample's shape -> batch-invariant summation order)
  const int ne = C / d.stat_gran, rbs = HW / 32;
  int RL = 1;
  while (RL * 2 * ne <= 512 && RL * 2 <= rbs && RL < 16) RL *= 2;
  return launch_pdl(gn_finalize_kernel, dim3(d.B), dim3(ne * RL), 0, s, 1, d.stats0, d.C0, d.stats1, d.C1, HW, d.stat_gran, RL,
                    d.final, d.src_mod0, d.src_mod1) == cudaSuccess ? 0 : 1;
}

int gn_stats_launch(const GnDesc& d, cudaStream_t s) {
  if (gn_check(d) || (d.src0_is_op && d.C1)) return 1;  // a 16-bit concat only comes with producer statistics
  if (!d.src0_is_op && d.C0 + d.C1 > 1024) return 1;    // one thread per 4 channels, 256 threads
  const int C = d.C0 + d.C1, HW = d.H * d.W;
  if (d.src0_is_op) {
    const int C8 = C / 8;
    const int PL = 256 / C8 > 0 ? 256 / C8 : 1;
    gn_stats_op_kernel<<<dim3(d.chunks, d.B), C8 * PL, 0, s>>>(static_cast<const op_t*>(d.src0), HW, C, d.chunks, PL,
                                                             d.partial);
  } else {
    const int C4 = C / 4;
    const int PL = 256 / C4 > 0 ? 256 / C4 : 1;
    gn_stats_kernel<false><<<dim3(d.chunks, d.B), C4 * PL, 0, s>>>(d.src0, static_cast<const float*>(d.src1), HW, d.C0, d.C1,
                                                                   d.chunks, PL, d.partial);
  }
  return SGDM_LAUNCH_OK();
}

int gn_apply_launch(const GnDesc& d, cudaStream_t s) {
  if (gn_check(d)) return 1;
  const int C = d.C0 + d.C1, HW = d.H * d.W;
  const int C8 = C / 8;
  const int PLa = 256 / C8 > 0 ? 256 / C8 : 1;
  const int n_iter = d.resample == 1 ? HW / 4 : HW;
  // pixels per block: >= 8 per thread when the grid stays large enough to fill the chip
  // the per-block prologue (group statistics, folded scale / offset per channel) is amortised over ppb pixels
  int ppb = PLa * (d.src0_is_op ? 64 : 32);
  while (ppb > PLa && static_cast<long>(d.B) * ((n_iter + ppb - 1) / ppb) < 8 * kNumSMs) ppb >>= 1;
  if (ppb > n_iter) ppb = n_iter;
  if ((reinterpret_cast<uintptr_t>(d.gamma) | reinterpret_cast<uintptr_t>(d.beta) | reinterpret_cast<uintptr_t>(d.film)) & 15 ||
      (d.film && (d.film_stride & 3)))
    return 1;
  GnApplyArgs a{d.src0, d.src1, d.H, d.W, d.C0, d.C1, d.gamma, d.beta, d.film, d.film_stride,
                d.silu, d.resample, d.chunks, ppb, PLa, d.partial, gn_fused(d) ? d.final : nullptr, d.out, d.raw_out,
                d.pool_out, d.src_mod0, d.src_mod1};
  if ((d.src_mod0 || d.src_mod1) && !gn_fused(d)) return 1;  // the shared-prefix sources always carry producer statistics
  const dim3 grid((n_iter + ppb - 1) / ppb, d.B);
  const int threads = C8 * PLa;
  if (d.split3) {  // split-precision operands (engine precision 1): fp32 sources only (h1 stays fp32 in that mode)
    if (d.src0_is_op) return 1;
    if (d.resample == 0) return launch_pdl(gn_apply_kernel<false, 0, 3, true>, grid, dim3(threads), 0, s, 1, a) == cudaSuccess ? 0 : 1;
    else if (d.resample == 1) return launch_pdl(gn_apply_kernel<false, 1, 3, true>, grid, dim3(threads), 0, s, 1, a) == cudaSuccess ? 0 : 1;
    else return launch_pdl(gn_apply_kernel<false, 2, 3, true>, grid, dim3(threads), 0, s, 1, a) == cudaSuccess ? 0 : 1;
  }
  // Three resident blocks of 256 threads per SM, spill-free (four at 64 registers spill ~200 bytes per thread:
  // measured on B200, config 2 at batch 256, same box: gn_apply 10.4 -> 9.5 ms per step with three).
  if (d.src0_is_op) {
    if (d.resample == 0) return launch_pdl(gn_apply_kernel<true, 0, 3>, grid, dim3(threads), 0, s, 1, a) == cudaSuccess ? 0 : 1;
    else if (d.resample == 1) return launch_pdl(gn_apply_kernel<true, 1, 3>, grid, dim3(threads), 0, s, 1, a) == cudaSuccess ? 0 : 1;
    else return launch_pdl(gn_apply_kernel<true, 2, 3>, grid, dim3(threads), 0, s, 1, a) == cudaSuccess ? 0 : 1;
  }
  if (d.resample == 0) return launch_pdl(gn_apply_kernel<false, 0, 3>, grid, dim3(threads), 0, s, 1, a) == cudaSuccess ? 0 : 1;
  else if (d.resample == 1) return launch_pdl(gn_apply_kernel<false, 1, 3>, grid, dim3(threads), 0, s, 1, a) == cudaSuccess ? 0 : 1;
  return launch_pdl(gn_apply_kernel<false, 2, 3>, grid, dim3(threads), 0, s, 1, a) == cudaSuccess ? 0 : 1;
}

int gn_launch(const GnDesc& d, cudaStream_t s) {
  return (gn_fused(d) ? gn_finalize_launch(d, s) : gn_stats_launch(d, s)) || gn_apply_launch(d, s);
}

// =========================================================================== LayerNorm
// One warp per row of C channels (C % 128 == 0, C <= 1024); two-pass in registers.
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const float* __restrict__ res,
                                 op_t* __restrict__ out_op, float* __restrict__ out_f32, long rows, int C, int split3) {
  const long row = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nv = C >> 7;  // float4 per lane
  // (fully unrolled with a predicate per slot: a runtime trip count put v[] into local memory — 128 B of stack)
  float4 v[8];
  float s = 0.f;
  const float4* xr = reinterpret_cast<const float4*>(x + row * C);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) v[j] = __ldg(xr + j * 32 + lane);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) s += v[j].x + v[j].y + v[j].z + v[j].w;
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + 1e-5f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j >= nv) break;
    const int c = (j * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 y = make_float4((v[j].x - mean) * rstd * g.x + b.x, (v[j].y - mean) * rstd * g.y + b.y,
                           (v[j].z - mean) * rstd * g.z + b.z, (v[j].w - mean) * rstd * g.w + b.w);
    if (out_f32) {
      if (res) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(res + row * C + c));
        y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
      }
      *reinterpret_cast<float4*>(out_f32 + row * C + c) = y;
    } else if (split3) {  // [hi | hi | lo] parts of a 3C-channel row (see store8_split3)
      const float4 hi = make_float4(from_op(to_op(y.x)), from_op(to_op(y.y)), from_op(to_op(y.z)), from_op(to_op(y.w)));
      const uint2 h = make_uint2(pack_op2(hi.x, hi.y), pack_op2(hi.z, hi.w));
      op_t* o = out_op + row * 3 * C + c;
      *reinterpret_cast<uint2*>(o) = h;
      *reinterpret_cast<uint2*>(o + C) = h;
      *reinterpret_cast<uint2*>(o + 2 * C) = make_uint2(pack_op2(y.x - hi.x, y.y - hi.y), pack_op2(y.z - hi.z, y.w - hi.w));
    } else {
      *reinterpret_cast<uint2*>(out_op + row * C + c) = make_uint2(pack_op2(y.x, y.y), pack_op2(y.z, y.w));
    }
  }
}
// out = res + LN(x) gamma + beta (the tail of Attention_LR, crossattetion_lr.py:139-142) that ALSO emits the
// GroupNorm partial statistics of its output in the conv epilogue's format (ConvDesc::stats: {sum, sum of squares}
// per 32-row block and `gran` adjacent channels), so the GroupNorm that follows needs no pass over the tensor.
// Block = 8 warps = 32 consecutive rows (4 per warp); a lane's float4 j covers channels (j * 32 + lane) * 4 .. + 3,
// i.e. one granule of 4 or two of 2; fixed summation order (rows of a warp, then the 8 warps in order).
template <int kGran>
__global__ void __launch_bounds__(256, 2) layernorm_res_stats_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, const float* __restrict__ res,
                                                                  float* __restrict__ out, float2* __restrict__ stats, long rows,
                                                                  int C) {
  constexpr int G = 4 / kGran;  // granules per float4
  __shared__ float2 part[8][256 * G];  // [warp][granule] (C <= 1024: C / kGran <= 256 * G)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = C >> 7;
  const long row0 = blockIdx.x * 32L;
  float sa[8][G], qa[8][G];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int g = 0; g < G; ++g) { sa[j][g] = 0.f; qa[j][g] = 0.f; }
  // (the loads of the warp's NEXT row are issued before the current row is reduced: two rows in flight)
  float4 vn[8];
  {
    const float4* xr = reinterpret_cast<const float4*>(x + (row0 + warp * 4) * C);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nv) vn[j] = __ldg(xr + j * 32 + lane);
  }
  for (int rr = 0; rr < 4; ++rr) {
    const long row = row0 + warp * 4 + rr;
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nv) v[j] = vn[j];
    // the residual row does not depend on the statistics: its loads go out before the two warp reductions
    float4 rv[8];
    {
      const float4* rr4 = reinterpret_cast<const float4*>(res + row * C);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nv) rv[j] = __ldg(rr4 + j * 32 + lane);
    }
    if (rr + 1 < 4) {
      const float4* xr = reinterpret_cast<const float4*>(x + (row + 1) * C);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nv) vn[j] = __ldg(xr + j * 32 + lane);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nv) s += v[j].x + v[j].y + v[j].z + v[j].w;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nv) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        q += a * a + b * b + c * c + d * d;
      }
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + 1e-5f);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nv) {
        const int c = (j * 32 + lane) * 4;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
        const float4 r = rv[j];
        const float4 y = make_float4((v[j].x - mean) * rstd * g.x + b.x + r.x, (v[j].y - mean) * rstd * g.y + b.y + r.y,
                                     (v[j].z - mean) * rstd * g.z + b.z + r.z, (v[j].w - mean) * rstd * g.w + b.w + r.w);
        *reinterpret_cast<float4*>(out + row * C + c) = y;
        if (G == 1) {
          sa[j][0] += (y.x + y.y) + (y.z + y.w);
          qa[j][0] += (y.x * y.x + y.y * y.y) + (y.z * y.z + y.w * y.w);
        } else {
          sa[j][0] += y.x + y.y; qa[j][0] += y.x * y.x + y.y * y.y;
          sa[j][G - 1] += y.z + y.w; qa[j][G - 1] += y.z * y.z + y.w * y.w;
        }
      }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv)
#pragma unroll
      for (int g = 0; g < G; ++g) part[warp][(j * 32 + lane) * G + g] = make_float2(sa[j][g], qa[j][g]);
  __syncthreads();
  const int ng = C / kGran;
  for (int e = threadIdx.x; e < ng; e += 256) {
    float2 t = part[0][e];
#pragma unroll
    for (int w = 1; w < 8; ++w) { t.x += part[w][e].x; t.y += part[w][e].y; }
    stats[blockIdx.x * static_cast<long>(ng) + e] = t;
  }
}
int layernorm_res_stats_launch(const float* x, const float* gamma, const float* beta, const float* res, float* out_f32,
                               float2* stats, int stat_gran, long rows, int C, cudaStream_t s) {
  if (C % 128 || C > 1024 || (rows % 32) || !res || !stats || (stat_gran != 2 && stat_gran != 4)) return 1;
  const unsigned grid = static_cast<unsigned>(rows / 32);
  if (stat_gran == 4) layernorm_res_stats_kernel<4><<<grid, 256, 0, s>>>(x, gamma, beta, res, out_f32, stats, rows, C);
  else layernorm_res_stats_kernel<2><<<grid, 256, 0, s>>>(x, gamma, beta, res, out_f32, stats, rows, C);
  return SGDM_LAUNCH_OK();
}
int layernorm_launch(const float* x, const float* gamma, const float* beta, const float* res, op_t* out_op,
                     float* out_f32, long rows, int C, cudaStream_t s, int split3) {
  if (C % 128 || C > 1024) return 1;
  const long threads = rows * 32;
  layernorm_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(x, gamma, beta, res, out_op, out_f32,
                                                                              rows, C, split3);
  return SGDM_LAUNCH_OK();
}

// =========================================================================== casts
__global__ void cast_kernel(const float* __restrict__ src, op_t* __restrict__ dst, int H, int W, int C, int up2,
                            long items, int split3) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= items) return;
  const int C8 = C >> 3;
  const int cg = idx % C8;
  const long pixg = idx / C8;  // global input pixel index (n*H*W + y*W + x)
  const float* p = src + pixg * C + cg * 8;
  const float4 lo = __ldg(reinterpret_cast<const float4*>(p));
  const float4 hi = __ldg(reinterpret_cast<const float4*>(p + 4));
  const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  auto put = [&](long row) {
    if (split3) store8_split3(dst + row * 3 * C + cg * 8, C, v);
    else store8_op(dst + row * C + cg * 8, v);
  };
  if (!up2) {
    put(pixg);
  } else {
    const long n = pixg / (static_cast<long>(H) * W);
    const int pix = pixg - n * H * W;
    const int y = pix / W, x = pix - y * W;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) put((n * 2 * H + 2 * y + dy) * (2 * W) + 2 * x + dx);
  }
}
int cast_launch(const float* src, op_t* dst, int B, int H, int W, int C, int up2, cudaStream_t s, int split3) {
  if (C % 8) return 1;
  const long items = static_cast<long>(B) * H * W * (C / 8);
  cast_kernel<<<static_cast<unsigned>((items + 255) / 256), 256, 0, s>>>(src, dst, H, W, C, up2, items, split3);
  return SGDM_LAUNCH_OK();
}

// dst[r, c] = op(silu(src[r, c])); split3: rows of 3C channels [hi | hi | lo]
__global__ void silu_cast_kernel(const float* __restrict__ src, op_t* __restrict__ dst, long n, int C, int split3) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float v = silu(src[i]);
  if (!split3) { dst[i] = to_op(v); return; }
  const long r = i / C;
  const int c = static_cast<int>(i - r * C);
  const op_t hi = to_op(v);
  op_t* o = dst + r * 3 * C + c;
  o[0] = hi;
  o[C] = hi;
  o[2 * C] = to_op(v - from_op(hi));
}
int silu_cast_launch(const float* src, op_t* dst, long n, cudaStream_t s, int C, int split3) {
  silu_cast_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(src, dst, n, C > 0 ? C : 1, split3);
  return SGDM_LAUNCH_OK();
}

// 16-bit [rows, C] -> [rows, 3C] = [v | v | 0]: a tensor that only exists in the operand type (the attention output)
// as the activation side of a split-precision GEMM (v w_hi + v w_lo).
__global__ void expand3_kernel(const op_t* __restrict__ src, op_t* __restrict__ dst, long items, int C8) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= items) return;
  const long row = idx / C8;
  const int cg = static_cast<int>(idx - row * C8);
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (row * C8 + cg) * 8));
  uint4* o = reinterpret_cast<uint4*>(dst + (row * 3 * C8 + cg) * 8);
  o[0] = v;
  o[C8] = v;
  o[2 * C8] = make_uint4(0u, 0u, 0u, 0u);
}
int expand3_launch(const op_t* src, op_t* dst, long rows, int C, cudaStream_t s) {
  if (C % 8) return 1;
  const long items = rows * (C / 8);
  expand3_kernel<<<static_cast<unsigned>((items + 255) / 256), 256, 0, s>>>(src, dst, items, C / 8);
  return SGDM_LAUNCH_OK();
}

// =========================================================================== fp32 linear
// 64x64 output tile per 256-thread CTA, 4x4 per thread, K staged 32 at a time through smem; the next K tile is prefetched
// into registers while the current one is multiplied (the K = 5000 cluster-condition MLP of config 3 was bound by the
// exposed load latency of 312 unpipelined 16-wide tiles: 0.8 ms).  Per output the products are added in ascending k.
__global__ void __launch_bounds__(256) linear_f32_kernel(const float* __restrict__ in, long in_stride, const float* __restrict__ W,
                                                         const float* __restrict__ bias, float* __restrict__ out, long out_stride, int M,
                                                         int N, int K, int silu_out, int accumulate, float* __restrict__ partial,
                                                         int k_chunk) {
  // split-K (gridDim.z > 1): this CTA multiplies k in [z * k_chunk, (z + 1) * k_chunk) and writes the raw partial sums
  // to partial[z][m][n]; linear_reduce_kernel adds them in ascending z (deterministic), then bias / SiLU / accumulate
  constexpr int KT = 32;
  const bool split = gridDim.z > 1;
  const int k_lo = split ? static_cast<int>(blockIdx.z) * k_chunk : 0;
  const int k_hi = split ? min(K, k_lo + k_chunk) : K;
  __shared__ float As[KT][65];
  __shared__ float Ws[KT][65];
  const int tm = blockIdx.y * 64, tn = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lk = threadIdx.x & 31, lr = threadIdx.x >> 5;  // loader: k within the tile, row (+ 8 j)
  float acc[4][4] = {};
  float pa[8], pw[8];
  auto fetch = [&](int k0) {
    const int k = k0 + lk;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int m = tm + lr + 8 * j, n = tn + lr + 8 * j;
      pa[j] = (m < M && k < k_hi) ? in[m * in_stride + k] : 0.f;
      pw[j] = (n < N && k < k_hi) ? W[static_cast<long>(n) * K + k] : 0.f;
    }
  };
  if (k_lo < k_hi) fetch(k_lo);
  for (int k0 = k_lo; k0 < k_hi; k0 += KT) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { As[lk][lr + 8 * j] = pa[j]; Ws[lk][lr + 8 * j] = pw[j]; }
    __syncthreads();
    if (k0 + KT < k_hi) fetch(k0 + KT);
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
  float* pz = split ? partial + static_cast<long>(blockIdx.z) * M * N : nullptr;
  for (int i = 0; i < 4; ++i) {
    const int m = tm + ty * 4 + i;
    if (m >= M) continue;
    for (int j = 0; j < 4; ++j) {
      const int n = tn + tx * 4 + j;
      if (n >= N) continue;
      if (split) { pz[static_cast<long>(m) * N + n] = acc[i][j]; continue; }
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (silu_out) v = silu(v);
      float* o = out + m * out_stride + n;
      *o = accumulate ? *o + v : v;
    }
  }
}
__global__ void linear_reduce_kernel(const float* __restrict__ partial, int splits, const float* __restrict__ bias,
                                     float* __restrict__ out, long out_stride, int M, int N, int silu_out, int accumulate) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long>(M) * N) return;
  const int m = static_cast<int>(i / N), n = static_cast<int>(i - static_cast<long>(m) * N);
  float acc = 0.f;
  for (int z = 0; z < splits; ++z) acc += partial[static_cast<long>(z) * M * N + i];
  float v = acc + (bias ? bias[n] : 0.f);
  if (silu_out) v = silu(v);
  float* o = out + m * out_stride + n;
  *o = accumulate ? *o + v : v;
}
// how many K slices a skinny, deep Linear is cut into (1 = no split): enough CTAs to cover the SMs, slices of >= 256
int linear_f32_splits(int M, int N, int K) {
  const int ctas = ((N + 63) / 64) * ((M + 63) / 64);
  if (K < 1024 || ctas >= 64) return 1;
  int s = (128 + ctas - 1) / ctas;
  if (s > K / 256) s = K / 256;
  if (s > 16) s = 16;
  return s < 2 ? 1 : s;
}
int linear_f32_launch(const float* in, long in_stride, const float* W, const float* bias, float* out,
                      long out_stride, int M, int N, int K, int silu_out, int accumulate, cudaStream_t s, float* partial,
                      int splits) {
  if (splits > 1 && partial != nullptr) {
    const int k_chunk = ((K + splits - 1) / splits + 31) / 32 * 32;
    linear_f32_kernel<<<dim3((N + 63) / 64, (M + 63) / 64, splits), 256, 0, s>>>(in, in_stride, W, bias, out, out_stride, M,
                                                                                N, K, silu_out, accumulate, partial, k_chunk);
    const long total = static_cast<long>(M) * N;
    linear_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(partial, splits, bias, out, out_stride, M,
                                                                                    N, silu_out, accumulate);
    return SGDM_LAUNCH_OK();
  }
  linear_f32_kernel<<<dim3((N + 63) / 64, (M + 63) / 64), 256, 0, s>>>(in, in_stride, W, bias, out, out_stride, M, N,
                                                                      K, silu_out, accumulate, nullptr, 0);
  return SGDM_LAUNCH_OK();
}

// =========================================================================== prologue
// x (NCHW fp32) -> NHWC op_t with 64 channels: [x_hi | x_lo | layout | 0].  x_lo = x - x_hi
// restores the precision lost by the 16-bit rounding of the image (its weights duplicate
// the image-channel weights), so the first conv sees x to ~2^-21.
// Eight threads per pixel, each builds one 16-byte slice (8 of the 64 slots) in registers: a warp writes 512 contiguous
// bytes per store.  grid.y = row (sample of the doubled batch): no 64-bit division per thread; the (tap, entry) decode of
// the 64 slots is a shared-memory table built once per block.  (The first version — one thread per pixel, the 64 slots
// in a local-memory array, 128-byte-strided stores, a runtime division per slot — took 0.33 ms per step at batch 256.)
// Only the rows the first conv consumes are produced (Bx: the shared rows of a guided plan).
__global__ void __launch_bounds__(256) prep_x_kernel(const PrepDesc d) {
  __shared__ int s_tab[64];  // per slot: (dy + 1) | (dx + 1) << 2 | entry << 4, or -1 for a slot that stays zero
  const int ce = 2 * d.Cimg + d.L;
  if (threadIdx.x < 64) {
    const int k = threadIdx.x;
    int code = -1;
    if (d.im2col) {
      const int tap = k / ce, j = k - tap * ce;
      if (tap < 9) code = (tap / 3) | ((tap % 3) << 2) | (j << 4);
    } else if (k < ce) {
      code = 1 | (1 << 2) | (k << 4);  // the pixel itself
    }
    s_tab[k] = code;
  }
  __syncthreads();
  const int HW = d.H * d.W;
  const int r = blockIdx.y;
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int pix = static_cast<int>(gid >> 3), t8 = static_cast<int>(gid & 7);
  if (pix >= HW) return;
  const int b = r % d.B;
  const bool drop = d.drop && d.drop[r];
  const int y = pix / d.W, x = pix - y * d.W;
  const float* xb = d.x + static_cast<long>(b) * d.Cimg * HW;
  float val[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int code = s_tab[t8 * 8 + i];
    float v = 0.f;
    if (code >= 0) {
      const int yy = y + (code & 3) - 1, xx = x + ((code >> 2) & 3) - 1, j = code >> 4;
      if (yy >= 0 && yy < d.H && xx >= 0 && xx < d.W) {
        const int pp = yy * d.W + xx;
        if (j < 2 * d.Cimg) {
          const float xv = xb[(j < d.Cimg ? j : j - d.Cimg) * HW + pp];
          v = j < d.Cimg ? xv : xv - from_op(to_op(xv));  // x_hi (rounded below) | x_lo = x - x_hi
        } else {
          v = drop ? d.null_layout[pp] : d.layout[(static_cast<long>(b) * d.L + (j - 2 * d.Cimg)) * HW + pp];
        }
      }
    }
    val[i] = v;
  }
  *reinterpret_cast<uint4*>(d.x_in + (static_cast<long>(r) * HW + pix) * 64 + t8 * 8) =
      make_uint4(pack_op2(val[0], val[1]), pack_op2(val[2], val[3]), pack_op2(val[4], val[5]), pack_op2(val[6], val[7]));
}
// Split-precision first-conv input (engine precision 1): xc channels per pixel,
// [hi(x, layout) | hi(x, layout) | lo(x, layout) | 0] with parts of Cimg + L channels (see store8_split3).
__global__ void prep_x_split_kernel(const PrepDesc d) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const int HW = d.H * d.W;
  if (idx >= static_cast<long>(d.Bp) * HW) return;
  const int r = idx / HW, pix = idx - static_cast<long>(r) * HW;
  const int b = r % d.B;
  const bool drop = d.drop && d.drop[r];
  const int cp = d.Cimg + d.L;
  op_t* o = d.x_in + idx * d.xc;
  for (int c = 0; c < cp; ++c) {
    const float v = c < d.Cimg ? d.x[(static_cast<long>(b) * d.Cimg + c) * HW + pix]
                               : (drop ? d.null_layout[pix] : d.layout[(static_cast<long>(b) * d.L + (c - d.Cimg)) * HW + pix]);
    const op_t hi = to_op(v);
    o[c] = hi;
    o[cp + c] = hi;
    o[2 * cp + c] = to_op(v - from_op(hi));
  }
  for (int c = 3 * cp; c < d.xc; ++c) o[c] = to_op(0.f);
}
__global__ void prep_emb_kernel(const PrepDesc d) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const int half = d.mc / 2;
  const long n_t = static_cast<long>(d.Bp) * half;
  const long n_c = static_cast<long>(d.Bp) * d.cond_dim;
  if (idx < n_t) {
    const int r = idx / half, i = idx - static_cast<long>(r) * half;
    const float arg = __fmul_rn(static_cast<float>(d.t[r % d.B]), d.freqs[i]);
    d.t_emb[static_cast<long>(r) * d.mc + i] = cosf(arg);
    d.t_emb[static_cast<long>(r) * d.mc + half + i] = sinf(arg);
    if ((d.mc & 1) && i == 0) d.t_emb[static_cast<long>(r) * d.mc + d.mc - 1] = 0.f;
  } else if (idx < n_t + n_c) {
    const long k = idx - n_t;
    const int r = k / d.cond_dim, j = k - static_cast<long>(r) * d.cond_dim;
    const bool drop = d.drop && d.drop[r];
    d.cond_masked[k] = drop ? d.null_cond[j] : d.cond[static_cast<long>(r % d.B) * d.cond_dim + j];
  }
}
int prep_launch(const PrepDesc& d, cudaStream_t s) {
  const long npx = static_cast<long>(d.Bp) * d.H * d.W;
  if (d.split3) {
    if (3 * (d.Cimg + d.L) > d.xc || d.im2col) return 1;
    prep_x_split_kernel<<<static_cast<unsigned>((npx + 127) / 128), 128, 0, s>>>(d);
  } else {
    if (2 * d.Cimg + d.L > 64 || (d.im2col && 9 * (2 * d.Cimg + d.L) > 64)) return 1;
    const int rows = d.Bx > 0 ? d.Bx : d.Bp;
    if (rows > 65535) return 1;  // grid.y
    prep_x_kernel<<<dim3(static_cast<unsigned>((d.H * d.W * 8 + 255) / 256), rows), 256, 0, s>>>(d);
  }
  const long ne = static_cast<long>(d.Bp) * (d.mc / 2) + static_cast<long>(d.Bp) * d.cond_dim;
  prep_emb_kernel<<<static_cast<unsigned>((ne + 255) / 256), 256, 0, s>>>(d);
  return SGDM_LAUNCH_OK();
}

// =========================================================================== context K/V
// grid = (samples, attention sites): the context tokens depend on (t, cond) only, so the K/V rows of ALL Attention_LR
// sites are produced by one launch in the prologue instead of one launch per site inside the chain.
__global__ void context_kv_kernel(const CtxDesc d) {
  extern __shared__ float sm[];
  const int ctx = d.ctx, dh = d.dh;
  const CtxSite& w = d.site[blockIdx.y];
  float* tok = sm;            // [n_tok][ctx]
  const int n = blockIdx.x, nt = d.n_tok, rows = nt + 1;
  for (int i = threadIdx.x; i < nt * ctx; i += blockDim.x) {
    const int t = i / ctx, j = i - t * ctx;
    tok[i] = t < 8 ? d.time_tokens[(static_cast<long>(n) * 8 + t) * ctx + j]
                   : d.cond_tokens[(static_cast<long>(n) * (nt - 8) + (t - 8)) * ctx + j];
  }
  __syncthreads();
  // two LayerNorms back to back (norm_cond, then to_context.0), one thread per token
  if (threadIdx.x < nt) {
    float* row = tok + threadIdx.x * ctx;
    for (int pass = 0; pass < 2; ++pass) {
      const float* g = pass == 0 ? d.norm_w : w.ln_w;
      const float* b = pass == 0 ? d.norm_b : w.ln_b;
      float mean = 0.f;
      for (int j = 0; j < ctx; ++j) mean += row[j];
      mean /= ctx;
      float var = 0.f;
      for (int j = 0; j < ctx; ++j) { const float dlt = row[j] - mean; var += dlt * dlt; }
      const float rstd = rsqrtf(var / ctx + 1e-5f);
      for (int j = 0; j < ctx; ++j) row[j] = (row[j] - mean) * rstd * g[j] + b[j];
    }
  }
  __syncthreads();
  // thread = output feature o: its weight row is read once per 16 tokens (16-byte loads of whole lines; token values
  // are shared-memory broadcasts); j-ascending summation order per output
  for (int o = threadIdx.x; o < 2 * dh; o += blockDim.x) {
    const float* wr = w.lin_w + static_cast<long>(o) * ctx;
    const float bias = w.lin_b[o];
    for (int t0 = 0; t0 < nt; t0 += 16) {
      float acc[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) acc[t] = bias;
      for (int j = 0; j < ctx; j += 4) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + j));
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          if (t0 + t >= nt) break;
          const float4 tv = *reinterpret_cast<const float4*>(tok + (t0 + t) * ctx + j);
          acc[t] += tv.x * wv.x;
          acc[t] += tv.y * wv.y;
          acc[t] += tv.z * wv.z;
          acc[t] += tv.w * wv.w;
        }
      }
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        if (t0 + t >= nt) break;
        if (o < dh) w.k_out[(static_cast<long>(n) * rows + t0 + t) * dh + o] = to_op(acc[t]);
        else w.v_out[(static_cast<long>(n) * rows + t0 + t) * dh + (o - dh)] = to_op(acc[t]);
      }
    }
  }
  for (int o = threadIdx.x; o < dh; o += blockDim.x) {
    w.k_out[(static_cast<long>(n) * rows + nt) * dh + o] = to_op(w.null_kv[o]);
    w.v_out[(static_cast<long>(n) * rows + nt) * dh + o] = to_op(w.null_kv[dh + o]);
  }
}
int context_kv_launch(const CtxDesc& d, cudaStream_t s) {
  if (d.n_sites < 1 || d.n_sites > kMaxCtxSites || (d.ctx % 4) || d.n_tok < 8 || d.n_tok > kMaxCtxTok) return 1;
  context_kv_kernel<<<dim3(d.Bp, d.n_sites), 128, d.n_tok * d.ctx * sizeof(float), s>>>(d);
  return SGDM_LAUNCH_OK();
}

__global__ void token_mean_kernel(const float* __restrict__ in, float* __restrict__ out, int n_tok, int dim, long total) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const long b = i / dim;
  const int j = static_cast<int>(i - b * dim);
  float acc = 0.f;
  for (int t = 0; t < n_tok; ++t) acc += in[(b * n_tok + t) * dim + j];
  out[i] = acc / static_cast<float>(n_tok);
}
int token_mean_launch(const float* in, float* out, int B, int n_tok, int dim, cudaStream_t s) {
  const long total = static_cast<long>(B) * dim;
  token_mean_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, out, n_tok, dim, total);
  return SGDM_LAUNCH_OK();
}

// =========================================================================== guidance mix + sampler updates
// All arithmetic uses explicit round-to-nearest intrinsics in the reference's operation
// order (no FMA contraction), so that given identical eps the update is bit-identical to
// the unfused torch ops it replaces.
__device__ __forceinline__ float mix1(const MixDesc& m, float ec, float eu, float w, float ow) {
  if (m.eps_u == nullptr) return ec;
  if (m.scale_type == 0) return __fadd_rn(__fmul_rn(ow, eu), __fmul_rn(w, ec));  // (1-w) z + w zc
  return __fsub_rn(__fmul_rn(ow, ec), __fmul_rn(w, eu));                        // (1+w) zc - w z
}
__device__ __forceinline__ void mix_coeffs(const MixDesc& m, int b, float& w, float& ow) {
  if (m.w_per_sample) {  // fp32 tensor cond_scale: (1 -/+ w) is an fp32 tensor op in the reference
    w = m.w_per_sample[b];
    ow = m.scale_type == 0 ? __fsub_rn(1.0f, w) : __fadd_rn(1.0f, w);
  } else {               // Python-number cond_scale: (1 -/+ w) evaluated in double, then cast
    w = m.w;
    ow = m.ow;
  }
}

__global__ void mix_kernel(const MixDesc m, float* __restrict__ out, long per_sample, long total) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  float w, ow;
  mix_coeffs(m, static_cast<int>(i / per_sample), w, ow);
  out[i] = mix1(m, m.eps_c[i], m.eps_u ? m.eps_u[i] : 0.f, w, ow);
}
int mix_launch(const MixDesc& m, float* eps_out, int B, long per_sample, cudaStream_t s) {
  const long total = B * per_sample;
  if (total == 0) return 0;  // an empty batch is a no-op, like the reference's torch ops
  return launch_pdl(mix_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, 1, m, eps_out, per_sample, total) == cudaSuccess ? 0 : 1;
}

// Optional extras of the update kernels (sampling_kwargs dtp < 1 / noise_dropout > 0):
//   x0_raw   phase 1 of dynamic thresholding: only the UNCLIPPED pred_x0 is written (then quantile_abs_kernel)
//   dyn_s    [B] per-sample threshold s = max(quantile(|x0|, dtp), 1): x0 <- clamp(x0, -s, s) / s, replacing the
//            clamp to [-1, 1] (clip_x0_minus_one_to_one, diffusion_utils/util.py:70-82)
//   noise_mul per-element F.dropout factor {0, 1/(1-p)} applied to the scaled noise last, like F.dropout does
__device__ __forceinline__ float clip_x0_dev(float x0, int clip, const float* dyn_s, int b) {
  if (dyn_s) {
    const float s = dyn_s[b];
    return __fdiv_rn(fminf(fmaxf(x0, -s), s), s);
  }
  return clip ? fminf(fmaxf(x0, -1.0f), 1.0f) : x0;
}

__global__ void ddim_step_kernel(const MixDesc m, const DdimCoef c, const StepExtras ex, const float* __restrict__ x,
                                 const float* __restrict__ noise, float* __restrict__ x_out,
                                 float* __restrict__ x0_out, float* __restrict__ eps_out, long per_sample,
                                 long total) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int b = static_cast<int>(i / per_sample);
  float w, ow;
  mix_coeffs(m, b, w, ow);
  const float e = mix1(m, m.eps_c[i], m.eps_u ? m.eps_u[i] : 0.f, w, ow);
  // pred_x0 = (x - sqrt(1-a_t) e) / sqrt(a_t)
  float x0 = __fdiv_rn(__fsub_rn(x[i], __fmul_rn(c.sqrt_one_minus_at, e)), c.sqrt_at);
  if (ex.x0_raw) { ex.x0_raw[i] = x0; return; }
  x0 = clip_x0_dev(x0, c.clip, ex.dyn_s, b);
  const float dir = __fmul_rn(c.dir_coef, e);                              // sqrt(1 - a_prev - sigma^2) e
  float nz = __fmul_rn(__fmul_rn(c.sigma_t, noise[i]), c.temperature);     // sigma * noise * temperature
  if (ex.noise_mul) nz = __fmul_rn(nz, ex.noise_mul[i]);
  x_out[i] = __fadd_rn(__fadd_rn(__fmul_rn(c.sqrt_a_prev, x0), dir), nz);
  if (x0_out) x0_out[i] = x0;
  if (eps_out) eps_out[i] = e;
}
int ddim_step_launch(const MixDesc& m, const DdimCoef& c, const StepExtras& ex, const float* x, const float* noise,
                     float* x_out, float* x0_out, float* eps_out, int B, long per_sample, cudaStream_t s) {
  const long total = B * per_sample;
  if (total == 0) return 0;
  return launch_pdl(ddim_step_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, 1, m, c, ex, x, noise,
                    x_out, x0_out, eps_out, per_sample, total) == cudaSuccess ? 0 : 1;
}

__global__ void ddpm_step_kernel(const MixDesc m, const DdpmCoef c, const StepExtras ex, const float* __restrict__ x,
                                 const float* __restrict__ noise, float* __restrict__ x_out,
                                 float* __restrict__ x0_out, long per_sample, long total) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int b = static_cast<int>(i / per_sample);
  float w, ow;
  mix_coeffs(m, b, w, ow);
  const float e = mix1(m, m.eps_c[i], m.eps_u ? m.eps_u[i] : 0.f, w, ow);
  const float xi = x[i];
  float x0 = __fsub_rn(__fmul_rn(c.sqrt_recip, xi), __fmul_rn(c.sqrt_recipm1, e));
  if (ex.x0_raw) { ex.x0_raw[i] = x0; return; }
  x0 = clip_x0_dev(x0, c.clip, ex.dyn_s, b);
  const float mean = __fadd_rn(__fmul_rn(c.coef1, x0), __fmul_rn(c.coef2, xi));
  float nz = __fmul_rn(noise[i], c.temperature);
  if (ex.noise_mul) nz = __fmul_rn(nz, ex.noise_mul[i]);
  x_out[i] = __fadd_rn(mean, __fmul_rn(c.nonzero_sigma, nz));  // nonzero_mask * exp(0.5 logvar) * noise
  if (x0_out) x0_out[i] = x0;
}
int ddpm_step_launch(const MixDesc& m, const DdpmCoef& c, const StepExtras& ex, const float* x, const float* noise,
                     float* x_out, float* x0_out, int B, long per_sample, cudaStream_t s) {
  const long total = B * per_sample;
  if (total == 0) return 0;
  return launch_pdl(ddpm_step_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, 1, m, c, ex, x, noise,
                    x_out, x0_out, per_sample, total) == cudaSuccess ? 0 : 1;
}

// s[b] = max(quantile(|x0[b, :]|, q), 1) with torch.quantile's default 'linear' interpolation
// (diffusion_utils/util.py:74-77).  One block per sample; the two order statistics around rank q (n-1) are
// found EXACTLY by a 4-pass most-significant-byte radix select on the bit patterns of |x| (non-negative
// floats order like their bits), then interpolated with torch's two-branch lerp.
__global__ void __launch_bounds__(256) quantile_abs_kernel(const float* __restrict__ x0, long n, float q,
                                                           float* __restrict__ s_out) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int sh_prefix, sh_rank;
  const float* v = x0 + blockIdx.x * n;
  const float ranks = __fmul_rn(q, static_cast<float>(n - 1));  // q * (n - 1) in fp32, like torch
  const float below = floorf(ranks);
  const float wgt = __fsub_rn(ranks, below);
  const long r_lo = static_cast<long>(below), r_hi = static_cast<long>(ceilf(ranks));
  float val[2];
  for (int which = 0; which < 2; ++which) {
    if (which == 1 && r_hi == r_lo) { val[1] = val[0]; break; }
    unsigned int prefix = 0, mask = 0;
    unsigned int rank = static_cast<unsigned int>(which == 0 ? r_lo : r_hi);  // 0-based rank among the n values
    for (int pass = 3; pass >= 0; --pass) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      for (long i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned int u = __float_as_uint(fabsf(v[i]));
        if ((u & mask) == prefix) atomicAdd(&hist[(u >> (8 * pass)) & 255u], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned int cum = 0;
        int bin = 0;
        for (; bin < 255; ++bin) {
          if (cum + hist[bin] > rank) break;
          cum += hist[bin];
        }
        sh_prefix = prefix | (static_cast<unsigned int>(bin) << (8 * pass));
        sh_rank = rank - cum;
      }
      __syncthreads();
      prefix = sh_prefix;
      rank = sh_rank;
      mask |= 0xFFu << (8 * pass);
      __syncthreads();
    }
    val[which] = __uint_as_float(prefix);
  }
  if (threadIdx.x == 0) {
    const float a = val[0], b = val[1];
    const float d = __fsub_rn(b, a);
    // at::lerp as ATen evaluates it on CPU (vectorised) and CUDA (contracted): one fused multiply-add,
    // weight < 0.5 ? fma(w, b - a, a) : fma(w - 1, b - a, b)   [checked against torch.quantile bit for bit]
    const float r = wgt < 0.5f ? __fmaf_rn(wgt, d, a) : __fmaf_rn(__fsub_rn(wgt, 1.0f), d, b);
    s_out[blockIdx.x] = fmaxf(r, 1.0f);
  }
}
int quantile_abs_launch(const float* x0, int B, long n, float q, float* s_out, cudaStream_t s) {
  if (B == 0) return 0;
  quantile_abs_kernel<<<B, 256, 0, s>>>(x0, n, q, s_out);
  return SGDM_LAUNCH_OK();
}

__global__ void to_uint8_kernel(const float* __restrict__ x, unsigned char* __restrict__ out, long n) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float v = __fmul_rn(__fadd_rn(x[i], 1.0f), 127.5f);
  v = fminf(fmaxf(v, 0.0f), 255.0f);
  out[i] = static_cast<unsigned char>(v);  // truncation, like .to(torch.uint8)
}
int to_uint8_launch(const float* x, unsigned char* out, long n, cudaStream_t s) {
  if (n == 0) return 0;
  to_uint8_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(x, out, n);
  return SGDM_LAUNCH_OK();
}


// PLMS multistep combination (ddim_plms_sampler.py:432-459): out = (c0 a0 + c1 a1 + ...) / div,
// evaluated left to right with separately rounded products, like the unfused torch expression.
// (pre_scale: out = pre_scale * (sum), the form PNDM's `(1 / 24) * (55 e1 - 59 e2 + ...)` takes, pndm_sampler.py:125)
struct LincombArgs { const float* a[4]; float c[4]; int n_terms; float div; float pre_scale; int use_scale; };
__global__ void lincomb_kernel(const LincombArgs g, float* __restrict__ out, long n) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float acc = __fmul_rn(g.c[0], g.a[0][i]);
  for (int k = 1; k < g.n_terms; ++k) acc = __fadd_rn(acc, __fmul_rn(g.c[k], g.a[k][i]));
  out[i] = g.use_scale ? __fmul_rn(g.pre_scale, acc) : __fdiv_rn(acc, g.div);
}
int lincomb_launch(const float* const* a, const float* c, int n_terms, float div, float* out, long n, cudaStream_t s,
                   int use_scale, float pre_scale) {
  if (n_terms < 1 || n_terms > 4) return 1;
  if (n == 0) return 0;
  LincombArgs g;
  for (int k = 0; k < 4; ++k) { g.a[k] = k < n_terms ? a[k] : nullptr; g.c[k] = k < n_terms ? c[k] : 0.f; }
  g.n_terms = n_terms;
  g.div = div;
  g.pre_scale = pre_scale;
  g.use_scale = use_scale;
  lincomb_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(g, out, n);
  return SGDM_LAUNCH_OK();
}

// PNDM transfer (pndm_sampler.py:128-141, Eq. 9 of the PNDM paper): x_next = x + d * (A * x - B * et) with the
// reference's separately rounded fp32 operations; d = a_next - a_t, A and B are its fp32 scalar sub-expressions.
__global__ void pndm_transfer_kernel(const float* __restrict__ x, const float* __restrict__ et, float d, float A, float B,
                                     float* __restrict__ out, long n) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float xi = x[i];
  out[i] = __fadd_rn(xi, __fmul_rn(d, __fsub_rn(__fmul_rn(A, xi), __fmul_rn(B, et[i]))));
}
int pndm_transfer_launch(const float* x, const float* et, float d, float A, float B, float* out, long n, cudaStream_t s) {
  if (n == 0) return 0;
  pndm_transfer_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(x, et, d, A, B, out, n);
  return SGDM_LAUNCH_OK();
}

// =========================================================================== weight packing
// cin_part > 0: split-precision packing (engine precision 1).  The K row of a tap then holds three parts of cin_part
// channels [w_hi | w_lo | w_hi] (w_hi = op(w), w_lo = op(w - w_hi)), zero-padded to cin_pad, matching activation rows
// [a_hi | a_hi | a_lo] (store8_split3).
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, op_t* __restrict__ dst, int Cout, int Cin, int ks,
                                        int cin_pad, int ktot, int k_off, const int* __restrict__ ci_map, int cin_part) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const int width = cin_part > 0 ? 3 * cin_part : cin_pad;
  const long total = static_cast<long>(Cout) * ks * ks * width;
  if (idx >= total) return;
  const int cj = idx % width;
  const int tap = (idx / width) % (ks * ks);
  const int co = idx / (static_cast<long>(width) * ks * ks);
  const int part = cin_part > 0 ? cj / cin_part : 0;
  const int c = cin_part > 0 ? cj - part * cin_part : cj;
  const int ci = ci_map ? ci_map[c] : (c < Cin ? c : -1);
  if (ci < 0) return;
  const int r = tap / ks, s = tap - r * ks;
  const float wv = w[((static_cast<long>(co) * Cin + ci) * ks + r) * ks + s];
  const op_t hi = to_op(wv);
  dst[static_cast<long>(co) * ktot + k_off + tap * cin_pad + cj] = part == 1 ? to_op(wv - from_op(hi)) : hi;
}
int pack_conv_weight_launch(const float* w, op_t* dst, int Cout, int Cin, int ks, int cin_pad, int ktot, int k_off,
                            const int* ci_map, cudaStream_t s, int cin_part) {
  if (cin_part > 0 && 3 * cin_part > cin_pad) return 1;
  const long total = static_cast<long>(Cout) * ks * ks * (cin_part > 0 ? 3 * cin_part : cin_pad);
  pack_conv_weight_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w, dst, Cout, Cin, ks, cin_pad,
                                                                                   ktot, k_off, ci_map, cin_part);
  return SGDM_LAUNCH_OK();
}
// Sub-pixel packing (ConvDesc::up2): dst[(2 dy + dx) Cout + co][(R * 3 + S) * cin_pad + ci] = sum of w[co][ci][r][s] over
// r in V(dy, R), s in V(dx, S) with V(0,0) = {0}, V(0,1) = {1,2}, V(1,1) = {0,1}, V(1,2) = {2}; the sums are formed in
// fp32 and rounded once.  Taps a parity does not use stay zero (they are never read).
__global__ void pack_conv_weight_up2_kernel(const float* __restrict__ w, op_t* __restrict__ dst, int Cout, int Cin, int cin_pad,
                                            int dense) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const long total = static_cast<long>(4) * Cout * 4 * Cin;
  if (idx >= total) return;
  const int ci = idx % Cin;
  const int ab = (idx / Cin) % 4;                    // (a, b): which of the parity's 2 x 2 taps
  const int co = (idx / (static_cast<long>(Cin) * 4)) % Cout;
  const int par = idx / (static_cast<long>(Cin) * 4 * Cout);
  const int dy = par >> 1, dx = par & 1, a = ab >> 1, b = ab & 1;
  const int R = dy + a, S = dx + b;
  // V(d, T): the original taps that land on low-resolution offset T - 1 for output parity d
  const int r_lo = dy == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), r_hi = dy == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
  const int s_lo = dx == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), s_hi = dx == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
  const float* wp = w + (static_cast<long>(co) * Cin + ci) * 9;
  float acc = 0.f;
  for (int r = r_lo; r <= r_hi; ++r)
    for (int s = s_lo; s <= s_hi; ++s) acc += wp[r * 3 + s];
  // dense (ConvDesc::up2 == 2): only the parity's own four taps, K = (a * 2 + b) * cin_pad + ci
  if (dense) dst[(static_cast<long>(par) * Cout + co) * (4 * cin_pad) + ab * cin_pad + ci] = to_op(acc);
  else dst[(static_cast<long>(par) * Cout + co) * (9 * cin_pad) + (R * 3 + S) * cin_pad + ci] = to_op(acc);
}
int pack_conv_weight_up2_launch(const float* w, op_t* dst, int Cout, int Cin, int cin_pad, cudaStream_t s, int dense) {
  const long total = static_cast<long>(4) * Cout * 4 * Cin;
  pack_conv_weight_up2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w, dst, Cout, Cin, cin_pad, dense);
  return SGDM_LAUNCH_OK();
}
__global__ void pack_first_conv_im2col_kernel(const float* __restrict__ w, op_t* __restrict__ dst, int Cout, int Cimg, int L) {
  const int ce = 2 * Cimg + L, idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * 9 * ce) return;
  const int j = idx % ce, tap = (idx / ce) % 9, co = idx / (9 * ce);
  const int ci = j < Cimg ? j : j < 2 * Cimg ? j - Cimg : Cimg + (j - 2 * Cimg);  // hi and lo halves share the weight
  dst[static_cast<long>(co) * 64 + tap * ce + j] = to_op(w[(static_cast<long>(co) * (Cimg + L) + ci) * 9 + tap]);
}
int pack_first_conv_im2col_launch(const float* w, op_t* dst, int Cout, int Cimg, int L, cudaStream_t s) {
  if (9 * (2 * Cimg + L) > 64) return 1;
  const int total = Cout * 9 * (2 * Cimg + L);
  pack_first_conv_im2col_kernel<<<(total + 255) / 256, 256, 0, s>>>(w, dst, Cout, Cimg, L);
  return SGDM_LAUNCH_OK();
}
// Output head with folded horizontal taps (ConvDesc::hfold): dst[(s * Cout + co)][r * cin_pad + ci] = w[co][ci][r][s],
// 16 rows of 3 * cin_pad columns (rows >= 3 * Cout and padded channels stay zero).
__global__ void pack_conv_weight_hfold_kernel(const float* __restrict__ w, op_t* __restrict__ dst, int Cout, int Cin,
                                              int cin_pad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * 9 * Cin) return;
  const int ci = idx % Cin, tap = (idx / Cin) % 9, co = idx / (9 * Cin);
  const int r = tap / 3, s = tap - 3 * r;
  dst[static_cast<long>(s * Cout + co) * (3 * cin_pad) + r * cin_pad + ci] = to_op(w[((static_cast<long>(co) * Cin + ci) * 3 + r) * 3 + s]);
}
int pack_conv_weight_hfold_launch(const float* w, op_t* dst, int Cout, int Cin, int cin_pad, cudaStream_t s) {
  if (3 * Cout > 16) return 1;
  const int total = Cout * 9 * Cin;
  pack_conv_weight_hfold_kernel<<<(total + 255) / 256, 256, 0, s>>>(w, dst, Cout, Cin, cin_pad);
  return SGDM_LAUNCH_OK();
}
// =========================================================================== parameter fingerprint
// out[t] = order-independent 64-bit hash of the bit pattern of tensor t (position-mixed, so permutations and
// single-bit changes move it).  Lets the host detect parameter writes that bypass autograd's version counter
// (`param.data.copy_`, as the reference's EMA swap does, dynamic/ema.py:46-53) with ONE launch and one small
// read-back instead of re-packing every tensor.  grid = (tensors, slices); `out` must be zeroed.
__global__ void fingerprint_kernel(const void* const* __restrict__ ptrs, const long long* __restrict__ numel,
                                   unsigned long long* __restrict__ out) {
  const int t = blockIdx.x;
  const long long n = numel[t];
  const unsigned int* p = static_cast<const unsigned int*>(ptrs[t]);
  unsigned long long h = 0;
  for (long long i = blockIdx.y * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.y) * blockDim.x) {
    unsigned long long v = static_cast<unsigned long long>(p[i]) + 0x9E3779B97F4A7C15ull * static_cast<unsigned long long>(i + 1);
    v ^= v >> 29;
    v *= 0xBF58476D1CE4E5B9ull;
    v ^= v >> 32;
    h += v;
  }
  for (int o = 16; o; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
  if ((threadIdx.x & 31) == 0 && h) atomicAdd(out + t, h);
}
int fingerprint_launch(const void* const* ptrs, const long long* numel, int n, unsigned long long* out, cudaStream_t s) {
  if (cudaMemsetAsync(out, 0, static_cast<size_t>(n) * sizeof(unsigned long long), s) != cudaSuccess) return 1;
  fingerprint_kernel<<<dim3(n, 8), 256, 0, s>>>(ptrs, numel, out);
  return SGDM_LAUNCH_OK();
}

__global__ void add_bias_kernel(const float* a, const float* b, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (a ? a[i] : 0.f) + (b ? b[i] : 0.f);
}
int add_bias_launch(const float* a, const float* b, float* out, int n, cudaStream_t s) {
  add_bias_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, b, out, n);
  return SGDM_LAUNCH_OK();
}

}  // namespace sgdm
