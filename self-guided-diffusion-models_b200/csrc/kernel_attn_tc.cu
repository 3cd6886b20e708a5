// Self-attention on tcgen05 for the UNet's attention sites (T = 256 tokens, head dim 64).
//
// Replaces QKVAttentionLegacy (openaimodel.py:403-420): per (sample, head)
//   S = (q ch^-1/4)(k ch^-1/4)^T,  P = softmax_fp32(S),  O = P V
// One persistent CTA per SM walks (sample, head) pairs:
//   warp 0 (one lane)  TMA producer: Q, K, V tiles [256 x 64] of the packed qkv matrix, 2-stage ring
//   warp 1 (one lane)  MMA issuer  : S_j = Q_j K^T (M128 x N256 x K64) for both query tiles into TMEM,
//                                    then O_j = P_j V (M128 x N64 x K256, V as an MN-major operand)
//   warps 2..5         softmax     : thread = query row; exact two-pass softmax over all 256 keys read from
//                                    TMEM (row max, then exp2 / row sum), P written as the 16-bit K-major
//                                    operand INTO the shared memory of Q and K (dead once S is complete),
//                                    O read back from TMEM, scaled by 1/row-sum and stored
// TMEM: S_0 in columns [0, 256), S_1 in [256, 512); O_j overwrites the first 64 columns of S_j.
// Roofline: the softmax (256 x 256 exponentials per pair on the 16-lane MUFU) bounds it at ~4 k clk per
// pair, the MMAs need ~2 k: FLOPs per launch = 4 * B * heads * T * T * D.
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

#include "attn.cuh"
#include "conv.cuh"

namespace sgdm {

constexpr int kTcT = 256, kTcD = 64;
constexpr int kTcTile = kTcT * kTcD * 2;   // 32 KB: one [256 x 64] 16-bit tile

struct alignas(64) AttnTcParams {
  CUtensorMap tm;  // the packed [B*T, row_stride] matrix holding q, k and v column blocks
  op_t* out;
  long o_row_stride;
  int pairs, heads;
  int q_col, k_col, v_col, head_stride;  // column of head 0 / per-head column step
  float scale_log2;                      // logits scale * log2(e)
};

__device__ __forceinline__ void sts128u_(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------------------
// TWO softmax warpgroups, one per query tile, running concurrently (10 warps).
//   warp 0 (one lane)  TMA producer: Q|K (64 KB) and V (32 KB) of the next pair as soon as the MMAs that read
//                                    them have retired (Q, K: after both S MMAs; V: after both PV MMAs)
//   warp 1 (one lane)  MMA issuer  : an event loop over the two query tiles; per tile strictly S_j, PV_j, S_j, ...
//   warps 2..5 / 6..9  softmax of query tile 0 / 1: thread = query row, exact two-pass softmax from TMEM, P_j
//                                    into its OWN 64 KB operand buffer, O_j read back, scaled, stored
// (A first generation, removed in round 2, ran the two tiles back to back on four warps and spent most of a pair waiting
// on the TMEM-load round trips of one warp per scheduler: 1.55 vs 1.35 ms per step.)  With a warp of each group on every scheduler the round trips
// of one tile hide behind the arithmetic of the other, and S / PV of one tile overlap the softmax of the other.
// Shared memory: Q | K | V | P_0 | P_1 = 224 KB (single Q/K/V stage: the loads of the next pair are issued a whole
// softmax ahead of their first use, so a second stage would buy nothing).
// kHalves = 2 (round 2, the default): TWO threads per query row — 8 softmax warps per tile, 18 warps per CTA.  The
// source-level profile of the one-thread-per-row version (profiles/r02_ncu_src_attn_tc2.txt) showed pass 2 issuing in
// 18 % of its cycles: 39 % fixed-latency dependency stalls, 20 % scoreboard (MUFU / TMEM results) with two softmax warps
// per scheduler to choose from.  Each thread now owns 128 of the row's 256 keys (P blocks 2 h, 2 h + 1) and 32 of the 64
// output dims; the row maximum and the row sum are exchanged through 2 KB of shared memory under a 64-thread named
// barrier per (tile, lane quarter).
constexpr int kTc2Threads = 320;
constexpr int kTc2P = 4 * 16384;  // P_j: four K-major [128 x 64] blocks
constexpr int kTc2Xch = 2 * 128 * 2 * 4;  // [tile][row][half] fp32 exchange slots
constexpr int kTc2Smem = 3 * kTcTile + 2 * kTc2P + 256 + kTc2Xch;

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int kHalves>
__global__ void __launch_bounds__(64 + 256 * kHalves, 1) attn_tc2_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  pdl_launch_dependents();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kTcTile + 2 * kTc2P);
  uint64_t* qk_full = bars;        // TMA -> MMA
  uint64_t* v_full = bars + 1;     // TMA -> MMA
  uint64_t* qk_free = bars + 2;    // both S MMAs of a pair retired -> TMA
  uint64_t* v_free = bars + 3;     // both PV MMAs of a pair retired -> TMA
  uint64_t* s_ready = bars + 4;    // [2] S_j complete -> softmax group j
  uint64_t* p_ready = bars + 6;    // [2] P_j staged (4 warps) -> MMA
  uint64_t* o_ready = bars + 8;    // [2] O_j complete -> softmax group j
  uint64_t* tfree = bars + 10;     // [2] O_j read out of TMEM (4 warps) -> MMA (S_j of the next pair)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm);
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    mbar_init(qk_free, 1);
    mbar_init(v_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_ready[i], 4 * kHalves);
      mbar_init(&o_ready[i], 1);
      mbar_init(&tfree[i], 4 * kHalves);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int n_my = blockIdx.x < p.pairs ? (p.pairs - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_my; ++it) {
        const int pair = blockIdx.x + it * gridDim.x;
        const int n = pair / p.heads, h = pair - n * p.heads;
        const uint32_t ph = it & 1;
        mbar_wait(qk_free, ph ^ 1);
        mbar_arrive_expect_tx(qk_full, 2 * kTcTile);
        tma_load_2d(&p.tm, qk_full, smem, p.q_col + h * p.head_stride, n * kTcT);
        tma_load_2d(&p.tm, qk_full, smem + kTcTile, p.k_col + h * p.head_stride, n * kTcT);
        mbar_wait(v_free, ph ^ 1);
        mbar_arrive_expect_tx(v_full, kTcTile);
        tma_load_2d(&p.tm, v_full, smem + 2 * kTcTile, p.v_col + h * p.head_stride, n * kTcT);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc(128, 256);
      const uint32_t idesc_o = umma_idesc(128, 64) | (1u << 16);  // B operand (V) is MN-major
      const uint32_t sbase = smem_u32(smem);
      // per query tile: index of the next pair and whether its S has been issued (then PV is next)
      int it_j[2] = {0, 0};
      bool s_done[2] = {false, false};
      uint32_t idle = 0;
      while (it_j[0] < n_my || it_j[1] < n_my) {
        bool progressed = false;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (it_j[j] >= n_my) continue;
          const uint32_t ph = it_j[j] & 1;
          if (!s_done[j]) {
            // S_j = Q_j K^T: needs this pair's Q, K and the TMEM columns of tile j (O_j of the previous pair drained)
            if (!mbar_try_wait(qk_full, ph) || !mbar_try_wait(&tfree[j], ph ^ 1)) continue;
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + j * 256, umma_smem_desc(sbase + j * 16384 + k * 32),
                       umma_smem_desc(sbase + kTcTile + k * 32), idesc_s, k != 0 ? 1u : 0u);
            umma_commit(&s_ready[j]);
            s_done[j] = true;
            // Q and K may be overwritten once the S MMAs of BOTH tiles of this pair have retired
            if (s_done[j ^ 1] ? it_j[j ^ 1] == it_j[j] : it_j[j ^ 1] > it_j[j]) umma_commit(qk_free);
            progressed = true;
          } else {
            // O_j = P_j V: P_j staged by the softmax group (which has then also finished reading S_j)
            if (!mbar_try_wait(&p_ready[j], ph) || !mbar_try_wait(v_full, ph)) continue;
            tc_fence_after();
            const uint32_t pbase = sbase + 3 * kTcTile + j * kTc2P;
#pragma unroll
            for (int kk = 0; kk < 16; ++kk)
              umma_f16(tmem_base + j * 256, umma_smem_desc(pbase + (kk >> 2) * 16384 + (kk & 3) * 32),
                       umma_smem_desc(sbase + 2 * kTcTile + kk * 2048), idesc_o, kk != 0 ? 1u : 0u);
            umma_commit(&o_ready[j]);
            s_done[j] = false;
            ++it_j[j];
            // V may be overwritten once the PV MMAs of both tiles of this pair have retired
            if (it_j[j ^ 1] >= it_j[j]) umma_commit(v_free);
            progressed = true;
          }
        }
        if (progressed) idle = 0;
        else if (++idle > (1u << 27)) __trap();  // bounded like mbar_wait: a protocol bug must not hang the GPU
      }
    }
  } else {
    const int sw = warp - 2;                 // softmax warp 0 .. 8 kHalves - 1
    const int j = sw / (4 * kHalves);        // query tile
    const int half = (sw >> 2) % kHalves;    // which half of the row's keys / output dims this thread owns
    const int quarter = warp & 3;            // == TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;       // query row within the tile == TMEM lane
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + j * 256;
    const uint32_t x7 = r & 7;
    const uint32_t pbase = smem_u32(smem) + 3 * kTcTile + j * kTc2P + r * 128;
    constexpr int NC = 8 / kHalves;          // 32-key chunks per thread
    const int c0 = half * NC;
    float* xch = reinterpret_cast<float*>(smem + 3 * kTcTile + 2 * kTc2P + 256) + (j * 128 + r) * 2;
    const int bar_id = 1 + j * 4 + quarter;  // the two warps that share this tile's lane quarter
    for (int it = 0; it < n_my; ++it) {
      const int pair = blockIdx.x + it * gridDim.x;
      const uint32_t ph = it & 1;
      const int n = pair / p.heads, h = pair - n * p.heads;
      mbar_wait(&s_ready[j], ph);
      tc_fence_after();
      // Both passes keep the TMEM load of the next 32 columns in flight while the current 32 are processed
      // (two register buffers; tcgen05.wait::ld waits for everything outstanding, hence the issue order).
      // pass 1: row maximum over the 256 keys, four independent chains
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      uint32_t va[32], vb[32];
      tmem_ld_32x32(taddr + 32 * c0, va);
#pragma unroll
      for (int c = 0; c < NC; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 32 * (c0 + c + 1), vb);
#pragma unroll
        for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(va[i]));
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 32 * (c0 + ((c + 2) % NC)), va);  // after the last chunk: the first again, for pass 2
#pragma unroll
        for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(vb[i]));
      }
      float mrow = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      if (kHalves == 2) {  // row maximum over both halves
        xch[half] = mrow;
        named_bar_sync(bar_id, 64);
        mrow = fmaxf(mrow, xch[half ^ 1]);
      }
      const float mb = mrow * p.scale_log2;
      // pass 2: exponentials (log2 domain), row sum, P as the 16-bit K-major operand
      float l4[4] = {0.f, 0.f, 0.f, 0.f};  // four independent partial row sums
      auto exp_chunk = [&](const uint32_t(&v)[32], int c) {
        const uint32_t blk = pbase + (c >> 1) * 16384;  // K block of 64 keys
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            e[i] = ex2f(__uint_as_float(v[8 * q + i]) * p.scale_log2 - mb);
            l4[i & 3] += e[i];
          }
          const uint32_t ci = (c & 1) * 4 + q;  // 16-byte chunk (8 keys) inside the 128-byte row
          sts128u_(blk + ((ci ^ x7) << 4), make_uint4(pack_op2(e[0], e[1]), pack_op2(e[2], e[3]), pack_op2(e[4], e[5]),
                                                      pack_op2(e[6], e[7])));
        }
      };
#pragma unroll
      for (int c = 0; c < NC; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 32 * (c0 + c + 1), vb);
        exp_chunk(va, c0 + c);
        tmem_ld_wait();
        if (c + 2 < NC) tmem_ld_32x32(taddr + 32 * (c0 + c + 2), va);
        exp_chunk(vb, c0 + c + 1);
      }
      float lrow = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[j]);
      if (kHalves == 2) {
        // row sum over both halves, through the same slots: the partner has read the maxima (it is past its pass 2 when it
        // arrives here), and the next pair's maxima are written only after both warps have released the tile (tfree)
        named_bar_sync(bar_id, 64);
        xch[half] = lrow;
        named_bar_sync(bar_id, 64);
        lrow += xch[half ^ 1];
      }
      const float inv_l = 1.0f / lrow;
      mbar_wait(&o_ready[j], ph);
      tc_fence_after();
      {
        // O_j row -> * 1/l -> this thread's 64 / kHalves output dims (16-bit)
        constexpr int ND = 2 / kHalves;  // 32-column TMEM loads per thread
        op_t* dst = p.out + (static_cast<long>(n) * kTcT + j * 128 + r) * p.o_row_stride + h * kTcD + (kHalves == 2 ? 32 * half : 0);
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(taddr + (kHalves == 2 ? 32 * half : 0), v0);
        if (ND == 2) tmem_ld_32x32(taddr + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tfree[j]);  // the values are in registers: tile j's TMEM columns are free
#pragma unroll
        for (int q = 0; q < ND; ++q) {
          const uint32_t(&v)[32] = q == 0 ? v0 : v1;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 o = make_uint4(pack_op2(__uint_as_float(v[8 * c]) * inv_l, __uint_as_float(v[8 * c + 1]) * inv_l),
                                       pack_op2(__uint_as_float(v[8 * c + 2]) * inv_l, __uint_as_float(v[8 * c + 3]) * inv_l),
                                       pack_op2(__uint_as_float(v[8 * c + 4]) * inv_l, __uint_as_float(v[8 * c + 5]) * inv_l),
                                       pack_op2(__uint_as_float(v[8 * c + 6]) * inv_l, __uint_as_float(v[8 * c + 7]) * inv_l));
            *reinterpret_cast<uint4*>(dst + 32 * q + 8 * c) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------
// Attention_LR on tcgen05 (crossattetion_lr.py:88-137): multi-query attention — ONE shared K/V head for all heads —
// over T = 256 self keys plus up to 32 "extra" keys (16 context tokens + the null key) that are the same for every
// head of a sample.  273 keys do not fit the two-tiles-in-flight TMEM plan of attn_tc2 (2 x (256 + 32) columns > 512)
// and P [128 x 288] x 2 does not fit shared memory, so the extra keys travel as an EXACT second softmax block:
//   1. S_x = Q_j K_ext^T (M128 x N32) first; the softmax group reads its 32 columns, keeps m_x, the exponentials
//      e_x[k] = 2^(s_x[k] - m_x) and their sum l_x in REGISTERS and releases the columns;
//   2. S = Q_j K^T (M128 x N256) over the same columns; pass 1 gives m_s, and m = max(m_s, m_x) is the exact row
//      maximum over all keys BEFORE any probability of the main block is formed: pass 2 runs as in attn_tc2;
//   3. P_x = e_x * 2^(m_x - m) becomes a fifth operand block [128 x 32] (K-major, SWIZZLE_64B) written over the dead Q_j
//      tile, l = l_s + l_x 2^(m_x - m), and O_j = P_j V + P_x V_ext accumulate in ONE TMEM tile (16 + 2 MMAs).
// Shared memory: Q | K | V | P_0 | P_1 = 224 KB as in attn_tc2; K_ext (4 KB) is staged in the first bytes of P_1 (dead
// until group 1's pass 2, which starts after every S_x MMA has retired), V_ext (4 KB) in the upper half of the dead Q_0
// tile.  Q / K_ext / V of the next pair are loaded when both PV MMAs have retired, K when both S MMAs have.
constexpr int kLrExtRows = 32;                       // extra keys, zero-padded by TMA
constexpr int kLrExtTile = kLrExtRows * kTcD * 2;    // 4 KB
constexpr int kLrPx = 128 * kLrExtRows * 2;          // P_x: [128 x 32] 16-bit, 64-byte rows = 8 KB

struct alignas(64) AttnLrParams {
  CUtensorMap tm;    // packed [B*T, row_stride] matrix: q heads, then the shared k and v column blocks
  CUtensorMap tmKx;  // k_extra [B][n_extra][64] as a 3-D map, box 64 x 32 x 1 (rows >= n_extra are zero-filled)
  CUtensorMap tmVx;
  op_t* out;
  long o_row_stride;
  int pairs, heads;
  int q_col, k_col, v_col, q_head_stride;
  int n_extra;
  float scale_log2;
};

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* d, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(d)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__global__ void __launch_bounds__(kTc2Threads, 1) attn_lr_tc_kernel(const __grid_constant__ AttnLrParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  pdl_launch_dependents();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kTcTile + 2 * kTc2P);
  uint64_t* q_full = bars;          // Q (both tiles) + K_ext landed
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* vx_full = bars + 3;
  uint64_t* s_free = bars + 4;      // both S MMAs of a pair retired: K may be reloaded, Q tiles are dead
  uint64_t* pv_free = bars + 5;     // both PV MMAs of a pair retired: Q / K_ext / V / V_ext regions may be reloaded
  uint64_t* sx_ready = bars + 6;    // [2] S_x of tile j complete -> softmax group j
  uint64_t* sx_done = bars + 8;     // [2] group j has read S_x (4 warps) -> MMA (S over the same columns)
  uint64_t* s_ready = bars + 10;    // [2]
  uint64_t* p_ready = bars + 12;    // [2] P_j and P_x staged (4 warps)
  uint64_t* o_ready = bars + 14;    // [2]
  uint64_t* tfree = bars + 16;      // [2] O_j read out of TMEM (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTcTile;
  uint8_t* sV = smem + 2 * kTcTile;
  uint8_t* sP = smem + 3 * kTcTile;                 // P_0 | P_1
  uint8_t* sKx = sP + kTc2P;                        // K_ext: first 4 KB of P_1
  uint8_t* sVx = sQ + kLrPx;                        // V_ext: upper half of the Q_0 tile (P_x of tile 0 takes the lower)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm);
    tma_prefetch_desc(&p.tmKx);
    tma_prefetch_desc(&p.tmVx);
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sx_ready[i], 1);
      mbar_init(&sx_done[i], 4);
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_ready[i], 4);
      mbar_init(&o_ready[i], 1);
      mbar_init(&tfree[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int n_my = blockIdx.x < p.pairs ? (p.pairs - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_my; ++it) {
        const int pair = blockIdx.x + it * gridDim.x;
        const int n = pair / p.heads, h = pair - n * p.heads;
        const uint32_t ph = it & 1;
        mbar_wait(s_free, ph ^ 1);  // previous pair's S MMAs retired: K is free
        mbar_arrive_expect_tx(k_full, kTcTile);
        tma_load_2d(&p.tm, k_full, sK, p.k_col, n * kTcT);
        mbar_wait(pv_free, ph ^ 1);  // previous pair's PV MMAs retired: Q tiles (P_x, V_ext), P_1 (K_ext) and V are free
        mbar_arrive_expect_tx(q_full, kTcTile + kLrExtTile);
        tma_load_2d(&p.tm, q_full, sQ, p.q_col + h * p.q_head_stride, n * kTcT);
        tma_load_3d(&p.tmKx, q_full, sKx, 0, 0, n);
        mbar_arrive_expect_tx(v_full, kTcTile);
        tma_load_2d(&p.tm, v_full, sV, p.v_col, n * kTcT);
        mbar_wait(s_free, ph);  // THIS pair's S MMAs retired: the Q_0 tile is dead -> V_ext into its upper half
        mbar_arrive_expect_tx(vx_full, kLrExtTile);
        tma_load_3d(&p.tmVx, vx_full, sVx, 0, 0, n);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_x = umma_idesc(128, kLrExtRows);
      const uint32_t idesc_s = umma_idesc(128, 256);
      const uint32_t idesc_o = umma_idesc(128, 64) | (1u << 16);  // B operand (V) is MN-major
      const uint64_t hi64 = umma_smem_desc_hi(true);               // P_x: 64-byte rows, SWIZZLE_64B
      const uint32_t sbase = smem_u32(smem);
      int it_j[2] = {0, 0};
      int st_j[2] = {0, 0};  // 0: S_x next | 1: S next | 2: PV next
      uint32_t idle = 0;
      while (it_j[0] < n_my || it_j[1] < n_my) {
        bool progressed = false;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (it_j[j] >= n_my) continue;
          const uint32_t ph = it_j[j] & 1;
          const uint32_t d_tmem = tmem_base + j * 256;
          if (st_j[j] == 0) {
            if (!mbar_try_wait(q_full, ph) || !mbar_try_wait(&tfree[j], ph ^ 1)) continue;
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(d_tmem, umma_smem_desc(sbase + j * 16384 + k * 32), umma_smem_desc(smem_u32(sKx) + k * 32), idesc_x,
                       k != 0 ? 1u : 0u);
            umma_commit(&sx_ready[j]);
            st_j[j] = 1;
            progressed = true;
          } else if (st_j[j] == 1) {
            // S_j over the columns S_x occupied: the group must have read S_x.  Tile 1 additionally waits until tile 0's
            // S_x of this pair has been issued: K_ext lives in P_1, which group 1 overwrites after s_ready[1]
            if (j == 1 && !(it_j[0] > it_j[1] || (it_j[0] == it_j[1] && st_j[0] >= 1))) continue;
            if (!mbar_try_wait(&sx_done[j], ph) || !mbar_try_wait(k_full, ph)) continue;
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(d_tmem, umma_smem_desc(sbase + j * 16384 + k * 32), umma_smem_desc(smem_u32(sK) + k * 32), idesc_s,
                       k != 0 ? 1u : 0u);
            umma_commit(&s_ready[j]);
            st_j[j] = 2;
            // both S of this pair issued -> K (and, for the V_ext load, the Q tiles) free once they retire
            if (it_j[j ^ 1] > it_j[j] || (it_j[j ^ 1] == it_j[j] && st_j[j ^ 1] == 2)) umma_commit(s_free);
            progressed = true;
          } else {
            if (!mbar_try_wait(&p_ready[j], ph) || !mbar_try_wait(v_full, ph) || !mbar_try_wait(vx_full, ph)) continue;
            tc_fence_after();
            const uint32_t pbase = sbase + 3 * kTcTile + j * kTc2P;
#pragma unroll
            for (int kk = 0; kk < 16; ++kk)
              umma_f16(d_tmem, umma_smem_desc(pbase + (kk >> 2) * 16384 + (kk & 3) * 32),
                       umma_smem_desc(smem_u32(sV) + kk * 2048), idesc_o, kk != 0 ? 1u : 0u);
            const uint32_t px = sbase + j * 16384;  // P_x over the dead Q_j tile
#pragma unroll
            for (int kk = 0; kk < kLrExtRows / 16; ++kk)
              umma_f16(d_tmem, hi64 | umma_smem_desc_lo(px + kk * 32), umma_smem_desc(smem_u32(sVx) + kk * 2048), idesc_o, 1u);
            umma_commit(&o_ready[j]);
            st_j[j] = 0;
            ++it_j[j];
            if (it_j[j ^ 1] >= it_j[j]) umma_commit(pv_free);
            progressed = true;
          }
        }
        if (progressed) idle = 0;
        else if (++idle > (1u << 27)) __trap();
      }
    }
  } else {
    const int j = warp >= 6 ? 1 : 0;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + j * 256;
    const uint32_t x7 = r & 7;
    const uint32_t pbase = smem_u32(smem) + 3 * kTcTile + j * kTc2P + r * 128;
    const uint32_t pxrow = smem_u32(smem) + j * 16384 + r * 64;  // P_x row: 64 bytes, chunk c at c ^ ((r >> 1) & 3)
    const uint32_t x3 = (r >> 1) & 3;
    for (int it = 0; it < n_my; ++it) {
      const int pair = blockIdx.x + it * gridDim.x;
      const uint32_t ph = it & 1;
      const int n = pair / p.heads, h = pair - n * p.heads;
      // ---- the extra keys: S_x -> m_x, e_x (registers), l_x
      mbar_wait(&sx_ready[j], ph);
      tc_fence_after();
      float ex[kLrExtRows];
      float mxb, lx = 0.f;
      {
        uint32_t v[32];
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sx_done[j]);  // the values are in registers: S may overwrite the columns
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < kLrExtRows; ++i)
          if (i < p.n_extra) m = fmaxf(m, __uint_as_float(v[i]));
        mxb = m * p.scale_log2;
#pragma unroll
        for (int i = 0; i < kLrExtRows; ++i) {
          ex[i] = i < p.n_extra ? ex2f(__uint_as_float(v[i]) * p.scale_log2 - mxb) : 0.f;
          lx += ex[i];
        }
      }
      // ---- the 256 self keys, as in attn_tc2, with the row maximum taken over BOTH blocks
      mbar_wait(&s_ready[j], ph);
      tc_fence_after();
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      uint32_t va[32], vb[32];
      tmem_ld_32x32(taddr, va);
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 32 * (c + 1), vb);
#pragma unroll
        for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(va[i]));
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 32 * ((c + 2) & 7), va);
#pragma unroll
        for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(vb[i]));
      }
      const float mb = fmaxf(fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2, mxb);
      // P_x = e_x 2^(m_x - m) into the dead Q_j tile (S of tile j has retired: s_ready[j]); 4 chunks of 8 keys per row
      const float fx = ex2f(mxb - mb);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        sts128u_(pxrow + ((c ^ x3) << 4),
                 make_uint4(pack_op2(ex[8 * c] * fx, ex[8 * c + 1] * fx), pack_op2(ex[8 * c + 2] * fx, ex[8 * c + 3] * fx),
                            pack_op2(ex[8 * c + 4] * fx, ex[8 * c + 5] * fx), pack_op2(ex[8 * c + 6] * fx, ex[8 * c + 7] * fx)));
      float l4[4] = {lx * fx, 0.f, 0.f, 0.f};
      auto exp_chunk = [&](const uint32_t(&v)[32], int c) {
        const uint32_t blk = pbase + (c >> 1) * 16384;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            e[i] = ex2f(__uint_as_float(v[8 * q + i]) * p.scale_log2 - mb);
            l4[i & 3] += e[i];
          }
          const uint32_t ci = (c & 1) * 4 + q;
          sts128u_(blk + ((ci ^ x7) << 4), make_uint4(pack_op2(e[0], e[1]), pack_op2(e[2], e[3]), pack_op2(e[4], e[5]),
                                                      pack_op2(e[6], e[7])));
        }
      };
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 32 * (c + 1), vb);
        exp_chunk(va, c);
        tmem_ld_wait();
        if (c + 2 < 8) tmem_ld_32x32(taddr + 32 * (c + 2), va);
        exp_chunk(vb, c + 1);
      }
      const float inv_l = 1.0f / ((l4[0] + l4[1]) + (l4[2] + l4[3]));
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[j]);
      mbar_wait(&o_ready[j], ph);
      tc_fence_after();
      {
        op_t* dst = p.out + (static_cast<long>(n) * kTcT + j * 128 + r) * p.o_row_stride + h * kTcD;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(taddr, v0);
        tmem_ld_32x32(taddr + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tfree[j]);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t(&v)[32] = half == 0 ? v0 : v1;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 o = make_uint4(pack_op2(__uint_as_float(v[8 * c]) * inv_l, __uint_as_float(v[8 * c + 1]) * inv_l),
                                       pack_op2(__uint_as_float(v[8 * c + 2]) * inv_l, __uint_as_float(v[8 * c + 3]) * inv_l),
                                       pack_op2(__uint_as_float(v[8 * c + 4]) * inv_l, __uint_as_float(v[8 * c + 5]) * inv_l),
                                       pack_op2(__uint_as_float(v[8 * c + 6]) * inv_l, __uint_as_float(v[8 * c + 7]) * inv_l));
            *reinterpret_cast<uint4*>(dst + 32 * half + 8 * c) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static PFN_cuTensorMapEncodeTiled_v12000 attn_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}
// [B][n_extra][64] 16-bit -> box 64 x 32 x 1, SWIZZLE_128B; rows >= n_extra are zero-filled by the TMA unit
static int encode_extra_map(CUtensorMap* tm, const op_t* base, int B, int n_extra) {
  auto fn = attn_encode_fn();
  if (!fn) return 1;
  cuuint64_t dims[3] = {(cuuint64_t)kTcD, (cuuint64_t)n_extra, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)kTcD * 2, (cuuint64_t)n_extra * kTcD * 2};
  cuuint32_t box[3] = {(cuuint32_t)kTcD, (cuuint32_t)kLrExtRows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
#ifdef SGDM_OPERAND_BF16
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#else
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#endif
  return fn(tm, dt, 3, const_cast<op_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0 : 1;
}

bool attn_lr_tc_applicable(const AttnDesc& a) {
  if (a.T != kTcT || a.D != kTcD || a.n_extra < 1 || a.n_extra > kLrExtRows || !a.k_extra || !a.v_extra) return false;
  if (a.k_head_stride != 0 || a.v_head_stride != 0 || a.q_head_stride <= 0) return false;  // multi-query
  if (a.q_row_stride != a.k_row_stride || a.q_row_stride != a.v_row_stride) return false;
  const long kq = a.k - a.q, vq = a.v - a.q;
  if (kq < 0 || vq < 0 || kq >= a.q_row_stride || vq >= a.q_row_stride) return false;
  if ((a.q_row_stride % 8) || (a.q_head_stride % 8) || (kq % 8) || (vq % 8) || (a.o_row_stride % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(a.k_extra) | reinterpret_cast<uintptr_t>(a.v_extra)) & 15) return false;
  return true;
}

int attn_lr_tc_launch(const AttnDesc& a, cudaStream_t s) {
  AttnLrParams p;
  char err[256];
  if (encode_matrix_map(&p.tm, a.q, false, static_cast<long>(a.B) * a.T, static_cast<int>(a.q_row_stride), kTcD, 128, err,
                        sizeof(err), kTcT))
    return 1;
  if (encode_extra_map(&p.tmKx, a.k_extra, a.B, a.n_extra) || encode_extra_map(&p.tmVx, a.v_extra, a.B, a.n_extra)) return 1;
  p.out = a.out;
  p.o_row_stride = a.o_row_stride;
  p.pairs = a.B * a.heads;
  p.heads = a.heads;
  p.q_col = 0;
  p.k_col = static_cast<int>(a.k - a.q);
  p.v_col = static_cast<int>(a.v - a.q);
  p.q_head_stride = a.q_head_stride;
  p.n_extra = a.n_extra;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(attn_lr_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTc2Smem) != cudaSuccess) return 1;
    attr_set = true;
  }
  const int grid = p.pairs < kNumSMs ? p.pairs : kNumSMs;
  return launch_pdl(attn_lr_tc_kernel, dim3(grid), dim3(kTc2Threads), kTc2Smem, s, 1, p) == cudaSuccess ? 0 : 1;
}

// The tcgen05 path covers the UNet's self-attention shape; everything else stays on the mma.sync kernel.
bool attn_tc_applicable(const AttnDesc& a) {
  if (a.T != kTcT || a.D != kTcD || a.n_extra != 0) return false;
  if (a.q_row_stride != a.k_row_stride || a.q_row_stride != a.v_row_stride) return false;
  if (a.q_head_stride != a.k_head_stride || a.q_head_stride != a.v_head_stride || a.q_head_stride <= 0) return false;
  const long kq = a.k - a.q, vq = a.v - a.q;  // all three are column blocks of one packed matrix
  if (kq < 0 || vq < 0 || kq >= a.q_row_stride || vq >= a.q_row_stride) return false;
  if ((a.q_row_stride % 8) || (a.q_head_stride % 8) || (kq % 8) || (vq % 8) || (a.o_row_stride % 8)) return false;
  return true;
}

int attn_tc_launch(const AttnDesc& a, cudaStream_t s) {
  AttnTcParams p;
  char err[256];
  if (encode_matrix_map(&p.tm, a.q, false, static_cast<long>(a.B) * a.T, static_cast<int>(a.q_row_stride), kTcD, 128, err,
                        sizeof(err), kTcT))
    return 1;
  p.out = a.out;
  p.o_row_stride = a.o_row_stride;
  p.pairs = a.B * a.heads;
  p.heads = a.heads;
  p.q_col = 0;
  p.k_col = static_cast<int>(a.k - a.q);
  p.v_col = static_cast<int>(a.v - a.q);
  p.head_stride = a.q_head_stride;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(attn_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTc2Smem) != cudaSuccess ||
        cudaFuncSetAttribute(attn_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTc2Smem) != cudaSuccess)
      return 1;
    attr_set = true;
  }
  const int grid = p.pairs < kNumSMs ? p.pairs : kNumSMs;
  static const int halves = [] { const char* ev = getenv("SGDM_ATTN_HALVES"); return ev && atoi(ev) == 1 ? 1 : 2; }();
  const cudaError_t e = halves == 2 ? launch_pdl(attn_tc2_kernel<2>, dim3(grid), dim3(64 + 512), kTc2Smem, s, 1, p)
                                    : launch_pdl(attn_tc2_kernel<1>, dim3(grid), dim3(64 + 256), kTc2Smem, s, 1, p);
  return e == cudaSuccess ? 0 : 1;
}

}  // namespace sgdm
