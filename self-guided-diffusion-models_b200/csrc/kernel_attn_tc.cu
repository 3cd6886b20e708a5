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

#include "attn.cuh"
#include "conv.cuh"

namespace sgdm {

constexpr int kTcT = 256, kTcD = 64;
constexpr int kTcTile = kTcT * kTcD * 2;   // 32 KB: one [256 x 64] 16-bit tile
constexpr int kTcStage = 3 * kTcTile;      // Q | K | V
constexpr int kTcSmem = 2 * kTcStage + 256;
constexpr int kTcThreads = 192;

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

__global__ void __launch_bounds__(kTcThreads, 1) attn_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kTcStage);
  uint64_t* full = bars;            // [2] TMA -> MMA
  uint64_t* stage_free = bars + 2;  // [2] MMA (O_1 done) -> TMA
  uint64_t* s_ready = bars + 4;     // both S tiles complete -> softmax
  uint64_t* p_ready = bars + 5;     // [2] P_j staged (4 warps) -> MMA
  uint64_t* o_ready = bars + 7;     // [2] O_j complete -> softmax / epilogue
  uint64_t* tfree = bars + 9;       // TMEM drained (4 warps) -> MMA of the next pair
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&stage_free[i], 1);
      mbar_init(&p_ready[i], 4);
      mbar_init(&o_ready[i], 1);
    }
    mbar_init(s_ready, 1);
    mbar_init(tfree, 4);
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

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int pair = blockIdx.x; pair < p.pairs; pair += gridDim.x, ++it) {
        const uint32_t st = it & 1;
        const int n = pair / p.heads, h = pair - n * p.heads;
        mbar_wait(&stage_free[st], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[st], kTcStage);
        uint8_t* base = smem + st * kTcStage;
        tma_load_2d(&p.tm, &full[st], base, p.q_col + h * p.head_stride, n * kTcT);
        tma_load_2d(&p.tm, &full[st], base + kTcTile, p.k_col + h * p.head_stride, n * kTcT);
        tma_load_2d(&p.tm, &full[st], base + 2 * kTcTile, p.v_col + h * p.head_stride, n * kTcT);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc(128, 256);
      const uint32_t idesc_o = umma_idesc(128, 64) | (1u << 16);  // B operand (V) is MN-major
      uint32_t it = 0;
      for (int pair = blockIdx.x; pair < p.pairs; pair += gridDim.x, ++it) {
        const uint32_t st = it & 1, ph = it & 1;
        const uint32_t base = smem_u32(smem + st * kTcStage);
        mbar_wait(&full[st], (it >> 1) & 1);
        mbar_wait(tfree, ph ^ 1);  // the previous pair's O tiles have been read out of TMEM
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + j * 256, umma_smem_desc(base + j * 16384 + k * 32), umma_smem_desc(base + kTcTile + k * 32),
                     idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(s_ready);
        for (int j = 0; j < 2; ++j) {
          mbar_wait(&p_ready[j], ph);
          tc_fence_after();
          // P_j: four K-major [128 x 64] blocks of 16 KB laid over the Q and K tiles; V: rows = keys
#pragma unroll
          for (int kk = 0; kk < 16; ++kk)
            umma_f16(tmem_base + j * 256, umma_smem_desc(base + (kk >> 2) * 16384 + (kk & 3) * 32),
                     umma_smem_desc(base + 2 * kTcTile + kk * 2048), idesc_o, kk != 0 ? 1u : 0u);
          umma_commit(&o_ready[j]);
        }
        umma_commit(&stage_free[st]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t x7 = r & 7;
    uint32_t it = 0;
    for (int pair = blockIdx.x; pair < p.pairs; pair += gridDim.x, ++it) {
      const uint32_t st = it & 1, ph = it & 1;
      const int n = pair / p.heads, h = pair - n * p.heads;
      const uint32_t base = smem_u32(smem + st * kTcStage);
      float inv_l[2];
      auto store_o = [&](int j) {
        // O_j row -> * 1/l -> 64 x 16-bit = 128 B of the output row
        op_t* dst = p.out + (static_cast<long>(n) * kTcT + j * 128 + r) * p.o_row_stride + h * kTcD;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(lane_taddr + j * 256, v0);
        tmem_ld_32x32(lane_taddr + j * 256 + 32, v1);
        tmem_ld_wait();
        const float s = inv_l[j];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t(&v)[32] = half == 0 ? v0 : v1;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 o = make_uint4(pack_op2(__uint_as_float(v[8 * c]) * s, __uint_as_float(v[8 * c + 1]) * s),
                                       pack_op2(__uint_as_float(v[8 * c + 2]) * s, __uint_as_float(v[8 * c + 3]) * s),
                                       pack_op2(__uint_as_float(v[8 * c + 4]) * s, __uint_as_float(v[8 * c + 5]) * s),
                                       pack_op2(__uint_as_float(v[8 * c + 6]) * s, __uint_as_float(v[8 * c + 7]) * s));
            *reinterpret_cast<uint4*>(dst + 32 * half + 8 * c) = o;
          }
        }
      };
      mbar_wait(s_ready, ph);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint32_t taddr = lane_taddr + j * 256;
        // pass 1: row maximum over the 256 keys
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + 32 * c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
        }
        const float mb = m * p.scale_log2;
        if (j == 1) {
          // the P buffers (Q / K tiles) are still read by the O_0 MMAs; O_0 is also ready to be stored then
          mbar_wait(&o_ready[0], ph);
          tc_fence_after();
          store_o(0);
        }
        // pass 2: exponentials (log2 domain), row sum, P as 16-bit K-major operand.  (Measured slower: double-
        // buffering the TMEM loads in registers, and 64-column loads — both end at 255 registers with spills.)
        float l = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + 32 * c, v);
          tmem_ld_wait();
          float e[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            e[i] = ex2f(__uint_as_float(v[i]) * p.scale_log2 - mb);
            l += e[i];
          }
          const uint32_t blk = base + (c >> 1) * 16384 + r * 128;  // K block of 64 keys
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t ci = (c & 1) * 4 + q;  // 16-byte chunk (8 keys) inside the 128-byte row
            sts128u_(blk + ((ci ^ x7) << 4), make_uint4(pack_op2(e[8 * q], e[8 * q + 1]), pack_op2(e[8 * q + 2], e[8 * q + 3]),
                                                        pack_op2(e[8 * q + 4], e[8 * q + 5]), pack_op2(e[8 * q + 6], e[8 * q + 7])));
          }
        }
        inv_l[j] = 1.0f / l;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[j]);
      }
      mbar_wait(&o_ready[1], ph);
      tc_fence_after();
      store_o(1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tfree);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
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
    if (cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem) != cudaSuccess) return 1;
    attr_set = true;
  }
  const int grid = p.pairs < kNumSMs ? p.pairs : kNumSMs;
  attn_tc_kernel<<<grid, kTcThreads, kTcSmem, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace sgdm
