// Fused attention (flash-style online softmax, warp-shuffle row reductions).
//
// Replaces QKVAttentionLegacy (openaimodel.py:403-420: 8 heads, per-head [q|k|v] channel
// blocks, q and k each scaled by ch^-1/4) and the attention core of Attention_LR
// (crossattetion_lr.py:88-137: multi-query, ONE shared k/v head, keys = [16 context |
// 1 null | T self], q scaled by d^-1/2) — SURVEY.md §2.3 rows K6/K7.
//
// One CTA = 128 queries of one (sample, head); 8 warps x 16 query rows.  K and V of the
// sample are staged in shared memory with 16-byte cp.async — all T_kv rows at once when they fit (T_kv <= 320 incl. the
// extra context/null rows at d = 64), otherwise in blocks of 256 keys that the online softmax walks one after the other
// (T = 1024: 128x128 images with attention at ds 4); S = QK^T and O = PV run on mma.sync m16n8k16 (16-bit operands, fp32 accumulate);
// the softmax is computed in fp32 over key chunks of 64 with running max / sum.
// Attention is ~1 % of the UNet FLOPs.
#include "attn.cuh"

namespace sgdm {

#ifdef SGDM_OPERAND_BF16
#define SGDM_MMA "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32"
#else
#define SGDM_MMA "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32"
#endif

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(SGDM_MMA " {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

constexpr int kAttnQ = 128;       // queries per CTA
constexpr int kAttnThreads = 256;  // 8 warps x 16 query rows

// D: head dim as the MMAs see it (multiple of 16); DR <= D: the real head dim (8 for 32 heads on 256 channels — the rows
// are zero-extended to 16 in shared memory)
template <int D, int DR = D>
__global__ void __launch_bounds__(kAttnThreads) attn_kernel(const AttnDesc a, int tkv_pad, int kv_blk) {
  constexpr int LD = D + 8;  // padded row: conflict-free ldmatrix
  extern __shared__ __align__(16) uint8_t smem_attn[];
  pdl_launch_dependents();
  pdl_wait();
  op_t* sK = reinterpret_cast<op_t*>(smem_attn);   // kv_blk rows (a multiple of 64; == tkv_pad when everything fits)
  op_t* sV = sK + static_cast<long>(kv_blk) * LD;
  op_t* sQ = sV + static_cast<long>(kv_blk) * LD;
  const int n = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kAttnQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tkv = a.n_extra + a.T;
  constexpr int CPR = D / 8;   // 16-byte chunks per row
  constexpr int CPRR = DR / 8;  // ... that exist in global memory

  // ---- stage K, V (extra rows first, then the T self rows) and the Q tile: 16-byte cp.async,
  //      zero-fill for the padding rows
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  auto stage_kv = [&](int kb0) {  // rows [kb0, kb0 + kv_blk) of the key / value sequence -> sK / sV
  for (int i = threadIdx.x; i < kv_blk * CPR; i += kAttnThreads) {
    const int lrow = i / CPR, ch = i - lrow * CPR, row = kb0 + lrow;
    op_t* dk = sK + lrow * LD + ch * 8;
    op_t* dv = sV + lrow * LD + ch * 8;
    if (ch >= CPRR) {
      *reinterpret_cast<uint4*>(dk) = zero4;
      *reinterpret_cast<uint4*>(dv) = zero4;
    } else if (row < a.n_extra) {
      const long o = (static_cast<long>(n) * a.n_extra + row) * DR + ch * 8;
      cp_async16(dk, a.k_extra + o);
      cp_async16(dv, a.v_extra + o);
    } else if (row < Tkv) {
      const long tok = static_cast<long>(n) * a.T + (row - a.n_extra);
      cp_async16(dk, a.k + tok * a.k_row_stride + h * a.k_head_stride + ch * 8);
      cp_async16(dv, a.v + tok * a.v_row_stride + h * a.v_head_stride + ch * 8);
    } else {
      *reinterpret_cast<uint4*>(dk) = zero4;
      *reinterpret_cast<uint4*>(dv) = zero4;
    }
  }
  };
  stage_kv(0);
  for (int i = threadIdx.x; i < kAttnQ * CPR; i += kAttnThreads) {
    const int row = i / CPR, ch = i - row * CPR;
    op_t* dq = sQ + row * LD + ch * 8;
    if (q0 + row < a.T && ch < CPRR)
      cp_async16(dq, a.q + (static_cast<long>(n) * a.T + q0 + row) * a.q_row_stride + h * a.q_head_stride + ch * 8);
    else
      *reinterpret_cast<uint4*>(dq) = zero4;
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // (a warp whose 16 query rows are all padding computes nothing but keeps taking part in the block-wide staging)
  const bool active = q0 + warp * 16 < a.T;

  // ---- Q fragments (A operand, 16 x D per warp)
  uint32_t qf[D / 16][4];
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks)
    ldsm_x4(qf[ks], sQ + (warp * 16 + (lane & 15)) * LD + ks * 16 + (lane >> 4) * 8);

  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float sl = a.scale * 1.4426950408889634f;  // logits in log2 units
  const int mi = lane >> 3, lr = lane & 7;

  const int nchunks = (Tkv + 63) / 64, cpb = kv_blk / 64;  // 64-key chunks in all / per staged block
  for (int kc = 0; kc < nchunks; ++kc) {
    if (kc > 0 && kc % cpb == 0) {  // next block of keys: everyone is done with the staged one
      __syncthreads();
      stage_kv(kc * 64);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    }
    if (!active) continue;
    const int kl = (kc % cpb) * 64;  // first row of this chunk inside the staged block
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        ldsm_x4(b, sK + (kl + np * 16 + (mi >> 1) * 8 + lr) * LD + ks * 16 + (mi & 1) * 8);
        mma16816(s[2 * np], qf[ks], b[0], b[1]);
        mma16816(s[2 * np + 1], qf[ks], b[2], b[3]);
      }
    }
    // scale, mask, running max
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = kc * 64 + nt * 8 + (lane & 3) * 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = key + (j & 1) < Tkv;
        s[nt][j] = ok ? s[nt][j] * sl : -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = fast_exp2(m0 - mx0), c1 = fast_exp2(m1 - mx1);
    m0 = mx0;
    m1 = mx1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];  // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = fast_exp2(s[nt][0] - m0), p1 = fast_exp2(s[nt][1] - m0);
      const float p2 = fast_exp2(s[nt][2] - m1), p3 = fast_exp2(s[nt][3] - m1);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_op2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_op2(p2, p3);
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int ndp = 0; ndp < D / 16; ++ndp) {
        uint32_t b[4];
        ldsm_x4_t(b, sV + (kl + kk * 16 + (mi & 1) * 8 + lr) * LD + (ndp * 2 + (mi >> 1)) * 8);
        mma16816(o[2 * ndp], pf[kk], b[0], b[1]);
        mma16816(o[2 * ndp + 1], pf[kk], b[2], b[3]);
      }
    }
  }
  if (!active) return;
  // ---- finalise: row sums across the 4 lanes of a row, normalise, store
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
  for (int nd = 0; nd < D / 8; ++nd) {
    const int d = nd * 8 + (lane & 3) * 2;
    if (d >= DR) continue;
    if (r0 < a.T)
      *reinterpret_cast<uint32_t*>(a.out + (static_cast<long>(n) * a.T + r0) * a.o_row_stride + h * DR + d) =
          pack_op2(o[nd][0] * i0, o[nd][1] * i0);
    if (r1 < a.T)
      *reinterpret_cast<uint32_t*>(a.out + (static_cast<long>(n) * a.T + r1) * a.o_row_stride + h * DR + d) =
          pack_op2(o[nd][2] * i1, o[nd][3] * i1);
  }
}

int attn_launch(const AttnDesc& a, cudaStream_t s) {
  if (a.use_tc != 0 && attn_tc_applicable(a)) return attn_tc_launch(a, s);
  if (a.use_tc != 0 && attn_lr_tc_applicable(a)) return attn_lr_tc_launch(a, s);
  const int Tkv = a.n_extra + a.T;
  const int tkv_pad = (Tkv + 63) / 64 * 64;
  // head dims of the reference configs (mc 64 / 128 / 256 with 8 heads: 32, 64, 128) and of num_heads = 32 (8, 16)
  if (a.D != 8 && a.D != 16 && a.D != 32 && a.D != 64 && a.D != 128) return 1;
  const int LD = (a.D < 16 ? 16 : a.D) + 8;
  // everything staged at once when it fits in 200 KB, else blocks of 256 keys
  int kv_blk = tkv_pad;
  if ((static_cast<size_t>(tkv_pad) * 2 + kAttnQ) * LD * sizeof(op_t) > 200 * 1024) kv_blk = 256;
  const size_t smem = (static_cast<size_t>(kv_blk) * 2 + kAttnQ) * LD * sizeof(op_t);
  if (smem > 200 * 1024) return 1;
  const dim3 grid((a.T + kAttnQ - 1) / kAttnQ, a.heads, a.B);
  static size_t max_set[5] = {0, 0, 0, 0, 0};  // opt-in dynamic smem limit, raised on demand
  auto run = [&](auto kernel, size_t& limit) {
    if (smem > limit) {
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
        return 1;
      limit = smem;
    }
    return launch_pdl(kernel, grid, dim3(kAttnThreads), smem, s, 1, a, tkv_pad, kv_blk) == cudaSuccess ? 0 : 1;
  };
  if (a.D == 64 ? run(attn_kernel<64>, max_set[0]) : a.D == 32 ? run(attn_kernel<32>, max_set[1]) :
      a.D == 128 ? run(attn_kernel<128>, max_set[2]) : a.D == 16 ? run(attn_kernel<16>, max_set[3]) : run(attn_kernel<16, 8>, max_set[4]))
    return 1;
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace sgdm
