// tcgen05 + TMA implicit-GEMM convolution (see conv.cuh).
//
// CTA = 192 threads, persistent over (m_tile, n_tile) work items, 1 CTA / SM:
//   warp 0 (one lane)  TMA producer : 4-stage ring of {A 128x64, B block_n x 64} tiles
//   warp 1 (one lane)  MMA issuer   : 4 x tcgen05.mma (M128, N=block_n, K16) per stage,
//                                     accumulating in one of two TMEM accumulator stages
//   warps 2..5         epilogue     : tcgen05.ld -> +bias (+residual) -> global store,
//                                     overlapping the next tile's MMAs
// Roofline: tensor pipe.  FLOPs per launch = 2 * M_total * Cout * Ktot.
#include <cudaTypedefs.h>
#include <stdio.h>
#include <string.h>

#include "conv.cuh"

namespace sgdm {

constexpr int kStages = 4;
constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 64 * 2;   // 16 KB
constexpr int kBBytesMax = 256 * 64 * 2;   // 32 KB
constexpr int kBarBytes = 256;
constexpr int kStgLd = 36;                          // staging row stride in floats (32 + 4: conflict-free)
constexpr int kStgBytes = 4 * 32 * kStgLd * 4;      // one 32x32 fp32 sub-tile per epilogue warp
constexpr int kConvSmem = 1024 + kStages * (kABytes + kBBytesMax) + kBarBytes + kStgBytes;
constexpr int kConvThreads = 192;

// kCtas == 2: CTA-pair mode.  The two CTAs of a cluster own adjacent 128-row m-tiles of one 256-row MMA
// (tcgen05 cta_group::2): each loads its own A tile and HALF of the B (weight) tile, the leader CTA
// issues the MMAs for both and the accumulator rows land in each CTA's own TMEM.  Per CTA this halves
// the weight bytes pulled through L2 -> SM and the shared-memory reads per MMA.
template <int kCtas>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kABytes + kBBytesMax));
  uint64_t* full = bars;             // [kStages] TMA -> MMA
  uint64_t* empty = bars + kStages;  // [kStages] MMA -> TMA
  uint64_t* tfull = bars + 2 * kStages;       // [2] MMA -> epilogue
  uint64_t* tempty = bars + 2 * kStages + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  float* stg_all = reinterpret_cast<float*>(smem + kStages * (kABytes + kBBytesMax) + kBarBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int block_n = p.block_n;
  // two accumulator stages of block_n fp32 columns each; allocation must be a power of 2 >= 32
  const uint32_t acc_cols = p.swap_ab ? 256u : static_cast<uint32_t>(block_n);  // fp32 columns per accumulator
  uint32_t ncols = 32;
  while (ncols < 2u * acc_cols) ncols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if (p.kc2) tma_prefetch_desc(&p.tmA2);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4 * kCtas);  // one arrive per epilogue warp (of both CTAs in pair mode)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kCtas == 2) { tmem_alloc_pair(tmem_slot, ncols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, ncols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all();  // the peer's barriers must be initialised before any remote arrive
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int KB = p.taps * p.kc1 + p.kc2;
  const int cta_rank = kCtas == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  // work items: (m_tile, n_tile), or (pair of adjacent m_tiles, n_tile) per 2-CTA cluster
  const int total_tiles = (kCtas == 2 ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;
  const int tile_begin = blockIdx.x / kCtas, tile_step = gridDim.x / kCtas;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      uint32_t stage = 0, phase = 0;
      // swap_ab: the activation tile (256 pixels, 32 KB) lives in the big slot and is the MMA B operand,
      // the weight tile (128 x 64, 16 KB) in the small slot is the A operand.
      const uint32_t tx_bytes = p.swap_ab ? (kABytes + kBBytesMax) : (kCtas * kABytes + block_n * 128);  // of the whole pair in pair mode
      const int b_rows = block_n / kCtas;  // weight rows this CTA loads
      for (int tile = tile_begin; tile < total_tiles; tile += tile_step) {
        const int m_tile = (tile / p.n_tiles) * kCtas + cta_rank;
        const int n_tile = tile % p.n_tiles;
        const int p0 = m_tile * p.tile_px;
        const int img = p0 / p.HW;
        const int y0 = (p0 - img * p.HW) / p.Wout;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (cta_rank == 0) mbar_arrive_expect_tx(&full[stage], tx_bytes);
          uint8_t* act_dst = p.swap_ab ? sB + stage * kBBytesMax : sA + stage * kABytes;
          uint8_t* wgt_dst = p.swap_ab ? sA + stage * kABytes : sB + stage * kBBytesMax;
          if (kb < p.taps * p.kc1) {
            const int tap = kb / p.kc1;
            const int cc = kb - tap * p.kc1;
            const int r = tap / p.ks;
            const int s = tap - r * p.ks;
            if (kCtas == 2) tma_load_4d_pair(&p.tmA, &full[stage], act_dst, cc * 64, s - p.pad, y0 * p.stride + r - p.pad, img);
            else tma_load_4d(&p.tmA, &full[stage], act_dst, cc * 64, s - p.pad, y0 * p.stride + r - p.pad, img);
          } else {
            const int cc = kb - p.taps * p.kc1;
            if (kCtas == 2) tma_load_4d_pair(&p.tmA2, &full[stage], act_dst, cc * 64, 0, y0, img);
            else tma_load_4d(&p.tmA2, &full[stage], act_dst, cc * 64, 0, y0, img);
          }
          if (kCtas == 2) tma_load_2d_pair(&p.tmB, &full[stage], wgt_dst, kb * 64, n_tile * block_n + cta_rank * b_rows);
          else tma_load_2d(&p.tmB, &full[stage], wgt_dst, kb * 64, n_tile * block_n);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank == 0) {
      // ------------------------------------------------------------ MMA issuer (the leader CTA in pair mode)
      const uint32_t idesc = umma_idesc(kTileM * kCtas, p.swap_ab ? 256 : block_n);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = tile_begin; tile < total_tiles; tile += tile_step, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_cols;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * kABytes);
          const uint32_t b_addr = smem_u32(sB + stage * kBBytesMax);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (kCtas == 2)
              umma_f16_pair(d_tmem, umma_smem_desc(a_addr + k * 32), umma_smem_desc(b_addr + k * 32), idesc,
                            (kb | k) != 0 ? 1u : 0u);
            else
              umma_f16(d_tmem, umma_smem_desc(a_addr + k * 32), umma_smem_desc(b_addr + k * 32), idesc,
                       (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs) when these MMAs retire; accumulator complete -> epilogue(s)
          if (kCtas == 2) {
            umma_commit_pair(&empty[stage]);
            if (kb == KB - 1) umma_commit_pair(&tfull[acc]);
          } else {
            umma_commit(&empty[stage]);
            if (kb == KB - 1) umma_commit(&tfull[acc]);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5)
    // Each warp owns one TMEM lane quarter and walks its part of the accumulator in 32x32 fp32
    // chunks.  A chunk is read from TMEM, put into a private smem staging tile as [pixel][channel]
    // (normal mode: thread = pixel row, 8 x 16-byte stores; swap-AB: thread = channel, 32 scalar
    // stores), and written back with thread = (row group, column quad) so that every global load /
    // store instruction of the warp touches whole 128-byte row segments (bias, residual and output
    // alike).  The same pass optionally emits GroupNorm partial statistics of the FINAL values:
    // per (32-row block, `stat_gran` channels) sum and sum of squares (see conv.cuh).
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    float* stg = stg_all + (warp - 2) * 32 * kStgLd;
    const int rg = lane >> 3, cq4 = (lane & 7) * 4;  // fp32 path: 8 lanes per row (float4 each), 4 rows per instruction
    const int rl8 = lane >> 2, cq8 = (lane & 3) * 8;  // 16-bit path: 4 lanes per row (8 columns each), 8 rows per instruction
    const bool use_res = p.res_mode != 0 && p.out_f32 != nullptr;
    const int n_chunks = p.swap_ab ? p.tile_px / 32 : (block_n + 31) / 32;
    const int sg_shift = p.stat_gran == 4 ? 2 : 1;
    const int stat_ld = p.N_total >> sg_shift;  // stat entries per 32-row block
    uint32_t it = 0;
    for (int tile = tile_begin; tile < total_tiles; tile += tile_step, ++it) {
      const int m_tile = (tile / p.n_tiles) * kCtas + cta_rank;
      const int n_tile = tile % p.n_tiles;
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      // chunk i covers rows [row0(i), +32) x columns [col0(i), +nc(i)) of the output matrix
      const int tile_row0 = p.swap_ab ? m_tile * p.tile_px : m_tile * kTileM + quarter * 32;
      const int tile_col0 = p.swap_ab ? quarter * 32 : n_tile * block_n;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * acc_cols;
      // Residual rows are known before the accumulator is ready: their loads are issued one chunk
      // ahead (the first before waiting on the MMA), so that ~32 KB of residual reads per SM are in
      // flight instead of one dependent 16-byte load per lane.
      int rrow[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int m = tile_row0 + 4 * k + rg;
        int r = -1;
        if (use_res && m < p.M_total) {
          r = m;
          if (p.res_mode == 2) {
            const int img = m / p.HW, pix = m - img * p.HW;
            const int y = pix / p.Wout, x = pix - y * p.Wout;
            r = (img * (p.Hout >> 1) + (y >> 1)) * (p.Wout >> 1) + (x >> 1);
          }
        }
        rrow[k] = r;
      }
      float4 rcur[8], rnext[8];
      auto load_res = [&](int i, float4 (&r)[8]) {
        const int nc = p.swap_ab ? 32 : min(32, block_n - 32 * i);
        const int col = tile_col0 + (p.swap_ab ? 0 : 32 * i) + cq4;
        const bool ok = cq4 < nc && col < p.N_total;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          int rr = rrow[k];
          if (p.swap_ab) {  // rows advance with the chunk (res_mode 1 only)
            rr = tile_row0 + 32 * i + 4 * k + rg;
            if (rr >= p.M_total) rr = -1;
          }
          r[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok && rr >= 0) r[k] = __ldg(reinterpret_cast<const float4*>(p.res + static_cast<long>(rr) * p.N_total + col));
        }
      };
      if (use_res) load_res(0, rcur);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      for (int i = 0; i < n_chunks; ++i) {
        uint32_t v[32];
        const int nc = p.swap_ab ? 32 : min(32, block_n - 32 * i);
        const int row0 = p.swap_ab ? tile_row0 + 32 * i : tile_row0;
        const int col0 = p.swap_ab ? tile_col0 : tile_col0 + 32 * i;
        if (nc == 32) tmem_ld_32x32(taddr + 32 * i, v);
        else tmem_ld_32x16(taddr + 32 * i, v);
        if (use_res && i + 1 < n_chunks) load_res(i + 1, rnext);
        tmem_ld_wait();
        if (p.out_nchw != nullptr) {  // final conv: lanes = adjacent pixels -> already coalesced
          const int m = row0 + lane;
          if (m < p.M_total) {
            const int img = m / p.HW, pix = m - img * p.HW;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              if (j < nc && col < p.N_total) {
                float a = __uint_as_float(v[j]);
                if (p.bias) a += p.bias[col];
                p.out_nchw[(static_cast<long>(img) * p.N_total + col) * p.HW + pix] = a;
              }
            }
          }
          continue;
        }
        if (!p.swap_ab) {
          float4* srow = reinterpret_cast<float4*>(stg + lane * kStgLd);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (4 * j < nc)
              srow[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                    __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[j * kStgLd + lane] = __uint_as_float(v[j]);
        }
        __syncwarp();
        // pair sums for the GroupNorm statistics: sg/qg[u] covers columns {2u, 2u+1} of this lane's slice
        float sg[4] = {0.f, 0.f, 0.f, 0.f}, qg[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.out_f32) {
          const int col = col0 + cq4;
          const bool col_ok = cq4 < nc && col < p.N_total;
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col_ok && p.bias) b = *reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rl = 4 * k + rg;
            const int m = row0 + rl;
            if (!col_ok || m >= p.M_total) continue;
            float4 a = *reinterpret_cast<const float4*>(stg + rl * kStgLd + cq4);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            if (use_res) { a.x += rcur[k].x; a.y += rcur[k].y; a.z += rcur[k].z; a.w += rcur[k].w; }
            *reinterpret_cast<float4*>(p.out_f32 + static_cast<long>(m) * p.N_total + col) = a;
            sg[0] += a.x + a.y; qg[0] += a.x * a.x + a.y * a.y;
            sg[1] += a.z + a.w; qg[1] += a.z * a.z + a.w * a.w;
          }
          if (use_res) {
#pragma unroll
            for (int k = 0; k < 8; ++k) rcur[k] = rnext[k];
          }
          if (p.stats) {
            if (p.stat_gran == 4) { sg[0] += sg[1]; qg[0] += qg[1]; }
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                sg[u] += __shfl_xor_sync(0xffffffffu, sg[u], o);
                qg[u] += __shfl_xor_sync(0xffffffffu, qg[u], o);
              }
            }
            if (rg == 0 && col_ok && row0 < p.M_total) {
              float2* dst = p.stats + static_cast<long>(row0 >> 5) * stat_ld + (col >> sg_shift);
              if (p.stat_gran == 4) dst[0] = make_float2(sg[0], qg[0]);
              else *reinterpret_cast<float4*>(dst) = make_float4(sg[0], qg[0], sg[1], qg[1]);
            }
          }
        } else {
          const int col = col0 + cq8;
          const bool col_ok = cq8 < nc && col < p.N_total;
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
          if (col_ok && p.bias) {
            b0 = *reinterpret_cast<const float4*>(p.bias + col);
            b1 = *reinterpret_cast<const float4*>(p.bias + col + 4);
          }
#pragma unroll
          for (int rr = 0; rr < 32; rr += 8) {
            const int rl = rr + rl8;
            const int m = row0 + rl;
            if (!col_ok || m >= p.M_total) continue;
            float4 a0 = *reinterpret_cast<const float4*>(stg + rl * kStgLd + cq8);
            float4 a1 = *reinterpret_cast<const float4*>(stg + rl * kStgLd + cq8 + 4);
            a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
            a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
            if (p.res_mode == 1) {
              const float* rp = p.res + static_cast<long>(m) * p.N_total + col;
              const float4 r0 = __ldg(reinterpret_cast<const float4*>(rp)), r1 = __ldg(reinterpret_cast<const float4*>(rp + 4));
              a0.x += r0.x; a0.y += r0.y; a0.z += r0.z; a0.w += r0.w;
              a1.x += r1.x; a1.y += r1.y; a1.z += r1.z; a1.w += r1.w;
            }
            const uint4 h = make_uint4(pack_op2(a0.x, a0.y), pack_op2(a0.z, a0.w), pack_op2(a1.x, a1.y), pack_op2(a1.z, a1.w));
            *reinterpret_cast<uint4*>(p.out_op + static_cast<long>(m) * p.N_total + col) = h;
            sg[0] += a0.x + a0.y; qg[0] += a0.x * a0.x + a0.y * a0.y;
            sg[1] += a0.z + a0.w; qg[1] += a0.z * a0.z + a0.w * a0.w;
            sg[2] += a1.x + a1.y; qg[2] += a1.x * a1.x + a1.y * a1.y;
            sg[3] += a1.z + a1.w; qg[3] += a1.z * a1.z + a1.w * a1.w;
          }
          if (p.stats) {
            if (p.stat_gran == 4) {
              sg[0] += sg[1]; qg[0] += qg[1];
              sg[1] = sg[2] + sg[3]; qg[1] = qg[2] + qg[3];
            }
#pragma unroll
            for (int o = 4; o <= 16; o <<= 1) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (u < 2 || p.stat_gran != 4) {
                  sg[u] += __shfl_xor_sync(0xffffffffu, sg[u], o);
                  qg[u] += __shfl_xor_sync(0xffffffffu, qg[u], o);
                }
              }
            }
            if (rl8 == 0 && col_ok && row0 < p.M_total) {
              float4* dst = reinterpret_cast<float4*>(p.stats + static_cast<long>(row0 >> 5) * stat_ld + (col >> sg_shift));
              dst[0] = make_float4(sg[0], qg[0], sg[1], qg[1]);
              if (p.stat_gran != 4) dst[1] = make_float4(sg[2], qg[2], sg[3], qg[3]);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCtas == 2) mbar_arrive_leader(&tempty[acc]);
        else mbar_arrive(&tempty[acc]);
      }
    }
  }
  tc_fence_before();
  __syncwarp();
  if (kCtas == 2) {
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while the peer can still touch it
    if (warp == 1) tmem_dealloc_pair(tmem_base, ncols);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols);
  }
}

// ------------------------------------------------------------------------------------ host
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
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

#ifdef SGDM_OPERAND_BF16
#define SGDM_TMA_DTYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#else
#define SGDM_TMA_DTYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#endif

static int encode_nhwc(CUtensorMap* tm, const op_t* base, int B, int H, int W, int C, int bw, int bh, int bn,
                       int stride, char* err, int errlen) {
  auto fn = get_encode_fn();
  if (!fn) { snprintf(err, errlen, "cuTensorMapEncodeTiled entry point unavailable"); return 1; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = fn(tm, SGDM_TMA_DTYPE, 4, const_cast<op_t*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled(A: B%d H%d W%d C%d box %dx%dx%d s%d) failed: %d", B, H, W, C, bw,
             bh, bn, stride, (int)r);
    return 1;
  }
  return 0;
}

int conv_prepare(const ConvDesc& d, ConvLaunch* out, char* err, int errlen) {
  memset(&out->p, 0, sizeof(out->p));
  out->desc = d;
  ConvKernelParams& p = out->p;
  auto fail = [&](const char* msg) { snprintf(err, errlen, "conv_prepare: %s", msg); return 1; };
  if (d.Cin % 64 || (d.in2 && d.C2 % 64)) return fail("channel counts must be multiples of 64");
  if (!(d.ks == 1 || d.ks == 3) || !(d.stride == 1 || d.stride == 2)) return fail("unsupported ks/stride");
  if (d.block_n != 16 && (d.block_n % 32 || d.block_n > 256 || d.block_n <= 0)) return fail("bad block_n");
  const int HW = d.Hout * d.Wout;
  if (!d.swap_ab) {
    if (d.Wout > 128 || (128 % d.Wout) != 0) return fail("Wout must divide 128");
    if (!((HW % 128) == 0 || (128 % HW) == 0)) return fail("Hout*Wout must divide or be a multiple of 128");
  } else if (d.block_n != 128) {
    return fail("swap_ab uses block_n == 128 (all output channels in one tile)");
  }
  if (d.out_nchw == nullptr && (d.Cout % 8)) return fail("Cout % 8 != 0 needs the NCHW epilogue");
  if ((d.out_f32 != nullptr) + (d.out_op != nullptr) + (d.out_nchw != nullptr) != 1) return fail("exactly one output");
  if (d.out_op && d.res && d.res_mode == 2) return fail("res_mode 2 needs the fp32 output");
  if (d.res_mode == 2 && ((d.Hout | d.Wout) & 1)) return fail("res_mode 2 needs even output size");
  if (d.swap_ab && !conv_can_swap(d)) return fail("swap_ab needs Cout == 128, no NCHW / upsampled-residual epilogue");
  if (d.stats && (d.out_nchw || (d.Cout % 8) || (d.stat_gran != 2 && d.stat_gran != 4)))
    return fail("stats need an NHWC output, Cout % 8 == 0 and stat_gran in {2, 4}");
  const int tile_px = d.swap_ab ? 256 : kTileM;
  const bool pair = conv_use_pair(d);
  const int bw = d.Wout;
  const int bh = min(d.Hout, tile_px / bw);
  const int bn = tile_px / (bw * bh);
  if (encode_nhwc(&p.tmA, d.in, d.B, d.Hin, d.Win, d.Cin, bw, bh, bn, d.stride, err, errlen)) return 1;
  if (d.in2) {
    if (encode_nhwc(&p.tmA2, d.in2, d.B, d.Hout, d.Wout, d.C2, bw, bh, bn, 1, err, errlen)) return 1;
  }
  const int Ktot = d.ks * d.ks * d.Cin + (d.in2 ? d.C2 : 0);
  const int npad = conv_npad(d.Cout, d.block_n);
  {
    auto fn = get_encode_fn();
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)npad};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(d.block_n / (pair ? 2 : 1))};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&p.tmB, SGDM_TMA_DTYPE, 2, const_cast<op_t*>(d.w), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      snprintf(err, errlen, "cuTensorMapEncodeTiled(B: K%d N%d) failed: %d", Ktot, npad, (int)r);
      return 1;
    }
  }
  p.M_total = d.B * HW;
  p.HW = HW;
  p.Wout = d.Wout;
  p.Hout = d.Hout;
  p.stride = d.stride;
  p.pad = d.pad;
  p.ks = d.ks;
  p.taps = d.ks * d.ks;
  p.kc1 = d.Cin / 64;
  p.kc2 = d.in2 ? d.C2 / 64 : 0;
  p.N_total = d.Cout;
  p.block_n = d.block_n;
  p.n_tiles = npad / d.block_n;
  p.swap_ab = d.swap_ab;
  p.tile_px = tile_px;
  p.m_tiles = (p.M_total + tile_px - 1) / tile_px;
  p.bias = d.bias;
  p.res = d.res;
  p.res_mode = d.res ? d.res_mode : 0;
  p.out_f32 = d.out_f32;
  p.out_op = d.out_op;
  p.out_nchw = d.out_nchw;
  p.stats = d.stats;
  p.stat_gran = d.stat_gran;
  out->pair = pair ? 1 : 0;
  if (pair) {
    const int total = (p.m_tiles + 1) / 2 * p.n_tiles;
    out->grid = 2 * (total < kNumSMs / 2 ? total : kNumSMs / 2);
  } else {
    const int total = p.m_tiles * p.n_tiles;
    out->grid = total < kNumSMs ? total : kNumSMs;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmem);
    if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    attr_set = true;
  }
  return 0;
}

int conv_launch(const ConvLaunch& l, cudaStream_t stream) {
  if (l.pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(l.grid);
    cfg.blockDim = dim3(kConvThreads);
    cfg.dynamicSmemBytes = kConvSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<2>, l.p) == cudaSuccess ? 0 : 1;
  }
  conv_gemm_kernel<1><<<l.grid, kConvThreads, kConvSmem, stream>>>(l.p);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// --------------------------------------------------------------------- CUDA-core checker
__global__ void conv_naive_kernel(ConvDesc d, int npad) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const long M = static_cast<long>(d.B) * d.Hout * d.Wout;
  if (idx >= M * d.Cout) return;
  const int col = idx % d.Cout;
  const long m = idx / d.Cout;
  const int HW = d.Hout * d.Wout;
  const int img = m / HW, pix = m % HW, y = pix / d.Wout, x = pix % d.Wout;
  const int Ktot = d.ks * d.ks * d.Cin + (d.in2 ? d.C2 : 0);
  const op_t* wrow = d.w + static_cast<long>(col) * Ktot;
  float acc = 0.f;
  for (int r = 0; r < d.ks; ++r)
    for (int s = 0; s < d.ks; ++s) {
      const int iy = y * d.stride + r - d.pad, ix = x * d.stride + s - d.pad;
      if (iy < 0 || iy >= d.Hin || ix < 0 || ix >= d.Win) continue;
      const op_t* a = d.in + ((static_cast<long>(img) * d.Hin + iy) * d.Win + ix) * d.Cin;
      const op_t* w = wrow + (r * d.ks + s) * d.Cin;
      for (int c = 0; c < d.Cin; ++c) acc += from_op(a[c]) * from_op(w[c]);
    }
  if (d.in2) {
    const op_t* a = d.in2 + m * d.C2;
    const op_t* w = wrow + d.ks * d.ks * d.Cin;
    for (int c = 0; c < d.C2; ++c) acc += from_op(a[c]) * from_op(w[c]);
  }
  if (d.bias) acc += d.bias[col];
  if (d.res && d.res_mode == 1) acc += d.res[m * d.Cout + col];
  if (d.res && d.res_mode == 2)
    acc += d.res[((static_cast<long>(img) * (d.Hout / 2) + y / 2) * (d.Wout / 2) + x / 2) * d.Cout + col];
  if (d.out_f32) d.out_f32[m * d.Cout + col] = acc;
  if (d.out_op) d.out_op[m * d.Cout + col] = to_op(acc);
  if (d.out_nchw) d.out_nchw[(static_cast<long>(img) * d.Cout + col) * HW + pix] = acc;
}

int conv_launch_naive(const ConvDesc& d, cudaStream_t stream) {
  const long total = static_cast<long>(d.B) * d.Hout * d.Wout * d.Cout;
  const int threads = 256;
  conv_naive_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, stream>>>(
      d, conv_npad(d.Cout, d.block_n));
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace sgdm
