// tcgen05 + TMA implicit-GEMM convolution (see conv.cuh).
//
// CTA = 192 threads, persistent over (m_tile, n_tile) work items, 1 CTA / SM:
//   warp 0 (one lane)  TMA producer : ring of {A 128x64, B rows x 64} K-block stages (as many as fit)
//   warp 1 (one lane)  MMA issuer   : 4 x tcgen05.mma (K16) per stage into one of two TMEM
//                                     accumulator stages
//   warps 2..5         epilogue     : tcgen05.ld -> +bias (+residual) -> swizzled smem tile -> TMA store,
//                                     overlapping the next tile's MMAs
// Three geometries share the code:
//   normal   D[128 px, block_n ch]            one CTA
//   pair     D[2 x 128 px, block_n ch]        tcgen05 cta_group::2 over a 2-CTA cluster (M = 256)
//   swap-AB  D^T[128 ch, 256 px] (Cout = 128) weights are the M operand, a 256-pixel tile the N operand
// Halo mode (3x3, stride 1, tiles of whole image rows): one stage holds the activation rows y0-1 .. y0+bh
// of ONE horizontal tap and 64 channels, plus the three weight tiles of the vertical taps; the three MMA
// groups read that one tile at row offsets 0 / W / 2W (whole 1024-byte swizzle atoms), so every activation
// row is pulled through L2 -> SM once per horizontal tap instead of once per tap.
// Roofline: tensor pipe.  FLOPs per launch = 2 * M_total * Cout * Ktot.
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "conv.cuh"

namespace sgdm {

constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
constexpr int kABytes = kTileM * 64 * 2;  // 16 KB: the M-side operand tile of one K block
constexpr int kConvThreads = 192;
constexpr int kSmemLimit = 232448;  // 227 KB opt-in maximum per CTA
constexpr int kBarBytes = 512;
constexpr int kBiasBytes = 2 * 256 * 4;  // bias of the tile's columns, double-buffered by tile parity
constexpr int kEpiBuf = 4096;            // one epilogue staging buffer: 32 rows x 128 B
constexpr int kMaxEpiBufs = 6;
constexpr int kEpi2Bytes = 4 * 2 * 2048;  // staging of the optional 16-bit copy: 32 rows x 64 B, double-buffered per warp

// ---- shared-window accessors (explicit state space: the 1024-byte alignment of the dynamic smem base
//      goes through an integer and the compiler would otherwise emit generic LD/ST) -------------------
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ void sts128u(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts16(uint32_t a, op_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(*reinterpret_cast<unsigned short*>(&v)));
}
// TMA tile store smem -> global (bulk async-group of the issuing thread) and tile load with mbarrier
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* d, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(d)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* d, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(d)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// pull one tile of a tensor map into L2 (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* d, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(d)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_a(const CUtensorMap* d, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(d)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// kCtas == 2: CTA-pair mode.  The two CTAs of a cluster own adjacent 128-row m-tiles of one 256-row MMA
// (tcgen05 cta_group::2): each loads its own A tile and HALF of the B (weight) tile, the leader CTA
// issues the MMAs for both and the accumulator rows land in each CTA's own TMEM.  Per CTA this halves
// the weight bytes pulled through L2 -> SM and the shared-memory reads per MMA.
// Optional cycle accounting of the epilogue warps (ConvDesc::timing != nullptr; bring-up / tuning only):
// slots 0 wait-for-accumulator | 1 wait-for-staging-buffer (TMA store drained) | 2 wait-for-residual |
// 3 TMEM load | 4 bias/residual/stage | 5 statistics | 6 fence + TMA store issue | 7 chunks | 8 MMA-issuer
// stalls on smem stages | 9 MMA-issuer stalls on the accumulator | 10 producer stalls on free stages
#define SGDM_T0() long long t_prev_ = p.timing ? clock64() : 0
#define SGDM_T(slot)                                   \
  do {                                                 \
    if (p.timing) {                                    \
      const long long t_now_ = clock64();              \
      t_acc_[slot] += t_now_ - t_prev_;                \
      t_prev_ = t_now_;                                \
    }                                                  \
  } while (0)

template <int kCtas>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvKernelParams p) {
  // No static shared memory in this kernel: the dynamic window starts 1024-byte aligned (checked), which the
  // 128-byte swizzle of the TMA / UMMA tiles needs.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  pdl_launch_dependents();
  // [ring: n_stages x (act | weights)] [epilogue staging: 4 warps x epi_bufs x 4 KB]
  // [16-bit copy staging: 4 warps x 2 x 2 KB, only with a second output] [bias: 2 x 1 KB] [barriers]
  const int stage_bytes = p.act_bytes + p.tps * p.wgt_bytes;  // [activation slot][tps weight slots]
  // A-stationary mode (1x1 GEMMs with several n-tiles): the kc1 activation K blocks of the current m-tile stay
  // resident in front of the ring, whose stages then hold weight tiles only (act_bytes = 0)
  uint8_t* a_res = smem;
  uint8_t* ring = smem + p.a_stat * p.kc1 * p.act_tx;
  uint8_t* after_ring = ring + p.n_stages * stage_bytes;
  const int epi2_bytes = p.out2 ? kEpi2Bytes : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(after_ring + 4 * p.epi_bufs * kEpiBuf + epi2_bytes + kBiasBytes);
  uint64_t* full = bars;                         // [kMaxStages] TMA -> MMA
  uint64_t* empty = bars + kMaxStages;           // [kMaxStages] MMA -> TMA
  uint64_t* tfull = bars + 2 * kMaxStages;       // [2] MMA -> epilogue
  uint64_t* tempty = bars + 2 * kMaxStages + 2;  // [2] epilogue -> MMA
  uint64_t* rbars = bars + 2 * kMaxStages + 4;   // [4 warps][kMaxEpiBufs] residual tile landed
  uint64_t* bfree = bars + 2 * kMaxStages + 4 + 4 * kMaxEpiBufs;  // [2] all epilogue warps are done with a bias buffer
  uint64_t* a_full = bars + 2 * kMaxStages + 6 + 4 * kMaxEpiBufs;  // [kMaxStages] A-stationary mode: resident A K block landed
  uint64_t* a_empty = a_full + kMaxStages;                         // [kMaxStages] ... and may be replaced (last n-tile's MMA retired)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + kMaxStages);
  const uint32_t epi_all = smem_u32(after_ring);
  const uint32_t epi2_all = epi_all + 4 * p.epi_bufs * kEpiBuf;
  const uint32_t sbias_all = epi2_all + epi2_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int block_n = p.block_n;
  const int n_stages = p.n_stages;
  // two accumulator stages of acc_cols fp32 columns each; allocation must be a power of 2 >= 32
  const uint32_t acc_cols = p.swap_ab ? 256u : static_cast<uint32_t>(block_n);
  uint32_t ncols = 32;
  while (ncols < 2u * acc_cols) ncols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if (p.kc2) tma_prefetch_desc(&p.tmA2);
    if (p.kc2 > p.kc2a) tma_prefetch_desc(&p.tmA2b);
    if (p.epi_mode != 0) tma_prefetch_desc(&p.tmOut);
    if (p.out2) tma_prefetch_desc(&p.tmOut2);
    if (p.res_mode == 1) tma_prefetch_desc(&p.tmRes);
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4 * kCtas);  // one arrive per epilogue warp (of both CTAs in pair mode)
    }
    for (int i = 0; i < 4 * kMaxEpiBufs; ++i) mbar_init(&rbars[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&bfree[i], 4);
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
  pdl_wait();  // everything above is CTA-private set-up; the preceding kernel's outputs are read from here on

  const int cta_rank = kCtas == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  // work items: (m_tile, n_tile), or (pair of adjacent m_tiles, n_tile) per 2-CTA cluster
  const int total_tiles = (kCtas == 2 ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;
  const int tile_begin = blockIdx.x / kCtas, tile_step = gridDim.x / kCtas;
  // this CTA's (pair's) tile sequence: round-robin over all (m, n) tiles, or — A-stationary — round-robin over the
  // m-tiles with all n-tiles of an m-tile back to back
  const int tile_first = p.a_stat ? tile_begin * p.n_tiles : tile_begin;
  // A-stationary: every CTA walks the n-tiles of its m-tile in a different rotation, so that at any moment the CTAs
  // pull different weight tiles (all CTAs reading the same 16 KB of weights in lockstep serialises on its L2 slices)
  const int n_rot = p.a_stat ? tile_begin % p.n_tiles : 0;
  auto n_of = [&](int tile) {
    const int n = tile % p.n_tiles + n_rot;
    return n >= p.n_tiles ? n - p.n_tiles : n;
  };
  auto tile_next = [&](int tile) {
    if (!p.a_stat) return tile + tile_step;
    const int n = tile % p.n_tiles;
    return n + 1 < p.n_tiles ? tile + 1 : tile - n + tile_step * p.n_tiles;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      uint32_t stage = 0, phase = 0;
      long long t_prod = 0;
      const int b_rows = block_n / kCtas;  // weight rows this CTA loads
      // main K blocks: halo stages hold the three vertical taps of one (horizontal tap, chunk); plain stages hold
      // p.kps consecutive (tap, chunk) K blocks, each with its own activation tile and weight slot
      // (sub-pixel mode, halo geometry: two horizontal taps per parity; dense geometry: a plain 2x2 conv, p.ks = 2)
      const int n_blk = (p.hfold ? 1 : p.up2 == 1 ? 2 : p.halo ? p.ks : p.taps) * p.kc1;
      const int n_main = p.halo ? n_blk : (n_blk + p.kps - 1) / p.kps, n_st = n_main + (p.kc2 + p.tps2 - 1) / p.tps2;
      uint32_t a_it = 0;  // A-stationary: m-tiles loaded so far (parity of the resident slots)
      for (int tile = tile_first; tile < total_tiles; tile = tile_next(tile)) {
        const int m_tile = (tile / p.n_tiles) * kCtas + cta_rank;
        const int n_pos = tile % p.n_tiles;  // position in this CTA's walk over the m-tile (A-stationary)
        const int n_tile = n_of(tile);
        const int p0 = m_tile * p.tile_px;
        const int img = p0 / p.HW;
        const int y0 = (p0 - img * p.HW) / p.Wout;
        // sub-pixel mode: this n-tile's output parity (dy, dx) selects the vertical taps {dy, dy+1} and horizontal {dx, dx+1}
        const int par = p.up2 ? n_tile / p.ntpp : 0, up_dy = par >> 1, up_dx = par & 1;
        int q_tap = 0, q_cc = 0, q_r = 0;
        for (int q = 0; q < n_st; ++q) {
          const long long tw0 = p.timing ? clock64() : 0;
          mbar_wait(&empty[stage], phase ^ 1);
          if (p.timing) t_prod += clock64() - tw0;
          const bool main_st = q < n_main;
          if (p.a_stat) {
            // resident A K block q of this m-tile: loaded with the first n-tile, into the slot the previous m-tile's
            // last n-tile has just released; the ring stage carries the weight tile only
            if (n_pos == 0) {
              mbar_wait(&a_empty[q], (a_it & 1) ^ 1);
              if (cta_rank == 0) mbar_arrive_expect_tx(&a_full[q], kCtas * p.act_tx);
              if (kCtas == 2) tma_load_4d_pair(&p.tmA, &a_full[q], a_res + q * p.act_tx, q * p.kblk, 0, y0, img);
              else tma_load_4d(&p.tmA, &a_full[q], a_res + q * p.act_tx, q * p.kblk, 0, y0, img);
            }
            if (cta_rank == 0) mbar_arrive_expect_tx(&full[stage], kCtas * p.wgt_tx);
            uint8_t* wdst = ring + stage * stage_bytes;
            if (kCtas == 2) tma_load_2d_pair(&p.tmB, &full[stage], wdst, q * p.kblk, n_tile * block_n + cta_rank * b_rows);
            else tma_load_2d(&p.tmB, &full[stage], wdst, q * p.kblk, n_tile * block_n);
            if (++stage == static_cast<uint32_t>(n_stages)) { stage = 0; phase ^= 1; }
            continue;
          }
          // a stage of the fused 1x1-skip source holds up to tps2 plain (tile, weight) K blocks
          const int ntap = main_st ? (p.halo ? p.tps : min(p.kps, n_blk - q * p.kps)) : min(p.tps2, p.kc2 - (q - n_main) * p.tps2);
          const uint32_t act_tx = (main_st && p.halo) ? p.act_tx_halo : ntap * p.act_tx;
          if (cta_rank == 0)
            mbar_arrive_expect_tx(&full[stage], kCtas * (act_tx + ntap * p.wgt_tx));
          uint8_t* act_dst = ring + stage * stage_bytes;
          uint8_t* wgt_dst = act_dst + p.act_bytes;
          int cc, kb0;
          if (main_st) {
            // (tap, chunk) and (r, s) of the tap are walked with counters: no integer divisions per K block
            kb0 = 0;
            for (int u = 0; u < (p.halo ? 1 : ntap); ++u) {
              const int tap = q_tap;  // halo: the horizontal tap s; otherwise the tap index r*ks+s
              cc = q_cc;
              const int r = p.halo ? 0 : q_r;
              const int s_tap = p.hfold ? p.pad : p.halo ? tap + up_dx : tap - r * p.ks + up_dx;  // hfold: no horizontal shift
              const int r_ld = r + (p.halo ? 0 : up_dy);  // dense sub-pixel mode: the parity's vertical offset goes into the load
              if (++q_cc == p.kc1) {
                q_cc = 0;
                ++q_tap;
                if (q_tap - q_r * p.ks == p.ks) ++q_r;
              }
              if (u == 0) kb0 = p.hfold ? cc : p.halo ? s_tap * p.kc1 + cc : q * p.kps;
              uint8_t* dst = act_dst + u * p.act_tx;
              if (kCtas == 2) tma_load_4d_pair(&p.tmA, &full[stage], dst, cc * p.kblk, s_tap - p.pad, y0 * p.stride + r_ld - p.pad, img);
              else tma_load_4d(&p.tmA, &full[stage], dst, cc * p.kblk, s_tap - p.pad, y0 * p.stride + r_ld - p.pad, img);
            }
          } else {
            cc = (q - n_main) * p.tps2;
            kb0 = p.taps * p.kc1 + cc;
            for (int t = 0; t < ntap; ++t) {
              // the skip source may be a channel concat of two tensors: [kc2a chunks of tmA2 | rest of tmA2b]
              const CUtensorMap* tm2 = (cc + t) < p.kc2a ? &p.tmA2 : &p.tmA2b;
              const int c2 = (cc + t) < p.kc2a ? (cc + t) : (cc + t) - p.kc2a;
              if (kCtas == 2) tma_load_4d_pair(tm2, &full[stage], act_dst + t * p.act_tx, c2 * p.kblk, 0, y0, img);
              else tma_load_4d(tm2, &full[stage], act_dst + t * p.act_tx, c2 * p.kblk, 0, y0, img);
            }
          }
          for (int t = 0; t < ntap; ++t) {
            // halo: vertical tap t -> K block (t*ks + s)*kc1 + cc; skip source: consecutive K blocks
            // (hfold: the packed K axis is (vertical tap, channel) only)
            // (plain stages: kps consecutive K blocks)
            const int kb = kb0 + ((main_st && p.halo) ? (t + up_dy) * (p.hfold ? 1 : p.ks) * p.kc1 : t);
            if (kCtas == 2) tma_load_2d_pair(&p.tmB, &full[stage], wgt_dst + t * p.wgt_bytes, kb * p.kblk, n_tile * block_n + cta_rank * b_rows);
            else tma_load_2d(&p.tmB, &full[stage], wgt_dst + t * p.wgt_bytes, kb * p.kblk, n_tile * block_n);
          }
          if (++stage == static_cast<uint32_t>(n_stages)) { stage = 0; phase ^= 1; }
        }
        if (p.a_stat && n_pos == 0) ++a_it;
      }
      if (p.timing) atomicAdd(reinterpret_cast<unsigned long long*>(p.timing) + 10, static_cast<unsigned long long>(t_prod));
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank == 0) {
      // ------------------------------------------------------------ MMA issuer (the leader CTA in pair mode)
      const uint32_t idesc = umma_idesc(kTileM * kCtas, p.swap_ab ? 256 : block_n);
      const int nk16 = p.kblk >> 4;
      const uint64_t desc_hi = umma_smem_desc_hi(p.kblk == 32);
      uint32_t stage = 0, phase = 0, it = 0, a_it = 0;
      long long t_full = 0, t_acc = 0;
      const int n_blk = (p.hfold ? 1 : p.up2 == 1 ? 2 : p.halo ? p.ks : p.taps) * p.kc1;
      const int n_main = p.halo ? n_blk : (n_blk + p.kps - 1) / p.kps, n_st = n_main + (p.kc2 + p.tps2 - 1) / p.tps2;
      for (int tile = tile_first; tile < total_tiles; tile = tile_next(tile), ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        const int up_dy = p.up2 ? (n_of(tile) / p.ntpp) >> 1 : 0;  // sub-pixel mode: first vertical tap of this tile's parity
        const long long ta0 = p.timing ? clock64() : 0;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        if (p.timing) t_acc += clock64() - ta0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_cols;
        const int n_tile_a = p.a_stat ? tile % p.n_tiles : 0;  // A-stationary: position inside the m-tile
        for (int q = 0; q < n_st; ++q) {
          const long long tw0 = p.timing ? clock64() : 0;
          if (p.a_stat && n_tile_a == 0) mbar_wait(&a_full[q], a_it & 1);  // this m-tile's resident A K block
          mbar_wait(&full[stage], phase);
          if (p.timing) t_full += clock64() - tw0;
          tc_fence_after();  // (measured in round 1: free)
          const bool main_st = q < n_main;
          const int ntap = main_st ? (p.halo ? p.tps : min(p.kps, n_blk - q * p.kps)) : min(p.tps2, p.kc2 - (q - n_main) * p.tps2);
          const uint32_t act_addr = p.a_stat ? smem_u32(a_res + q * p.act_tx) : smem_u32(ring + stage * stage_bytes);
          const uint32_t wgt_addr = p.a_stat ? smem_u32(ring + stage * stage_bytes) : act_addr + p.act_bytes;
          for (int t = 0; t < ntap; ++t) {
            // halo: vertical tap t reads the staged rows starting t image rows further down
            const uint32_t act_t = act_addr + ((main_st && p.halo) ? (t + up_dy) * p.halo_row_bytes : t * p.act_tx);
            const uint32_t wgt_t = wgt_addr + t * p.wgt_bytes;
            const uint32_t a_addr = p.swap_ab ? wgt_t : act_t;  // M-side operand
            const uint32_t b_addr = p.swap_ab ? act_t : wgt_t;  // N-side operand
            // K block = 64 channels (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B): 4 or 2 K16 slices
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k < nk16) {
                const uint64_t da = desc_hi | umma_smem_desc_lo(a_addr + k * 32), db = desc_hi | umma_smem_desc_lo(b_addr + k * 32);
                if (kCtas == 2) umma_f16_pair(d_tmem, da, db, idesc, (q | t | k) != 0 ? 1u : 0u);
                else umma_f16(d_tmem, da, db, idesc, (q | t | k) != 0 ? 1u : 0u);
              }
            }
          }
          // frees the smem stage (in both CTAs) when these MMAs retire; accumulator complete -> epilogue(s)
          const bool a_done = p.a_stat && n_tile_a == p.n_tiles - 1;  // last use of the resident A K block q
          if (kCtas == 2) {
            umma_commit_pair(&empty[stage]);
            if (a_done) umma_commit_pair(&a_empty[q]);
            if (q == n_st - 1) umma_commit_pair(&tfull[acc]);
          } else {
            umma_commit(&empty[stage]);
            if (a_done) umma_commit(&a_empty[q]);
            if (q == n_st - 1) umma_commit(&tfull[acc]);
          }
          if (++stage == static_cast<uint32_t>(n_stages)) { stage = 0; phase ^= 1; }
        }
        if (p.a_stat && n_tile_a == p.n_tiles - 1) ++a_it;
      }
      if (p.timing) {
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing) + 8, static_cast<unsigned long long>(t_full));
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing) + 9, static_cast<unsigned long long>(t_acc));
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5)
    // Each warp owns one TMEM lane quarter and walks its part of the accumulator in chunks of 32 rows x
    // 128 output bytes (32 fp32 or 64 16-bit columns).  A chunk goes TMEM -> registers (thread = row, or
    // thread = channel in swap-AB mode) -> +bias (+residual) -> a 128-byte-swizzled smem tile -> one TMA
    // tile store.  The residual tile is TMA-LOADED into that same smem tile two chunks ahead and added in
    // place, so the loop contains no global load, no bounds check (TMA clips / zero-fills) and no address
    // arithmetic per row.  GroupNorm partial statistics of the final values (ConvDesc::stats) are read back
    // from the staged tile: per (32-row block, stat_gran channels) sum and sum of squares.
    const int wq = warp - 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const uint32_t ebuf0 = epi_all + wq * p.epi_bufs * kEpiBuf;
    uint64_t* rbar = rbars + wq * kMaxEpiBufs;
    const uint32_t NB = p.epi_bufs;         // staging buffers per warp (2 without a residual, else 4..6)
    const bool tma_res = p.res_mode == 1;   // residual tile TMA-loaded INTO the staging buffer, added in place
    const bool tma_res2 = p.res_mode == 2;  // nearest-2x upsampled source: 16 source rows into a side slot
    const int sg_shift = p.stat_gran == 4 ? 2 : 1;
    const int stat_ld = (p.up2 ? p.cout_real : p.N_total) >> sg_shift;  // stat entries per 32-row block
    const int x7 = lane & 7;
    const int rg = lane >> 3, cc = lane & 7;  // read-back layout: row group, 16-byte chunk
    // chunks per tile, columns per chunk
    const int cpc = (p.epi_mode == 2 && !p.swap_ab) ? 64 : 32;
    const int n_chunks = p.swap_ab ? p.tile_px / 32 : (block_n + cpc - 1) / cpc;

    if (p.epi_mode == 0) {
      // final conv (Cout = 3 padded to N = 16), NCHW fp32 output: lanes = adjacent pixels -> coalesced already
      uint32_t it = 0;
      for (int tile = tile_first; tile < total_tiles; tile = tile_next(tile), ++it) {
        const int m_tile = (tile / p.n_tiles) * kCtas + cta_rank;
        const int n_tile = n_of(tile);
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * acc_cols;
        float bias_r[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = n_tile * block_n + j;
          bias_r[j] = (p.bias != nullptr && col < p.N_total) ? __ldg(p.bias + col) : 0.f;
        }
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x16(taddr, v);
        tmem_ld_wait();
        const int m = m_tile * kTileM + quarter * 32 + lane;
        if (p.hfold) {
          // v[s * Cout + co] = partial sum of horizontal tap s at THIS pixel; out(x) = P[x-1][s=0] + P[x][s=1] +
          // P[x+1][s=2].  Tiles are whole image rows, so both neighbours live in this tile (or are padding).
          // Exchange through shared memory, double-buffered by tile parity: one named barrier per tile.
          const int nco = p.N_total, row = quarter * 32 + lane;
          const uint32_t xb = epi_all + (it & 1) * (kTileM * 60);  // [128 pixels][15 floats]: odd stride, no conflicts
#pragma unroll
          for (int j = 0; j < 15; ++j) sts32(xb + (row * 15 + j) * 4, __uint_as_float(v[j]));
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const int x = m % p.Wout;
          const int img = m / p.HW, pix = m - img * p.HW;
          for (int co = 0; co < nco; ++co) {
            float o = lds32(xb + (row * 15 + nco + co) * 4) + (p.bias ? __ldg(p.bias + co) : 0.f);
            if (x > 0) o += lds32(xb + ((row - 1) * 15 + co) * 4);
            if (x < p.Wout - 1) o += lds32(xb + ((row + 1) * 15 + 2 * nco + co) * 4);
            p.out_nchw[(static_cast<long>(img) * nco + co) * p.HW + pix] = o;
          }
        } else if (m < p.M_total) {
          const int img = m / p.HW, pix = m - img * p.HW;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = n_tile * block_n + j;
            if (col < p.N_total)
              p.out_nchw[(static_cast<long>(img) * p.N_total + col) * p.HW + pix] = __uint_as_float(v[j]) + bias_r[j];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kCtas == 2) mbar_arrive_leader(&tempty[acc]);
          else mbar_arrive(&tempty[acc]);
        }
      }
    } else {
      // coordinates of chunk i of a tile: rows [row0, +32) x columns [col0, +cpc) of the output matrix
      auto chunk_coords = [&](int tile, int i, int& row0, int& col0) {
        const int m_tile = (tile / p.n_tiles) * kCtas + cta_rank;
        const int n_tile = n_of(tile);
        if (p.swap_ab) {
          row0 = m_tile * p.tile_px + 32 * i;
          col0 = quarter * 32;
        } else {
          row0 = m_tile * kTileM + quarter * 32;
          col0 = n_tile * block_n + cpc * i;
        }
      };
      // row of the nearest-2x upsampled residual source that output row m reads (res_mode 2)
      auto src_row = [&](int m) {
        const int img = m / p.HW, pix = m - img * p.HW;
        const int y = pix / p.Wout, x = pix - y * p.Wout;
        return (img * (p.Hout >> 1) + (y >> 1)) * (p.Wout >> 1) + (x >> 1);
      };
      // residual prefetch cursor: runs NB-2 chunks ahead of the processing cursor (in-flight residual bytes
      // per SM = 4 warps x (NB-2) x 4 KB must cover HBM latency x the residual read rate).  NB slots per warp:
      // res_mode 1: the 4 KB staging buffers themselves; res_mode 2: 2 KB slots behind two staging buffers.
      // The cursor's tile coordinates (integer divisions: ~600 clk on one lane, measured as a constant residual
      // wait per chunk when they were redone for every chunk) are derived once per tile.
      int pre_tile = tile_first, pre_i = 0;
      int pre_r0 = 0, pre_c0 = 0;  // output row (or, res_mode 2, source row) / column of chunk 0 of pre_tile
      auto set_pre = [&]() {
        if (pre_tile >= total_tiles) return;
        chunk_coords(pre_tile, 0, pre_r0, pre_c0);
        if (tma_res2) pre_r0 = src_row(pre_r0);
      };
      set_pre();
      // chunks processed (g; slot g % NB and ring-pass parity kept incrementally: NB is a runtime value) and the
      // slot of the next residual request
      uint32_t g = 0, gslot = 0, gphase = 0, pslot = 0;
      auto res_slot_addr = [&](uint32_t slot) { return tma_res ? ebuf0 + slot * kEpiBuf : ebuf0 + 2 * kEpiBuf + slot * 2048; };
      auto issue_res = [&]() {
        if (pre_tile >= total_tiles) return;
        if (lane == 0) {
          // chunk pre_i of the tile: 32 rows further down (swap-AB) or cpc columns further right
          int r0 = p.swap_ab ? pre_r0 + 32 * pre_i : pre_r0;
          const int c0 = p.swap_ab ? pre_c0 : pre_c0 + cpc * pre_i;
          if (p.res_rows) r0 %= p.res_rows;  // residual shared by the conditional / unconditional halves (tiles never straddle)
          const uint32_t slot = pslot;
          mbar_arrive_expect_tx(&rbar[slot], tma_res ? kEpiBuf : 2048);
          tma_load_2d_a(&p.tmRes, smem_u32(&rbar[slot]), res_slot_addr(slot), c0, r0);
        }
        if (++pslot == NB) pslot = 0;
        if (++pre_i == n_chunks) { pre_i = 0; pre_tile = tile_next(pre_tile); set_pre(); }
      };
      if (tma_res || tma_res2)
        for (uint32_t k = 0; k + 2 < NB; ++k) issue_res();

      uint32_t it = 0;
      long long t_acc_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int tile = tile_first; tile < total_tiles; tile = tile_next(tile), ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * acc_cols;
        int tile_row0, tile_col0;
        chunk_coords(tile, 0, tile_row0, tile_col0);
        // bias of this tile's columns -> smem, before the wait on the MMA.  The four warps write identical
        // values into the buffer of this tile's parity, once every warp has left the tile that used it last
        // (a warp may run up to two tiles ahead of another).  swap-AB: the warp's own 32 channels.
        const uint32_t sbias = sbias_all + acc * 1024 + (p.swap_ab ? quarter * 128 : 0);
        mbar_wait(&bfree[acc], acc_phase ^ 1);
        {
          const int ncol = p.swap_ab ? 32 : block_n;
          for (int c = lane; c < ncol; c += 32)
            sts32(sbias + c * 4, (p.bias != nullptr && tile_col0 + c < p.N_total)
                                     ? __ldg(p.bias + (p.up2 ? (tile_col0 + c) % p.cout_real : tile_col0 + c)) : 0.f);
          __syncwarp();
        }
        // res_mode 2: which of the 16 staged source rows this thread's output row reads (the rows a 32-row
        // chunk needs are contiguous in the source, starting at the source row of the chunk's first row)
        int s_r = 0;
        if (tma_res2) {
          s_r = src_row(min(tile_row0 + lane, p.M_total - 1)) - src_row(min(tile_row0, p.M_total - 1));
          s_r = max(0, min(s_r, 15));
        }
        const float bias_ch = p.swap_ab ? lds32(sbias + lane * 4) : 0.f;
        SGDM_T0();
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        SGDM_T(0);
        for (int i = 0; i < n_chunks; ++i, ++g) {
          const uint32_t b = tma_res2 ? (g & 1) : gslot;
          const uint32_t baddr = ebuf0 + b * kEpiBuf;
          const uint32_t b2addr = epi2_all + (wq * 2 + (g & 1)) * 2048;  // 16-bit copy tile (second output)
          const int row0 = p.swap_ab ? tile_row0 + 32 * i : tile_row0;
          const int col0 = p.swap_ab ? tile_col0 : tile_col0 + cpc * i;
          // sub-pixel mode: parity of this tile, output channel of the chunk, statistics row block (contiguous per sample)
          const int up_par = p.up2 ? col0 / p.cout_real : 0;
          const int ocol0 = p.up2 ? col0 - up_par * p.cout_real : col0;
          const long srow = p.up2 ? (static_cast<long>(row0 / p.HW) * 4 + up_par) * (p.HW >> 5) + ((row0 % p.HW) >> 5)
                                  : static_cast<long>(row0 >> 5);
          // the buffer written next (by the residual load for chunk g+NB-2, or by this chunk when there is
          // no residual) was last read by the TMA store of chunk g-2: all but the newest store must be done
          // (without a residual all NB buffers rotate as store sources: the store of chunk g-NB must be done)
          if (lane == 0) {
            if (tma_res || tma_res2 || NB == 2 || p.out2) bulk_wait_read<1>();  // (the 16-bit copy staging is 2 deep)
            else if (NB == 3) bulk_wait_read<2>();
            else bulk_wait_read<3>();
          }
          __syncwarp();
          SGDM_T(1);
          if (tma_res || tma_res2) {
            issue_res();
            mbar_wait(&rbar[gslot], gphase);
          }
          SGDM_T(2);
          // All shared-memory loads of a phase are issued back to back before their first use (the accessors
          // are volatile asm: the compiler keeps their order, so a load placed after a store would wait for it).
          float sg[4] = {0.f, 0.f, 0.f, 0.f}, qg[4] = {0.f, 0.f, 0.f, 0.f};
          // (each chunk is walked in two 16-column halves: halves the live registers of the accumulator /
          //  bias / residual values, which is what bounds the register count of the whole kernel)
          if (p.swap_ab) {
            // thread = output channel, registers = 16 consecutive pixels per half
            float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
            const uint32_t lane_off = (lane & 3) << 2;
            const uint32_t lane_chunk = lane >> 2;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint32_t v[32];
              tmem_ld_32x16(taddr + 32 * i + 16 * hf, v);
              float val[16];
              if (p.epi_mode == 1) {
                float rr[16];
                if (tma_res) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const int px = 16 * hf + j;
                    rr[j] = lds32(baddr + px * 128 + (((lane_chunk ^ (px & 7)) << 4) | lane_off));
                  }
                }
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  val[j] = __uint_as_float(v[j]) + bias_ch;
                  if (tma_res) val[j] += rr[j];
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int px = 16 * hf + j;
                  sts32(baddr + px * 128 + (((lane_chunk ^ (px & 7)) << 4) | lane_off), val[j]);
                }
                if (p.out2) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) sts16(b2addr + (16 * hf + j) * 64 + lane * 2, to_op(val[j]));
                }
              } else {
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const op_t h = to_op(__uint_as_float(v[j]) + bias_ch);
                  sts16(baddr + (16 * hf + j) * 64 + lane * 2, h);
                  val[j] = from_op(h);
                }
              }
              if (p.stats) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  s4[j & 3] += val[j];
                  q4[j & 3] += val[j] * val[j];
                }
              }
            }
            SGDM_T(4);
            if (p.stats) {
              sg[0] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
              qg[0] = (q4[0] + q4[1]) + (q4[2] + q4[3]);
              sg[0] += __shfl_xor_sync(0xffffffffu, sg[0], 1);
              qg[0] += __shfl_xor_sync(0xffffffffu, qg[0], 1);
              if (p.stat_gran == 4) {
                sg[0] += __shfl_xor_sync(0xffffffffu, sg[0], 2);
                qg[0] += __shfl_xor_sync(0xffffffffu, qg[0], 2);
              }
              const int ch = col0 + lane;
              if ((lane & (p.stat_gran - 1)) == 0 && row0 < p.M_total)
                p.stats[static_cast<long>(row0 >> 5) * stat_ld + (ch >> sg_shift)] = make_float2(sg[0], qg[0]);
            }
          } else if (p.epi_mode == 1) {
            // thread = output row, 32 consecutive channels = 8 x 16-byte chunks (XOR-swizzled), 4 per half
            const uint32_t rowaddr = baddr + lane * 128;
            const uint32_t row2 = b2addr + lane * 64;  // 16-bit copy: 64-byte rows, chunk j at j ^ ((row >> 1) & 3)
            const uint32_t x3 = (lane >> 1) & 3;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint32_t v[32];
              tmem_ld_32x16(taddr + 32 * i + 16 * hf, v);
              float4 bb[4], rr[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) bb[c] = lds128(sbias + (32 * i + 16 * hf + 4 * c) * 4);
              if (tma_res) {
#pragma unroll
                for (int c = 0; c < 4; ++c) rr[c] = lds128(rowaddr + (((4 * hf + c) ^ x7) << 4));
              } else if (tma_res2) {
                const uint32_t ra = res_slot_addr(gslot) + s_r * 128;
#pragma unroll
                for (int c = 0; c < 4; ++c) rr[c] = lds128(ra + (((4 * hf + c) ^ (s_r & 7)) << 4));
              }
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                float4 a = make_float4(__uint_as_float(v[4 * c]) + bb[c].x, __uint_as_float(v[4 * c + 1]) + bb[c].y,
                                       __uint_as_float(v[4 * c + 2]) + bb[c].z, __uint_as_float(v[4 * c + 3]) + bb[c].w);
                if (tma_res || tma_res2) { a.x += rr[c].x; a.y += rr[c].y; a.z += rr[c].z; a.w += rr[c].w; }
                sts128(rowaddr + (((4 * hf + c) ^ x7) << 4), a);
                bb[c] = a;  // kept for the 16-bit copy
              }
              if (p.out2) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const float4 lo = bb[2 * j], hi = bb[2 * j + 1];
                  sts128u(row2 + (((2 * hf + j) ^ x3) << 4),
                          make_uint4(pack_op2(lo.x, lo.y), pack_op2(lo.z, lo.w), pack_op2(hi.x, hi.y), pack_op2(hi.z, hi.w)));
                }
              }
            }
            SGDM_T(4);
            if (p.stats) {
              __syncwarp();
#pragma unroll
              for (int kh = 0; kh < 2; ++kh) {
                float4 a[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int row = 4 * (4 * kh + k) + rg;
                  a[k] = lds128(baddr + row * 128 + ((cc ^ (row & 7)) << 4));
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  sg[0] += a[k].x + a[k].y; qg[0] += a[k].x * a[k].x + a[k].y * a[k].y;
                  sg[1] += a[k].z + a[k].w; qg[1] += a[k].z * a[k].z + a[k].w * a[k].w;
                }
              }
              if (p.stat_gran == 4) { sg[0] += sg[1]; qg[0] += qg[1]; }
#pragma unroll
              for (int o = 8; o <= 16; o <<= 1) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  sg[u] += __shfl_xor_sync(0xffffffffu, sg[u], o);
                  qg[u] += __shfl_xor_sync(0xffffffffu, qg[u], o);
                }
              }
              const int col = ocol0 + 4 * cc;
              if (rg == 0 && col0 + 4 * cc < p.N_total && row0 < p.M_total) {
                float2* dst = p.stats + srow * stat_ld + (col >> sg_shift);
                if (p.stat_gran == 4) dst[0] = make_float2(sg[0], qg[0]);
                else *reinterpret_cast<float4*>(dst) = make_float4(sg[0], qg[0], sg[1], qg[1]);
              }
            }
          } else {
            // 16-bit output: thread = output row, 64 consecutive channels = 8 x 16-byte chunks of 8, 2 per quarter
            const uint32_t rowaddr = baddr + lane * 128;
#pragma unroll
            for (int qt = 0; qt < 4; ++qt) {
              uint32_t v[32];
              tmem_ld_32x16(taddr + 64 * i + 16 * qt, v);
              float4 bb[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) bb[c] = lds128(sbias + (64 * i + 16 * qt + 4 * c) * 4);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                const float4 b0 = bb[2 * c], b1 = bb[2 * c + 1];
                const uint4 h = make_uint4(
                    pack_op2(__uint_as_float(v[8 * c]) + b0.x, __uint_as_float(v[8 * c + 1]) + b0.y),
                    pack_op2(__uint_as_float(v[8 * c + 2]) + b0.z, __uint_as_float(v[8 * c + 3]) + b0.w),
                    pack_op2(__uint_as_float(v[8 * c + 4]) + b1.x, __uint_as_float(v[8 * c + 5]) + b1.y),
                    pack_op2(__uint_as_float(v[8 * c + 6]) + b1.z, __uint_as_float(v[8 * c + 7]) + b1.w));
                sts128u(rowaddr + (((2 * qt + c) ^ x7) << 4), h);
              }
            }
            SGDM_T(4);
            if (p.stats) {
              __syncwarp();
#pragma unroll
              for (int kh = 0; kh < 2; ++kh) {
                uint4 hh[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int row = 4 * (4 * kh + k) + rg;
                  hh[k] = lds128u(baddr + row * 128 + ((cc ^ (row & 7)) << 4));
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 p0 = unpack_op2(hh[k].x), p1 = unpack_op2(hh[k].y), p2 = unpack_op2(hh[k].z), p3 = unpack_op2(hh[k].w);
                  sg[0] += p0.x + p0.y; qg[0] += p0.x * p0.x + p0.y * p0.y;
                  sg[1] += p1.x + p1.y; qg[1] += p1.x * p1.x + p1.y * p1.y;
                  sg[2] += p2.x + p2.y; qg[2] += p2.x * p2.x + p2.y * p2.y;
                  sg[3] += p3.x + p3.y; qg[3] += p3.x * p3.x + p3.y * p3.y;
                }
              }
              if (p.stat_gran == 4) {
                sg[0] += sg[1]; qg[0] += qg[1];
                sg[1] = sg[2] + sg[3]; qg[1] = qg[2] + qg[3];
              }
#pragma unroll
              for (int o = 8; o <= 16; o <<= 1) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (u < 2 || p.stat_gran != 4) {
                    sg[u] += __shfl_xor_sync(0xffffffffu, sg[u], o);
                    qg[u] += __shfl_xor_sync(0xffffffffu, qg[u], o);
                  }
                }
              }
              const int col = ocol0 + 8 * cc;
              if (rg == 0 && col0 + 8 * cc < p.N_total && row0 < p.M_total) {
                float4* dst = reinterpret_cast<float4*>(p.stats + srow * stat_ld + (col >> sg_shift));
                dst[0] = make_float4(sg[0], qg[0], sg[1], qg[1]);
                if (p.stat_gran != 4) dst[1] = make_float4(sg[2], qg[2], sg[3], qg[3]);
              }
            }
          }
          SGDM_T(5);
          // generic-proxy writes -> visible to the async proxy, then one lane stores the tile
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (p.up2)  // (c, dx, x, dy, image row): pixel (2y + dy, 2x + dx) of the 2x larger output
              tma_store_5d(&p.tmOut, baddr, ocol0, up_par & 1, row0 % p.Wout, up_par >> 1, row0 / p.Wout);
            else tma_store_2d(&p.tmOut, baddr, col0, row0);
            if (p.out2) tma_store_2d(&p.tmOut2, b2addr, col0, row0);  // same bulk group: one wait covers both
            bulk_commit();
          }
          SGDM_T(6);
          t_acc_[7] += 1;
          if (++gslot == NB) { gslot = 0; gphase ^= 1; }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kCtas == 2) mbar_arrive_leader(&tempty[acc]);
          else mbar_arrive(&tempty[acc]);
          mbar_arrive(&bfree[acc]);
        }
      }
      if (lane == 0) bulk_wait_all();  // the staging smem must outlive the last stores
      if (p.timing && lane == 0)
        for (int k = 0; k < 8; ++k)
          atomicAdd(reinterpret_cast<unsigned long long*>(p.timing) + k, static_cast<unsigned long long>(t_acc_[k]));
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
                       int stride, char* err, int errlen, int kblk = 64) {
  auto fn = get_encode_fn();
  if (!fn) { snprintf(err, errlen, "cuTensorMapEncodeTiled entry point unavailable"); return 1; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kblk, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = fn(tm, SGDM_TMA_DTYPE, 4, const_cast<op_t*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, kblk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled(A: B%d H%d W%d C%d box %dx%dx%d s%d) failed: %d", B, H, W, C, bw,
             bh, bn, stride, (int)r);
    return 1;
  }
  return 0;
}

// row-major [rows, cols] matrix, box = box_rows rows x box_cols columns; swizzle = 0 | 64 | 128 (= the box row bytes)
int encode_matrix_map(CUtensorMap* tm, const void* base, bool f32, long rows, int cols, int box_cols, int swizzle,
                      char* err, int errlen, int box_rows) {
  auto fn = get_encode_fn();
  if (!fn) { snprintf(err, errlen, "cuTensorMapEncodeTiled entry point unavailable"); return 1; }
  const int es = f32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * es};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : SGDM_TMA_DTYPE, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled(matrix %ld x %d, es %d, box %d) failed: %d", rows, cols, es, box_cols, (int)r);
    return 1;
  }
  return 0;
}

int conv_prepare(const ConvDesc& d, ConvLaunch* out, char* err, int errlen) {
  memset(&out->p, 0, sizeof(out->p));
  out->desc = d;
  ConvKernelParams& p = out->p;
  auto fail = [&](const char* msg) { snprintf(err, errlen, "conv_prepare: %s", msg); return 1; };
  if (d.Cin % 64 || (d.in2 && d.C2 % 64) || (d.in2b && (d.C2b % 64 || !d.in2)))
    return fail("channel counts must be multiples of 64");
  if (!(d.ks == 1 || d.ks == 3) || !(d.stride == 1 || d.stride == 2)) return fail("unsupported ks/stride");
  // sub-pixel mode, dense geometry (up2 == 2): a plain 2x2 conv per parity over [4 Cout][4 Cin] weights
  const int ks = d.up2 == 2 ? 2 : d.ks;
  if (d.block_n != 16 && (d.block_n % 32 || d.block_n > 256 || d.block_n <= 0)) return fail("bad block_n");
  const int HW = d.Hout * d.Wout;
  if (!d.swap_ab) {
    if (d.Wout > 128 || (128 % d.Wout) != 0) return fail("Wout must divide 128");
    if (!((HW % 128) == 0 || (128 % HW) == 0)) return fail("Hout*Wout must divide or be a multiple of 128");
  } else if (d.block_n != 128) {
    return fail("swap_ab uses block_n == 128 (all output channels in one tile)");
  }
  if (d.out_nchw == nullptr && (d.Cout % 8)) return fail("Cout % 8 != 0 needs the NCHW epilogue");
  if (d.out_nchw != nullptr && d.block_n != 16) return fail("the NCHW epilogue is the block_n == 16 (Cout <= 16) path");
  if (d.out_nchw == nullptr && d.block_n == 16) return fail("block_n == 16 needs the NCHW epilogue");
  if ((d.out_f32 != nullptr) + (d.out_op != nullptr) + (d.out_nchw != nullptr) != 1) return fail("exactly one output");
  if (d.out_op && d.res) return fail("a residual needs the fp32 output");
  if (d.out_nchw && d.res) return fail("the NCHW epilogue has no residual");
  if (d.out_op && !d.swap_ab && (d.block_n % 64)) return fail("16-bit output needs block_n % 64 == 0");
  if (d.res && d.res_mode == 2 && ((d.Hout | d.Wout) & 1)) return fail("res_mode 2 needs even output size");
  if (d.swap_ab && !conv_can_swap(d)) return fail("swap_ab needs Cout == 128, no NCHW / upsampled-residual epilogue");
  if (d.stats && (d.out_nchw || (d.Cout % 8) || (d.stat_gran != 2 && d.stat_gran != 4)))
    return fail("stats need an NHWC output, Cout % 8 == 0 and stat_gran in {2, 4}");
  const int tile_px = d.swap_ab ? 256 : kTileM;
  const bool pair = conv_use_pair(d);
  const int bw = d.Wout;
  const int bh = min(d.Hout, tile_px / bw);
  const int bn = tile_px / (bw * bh);
  // halo mode: 3x3 stride 1, tiles made of whole rows of one image, row pitch = whole swizzle atoms
  bool halo = d.halo != 0 && d.up2 != 2 && d.ks == 3 && d.stride == 1 && bn == 1 && bw * bh == tile_px && (HW % tile_px) == 0 &&
              (d.Wout % 8) == 0 && bh + 2 <= 256;
  if (d.halo == 1 && !halo) return fail("halo mode needs a 3x3 stride-1 conv whose tiles are whole rows of one image");
  if (d.hfold && (!halo || !d.out_nchw || 3 * d.Cout > 16 || d.in2 || d.swap_ab || pair))
    return fail("hfold needs the NCHW head (3 * Cout <= 16) in halo geometry, one CTA per tile");
  const int b_rows = d.block_n / (pair ? 2 : 1);  // weight rows per CTA and stage slot
  p.epi_mode = d.out_nchw ? 0 : d.out_f32 ? 1 : 2;
  p.res_mode = d.res ? d.res_mode : 0;
  if (d.out_op2 && !d.out_f32) return fail("the 16-bit copy (out_op2) accompanies the fp32 output");
  p.out2 = d.out_op2 ? 1 : 0;
  const int smem_cap = kSmemLimit;
  const int budget = smem_cap - kBarBytes - kBiasBytes - (p.out2 ? kEpi2Bytes : 0);
  // staging buffers per epilogue warp: 2 without a residual; res_mode 1 adds the in-place residual ring (>= 3,
  // up to 6: in-flight residual bytes per SM must cover HBM latency); res_mode 2 uses four 2 KB slots in two more
  // (the NCHW head needs no staging; with hfold its pixel-exchange buffers, 2 x 4.5 KB, live in that region)
  const int min_bufs = p.epi_mode == 0 ? (d.hfold ? 1 : 0) : p.res_mode == 1 ? 3 : p.res_mode == 2 ? 4 : 2;
  int n_stages = 0;
  bool pack_skip = true;  // three skip-source K blocks per halo stage (needs a 48 KB activation slot)
  // K block: 64 channels, or 32 (half-size stages: 64-byte rows, SWIZZLE_64B) when that is what lets a halo
  // ring fit beside the epilogue staging — the Cout = 128 swap-AB convs with a residual, whose 96 KB halo stages
  // do not: per tile they pull 864 KB through L2 -> SM without the halo (measured at the ~11 TB/s L2 -> SM limit,
  // tensor pipe 47 %), 576 KB with it.  ConvDesc::k32: -1 policy, 0 never, 1 force.
  const bool halo_ok = halo;
  if (d.up2) {
    if ((d.up2 == 1 && !halo) || d.ks != 3 || d.stride != 1 || d.swap_ab || d.res || d.in2 || d.out_nchw || d.out_op2 || d.hfold ||
        d.k32 == 1 || d.a_stat == 1 || (d.Cout % d.block_n) || (HW % 32) || (d.Wout < 32 && (32 % d.Wout)) || (d.Wout > 32 && (d.Wout % 32)))
      return fail("up2 (sub-pixel) mode needs a 3x3 stride-1 conv, Cout % block_n == 0, no residual / skip / NCHW output "
                  "(up2 == 1: the halo geometry)");
  }
  int kblk = d.k32 == 1 ? 32 : 64;
  if (kblk == 32 && ((d.Cin % 32) || !halo_ok || pair || d.hfold)) return fail("k32 needs a halo-mode conv, one CTA per tile");
  for (int pass = 0; pass < 6; ++pass) {
    const int row_bytes = kblk * 2;
    p.kblk = kblk;
    p.wgt_tx = (d.swap_ab ? kTileM : b_rows) * row_bytes;
    p.wgt_bytes = (p.wgt_tx + 1023) / 1024 * 1024;
    p.act_tx = tile_px * row_bytes;
    p.halo = halo ? 1 : 0;
    p.tps = halo ? (d.up2 ? 2 : 3) : 1;  // sub-pixel mode: two vertical taps per parity
    p.act_tx_halo = (bh + 2) * bw * row_bytes;
    p.act_bytes = halo ? p.act_tx_halo : p.act_tx;
    p.halo_row_bytes = bw * row_bytes;
    // halo stages have three weight slots: let the fused 1x1-skip K blocks use them three at a time as well
    // (needs room for three plain activation tiles in the activation slot)
    p.tps2 = (halo && d.in2 && !d.swap_ab && pack_skip) ? 3 : 1;
    if (p.tps2 == 3 && p.act_bytes < 3 * p.act_tx) p.act_bytes = 3 * p.act_tx;
    const int stage_bytes = p.act_bytes + p.tps * p.wgt_bytes;
    const int min_stages = halo ? (kblk == 32 ? 3 : 2) : (stage_bytes > 32768 ? 3 : 4);
    n_stages = (budget - 4 * min_bufs * kEpiBuf) / stage_bytes;
    if (n_stages > kMaxStages) n_stages = kMaxStages;
    if (n_stages < min_stages && p.tps2 == 3) { pack_skip = false; continue; }     // first give up the skip packing,
    // (policy: not with a fused 1x1-skip source in swap-AB mode — each of its 32-channel K blocks would take a whole
    //  halo stage; measured +0.02 ms per such layer, against -0.055 ms for the plain ones)
    if (n_stages < min_stages && d.up2) return fail("up2 mode: the halo ring does not fit");
    if (n_stages < min_stages && halo && kblk == 64 && d.k32 != 0 && !d.hfold && !pair &&
        (d.k32 == 1 || !(d.in2 && d.swap_ab))) {
      kblk = 32;  // then halve the K block,
      pack_skip = true;
      continue;
    }
    if (n_stages < min_stages && halo && d.halo != 1 && !d.hfold) {  // then fall back to per-tap stages of 64 channels
      halo = false;
      kblk = 64;
      continue;
    }
    if (n_stages < 2) return fail("shared memory budget: fewer than 2 stages");
    p.epi_bufs = min_bufs;
    if (p.res_mode == 1) {
      // trade ring depth beyond the minimum for a deeper residual ring
      const int max_bufs = kMaxEpiBufs;  // (ring depths 4 / 6 / 8 measured in round 1: no difference)
      while (n_stages > min_stages && (budget - n_stages * stage_bytes) / (4 * kEpiBuf) < max_bufs) --n_stages;
      p.epi_bufs = (budget - n_stages * stage_bytes) / (4 * kEpiBuf);
      if (p.epi_bufs > max_bufs) p.epi_bufs = max_bufs;
    }
    break;
  }
  // Plain (non-halo) stages with TWO K blocks: the per-stage cost of the single MMA-issuing lane (barrier wait, fence,
  // commit) is paid once per 8 MMAs instead of once per 4.  Only where three such stages fit.  SGDM_CONV_KPS=1: A/B.
  const int kps_max = 3;
  p.kps = 1;
  if (!p.halo && p.kblk == 64 && d.a_stat != 1) {
    const int n_blk = ks * ks * (d.Cin / 64);
    // (three blocks per stage only as two 96..144 KB stages, like the halo ring; two blocks need three stages)
    for (int kps = min(kps_max, min(n_blk, 3)); kps >= 2; --kps) {
      const int stage_k = kps * (p.act_tx + p.wgt_bytes);
      int ns = (budget - 4 * min_bufs * kEpiBuf) / stage_k;
      if (ns > kMaxStages) ns = kMaxStages;
      if (ns < (kps == 3 ? 2 : 3)) continue;
      p.kps = kps;
      p.tps = kps;
      p.act_bytes = kps * p.act_tx;
      p.tps2 = d.in2 ? kps : 1;
      n_stages = ns;
      p.epi_bufs = min_bufs;
      if (p.res_mode == 1) p.epi_bufs = max(min_bufs, min(kMaxEpiBufs, (budget - ns * stage_k) / (4 * kEpiBuf)));
      break;
    }
  }
  // A-stationary main loop: a 1x1 GEMM with several n-tiles re-reads its activation rows once per n-tile (qkv: 6
  // times, 1.6 GB through L2 -> SM for 0.13 GB of input).  With K <= 512 the m-tile's K blocks fit in shared memory
  // (kc1 x 16 KB): they are loaded once per m-tile and the ring carries weight tiles only.  Policy: >= 3 n-tiles.
  int a_region = 0;
  p.a_stat = 0;
  {
    const int n_tiles = conv_npad(d.Cout, d.block_n) / d.block_n;
    const bool want = d.a_stat != 0 && d.ks == 1 && d.stride == 1 && !d.in2 && !d.swap_ab && !d.hfold &&
                      p.epi_mode != 0 && p.kblk == 64 && !p.halo && p.kps == 1 && d.Cin / 64 <= kMaxStages &&
                      (d.a_stat == 1 || n_tiles >= 3);
    if (want) {
      const int region = (d.Cin / 64) * p.act_tx, stage = p.wgt_bytes;
      int ns = (budget - 4 * min_bufs * kEpiBuf - region) / stage;
      if (ns > kMaxStages) ns = kMaxStages;
      if (ns >= 3) {
        p.a_stat = 1;
        a_region = region;
        p.act_bytes = 0;
        p.tps = 1;
        p.tps2 = 1;
        n_stages = ns;
        p.epi_bufs = min_bufs;
        if (p.res_mode == 1) p.epi_bufs = max(min_bufs, min(6, (budget - region - ns * stage) / (4 * kEpiBuf)));
      } else if (d.a_stat == 1) {
        return fail("A-stationary mode: the resident K blocks leave fewer than 3 ring stages");
      }
    } else if (d.a_stat == 1) {
      return fail("A-stationary mode needs a plain 1x1 GEMM with K <= 512");
    }
  }
  // Without a residual the staging buffers only rotate as TMA-store sources: two suffice, up to two more are
  // taken from shared memory the K-block ring left over (never from the ring itself: measured, reserving them
  // up front costs halo-mode stages and 3.5 ms per step).
  if (p.epi_mode != 0 && p.res_mode == 0) {
    const int left = budget - a_region - n_stages * (p.act_bytes + p.tps * p.wgt_bytes) - 4 * min_bufs * kEpiBuf;
    p.epi_bufs = min(4, min_bufs + left / (4 * kEpiBuf));
  }
  p.n_stages = n_stages;
  out->smem = smem_cap - budget + 4 * p.epi_bufs * kEpiBuf + n_stages * (p.act_bytes + p.tps * p.wgt_bytes) + a_region;
  if (encode_nhwc(&p.tmA, d.in, d.B, d.Hin, d.Win, d.Cin, bw, p.halo ? bh + 2 : bh, bn, d.stride, err, errlen, kblk)) return 1;
  if (d.in2) {
    if (encode_nhwc(&p.tmA2, d.in2, d.B, d.Hout, d.Wout, d.C2, bw, bh, bn, 1, err, errlen, kblk)) return 1;
    if (d.in2b && encode_nhwc(&p.tmA2b, d.in2b, d.B, d.Hout, d.Wout, d.C2b, bw, bh, bn, 1, err, errlen, kblk)) return 1;
  }
  const int Ktot = d.hfold ? d.ks * d.Cin : ks * ks * d.Cin + (d.in2 ? d.C2 + (d.in2b ? d.C2b : 0) : 0);
  const int npad = d.up2 ? 4 * d.Cout : conv_npad(d.Cout, d.block_n);
  {
    auto fn = get_encode_fn();
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)npad};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)kblk, (cuuint32_t)(d.swap_ab ? kTileM : b_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&p.tmB, SGDM_TMA_DTYPE, 2, const_cast<op_t*>(d.w), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, kblk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
  p.ks = ks;
  p.taps = ks * ks;
  p.kc1 = d.Cin / kblk;
  p.kc2a = d.in2 ? d.C2 / kblk : 0;
  p.kc2 = p.kc2a + (d.in2 && d.in2b ? d.C2b / kblk : 0);
  p.N_total = d.up2 ? 4 * d.Cout : d.Cout;
  p.up2 = d.up2;
  p.cout_real = d.Cout;
  p.ntpp = d.up2 ? d.Cout / d.block_n : 0;
  p.block_n = d.block_n;
  p.n_tiles = npad / d.block_n;
  p.swap_ab = d.swap_ab;
  p.hfold = d.hfold;
  p.tile_px = tile_px;
  p.m_tiles = (p.M_total + tile_px - 1) / tile_px;
  p.bias = d.bias;
  p.res = d.res;
  p.res_rows = 0;
  if (d.res && d.res_mode == 1 && d.res_batch > 0 && d.res_batch < d.B) {
    if (d.B % d.res_batch || (static_cast<long>(d.res_batch) * HW) % tile_px) return fail("res_batch must divide B and cover whole tiles");
    p.res_rows = d.res_batch * HW;
  }
  p.out_nchw = d.out_nchw;
  p.stats = d.stats;
  p.stat_gran = d.stat_gran;
  p.timing = d.timing;
  // epilogue: output / residual tile maps
  if (d.up2) {
    // the 2x larger output [B, 2H, 2W, C] seen as (c, dx, x, dy, image row): one parity plane per store
    auto fn = get_encode_fn();
    const bool f32 = p.epi_mode == 1;
    const int es = f32 ? 4 : 2, box_c = f32 ? 32 : 64;
    const cuuint64_t C = (cuuint64_t)d.Cout, W = (cuuint64_t)d.Wout;
    cuuint64_t dims[5] = {C, 2, W, 2, (cuuint64_t)d.B * d.Hout};
    cuuint64_t strides[4] = {C * es, 2 * C * es, 2 * W * C * es, 4 * W * C * es};
    const int bx = d.Wout < 32 ? d.Wout : 32;
    cuuint32_t box[5] = {(cuuint32_t)box_c, 1, (cuuint32_t)bx, 1, (cuuint32_t)(32 / bx)};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    void* base = f32 ? static_cast<void*>(d.out_f32) : static_cast<void*>(d.out_op);
    CUresult r = fn(&p.tmOut, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : SGDM_TMA_DTYPE, 5, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(up2 output) failed: %d", (int)r); return 1; }
  } else if (p.epi_mode == 1) {
    if (encode_matrix_map(&p.tmOut, d.out_f32, true, p.M_total, d.Cout, 32, 128, err, errlen)) return 1;
    if (p.res_mode == 1 &&
        encode_matrix_map(&p.tmRes, d.res, true, p.res_rows ? p.res_rows : p.M_total, d.Cout, 32, 128, err, errlen))
      return 1;
    if (p.res_mode == 2 &&
        encode_matrix_map(&p.tmRes, d.res, true, static_cast<long>(d.B) * (d.Hout / 2) * (d.Wout / 2), d.Cout, 32, 128, err,
                      errlen, 16))
      return 1;
    // second output: the same values rounded to the 16-bit operand type (32 channels = 64-byte rows)
    if (d.out_op2 && encode_matrix_map(&p.tmOut2, d.out_op2, false, p.M_total, d.Cout, 32, d.swap_ab ? 0 : 64, err, errlen)) return 1;
  } else if (p.epi_mode == 2) {
    // swap-AB: a warp owns 32 channels = 64-byte rows (dense); normal: 64 channels = 128-byte swizzled rows
    if (encode_matrix_map(&p.tmOut, d.out_op, false, p.M_total, d.Cout, d.swap_ab ? 32 : 64, d.swap_ab ? 0 : 128, err, errlen)) return 1;
  }
  // shared memory: as many K-block stages as fit beside the epilogue staging
  out->pair = pair ? 1 : 0;
  // (A-stationary: the work items of a CTA / pair are whole m-tiles)
  if (d.up2 && (p.kblk != 64 || (d.up2 == 1) != (p.halo != 0) || p.a_stat)) return fail("up2 mode: unexpected geometry");
  if (pair) {
    const int total = (p.m_tiles + 1) / 2 * (p.a_stat ? 1 : p.n_tiles);
    out->grid = 2 * (total < kNumSMs / 2 ? total : kNumSMs / 2);
  } else {
    const int total = p.m_tiles * (p.a_stat ? 1 : p.n_tiles);
    out->grid = total < kNumSMs ? total : kNumSMs;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    attr_set = true;
  }
  return 0;
}

int conv_up2_applicable(int H, int W, int Cin, int Cout, int block_n) {
  const int HW = H * W;
  if ((Cin % 64) || block_n < 64 || (Cout % block_n)) return 0;
  if ((HW % 32) || (W < 32 ? (32 % W) != 0 : (W % 32) != 0)) return 0;         // 32-pixel store chunks, statistics blocks
  if (HW < kTileM) return (kTileM % HW) == 0 && W >= 4 ? 2 : 0;                 // tiles of whole images: the dense geometry
  if (W < 8 || (W % 8) || W > kTileM || (kTileM % W) || (HW % kTileM)) return 0;  // whole rows of one image per tile
  const int bh = kTileM / W;
  // two halo stages of (bh + 2) rows + two weight slots each (one CTA per tile: the larger case) beside 2 staging buffers
  const int stage = (bh + 2) * W * 128 + 2 * block_n * 128;
  return 2 * stage + 4 * 2 * kEpiBuf + kBarBytes + kBiasBytes <= kSmemLimit ? 1 : 0;
}

int conv_launch(const ConvLaunch& l, cudaStream_t stream) {
  const cudaError_t e = l.pair ? launch_pdl(conv_gemm_kernel<2>, dim3(l.grid), dim3(kConvThreads), l.smem, stream, 2, l.p)
                               : launch_pdl(conv_gemm_kernel<1>, dim3(l.grid), dim3(kConvThreads), l.smem, stream, 1, l.p);
  return e == cudaSuccess ? 0 : 1;
}

// --------------------------------------------------------------------- CUDA-core checker
__global__ void conv_naive_kernel(ConvDesc d, int npad) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const long M = static_cast<long>(d.B) * d.Hout * d.Wout;
  if (d.up2) {  // sub-pixel mode: one thread per output element of the 2x larger tensor, straight from the definition
    const long total = M * 4 * d.Cout;
    if (idx >= total) return;
    const int co = idx % d.Cout;
    const long px = idx / d.Cout;                      // output pixel in NHWC order of [B, 2H, 2W]
    const int W2 = 2 * d.Wout, H2 = 2 * d.Hout;
    const int X = px % W2, Y = (px / W2) % H2, b = px / (static_cast<long>(W2) * H2);
    const int dy = Y & 1, dx = X & 1, yy = Y >> 1, xx = X >> 1;
    const int taps_k = d.up2 == 2 ? 4 : 9;  // dense geometry: [4 Cout][4 Cin], tap (a, b) at (a * 2 + b) * Cin
    const op_t* wr = d.w + static_cast<long>((2 * dy + dx) * d.Cout + co) * (taps_k * d.Cin);
    float acc = 0.f;
    for (int R = dy; R <= dy + 1; ++R)
      for (int S = dx; S <= dx + 1; ++S) {
        const int iy = yy + R - 1, ix = xx + S - 1;
        if (iy < 0 || iy >= d.Hin || ix < 0 || ix >= d.Win) continue;
        const op_t* a = d.in + ((static_cast<long>(b) * d.Hin + iy) * d.Win + ix) * d.Cin;
        const op_t* w = wr + (d.up2 == 2 ? (R - dy) * 2 + (S - dx) : R * 3 + S) * d.Cin;
        for (int c = 0; c < d.Cin; ++c) acc += from_op(a[c]) * from_op(w[c]);
      }
    if (d.bias) acc += d.bias[co];
    if (d.out_f32) d.out_f32[px * d.Cout + co] = acc;
    if (d.out_op) d.out_op[px * d.Cout + co] = to_op(acc);
    return;
  }
  if (idx >= M * d.Cout) return;
  const int col = idx % d.Cout;
  const long m = idx / d.Cout;
  const int HW = d.Hout * d.Wout;
  const int img = m / HW, pix = m % HW, y = pix / d.Wout, x = pix % d.Wout;
  const int Ktot = d.ks * d.ks * d.Cin + (d.in2 ? d.C2 + (d.in2b ? d.C2b : 0) : 0);
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
    if (d.in2b) {
      const op_t* a2 = d.in2b + m * d.C2b;
      for (int c = 0; c < d.C2b; ++c) acc += from_op(a2[c]) * from_op(w[d.C2 + c]);
    }
  }
  if (d.bias) acc += d.bias[col];
  if (d.res && d.res_mode == 1) acc += d.res[m * d.Cout + col];
  if (d.res && d.res_mode == 2)
    acc += d.res[((static_cast<long>(img) * (d.Hout / 2) + y / 2) * (d.Wout / 2) + x / 2) * d.Cout + col];
  if (d.out_f32) d.out_f32[m * d.Cout + col] = acc;
  if (d.out_op2) d.out_op2[m * d.Cout + col] = to_op(acc);
  if (d.out_op) d.out_op[m * d.Cout + col] = to_op(acc);
  if (d.out_nchw) d.out_nchw[(static_cast<long>(img) * d.Cout + col) * HW + pix] = acc;
}

int conv_launch_naive(const ConvDesc& d, cudaStream_t stream) {
  const long total = static_cast<long>(d.B) * d.Hout * d.Wout * d.Cout * (d.up2 ? 4 : 1);
  const int threads = 256;
  conv_naive_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, stream>>>(
      d, conv_npad(d.Cout, d.block_n));
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace sgdm
