// Implicit-GEMM convolution / GEMM on tcgen05 (sm_100a): host-side description.
//
// Replaces the reference's nn.Conv2d 3x3 (s1/s2, pad 1), nn.Conv2d 1x1, nn.Conv1d(k=1)
// and nn.Linear call sites (dynamic/diffusionmodules/openaimodel.py:245-287,349-357,
// openaimodel_ca.py:101-181, crossattetion_lr.py:70-79) — SURVEY.md §2.3 rows K3/K4.
//
//   D[M = B*Hout*Wout, N = Cout] = sum over K-blocks of A[M, 64] * W[N, 64]^T
//   K-blocks = (tap r,s) x (64-channel chunk of the NHWC source)  [+ 1x1 chunks of a
//   second "skip" source accumulated into the same TMEM tile: fused skip_connection]
//
// A tiles are fetched with 4-D tiled TMA boxes (64ch, bw, bh, bn) at coordinates shifted
// by the tap offset; out-of-bounds pixels are zero-filled by TMA, which implements the
// conv padding for free (no im2col buffer).  Stride-2 convs use TMA elementStrides.
#pragma once
#include "common.cuh"

namespace sgdm {

struct ConvDesc {
  // main source: NHWC op_t [B, Hin, Win, Cin], Cin % 64 == 0
  const op_t* in = nullptr;
  int B = 0, Hin = 0, Win = 0, Cin = 0;
  // optional 1x1 skip source at OUTPUT resolution: NHWC op_t [B, Hout, Wout, C2], C2 % 64 == 0
  const op_t* in2 = nullptr;
  int C2 = 0;
  // ... which may itself be a channel concat [in2 (C2) | in2b (C2b)] of two tensors
  const op_t* in2b = nullptr;
  int C2b = 0;
  // packed weights: op_t [Npad][Ktot], Ktot = ks*ks*Cin + C2, K order = (tap, cin) then skip cin
  const op_t* w = nullptr;
  int ks = 3, stride = 1, pad = 1;
  int Hout = 0, Wout = 0, Cout = 0;
  // epilogue
  const float* bias = nullptr;  // [Cout]
  const float* res = nullptr;   // fp32 NHWC residual, see res_mode
  int res_mode = 0;             // 0 none | 1 same shape | 2 nearest-2x upsampled source [B,Hout/2,Wout/2,Cout]
  int res_batch = 0;            // res_mode 1: the residual tensor has only this many samples; output sample n adds residual
                                // sample n % res_batch (0 = B).  The shared CFG prefix of a guided plan (see GnDesc::src_mod0).
  float* out_f32 = nullptr;     // NHWC fp32 [B,Hout,Wout,Cout]
  op_t* out_op = nullptr;       // NHWC op_t
  float* out_nchw = nullptr;    // NCHW fp32 [B,Cout,Hout,Wout] (final conv, Cout = 3)
  op_t* out_op2 = nullptr;      // with out_f32: the same values rounded to op_t as a second NHWC tensor (GroupNorm input copy)
  int block_n = 128;            // 16 or a multiple of 32, <= 256; Npad = roundup(Cout, block_n)
  // Cout == 128 only: compute D^T = W * X^T, i.e. the 128 output channels are the MMA M dimension and a
  // tile of 256 PIXELS is the MMA N dimension.  An M128xN128 SS-MMA needs 128 B/clk of shared-memory
  // reads (the port limit, ~50 % tensor rate); M128xN256 needs 96 B/clk and runs at full rate.
  int swap_ab = 0;
  // Optional GroupNorm partial statistics of the FINAL output values (after bias / residual, before the
  // 16-bit rounding): stats[(row / 32) * (Cout / stat_gran) + channel / stat_gran] = {sum, sum of squares}
  // over that 32-row block and `stat_gran` (2 or 4) adjacent channels.  Rows are output pixels in NHWC
  // order, so with Hout*Wout % 32 == 0 every block lies inside one sample and any GroupNorm whose groups
  // are unions of such channel granules (also across a channel concat) is finalised from these sums
  // without re-reading the tensor (gn_finalize_kernel).  Buffer: ceil(M/32) * Cout/stat_gran float2.
  float2* stats = nullptr;
  int stat_gran = 4;
  // CTA-pair mode (tcgen05 cta_group::2, M = 256 over two CTAs): -1 = policy (conv_use_pair), 0 = off, 1 = on
  int pair = -1;
  // halo mode (one activation load per horizontal tap, shared by the three vertical taps): -1 policy, 0 off, 1 on
  int halo = -1;
  // Horizontal-tap folding for the output head (3x3, stride 1, 3 * Cout <= 16, NCHW epilogue, halo geometry):
  // the GEMM computes, per pixel p and horizontal tap s, the partial sums P[p, s * Cout + co] = sum over
  // (vertical tap r, channel c) of in(y + r - 1, x, c) * w[co, c, r, s] — one staged activation tile (rows
  // y0-1 .. y0+bh, NO horizontal shift) serves all nine taps — and the epilogue adds the three shifted
  // partials: out(y, x) = P[(y, x-1), s=0] + P[(y, x), s=1] + P[(y, x+1), s=2].  Activation bytes through
  // L2 -> SM drop 3x against halo mode, the MMA count 3x.  `w` is then packed [16][3 * Cin]:
  // row s * Cout + co, column r * Cin + c (pack_conv_weight_hfold_launch).
  int hfold = 0;
  // Sub-pixel execution of "nearest-2x upsample, then 3x3 conv" (the up ResBlocks of unet_fast, openaimodel.py:253-258,
  // 301-306; Upsample of unetca_fast, openaimodel_ca.py:121-131).  Output pixel (2y+dy, 2x+dx) of that conv only ever sees
  // the 2x2 low-resolution neighbourhood {y+dy-1, y+dy} x {x+dx-1, x+dx}, each reached by 1, 2 or 4 of the nine taps:
  // four parity convs with 2x2 summed-tap kernels on the LOW-resolution tensor give the same sums with 16 instead of 36
  // MACs per low-resolution pixel and channel pair (2.25x less tensor work, no upsampled operand tensor).
  //   `in` is the low-resolution tensor [B, Hin, Win, Cin]; Hout = Hin, Wout = Win are the GEMM's pixel grid; the OUTPUT
  //   tensor is [B, 2 Hout, 2 Wout, Cout]; the GEMM's N is 4 Cout (parity-major: n = (2 dy + dx) Cout + co), `w` is packed by
  //   pack_conv_weight_up2_launch as [4 Cout][9 Cin] with tap (R, S) of parity (dy, dx) = sum of W[r, s] over
  //   r in V(dy, R), s in V(dx, S), V(0,0) = {0}, V(0,1) = {1,2}, V(1,1) = {0,1}, V(1,2) = {2} (other taps unused);
  //   statistics: sample n's row blocks are [(4 n + parity) HW/32 + block], i.e. contiguous per sample as gn_finalize expects.
  // Needs the halo geometry (3x3, stride 1, tiles of whole rows), Cout % block_n == 0, no residual / skip source.
  int up2 = 0;
  int a_stat = -1;              // A-stationary main loop for 1x1 GEMMs with >= 3 n-tiles and K <= 512: -1 policy, 0 off, 1 force
  int k32 = -1;                 // K block of 32 channels (SWIZZLE_64B halo stages): -1 policy, 0 never, 1 force
  long long* timing = nullptr;  // optional device array of 16 cycle counters (kernel_conv.cu, tuning only)
};

struct alignas(64) ConvKernelParams {
  CUtensorMap tmA;
  CUtensorMap tmA2;
  CUtensorMap tmA2b;
  CUtensorMap tmB;
  CUtensorMap tmOut;  // output matrix [M_total, Cout]: 32-row x 128-byte tiles (TMA store)
  CUtensorMap tmRes;  // residual matrix (res_mode 1): same tiling (TMA load)
  CUtensorMap tmOut2; // optional 16-bit copy of the fp32 output: 32-row x 64-byte tiles
  int M_total, HW, Wout, Hout;
  int stride, pad, ks, taps;
  int kc1, kc2, kc2a;  // 64-channel chunks: main source; skip source(s) in total; of the first skip tensor
  int N_total, block_n, n_tiles, m_tiles;
  int swap_ab, tile_px;  // tile_px: pixels per tile (128, or 256 when swap_ab)
  // K-block ring: stage = [activation slot: act_bytes][tps weight slots: wgt_bytes each]; *_tx = bytes TMA delivers
  int n_stages, act_bytes, act_tx, act_tx_halo, wgt_bytes, wgt_tx;
  int halo, tps, halo_row_bytes;  // halo mode: 3 vertical taps per stage read one staged tile at row offsets
  int hfold;                      // horizontal taps folded into the N dimension (output head)
  int up2, cout_real, ntpp;       // sub-pixel mode (ConvDesc::up2): real output channels, n-tiles per parity
  int kps;                        // K blocks per plain (non-halo) main stage: 1 or 2
  int a_stat;                     // A-stationary 1x1 GEMM: the m-tile's activation K blocks stay resident across its n-tiles
  int kblk;                       // channels per K block: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B)
  int tps2;                       // K blocks of the fused 1x1-skip source per stage (3 in halo mode, else 1)
  int epi_mode, epi_bufs;       // 0 NCHW direct | 1 fp32 NHWC | 2 16-bit NHWC; staging buffers per epilogue warp
  int out2;                     // epi_mode 1: also store the 16-bit copy
  const float* bias;
  const float* res;
  int res_mode;
  int res_rows;  // rows of the residual matrix when it is shorter than the output (ConvDesc::res_batch), else 0
  float* out_nchw;
  float2* stats;
  int stat_gran;
  long long* timing;
};

struct ConvLaunch {
  ConvKernelParams p;
  int grid = 0;
  int pair = 0;
  int smem = 0;
  ConvDesc desc;  // kept for the naive checker path
};

// Tiled tensor map of a row-major [rows, cols] matrix (fp32 or op_t), box = box_rows x box_cols,
// swizzle = 0 | 64 | 128 (= the box row bytes).  Returns 0 on success.
int encode_matrix_map(CUtensorMap* tm, const void* base, bool f32, long rows, int cols, int box_cols, int swizzle,
                      char* err, int errlen, int box_rows = 32);

// Builds the TMA descriptors and launch geometry.  Returns 0 on success; on failure
// writes a message into err (size errlen).
int conv_prepare(const ConvDesc& d, ConvLaunch* out, char* err, int errlen);
int conv_launch(const ConvLaunch& l, cudaStream_t stream);
// CUDA-core checker with identical semantics (test / bring-up only; never on the product path).
int conv_launch_naive(const ConvDesc& d, cudaStream_t stream);

inline int conv_npad(int cout, int block_n) { return (cout + block_n - 1) / block_n * block_n; }
// policy used by the engine and by block_n = 0 in sgdm_k_conv
inline bool conv_can_swap(const ConvDesc& d) {
  const int HW = d.Hout * d.Wout;
  return d.Cout == 128 && d.out_nchw == nullptr && !(d.res && d.res_mode == 2) && d.Wout <= 256 &&
         (256 % d.Wout) == 0 && ((HW % 256) == 0 || (256 % HW) == 0);
}
// (Measured, round 2: NOT swapping the one layer that is pure epilogue — the first conv as a K = 64 GEMM — in the hope
//  that the thread-per-row epilogue is cheaper: 0.484 -> 0.534 ms.  Swap-AB stays the policy for every Cout = 128 layer.)
inline bool conv_should_swap(const ConvDesc& d) { return conv_can_swap(d); }

// Can a "nearest-2x upsample, then 3x3 conv" whose LOW-resolution input is [*, H, W, Cin] run in the sub-pixel mode
// (ConvDesc::up2)?  Independent of the batch size (halo tiles never span images).
// Returns 0 (no), 1 (halo geometry: >= 128 pixels per image, [4 Cout][9 Cin] weights) or 2 (dense geometry: tiles of whole
// small images, a plain 2x2 conv per parity over [4 Cout][4 Cin] weights) = the value for ConvDesc::up2.
int conv_up2_applicable(int H, int W, int Cin, int Cout, int block_n);

// CTA pairs need two m-tiles to share a weight tile of >= 64 output channels (each CTA stages block_n/2 rows).
inline bool conv_pair_ok(const ConvDesc& d) { return !d.swap_ab && d.block_n >= 64 && (d.block_n % 32) == 0; }
inline bool conv_use_pair(const ConvDesc& d) {
  if (d.pair == 0 || !conv_pair_ok(d)) return false;
  if (d.pair == 1) return true;
  const long m_tiles = (static_cast<long>(d.B) * d.Hout * d.Wout + 127) / 128;
  return m_tiles >= 2;
}

}  // namespace sgdm
