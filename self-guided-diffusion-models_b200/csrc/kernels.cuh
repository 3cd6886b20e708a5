// Launch wrappers of the non-GEMM kernels (HBM-bound elementwise / reduction work).
// All activations are NHWC; the residual stream is fp32, GEMM operands are op_t.
#pragma once
#include "common.cuh"

namespace sgdm {

// ---- K1/K2: GroupNorm(32) [+FiLM] [+SiLU] [+2x2 avg-pool | nearest 2x] ------------------
// Reference: GroupNorm32 (diffusionmodules/util.py:199-216) + SiLU + ResBlock FiLM
// `out_norm(h) * (1 + scale) + shift` (openaimodel.py:312-316), avg_pool2d / nearest
// upsample of up/down ResBlocks (:253-258,301-306), skip concat th.cat([h, hs.pop()], 1) (:950).
struct GnDesc {
  const void* src0 = nullptr;   // NHWC [B, H, W, C0]: fp32, or op_t when src0_is_op
  int src0_is_op = 0;           // applies to BOTH sources; a 16-bit concat (C1 > 0) needs producer statistics (stats0/1)
  const void* src1 = nullptr;   // optional second concat source [B, H, W, C1]
  int B = 0, H = 0, W = 0, C0 = 0, C1 = 0;
  const float* gamma = nullptr;  // [C0+C1]
  const float* beta = nullptr;
  const float* film = nullptr;   // optional fp32 [B, film_stride]: scale at [c], shift at [C + c]
  long film_stride = 0;
  int silu = 1;
  int resample = 0;              // 0 none | 1 avg-pool 2x2 | 2 nearest 2x
  double* partial = nullptr;     // scratch [B][chunks][32][2]
  int chunks = 1;
  op_t* out = nullptr;           // NHWC op_t [B, H', W', C]
  op_t* raw_out = nullptr;       // optional op_t copy of the un-normalised concat input (for the 1x1 skip conv)
  float* pool_out = nullptr;     // optional fp32 avg-pooled raw input (residual of a down ResBlock)
  // Statistics emitted by the producing conv's epilogue (ConvDesc::stats, conv.cuh).  When stats0 is set
  // (and stats1 whenever C1 > 0) the standalone statistics pass over the tensor is skipped: a tiny
  // finalise kernel reduces the partial sums into `final` = {mean, rstd} per (sample, group).
  const float2* stats0 = nullptr;
  const float2* stats1 = nullptr;
  int stat_gran = 4;
  float2* final = nullptr;       // scratch [B][32]
  // Source batch modulo: sample n of the OUTPUT reads sample n % src_mod of source 0 / 1 (0 = none).  The guided plan
  // computes everything in front of the first FiLM once for the B rows that the conditional and the unconditional half
  // share (identical x, identical weights); the first consumers that differ between the halves read it through this.
  int src_mod0 = 0, src_mod1 = 0;
  // Split-precision operand output (engine precision 1): `out` / `raw_out` rows have 3 (C0 + C1) channels
  // [hi | hi | lo], hi = op(v), lo = op(v - hi) (kernels_misc.cu: store8_split3).  fp32 sources only.
  int split3 = 0;
};
int gn_chunks_for(int B, int HW, int C);
int gn_launch(const GnDesc& d, cudaStream_t s);        // stats (or finalise) + apply
int gn_finalize_launch(const GnDesc& d, cudaStream_t s);  // fused-statistics path: partial sums -> {mean, rstd}
int gn_stats_launch(const GnDesc& d, cudaStream_t s);  // pass 1 only
int gn_apply_launch(const GnDesc& d, cudaStream_t s);  // pass 2 only

// ---- LayerNorm over channels (Attention_LR, crossattetion_lr.py:36-43) -------------------
// mode 0: out_op = LN(x)*gamma+beta ; mode 1: out_f32 = res + LN(x)*gamma+beta
int layernorm_launch(const float* x, const float* gamma, const float* beta, const float* res, op_t* out_op,
                     float* out_f32, long rows, int C, cudaStream_t s, int split3 = 0);
// out_f32 = res + LN(x) gamma + beta, plus the GroupNorm partial statistics of the output in ConvDesc::stats format
// ({sum, sum sq} per 32-row block x stat_gran channels); rows % 32 == 0
int layernorm_res_stats_launch(const float* x, const float* gamma, const float* beta, const float* res, float* out_f32,
                               float2* stats, int stat_gran, long rows, int C, cudaStream_t s);

// ---- casts --------------------------------------------------------------------------------
// fp32 NHWC -> op_t NHWC, optionally nearest-2x upsampled (Upsample, openaimodel_ca.py:121-131)
// (split3: rows of 3C channels [hi | hi | lo], the split-precision operand layout of engine precision 1)
int cast_launch(const float* src, op_t* dst, int B, int H, int W, int C, int up2, cudaStream_t s, int split3 = 0);
// dst[i] = op(silu(src[i]))   (ResBlock.emb_layers[0], openaimodel.py:262-263); split3: src rows of C channels
int silu_cast_launch(const float* src, op_t* dst, long n, cudaStream_t s, int C = 0, int split3 = 0);
// op_t [rows, C] -> [rows, 3C] = [v | v | 0] (a 16-bit-only tensor as the activation side of a split-precision GEMM)
int expand3_launch(const op_t* src, op_t* dst, long rows, int C, cudaStream_t s);

// ---- fp32 small linear: out[m, n] (+)= act(bias[n] + sum_k in[m,k] W[n,k]) -----------------
int linear_f32_splits(int M, int N, int K);  // split-K policy (1 = none); partial: splits * M * N floats of scratch
int linear_f32_launch(const float* in, long in_stride, const float* W, const float* bias, float* out,
                      long out_stride, int M, int N, int K, int silu_out, int accumulate, cudaStream_t s,
                      float* partial = nullptr, int splits = 1);

// ---- prologue: CFG batch assembly (C1/C3/C4 rows of SURVEY §8a) ------------------------------
struct PrepDesc {
  const float* x = nullptr;          // fp32 NCHW [B, Cimg, H, W]
  const long long* t = nullptr;      // int64 [B]
  const float* cond = nullptr;       // fp32 [B, cond_dim] (may be null when cond_dim == 0)
  const float* layout = nullptr;     // fp32 [B, L, H, W]   (null when L == 0)
  const unsigned char* drop = nullptr;  // [Bp] 1 = replace cond/layout by the null embeddings
  const float* null_cond = nullptr;  // [cond_dim]
  const float* null_layout = nullptr;   // [H*W]
  const float* freqs = nullptr;      // [mc/2] host-computed exp(-ln(1e4) i / half) (util.py:160-163)
  int B = 0, Bp = 0;                 // Bp = B or 2B; row r reads sample r % B
  int Bx = 0;                        // rows of x_in to produce (0 = Bp): the shared rows of a guided plan
  int Cimg = 3, H = 0, W = 0, L = 0, cond_dim = 0, mc = 0;
  op_t* x_in = nullptr;              // NHWC op_t [Bp, H, W, 64]: [x_hi(Cimg) | x_lo(Cimg) | layout(L) | 0]
  // im2col: the 64 channels of pixel (y, x) hold the 3x3 neighbourhood instead, channel tap * (2 Cimg + L) + j =
  // entry j of the row above at pixel (y + r - 1, x + s - 1) (zero outside the image), tap = 3 r + s: the first conv
  // becomes a 1x1 GEMM with ONE 64-wide K block (9 (2 Cimg + L) <= 64; pack_first_conv_im2col_launch)
  int im2col = 0;
  // split-precision input (engine precision 1): x_in has xc channels [hi(x, layout) | hi(x, layout) | lo(x, layout) | 0]
  int split3 = 0, xc = 64;
  float* t_emb = nullptr;            // [Bp, mc]  [cos | sin]
  float* cond_masked = nullptr;      // [Bp, cond_dim]
};
int prep_launch(const PrepDesc& d, cudaStream_t s);

// context K/V for Attention_LR: norm_cond LayerNorm over [time tokens | cond tokens]
// (openaimodel_ca.py:973,1017) then per-site to_context = LayerNorm + Linear(ctx -> 2*dh)
// (crossattetion_lr.py:75,103-106), null_kv appended as key 16 (:95-97).
constexpr int kMaxCtxSites = 12;
struct CtxSite {                       // per Attention_LR site
  const float* ln_w = nullptr;         // to_context.0 [ctx]
  const float* ln_b = nullptr;
  const float* lin_w = nullptr;        // to_context.1 [2*dh, ctx]
  const float* lin_b = nullptr;        // [2*dh]
  const float* null_kv = nullptr;      // [2, dh]
  op_t* k_out = nullptr;               // [Bp, n_tok + 1, dh]
  op_t* v_out = nullptr;
};
struct CtxDesc {
  const float* time_tokens = nullptr;  // [Bp, 8*ctx]
  const float* cond_tokens = nullptr;  // [Bp, (n_tok - 8)*ctx]
  const float* norm_w = nullptr;       // norm_cond [ctx]
  const float* norm_b = nullptr;
  int Bp = 0, ctx = 32, dh = 64;
  int n_tok = 16;                      // context tokens: 8 time + the condition tokens (8 for cond_token_num 1, N for
                                       // cond_token_num N > 1, none for 0); k_out / v_out have n_tok + 1 rows per sample
  int n_sites = 0;                     // all sites in one launch (grid.y)
  CtxSite site[kMaxCtxSites];
};
int context_kv_launch(const CtxDesc& d, cudaStream_t s);
constexpr int kMaxCtxTok = 8 + 120;    // shared memory / one-thread-per-token LayerNorm of context_kv_kernel
// out[b][j] = mean_t in[b][t][j] (cond_token_num > 1 with use_cls_token_as_pooled = False, openaimodel_ca.py:1004-1006)
int token_mean_launch(const float* in, float* out, int B, int n_tok, int dim, cudaStream_t s);

// ---- K10: guidance mix + sampler updates (fp32, bit-faithful op order) ----------------------
// eps = (1-w) eps_u + w eps_c  ('imagen') | (1+w) eps_c - w eps_u ('cfg')   (openaimodel.py:853-859)
struct MixDesc {
  const float* eps_c = nullptr;
  const float* eps_u = nullptr;  // null: eps = eps_c (single pass)
  float w = 0.f;   // fp32(w)
  float ow = 1.f;  // host-computed fp32(1 - w) ('imagen') or fp32(1 + w) ('cfg'), from the DOUBLE w like torch
  const float* w_per_sample = nullptr;  // optional [B] (tensor cond_scale [B,1,1,1])
  int scale_type = 0;                    // 0 imagen | 1 cfg
};
int mix_launch(const MixDesc& m, float* eps_out, int B, long per_sample, cudaStream_t s);

// Host-computed fp32 scalars, each produced with the reference's own dtype chain:
//   sqrt_one_minus_at = ddim_sqrt_one_minus_alphas[i]; sqrt_at = fp32 sqrt(ddim_alphas[i]);
//   sqrt_a_prev = fp32 sqrt(fp32(ddim_alphas_prev[i])); dir_coef = fp32 sqrt(1 - a_prev - sigma^2)
struct DdimCoef { float sqrt_one_minus_at, sqrt_at, sqrt_a_prev, dir_coef, sigma_t, temperature; int clip; };
// Optional extras of both update kernels (all null = the plain update):
struct StepExtras {
  float* x0_raw = nullptr;           // write ONLY the unclipped pred_x0 here (phase 1 of dynamic thresholding)
  const float* dyn_s = nullptr;      // [B] dynamic threshold s: x0 <- clamp(x0, -s, s) / s instead of the [-1, 1] clamp
  const float* noise_mul = nullptr;  // per-element F.dropout factor {0, 1/(1-p)} on the scaled noise
};
// s[b] = max(quantile(|x0[b]|, q), 1): torch.quantile 'linear' (clip_x0_minus_one_to_one, diffusion_utils/util.py:70-82)
int quantile_abs_launch(const float* x0, int B, long n, float q, float* s_out, cudaStream_t s);
// p_sample_ddim arithmetic (ddim_plms_sampler.py:360-391) fused with the mix.
int ddim_step_launch(const MixDesc& m, const DdimCoef& c, const StepExtras& ex, const float* x, const float* noise,
                     float* x_out, float* x0_out, float* eps_out, int B, long per_sample, cudaStream_t s);

// nonzero_sigma = (t != 0) * exp(0.5 * posterior_log_variance_clipped[t]) (fp32, host-computed)
struct DdpmCoef { float sqrt_recip, sqrt_recipm1, coef1, coef2, nonzero_sigma, temperature; int clip; };
// p_mean_variance + p_sample arithmetic (ddpm_sampler.py:154-192) fused with the mix.
int ddpm_step_launch(const MixDesc& m, const DdpmCoef& c, const StepExtras& ex, const float* x, const float* noise,
                     float* x_out, float* x0_out, int B, long per_sample, cudaStream_t s);

// ((x+1)*127.5).clamp(0,255).to(uint8)  (diffusion_utils/util.py:99-100)
int to_uint8_launch(const float* x, unsigned char* out, long n, cudaStream_t s);

// out = (c0*a0 + c1*a1 + ...)/div, left to right (PLMS, ddim_plms_sampler.py:432-459)
int lincomb_launch(const float* const* a, const float* c, int n_terms, float div, float* out, long n, cudaStream_t s,
                   int use_scale = 0, float pre_scale = 1.f);
// x_next = x + d * (A * x - B * et)   (PNDM transfer, pndm_sampler.py:128-141)
int pndm_transfer_launch(const float* x, const float* et, float d, float A, float B, float* out, long n, cudaStream_t s);

// ---- weight packing --------------------------------------------------------------------------
// torch conv weight fp32 [Cout, Cin, ks, ks] -> op_t dst[co][k_off + (r*ks+s)*cin_pad + ci] (row length ktot);
// channels ci >= Cin (padding) are left untouched (buffers are zero-initialised).
// first conv as a 1x1 GEMM over PrepDesc::im2col input: dst[co][tap * ce + j], ce = 2 Cimg + L, from w [Cout, Cimg + L, 3, 3]
int pack_first_conv_im2col_launch(const float* w, op_t* dst, int Cout, int Cimg, int L, cudaStream_t s);
int pack_conv_weight_hfold_launch(const float* w, op_t* dst, int Cout, int Cin, int cin_pad, cudaStream_t s);
// cin_part > 0: split-precision packing, three parts of cin_part channels [w_hi | w_lo | w_hi] per tap (3 cin_part <= cin_pad)
int pack_conv_weight_launch(const float* w, op_t* dst, int Cout, int Cin, int ks, int cin_pad, int ktot, int k_off,
                            const int* ci_map, cudaStream_t s, int cin_part = 0);
// sub-pixel packing of a 3x3 conv that follows a nearest-2x upsample (ConvDesc::up2): dst [4 Cout][9 cin_pad]
int pack_conv_weight_up2_launch(const float* w, op_t* dst, int Cout, int Cin, int cin_pad, cudaStream_t s, int dense = 0);
int add_bias_launch(const float* a, const float* b, float* out, int n, cudaStream_t s);
// out[t] = 64-bit hash of the bits of fp32 tensor t (device pointer / element-count tables on the device)
int fingerprint_launch(const void* const* ptrs, const long long* numel, int n, unsigned long long* out, cudaStream_t s);

}  // namespace sgdm
