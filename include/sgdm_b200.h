/* sgdm_b200.h — C ABI of libsgdm_b200.so (sm_100a).
 *
 * The drop-in boundary of the guided reverse-diffusion hot path.  Everything crossing it
 * is a plain pointer, size or scalar: no torch types, no C++ types, no exceptions.  All
 * device pointers are CUDA device memory of the current device; `stream` is a
 * cudaStream_t passed as void* (the host passes torch.cuda.current_stream().cuda_stream).
 * Every function returns 0 on success and non-zero on failure; sgdm_last_error() then
 * returns a message.  No function synchronises the device or the stream.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   sgdm_create / sgdm_load_param   UNetModel.__init__ + load_state_dict
 *                                   (dynamic/diffusionmodules/openaimodel.py:496-835,
 *                                    openaimodel_ca.py:479-836)
 *   sgdm_forward                    UNetModel.forward (openaimodel.py:904-956,
 *                                    openaimodel_ca.py:917-1033)
 *   sgdm_forward_guided + sgdm_mix  UNetModel.forward_with_cond_scale + get_guided_score
 *                                   (openaimodel.py:853-902, openaimodel_ca.py:871-915)
 *   sgdm_ddim_step                  DDIMSampler.p_sample_ddim / p_sample_plms
 *                                   (diffusion/sampler/ddim_plms_sampler.py:345-391,482-525)
 *   sgdm_ddpm_step                  Schedule_DDPM.p_mean_variance + p_sample
 *                                   (diffusion/sampler/ddpm_sampler.py:154-192)
 *   sgdm_to_uint8                   clip_unnormalize_to_zero_to_255 (diffusion_utils/util.py:99-100)
 *   sgdm_k_*                        single-kernel entry points used by the unit parity tests
 */
#ifndef SGDM_B200_H
#define SGDM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgdm_engine* sgdm_handle;

#define SGDM_KIND_UNET_FAST 0   /* dynamic.diffusionmodules.openaimodel.UNetModel    */
#define SGDM_KIND_UNETCA_FAST 1 /* dynamic.diffusionmodules.openaimodel_ca.UNetModel */
#define SGDM_SCALE_IMAGEN 0
#define SGDM_SCALE_CFG 1

typedef struct sgdm_config {
  int32_t kind;
  int32_t image_size, in_channels, out_channels, model_channels, num_res_blocks;
  int32_t n_channel_mult;
  int32_t channel_mult[8];
  int32_t n_attention_resolutions;
  int32_t attention_resolutions[8];
  int32_t num_heads;
  int32_t resblock_updown;
  int32_t cond_dim;
  int32_t layout_dim;     /* 0 | clusterlayout: 1 | stegoclusterlayout: 27 */
  int32_t context_dim;    /* unetca_fast: 32 */
  int32_t cond_token_num; /* unetca_fast: 1 (README runs) | 0 (layout-only) | N > 1: cond is [B, N, cond_dim] tokens */
  /* 0 (default): 16-bit GEMM operands, fp32 accumulation — the throughput path.
   * 1 ("fp16 x3"): every conv / GEMM on split operands (a_hi w_hi + a_hi w_lo + a_lo w_hi, ~22-bit operands, fp32
   *    accumulation) at ~3x the tensor work: for deterministic samplers (DDIM eta=0, PLMS) on ill-conditioned
   *    networks, where 16-bit operand rounding is amplified along the trajectory. */
  int32_t precision;
  /* cond_token_num > 1 only: the vector fed to cond_mlp is token 0 (1, the reference config's value) or the mean over
   * the tokens (0) (openaimodel_ca.py:1000-1006) */
  int32_t use_cls_token_as_pooled;
} sgdm_config;

const char* sgdm_last_error(void);
const char* sgdm_version(void);
/* "f16" (default build) or "bf16": the 16-bit GEMM operand type; accumulation is fp32 */
const char* sgdm_operand_dtype(void);

/* Host-only: builds the layer list and the parameter inventory. No CUDA call is made. */
int sgdm_create(const sgdm_config* cfg, sgdm_handle* out);
int sgdm_destroy(sgdm_handle h);

/* Parameter inventory == the reference module's state_dict (same names, same shapes). */
int sgdm_param_count(sgdm_handle h);
const char* sgdm_param_name(sgdm_handle h, int i);
int sgdm_param_shape(sgdm_handle h, int i, int64_t* dims /* >= 4 */, int* ndim);

/* Copies / packs one fp32 parameter (device pointer, contiguous) into the engine.
 * The engine never keeps `data`; call again after the parameter changes (EMA swap). */
int sgdm_load_param(sgdm_handle h, const char* name, const float* data, const int64_t* shape, int ndim,
                    void* stream);
/* Number of parameters not loaded yet (0 = ready). */
int sgdm_params_missing(sgdm_handle h);
/* freqs[i] = exp(-ln(10000) * i / half), i < model_channels/2, computed by the host with the
 * reference's own fp32 expression (dynamic/diffusionmodules/util.py:160-163). Host pointer. */
int sgdm_set_timestep_freqs(sgdm_handle h, const float* host_freqs, int n);

/* eps = UNet(x, t, cond, layout) with a per-sample drop mask (1 = use the null embeddings).
 *   x [B, C, H, W] fp32 NCHW | t [B] int64 | cond [B, cond_dim] fp32 or NULL |
 *   layout [B, L, H, W] fp32 or NULL | drop [B] uint8 or NULL (= keep all) | eps_out [B, C, H, W] fp32 */
int sgdm_forward(sgdm_handle h, void* stream, const float* x, const int64_t* t, const float* cond,
                 const float* layout, const uint8_t* drop, int B, float* eps_out);

/* Conditional and unconditional passes as ONE batched launch sequence (rows [0,B) keep
 * the condition, rows [B,2B) use the null embeddings).  Writes the engine-owned result
 * pointers (valid until the next forward on this handle): eps_c, eps_u [B, C, H, W] fp32. */
int sgdm_forward_guided(sgdm_handle h, void* stream, const float* x, const int64_t* t, const float* cond,
                        const float* layout, int B, const float** eps_c, const float** eps_u);

/* eps = (1-w) eps_u + w eps_c (imagen) | (1+w) eps_c - w eps_u (cfg); w scalar (a double, like the
 * Python number the reference multiplies with: 1-w is formed in double, then rounded to fp32), or
 * per sample when w_per_sample != NULL ([B] fp32 device; 1-w is then an fp32 op). */
int sgdm_mix(void* stream, const float* eps_c, const float* eps_u, double w, const float* w_per_sample,
             int scale_type, float* eps_out, int B, int64_t per_sample);

/* One DDIM / PLMS update, optionally fused with the guidance mix (eps_u may be NULL: eps = eps_c).
 * coef = {sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma_t, temperature}. */
int sgdm_ddim_step(void* stream, const float* eps_c, const float* eps_u, double w, const float* w_per_sample,
                   int scale_type, const float* coef6, int clip_denoised, const float* x, const float* noise,
                   float* x_out, float* x0_out /* may be NULL */, float* eps_out /* may be NULL */, int B,
                   int64_t per_sample);
/* One ancestral DDPM update. coef = {sqrt_recip_ac[t], sqrt_recipm1_ac[t], post_mean_coef1[t],
 * post_mean_coef2[t], (t!=0)*exp(0.5*post_logvar[t]), temperature}. */
int sgdm_ddpm_step(void* stream, const float* eps_c, const float* eps_u, double w, const float* w_per_sample,
                   int scale_type, const float* coef6, int clip_denoised, const float* x, const float* noise,
                   float* x_out, float* x0_out /* may be NULL */, int B, int64_t per_sample);
/* The same updates with the optional sampling_kwargs extras (either pointer may be NULL):
 *   dyn_s     [B] per-sample dynamic threshold from sgdm_dyn_threshold (dtp < 1): pred_x0 <- clamp(x0,-s,s)/s
 *             replaces the clamp to [-1,1] (clip_x0_minus_one_to_one, diffusion_utils/util.py:70-82)
 *   noise_mul [B, per_sample] F.dropout factor {0, 1/(1-p)} multiplied onto the scaled noise (noise_dropout > 0,
 *             ddpm_sampler.py:184-185, ddim_plms_sampler.py:388-389); drawn by the host like the noise itself */
int sgdm_ddim_step_ex(void* stream, const float* eps_c, const float* eps_u, double w, const float* w_per_sample,
                      int scale_type, const float* coef6, int clip_denoised, const float* x, const float* noise,
                      float* x_out, float* x0_out, float* eps_out, int B, int64_t per_sample, const float* dyn_s,
                      const float* noise_mul);
int sgdm_ddpm_step_ex(void* stream, const float* eps_c, const float* eps_u, double w, const float* w_per_sample,
                      int scale_type, const float* coef6, int clip_denoised, const float* x, const float* noise,
                      float* x_out, float* x0_out, int B, int64_t per_sample, const float* dyn_s,
                      const float* noise_mul);
/* Dynamic thresholding, phase 1: s_out[b] = max(quantile(|pred_x0[b]|, dtp), 1) (torch.quantile, 'linear') for the
 * update described by the same eps / coef6 (kind 0 = DDPM coefficients, 1 = DDIM); scratch_x0 [B, per_sample]. */
int sgdm_dyn_threshold(void* stream, int kind, const float* eps_c, const float* eps_u, double w, const float* w_per_sample,
                       int scale_type, const float* coef6, const float* x, double dtp, float* scratch_x0, float* s_out, int B,
                       int64_t per_sample);
/* PLMS multistep eps combination (ddim_plms_sampler.py:432-459): out = (sum_k coefs[k]*terms[k]) / div,
 * n_terms <= 4; `terms` and `coefs` are HOST arrays (of device pointers / floats). */
int sgdm_lincomb(void* stream, int n_terms, const float* const* terms, const float* coefs, float div, float* out,
                 int64_t n);
/* out = scale * (sum_k coefs[k]*terms[k]) — PNDM's multistep form `(1 / 24) * (55 e1 - 59 e2 + 37 e3 - 9 e4)`
 * (diffusion/sampler/pndm_sampler.py:118-127) */
int sgdm_lincomb_scaled(void* stream, int n_terms, const float* const* terms, const float* coefs, float scale, float* out,
                        int64_t n);
/* PNDM transfer x_next = x + d * (A * x - B * et) with separately rounded fp32 operations
 * (PNDMScheduler.transfer, diffusion/sampler/pndm_sampler.py:128-141); d, A, B: host-computed fp32 scalars */
int sgdm_pndm_transfer(void* stream, const float* x, const float* et, float d, float A, float B, float* out, int64_t n);
int sgdm_to_uint8(void* stream, const float* x, uint8_t* out, int64_t n);

/* Per-launch timing of one forward (bench.py roofline): with profiling on, the next forward
 * brackets every launch of its plan with CUDA events on `stream` and synchronises at the end.
 * Each record: kernel family, measured ms, algorithmic FLOPs and algorithmic HBM bytes. */
/* Guided plans compute the part of the network the conditional and the unconditional half have in common (the first
 * conv and the first ResBlock's GroupNorm + conv: identical x, no embedding yet) once for B rows instead of 2B
 * (models without a layout input).  Same bits.  Default on; 0 switches it off for plans used afterwards (A/B, tests). */
int sgdm_set_share_prefix(sgdm_handle h, int on);
/* CUDA-graph replay of a plan's static launch list (captured once per plan on first use; the prologue that reads
 * the caller's inputs stays outside): -1 = policy (every plan; falls back to stream replay if the capture fails),
 * 0 = off, 1 = on (a failed capture is an error).  Same kernels, same order, same values as the stream replay. */
int sgdm_set_graph_mode(sgdm_handle h, int mode);
/* dev_out[t] = order-independent 64-bit hash of the bit pattern of the n fp32 tensors whose device pointers /
 * element counts are in the DEVICE arrays dev_ptrs / dev_numel.  One launch: lets the host detect parameter writes
 * that bypass its bookkeeping (`.data` copies of the reference's EMA swap, dynamic/ema.py:46-53). */
int sgdm_fingerprint(void* stream, const void* const* dev_ptrs, const int64_t* dev_numel, int n, uint64_t* dev_out);
int sgdm_set_profiling(sgdm_handle h, int on);
int sgdm_profile_count(sgdm_handle h);
/* launch i of the last profiled replay: kernel family, duration, ALGORITHMIC FLOPs (2 x MAC of the reference computation
 * the launch stands for) and HBM bytes; sgdm_profile_executed_flops: what the tensor pipe executes for it (4/9 for a
 * sub-pixel up-conv, half for a launch on the shared rows of a guided plan, 3x in the fp16x3 mode) */
int sgdm_profile_get(sgdm_handle h, int i, const char** kind, double* ms, double* flops, double* bytes);
int sgdm_profile_executed_flops(sgdm_handle h, int i, double* flops);

/* Count of kernels launched by this library since load (claim for bench.py `gpu_launches`). */
int64_t sgdm_launch_count(void);
/* Mode overrides for the SINGLE-KERNEL entry points below (sgdm_k_conv*, sgdm_k_attention) made afterwards by the
 * calling thread — unit tests force every geometry of the conv / attention kernels through them.  Thread-local; an
 * engine's plans never read them (they follow the kernels' own policy).
 * CTA-pair (tcgen05 cta_group::2) conv mode: -1 = library policy (default), 0 = never, 1 = whenever the shape allows */
int sgdm_debug_set_conv_pair(int mode);
/* tcgen05 self-attention kernel (T = 256, head dim 64): -1 = whenever applicable (default), 0 = mma.sync kernel */
int sgdm_debug_set_attn_tc(int mode);
/* halo mode of 3x3 stride-1 convs (one staged activation tile shared by the three vertical taps), same values */
int sgdm_debug_set_conv_halo(int mode);
/* K blocks of 32 channels (64-byte rows, SWIZZLE_64B) for halo-mode convs: -1 = only where 64-channel halo stages do
 * not fit beside the epilogue staging (default), 0 = never, 1 = every halo-mode conv (tests) */
int sgdm_debug_set_conv_k32(int mode);
/* A-stationary main loop of 1x1 GEMMs (the m-tile's K blocks stay in shared memory across its n-tiles): -1 = policy
 * (K <= 512 and >= 3 n-tiles), 0 = never, 1 = whenever K <= 512 (tests) */
int sgdm_debug_set_conv_astat(int mode);
/* tuning aid: single-kernel conv calls made afterwards add per-role stall cycle counts to this device array
 * of 16 int64 (NULL = off); slot meaning in csrc/kernel_conv.cu */
int sgdm_debug_set_conv_timing(void* device_counters16);

/* ---- single-kernel entry points (unit parity tests). 16-bit tensors are `op` = fp16 (or bf16). ---- */
/* conv / GEMM: in [B,Hin,Win,Cin] op NHWC; in2 optional [B,Hout,Wout,C2]; w packed [Npad][ks*ks*Cin + C2] op */
int sgdm_k_conv(void* stream, const void* in, int B, int Hin, int Win, int Cin, const void* in2, int C2,
                const void* w, int ks, int stride, int Hout, int Wout, int Cout, const float* bias,
                const float* res, int res_mode, float* out_f32, void* out_op, float* out_nchw, int block_n,
                int naive);
/* same, with the optional extras of the engine's fused layers:
 *  - stats: GroupNorm partial statistics of the final output values,
 *    stats[(row/32) * (Cout/stat_gran) + c/stat_gran] = {sum, sum of squares} (float2) over 32 output rows
 *    (NHWC pixels) x stat_gran (2 or 4) channels; buffer of ceil(B*Hout*Wout/32) * Cout/stat_gran float2
 *  - out_op2 (with out_f32): a second NHWC tensor holding the same values rounded to the 16-bit operand type
 *  - in2b / C2b: the 1x1 skip source is the channel concat [in2 (C2) | in2b (C2b)] of two tensors */
int sgdm_k_conv_stats(void* stream, const void* in, int B, int Hin, int Win, int Cin, const void* in2, int C2,
                      const void* w, int ks, int stride, int Hout, int Wout, int Cout, const float* bias,
                      const float* res, int res_mode, float* out_f32, void* out_op, float* out_nchw, int block_n,
                      int naive, float* stats, int stat_gran, void* out_op2, const void* in2b, int C2b);
/* the output head (3x3, stride 1, pad 1, 3 * Cout <= 16, NCHW fp32 result) with the horizontal taps folded into the
 * GEMM's N dimension (openaimodel.py:830-835: the conv of self.out): w is the torch weight [Cout,Cin,3,3] fp32,
 * w_scratch 16 * 3 * Cin op elements for its packed form; needs tiles of whole image rows (W | 128, H*W % 128 == 0) */
/* "nearest-2x upsample, then 3x3 conv" executed as four 2x2 parity convs on the low-resolution tensor (ConvDesc::up2):
 * in [B, H, W, Cin] 16-bit NHWC, w fp32 [Cout, Cin, 3, 3] (packed into w_scratch: 4 * Cout * 9 * Cin 16-bit), output
 * [B, 2H, 2W, Cout] fp32 or 16-bit, optional GroupNorm partial statistics of the output (sample n: row blocks
 * [n * 4HW/32, (n+1) * 4HW/32)) */
int sgdm_k_conv_up2(void* stream, const void* in, int B, int H, int W, int Cin, const float* w, void* w_scratch,
                    const float* bias, float* out_f32, void* out_op, int Cout, float* stats, int stat_gran, int naive);
int sgdm_k_conv_head_hfold(void* stream, const void* in, int B, int H, int W, int Cin, const float* w, void* w_scratch,
                           const float* bias, float* out_nchw, int Cout);
/* packs a torch conv weight [Cout,Cin,ks,ks] fp32 into dst[co][k_off + tap*cin_pad + ci] (row length ktot) */
int sgdm_k_pack_weight(void* stream, const float* w, void* dst, int Cout, int Cin, int ks, int cin_pad,
                       int ktot, int k_off);
/* src0 / src1: fp32 NHWC, or both op NHWC when src0_is_op (a 16-bit concat needs sgdm_k_groupnorm_fused) */
int sgdm_k_groupnorm(void* stream, const void* src0, int src0_is_op, const void* src1, int B, int H, int W, int C0,
                     int C1, const float* gamma, const float* beta, const float* film, int64_t film_stride, int silu,
                     int resample, void* out_op, void* raw_out_op, float* pool_out);
/* same, with the statistics taken from conv-epilogue partial sums (sgdm_k_conv_stats) of the producer(s)
 * of src0 / src1 instead of a pass over the tensor; requires H*W % 32 == 0. */
int sgdm_k_groupnorm_fused(void* stream, const void* src0, int src0_is_op, const void* src1, int B, int H, int W,
                           int C0, int C1, const float* gamma, const float* beta, const float* film,
                           int64_t film_stride, int silu, int resample, const float* stats0, const float* stats1,
                           int stat_gran, void* out_op, void* raw_out_op, float* pool_out);
int sgdm_k_layernorm(void* stream, const float* x, const float* gamma, const float* beta, const float* res,
                     void* out_op, float* out_f32, int64_t rows, int C);
/* out_f32 = res + LN(x) gamma + beta, plus the GroupNorm partial statistics of the output: stats[(row / 32) *
 * (C / stat_gran) + c / stat_gran] = {sum, sum of squares} over 32 rows x stat_gran (2 | 4) channels; rows % 32 == 0 */
int sgdm_k_layernorm_stats(void* stream, const float* x, const float* gamma, const float* beta, const float* res,
                           float* out_f32, float* stats, int stat_gran, int64_t rows, int C);
/* split-precision operand layout of sgdm_config.precision = 1: out_op [rows, 3C] = [hi | hi | lo] of LN(x) gamma + beta */
int sgdm_k_layernorm_split3(void* stream, const float* x, const float* gamma, const float* beta, void* out_op,
                            int64_t rows, int C);
int sgdm_k_attention(void* stream, const void* q, int64_t q_row_stride, int q_head_stride, const void* k,
                     int64_t k_row_stride, int k_head_stride, const void* v, int64_t v_row_stride,
                     int v_head_stride, const void* k_extra, const void* v_extra, int n_extra, void* out,
                     int64_t o_row_stride, int B, int T, int heads, int D, float scale);
/* the same through `splits` K slices (partial: splits * M * N floats of scratch; slices summed in ascending order, so
 * the result is deterministic); splits < 0 = the engine's policy (skinny M x N, K >= 1024) */
int sgdm_k_linear_f32_splitk(void* stream, const float* in, int64_t in_stride, const float* W, const float* bias,
                             float* out, int64_t out_stride, int M, int N, int K, int silu_out, int accumulate,
                             float* partial, int splits);
int sgdm_k_linear_f32(void* stream, const float* in, int64_t in_stride, const float* W, const float* bias,
                      float* out, int64_t out_stride, int M, int N, int K, int silu_out, int accumulate);
int sgdm_k_cast(void* stream, const float* src, void* dst_op, int B, int H, int W, int C, int up2);
/* s_out[b] = max(quantile(|v[b, 0:n]|, q), 1), torch.quantile 'linear' (diffusion_utils/util.py:74-77) */
int sgdm_k_quantile_abs(void* stream, const float* v, int B, int64_t n, double q, float* s_out);

#ifdef __cplusplus
}
#endif
#endif /* SGDM_B200_H */
